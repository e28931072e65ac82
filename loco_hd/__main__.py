"""``python -m loco_hd``: LoCoHD scores of chosen anchor pairs between two PDB structures.

Same command line as the reference's experimental tool (``loco_hd/__main__.py:50-133`` upstream):

    python -m loco_hd -s1 a.pdb -s2 b.pdb -pts primitive_typings/all_atom_with_centroid.config.json -apf anchors.txt
                      [-mn 0] [-nt N] [-udc 10.0] [-tpra '{"accept_same": false}']
                      [-wfa '{"function_name": "uniform", "parameters": [3.0, 10.0]}']

The anchor pairing file lists ``chain/number-RESNAME/ATOM[,ATOM...]:chain/number-RESNAME/ATOM[,...]`` entries
separated by semicolons: each side names one primitive by its chain, its residue and the set of atoms it was built
from.  The structures are read by the small PDB reader of this package (the reference uses Bio.PDB), the scoring runs
on the GPU.  One line ``LoCoHD(<entry>) = <score>`` is printed per anchor pair.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path
from typing import Dict, FrozenSet, List, Sequence, Tuple

PrimitiveKey = Tuple[str, str, FrozenSet[str]]   # chain id, "number-RESNAME", names of the contributing atoms


def read_anchor_entries(path: Path) -> List[str]:
    text = Path(path).read_text().replace("\n", "")
    return [entry.strip() for entry in text.split(";") if entry.strip()]


def entry_to_keys(entry: str) -> Tuple[PrimitiveKey, PrimitiveKey]:
    try:
        left, right = entry.split(":")
        keys = []
        for side in (left, right):
            chain, residue, atoms = side.split("/")
            keys.append((chain, residue, frozenset(name.strip() for name in atoms.split(","))))
    except ValueError as exc:
        raise ValueError(f"anchor pair '{entry}' is not of the form A/123-TYR/CG,CZ:B/45-ALA/CB") from exc
    return keys[0], keys[1]


def index_by_key(templates: Sequence) -> Dict[PrimitiveKey, int]:
    """Primitive templates -> {(chain, "number-RESNAME", atom set): position in the list}."""
    table: Dict[PrimitiveKey, int] = {}
    for position, template in enumerate(templates):
        source = template.atom_source
        residue = source.source_residue
        table[(residue[2], f"{residue[3][1]}-{source.source_residue_name}", frozenset(source.source_atom))] = position
    return table


def tag_of(template) -> str:
    """The tag callers give a primitive: "chain/number-RESNAME" (one tag per residue)."""
    residue = template.atom_source.source_residue
    return f"{residue[2]}/{residue[3][1]}-{template.atom_source.source_residue_name}"


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="python -m loco_hd", description=__doc__.split("\n\n")[0])
    parser.add_argument("-s1", "--structure1", required=True, help="first PDB file")
    parser.add_argument("-s2", "--structure2", required=True, help="second PDB file")
    parser.add_argument("-pts", "--primitive_typing_scheme", required=True, help="primitive typing scheme (JSON)")
    parser.add_argument("-apf", "--anchor_pairing_file", required=True, type=Path,
                        help="text file with semicolon-separated anchor pairs, e.g. A/123-TYR/CG,CZ:B/45-ALA/CB")
    parser.add_argument("-mn", "--model_number", type=int, default=0, help="model of the PDB files to use (default 0)")
    parser.add_argument("-nt", "--number_of_threads", type=int, default=None,
                        help="accepted for compatibility; the GPU grid replaces the thread pool")
    parser.add_argument("-udc", "--upper_distance_cutoff", type=float, default=10.0, help="environment radius (default 10)")
    parser.add_argument("-tpra", "--tag_pairing_rule_args", type=json.loads, default={"accept_same": False},
                        help="JSON for TagPairingRule (default '{\"accept_same\": false}')")
    parser.add_argument("-wfa", "--weight_function_args", type=json.loads,
                        default={"function_name": "uniform", "parameters": [3.0, 10.0]},
                        help="JSON for WeightFunction (default uniform [3, 10])")
    return parser


def prepare(args):
    """Everything up to the scoring call: (LoCoHD arguments, primitives of both structures, anchor index pairs,
    anchor entries).  Needs no GPU."""
    from loco_hd import PrimitiveAssigner, PrimitiveAtom
    from loco_hd_b200.pdb_io import load_model

    entries = read_anchor_entries(args.anchor_pairing_file)
    assigner = PrimitiveAssigner(Path(args.primitive_typing_scheme))
    sides = []
    for path in (args.structure1, args.structure2):
        templates = assigner.assign_primitive_structure(load_model(path, args.model_number))
        atoms = [PrimitiveAtom(t.primitive_type, tag_of(t), [float(v) for v in t.coordinates]) for t in templates]
        sides.append((index_by_key(templates), atoms))
    pairs = []
    for entry in entries:
        key_a, key_b = entry_to_keys(entry)
        for key, (table, _), which in ((key_a, sides[0], "first"), (key_b, sides[1], "second")):
            if key not in table:
                raise KeyError(f"anchor '{key[0]}/{key[1]}/{','.join(sorted(key[2]))}' of '{entry}' names no primitive of the "
                               f"{which} structure")
        pairs.append((sides[0][0][key_a], sides[1][0][key_b]))
    rule = dict(args.tag_pairing_rule_args)
    if "tag_pairs" in rule:   # JSON has no sets / tuples
        rule["tag_pairs"] = {tuple(pair) for pair in rule["tag_pairs"]}
    return assigner.all_primitive_types, rule, sides[0][1], sides[1][1], pairs, entries


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from loco_hd import LoCoHD, TagPairingRule, WeightFunction

    types, rule, atoms_a, atoms_b, pairs, entries = prepare(args)
    lchd = LoCoHD(types, WeightFunction(**args.weight_function_args), TagPairingRule(rule), args.number_of_threads)
    scores = lchd.from_primitives(atoms_a, atoms_b, pairs, args.upper_distance_cutoff)
    for entry, score in zip(entries, scores):
        print(f"LoCoHD({entry}) = {score}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
