"""``python -m loco_hd`` (the reference's experimental tool, loco_hd/__main__.py:50-204 upstream) and the PDB reader
behind it: record parsing, Bio.PDB-style residue ids, anchor identifiers -> primitive indices (CPU), and the printed
scores against the oracle (GPU)."""
import json

import numpy as np
import pytest

from loco_hd_b200.atom_converter_utils import TYPING_DIR, PrimitiveAssigner
from loco_hd_b200.pdb_io import load_model, parse_pdb
from test_primitive_assigner import make_structure

SCHEME = TYPING_DIR / "coarse_grained_with_centroid.config.json"


def pdb_text(structure, chain_of=lambda k: "AB"[k >= 4], shift=(0.0, 0.0, 0.0), model=None, water=True):
    lines, serial = [], 1
    if model is not None:
        lines.append(f"MODEL     {model:>4}")
    for k, residue in enumerate(structure.get_residues()):
        for atom in residue.get_atoms():
            x, y, z = (float(c) + s for c, s in zip(atom.coord, shift))
            lines.append(f"ATOM  {serial:>5} {atom.name:<4} {residue.resname:>3} {chain_of(k)}{residue.full_id[3][1]:>4}    "
                         f"{x:8.3f}{y:8.3f}{z:8.3f}  1.00 20.00          {atom.name[0]:>2}")
            serial += 1
    if water:
        lines.append(f"HETATM{serial:>5}  O   HOH A 201    {1.0:8.3f}{2.0:8.3f}{3.0:8.3f}  1.00 30.00           O")
        lines.append(f"ATOM  {serial + 1:>5}  CA AALA A 202    {4.0:8.3f}{2.0:8.3f}{3.0:8.3f}  0.60 30.00           C")
        lines.append(f"ATOM  {serial + 2:>5}  CA BALA A 202    {4.5:8.3f}{2.0:8.3f}{3.0:8.3f}  0.40 30.00           C")
    if model is not None:
        lines.append("ENDMDL")
    return "\n".join(lines) + "\n"


def test_pdb_reader_records_and_ids(tmp_path):
    structure = make_structure(4)
    text = pdb_text(structure)
    (model,) = parse_pdb(text)
    residues = list(model.get_residues())
    originals = list(structure.get_residues())
    assert len(residues) == len(originals) + 2
    for got, want in zip(residues, originals):
        assert got.resname == want.resname and got.full_id[3] == (" ", want.full_id[3][1], " ")
        assert [a.name for a in got.get_atoms()] == [a.name for a in want.get_atoms()]
        for a, b in zip(got.get_atoms(), want.get_atoms()):
            assert a.coord.dtype == np.float32 and np.abs(a.coord - b.coord).max() <= 5.1e-4   # %8.3f columns
    assert residues[0].full_id[:3] == ("s", 0, "A") and residues[5].full_id[2] == "B"
    water, alt = residues[-2], residues[-1]
    assert water.full_id[3] == ("W", 201, " ") and water.resname == "HOH"
    assert len(alt) == 1 and next(alt.get_atoms()).coord[0] == np.float32(4.0)      # first alternate location kept
    assert model.atom_coordinates().shape == (sum(len(r) for r in residues), 3)
    # several models, selection by number, file input
    path = tmp_path / "two_models.pdb"
    path.write_text(pdb_text(structure, model=1, water=False) + pdb_text(structure, shift=(1.0, 0.0, 0.0), model=2, water=False))
    m0, m1 = load_model(path, 0), load_model(path, 1)
    assert m1.index == 1 and np.allclose(m1.atom_coordinates()[:, 0] - m0.atom_coordinates()[:, 0], 1.0, atol=2e-3)
    with pytest.raises(IndexError):
        load_model(path, 2)
    with pytest.raises(ValueError):
        parse_pdb("REMARK nothing here\n")
    with pytest.raises(ValueError):
        parse_pdb("ATOM      1  CA  ALA A   1      xx.xxx   1.000   1.000\n")


def _write_case(tmp_path):
    a, b = make_structure(5), make_structure(5)
    rng = np.random.default_rng(2)
    for k, residue in enumerate(b.get_residues()):      # the second structure: same topology, moved atoms
        for atom in residue.get_atoms():
            atom.coord = (atom.coord + rng.normal(0, 0.7, 3)).astype(np.float32)
    # pull the residues together so that environments are not empty at 10 A
    for structure in (a, b):
        for k, residue in enumerate(structure.get_residues()):
            centre = np.mean([atom.coord for atom in residue.get_atoms()], axis=0)
            target = np.array([4.0 * (k % 3), 4.0 * (k // 3), 0.0], dtype=np.float32)
            for atom in residue.get_atoms():
                atom.coord = (atom.coord - centre + target).astype(np.float32)
    p1, p2 = tmp_path / "a.pdb", tmp_path / "b.pdb"
    p1.write_text(pdb_text(a, water=False))
    p2.write_text(pdb_text(b, water=False))
    assigner = PrimitiveAssigner(SCHEME)
    templates = assigner.assign_primitive_structure(load_model(p1))
    entries = []
    for t in templates:
        if t.primitive_type == "Cent":
            res = t.atom_source.source_residue
            ident = f"{res[2]}/{res[3][1]}-{t.atom_source.source_residue_name}/{','.join(t.atom_source.source_atom)}"
            entries.append(f"{ident}:{ident}")
    apf = tmp_path / "anchors.txt"
    apf.write_text(";\n".join(entries) + ";\n")
    return p1, p2, apf, entries


def test_cli_prepare_maps_anchor_identifiers(tmp_path):
    import loco_hd.__main__ as cli

    p1, p2, apf, entries = _write_case(tmp_path)
    args = cli.build_parser().parse_args(["-s1", str(p1), "-s2", str(p2), "-pts", str(SCHEME), "-apf", str(apf)])
    assert args.upper_distance_cutoff == 10.0 and args.tag_pairing_rule_args == {"accept_same": False}
    assert args.weight_function_args == {"function_name": "uniform", "parameters": [3.0, 10.0]} and args.model_number == 0
    types, rule, atoms_a, atoms_b, pairs, got_entries = cli.prepare(args)
    assert got_entries == entries and len(pairs) == len(entries) == 8
    assert all(atoms_a[i].primitive_type == "Cent" and atoms_b[j].primitive_type == "Cent" and atoms_a[i].tag == atoms_b[j].tag
               for i, j in pairs)
    assert atoms_a[pairs[0][0]].tag == "A/1-GLY" and "Cent" in types and rule == {"accept_same": False}
    bad = tmp_path / "bad.txt"
    bad.write_text("A/1-GLY/XX:A/1-GLY/N,CA,C,O")
    args.anchor_pairing_file = bad
    with pytest.raises(KeyError):
        cli.prepare(args)
    bad.write_text("A/1-GLY")
    with pytest.raises(ValueError):
        cli.prepare(args)
    args2 = cli.build_parser().parse_args(["-s1", str(p1), "-s2", str(p2), "-pts", str(SCHEME), "-apf", str(apf), "-tpra",
                                          json.dumps({"tag_pairs": [["A/1-GLY", "A/2-ALA"]], "accepted_pairs": False, "ordered": False}),
                                          "-wfa", json.dumps({"function_name": "kumaraswamy", "parameters": [3, 10, 2, 5]}), "-udc", "8"])
    _, rule2, *_ = cli.prepare(args2)
    assert rule2["tag_pairs"] == {("A/1-GLY", "A/2-ALA")} and args2.upper_distance_cutoff == 8.0


@pytest.mark.gpu
def test_cli_scores_match_the_oracle(tmp_path, capsys, oracle_mod):
    import loco_hd.__main__ as cli

    p1, p2, apf, entries = _write_case(tmp_path)
    argv = ["-s1", str(p1), "-s2", str(p2), "-pts", str(SCHEME), "-apf", str(apf)]
    assert cli.main(argv) == 0
    lines = capsys.readouterr().out.strip().splitlines()
    assert len(lines) == len(entries)
    got = []
    for line, entry in zip(lines, entries):
        head, value = line.rsplit(" = ", 1)
        assert head == f"LoCoHD({entry})"
        got.append(float(value))
    types, rule, atoms_a, atoms_b, pairs, _ = cli.prepare(cli.build_parser().parse_args(argv))
    tags = sorted({a.tag for a in atoms_a} | {b.tag for b in atoms_b})
    flat = lambda atoms: (np.array([a.coordinates for a in atoms]), np.array([types.index(a.primitive_type) for a in atoms]),
                          np.array([tags.index(a.tag) for a in atoms]))
    op = oracle_mod.Params(len(types), [("uniform", [3.0, 10.0])], tag_rule=rule)
    want = oracle_mod.from_primitives(op, *flat(atoms_a), *flat(atoms_b), np.array(pairs, np.uint32), 10.0)
    assert np.abs(np.array(got) - want).max() <= 1e-9 and np.std(want) > 0
