"""Python side of the reference API: ``PrimitiveAssigner`` turns an atomistic structure into typed centroid
primitives according to a JSON typing scheme.  Written from scratch to the behaviour of the reference's
``loco_hd/atom_converter_utils.py:19-168`` (same class, method and field names, same outputs).

The structure objects are duck-typed: anything with ``get_residues()`` yielding objects that have ``resname``,
``full_id`` and ``get_atoms()`` (atoms with ``name`` and ``coord``) works — Bio.PDB entities do, so does any
light-weight stand-in (BioPython is not required to import this module).

``PrimitiveAssigner.compile_topology`` / ``assign_from_coordinates`` are additions for trajectories (SURVEY.md
§8(f) N3): the regex matching is done once per topology, every further frame is a segmented mean over numpy
arrays.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Any, List, Sequence, Tuple, Union

import numpy as np

# structure id, model id, chain id, (hetero flag, residue number, insertion code)
ResiFullIdType = Tuple[str, int, str, Tuple[str, int, str]]

TYPING_DIR = Path(__file__).resolve().parent / "primitive_typings"


@dataclass
class PrimitiveAtomSource:
    """Where a primitive atom came from: the residue, its name and the names of the contributing atoms
    (one name for a plain atom, several for a centroid)."""

    source_residue: ResiFullIdType
    source_residue_name: str
    source_atom: List[str]


@dataclass
class PrimitiveAtomTemplate:
    """Intermediate between a structure's atoms and ``PrimitiveAtom``: a type, a position and the source."""

    primitive_type: str
    coordinates: np.ndarray
    atom_source: PrimitiveAtomSource


@dataclass
class TypingSchemeElement:
    """One rule of a typing scheme: residues whose name fully matches ``residue_matcher`` get a primitive of
    ``primitive_type`` at the centroid of the atoms whose names fully match ``atom_matcher``, provided the number
    of matched atoms equals ``atom_counter`` (or ``atom_counter`` is ``"any"``)."""

    primitive_type: str
    residue_matcher: re.Pattern
    atom_matcher: re.Pattern
    atom_counter: Union[int, str]

    def match_resi(self, resi_name: str) -> bool:
        return self.residue_matcher.fullmatch(resi_name) is not None

    def match_atom(self, atom_name: str) -> bool:
        return self.atom_matcher.fullmatch(atom_name) is not None


@dataclass
class CompiledTopology:
    """Result of ``PrimitiveAssigner.compile_topology``: which atoms feed which primitive."""

    primitive_types: List[str]
    sources: List[PrimitiveAtomSource]
    atom_index: np.ndarray      # flat atom indices, grouped by primitive
    segment_start: np.ndarray   # [n_primitives + 1] offsets into atom_index
    n_atoms: int


class PrimitiveAssigner:
    """Holds a primitive typing scheme (a JSON file ``{type: [[residue_regex, atom_regex(, count)], ...]}``,
    see ``primitive_typings/``) and converts structures to lists of ``PrimitiveAtomTemplate``."""

    def __init__(self, config_path: Union[str, Path]):
        with open(config_path, "r") as handle:
            table = json.load(handle)
        self.scheme: List[TypingSchemeElement] = []
        for primitive_type, rules in table.items():
            for rule in rules:
                counter = rule[2] if len(rule) > 2 else 1
                self.scheme.append(
                    TypingSchemeElement(primitive_type, re.compile(rule[0]), re.compile(rule[1]), counter))

    @property
    def all_primitive_types(self) -> List[str]:
        # an unordered collection upstream (list(set(...))); the order here is the scheme's first-appearance order,
        # which is one of the orders the reference can return and is stable across processes
        return list(dict.fromkeys(element.primitive_type for element in self.scheme))

    @all_primitive_types.setter
    def all_primitive_types(self, value):
        raise Exception("Cannot set all_primitive_types directly, since it depends on the config file!")

    def assign_primitive_structure(self, structure: Any) -> List[PrimitiveAtomTemplate]:
        templates: List[PrimitiveAtomTemplate] = []
        for residue in structure.get_residues():
            name, full_id = residue.resname, residue.full_id
            atoms = list(residue.get_atoms())
            for element in self.scheme:
                if not element.match_resi(name):
                    continue
                hits = [atom for atom in atoms if element.match_atom(atom.name)]
                if element.atom_counter != "any" and element.atom_counter != len(hits):
                    continue
                centroid = np.mean([atom.coord for atom in hits], axis=0)
                source = PrimitiveAtomSource(full_id, name, [atom.name for atom in hits])
                templates.append(PrimitiveAtomTemplate(element.primitive_type, centroid, source))
        return templates

    # ---- trajectory helpers (not in the reference) ---------------------------------------------------------
    def compile_topology(self, structure: Any) -> CompiledTopology:
        """Resolve the regex scheme once for a topology.  Atoms are numbered in ``get_residues()`` /
        ``get_atoms()`` iteration order; ``assign_from_coordinates`` expects coordinates in that order."""
        types: List[str] = []
        sources: List[PrimitiveAtomSource] = []
        flat: List[int] = []
        starts = [0]
        cursor = 0
        for residue in structure.get_residues():
            atoms = list(residue.get_atoms())
            for element in self.scheme:
                if not element.match_resi(residue.resname):
                    continue
                hits = [k for k, atom in enumerate(atoms) if element.match_atom(atom.name)]
                if element.atom_counter != "any" and element.atom_counter != len(hits):
                    continue
                types.append(element.primitive_type)
                sources.append(PrimitiveAtomSource(residue.full_id, residue.resname, [atoms[k].name for k in hits]))
                flat.extend(cursor + k for k in hits)
                starts.append(len(flat))
            cursor += len(atoms)
        return CompiledTopology(types, sources, np.asarray(flat, dtype=np.int64), np.asarray(starts, dtype=np.int64),
                                cursor)

    @staticmethod
    def assign_from_coordinates(topology: CompiledTopology, coordinates: np.ndarray) -> np.ndarray:
        """Centroids [n_primitives, 3] of one frame ([n_atoms, 3], the dtype is kept: float32 positions give the
        float32 means the per-residue path produces)."""
        coordinates = np.asarray(coordinates)
        if coordinates.shape != (topology.n_atoms, 3):
            raise ValueError(f"expected coordinates of shape ({topology.n_atoms}, 3), got {coordinates.shape}")
        counts = np.diff(topology.segment_start)
        if len(counts) == 0:
            return np.zeros((0, 3), dtype=coordinates.dtype)
        gathered = coordinates[topology.atom_index]
        out = np.full((len(counts), 3), np.nan, dtype=coordinates.dtype)
        filled = counts > 0
        if filled.any():
            sums = np.add.reduceat(gathered, topology.segment_start[:-1][filled], axis=0)
            out[filled] = sums / counts[filled, None].astype(coordinates.dtype)
        return out

    def generate_primitive_pdb(self, primitive_structure: Sequence[PrimitiveAtomTemplate],
                               b_labels: Union[None, Sequence[float], np.ndarray] = None) -> str:
        """PDB text with one pseudo-atom per primitive (atom name = letter of its type, element ``Pr``)."""
        type_order = self.all_primitive_types
        lines = []
        previous_residue = None
        residue_number = 0
        for serial, template in enumerate(primitive_structure, start=1):
            source = template.atom_source
            if source.source_residue != previous_residue:
                residue_number += 1
                previous_residue = source.source_residue
            b_factor = 1.0 if b_labels is None else b_labels[serial - 1]
            letter = chr(ord("A") + type_order.index(template.primitive_type))
            x, y, z = template.coordinates[:3]
            lines.append(
                "ATOM  " + f"{serial: >5} " + f"{letter: >4}" + " " + f"{source.source_residue_name} "
                + f"{source.source_residue[1]}" + f"{residue_number: >4}" + "    "
                + f"{x:8.3f}{y:8.3f}{z:8.3f}" + f"{1.0:6.2f}" + f"{b_factor:6.2f}" + " " * 10 + "Pr" + "  ")
        return "".join(line + "\n" for line in lines)
