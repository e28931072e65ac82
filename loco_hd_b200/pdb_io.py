"""A small PDB reader for the command-line tool (``python -m loco_hd``): ATOM / HETATM records of one model ->
the duck-typed structure ``PrimitiveAssigner`` works on (``get_residues()`` -> residues with ``resname``, ``full_id``
and ``get_atoms()``; atoms with ``name`` and a float32 ``coord``), i.e. the part of Bio.PDB's object tree the
reference's tool touches (``loco_hd/__main__.py:149-170`` upstream uses ``Bio.PDB.PDBParser``; BioPython is not
available in this image and is not needed here).

Conventions taken over from Bio.PDB so that primitive tags and anchor identifiers come out the same:
``full_id = (structure_id, model_index, chain_id, (hetero_flag, residue_number, insertion_code))`` with the hetero flag
``" "`` for ATOM records, ``"W"`` for waters and ``"H_<resname>"`` for other HETATM residues; models are numbered by
their order in the file; coordinates are float32; of the alternate locations of an atom only the first one is kept.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterator, List, Tuple, Union

import numpy as np


class PdbAtom:
    __slots__ = ("name", "coord", "element", "altloc", "bfactor", "occupancy")

    def __init__(self, name: str, coord, element: str = "", altloc: str = " ", bfactor: float = 0.0,
                 occupancy: float = 1.0):
        self.name, self.element, self.altloc = name, element, altloc
        self.coord = np.asarray(coord, dtype=np.float32)
        self.bfactor, self.occupancy = bfactor, occupancy


class PdbResidue:
    def __init__(self, resname: str, full_id):
        self.resname, self.full_id = resname, full_id
        self._atoms: List[PdbAtom] = []
        self._names: Dict[str, int] = {}

    def get_atoms(self) -> Iterator[PdbAtom]:
        return iter(self._atoms)

    def __len__(self):
        return len(self._atoms)


class PdbModel:
    def __init__(self, index: int):
        self.index = index
        self._residues: List[PdbResidue] = []
        self._by_key: Dict[Tuple, PdbResidue] = {}

    def get_residues(self) -> Iterator[PdbResidue]:
        return iter(self._residues)

    def get_atoms(self) -> Iterator[PdbAtom]:
        for residue in self._residues:
            yield from residue.get_atoms()

    def atom_coordinates(self) -> np.ndarray:
        """[n_atoms, 3] float32 in ``get_residues()`` / ``get_atoms()`` order (what ``compile_topology`` numbers)."""
        return np.array([a.coord for a in self.get_atoms()], dtype=np.float32).reshape(-1, 3)


def parse_pdb(source: Union[str, Path], structure_id: str = "s") -> List[PdbModel]:
    """All models of a PDB file (or of PDB-formatted text containing a newline) in file order."""
    text = source if isinstance(source, str) and "\n" in source else Path(source).read_text()
    models: List[PdbModel] = []
    current: Union[PdbModel, None] = None
    for line in text.splitlines():
        record = line[:6]
        if record.startswith("MODEL"):
            current = PdbModel(len(models))
            models.append(current)
            continue
        if record.startswith("ENDMDL"):
            current = None
            continue
        if record not in ("ATOM  ", "HETATM"):
            continue
        if current is None:   # a file without MODEL records is one model
            current = PdbModel(len(models))
            models.append(current)
        try:
            name = line[12:16].strip()
            altloc = line[16]
            resname = line[17:20].strip()
            chain = line[21]
            number = int(line[22:26])
            icode = line[26]
            coord = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
        except (ValueError, IndexError) as exc:
            raise ValueError(f"malformed PDB coordinate record: {line!r}") from exc
        occupancy = float(line[54:60]) if line[54:60].strip() else 1.0
        bfactor = float(line[60:66]) if line[60:66].strip() else 0.0
        element = line[76:78].strip() if len(line) >= 78 else ""
        hetero = " " if record == "ATOM  " else ("W" if resname in ("HOH", "WAT") else f"H_{resname}")
        key = (chain, hetero, number, icode)
        residue = current._by_key.get(key)
        if residue is None:
            residue = PdbResidue(resname, (structure_id, current.index, chain, (hetero, number, icode)))
            current._by_key[key] = residue
            current._residues.append(residue)
        if name in residue._names:       # a further alternate location (or a duplicate record): the first one stays
            continue
        residue._names[name] = len(residue._atoms)
        residue._atoms.append(PdbAtom(name, coord, element, altloc, bfactor, occupancy))
    if not models:
        raise ValueError("no ATOM / HETATM records found")
    return models


def load_model(path: Union[str, Path], model_number: int = 0, structure_id: str = "s") -> PdbModel:
    models = parse_pdb(path, structure_id)
    if not 0 <= model_number < len(models):
        raise IndexError(f"{path}: model {model_number} requested, the file has {len(models)} model(s)")
    return models[model_number]
