"""PrimitiveAssigner and the primitive_typings JSON schemes (Python side of the reference API,
/root/reference/loco_hd/atom_converter_utils.py:19-168, primitive_typings/*.config.json) on duck-typed structures -
BioPython is not needed - plus the trajectory helpers compile_topology / assign_from_coordinates."""
import json

import numpy as np
import pytest

from loco_hd_b200.atom_converter_utils import (TYPING_DIR, PrimitiveAssigner, PrimitiveAtomSource, PrimitiveAtomTemplate,
                                               TypingSchemeElement)


class Atom:
    def __init__(self, name, coord):
        self.name, self.coord = name, np.asarray(coord, dtype=np.float32)   # Bio.PDB coordinates are float32


class Residue:
    def __init__(self, resname, number, atoms, chain="A"):
        self.resname = resname
        self.full_id = ("s", 0, chain, (" ", number, " "))
        self._atoms = atoms

    def get_atoms(self):
        return iter(self._atoms)


class Structure:
    def __init__(self, residues):
        self._residues = residues

    def get_residues(self):
        return iter(self._residues)


BACKBONE = ["N", "CA", "C", "O"]
SIDE = {"GLY": [], "ALA": ["CB"], "SER": ["CB", "OG"], "LYS": ["CB", "CG", "CD", "CE", "NZ"],
        "ASP": ["CB", "CG", "OD1", "OD2"], "PHE": ["CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ"]}


def make_structure(seed=0, names=("GLY", "ALA", "SER", "LYS", "ASP", "PHE", "ALA", "GLY")):
    rng = np.random.default_rng(seed)
    residues = []
    for k, name in enumerate(names):
        centre = rng.uniform(-20, 20, 3)
        atoms = [Atom(a, centre + rng.normal(0, 1.5, 3)) for a in BACKBONE + SIDE[name]]
        atoms.append(Atom("H", centre + rng.normal(0, 1.5, 3)))   # hydrogens never enter a centroid
        residues.append(Residue(name, k + 1, atoms))
    return Structure(residues)


# scheme name -> (number of primitive types, number of rules): SURVEY.md 2.1 row 13
SCHEMES = {"all_atom": (7, 79), "all_atom_with_centroid": (8, 80), "coarse_grained": (7, 31),
           "coarse_grained_with_centroid": (8, 32)}


@pytest.mark.parametrize("scheme", sorted(SCHEMES))
def test_typing_schemes_load(scheme):
    path = TYPING_DIR / f"{scheme}.config.json"
    table = json.loads(path.read_text())
    assigner = PrimitiveAssigner(path)
    n_types, n_rules = SCHEMES[scheme]
    assert len(table) == n_types and len(assigner.all_primitive_types) == n_types
    assert len(assigner.scheme) == n_rules == sum(len(rules) for rules in table.values())
    assert all(isinstance(e, TypingSchemeElement) for e in assigner.scheme)
    assert ("Cent" in assigner.all_primitive_types) == scheme.endswith("with_centroid")
    with pytest.raises(Exception):
        assigner.all_primitive_types = ["x"]


@pytest.mark.parametrize("scheme", sorted(SCHEMES))
def test_assign_primitive_structure(scheme):
    assigner = PrimitiveAssigner(TYPING_DIR / f"{scheme}.config.json")
    structure = make_structure(1)
    templates = assigner.assign_primitive_structure(structure)
    assert templates and all(isinstance(t, PrimitiveAtomTemplate) for t in templates)
    by_residue = {}
    for t in templates:
        assert isinstance(t.atom_source, PrimitiveAtomSource)
        assert t.primitive_type in assigner.all_primitive_types
        assert all(not a.startswith("H") for a in t.atom_source.source_atom)   # no rule of the schemes takes hydrogens
        by_residue.setdefault(t.atom_source.source_residue, []).append(t)
    residues = list(structure.get_residues())
    assert set(by_residue) == {r.full_id for r in residues}      # every residue yields at least one primitive
    for residue in residues:
        atoms = {a.name: a.coord for a in residue.get_atoms()}
        for t in by_residue[residue.full_id]:
            # the position is the plain mean of the contributing atoms (float32, as np.mean of Bio.PDB coordinates)
            expected = np.mean([atoms[a] for a in t.atom_source.source_atom], axis=0)
            assert t.coordinates.dtype == np.float32
            assert np.array_equal(t.coordinates, expected)
            assert t.atom_source.source_residue_name == residue.resname
    if scheme.endswith("with_centroid"):
        cents = [t for t in templates if t.primitive_type == "Cent"]
        assert len(cents) == len(residues)                      # one centroid per residue ...
        for residue, t in zip(residues, cents):
            heavy = [a.name for a in residue.get_atoms() if not a.name.startswith("H")]
            assert t.atom_source.source_atom == heavy           # ... over all non-hydrogen atoms


@pytest.mark.parametrize("scheme", sorted(SCHEMES))
def test_compiled_topology_matches_per_residue_path(scheme):
    assigner = PrimitiveAssigner(TYPING_DIR / f"{scheme}.config.json")
    structure = make_structure(2)
    templates = assigner.assign_primitive_structure(structure)
    topology = assigner.compile_topology(structure)
    assert topology.primitive_types == [t.primitive_type for t in templates]
    assert [s.source_atom for s in topology.sources] == [t.atom_source.source_atom for t in templates]
    coords = np.array([a.coord for r in structure.get_residues() for a in r.get_atoms()], dtype=np.float32)
    assert topology.n_atoms == len(coords)
    frame0 = PrimitiveAssigner.assign_from_coordinates(topology, coords)
    assert frame0.dtype == np.float32
    np.testing.assert_allclose(frame0, np.array([t.coordinates for t in templates]), rtol=0, atol=1e-5)
    # a second "frame": same topology, moved atoms
    moved = coords + np.float32(0.25)
    np.testing.assert_allclose(PrimitiveAssigner.assign_from_coordinates(topology, moved), frame0 + np.float32(0.25),
                               rtol=0, atol=1e-5)
    with pytest.raises(ValueError):
        PrimitiveAssigner.assign_from_coordinates(topology, coords[:-1])


def test_generate_primitive_pdb_format():
    assigner = PrimitiveAssigner(TYPING_DIR / "coarse_grained_with_centroid.config.json")
    templates = assigner.assign_primitive_structure(make_structure(3))
    text = assigner.generate_primitive_pdb(templates, b_labels=np.linspace(0, 1, len(templates)))
    lines = text.splitlines()
    assert len(lines) == len(templates) and all(line.startswith("ATOM  ") and line.rstrip().endswith("Pr") for line in lines)
    # fixed columns of the PDB record: coordinates in columns 31-54 with three decimals
    x, y, z = templates[0].coordinates[:3]
    assert lines[0][30:54] == f"{x:8.3f}{y:8.3f}{z:8.3f}"
