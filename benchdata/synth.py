"""Synthetic labelled point clouds of the BASELINE.json shapes (SURVEY.md §8(d) generator ``GEN``).

Protein-like: residue centres uniform in a ball of protein density (135 A^3 per residue), k primitives per residue
scattered with sigma = 1.5 A around the centre, categories uniform over C types, one tag per residue (the reference's
callers use "chain/resnum-RESNAME" strings, casp14_extend_with_locohd.py:13-18; here the residue number is the
interned tag id).  Deterministic for a given seed (numpy PCG64).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Cloud:
    xyz: np.ndarray   # [N, 3] f64
    cat: np.ndarray   # [N] u16
    tag: np.ndarray   # [N] u32 (residue id)
    k: int            # primitives per residue
    centroid_cat: int = -1  # category id of the "Cent" primitive (slot 0 of every residue) or -1

    @property
    def n(self) -> int:
        return len(self.cat)

    def centroid_anchors(self) -> np.ndarray:
        return np.arange(0, self.n, self.k, dtype=np.uint32)


def gen(seed: int, n_residues: int, k: int, n_categories: int, sigma: float = 1.5, v_res: float = 135.0,
        with_centroid: bool = False, f32_exact: bool = False) -> Cloud:
    rng = np.random.default_rng(seed)
    rho = (3.0 * n_residues * v_res / (4.0 * np.pi)) ** (1.0 / 3.0)
    direction = rng.normal(size=(n_residues, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    centres = direction * (rho * rng.random(n_residues) ** (1.0 / 3.0))[:, None]
    xyz = np.repeat(centres, k, axis=0) + rng.normal(0.0, sigma, size=(n_residues * k, 3))
    centroid_cat = -1
    if with_centroid:
        # slot 0 of each residue is the "Cent" category, placed at the mean of the residue's other slots
        centroid_cat = n_categories - 1
        cat = rng.integers(0, n_categories - 1, size=n_residues * k)
        xyz3 = xyz.reshape(n_residues, k, 3)
        xyz3[:, 0, :] = xyz3[:, 1:, :].mean(axis=1)
        cat.reshape(n_residues, k)[:, 0] = centroid_cat
    else:
        cat = rng.integers(0, n_categories, size=n_residues * k)
    if f32_exact:
        xyz = xyz.astype(np.float32).astype(np.float64)
    tag = np.repeat(np.arange(n_residues, dtype=np.uint32), k)
    return Cloud(np.ascontiguousarray(xyz), cat.astype(np.uint16), tag, k, centroid_cat)


def partner(base: Cloud, delta: float, seed: int, f32_exact: bool = False) -> Cloud:
    """Same types and tags, coordinates + normal(0, delta) (a model / frame / ensemble member of ``base``)."""
    rng = np.random.default_rng(seed)
    xyz = base.xyz + rng.normal(0.0, delta, size=base.xyz.shape)
    if f32_exact:
        xyz = xyz.astype(np.float32).astype(np.float64)
    return Cloud(np.ascontiguousarray(xyz), base.cat.copy(), base.tag.copy(), base.k, base.centroid_cat)


# ---- the five BASELINE.json configurations --------------------------------------------
def config1():
    a = gen(1, 150, 3, 7)
    return a, partner(a, 1.0, 2)


def config2():
    a = gen(3, 1250, 8, 7)
    return a, partner(a, 1.0, 4)


def config2_pair(i: int):
    """i-th structure pair of the benchmark batch of configuration-2-shaped pairs (pair 0 is config2())."""
    a = gen(3 + 2 * i, 1250, 8, 7)
    return a, partner(a, 1.0, 4 + 2 * i)


def config3_reference():
    return gen(5, 300, 9, 8, with_centroid=True)


def config3_model(ref: Cloud, m: int):
    return partner(ref, 2.0, 1000 + m)


def config4_frame0():
    return gen(7, 1250, 4, 8, with_centroid=True)


def config4_frame(frame0: Cloud, t: int):
    return frame0 if t == 0 else partner(frame0, 0.5, 2000 + t)


def config5_base():
    return gen(9, 625, 8, 7)


def config5_member(base: Cloud, i: int):
    return partner(base, 1.5, 3000 + i)
