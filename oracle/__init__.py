"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/locohd_oracle.cpp``, the CPU restatement of the reference's
per-anchor scoring path (``/root/reference/src/locohd.rs:61-226, 479-567`` and the leaf
files cited inside the C++ source).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product (``loco_hd_b200``, ``loco_hd``) never does.

Parity status: pinned against the reference's known-answer tests only (the Rust crate
cannot be built in this image and its golden outputs are missing upstream); the
neighbour-membership boundary decided by the un-vendored ``kd-tree`` 0.6 crate is
"parity unpinned".  See DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "locohd_oracle.cpp"
_LIB = _HERE / "_build" / "liblocohd_oracle.so"

UNKNOWN_CAT = 0xFFFF
WF_KINDS = {"hyper_exp": 0, "dagum": 1, "uniform": 2, "kumaraswamy": 3}
SD_KINDS = {"Hellinger": 0, "Kolmogorov-Smirnov": 1, "Kullback-Leibler": 2, "Renyi": 3}
STATUS = {
    0: "ok", 1: "len mismatch", 2: "dists must start with 0", 3: "unknown category", 4: "zero norm",
    5: "negative integral point", 6: "NaN", 7: "empty environment", 8: "index out of range",
    9: "distance matrix shape", 10: "bad parameter",
}


class OracleError(ValueError):
    def __init__(self, status: int):
        super().__init__(f"oracle status {status}: {STATUS.get(status, '?')}")
        self.status = status


def build(force: bool = False) -> Path:
    """Compile the restatement with g++ (no FMA contraction, OpenMP) into oracle/_build/."""
    if _LIB.exists() and not force and _LIB.stat().st_mtime >= _SRC.stat().st_mtime:
        return _LIB
    _LIB.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-o", str(_LIB), str(_SRC)]
    subprocess.run(cmd, check=True)
    return _LIB


class _Params(C.Structure):
    _fields_ = [
        ("n_categories", C.c_int32),
        ("category_weights", C.POINTER(C.c_double)),
        ("sd_kind", C.c_int32),
        ("sd_params", C.c_double * 2),
        ("n_wf", C.c_int32),
        ("wf_kind", C.POINTER(C.c_int32)),
        ("wf_nparams", C.POINTER(C.c_int32)),
        ("wf_offset", C.POINTER(C.c_int32)),
        ("wf_params", C.POINTER(C.c_double)),
        ("tpr_kind", C.c_int32),
        ("tpr_accept_same", C.c_int32),
        ("tpr_accepted_pairs", C.c_int32),
        ("tpr_ordered", C.c_int32),
        ("n_tag_pairs", C.c_uint64),
        ("tag_pairs", C.POINTER(C.c_uint64)),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists():
            build()
        _lib = C.CDLL(str(_LIB))
        _lib.oracle_environment.restype = C.c_int64
    return _lib


def _ptr(a: Optional[np.ndarray], ty):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ty))


@dataclass
class Params:
    """Interned parameter block: what ``LoCoHD::build`` holds (locohd.rs:42-55, 289-389)."""

    n_categories: int
    weight_functions: List[Tuple[str, Sequence[float]]] = field(default_factory=lambda: [("uniform", [3.0, 10.0])])
    category_weights: Optional[Sequence[float]] = None
    statistical_distance: Tuple[str, Sequence[float]] = ("Hellinger", [2.0])
    # tag rule on interned tag ids: {"accept_same": bool} or {"tag_pairs": [(a, b)], "accepted_pairs", "ordered"}
    tag_rule: Optional[dict] = None

    def pack(self):
        keep = []
        p = _Params()
        p.n_categories = self.n_categories
        cw = np.ascontiguousarray(
            np.ones(self.n_categories) if self.category_weights is None else self.category_weights, dtype=np.float64)
        keep.append(cw)
        p.category_weights = _ptr(cw, C.c_double)
        name, sdp = self.statistical_distance
        p.sd_kind = SD_KINDS[name]
        sdp = list(sdp) + [0.0, 0.0]
        p.sd_params[0], p.sd_params[1] = sdp[0], sdp[1]
        kinds = np.array([WF_KINDS[n] for n, _ in self.weight_functions], dtype=np.int32)
        nps = np.array([len(q) for _, q in self.weight_functions], dtype=np.int32)
        offs = np.concatenate([[0], np.cumsum(nps)[:-1]]).astype(np.int32)
        flat = np.ascontiguousarray(
            np.concatenate([np.asarray(q, dtype=np.float64).ravel() for _, q in self.weight_functions] + [np.zeros(1)]))
        keep += [kinds, nps, offs, flat]
        p.n_wf = len(self.weight_functions)
        p.wf_kind, p.wf_nparams, p.wf_offset = _ptr(kinds, C.c_int32), _ptr(nps, C.c_int32), _ptr(offs, C.c_int32)
        p.wf_params = _ptr(flat, C.c_double)
        rule = self.tag_rule if self.tag_rule is not None else {"accept_same": True}  # locohd.rs:357-362
        if "accept_same" in rule:
            p.tpr_kind, p.tpr_accept_same = 0, int(bool(rule["accept_same"]))
            pairs = np.zeros(1, dtype=np.uint64)
            p.n_tag_pairs = 0
        else:
            p.tpr_kind = 1
            p.tpr_accepted_pairs = int(bool(rule["accepted_pairs"]))
            p.tpr_ordered = int(bool(rule["ordered"]))
            pairs = np.array(sorted({(int(a) << 32) | int(b) for a, b in rule["tag_pairs"]}) or [0], dtype=np.uint64)
            p.n_tag_pairs = len(set(rule["tag_pairs"]))
        keep.append(pairs)
        p.tag_pairs = _ptr(pairs, C.c_uint64)
        return p, keep


def _check(st: int):
    if st != 0:
        raise OracleError(st)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def wf_integral_point(name: str, params: Sequence[float], x: float) -> float:
    q = np.asarray(params, dtype=np.float64)
    out = C.c_double()
    _check(lib().oracle_wf_integral_point(WF_KINDS[name], len(q), _ptr(q, C.c_double), C.c_double(x), C.byref(out)))
    return out.value


def wf_integral_range(name: str, params: Sequence[float], a: float, b: float) -> float:
    q = np.asarray(params, dtype=np.float64)
    out = C.c_double()
    _check(lib().oracle_wf_integral_range(WF_KINDS[name], len(q), _ptr(q, C.c_double), C.c_double(a), C.c_double(b),
                                          C.byref(out)))
    return out.value


def sd_run(name: str, params: Sequence[float], p1: Sequence[float], p2: Sequence[float]) -> float:
    sdp = np.asarray(list(params) + [0.0, 0.0], dtype=np.float64)
    a, b = np.ascontiguousarray(p1, dtype=np.float64), np.ascontiguousarray(p2, dtype=np.float64)
    out = C.c_double()
    _check(lib().oracle_sd_run(SD_KINDS[name], _ptr(sdp, C.c_double), len(a), _ptr(a, C.c_double),
                               _ptr(b, C.c_double), C.byref(out)))
    return out.value


def tag_pair_accepted(params: Params, anchor_tag: int, other_tag: int) -> bool:
    p, keep = params.pack()
    return bool(lib().oracle_tag_pair_accepted(C.byref(p), C.c_uint32(anchor_tag), C.c_uint32(other_tag)))


def from_anchors(params: Params, seq_a, seq_b, dists_a, dists_b, wf_idx: int = 0, return_steps=False):
    p, keep = params.pack()
    sa, sb = np.ascontiguousarray(seq_a, dtype=np.uint16), np.ascontiguousarray(seq_b, dtype=np.uint16)
    da, db = np.ascontiguousarray(dists_a, dtype=np.float64), np.ascontiguousarray(dists_b, dtype=np.float64)
    out, steps = C.c_double(), C.c_uint64()
    _check(lib().oracle_from_anchors(C.byref(p), _ptr(sa, C.c_uint16), C.c_uint64(len(sa)), _ptr(sb, C.c_uint16),
                                     C.c_uint64(len(sb)), _ptr(da, C.c_double), C.c_uint64(len(da)),
                                     _ptr(db, C.c_double), C.c_uint64(len(db)), wf_idx, C.byref(out),
                                     C.byref(steps)))
    return (out.value, steps.value) if return_steps else out.value


def environment(params: Params, xyz, cat, tag, anchor: int, threshold: float, use_tree: bool = True):
    """Sorted environment of one anchor: (orig indices, distances, categories) in reference order."""
    p, keep = params.pack()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = len(xyz)
    cat = np.ascontiguousarray(cat, dtype=np.uint16)
    tag = np.ascontiguousarray(tag, dtype=np.uint32)
    idx, d, c = np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.uint16)
    m = lib().oracle_environment(C.byref(p), C.c_uint64(n), _ptr(xyz, C.c_double), _ptr(cat, C.c_uint16),
                                 _ptr(tag, C.c_uint32), C.c_uint64(anchor), C.c_double(threshold), int(use_tree),
                                 _ptr(idx, C.c_uint32), _ptr(d, C.c_double), _ptr(c, C.c_uint16))
    if m < 0:
        raise OracleError(int(-m))
    return idx[:m].copy(), d[:m].copy(), c[:m].copy()


def from_primitives(params: Params, xyz_a, cat_a, tag_a, xyz_b, cat_b, tag_b, anchors, threshold: float,
                    wf_idx=None, use_tree: bool = True, n_threads: int = 0, debug: bool = False):
    """LoCoHD::from_primitives (locohd.rs:479-567). Returns scores, or with ``debug`` a dict that also
    holds env sizes [P,2], final category counts [P,2,C] and walk step counts [P]."""
    p, keep = params.pack()
    xa = np.ascontiguousarray(xyz_a, dtype=np.float64).reshape(-1, 3)
    xb = np.ascontiguousarray(xyz_b, dtype=np.float64).reshape(-1, 3)
    ca, cb = np.ascontiguousarray(cat_a, dtype=np.uint16), np.ascontiguousarray(cat_b, dtype=np.uint16)
    ta, tb = np.ascontiguousarray(tag_a, dtype=np.uint32), np.ascontiguousarray(tag_b, dtype=np.uint32)
    an = np.ascontiguousarray(anchors, dtype=np.uint32).reshape(-1, 2)
    P = len(an)
    wf = None if wf_idx is None else np.ascontiguousarray(wf_idx, dtype=np.int32)
    out = np.zeros(P, np.float64)
    sizes = np.zeros((P, 2), np.uint32) if debug else None
    counts = np.zeros((P, 2, params.n_categories), np.uint32) if debug else None
    steps = np.zeros(P, np.uint64) if debug else None
    st = lib().oracle_from_primitives(
        C.byref(p), C.c_uint64(len(xa)), _ptr(xa, C.c_double), _ptr(ca, C.c_uint16), _ptr(ta, C.c_uint32),
        C.c_uint64(len(xb)), _ptr(xb, C.c_double), _ptr(cb, C.c_uint16), _ptr(tb, C.c_uint32), C.c_uint64(P),
        _ptr(an, C.c_uint32), _ptr(wf, C.c_int32), C.c_double(threshold), int(use_tree), int(n_threads),
        _ptr(out, C.c_double), _ptr(sizes, C.c_uint32), _ptr(counts, C.c_uint32), _ptr(steps, C.c_uint64))
    _check(st)
    if debug:
        return {"scores": out, "env_sizes": sizes, "counts": counts, "steps": steps}
    return out


def from_dmxs(params: Params, seq_a, seq_b, dmx_a, dmx_b, wf_idx=None, n_threads: int = 0):
    p, keep = params.pack()
    sa, sb = np.ascontiguousarray(seq_a, dtype=np.uint16), np.ascontiguousarray(seq_b, dtype=np.uint16)
    ma, mb = np.ascontiguousarray(dmx_a, dtype=np.float64), np.ascontiguousarray(dmx_b, dtype=np.float64)
    ma, mb = ma.reshape(len(ma), -1), mb.reshape(len(mb), -1)
    wf = None if wf_idx is None else np.ascontiguousarray(wf_idx, dtype=np.int32)
    out = np.zeros(len(ma), np.float64)
    _check(lib().oracle_from_dmxs(
        C.byref(p), _ptr(sa, C.c_uint16), C.c_uint64(len(sa)), _ptr(sb, C.c_uint16), C.c_uint64(len(sb)),
        _ptr(ma, C.c_double), C.c_uint64(ma.shape[0]), C.c_uint64(ma.shape[1]), _ptr(mb, C.c_double),
        C.c_uint64(mb.shape[0]), C.c_uint64(mb.shape[1]), _ptr(wf, C.c_int32), int(n_threads), _ptr(out, C.c_double)))
    return out


def from_coords(params: Params, seq_a, seq_b, xyz_a, xyz_b, wf_idx=None, n_threads: int = 0):
    p, keep = params.pack()
    sa, sb = np.ascontiguousarray(seq_a, dtype=np.uint16), np.ascontiguousarray(seq_b, dtype=np.uint16)
    xa = np.ascontiguousarray(xyz_a, dtype=np.float64).reshape(-1, 3)
    xb = np.ascontiguousarray(xyz_b, dtype=np.float64).reshape(-1, 3)
    wf = None if wf_idx is None else np.ascontiguousarray(wf_idx, dtype=np.int32)
    out = np.zeros(len(xa), np.float64)
    _check(lib().oracle_from_coords(
        C.byref(p), _ptr(sa, C.c_uint16), C.c_uint64(len(sa)), _ptr(sb, C.c_uint16), C.c_uint64(len(sb)),
        _ptr(xa, C.c_double), C.c_uint64(len(xa)), _ptr(xb, C.c_double), C.c_uint64(len(xb)), _ptr(wf, C.c_int32),
        int(n_threads), _ptr(out, C.c_double)))
    return out
