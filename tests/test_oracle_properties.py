"""CPU: size-independent properties of the LoCoHD score that the oracle must have (the GPU twins of these checks are
tests/test_gpu_parity.py::test_symmetry_and_permutation_invariance and the exact-zero tests): symmetry under swapping
the two structures (Hellinger, Kolmogorov-Smirnov), exactly 0 for identical environments, invariance under a
permutation of the primitives, 0 <= score <= 1 for a CDF weight and a bounded distance, agreement of the kd-tree and
brute-force neighbour searches, and invariance under rigid translation (to rounding).  Randomised with hypothesis
(src/locohd.rs:61-226, 479-567 are what these properties are consequences of)."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

WFS = [("uniform", [3.0, 10.0]), ("kumaraswamy", [3.0, 10.0, 2.0, 5.0]), ("hyper_exp", [1.0, 0.3]), ("dagum", [2.0, 5.0, 1.0])]


def _cloud(rng, n, C, n_tags):
    xyz = rng.uniform(-8.0, 8.0, (n, 3))
    return xyz, rng.integers(0, C, n).astype(np.uint16), rng.integers(0, n_tags, n).astype(np.uint32)


@settings(max_examples=60, deadline=None, derandomize=True, database=None)   # the same examples on every box
@given(seed=st.integers(0, 2 ** 31), n=st.integers(2, 70), C=st.integers(1, 9), wf=st.integers(0, 3),
       sd=st.sampled_from([("Hellinger", [2.0]), ("Hellinger", [3.0]), ("Kolmogorov-Smirnov", [])]),
       rule=st.sampled_from([None, {"accept_same": False}, {"accept_same": True}]), thr=st.sampled_from([4.0, 7.5, 30.0]))
def test_oracle_properties(oracle_mod, seed, n, C, wf, sd, rule, thr):
    rng = np.random.default_rng(seed)
    A, B = _cloud(rng, n, C, 5), _cloud(rng, n, C, 5)
    p = oracle_mod.Params(C, [WFS[wf]], None, sd, rule)
    k = min(n, 12)
    an = np.stack([rng.integers(0, n, k), rng.integers(0, n, k)], axis=1).astype(np.uint32)
    s = oracle_mod.from_primitives(p, *A, *B, an, thr)
    # bounded: a CDF weight integrates to at most 1 and these distances live in [0, 1]
    assert np.all(s >= 0.0) and np.all(s <= 1.0 + 1e-12)
    # symmetric in the two structures
    t = oracle_mod.from_primitives(p, *B, *A, an[:, ::-1].copy(), thr)
    assert np.abs(s - t).max() <= 1e-14
    # kd-tree and brute-force neighbour search agree exactly
    assert np.array_equal(s, oracle_mod.from_primitives(p, *A, *B, an, thr, use_tree=False))
    # a structure against itself, same anchors: exactly 0
    same = np.stack([an[:, 0], an[:, 0]], axis=1)
    assert np.all(oracle_mod.from_primitives(p, *A, *A, same, thr) == 0.0)
    # permuting the primitives of A (anchors follow) changes nothing but the order of tied neighbours
    perm = rng.permutation(n)
    inv = np.argsort(perm)
    Ap = (A[0][perm], A[1][perm], A[2][perm])
    anp = np.stack([inv[an[:, 0]], an[:, 1]], axis=1).astype(np.uint32)
    assert np.abs(oracle_mod.from_primitives(p, *Ap, *B, anp, thr) - s).max() <= 1e-14
    # rigid translation by a vector that is exact in binary: distances change by rounding only; a neighbour sitting
    # within that rounding of the sphere may enter or leave, so only anchors with a clear margin are compared
    shift = np.array([64.0, -32.0, 16.0])
    moved = oracle_mod.from_primitives(p, A[0] + shift, A[1], A[2], B[0] + shift, B[1], B[2], an, thr)
    clear = np.ones(k, dtype=bool)
    for q in range(k):
        for X, i in ((A[0], an[q, 0]), (B[0], an[q, 1])):
            d = np.sqrt(((X - X[i]) ** 2).sum(axis=1))
            clear[q] &= bool(np.all(np.abs(d - thr) > 1e-9))
    # W is Lipschitz (slope <= ~1 per Angstrom for these parameters): rounding of the distances moves the score by ~1e-13
    assert np.abs(moved[clear] - s[clear]).max(initial=0.0) <= 1e-11
