"""Shared helpers of the parity tests: run the same case through the CUDA C ABI and through the CPU oracle."""
from __future__ import annotations

import numpy as np

SCORE_TOL = 1e-9  # BASELINE.json north_star: per-anchor scores within 1e-9 absolute of the reference's f64 results


def assert_scores_close(got, ref, tol=SCORE_TOL):
    """Finite scores within tol.  Where the reference result is not finite (e.g. Renyi alpha=0 on disjoint
    compositions gives H = inf; the integral then holds inf or, through 0 * inf on a tied step, NaN - which of
    the two depends on how tied events are grouped) the CUDA result must be non-finite as well."""
    got, ref = np.atleast_1d(np.asarray(got, dtype=np.float64)), np.atleast_1d(np.asarray(ref, dtype=np.float64))
    assert got.shape == ref.shape
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), "non-finite scores in different places"
    if fin.any():
        err = np.abs(got[fin] - ref[fin])
        assert err.max() <= tol, f"max |score diff| = {err.max()} at pair {np.flatnonzero(fin)[err.argmax()]}"


def canonical_env(idx, dist):
    """Neighbour list in canonical order (distance, primitive index): tie order is immaterial to the score
     and differs between implementations."""
    order = np.lexsort((idx, dist))
    return idx[order], dist[order]


def set_both(ctx, oracle, n_categories, weight_functions=(("uniform", (3.0, 10.0)),), category_weights=None,
             statistical_distance=("Hellinger", (2.0,)), tag_rule=None):
    ctx.set_params(n_categories, weight_functions, category_weights, statistical_distance, tag_rule)
    return oracle.Params(n_categories, [(n, list(p)) for n, p in weight_functions], category_weights,
                         (statistical_distance[0], list(statistical_distance[1])), tag_rule)


def check_from_primitives(ctx, oracle, op, A, B, anchors, threshold, wf_idx=None, tol=SCORE_TOL, check_envs=True):
    """A, B: (xyz, cat, tag).  Compares scores (<= tol), and with check_envs the neighbour lists (bit-exact
    membership, canonical order), distances (bit-exact) and per-category counts."""
    anchors = np.asarray(anchors, dtype=np.uint32).reshape(-1, 2)
    got = ctx.from_primitives(A[0], A[1], A[2], B[0], B[1], B[2], anchors, threshold, wf_idx=wf_idx)
    ref = oracle.from_primitives(op, A[0], A[1], A[2], B[0], B[1], B[2], anchors, threshold, wf_idx=wf_idx,
                                 debug=True)
    assert_scores_close(got, ref["scores"], tol)
    if check_envs and len(anchors):
        for side, (S, col) in enumerate(((A, 0), (B, 1))):
            st = ctx.structure(*S)
            env = ctx.envset_build(st, anchors[:, col], threshold, keep_indices=True)
            off, d, c, ix = env.dump()
            sizes = np.diff(off).astype(np.int64)
            assert np.array_equal(sizes, ref["env_sizes"][:, side]), "environment sizes differ"
            C = op.n_categories
            # category counts: bit-exact integers
            for p in range(len(anchors)):
                cc = np.bincount(c[off[p]:off[p + 1]], minlength=C)
                assert np.array_equal(cc, ref["counts"][p, side]), f"category counts differ at pair {p}"
            # neighbour lists for a sample of anchors (the oracle call is per anchor)
            sample = np.unique(np.linspace(0, len(anchors) - 1, min(len(anchors), 24)).astype(int))
            for p in sample:
                oi, od, oc = oracle.environment(op, S[0], S[1], S[2], int(anchors[p, col]), threshold)
                gi, gd = canonical_env(ix[off[p]:off[p + 1]], d[off[p]:off[p + 1]])
                ri, rd = canonical_env(oi, od)
                assert np.array_equal(gi, ri), f"neighbour list differs at pair {p} side {side}"
                assert np.array_equal(gd, rd), f"neighbour distances differ at pair {p} side {side}"
                assert np.all(np.diff(d[off[p]:off[p + 1]]) >= 0), "environment not sorted"
            env.close()
            st.close()
    return got, ref


def random_cloud(rng, n, n_categories, extent=50.0, n_tags=None):
    xyz = rng.uniform(-extent, extent, size=(n, 3))
    cat = rng.integers(0, n_categories, size=n).astype(np.uint16)
    tag = (rng.integers(0, n_tags, size=n) if n_tags else np.zeros(n)).astype(np.uint32)
    return xyz, cat, tag
