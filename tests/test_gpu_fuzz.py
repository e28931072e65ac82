"""Randomised differential test of the CUDA path against the oracle: random sizes (including anchor counts around the
boundaries of the work distribution - runs of 16 pairs, 2 048 / 32 768-anchor thresholds of the gather), densities,
category counts, weight functions, statistical distances, category weights, tag rules, thresholds and anchor
pairings, through the batch entry points (structures / environments / jobs) as well as the one-call entry point.
Fixed seed; LOCOHD_FUZZ_CASES raises the number of cases (default 36)."""
import os

import numpy as np
import pytest

from helpers import SCORE_TOL, assert_scores_close, set_both

pytestmark = pytest.mark.gpu

JOB = [("a_first", "<u8"), ("b_first", "<u8"), ("n", "<u8")]
WFS = [("uniform", (3.0, 10.0)), ("uniform", (0.0, 7.5)), ("kumaraswamy", (3.0, 10.0, 2.0, 5.0)), ("kumaraswamy", (1.0, 12.0, 3.0, 2.0)),
       ("kumaraswamy", (2.0, 9.0, 1.7, 3.3)), ("hyper_exp", (1.0, 0.2)), ("hyper_exp", (0.5, 0.9, 0.6, 0.1)), ("dagum", (1.7, 2.9, 11.0))]
SDS = [("Hellinger", (2.0,)), ("Hellinger", (2.0,)), ("Hellinger", (2.0,)), ("Hellinger", (3.3,)), ("Kolmogorov-Smirnov", ()),
       ("Kullback-Leibler", (0.7,)), ("Renyi", (2.3, 0.7))]


def _case(rng):
    C = int(rng.integers(2, 12))
    n_a = int(rng.choice([40, 150, 600, 1500, 2049, 4000]))
    n_b = max(20, n_a + int(rng.integers(-15, 16)))
    density = rng.choice([0.01, 0.03, 0.06])                     # primitives per cubic angstrom
    extent = 0.5 * (max(n_a, n_b) / density) ** (1.0 / 3.0)
    mk = lambda n: (rng.uniform(-extent, extent, (n, 3)), rng.integers(0, C, n).astype(np.uint16),
                    rng.integers(0, max(2, n // int(rng.choice([1, 4, 9]))), n).astype(np.uint32))
    A, B = mk(n_a), mk(n_b)
    n_pairs = int(rng.choice([1, 15, 16, 17, 33, 255, 1000, min(n_a, n_b)]))
    anchors = np.stack([rng.integers(0, n_a, n_pairs), rng.integers(0, n_b, n_pairs)], axis=1).astype(np.uint32)
    n_wf = int(rng.choice([1, 1, 1, 3]))
    wfs = [WFS[k] for k in rng.choice(len(WFS), n_wf, replace=False)]
    sd = SDS[int(rng.integers(len(SDS)))]
    weights = None if rng.random() < 0.7 else rng.uniform(0.3, 3.0, C)
    kind = rng.random()
    if kind < 0.5:
        rule = {"accept_same": False}
    elif kind < 0.7:
        rule = None
    else:
        n_tags = int(max(A[2].max(), B[2].max())) + 1
        pairs = [(int(x), int(y)) for x, y in rng.integers(0, n_tags, (int(rng.integers(1, 40)), 2))]
        rule = {"tag_pairs": pairs, "accepted_pairs": bool(rng.random() < 0.5), "ordered": bool(rng.random() < 0.5)}
        if not rule["accepted_pairs"] or len(pairs) > 10:
            pass
    threshold = float(rng.choice([6.0, 10.0, 14.5]))
    wf_idx = rng.integers(0, n_wf, n_pairs).astype(np.uint32) if n_wf > 1 else None
    return C, A, B, anchors, wfs, sd, weights, rule, threshold, wf_idx


def test_random_configurations_against_the_oracle(gpu_ctx, oracle_mod):
    rng = np.random.default_rng(20261017)
    n_cases = int(os.environ.get("LOCOHD_FUZZ_CASES", "36"))
    worst, executed = 0.0, 0
    for case in range(n_cases):
        C, A, B, anchors, wfs, sd, weights, rule, threshold, wf_idx = _case(rng)
        op = set_both(gpu_ctx, oracle_mod, C, wfs, weights, sd, rule)
        try:
            ref = oracle_mod.from_primitives(op, *A, *B, anchors, threshold, wf_idx=wf_idx, debug=True)
        except oracle_mod.OracleError:
            continue                     # e.g. a rule that leaves an anchor without any weight: covered elsewhere
        executed += 1
        got = gpu_ctx.from_primitives(*A, *B, anchors, threshold, wf_idx=wf_idx)
        assert_scores_close(got, ref["scores"])
        fin = np.isfinite(ref["scores"])
        if fin.any():
            worst = max(worst, float(np.abs(got[fin] - ref["scores"][fin]).max()))
        # the same through resident structures / environments / jobs, environments checked by their sizes
        offs = np.array([0, len(A[1]), len(A[1]) + len(B[1])], dtype=np.uint64)
        st = gpu_ctx.structs_create(offs, np.concatenate([A[0], B[0]]), np.concatenate([A[1], B[1]]), np.concatenate([A[2], B[2]]))
        a_struct = np.concatenate([np.zeros(len(anchors), np.uint32), np.ones(len(anchors), np.uint32)])
        env = gpu_ctx.envset_build(st, np.concatenate([anchors[:, 0], anchors[:, 1]]), threshold, anchor_struct=a_struct)
        from loco_hd_b200 import _capi
        off = np.empty(2 * len(anchors) + 1, np.uint64)
        gpu_ctx._check(gpu_ctx.lib.locohd_envset_dump(gpu_ctx.h, env.h, _capi._p(off), None, None, None))
        sizes = np.diff(off).astype(np.int64).reshape(2, -1).T
        assert np.array_equal(sizes, ref["env_sizes"]), f"case {case}: environment sizes differ"
        jobs = np.array([(0, len(anchors), len(anchors))], dtype=JOB)
        batch = gpu_ctx.score_jobs(env, env, jobs, wf_idx=wf_idx)
        assert np.array_equal(np.isfinite(batch), np.isfinite(got)) and np.array_equal(batch[fin], got[fin]), f"case {case}: batch entry point differs"
        env.close(); st.close()
    print(f"fuzz: {executed} of {n_cases} cases executed, max |score - oracle| = {worst:.3e}")
    assert worst <= SCORE_TOL and executed >= 0.9 * n_cases
