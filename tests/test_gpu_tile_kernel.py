"""score_tile_kernel: job lists whose jobs share runs of environments (the all-vs-all ensemble of
compare_ensembles.py:250-296 in batch.blocked_pairs order) are scored tile by tile - the <= 8 environments of
(tile, anchor) staged once, <= 16 anchor pairs by a team of four warps with 8 lanes per pair.  Same arithmetic as the
one-pair-per-warp kernel (src/locohd.rs:61-226 as a flat prefix scan), different chunking: checked against the CPU
oracle (1e-9 on every per-anchor score), against the one-pair-per-warp kernel on the same jobs, and on the cases
where the two paths differ in structure (partial tiles, odd / even environment sizes, identical members, the
small-H^2 branch, the unknown-category error).
"""
import numpy as np
import pytest

from benchdata import synth
from helpers import SCORE_TOL, assert_scores_close, set_both
from loco_hd_b200 import batch

pytestmark = pytest.mark.gpu

JOB = [("a_first", "<u8"), ("b_first", "<u8"), ("n", "<u8")]


def _resident(ctx, clouds, anchors, threshold):
    offs = np.cumsum([0] + [c.n for c in clouds]).astype(np.uint64)
    st = ctx.structs_create(offs, np.concatenate([c.xyz for c in clouds]), np.concatenate([c.cat for c in clouds]),
                            np.concatenate([c.tag for c in clouds]))
    a_struct = np.repeat(np.arange(len(clouds), dtype=np.uint32), len(anchors))
    env = ctx.envset_build(st, np.tile(np.asarray(anchors, np.uint32), len(clouds)), threshold, anchor_struct=a_struct)
    return st, env


def _oracle_scores(oracle, op, a, b, anchors_a, anchors_b, threshold):
    an = np.stack([anchors_a, anchors_b], axis=1).astype(np.uint32)
    return oracle.from_primitives(op, a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, an, threshold)


def _score_both_ways(ctx, monkeypatch, env_a, env_b, jobs, expect_tiles=True, **kw):
    before = ctx.tile_launches
    tiled = ctx.score_jobs_stats(env_a, env_b, jobs, scores=True, **kw)
    used = ctx.tile_launches - before
    assert expect_tiles is None or (used > 0) == expect_tiles, f"tile kernel launches: {used}"
    monkeypatch.setenv("LOCOHD_NO_TILES", "1")
    before = ctx.tile_launches
    plain = ctx.score_jobs_stats(env_a, env_b, jobs, scores=True, **kw)
    assert ctx.tile_launches == before
    monkeypatch.delenv("LOCOHD_NO_TILES")
    return tiled, plain


def test_ensemble_in_blocked_order_uses_the_tile_kernel(gpu_ctx, oracle_mod, monkeypatch):
    """11 members of 960 primitives (odd member count: the last block row and column are partial), member 7 an exact
    copy of member 2; all 55 structure pairs, every primitive an anchor."""
    base = synth.gen(21, 120, 8, 7)
    members = [synth.partner(base, 1.5, 500 + i) for i in range(11)]
    members[7] = members[2]
    anchors = np.arange(base.n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    pairs = batch.blocked_pairs(len(members), 4)
    assert len(pairs) == 55
    jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in pairs], dtype=JOB)
    tiled, plain = _score_both_ways(gpu_ctx, monkeypatch, env, env, jobs, job_means=True, anchor_means=True,
                                    anchor_stds=True)
    got = tiled["scores"].reshape(len(pairs), base.n)
    want = np.stack([_oracle_scores(oracle_mod, op, members[i], members[j], anchors, anchors, 10.0) for i, j in pairs])
    assert_scores_close(got.ravel(), want.ravel())
    # same tables, same expanded form; only the chunk boundaries (where D and R are rebuilt) differ
    assert np.abs(tiled["scores"] - plain["scores"]).max() <= 1e-12
    k = [tuple(p) for p in pairs.tolist()].index((2, 7))
    assert np.all(got[k] == 0.0), "identical members must score exactly 0"
    assert np.abs(tiled["job_means"] - want.mean(axis=1)).max() <= SCORE_TOL
    assert np.abs(tiled["anchor_means"] - want.mean(axis=0)).max() <= SCORE_TOL
    assert np.abs(tiled["anchor_stds"] - want.std(axis=0)).max() <= SCORE_TOL
    # a contiguous share of the list (what one rank of a multi-GPU run scores) starts and ends inside tiles
    sl = batch.contiguous_share(len(jobs), 1, 3)
    part, _ = _score_both_ways(gpu_ctx, monkeypatch, env, env, jobs[sl], expect_tiles=None)
    assert np.abs(part["scores"].reshape(-1, base.n) - got[sl]).max() <= 1e-12   # whichever kernel the share takes
    env.close(); st.close()


def test_job_orders_the_tiles_do_not_fit_take_the_pair_kernel(gpu_ctx, oracle_mod, monkeypatch):
    """Plain (i, j) order of few structures and one-against-many lists give tiles of one row: below the fill the tile
    kernel needs, so the one-pair-per-warp kernel scores them - same values either way."""
    base = synth.gen(22, 60, 8, 7)
    members = [synth.partner(base, 1.0, 700 + i) for i in range(6)]
    anchors = np.arange(base.n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    jobs = np.array([(0, j * base.n, base.n) for j in range(1, 6)], dtype=JOB)   # frame 0 against the others
    tiled, plain = _score_both_ways(gpu_ctx, monkeypatch, env, env, jobs, expect_tiles=False)
    assert np.array_equal(tiled["scores"], plain["scores"])
    want = np.stack([_oracle_scores(oracle_mod, op, members[0], members[j], anchors, anchors, 10.0) for j in range(1, 6)])
    assert_scores_close(tiled["scores"], want.ravel())
    env.close(); st.close()


def test_shuffled_job_list_and_repeated_jobs(gpu_ctx, oracle_mod, monkeypatch):
    """Lists whose order does not tile are regrouped on the host by the ranks of their environment runs (every cell of
    a tile carries its own output position): plain row-major (i, j) order (compare_ensembles.py:273-296 loops that
    way), a shuffled list and a list that names every job twice take the tile kernel too and are scored job by job."""
    rng = np.random.default_rng(5)
    base = synth.gen(23, 50, 8, 7)
    members = [synth.partner(base, 1.5, 900 + i) for i in range(8)]
    anchors = np.arange(base.n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    pairs = batch.blocked_pairs(8, 4)
    want = {tuple(p): _oracle_scores(oracle_mod, op, members[p[0]], members[p[1]], anchors, anchors, 10.0)
            for p in pairs.tolist()}
    row_major = np.lexsort((pairs[:, 1], pairs[:, 0]))
    assert np.array_equal(pairs[row_major], batch.all_pairs(8))
    for order in (row_major, rng.permutation(len(pairs)), np.r_[np.arange(len(pairs)), np.arange(len(pairs))][::-1]):
        sel = pairs[order]
        jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in sel], dtype=JOB)
        before = gpu_ctx.tile_launches
        got = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True)["scores"].reshape(len(sel), base.n)
        assert gpu_ctx.tile_launches > before
        assert_scores_close(got.ravel(), np.concatenate([want[tuple(p)] for p in sel.tolist()]))
    env.close(); st.close()


def test_tile_kernel_small_h2_branch_and_two_env_sets(gpu_ctx, oracle_mod, monkeypatch):
    """Rows from one environment set, columns from another (structures of different sizes): B holds every point of A
    twice, so the compositions are proportional after every complete triple of events and H^2 = 1 - D / sqrt(nA nB)
    is a rounding residue there - the difference-form branch has to take over (see
    test_proportional_compositions_small_h2_branch for the one-pair-per-warp kernel)."""
    rng = np.random.default_rng(77)
    C, n_points = 5, 280
    direction = rng.normal(size=(n_points, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    xyz_a = direction * (8.0 * rng.random(n_points) ** (1.0 / 3.0))[:, None]
    cat_a = rng.integers(0, C, n_points).astype(np.uint16)
    xyz_b = np.repeat(xyz_a, 2, axis=0)
    xyz_b[1::2, 0] += 1e-7
    cat_b = np.repeat(cat_a, 2)
    a = synth.Cloud(xyz_a, cat_a, np.arange(n_points, dtype=np.uint32), 1)
    b = synth.Cloud(xyz_b, cat_b, np.arange(2 * n_points, dtype=np.uint32) + n_points, 1)
    an_a = np.arange(0, n_points, 3, dtype=np.uint32)
    an_b = 2 * an_a
    op = set_both(gpu_ctx, oracle_mod, C, [("uniform", (0.0, 30.0))], tag_rule={"accept_same": False})
    st_a, env_a = _resident(gpu_ctx, [a] * 4, an_a, 20.0)
    st_b, env_b = _resident(gpu_ctx, [b] * 4, an_b, 20.0)
    n = len(an_a)
    jobs = np.array([(i * n, j * n, n) for i in range(4) for j in range(4)], dtype=JOB)
    tiled, plain = _score_both_ways(gpu_ctx, monkeypatch, env_a, env_b, jobs)
    want = _oracle_scores(oracle_mod, op, a, b, an_a, an_b, 20.0)
    got = tiled["scores"].reshape(16, n)
    assert_scores_close(got.ravel(), np.tile(want, 16))
    assert np.all(got == got[0]), "the 16 cells of the tile hold the same pair of structures"
    assert got.max() < 0.2
    assert np.abs(tiled["scores"] - plain["scores"]).max() <= 1e-12
    for h in (env_a, env_b, st_a, st_b):
        h.close()


def test_tile_kernel_reports_unknown_categories(gpu_ctx, oracle_mod):
    """pmf.rs:38-42: a primitive type the instance does not know is an error once it is met inside an environment."""
    from loco_hd_b200._capi import LocoHDError, UNKNOWN_CATEGORY
    base = synth.gen(24, 40, 8, 7)
    members = [synth.partner(base, 1.0, 40 + i) for i in range(8)]
    members[5].cat[17] = UNKNOWN_CATEGORY
    anchors = np.arange(base.n, dtype=np.uint32)
    set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in batch.blocked_pairs(8, 4)], dtype=JOB)
    before = gpu_ctx.tile_launches
    with pytest.raises(LocoHDError) as ei:
        gpu_ctx.score_jobs_stats(env, env, jobs, scores=True)
    assert ei.value.status == 3 and gpu_ctx.tile_launches > before
    env.close(); st.close()


def test_sliced_unit_order_scores_the_same_units(gpu_ctx, oracle_mod, monkeypatch):
    """The tile kernel visits all tiles for one slice of anchors before the next slice (environments of a slice stay in
    L2; on the 1000-structure ensemble the library picks the slice from the store size).  Forced here on a small
    ensemble: slices of 8, 24 and 200 anchors over 203 anchors (a short last slice of 3 / 11 anchors, runs that cross
    tile boundaries), against the tile-major order bit for bit - the same units, only in another order - and the
    oracle."""
    base = synth.gen(31, 29, 7, 7)
    assert base.n == 203
    members = [synth.partner(base, 1.5, 900 + i) for i in range(12)]   # 12: the 66 structure pairs fill their tiles well enough
    anchors = np.arange(base.n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    pairs = batch.blocked_pairs(len(members), 4)
    jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in pairs], dtype=JOB)
    monkeypatch.setenv("LOCOHD_TILE_SLICE", "0")
    before = gpu_ctx.tile_launches
    plain = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True, job_means=True)
    assert gpu_ctx.tile_launches > before
    want = np.stack([_oracle_scores(oracle_mod, op, members[i], members[j], anchors, anchors, 10.0) for i, j in pairs])
    assert_scores_close(plain["scores"], want.ravel())
    for s in ("8", "24", "200", "203", "1000"):
        monkeypatch.setenv("LOCOHD_TILE_SLICE", s)
        before = gpu_ctx.tile_launches
        sliced = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True, job_means=True)
        assert gpu_ctx.tile_launches > before
        assert np.array_equal(sliced["scores"], plain["scores"]), f"slice {s}"
        assert np.array_equal(sliced["job_means"], plain["job_means"])
    monkeypatch.delenv("LOCOHD_TILE_SLICE")
    monkeypatch.setenv("LOCOHD_TILE_SLICE_MB", "0.25")       # the automatic choice, with a budget this store exceeds
    auto = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True)
    assert np.array_equal(auto["scores"], plain["scores"])
    monkeypatch.delenv("LOCOHD_TILE_SLICE_MB")
    env.close(); st.close()
