"""GPU parity at the BASELINE.json shapes of configurations 3, 4 and 5 (SURVEY.md §8(d) instances, reduced only in
the NUMBER of models / frames / ensemble members - every structure has its full size and every anchor of the
configuration is scored), against the CPU oracle: per-anchor scores within 1e-9, neighbour lists / distances /
category counts bit-exact; the on-device reductions against numpy on the oracle's scores; and a case
that drives the expanded-form Hellinger of the fast kernel into its small-H^2 branch.

Reference call sites: casp14_extend_with_locohd.py:58-88 (config 3), trajectory_analyzer.py:89-129, 310 (config 4),
compare_ensembles.py:250-299 (config 5); the walk itself is src/locohd.rs:61-226.
"""
import numpy as np
import pytest

from benchdata import synth
from helpers import SCORE_TOL, assert_scores_close, check_from_primitives, set_both

pytestmark = pytest.mark.gpu

JOB = [("a_first", "<u8"), ("b_first", "<u8"), ("n", "<u8")]


def _resident(ctx, clouds, anchors, threshold):
    """Structures uploaded once, one environment per (structure, anchor): the layout bench.py uses."""
    offs = np.cumsum([0] + [c.n for c in clouds]).astype(np.uint64)
    st = ctx.structs_create(offs, np.concatenate([c.xyz for c in clouds]), np.concatenate([c.cat for c in clouds]),
                            np.concatenate([c.tag for c in clouds]))
    a_struct = np.repeat(np.arange(len(clouds), dtype=np.uint32), len(anchors))
    env = ctx.envset_build(st, np.tile(np.asarray(anchors, np.uint32), len(clouds)), threshold, anchor_struct=a_struct)
    return st, env


def _oracle_scores(oracle, op, a, b, anchors, threshold):
    an = np.stack([anchors, anchors], axis=1).astype(np.uint32)
    return oracle.from_primitives(op, a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, an, threshold)


def test_config3_reference_vs_models_full_size(gpu_ctx, oracle_mod):
    """Config 3: the 300-residue reference (2 700 primitives, all_atom_with_centroid, C = 8) against 8 models, all 300
    Cent anchors, uniform [3, 10], hetero contacts, threshold 10."""
    ref = synth.config3_reference()
    models = [synth.config3_model(ref, m) for m in range(8)]
    assert ref.n == 2700
    cent = ref.centroid_anchors()
    assert len(cent) == 300
    op = set_both(gpu_ctx, oracle_mod, 8, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, [ref] + models, cent, 10.0)
    jobs = np.array([(0, (m + 1) * len(cent), len(cent)) for m in range(len(models))], dtype=JOB)
    res = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True, job_means=True, anchor_means=True, anchor_stds=True)
    scores = res["scores"].reshape(len(models), len(cent))
    want = np.stack([_oracle_scores(oracle_mod, op, ref, mdl, cent, 10.0) for mdl in models])
    assert_scores_close(scores.ravel(), want.ravel())
    assert np.abs(res["job_means"] - want.mean(axis=1)).max() <= SCORE_TOL        # casp14_extend_with_locohd.py:88
    assert np.abs(res["anchor_means"] - want.mean(axis=0)).max() <= SCORE_TOL
    assert np.abs(res["anchor_stds"] - want.std(axis=0)).max() <= SCORE_TOL
    env.close(); st.close()
    # environments bit-exact (neighbour lists, distances, counts) on two of the pairs through the one-call entry point
    an = np.stack([cent, cent], axis=1)
    for mdl in models[:2]:
        check_from_primitives(gpu_ctx, oracle_mod, op, (ref.xyz, ref.cat, ref.tag), (mdl.xyz, mdl.cat, mdl.tag), an, 10.0)


@pytest.mark.parametrize("variant", ["cent_anchors", "all_anchors"])
def test_config4_frames_vs_frame0_full_size(gpu_ctx, oracle_mod, variant):
    """Config 4: frame 0 (5 000 primitives, coarse_grained_with_centroid, C = 8) against perturbed frames
    (delta = 0.5 A: near-identical compositions, where an expanded-form Hellinger is weakest), the 1 250 Cent anchors
    of trajectory_analyzer.py:104 and the 5 000-anchor stress variant; frame 0 against itself scores exactly 0."""
    f0 = synth.config4_frame0()
    assert f0.n == 5000
    n_frames = 8 if variant == "cent_anchors" else 3
    frames = [synth.config4_frame(f0, t) for t in range(0, n_frames + 1)]      # frames[0] is frame 0 itself
    anchors = f0.centroid_anchors() if variant == "cent_anchors" else np.arange(f0.n, dtype=np.uint32)
    assert len(anchors) == (1250 if variant == "cent_anchors" else 5000)
    op = set_both(gpu_ctx, oracle_mod, 8, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, frames, anchors, 10.0)
    n = len(anchors)
    jobs = np.array([(0, t * n, n) for t in range(len(frames))], dtype=JOB)
    res = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True, anchor_stds=True, anchor_means=True)
    scores = res["scores"].reshape(len(frames), n)
    assert np.all(scores[0] == 0.0)                                              # frame 0 against itself
    want = np.stack([_oracle_scores(oracle_mod, op, f0, fr, anchors, 10.0) for fr in frames])
    assert_scores_close(scores.ravel(), want.ravel())
    # np.std(all_points[1:], axis=0) of trajectory_analyzer.py:310 is taken over the frames after the first
    res1 = gpu_ctx.score_jobs_stats(env, env, jobs[1:], anchor_stds=True, anchor_means=True)
    assert np.abs(res1["anchor_stds"] - want[1:].std(axis=0)).max() <= SCORE_TOL
    assert np.abs(res1["anchor_means"] - want[1:].mean(axis=0)).max() <= SCORE_TOL
    env.close(); st.close()
    if variant == "cent_anchors":
        an = np.stack([anchors, anchors], axis=1)
        check_from_primitives(gpu_ctx, oracle_mod, op, (f0.xyz, f0.cat, f0.tag),
                              (frames[1].xyz, frames[1].cat, frames[1].tag), an, 10.0)


def test_config5_ensemble_members_full_size(gpu_ctx, oracle_mod):
    """Config 5: 6 full members (5 000 primitives, all_atom, C = 7) of the ensemble, all 15 structure pairs, all
    5 000 anchors per pair = 75 000 per-anchor scores (not only the means), plus the reductions compare_ensembles.py
    computes from them (:293 per-pair mean, :299 per-atom mean over the pairs)."""
    base = synth.config5_base()
    members = [synth.config5_member(base, i) for i in range(6)]
    assert base.n == 5000
    anchors = np.arange(base.n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    st, env = _resident(gpu_ctx, members, anchors, 10.0)
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in pairs], dtype=JOB)
    res = gpu_ctx.score_jobs_stats(env, env, jobs, scores=True, job_means=True, anchor_means=True, anchor_stds=True)
    scores = res["scores"].reshape(len(pairs), base.n)
    want = np.stack([_oracle_scores(oracle_mod, op, members[i], members[j], anchors, 10.0) for i, j in pairs])
    assert_scores_close(scores.ravel(), want.ravel())
    assert np.abs(res["job_means"] - want.mean(axis=1)).max() <= SCORE_TOL
    assert np.abs(res["anchor_means"] - want.mean(axis=0)).max() <= SCORE_TOL
    assert np.abs(res["anchor_stds"] - want.std(axis=0)).max() <= SCORE_TOL
    # means only (what the benchmark copies out): identical values without the scores leaving the device
    only = gpu_ctx.score_jobs_stats(env, env, jobs, job_means=True)
    assert set(only) == {"job_means"} and np.array_equal(only["job_means"], res["job_means"])
    env.close(); st.close()
    an = np.stack([anchors[::7], anchors[::7]], axis=1)
    check_from_primitives(gpu_ctx, oracle_mod, op, (members[0].xyz, members[0].cat, members[0].tag),
                          (members[1].xyz, members[1].cat, members[1].tag), an, 10.0)


@pytest.mark.parametrize("n_points", [600, 2600])
def test_proportional_compositions_small_h2_branch(gpu_ctx, oracle_mod, n_points):
    """Structure B holds every point of A twice (the twin 1e-7 A away): walking outwards, the composition of B is
    exactly twice A's after every complete triple of events, so H^2 = 1 - D / sqrt(nA nB) is a rounding residue there
    and the fast kernel has to take its difference-form branch (returning ~0, not sqrt(1e-16) = 1e-8); in between the
    compositions differ by one count in ~1e3 (small, not tiny, H^2).  600 points: both environments fit the fast
    kernel's stage (1 800 members per pair); 2 600 points: environments of 2 600 / 5 200 members (> 2 000) take the
    staged generic kernel.  The weight is uniform over [0, 30] so that every event carries weight."""
    rng = np.random.default_rng(4242 + n_points)
    C = 5
    direction = rng.normal(size=(n_points, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    xyz_a = direction * (8.0 * rng.random(n_points) ** (1.0 / 3.0))[:, None]
    cat_a = rng.integers(0, C, n_points).astype(np.uint16)
    tag_a = np.arange(n_points, dtype=np.uint32)
    xyz_b = np.repeat(xyz_a, 2, axis=0)
    xyz_b[1::2, 0] += 1e-7
    cat_b = np.repeat(cat_a, 2)
    tag_b = np.arange(2 * n_points, dtype=np.uint32) + n_points
    anchors = np.stack([np.arange(0, n_points, 4), 2 * np.arange(0, n_points, 4)], axis=1).astype(np.uint32)
    op = set_both(gpu_ctx, oracle_mod, C, [("uniform", (0.0, 30.0))], tag_rule={"accept_same": False})
    got, ref = check_from_primitives(gpu_ctx, oracle_mod, op, (xyz_a, cat_a, tag_a), (xyz_b, cat_b, tag_b), anchors, 20.0,
                                     check_envs=(n_points == 600))
    assert ref["env_sizes"][:, 0].min() == n_points and ref["env_sizes"][:, 1].min() == 2 * n_points
    # the scores are tiny (the compositions agree up to one count): a broken small-H^2 branch would add ~3e-9
    assert got.max() < 0.2


# ------------------------------------------------------------------------------------- SURVEY §8(f) N1 / N3 / N4
def _as_atoms(lchd_mod, cloud, prefix="T"):
    return [lchd_mod.PrimitiveAtom(f"{prefix}{c}", f"A/{t}-RES", x) for x, c, t in zip(cloud.xyz.tolist(), cloud.cat, cloud.tag)]


def test_public_class_resident_batch_equals_per_call_api(oracle_mod):
    """LoCoHD.structures / environments / score_batch (N1 on the public class): one reference against several
    models, uploaded once, equals the reference API's per-pair from_primitives loop (casp14_extend_with_locohd.py:58-79)
    bit for bit; the device reductions equal numpy on those scores."""
    import loco_hd

    ref = synth.gen(5, 80, 9, 8, with_centroid=True)
    models = [synth.config3_model(ref, m) for m in range(4)]
    types = [f"T{i}" for i in range(8)]
    lchd = loco_hd.LoCoHD(types, loco_hd.WeightFunction("uniform", [3.0, 10.0]), loco_hd.TagPairingRule({"accept_same": False}))
    cent = ref.centroid_anchors()
    pa = _as_atoms(loco_hd, ref)
    per_call = np.array([lchd.from_primitives(pa, _as_atoms(loco_hd, m), [(int(i), int(i)) for i in cent], 10.0) for m in models])
    clouds = [ref] + models
    offs = np.cumsum([0] + [c.n for c in clouds]).astype(np.uint64)
    cats = lchd.category_ids([f"T{c}" for c in np.concatenate([c.cat for c in clouds])])
    tags = lchd.intern_tags([f"A/{t}-RES" for t in np.concatenate([c.tag for c in clouds])])
    st = lchd.structures(offs, np.concatenate([c.xyz for c in clouds]), cats, tags)
    assert st.n_structures == 5 and st.n_primitives == int(offs[-1])
    env = lchd.environments(st, np.tile(cent, len(clouds)), 10.0,
                            anchor_struct=np.repeat(np.arange(len(clouds), dtype=np.uint32), len(cent)))
    assert len(env) == len(clouds) * len(cent)
    jobs = np.array([(0, (m + 1) * len(cent), len(cent)) for m in range(len(models))], dtype=np.uint64)
    flat = lchd.score_batch(env, env, jobs)
    assert np.array_equal(flat.reshape(per_call.shape), per_call)
    red = lchd.score_batch(env, env, jobs, reduce=["scores", "job_mean", "anchor_mean", "anchor_std"])
    assert np.array_equal(red["scores"], flat)
    assert np.abs(red["job_mean"] - per_call.mean(axis=1)).max() <= 1e-14
    assert np.abs(red["anchor_mean"] - per_call.mean(axis=0)).max() <= 1e-14
    assert np.abs(red["anchor_std"] - per_call.std(axis=0)).max() <= 1e-14
    assert set(lchd.score_batch(env, env, jobs, reduce="job_mean")) == {"job_mean"}
    # trajectory use: new coordinates for the same topology (float32 travels as float32)
    moved = np.concatenate([c.xyz for c in clouds]).astype(np.float32)
    st.update_xyz(moved)
    env2 = lchd.environments(st, np.tile(cent, len(clouds)), 10.0,
                             anchor_struct=np.repeat(np.arange(len(clouds), dtype=np.uint32), len(cent)))
    again = lchd.score_batch(env2, env2, jobs)
    st64 = lchd.structures(offs, moved.astype(np.float64), cats, tags)
    env3 = lchd.environments(st64, np.tile(cent, len(clouds)), 10.0,
                             anchor_struct=np.repeat(np.arange(len(clouds), dtype=np.uint32), len(cent)))
    assert np.array_equal(again, lchd.score_batch(env3, env3, jobs))
    with pytest.raises(ValueError):
        lchd.score_batch(env, env, np.array([[0, 0, 10 ** 9]], dtype=np.uint64))
    with pytest.raises(ValueError):
        lchd.score_batch(env, env, jobs, reduce="median")
    for obj in (env, env2, env3, st, st64):
        obj.close()


def test_ensemble_callers_dmx_form_equals_the_cutoff_form(gpu_ctx, oracle_mod):
    """compare_ensembles.py:250-296 calls from_dmxs on full distance matrices with +inf at homo-residue entries; the
    benchmark states the same computation as the cutoff form (threshold 10, hetero contacts).  Both forms on the GPU,
    against each other and against the oracle (uniform [3, 10] is constant from the cutoff on)."""
    base = synth.gen(9, 60, 8, 7)
    a, b = synth.config5_member(base, 0), synth.config5_member(base, 1)
    op = set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    anchors = np.stack([np.arange(a.n)] * 2, axis=1).astype(np.uint32)
    cut = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    want = oracle_mod.from_primitives(op, a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    assert_scores_close(cut, want)

    def dmx(c):
        d = c.xyz[None, :, :] - c.xyz[:, None, :]
        d = np.sqrt(np.sum(d ** 2, axis=2))
        same = c.tag[:, None] == c.tag[None, :]
        np.fill_diagonal(same, False)
        d[same] = np.inf
        return d

    ea, eb = gpu_ctx.envset_from_rows(dmx(a), a.cat), gpu_ctx.envset_from_rows(dmx(b), b.cat)
    full = gpu_ctx.score_pairs(ea, eb, anchors)
    ea.close(); eb.close()
    assert_scores_close(full, want)
    assert np.abs(full - cut).max() <= 1e-12


def test_models_against_reference_helper(oracle_mod):
    """batch.models_against_reference (the loop of casp14_extend_with_locohd.py:44-88 as one resident batch): models that
    lack residues or list them in another order, anchors paired by tag; per-residue scores against the per-model
    from_arrays calls and against the oracle, the per-model mean against numpy."""
    import loco_hd
    from loco_hd_b200 import batch

    ref = synth.gen(5, 120, 9, 8, with_centroid=True)
    lchd = loco_hd.LoCoHD([f"T{i}" for i in range(8)], loco_hd.WeightFunction("uniform", [3.0, 10.0]),
                          loco_hd.TagPairingRule({"accept_same": False}))
    op = oracle_mod.Params(8, [("uniform", [3.0, 10.0])], tag_rule={"accept_same": False})
    rng = np.random.default_rng(17)
    models = []
    for m in range(5):
        full = synth.config3_model(ref, m)
        res = np.arange(120)
        if m == 1:
            res = res[rng.random(120) < 0.85]                 # missing residues
        if m == 2:
            res = rng.permutation(res)                        # another residue order
        if m == 3:
            res = res[:0]                                     # nothing in common: no anchors
        idx = (res[:, None] * 9 + np.arange(9)[None, :]).ravel()
        tag = full.tag[idx] if m != 3 else full.tag[:18] + 1000
        idx = idx if m != 3 else np.arange(18)
        models.append((full.xyz[idx], full.cat[idx], tag.astype(np.uint32)))
    cent = int(ref.centroid_cat)
    out = batch.models_against_reference(lchd, (ref.xyz, ref.cat, ref.tag), models, cent, 10.0)
    assert len(out) == 5
    for m, (pairs, scores, mean) in enumerate(out):
        mx, mc, mt = models[m]
        if m == 3:
            assert len(pairs) == 0 and len(scores) == 0 and np.isnan(mean)
            continue
        assert len(pairs) == (120 if m != 1 else len(np.unique(mt)))
        assert np.all(ref.cat[pairs[:, 0]] == cent) and np.all(mc[pairs[:, 1]] == cent)
        assert np.array_equal(ref.tag[pairs[:, 0]], mt[pairs[:, 1]])
        per_call = lchd.from_arrays(ref.xyz, ref.cat, ref.tag, mx, mc, mt, pairs, 10.0)
        assert np.abs(scores - per_call).max() <= 1e-12
        want = oracle_mod.from_primitives(op, ref.xyz, ref.cat, ref.tag, mx, mc, mt, pairs, 10.0)
        assert_scores_close(scores, want)
        assert abs(mean - want.mean()) <= SCORE_TOL


def test_from_arrays_with_tag_pair_list_rule():
    """from_arrays accepts a WithList rule when the integer tags come from intern_tags (N1)."""
    import loco_hd

    rng = np.random.default_rng(5)
    n = 260
    xyz_a, xyz_b = rng.uniform(-9, 9, (n, 3)), rng.uniform(-9, 9, (n, 3))
    types = ["A", "B", "C", "D"]
    ta, tb = rng.choice(types, n), rng.choice(types, n)
    names = ["r1", "r2", "r3", "r4"]
    ga, gb = rng.choice(names, n), rng.choice(names, n)
    rule = loco_hd.TagPairingRule({"tag_pairs": {("r1", "r2"), ("r3", "r3"), ("r4", "r1")}, "accepted_pairs": True, "ordered": False})
    lchd = loco_hd.LoCoHD(types, loco_hd.WeightFunction("uniform", [2.0, 9.0]), rule)
    pa = [loco_hd.PrimitiveAtom(t, g, x) for t, g, x in zip(ta, ga, xyz_a.tolist())]
    pb = [loco_hd.PrimitiveAtom(t, g, x) for t, g, x in zip(tb, gb, xyz_b.tolist())]
    anchors = [(i, i) for i in range(0, n, 2)]
    want = np.array(lchd.from_primitives(pa, pb, anchors, 9.0))
    got = lchd.from_arrays(xyz_a, lchd.category_ids(ta), lchd.intern_tags(ga), xyz_b, lchd.category_ids(tb),
                           lchd.intern_tags(gb), np.array(anchors, np.uint32), 9.0)
    assert np.array_equal(got, want) and want.std() > 0


def test_two_threads_share_one_instance():
    """Two Python threads calling the same LoCoHD instance (the reference allows it: its methods take &self and hold
    the GIL) must both finish: the host module never waits for its device mutex while holding the GIL."""
    import threading

    import loco_hd

    a, b = synth.config1()
    types = [f"T{i}" for i in range(7)]
    lchd = loco_hd.LoCoHD(types, loco_hd.WeightFunction("uniform", [3.0, 10.0]), loco_hd.TagPairingRule({"accept_same": False}))
    pa, pb = _as_atoms(loco_hd, a), _as_atoms(loco_hd, b)
    anchors = [(i, i) for i in range(0, a.n, a.k)]
    want = lchd.from_primitives(pa, pb, anchors, 10.0)
    results, errors = {}, []

    def work(k):
        try:
            for _ in range(40):
                results[k] = lchd.from_primitives(pa, pb, anchors, 10.0)
                lchd.from_anchors(["T0", "T1"], ["T0", "T2"], [0.0, 4.0], [0.0, 5.0])
        except Exception as exc:   # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "deadlock: a thread is still waiting"
    assert not errors, errors
    assert results[0] == want and results[1] == want


def test_device_centroids_match_the_python_assigner(gpu_ctx, oracle_mod):
    """N3: primitive assignment for a compiled topology on the device (locohd_structs_update_from_atoms) gives the
    coordinates PrimitiveAssigner.assign_primitive_structure computes per frame (float32 means,
    atom_converter_utils.py:92-131): the scores of frame 0 against every frame are bit-identical to scoring the
    Python-assigned primitives, i.e. trajectory_analyzer.py:89-129 without the per-frame regex pass."""
    from loco_hd_b200.atom_converter_utils import TYPING_DIR, PrimitiveAssigner
    from test_primitive_assigner import SIDE, make_structure

    assigner = PrimitiveAssigner(TYPING_DIR / "coarse_grained_with_centroid.config.json")
    rng = np.random.default_rng(12)
    names = tuple(rng.choice(sorted(SIDE), 90))
    structure = make_structure(7, names)
    for k, res in enumerate(structure.get_residues()):   # pack the residues closer than make_structure's 40 A box
        shift = np.float32(0.6) * np.array([k % 5, (k // 5) % 5, k // 25], dtype=np.float32) * 7 - res._atoms[0].coord * np.float32(0.8)
        for atom in res._atoms:
            atom.coord = (atom.coord + shift).astype(np.float32)
    topo = assigner.compile_topology(structure)
    types = assigner.all_primitive_types
    cat = np.array([types.index(t) for t in topo.primitive_types], dtype=np.uint16)
    tag = np.array([s.source_residue[3][1] for s in topo.sources], dtype=np.uint32)
    atoms0 = np.array([a.coord for r in structure.get_residues() for a in r.get_atoms()], dtype=np.float32)
    n_frames = 5
    frames = np.stack([atoms0] + [atoms0 + rng.normal(0, 0.4, atoms0.shape).astype(np.float32) for _ in range(n_frames - 1)])
    n_prims = len(cat)
    # Python path: move the atoms, run the assigner, widen to float64
    py_xyz = []
    for f in range(n_frames):
        it = iter(frames[f])
        for res in structure.get_residues():
            for atom in res._atoms:
                atom.coord = next(it)
        templates = assigner.assign_primitive_structure(structure)
        assert [t.primitive_type for t in templates] == topo.primitive_types
        py_xyz.append(np.array([t.coordinates for t in templates]))
    assert py_xyz[0].dtype == np.float32
    py_xyz = np.stack(py_xyz).astype(np.float64)
    # device path: structures created with placeholder coordinates, then filled from the atoms in one call
    op = set_both(gpu_ctx, oracle_mod, len(types), [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    offs = (np.arange(n_frames + 1) * n_prims).astype(np.uint64)
    st = gpu_ctx.structs_create(offs, np.zeros((n_frames * n_prims, 3)), np.tile(cat, n_frames), np.tile(tag, n_frames))
    st.update_from_atoms(frames, topo.segment_start, topo.atom_index, first_struct=0, n_frames=n_frames)
    anchors = np.flatnonzero(cat == types.index("Cent")).astype(np.uint32)
    a_struct = np.repeat(np.arange(n_frames, dtype=np.uint32), len(anchors))
    env = gpu_ctx.envset_build(st, np.tile(anchors, n_frames), 10.0, anchor_struct=a_struct)
    jobs = np.array([(0, f * len(anchors), len(anchors)) for f in range(n_frames)], dtype=JOB)
    dev = gpu_ctx.score_jobs(env, env, jobs).reshape(n_frames, len(anchors))
    env.close()
    # one frame at a time as well (what a streaming caller does)
    st.update_from_atoms(frames[2], topo.segment_start, topo.atom_index, first_struct=1, n_frames=1)
    env = gpu_ctx.envset_build(st, np.tile(anchors, n_frames), 10.0, anchor_struct=a_struct)
    dev2 = gpu_ctx.score_jobs(env, env, jobs).reshape(n_frames, len(anchors))
    env.close(); st.close()
    an = np.stack([anchors, anchors], axis=1)
    for f in range(n_frames):
        host = gpu_ctx.from_primitives(py_xyz[0], cat, tag, py_xyz[f], cat, tag, an, 10.0)
        assert np.array_equal(dev[f], host), f"frame {f}"
        ref = oracle_mod.from_primitives(op, py_xyz[0], cat, tag, py_xyz[f], cat, tag, an, 10.0)
        assert np.abs(dev[f] - ref).max() <= SCORE_TOL
    assert np.array_equal(dev2[1], dev[2]) and np.array_equal(dev2[3], dev[3])
    assert np.all(dev[0] == 0.0) and dev[1:].max() > 0.0
    with pytest.raises(ValueError):   # a topology with another primitive count
        st2 = gpu_ctx.structs_create(offs, np.zeros((n_frames * n_prims, 3)), np.tile(cat, n_frames), np.tile(tag, n_frames))
        st2.update_from_atoms(frames[0], topo.segment_start[:-1], topo.atom_index, first_struct=0, n_frames=1)


def test_f32_wire_format_is_exact(gpu_ctx, oracle_mod):
    """locohd_structs_create_f32 / update_xyz_f32: float32 coordinates uploaded as float32 give the scores of the
    same values passed as float64 (the widening on the device is exact)."""
    a = synth.gen(61, 180, 8, 7, f32_exact=True)
    b = synth.partner(a, 1.0, 62, f32_exact=True)
    set_both(gpu_ctx, oracle_mod, 7, [("kumaraswamy", (3.0, 10.0, 2.0, 5.0))], tag_rule={"accept_same": False})
    offs = np.array([0, a.n, a.n + b.n], dtype=np.uint64)
    xyz = np.concatenate([a.xyz, b.xyz])
    cat, tag = np.concatenate([a.cat, b.cat]), np.concatenate([a.tag, b.tag])
    anchors = np.arange(a.n, dtype=np.uint32)
    a_struct = np.repeat(np.arange(2, dtype=np.uint32), a.n)
    jobs = np.array([(0, a.n, a.n)], dtype=JOB)
    out = []
    for arr in (xyz, xyz.astype(np.float32)):
        st = gpu_ctx.structs_create(offs, arr, cat, tag)
        env = gpu_ctx.envset_build(st, np.tile(anchors, 2), 10.0, anchor_struct=a_struct)
        out.append(gpu_ctx.score_jobs(env, env, jobs))
        env.close()
        st.update_xyz(arr)     # the same values again through the update entry point
        env = gpu_ctx.envset_build(st, np.tile(anchors, 2), 10.0, anchor_struct=a_struct)
        out.append(gpu_ctx.score_jobs(env, env, jobs))
        env.close(); st.close()
    assert all(np.array_equal(out[0], o) for o in out[1:]) and out[0].std() > 0


def test_device_resident_wf_indices_are_validated(gpu_ctx, oracle_mod):
    """Per-pair weight-function indices that already live on the device are range-checked by a kernel."""
    import torch

    from loco_hd_b200 import _capi

    a, b = synth.config1()
    set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0)), ("uniform", (0.0, 8.0))], tag_rule={"accept_same": False})
    anchors = np.stack([np.arange(0, a.n, a.k, dtype=np.uint32)] * 2, axis=1)
    good = torch.zeros(len(anchors), dtype=torch.int32, device="cuda")
    good[::2] = 1
    ok = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0, wf_idx=good.data_ptr())
    host = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0, wf_idx=good.cpu().numpy().astype(np.uint32))
    assert np.array_equal(ok, host)
    bad = good.clone()
    bad[3] = 7
    with pytest.raises(_capi.LocoHDError) as e:
        gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0, wf_idx=bad.data_ptr())
    assert e.value.status == 10
    # the context stays usable
    again = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0, wf_idx=good.data_ptr())
    assert np.array_equal(again, ok)


def test_ragged_distance_rows(gpu_ctx, oracle_mod):
    """from_dmxs with rows of different lengths, as the reference's Vec<Vec<f64>> allows: row r is co-sorted with the
    first len_r categories (utils.rs:25-39).  Expected values: the oracle's from_anchors on every co-sorted row pair."""
    import loco_hd

    rng = np.random.default_rng(99)
    types = ["A", "B", "C", "D"]
    n_rows = 40
    seq_a, seq_b = rng.choice(types, 70), rng.choice(types, 64)
    rows_a, rows_b = [], []
    for r in range(n_rows):
        la, lb = int(rng.integers(1, 71)), int(rng.integers(1, 65))
        ra, rb = rng.uniform(0.5, 14.0, la), rng.uniform(0.5, 14.0, lb)
        za = int(rng.integers(la))
        ra[za] = 0.0
        rb[rng.integers(lb)] = 0.0
        if r % 7 == 0 and la > 2:
            ra[(za + 1) % la] = np.inf        # banned contacts of compare_ensembles.py:261-263
        rows_a.append(ra.tolist()); rows_b.append(rb.tolist())
    wf, sd = ("uniform", [3.0, 10.0]), ("Hellinger", [2.0])
    lchd = loco_hd.LoCoHD(types, loco_hd.WeightFunction(*wf))
    got = np.array(lchd.from_dmxs(seq_a, seq_b, rows_a, rows_b))
    op = set_both(gpu_ctx, oracle_mod, 4, [(wf[0], tuple(wf[1]))], statistical_distance=(sd[0], tuple(sd[1])))
    cat = {t: i for i, t in enumerate(types)}
    ca, cb = np.array([cat[t] for t in seq_a]), np.array([cat[t] for t in seq_b])
    want = []
    for ra, rb in zip(rows_a, rows_b):
        oa, ob = np.argsort(ra, kind="stable"), np.argsort(rb, kind="stable")
        want.append(oracle_mod.from_anchors(op, ca[:len(ra)][oa], cb[:len(rb)][ob], np.array(ra)[oa], np.array(rb)[ob]))
    assert_scores_close(got, np.array(want))
    # the C ABI directly
    ea = gpu_ctx.envset_from_ragged_rows(rows_a, ca)
    eb = gpu_ctx.envset_from_ragged_rows(rows_b, cb)
    jobs = np.array([(0, 0, n_rows)], dtype=JOB)
    assert np.array_equal(gpu_ctx.score_jobs(ea, eb, jobs), got)
    ea.close(); eb.close()
    with pytest.raises(ValueError):   # a row longer than the category sequence (the reference panics)
        lchd.from_dmxs(seq_a[:5], seq_b, rows_a, rows_b)
    with pytest.raises(ValueError):
        lchd.from_dmxs(seq_a, seq_b, rows_a + [[]], rows_b + [[0.0]])
    with pytest.raises(ValueError):   # a row without a zero distance (locohd.rs:74-77)
        lchd.from_dmxs(seq_a, seq_b, [[1.0, 2.0]] + rows_a[1:], rows_b)


def test_store_sizing_history_recovers_from_a_denser_call(gpu_ctx, oracle_mod):
    """Calls of the same shape reuse the store size of the previous call (no sizing pass); when the same number of
    anchors suddenly has much larger environments the store overflows, the history is dropped and the call is redone
    with a sampled size - the environments must come out complete either way."""
    rng = np.random.default_rng(31)
    n = 40000                                     # above the 32 768-anchor bound below which no sizing is needed
    cat = rng.integers(0, 5, n).astype(np.uint16)
    tag = np.arange(n, dtype=np.uint32)
    anchors = np.arange(n, dtype=np.uint32)
    op = set_both(gpu_ctx, oracle_mod, 5, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False})
    sizes = {}
    for label, extent in (("sparse", 120.0), ("sparse again", 120.0), ("dense", 42.0), ("dense again", 42.0), ("sparse once more", 120.0)):
        xyz = np.random.default_rng(7 if extent > 100 else 8).uniform(-extent, extent, (n, 3))
        st = gpu_ctx.structure(xyz, cat, tag)
        env = gpu_ctx.envset_build(st, anchors, 10.0)
        off = np.empty(n + 1, np.uint64)
        from loco_hd_b200 import _capi
        gpu_ctx._check(gpu_ctx.lib.locohd_envset_dump(gpu_ctx.h, env.h, _capi._p(off), None, None, None))
        got = np.diff(off).astype(np.int64)
        env.close(); st.close()
        for a in (0, 1234, 39999):
            idx, _, _ = oracle_mod.environment(op, xyz, cat, tag, a, 10.0)
            assert got[a] == len(idx), (label, a)
        sizes[label] = got
    assert np.array_equal(sizes["sparse"], sizes["sparse again"]) and np.array_equal(sizes["sparse"], sizes["sparse once more"])
    assert np.array_equal(sizes["dense"], sizes["dense again"])
    assert sizes["dense"].mean() > 10 * sizes["sparse"].mean()


def test_trajectory_helper_from_atoms_to_scores(oracle_mod):
    """batch.trajectory_scores: frame 0 against blocks of frames, atom coordinates in, scores out (N1 + N3 + N4 on the
    public class) - equal to the reference API's loop of assign + from_primitives per frame
    (trajectory_analyzer.py:89-129), bit for bit."""
    import loco_hd
    from loco_hd_b200 import batch
    from loco_hd_b200.atom_converter_utils import TYPING_DIR, PrimitiveAssigner
    from test_primitive_assigner import SIDE, make_structure

    assigner = PrimitiveAssigner(TYPING_DIR / "coarse_grained_with_centroid.config.json")
    rng = np.random.default_rng(21)
    structure = make_structure(9, tuple(rng.choice(sorted(SIDE), 60)))
    for k, res in enumerate(structure.get_residues()):
        centre = np.mean([a.coord for a in res._atoms], axis=0)
        target = np.array([5.0 * (k % 4), 5.0 * ((k // 4) % 4), 5.0 * (k // 16)], dtype=np.float32)
        for atom in res._atoms:
            atom.coord = (atom.coord - centre + target).astype(np.float32)
    topo = assigner.compile_topology(structure)
    types = assigner.all_primitive_types
    lchd = loco_hd.LoCoHD(types, loco_hd.WeightFunction("uniform", [3.0, 10.0]), loco_hd.TagPairingRule({"accept_same": False}))
    cats = lchd.category_ids(topo.primitive_types)
    tag_names = [f"{s.source_residue[2]}/{s.source_residue[3][1]}-{s.source_residue_name}" for s in topo.sources]
    tags = lchd.intern_tags(tag_names)
    atoms0 = np.array([a.coord for r in structure.get_residues() for a in r.get_atoms()], dtype=np.float32)
    frames = np.stack([atoms0 + rng.normal(0, 0.5, atoms0.shape).astype(np.float32) for _ in range(7)])
    anchors = np.flatnonzero(np.array(topo.primitive_types) == "Cent").astype(np.uint32)
    got = np.concatenate(list(batch.trajectory_scores(lchd, topo, cats, tags, atoms0, [frames[:3], frames[3:6], frames[6:]],
                                                      anchors, 10.0)))
    assert got.shape == (7, len(anchors))

    def primitive_atoms(atom_xyz):      # the reference's way: move the atoms, run the assigner, build PrimitiveAtoms
        it = iter(atom_xyz)
        for res in structure.get_residues():
            for atom in res._atoms:
                atom.coord = next(it)
        return [loco_hd.PrimitiveAtom(t.primitive_type, tag, [float(v) for v in t.coordinates])
                for t, tag in zip(assigner.assign_primitive_structure(structure), tag_names)]

    pa0 = primitive_atoms(atoms0)
    pairs = [(int(a), int(a)) for a in anchors]
    for f in range(7):
        want = np.array(lchd.from_primitives(pa0, primitive_atoms(frames[f]), pairs, 10.0))
        assert np.array_equal(got[f], want), f"frame {f}"
    stats = list(batch.trajectory_scores(lchd, topo, cats, tags, atoms0, [frames], anchors, 10.0,
                                         reduce=["anchor_mean", "anchor_std"]))
    assert len(stats) == 1 and np.abs(stats[0]["anchor_std"] - got.std(axis=0)).max() <= 1e-14
    assert np.abs(stats[0]["anchor_mean"] - got.mean(axis=0)).max() <= 1e-14
    assert list(batch.trajectory_scores(lchd, topo, cats, tags, atoms0, [], anchors, 10.0)) == []


def test_category_weights_on_the_fast_path(gpu_ctx, oracle_mod):
    """Hellinger-2 with category weights (pisces_random_pairs.py:32-41: 1 / abundance per primitive type) runs in the
    TMA-staged kernel too (weighted D and totals); config-2 shape, 10 000 anchors, and the identical-structure case."""
    a, b = synth.config2()
    w = [1 / 30.61875, 1 / 1.06906, 1 / 19.46292, 1 / 5.2, 1 / 8.9, 1 / 2.7, 1 / 12.4]
    op = set_both(gpu_ctx, oracle_mod, 7, [("kumaraswamy", (3.0, 10.0, 2.0, 5.0))], category_weights=w,
                  tag_rule={"accept_same": False})
    anchors = np.stack([np.arange(a.n, dtype=np.uint32)] * 2, axis=1)
    got = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    ref = oracle_mod.from_primitives(op, a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    assert_scores_close(got, ref)
    unit = set_both(gpu_ctx, oracle_mod, 7, [("kumaraswamy", (3.0, 10.0, 2.0, 5.0))], tag_rule={"accept_same": False})
    assert np.abs(got - gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)).max() > 1e-3
    set_both(gpu_ctx, oracle_mod, 7, [("uniform", (3.0, 10.0))], category_weights=w, tag_rule={"accept_same": False})
    same = gpu_ctx.from_primitives(a.xyz, a.cat, a.tag, a.xyz, a.cat, a.tag, anchors[:2000], 10.0)
    assert np.all(same == 0.0)
