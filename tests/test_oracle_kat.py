"""CPU: the oracle (oracle/locohd_oracle.cpp) against every known answer the reference's own tests hold for the
scoring path: /root/reference/tests/test_locohd.py:27-52, tests/test_tag_pairing_rule.py:8-157,
tests/test_wfs.py:8-138, and against an independent pure-Python twin (oracle/py_twin.py)."""
import numpy as np
import pytest


def test_small_locohd(oracle_mod):
    P = oracle_mod.Params
    p = P(4, [("uniform", [0., 4.])])
    seq = [0, 1, 2, 3]
    assert oracle_mod.from_anchors(p, seq, seq, [0., 1., 2., 3.], [0., 1., 1., 1.]) == pytest.approx(0.2268, abs=5e-5)
    assert oracle_mod.from_anchors(p, seq, seq, [0., 1., 1., 1.], [0., 1., 2., 3.]) == pytest.approx(0.2268, abs=5e-5)
    p = P(3, [("kumaraswamy", [3., 10., 2., 5.])])
    assert oracle_mod.from_anchors(p, [0, 1, 0, 2], [0, 2], [0., 1., 5., 9.], [0., 7.]) == pytest.approx(0.4979, abs=5e-5)


def test_tag_pairing_truth_tables(oracle_mod):
    P, acc = oracle_mod.Params, oracle_mod.tag_pair_accepted
    A, B, C = 0, 1, 2
    p = P(1, tag_rule={"accept_same": True})
    assert acc(p, A, A) and not acc(p, A, B)
    p = P(1, tag_rule={"accept_same": False})
    assert not acc(p, A, A) and acc(p, A, B)
    pairs = [(A, B), (A, C), (B, C)]
    fwd, rev, same = [(A, B), (A, C), (B, C)], [(B, A), (C, A), (C, B)], [(A, A), (B, B), (C, C)]
    table = {  # (accepted_pairs, ordered) -> expected for (same, forward, reverse)
        (True, True): (False, True, False), (True, False): (False, True, True),
        (False, True): (True, False, True), (False, False): (True, False, False),
    }
    for (accepted, ordered), (e_same, e_fwd, e_rev) in table.items():
        p = P(1, tag_rule={"tag_pairs": pairs, "accepted_pairs": accepted, "ordered": ordered})
        assert all(acc(p, *q) == e_same for q in same)
        assert all(acc(p, *q) == e_fwd for q in fwd)
        assert all(acc(p, *q) == e_rev for q in rev)


def test_tag_rule_in_locohd(oracle_mod):
    xyz = [[0, 0, 0], [0, 1, 0], [2, 0, 0], [2, 2, 0], [1, 2, 0], [1, 3, 0], [3, 2, 0], [3, 3, 0], [2, 1, 0]]
    cat = [0, 0, 0, 0, 1, 1, 1, 1, 2]
    an = [(0, 3), (4, 5), (0, 4), (0, 8), (4, 8)]
    for tree in (True, False):
        p = oracle_mod.Params(3, [("uniform", [1., 1.001])], tag_rule={"accept_same": True})
        s = oracle_mod.from_primitives(p, xyz, cat, cat, xyz, cat, cat, an, 1.002, use_tree=tree)
        assert np.allclose(s, [0., 0., 1., 1., 1.], atol=1e-15, rtol=0)
        p = oracle_mod.Params(3, [("uniform", [1., 1.001])], tag_rule={"accept_same": False})
        s = oracle_mod.from_primitives(p, xyz, cat, cat, xyz, cat, cat, an, 1.002, use_tree=tree)
        assert np.allclose(s, [0.7071, 0.5412, 0.5412, 0.4284, 0.6501], atol=5e-5, rtol=0)


WF_KAT = [  # tests/test_wfs.py:8-27, 47-66, 86-105, 119-138
    ("hyper_exp", [1., 1.], [(0., 1., 0.6321), (1., 3., 0.3181), (5., 10., 0.0067)]),
    ("hyper_exp", [0.5, 0.5, 1 / 2., 1 / 3.], [(0., 1., 0.3385), (1., 3., 0.3660), (5., 10., 0.1143)]),
    ("hyper_exp", [3., 5., 2., 1 / 3., 1 / 5., 1 / 10.], [(0., 1., 0.1947), (1., 3., 0.2724), (5., 10., 0.2100)]),
    ("dagum", [1., 1., 1.], [(0., 1., 0.5), (1., 3., 0.25), (5., 10., 0.0758)]),
    ("dagum", [2., 5., 1.], [(0., 1., 0.0385), (1., 3., 0.2262), (5., 10., 0.3000)]),
    ("dagum", [10., 5., 2.], [(0., 1., 0.), (1., 3., 0.), (5., 10., 0.7480)]),
    ("uniform", [0., 1.], [(0., 1., 1.), (1., 3., 0.), (5., 10., 0.)]),
    ("uniform", [3., 10.], [(0., 1., 0.), (1., 3., 0.), (5., 10., 0.7143)]),
    ("uniform", [2., 16.], [(0., 1., 0.), (1., 3., 0.0714), (5., 10., 0.3571)]),
    ("kumaraswamy", [1., 2., 2., 2.], [(1.0, 2.0, 1.0), (1.25, 1.75, 0.6875), (1.4, 10.0, 0.7056)]),
    ("kumaraswamy", [5., 10., 2., 3.], [(5., 7., 0.4073), (1., 17., 1.0), (6.4, 6.7, 0.0910)]),
    ("kumaraswamy", [5., 9., 7., 7.], [(5., 7., 0.0534), (1., 17., 1.0), (6.4, 6.7, 0.0129)]),
]


@pytest.mark.parametrize("name,params,cases", WF_KAT)
def test_weight_function_known_answers(oracle_mod, name, params, cases):
    for a, b, want in cases:
        assert oracle_mod.wf_integral_range(name, params, a, b) == pytest.approx(want, abs=5e-5)
    assert oracle_mod.wf_integral_point(name, params, float("inf")) == pytest.approx(1.0, abs=1e-15)
    with pytest.raises(oracle_mod.OracleError):
        oracle_mod.wf_integral_point(name, params, -1.0)


def test_tree_query_equals_exhaustive_scan(oracle_mod):
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-20, 20, size=(600, 3))
    xyz[50] = xyz[10] + [7.0, 0.0, 0.0]  # a point exactly on the radius in exact arithmetic
    cat = rng.integers(0, 4, 600)
    tag = rng.integers(0, 9, 600)
    p = oracle_mod.Params(4, tag_rule={"accept_same": False})
    for anchor in (0, 10, 50, 333):
        for thr in (3.0, 7.0, 11.5, 80.0):
            i1, d1, c1 = oracle_mod.environment(p, xyz, cat, tag, anchor, thr, use_tree=True)
            i2, d2, c2 = oracle_mod.environment(p, xyz, cat, tag, anchor, thr, use_tree=False)
            assert sorted(i1.tolist()) == sorted(i2.tolist())
            assert np.array_equal(np.sort(d1), np.sort(d2))
            assert d1[0] == 0.0 and np.all(np.diff(d1) >= 0)


def test_against_python_twin(oracle_mod):
    from oracle import py_twin
    rng = np.random.default_rng(1)
    wfs = [("hyper_exp", [0.4, 0.6, 0.2, 0.05]), ("dagum", [1.5, 3.0, 7.0]), ("uniform", [2.0, 9.0]),
           ("kumaraswamy", [1.0, 12.0, 2.5, 3.5])]
    sds = [("Hellinger", [2.0]), ("Hellinger", [3.3]), ("Kolmogorov-Smirnov", []), ("Kullback-Leibler", [0.3]),
           ("Renyi", [2.5, 0.2])]
    for trial in range(30):
        wf, sd = wfs[trial % 4], sds[trial % 5]
        w = rng.uniform(0.5, 2.0, 4).tolist() if trial % 2 else None
        p = oracle_mod.Params(4, [wf], category_weights=w, statistical_distance=sd)
        na, nb = rng.integers(1, 40, 2)
        da = np.concatenate([[0.0], np.sort(rng.uniform(0, 14, na - 1).round(1))])
        db = np.concatenate([[0.0], np.sort(rng.uniform(0, 14, nb - 1).round(1))])
        sa, sb = rng.integers(0, 4, na), rng.integers(0, 4, nb)
        got = oracle_mod.from_anchors(p, sa, sb, da, db)
        ref = py_twin.stat_dist_integral(sa, sb, da, db, wf, sd, w or [1.0] * 4)
        flat = py_twin.flat_scan(sa, sb, da, db, wf, sd, w or [1.0] * 4)
        assert got == pytest.approx(ref, abs=1e-13)
        assert got == pytest.approx(flat, abs=1e-12)


def test_constructor_level_errors(oracle_mod):
    p = oracle_mod.Params(3)
    with pytest.raises(oracle_mod.OracleError):  # locohd.rs:70-73
        oracle_mod.from_anchors(p, [0, 1], [0], [0.0], [0.0])
    with pytest.raises(oracle_mod.OracleError):  # locohd.rs:74-77
        oracle_mod.from_anchors(p, [0, 1], [0], [0.1, 1.0], [0.0])
    with pytest.raises(oracle_mod.OracleError):  # pmf.rs:38-42
        oracle_mod.from_anchors(p, [0, 5], [0], [0.0, 1.0], [0.0])


def test_ensemble_callers_dmx_form_equals_the_cutoff_form(oracle_mod):
    """compare_ensembles.py:250-296 scores an ensemble with from_dmxs on full distance matrices whose homo-residue
    entries are set to +inf; BASELINE config 5 states the same computation as from_primitives with threshold 10 and
    the rule {"accept_same": False}.  With uniform [3, 10] the weight function is constant from 10 on, so members at or
    beyond the cutoff (and the banned ones at +inf) carry zero weight: the two forms give identical scores."""
    from benchdata import synth

    base = synth.gen(9, 40, 8, 7)
    a, b = synth.config5_member(base, 0), synth.config5_member(base, 1)
    p = oracle_mod.Params(7, [("uniform", [3.0, 10.0])], tag_rule={"accept_same": False})
    anchors = np.stack([np.arange(a.n)] * 2, axis=1).astype(np.uint32)
    cut = oracle_mod.from_primitives(p, a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)

    def dmx(c):
        d = c.xyz[None, :, :] - c.xyz[:, None, :]
        d = np.sqrt(np.sum(d ** 2, axis=2))
        same = c.tag[:, None] == c.tag[None, :]
        np.fill_diagonal(same, False)
        d[same] = np.inf                                      # "Ban homo-residue contacts"
        return d

    full = oracle_mod.from_dmxs(oracle_mod.Params(7, [("uniform", [3.0, 10.0])]), a.cat, b.cat, dmx(a), dmx(b))
    assert np.abs(full - cut).max() <= 1e-15
