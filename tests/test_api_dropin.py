"""The Python drop-in API (`import loco_hd`) against the reference's own test-suite values:
/root/reference/tests/test_wfs.py, test_tag_pairing_rule.py, test_locohd.py.  Constructor / host-utility tests run
without a GPU; everything that scores is marked gpu (the API has no CPU scoring path)."""
import numpy as np
import pytest

import loco_hd
from loco_hd import LoCoHD, PrimitiveAtom, StatisticalDistance, TagPairingRule, WeightFunction
from test_oracle_kat import WF_KAT


# ------------------------------------------------------------------------------------------- host-side (CPU)
def test_import_surface():
    # loco_hd/__init__.py:1-2 upstream
    for name in ("WeightFunction", "PrimitiveAtom", "TagPairingRule", "LoCoHD", "StatisticalDistance",
                 "PrimitiveAssigner", "PrimitiveAtomTemplate", "PrimitiveAtomSource", "TypingSchemeElement"):
        assert hasattr(loco_hd, name)
    from loco_hd.loco_hd import LoCoHD as native  # the native module path of the reference
    assert native is LoCoHD


@pytest.mark.parametrize("name,params,cases", WF_KAT)
def test_weight_function_values(name, params, cases):
    wf = WeightFunction(name, params)
    assert wf.function_name == name and wf.parameters == [float(p) for p in params]
    for a, b, want in cases:
        assert wf.integral_range(a, b) == pytest.approx(want, abs=5e-5)
    assert wf.integral_vec([0.0, 1.0]) == [wf.integral_point(0.0), wf.integral_point(1.0)]
    with pytest.raises(ValueError):
        wf.integral_point(-0.5)


@pytest.mark.parametrize("name,params", [
    ("hyper_exp", [1.]), ("hyper_exp", [1., 2., 3.]), ("hyper_exp", [-1., 1.]), ("hyper_exp", [1., -1.]),
    ("hyper_exp", [1., -1., 2.]), ("dagum", [1.]), ("dagum", [1., 2.]), ("dagum", [-1., 2., 3.]),
    ("dagum", [1., -2., 3.]), ("dagum", [1., 2., -3.]), ("uniform", [1.]), ("uniform", [1., 0.]),
    ("uniform", [-1., 0.]), ("kumaraswamy", [1.]), ("kumaraswamy", [1., 2.]), ("kumaraswamy", [3., 1., 2., 2.]),
    ("kumaraswamy", [1., 3., -2., 2.]), ("kumaraswamy", [0., 3., 2., -2.]), ("nonsense", [1., 2.]),
])
def test_weight_function_errors(name, params):
    with pytest.raises(ValueError):  # tests/test_wfs.py:29-45, 68-84, 107-117, 140-156
        WeightFunction(name, params)


def test_tag_pairing_rule_truth_tables():
    tpr = TagPairingRule({"accept_same": True})
    assert tpr.pair_accepted(("A", "A")) and not tpr.pair_accepted(("A", "B"))
    tpr = TagPairingRule({"accept_same": False})
    assert not tpr.pair_accepted(("A", "A")) and tpr.pair_accepted(("A", "B"))
    pairs = {("A", "B"), ("A", "C"), ("B", "C")}
    fwd, rev, same = [("A", "B"), ("A", "C"), ("B", "C")], [("B", "A"), ("C", "A"), ("C", "B")], [("A", "A"), ("B", "B"), ("C", "C")]
    table = {(True, True): (False, True, False), (True, False): (False, True, True),
             (False, True): (True, False, True), (False, False): (True, False, False)}
    for (accepted, ordered), (e_same, e_fwd, e_rev) in table.items():
        tpr = TagPairingRule({"tag_pairs": pairs, "accepted_pairs": accepted, "ordered": ordered})
        assert all(tpr.pair_accepted(q) == e_same for q in same)
        assert all(tpr.pair_accepted(q) == e_fwd for q in fwd)
        assert all(tpr.pair_accepted(q) == e_rev for q in rev)
    assert "WithoutList" in TagPairingRule({"accept_same": True}).get_dbg_str()
    with pytest.raises(TypeError):
        TagPairingRule({"something": 1})


def test_locohd_constructor_errors_and_getters():
    w_func = WeightFunction("uniform", [0., 4.])
    types = ["O", "A", "B", "C"]
    for kwargs in (dict(categories=[]), dict(categories=types, category_weights=[1., 1., 1.]),
                   dict(categories=types, category_weights=[1.] * 5), dict(categories=types, category_weights=[1., -1., 1., 1.]),
                   dict(categories=types, category_weights=[1., 0., 1., 1.])):
        with pytest.raises(ValueError):  # tests/test_locohd.py:54-73
            LoCoHD(w_func=w_func, **kwargs)
    lchd = LoCoHD(types, w_func, n_of_threads=4, category_weights=[1., 2., 3., 4.],
                  statistical_distance=StatisticalDistance("Hellinger", [2.]))
    assert lchd.categories == {"O": 0, "A": 1, "B": 2, "C": 3}
    assert lchd.category_weights == [1., 2., 3., 4.]
    assert lchd.w_func.function_name == "uniform"
    assert lchd.tag_pairing_rule.pair_accepted(("x", "x")) and not lchd.tag_pairing_rule.pair_accepted(("x", "y"))
    multi = LoCoHD(types, {"near": WeightFunction("uniform", [0., 4.]), "far": WeightFunction("uniform", [3., 10.])})
    assert set(multi.w_func) == {"near", "far"}
    default = LoCoHD(types)  # locohd.rs:349-354: uniform [3, 10]
    assert default.w_func.parameters == [3., 10.]
    with pytest.raises(ValueError):
        StatisticalDistance("Hellinger", [])
    with pytest.raises(ValueError):
        StatisticalDistance("Euclid", [1.])
    assert StatisticalDistance("Kolmogorov-Smirnov", []).run([0.2, 0.8], [0.5, 0.5]) == pytest.approx(0.3)


def test_primitive_atom_fields():
    p = PrimitiveAtom("Cent", "A/12-GLY", np.array([1., 2., 3.], dtype=np.float32))
    assert (p.primitive_type, p.tag, p.coordinates) == ("Cent", "A/12-GLY", [1., 2., 3.])
    p.primitive_type, p.tag, p.coordinates = "O_neg", "B/1-ASP", [4, 5, 6]
    assert (p.primitive_type, p.tag, p.coordinates) == ("O_neg", "B/1-ASP", [4., 5., 6.])
    with pytest.raises((TypeError, ValueError)):
        PrimitiveAtom("x", "y", [1., 2.])


def test_weight_function_key_pairing_errors_need_no_gpu():
    # keys_to_weight_functions (locohd.rs:230-283) is checked before any device work
    single = LoCoHD(["A"], WeightFunction("uniform", [0., 4.]))
    multi = LoCoHD(["A"], {"k": WeightFunction("uniform", [0., 4.])})
    with pytest.raises(ValueError):
        single.from_anchors(["A"], ["A"], [0.], [0.], "k")
    with pytest.raises(ValueError):
        multi.from_anchors(["A"], ["A"], [0.], [0.])
    with pytest.raises(ValueError):
        multi.from_anchors(["A"], ["A"], [0.], [0.], "missing")
    with pytest.raises(ValueError):  # an empty anchor list counts as "with keys" upstream (locohd.rs:34-40, 276-281)
        single.from_primitives([], [], [], 10.0)
    assert multi.from_primitives([], [], [], 10.0) == []


# ------------------------------------------------------------------------------------------------ scoring (GPU)
@pytest.mark.gpu
def test_small_locohd():
    # tests/test_locohd.py:27-52
    lchd = LoCoHD(["O", "A", "B", "C"], WeightFunction("uniform", [0., 4.]))
    seq = ["O", "A", "B", "C"]
    assert lchd.from_anchors(seq, seq, [0., 1., 2., 3.], [0., 1., 1., 1.]) == pytest.approx(0.2268, abs=5e-5)
    assert lchd.from_anchors(seq, seq, [0., 1., 1., 1.], [0., 1., 2., 3.]) == pytest.approx(0.2268, abs=5e-5)
    lchd = LoCoHD(["A", "B", "C"], WeightFunction("kumaraswamy", [3., 10., 2., 5.]))
    assert lchd.from_anchors(["A", "B", "A", "C"], ["A", "C"], [0., 1., 5., 9.], [0., 7.]) == pytest.approx(0.4979, abs=5e-5)
    with pytest.raises(ValueError):
        lchd.from_anchors(["A", "B"], ["A"], [0.], [0.])
    with pytest.raises(ValueError):
        lchd.from_anchors(["A", "Z"], ["A"], [0., 1.], [0.])  # unknown category (pmf.rs:38-42)


@pytest.mark.gpu
def test_in_locohd():
    # tests/test_tag_pairing_rule.py:100-157
    structure = [PrimitiveAtom(t, t, c) for t, c in [
        ("A", [0., 0., 0.]), ("A", [0., 1., 0.]), ("A", [2., 0., 0.]), ("A", [2., 2., 0.]),
        ("B", [1., 2., 0.]), ("B", [1., 3., 0.]), ("B", [3., 2., 0.]), ("B", [3., 3., 0.]), ("C", [2., 1., 0.])]]
    anchors = [(0, 3), (4, 5), (0, 4), (0, 8), (4, 8)]
    wf = WeightFunction("uniform", [1., 1.001])
    scores = LoCoHD(["A", "B", "C"], wf, TagPairingRule({"accept_same": True})).from_primitives(structure, structure, anchors, 1.002)
    for got, want in zip(scores, [0., 0., 1., 1., 1.]):
        assert got == pytest.approx(want, abs=1e-15)
    scores = LoCoHD(["A", "B", "C"], wf, TagPairingRule({"accept_same": False})).from_primitives(structure, structure, anchors, 1.002)
    for got, want in zip(scores, [0.7071, 0.5412, 0.5412, 0.4284, 0.6501]):
        assert got == pytest.approx(want, abs=5e-5)
    assert isinstance(scores, list) and isinstance(scores[0], float)


@pytest.mark.gpu
def test_api_matches_oracle_on_fixture_style_clouds(oracle_mod):
    """Same construction as the reference's consistency test (tests/test_locohd.py:100-118): PrimitiveAtom(x, "", y),
    anchors (x, x), threshold 50, varying weight functions and statistical distances, plus per-anchor keys,
    tag lists, from_coords and from_dmxs through the Python API."""
    rng = np.random.default_rng(3)
    types = ["A", "B", "C", "D", "E"]
    clouds = []
    for n in (97, 140):
        clouds.append((rng.choice(types, n).tolist(), rng.uniform(-50, 50, size=(n, 3))))
    (s1, x1), (s2, x2) = clouds
    ids = {t: i for i, t in enumerate(types)}
    c1, c2 = [ids[t] for t in s1], [ids[t] for t in s2]
    pras1 = [PrimitiveAtom(t, "", c) for t, c in zip(s1, x1.tolist())]
    pras2 = [PrimitiveAtom(t, "", c) for t, c in zip(s2, x2.tolist())]
    anchors = [(x, x) for x in range(min(len(pras1), len(pras2)))]
    zeros1, zeros2 = np.zeros(len(s1), np.uint32), np.zeros(len(s2), np.uint32)
    cases = [(("hyper_exp", [1., 0.1]), ("Hellinger", [2.])), (("dagum", [1.2, 3.0, 9.0]), ("Hellinger", [3.26])),
             (("uniform", [4.0, 12.5]), ("Kolmogorov-Smirnov", [])), (("kumaraswamy", [2., 9., 3.3, 4.4]), ("Kullback-Leibler", [2.2])),
             (("uniform", [3., 10.]), ("Renyi", [1.7, 0.9]))]
    for wf, sd in cases:
        lchd = LoCoHD(categories=types, w_func=WeightFunction(*wf), statistical_distance=StatisticalDistance(*sd))
        got = np.array(lchd.from_primitives(pras1, pras2, anchors, 50.0))
        op = oracle_mod.Params(5, [wf], statistical_distance=sd)
        ref = oracle_mod.from_primitives(op, x1, c1, zeros1, x2, c2, zeros2, anchors, 50.0)
        assert np.abs(got - ref).max() <= 1e-9
        # from_coords / from_dmxs: every point an anchor, whole structure as environment (numpy inputs accepted)
        n = len(s1)
        got = np.array(lchd.from_coords(np.array(s1), np.array(s2[:n]), x1, x2[:n]))
        ref = oracle_mod.from_coords(op, c1, c2[:n], x1, x2[:n])
        assert np.abs(got - ref).max() <= 1e-9
        dm = lambda x: np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
        got = np.array(lchd.from_dmxs(s1, s2[:n], dm(x1).tolist(), dm(x2[:n])))
        ref = oracle_mod.from_dmxs(op, c1, c2[:n], dm(x1), dm(x2[:n]))
        assert np.abs(got - ref).max() <= 1e-9
    # per-anchor weight function keys + a tag-pair list rule
    wfs = {"near": WeightFunction("uniform", [0., 6.]), "far": WeightFunction("hyper_exp", [1., 0.05])}
    tags1 = rng.choice(["r1", "r2", "r3"], len(s1)).tolist()
    tags2 = rng.choice(["r1", "r2", "r3"], len(s2)).tolist()
    rule = {"tag_pairs": {("r1", "r2"), ("r3", "r3")}, "accepted_pairs": False, "ordered": False}
    lchd = LoCoHD(types, wfs, TagPairingRule(rule))
    pras1 = [PrimitiveAtom(t, g, c) for t, g, c in zip(s1, tags1, x1)]
    pras2 = [PrimitiveAtom(t, g, c) for t, g, c in zip(s2, tags2, x2)]
    keyed = [(i, i, "near" if i % 3 else "far") for i in range(len(anchors))]
    got = np.array(lchd.from_primitives(pras1, pras2, keyed, 30.0))
    tid = {"r1": 0, "r2": 1, "r3": 2}
    op = oracle_mod.Params(5, [("hyper_exp", [1., 0.05]), ("uniform", [0., 6.])],  # keys sorted: far, near
                           tag_rule={"tag_pairs": [(0, 1), (2, 2)], "accepted_pairs": False, "ordered": False})
    ref = oracle_mod.from_primitives(op, x1, c1, [tid[t] for t in tags1], x2, c2, [tid[t] for t in tags2], anchors, 30.0,
                                     wf_idx=[1 if i % 3 else 0 for i in range(len(anchors))])
    assert np.abs(got - ref).max() <= 1e-9
    with pytest.raises(ValueError):
        lchd.from_primitives(pras1, pras2, [(0, 0, "nope")], 30.0)
    with pytest.raises(ValueError):
        lchd.from_primitives(pras1, pras2, [(0, 10 ** 6, "near")], 30.0)
    # array API equals the object API
    single = LoCoHD(types, WeightFunction("uniform", [3., 10.]), TagPairingRule({"accept_same": False}))
    a = np.array(single.from_primitives(pras1, pras2, anchors, 30.0))
    b = single.from_arrays(x1, np.array(c1, np.uint16), np.array([tid[t] for t in tags1], np.uint32), x2,
                           np.array(c2, np.uint16), np.array([tid[t] for t in tags2], np.uint32),
                           np.array(anchors, np.uint32), 30.0)
    assert np.array_equal(a, b)
