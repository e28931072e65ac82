"""CPU-side checks of what this implementation adds around the reference API: the type stub describes the
extension module that is actually built, the id helpers work without a device, scoring without a device fails loudly
from several threads at once (no deadlock, no fallback), and the per-rank statistics pool exactly."""
import ast
import threading
from pathlib import Path

import numpy as np
import pytest

import loco_hd
from loco_hd_b200 import batch

ROOT = Path(__file__).resolve().parent.parent


def test_stub_matches_the_built_module():
    """loco_hd/loco_hd.pyi (the API contract, reference: loco_hd/loco_hd.pyi:7-290) names exactly the public classes,
    methods and functions of the built extension loco_hd.loco_hd; the reference's five classes keep their signatures."""
    native = loco_hd.loco_hd
    assert Path(native.__file__).parent == ROOT / "loco_hd" and Path(native.__file__).suffix == ".so"
    tree = ast.parse((ROOT / "loco_hd" / "loco_hd.pyi").read_text())
    classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
    assert {"WeightFunction", "PrimitiveAtom", "TagPairingRule", "StatisticalDistance", "LoCoHD"} <= set(classes)
    for cname, node in classes.items():
        cls = getattr(native, cname)
        stub_methods = {f.name for f in node.body if isinstance(f, ast.FunctionDef)}
        stub_attrs = {a.target.id for a in node.body if isinstance(a, ast.AnnAssign)}
        for name in (stub_methods | stub_attrs) - {"__init__", "__len__"}:
            assert hasattr(cls, name), f"{cname}.{name} is in the stub but not in the module"
        public = {n for n in vars(cls) if not n.startswith("_")}
        assert public <= (stub_methods | stub_attrs), f"{cname}: undocumented {public - stub_methods - stub_attrs}"
    for fn in (n.name for n in tree.body if isinstance(n, ast.FunctionDef)):
        assert callable(getattr(native, fn))

    def args_of(cname, fname):
        f = [f for f in classes[cname].body if isinstance(f, ast.FunctionDef) and f.name == fname][0]
        return [a.arg for a in f.args.args[1:]], len(f.args.defaults)

    # positional order, names and number of defaults of the reference's signatures (src/locohd.rs:290-302, 392, 410, 463, 479)
    assert args_of("LoCoHD", "__init__") == (["categories", "w_func", "tag_pairing_rule", "n_of_threads", "category_weights",
                                              "statistical_distance"], 5)
    assert args_of("LoCoHD", "from_anchors") == (["seq_a", "seq_b", "dists_a", "dists_b", "w_func_key"], 1)
    assert args_of("LoCoHD", "from_dmxs") == (["seq_a", "seq_b", "dmx_a", "dmx_b", "w_func_keys"], 1)
    assert args_of("LoCoHD", "from_coords") == (["seq_a", "seq_b", "coords_a", "coords_b", "w_func_keys"], 1)
    assert args_of("LoCoHD", "from_primitives") == (["prim_a", "prim_b", "anchor_pairs", "threshold_distance"], 0)
    assert args_of("WeightFunction", "__init__") == (["function_name", "parameters"], 0)
    assert args_of("PrimitiveAtom", "__init__") == (["primitive_type", "tag", "coordinates"], 0)
    assert args_of("StatisticalDistance", "__init__") == (["distance_name", "parameters"], 0)
    # the keyword names work on the built module
    lchd = loco_hd.LoCoHD(categories=["A", "B"], w_func=None, tag_pairing_rule=None, n_of_threads=2, category_weights=[1., 2.],
                          statistical_distance=loco_hd.StatisticalDistance(distance_name="Hellinger", parameters=[2.]))
    assert lchd.category_weights == [1., 2.]


def test_id_helpers_without_a_device():
    lchd = loco_hd.LoCoHD(["O", "N", "C"])
    ids = lchd.category_ids(np.array(["C", "O", "zzz", "N"]))
    assert ids.dtype == np.uint16 and ids.tolist() == [2, 0, 0xFFFF, 1]
    tags = lchd.intern_tags(["A/1-GLY", "A/2-ALA", "A/1-GLY"])
    assert tags.dtype == np.uint32 and tags[0] == tags[2] != tags[1]
    # list[PrimitiveAtom] -> arrays (lists, tuples and generators of temporaries alike), consistent with the helpers
    atoms = [loco_hd.PrimitiveAtom("ONC"[i % 3], f"A/{i // 4}-GLY", [float(i), 0.5, -1.0]) for i in range(50)]
    atoms[7] = loco_hd.PrimitiveAtom("unknown", "A/1-GLY", [7.0, 0.5, -1.0])
    for form in (atoms, tuple(atoms), (a for a in atoms), iter(atoms)):
        xyz, cat, tag = lchd.to_arrays(form)
        assert xyz.shape == (50, 3) and np.array_equal(xyz[:, 0], np.arange(50.0))
        assert np.array_equal(cat, lchd.category_ids([a.primitive_type for a in atoms])) and cat[7] == 0xFFFF
        assert np.array_equal(tag, lchd.intern_tags([a.tag for a in atoms]))
    gen = (loco_hd.PrimitiveAtom("O", f"t{i // 3}", [float(i), 0., 0.]) for i in range(3000))   # temporaries
    xyz, cat, tag = lchd.to_arrays(gen)
    assert np.array_equal(tag - tag[0], np.arange(3000) // 3)
    with pytest.raises(TypeError):
        lchd.to_arrays([1, 2, 3])


def test_scoring_without_a_device_fails_loudly_from_two_threads():
    if loco_hd.loco_hd.device_count() > 0:
        pytest.skip("a CUDA device is present")
    lchd = loco_hd.LoCoHD(["A", "B"])
    atoms = [loco_hd.PrimitiveAtom("A", "", [0., 0., 0.]), loco_hd.PrimitiveAtom("B", "", [1., 0., 0.])]
    seen = []

    def work():
        for _ in range(20):
            for call in (lambda: lchd.from_primitives(atoms, atoms, [(0, 0)], 5.0),
                         lambda: lchd.from_anchors(["A"], ["B"], [0.], [0.]),
                         lambda: lchd.from_coords(["A"], ["B"], [[0., 0., 0.]], [[0., 0., 0.]])):
                try:
                    call()
                    seen.append("returned")
                except RuntimeError as exc:
                    seen.append("no CPU fallback" in str(exc))

    threads = [threading.Thread(target=work, daemon=True) for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not any(t.is_alive() for t in threads)
    assert len(seen) == 120 and all(v is True for v in seen)


def test_input_errors_come_before_the_device_error():
    """from_primitives does its Python-side work (anchor parsing, flattening of any iterable of PrimitiveAtom) before
    it touches the device: bad inputs are reported as such even on a box without a GPU."""
    if loco_hd.loco_hd.device_count() > 0:
        pytest.skip("a CUDA device is present: covered by the gpu tests")
    lchd = loco_hd.LoCoHD(["A", "B"])
    gen = (loco_hd.PrimitiveAtom("AB"[i % 2], f"tag{i // 3}", [float(i), 0., 0.]) for i in range(5000))
    with pytest.raises(RuntimeError):   # the device is missing, but only after the Python-side work
        lchd.from_primitives(gen, [loco_hd.PrimitiveAtom("A", "t", [0., 0., 0.])], [(0, 0)], 5.0)
    with pytest.raises(TypeError):
        lchd.from_primitives([object()], [loco_hd.PrimitiveAtom("A", "t", [0., 0., 0.])], [(0, 0)], 5.0)


def test_combine_anchor_stats_pools_exactly():
    rng = np.random.default_rng(3)
    x = rng.random((37, 11))
    parts = [x[:5], x[5:20], x[20:]]
    mean, std = batch.combine_anchor_stats([len(q) for q in parts], [q.mean(axis=0) for q in parts],
                                           [q.std(axis=0) for q in parts])
    assert np.abs(mean - x.mean(axis=0)).max() < 1e-15 and np.abs(std - x.std(axis=0)).max() < 1e-15


def test_pair_anchors_by_tag_is_the_callers_pairing_loop():
    """batch.pair_anchors_by_tag against the dict-based pairing of casp14_extend_with_locohd.py:48,66-79, on models
    that lack residues, carry extra ones, list them in another order and repeat a tag."""
    rng = np.random.default_rng(11)
    cent = 7
    for trial in range(20):
        n_res = int(rng.integers(1, 40))
        k = int(rng.integers(1, 5))
        ref_tag = np.repeat(rng.permutation(100)[:n_res], k).astype(np.uint32)
        ref_cat = rng.integers(0, 7, n_res * k)
        ref_cat[::k] = cent
        keep = rng.random(n_res) < 0.8
        m_res = np.concatenate([np.unique(ref_tag)[rng.permutation(n_res)][: int(keep.sum())], [200, 201]])
        if trial % 3 == 0 and len(m_res) > 2:
            m_res = np.concatenate([m_res, m_res[:1]])            # a tag that occurs twice in the model
        model_tag = np.repeat(m_res, k).astype(np.uint32)
        model_cat = rng.integers(0, 7, len(model_tag))
        model_cat[::k] = cent
        lut = {int(model_tag[i]): i for i in range(len(model_tag)) if model_cat[i] == cent}
        want = [(i, lut[int(ref_tag[i])]) for i in range(len(ref_tag)) if ref_cat[i] == cent and int(ref_tag[i]) in lut]
        got = batch.pair_anchors_by_tag(ref_cat, ref_tag, model_cat, model_tag, cent)
        assert got.dtype == np.uint32 and got.shape == (len(want), 2)
        assert got.tolist() == [list(p) for p in want]
    assert batch.pair_anchors_by_tag([1, 2], [0, 1], [3], [0], 7).shape == (0, 2)
    assert batch.models_against_reference(None, ([], [], []), [], 7, 10.0) == []


class _OracleBackedLoCoHD:
    """Stand-in for the public class with the resident-batch methods answered by the CPU oracle (TEST ONLY): lets the
    host-side logic of batch.models_against_reference run without a device."""

    def __init__(self, oracle_mod, params):
        self.o, self.p = oracle_mod, params

    class _H:
        def __init__(self, **kw):
            self.__dict__.update(kw)
            self.closed = False

        def close(self):
            self.closed = True

    def structures(self, offsets, xyz, cats, tags):
        assert len(xyz) == len(cats) == len(tags) == int(offsets[-1])
        self.last = self._H(off=np.asarray(offsets, dtype=np.int64), xyz=np.asarray(xyz, dtype=np.float64), cat=cats, tag=tags)
        return self.last

    def environments(self, st, prim, thr, anchor_struct=None):
        return self._H(st=st, prim=np.asarray(prim), struct=np.asarray(anchor_struct), thr=thr)

    def score_batch(self, ea, eb, jobs, reduce=None):
        scores, means = [], []
        for a0, b0, n in np.asarray(jobs, dtype=np.int64):
            s = []
            for i in range(n):
                sa, sb = int(ea.struct[a0 + i]), int(eb.struct[b0 + i])
                A = slice(ea.st.off[sa], ea.st.off[sa + 1])
                B = slice(eb.st.off[sb], eb.st.off[sb + 1])
                # anchors inside a structure are indices into that structure
                s.append(self.o.from_primitives(self.p, ea.st.xyz[A], ea.st.cat[A], ea.st.tag[A], eb.st.xyz[B], eb.st.cat[B],
                                                eb.st.tag[B], [(int(ea.prim[a0 + i]), int(eb.prim[b0 + i]))], ea.thr)[0])
            scores += s
            means.append(np.mean(s))
        return {"scores": np.array(scores), "job_mean": np.array(means)}


def test_models_against_reference_host_logic(oracle_mod):
    """The job / anchor bookkeeping of batch.models_against_reference (casp14_extend_with_locohd.py:44-88) with the
    device calls answered by the oracle: per-model scores equal the per-model from_primitives calls."""
    from benchdata import synth

    ref = synth.gen(5, 30, 5, 8, with_centroid=True)
    rng = np.random.default_rng(17)
    models = []
    for m in range(4):
        full = synth.config3_model(ref, m)
        res = np.arange(30)
        if m == 1:
            res = res[rng.random(30) < 0.8]
        if m == 2:
            res = rng.permutation(res)
        idx = (res[:, None] * 5 + np.arange(5)[None, :]).ravel()
        if m == 3:
            models.append((full.xyz[:10], full.cat[:10], (full.tag[:10] + 1000).astype(np.uint32)))
        else:
            models.append((full.xyz[idx], full.cat[idx], full.tag[idx]))
    op = oracle_mod.Params(8, [("uniform", [3.0, 10.0])], tag_rule={"accept_same": False})
    fake = _OracleBackedLoCoHD(oracle_mod, op)
    out = batch.models_against_reference(fake, (ref.xyz, ref.cat, ref.tag), models, int(ref.centroid_cat), 10.0)
    assert fake.last.closed
    assert [len(p) for p, _, _ in out] == [30, len(np.unique(models[1][2])), 30, 0]
    for m, (pairs, scores, mean) in enumerate(out):
        if m == 3:
            assert len(scores) == 0 and np.isnan(mean)
            continue
        mx, mc, mt = models[m]
        assert np.array_equal(ref.tag[pairs[:, 0]], mt[pairs[:, 1]])
        want = oracle_mod.from_primitives(op, ref.xyz, ref.cat, ref.tag, mx, mc, mt, pairs, 10.0)
        assert np.array_equal(scores, want) and mean == pytest.approx(want.mean(), abs=1e-15)
