"""Dry run of GPU test BODIES on a box without a GPU: the `gpu_ctx` fixture is replaced by a stand-in with the same
methods answered by the CPU oracle (TEST INFRASTRUCTURE; nothing here is product code).  This checks the tests'
own logic - indexing, shapes, fixtures, tolerances against the exact values - before they meet the device; it says
nothing about the kernels.

    python tools/dryrun_gpu_tests.py            # the tests listed at the bottom
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402

oracle.build()


class _H:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def close(self):
        pass


class OracleCtx:
    tile_launches = 0

    def set_params(self, n_categories, weight_functions=(("uniform", (3.0, 10.0)),), category_weights=None,
                   statistical_distance=("Hellinger", (2.0,)), tag_rule=None):
        self.p = oracle.Params(n_categories, [(n, list(q)) for n, q in weight_functions], category_weights,
                               (statistical_distance[0], list(statistical_distance[1])), tag_rule)

    def from_primitives(self, xa, ca, ta, xb, cb, tb, anchors, thr, wf_idx=None):
        return oracle.from_primitives(self.p, xa, ca, ta, xb, cb, tb, anchors, thr, wf_idx=wf_idx)

    def structs_create(self, offs, xyz, cat, tag):
        return _H(off=np.asarray(offs, dtype=np.int64), xyz=np.asarray(xyz, dtype=np.float64), cat=np.asarray(cat), tag=np.asarray(tag))

    def structure(self, xyz, cat, tag):
        return self.structs_create([0, len(cat)], xyz, cat, tag)

    def envset_build(self, st, prim, thr, anchor_struct=None, keep_indices=False):
        prim = np.asarray(prim)
        s = np.zeros(len(prim), dtype=np.int64) if anchor_struct is None else np.asarray(anchor_struct, dtype=np.int64)
        return _H(st=st, prim=prim, struct=s, thr=thr)

    def _pair(self, ea, ia, eb, ib):
        sa, sb = int(ea.struct[ia]), int(eb.struct[ib])
        A, B = slice(ea.st.off[sa], ea.st.off[sa + 1]), slice(eb.st.off[sb], eb.st.off[sb + 1])
        return A, B

    def score_pairs(self, ea, eb, pairs):
        out = []
        for ia, ib in np.asarray(pairs).reshape(-1, 2):
            A, B = self._pair(ea, ia, eb, ib)
            out.append(oracle.from_primitives(self.p, ea.st.xyz[A], ea.st.cat[A], ea.st.tag[A], eb.st.xyz[B], eb.st.cat[B],
                                              eb.st.tag[B], [(int(ea.prim[ia]), int(eb.prim[ib]))], ea.thr)[0])
        return np.array(out)

    def score_jobs_stats(self, ea, eb, jobs, scores=False, job_means=False, anchor_means=False, anchor_stds=False):
        self.tile_launches += 1
        all_scores = []
        for a0, b0, n in [(int(j["a_first"]), int(j["b_first"]), int(j["n"])) for j in jobs]:
            A, B = self._pair(ea, a0, eb, b0)   # a job = one run of anchors of one structure pair
            assert len(set(ea.struct[a0:a0 + n])) == 1 and len(set(eb.struct[b0:b0 + n])) == 1
            an = np.stack([ea.prim[a0:a0 + n], eb.prim[b0:b0 + n]], axis=1).astype(np.uint32)
            all_scores.append(oracle.from_primitives(self.p, ea.st.xyz[A], ea.st.cat[A], ea.st.tag[A], eb.st.xyz[B],
                                                     eb.st.cat[B], eb.st.tag[B], an, ea.thr))
        res = {}
        if scores:
            res["scores"] = np.concatenate(all_scores)
        if job_means:
            res["job_means"] = np.array([s.mean() for s in all_scores])
        if anchor_means:
            res["anchor_means"] = np.mean(all_scores, axis=0)
        if anchor_stds:
            res["anchor_stds"] = np.std(all_scores, axis=0)
        return res


class _MonkeyPatch:
    def setenv(self, k, v):
        pass

    def delenv(self, k, raising=True):
        pass


if __name__ == "__main__":
    import test_gpu_tile_kernel as tk
    import test_highprec as hp

    ctx = OracleCtx()
    tk.test_sliced_unit_order_scores_the_same_units(ctx, oracle, _MonkeyPatch())
    print("ok  test_gpu_tile_kernel.py::test_sliced_unit_order_scores_the_same_units (test logic only)")
    hp.test_cuda_kernels_against_exact_values_at_the_ensemble_shape(ctx)
    print("ok  test_highprec.py::test_cuda_kernels_against_exact_values_at_the_ensemble_shape (test logic only)")
    hp.test_cuda_path_against_exact_values(ctx)
    hp.test_cuda_batch_path_against_exact_values(ctx)
    print("ok  test_highprec.py::test_cuda_path_against_exact_values / test_cuda_batch_path_against_exact_values (test logic only)")
