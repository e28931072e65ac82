"""Wall-clock and per-kernel-group timings of the host-buffer (e2e) call sequence; development probe."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import argparse
import bench
from loco_hd_b200 import _capi

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
args = argparse.Namespace(workload="cfg2", pairs=a.pairs, models=500, frames=1024, ensemble=96)
wl = bench.make_workload("cfg2", 0, 1, args)
ctx = _capi.Context(0)
ctx.set_params(wl.C, [wl.wf], tag_rule=wl.rule)
h_xyz = ctx.pinned_array(wl.xyz.shape, np.float64); h_xyz[...] = wl.xyz
h_cat = ctx.pinned_array(wl.cat.shape, np.uint16); h_cat[...] = wl.cat
h_tag = ctx.pinned_array(wl.tag.shape, np.uint32); h_tag[...] = wl.tag
h_as = ctx.pinned_array(wl.anchor_struct.shape, np.uint32); h_as[...] = wl.anchor_struct
h_ap = ctx.pinned_array(wl.anchor_prim.shape, np.uint32); h_ap[...] = wl.anchor_prim
h_out = ctx.pinned_array((wl.n_pairs,), np.float64)
ctx.profile_enable(True)
for it in range(a.steps):
    t0 = time.perf_counter()
    st = ctx.structs_create(wl.offsets, h_xyz, h_cat, h_tag); ctx.synchronize(); t1 = time.perf_counter()
    env = ctx.envset_build(st, h_ap, wl.threshold, anchor_struct=h_as); ctx.synchronize(); t2 = time.perf_counter()
    ctx.score_jobs(env, env, wl.jobs, out=h_out); ctx.synchronize(); t3 = time.perf_counter()
    env.close(); st.close(); ctx.synchronize(); t4 = time.perf_counter()
    prof = ctx.profile_read()
    print(f"step {it}: structs {1e3*(t1-t0):.1f} ms, envset {1e3*(t2-t1):.1f} ms, score {1e3*(t3-t2):.1f} ms, "
          f"close {1e3*(t4-t3):.1f} ms | kernels {{" + ", ".join(f"{k}: {v[0]:.2f}" for k, v in prof.items()) + "}")
ctx.close()
