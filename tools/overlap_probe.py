"""Do host-to-device copies of one context overlap the kernels of another?  Development probe for the e2e path
(two host threads, one C-ABI context each, the config-4 shape of bench.py)."""
import argparse
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from loco_hd_b200 import _capi  # noqa: E402

args = argparse.Namespace(workload="cfg4", pairs=256, models=500, frames=1024, ensemble=96)
wl = bench.make_workload("cfg4", 0, 1, args)


def make_ctx():
    c = _capi.Context(0)
    c.set_params(wl.C, [wl.wf], tag_rule=wl.rule)
    return c


def pinned(c, a):
    h = c.pinned_array(a.shape, a.dtype)
    h[...] = a
    return h


ca, cb = make_ctx(), make_ctx()
hx, hc, ht = pinned(ca, wl.xyz), pinned(ca, wl.cat), pinned(ca, wl.tag)
h32 = pinned(ca, wl.xyz.astype(np.float32))
st_b = cb.structs_create(wl.offsets, wl.xyz, wl.cat, wl.tag)
h_as, h_ap = pinned(cb, wl.anchor_struct), pinned(cb, wl.anchor_prim)
out_b = cb.pinned_array((wl.n_pairs,), np.float64)
N = 12


def copy_loop(xyz):
    for _ in range(N):
        st = ca.structs_create(wl.offsets, xyz, hc, ht)
        st.close()
    ca.synchronize()


def compute_loop():
    for _ in range(N):
        st_b.drop_cells()
        env = cb.envset_build(st_b, h_ap, wl.threshold, anchor_struct=h_as)
        cb.score_jobs(env, env, wl.jobs, out=out_b)
        env.close()
    cb.synchronize()


def timed(*fns):
    ths = [threading.Thread(target=f) for f in fns]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return 1e3 * (time.perf_counter() - t0) / N


copy_loop(hx); compute_loop()
print(f"copy alone (f64, {wl.xyz.nbytes / 1e6:.0f} MB xyz): {timed(lambda: copy_loop(hx)):.2f} ms / step")
print(f"copy alone (f32): {timed(lambda: copy_loop(h32)):.2f} ms / step")
print(f"compute alone: {timed(compute_loop):.2f} ms / step")
print(f"both (f64 copy): {timed(lambda: copy_loop(hx), compute_loop):.2f} ms / step")
print(f"both (f32 copy): {timed(lambda: copy_loop(h32), compute_loop):.2f} ms / step")
