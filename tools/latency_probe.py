"""Latency of single calls (one structure pair) through the C ABI and through the Python drop-in API; development probe."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from benchdata import synth
from loco_hd_b200 import _capi

ctx = _capi.Context(0)
for name, (a, b), C, wf in (("config 1 (450 primitives, 150 anchors)", synth.config1(), 7, ("uniform", (3.0, 10.0))),
                            ("config 2 (10 000 primitives, 10 000 anchors)", synth.config2(), 7, ("kumaraswamy", (3.0, 10.0, 2.0, 5.0)))):
    ctx.set_params(C, [wf], tag_rule={"accept_same": False})
    step = a.k if "config 1" in name else 1
    anchors = np.stack([np.arange(0, a.n, step, dtype=np.uint32)] * 2, axis=1)
    for _ in range(5):
        ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        ctx.from_primitives(a.xyz, a.cat, a.tag, b.xyz, b.cat, b.tag, anchors, 10.0)
    dt = (time.perf_counter() - t0) / n
    print(f"C ABI  {name}: {1e3 * dt:.3f} ms per call, {len(anchors) / dt / 1e6:.2f} M anchor pairs/s")
    import loco_hd
    lchd = loco_hd.LoCoHD([f"T{i}" for i in range(C)], loco_hd.WeightFunction(wf[0], list(wf[1])),
                          loco_hd.TagPairingRule({"accept_same": False}))
    pa = [loco_hd.PrimitiveAtom(f"T{c}", str(t), x) for x, c, t in zip(a.xyz.tolist(), a.cat, a.tag)]
    pb = [loco_hd.PrimitiveAtom(f"T{c}", str(t), x) for x, c, t in zip(b.xyz.tolist(), b.cat, b.tag)]
    ap = [tuple(map(int, p)) for p in anchors]
    for _ in range(3):
        lchd.from_primitives(pa, pb, ap, 10.0)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        lchd.from_primitives(pa, pb, ap, 10.0)
    dt = (time.perf_counter() - t0) / n
    print(f"Python {name}: {1e3 * dt:.3f} ms per call, {len(anchors) / dt / 1e6:.2f} M anchor pairs/s")
ctx.close()
