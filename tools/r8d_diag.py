"""One-shot diagnostic of the sliced unit order (prints everything, asserts nothing)."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle
from benchdata import synth
from loco_hd_b200 import _capi, batch

JOB = [("a_first", "<u8"), ("b_first", "<u8"), ("n", "<u8")]
ctx = _capi.Context(0)
base = synth.gen(31, 29, 7, 7)
members = [synth.partner(base, 1.5, 900 + i) for i in range(10)]
anchors = np.arange(base.n, dtype=np.uint32)
ctx.set_params(7, [("uniform", (3.0, 10.0))], None, ("Hellinger", (2.0,)), {"accept_same": False})
op = oracle.Params(7, [("uniform", [3.0, 10.0])], None, ("Hellinger", [2.0]), {"accept_same": False})
offs = np.cumsum([0] + [c.n for c in members]).astype(np.uint64)
st = ctx.structs_create(offs, np.concatenate([c.xyz for c in members]), np.concatenate([c.cat for c in members]), np.concatenate([c.tag for c in members]))
env = ctx.envset_build(st, np.tile(anchors, len(members)), 10.0, anchor_struct=np.repeat(np.arange(len(members), dtype=np.uint32), base.n))
pairs = batch.blocked_pairs(len(members), 4)
jobs = np.array([(i * base.n, j * base.n, base.n) for i, j in pairs], dtype=JOB)
want = np.stack([oracle.from_primitives(op, members[i].xyz, members[i].cat, members[i].tag, members[j].xyz, members[j].cat, members[j].tag,
                                        np.stack([anchors, anchors], axis=1), 10.0) for i, j in pairs]).ravel()
def run(label, **envv):
    for k in ("LOCOHD_TILE_SLICE", "LOCOHD_TILE_SLICE_MB", "LOCOHD_NO_TILES"):
        os.environ.pop(k, None)
    os.environ.update(envv)
    b = ctx.tile_launches
    try:
        r = ctx.score_jobs_stats(env, env, jobs, scores=True, job_means=True)
    except Exception as e:
        print(label, "EXCEPTION", repr(e)); return None
    s = r["scores"]
    print(label, "tile launches", ctx.tile_launches - b, "max|s-oracle|", float(np.nanmax(np.abs(s - want))), "nan", int(np.isnan(s).sum()), flush=True)
    return s, r["job_means"]
ref = run("slice=0 (1)", LOCOHD_TILE_SLICE="0")
ref2 = run("slice=0 (2)", LOCOHD_TILE_SLICE="0")
if ref and ref2:
    print("tile-major run-to-run identical:", bool(np.array_equal(ref[0], ref2[0])))
for s in ("8", "24", "200", "203", "1000"):
    got = run("slice=" + s, LOCOHD_TILE_SLICE=s)
    if got and ref:
        d = np.flatnonzero(got[0] != ref[0])
        print("   differing scores:", len(d), "max|diff|", float(np.abs(got[0] - ref[0]).max()), "means equal", bool(np.array_equal(got[1], ref[1])),
              "first (job, anchor):", [(int(x // base.n), int(x % base.n)) for x in d[:6]], flush=True)
got = run("auto 0.25MB", LOCOHD_TILE_SLICE_MB="0.25")
if got and ref:
    print("   differing scores:", int((got[0] != ref[0]).sum()))
got = run("pair kernel", LOCOHD_NO_TILES="1")
if got and ref:
    print("   pair kernel vs tile-major max|diff|", float(np.abs(got[0] - ref[0]).max()))
