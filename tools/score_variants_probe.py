"""K2 throughput of the scoring variants on one shape (ensemble of 48 structures, 5.64 M anchor pairs): the fast kernel
(Hellinger 2, unit weights) against the generic kernel's configurations (category weights, other exponents, other
statistical distances).  Development probe."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from benchdata import synth  # noqa: E402
from loco_hd_b200 import _capi, batch  # noqa: E402

S = 48
base = synth.config5_base()
clouds = [synth.config5_member(base, i) for i in range(S)]
ctx = _capi.Context(0)
pairs = batch.blocked_pairs(S, 4)
variants = [("Hellinger 2, unit weights (fast kernel)", dict()),
            ("Hellinger 2, category weights", dict(category_weights=[1 / 30.6, 1 / 1.07, 1 / 19.5, 1 / 5.0, 1 / 8.0, 1 / 3.0, 1 / 12.0])),
            ("Hellinger 3.3", dict(statistical_distance=("Hellinger", (3.3,)))),
            ("Kolmogorov-Smirnov", dict(statistical_distance=("Kolmogorov-Smirnov", ()))),
            ("Kullback-Leibler 0.5", dict(statistical_distance=("Kullback-Leibler", (0.5,)))),
            ("Renyi 2.3, 0.7", dict(statistical_distance=("Renyi", (2.3, 0.7))))]
for name, kw in variants:
    ctx.set_params(7, [("uniform", (3.0, 10.0))], tag_rule={"accept_same": False}, **kw)
    ens = batch.ResidentEnsemble.build(ctx, clouds, np.arange(base.n, dtype=np.uint32), 10.0)
    jobs = ens.job_table(pairs)
    out = ctx.pinned_array((len(jobs),), np.float64)
    ctx.score_jobs(ens.env, ens.env, jobs, want_scores=False, means_out=out)
    ctx.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        ctx.score_jobs(ens.env, ens.env, jobs, want_scores=False, means_out=out)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:45s} {1e3 * dt:8.2f} ms  {len(jobs) * base.n / dt / 1e6:8.1f} M anchor pairs/s   mean score {out.mean():.6f}")
    ens.close()
ctx.close()
