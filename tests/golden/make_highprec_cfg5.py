"""Builds tests/golden/highprec_cfg5.npz: 60-digit (mpmath) LoCoHD scores at the BASELINE configs[4] shape - two full
members of the 1000-structure ensemble (benchdata.synth.config5_member, 5000 primitives, C = 7), uniform [3, 10],
threshold 10, hetero-residue contacts only, 32 anchors spread over the structure (~380 merged events per anchor pair).
Only the anchors and the exact values are stored; the structures are regenerated from the seeded generator and their
checksum is stored with the values.  Same definition and code as make_highprec.py (independent of oracle/).

    python tests/golden/make_highprec_cfg5.py
"""
import hashlib
import json
import sys
from pathlib import Path

import mpmath as mp
import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
from benchdata import synth  # noqa: E402
from make_highprec import score  # noqa: E402

mp.mp.dps = 60


def checksum(*clouds):
    h = hashlib.sha256()
    for c in clouds:
        h.update(np.ascontiguousarray(c.xyz).tobytes()); h.update(c.cat.tobytes()); h.update(c.tag.tobytes())
    return h.hexdigest()


def main():
    base = synth.config5_base()
    members = {i: synth.config5_member(base, i) for i in (0, 1, 7)}
    rng = np.random.default_rng(55)
    anchors = np.sort(rng.choice(base.n, 32, replace=False)).astype(np.uint32)
    out = {}
    meta = {"provenance": "60-digit mpmath evaluation of the definition (tests/golden/make_highprec_cfg5.py)",
            "wf": ["uniform", [3.0, 10.0]], "sd": ["Hellinger", [2.0]], "C": 7, "threshold": 10.0, "accept_same": False,
            "pairs": [[0, 1], [0, 7], [1, 7]], "sha256": checksum(*members.values()), "events": {}}
    for i, j in meta["pairs"]:
        a, b = members[i], members[j]
        vals, ev = [], []
        for p in anchors:
            v, e = score((a.xyz, a.cat, a.tag), (b.xyz, b.cat, b.tag), (int(p), int(p)), 10.0, ("uniform", [3.0, 10.0]),
                         ("Hellinger", [2.0]), np.ones(7), False)
            vals.append(v); ev.append(e)
        out[f"truth_{i}_{j}"] = np.array([float(v) for v in vals])
        meta["events"][f"{i}_{j}"] = ev
        meta[f"truth_str_{i}_{j}"] = [mp.nstr(v, 30) for v in vals]
        print(i, j, "events", min(ev), max(ev), "first score", mp.nstr(vals[0], 25), flush=True)
    np.savez_compressed(HERE / "highprec_cfg5.npz", meta=np.array(json.dumps(meta)), anchors=anchors, **out)
    print("wrote", HERE / "highprec_cfg5.npz")


if __name__ == "__main__":
    main()
