"""Builds tests/golden/highprec.npz: LoCoHD scores evaluated in 60-digit arithmetic (mpmath) straight from the
mathematical definition the reference implements — NOT through oracle/ and not through its Python twin.

Why: the reference holds known answers to 4 decimal places only (tests/test_locohd.py:27-52,
tests/test_tag_pairing_rule.py:100-157), its golden outputs are missing upstream, and the Rust crate cannot be built
here.  These vectors pin the *numerical* side of the 1e-9 parity bar independently of any f64 implementation: a value
here is the exact real-number result of

    score = sum_k (W(t_k) - W(t_{k-1})) * SD(P_A, P_B before event k) + (W(inf) - W(t_K)) * SD(final)

(src/locohd.rs:61-226 as a flat scan over the merged neighbour events) with
* environments = anchor + {p : |x_p - x_a|_inf <= r, |x_p - x_a|^2 < r^2, rule(tag_a, tag_p)} (src/locohd.rs:514-542),
  evaluated exactly (the f64 inputs are exact rationals); the generator refuses a case in which a primitive sits
  within 1e-9 of the sphere, so no rounding of the membership test can matter;
* W = the four CDFs of src/locohd/weight_function/cdfs.rs:5-63, SD = the four statistical distances of
  src/locohd/pmf/statistical_distances.rs:4-78, compositions weighted by category_weights (src/locohd/pmf.rs:44-51).

Deterministic (numpy PCG64 seeds); needs only numpy + mpmath:  python tests/golden/make_highprec.py
"""
import json
from pathlib import Path

import mpmath as mp
import numpy as np

mp.mp.dps = 60
OUT = Path(__file__).resolve().parent / "highprec.npz"


def cdf(name, p, x):
    """x: mpf or mp.inf"""
    p = [mp.mpf(v) for v in p]
    if name == "hyper_exp":
        h = len(p) // 2
        if x == mp.inf:
            return mp.mpf(1)
        return 1 - sum(a * mp.e ** (-b * x) for a, b in zip(p[:h], p[h:])) / sum(p[:h])
    if name == "dagum":
        if x == 0:
            return mp.mpf(0)
        if x == mp.inf:
            return mp.mpf(1)
        return (1 + (x / p[1]) ** (-p[0])) ** (-p[2])
    if x < p[0]:
        return mp.mpf(0)
    if x > p[1]:
        return mp.mpf(1)
    z = (x - p[0]) / (p[1] - p[0])
    if name == "uniform":
        return z
    return 1 - (1 - z ** p[2]) ** p[3]


def sdist(name, q, a, b):
    """a, b: weighted counts (mpf lists), not yet normalised"""
    na, nb = sum(a), sum(b)
    p1, p2 = [x / na for x in a], [x / nb for x in b]
    q = [mp.mpf(v) for v in q]
    if name == "Hellinger":
        e = q[0]
        return (sum(abs(x ** (1 / e) - y ** (1 / e)) ** e for x, y in zip(p1, p2)) / 2) ** (1 / e)
    if name == "Kolmogorov-Smirnov":
        return max(abs(x - y) for x, y in zip(p1, p2))
    if name == "Kullback-Leibler":
        return sum(x * mp.log((x + q[0]) / (y + q[0])) for x, y in zip(p1, p2))
    alpha, eps = q
    if alpha == 1:        # statistical_distances.rs:36-38: the Kullback-Leibler limit
        return sum(x * mp.log((x + eps) / (y + eps)) for x, y in zip(p1, p2))
    if alpha == mp.inf:   # :42-51: log of the largest probability ratio
        return mp.log(max((x + eps) / (y + eps) for x, y in zip(p1, p2)))
    if alpha == 0:        # :55-64: minus log of the mass q puts where p is positive
        return -mp.log(sum(y for x, y in zip(p1, p2) if x > 0))
    return mp.log(sum(x * ((x + eps) / (y + eps)) ** (alpha - 1) for x, y in zip(p1, p2))) / (alpha - 1)


def environment(xyz, tag, anchor, r, accept_same):
    """[(d2 exact, index)] of the non-anchor members, plus a check that nothing sits on the sphere"""
    q = [mp.mpf(float(v)) for v in xyz[anchor]]
    r = mp.mpf(r)
    out = []
    for i in range(len(xyz)):
        if i == anchor:
            continue
        if accept_same is not None and (tag[i] == tag[anchor]) != accept_same:
            continue
        d = [mp.mpf(float(v)) - c for v, c in zip(xyz[i], q)]
        d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        assert abs(mp.sqrt(d2) - r) > mp.mpf("1e-9"), "primitive on the sphere: pick another seed"
        if d2 < r * r:
            out.append((d2, i))
    return out


def score(A, B, anchor, r, wf, sd, weights, accept_same):
    (xa, ca, ta), (xb, cb, tb) = A, B
    w = [mp.mpf(float(v)) for v in weights]
    a, b = [mp.mpf(0)] * len(w), [mp.mpf(0)] * len(w)
    a[ca[anchor[0]]] += w[ca[anchor[0]]]
    b[cb[anchor[1]]] += w[cb[anchor[1]]]
    ev = [(d2, 0, int(ca[i])) for d2, i in environment(xa, ta, anchor[0], r, accept_same)]
    ev += [(d2, 1, int(cb[i])) for d2, i in environment(xb, tb, anchor[1], r, accept_same)]
    ev.sort(key=lambda e: (e[0], e[1]))
    total, wprev = mp.mpf(0), cdf(*wf, mp.mpf(0))
    for d2, side, c in ev:
        wv = cdf(*wf, mp.sqrt(d2))
        total += (wv - wprev) * sdist(*sd, a, b)   # also when the weight is 0: 0 * inf = nan, as in the reference (locohd.rs:127-129)
        wprev = wv
        (a if side == 0 else b)[c] += w[c]
    return total + (cdf(*wf, mp.inf) - wprev) * sdist(*sd, a, b), len(ev)


def score_lists(sa, sb, da, db, wf, sd, weights):
    """from_anchors (src/locohd.rs:392-406): caller-sorted category / distance lists, member 0 = anchor"""
    w = [mp.mpf(float(v)) for v in weights]
    a, b = [mp.mpf(0)] * len(w), [mp.mpf(0)] * len(w)
    a[sa[0]] += w[sa[0]]
    b[sb[0]] += w[sb[0]]
    ev = [(mp.mpf(d), 0, c) for d, c in zip(da[1:], sa[1:])] + [(mp.mpf(d), 1, c) for d, c in zip(db[1:], sb[1:])]
    ev.sort(key=lambda e: (e[0], e[1]))
    total, wprev = mp.mpf(0), cdf(*wf, mp.mpf(0))
    for d, side, c in ev:
        wv = cdf(*wf, d)
        total += (wv - wprev) * sdist(*sd, a, b)
        wprev = wv
        (a if side == 0 else b)[c] += w[c]
    return total + (cdf(*wf, mp.inf) - wprev) * sdist(*sd, a, b)


def reference_kats():
    """The known answers the reference's own tests hold for the path, re-evaluated at 60 digits: the reference states
    them to 4 decimal places (tests/test_locohd.py:27-52, tests/test_tag_pairing_rule.py:100-157)."""
    h2 = ("Hellinger", [2.0])
    out = []
    wf = ("uniform", [0.0, 4.0])
    seq = [0, 1, 2, 3]
    out.append({"kind": "anchors", "wf": wf, "C": 4, "seq_a": seq, "seq_b": seq, "d_a": [0., 1., 2., 3.], "d_b": [0., 1., 1., 1.],
                "stated": 0.2268, "exact": score_lists(seq, seq, [0., 1., 2., 3.], [0., 1., 1., 1.], wf, h2, [1] * 4)})
    out.append({"kind": "anchors", "wf": wf, "C": 4, "seq_a": seq, "seq_b": seq, "d_a": [0., 1., 1., 1.], "d_b": [0., 1., 2., 3.],
                "stated": 0.2268, "exact": score_lists(seq, seq, [0., 1., 1., 1.], [0., 1., 2., 3.], wf, h2, [1] * 4)})
    wf = ("kumaraswamy", [3.0, 10.0, 2.0, 5.0])
    out.append({"kind": "anchors", "wf": wf, "C": 3, "seq_a": [0, 1, 0, 2], "seq_b": [0, 2], "d_a": [0., 1., 5., 9.], "d_b": [0., 7.],
                "stated": 0.4979, "exact": score_lists([0, 1, 0, 2], [0, 2], [0., 1., 5., 9.], [0., 7.], wf, h2, [1] * 3)})
    xyz = np.array([[0, 0, 0], [0, 1, 0], [2, 0, 0], [2, 2, 0], [1, 2, 0], [1, 3, 0], [3, 2, 0], [3, 3, 0], [2, 1, 0]], dtype=np.float64)
    cat = np.array([0, 0, 0, 0, 1, 1, 1, 1, 2], dtype=np.uint16)
    S = (xyz, cat, cat.astype(np.uint32))
    wf = ("uniform", [1.0, 1.001])
    for same, stated in ((True, [0., 0., 1., 1., 1.]), (False, [0.7071, 0.5412, 0.5412, 0.4284, 0.6501])):
        for an, st in zip([(0, 3), (4, 5), (0, 4), (0, 8), (4, 8)], stated):
            v, _ = score(S, S, an, 1.002, wf, h2, [1] * 3, same)
            out.append({"kind": "primitives", "wf": wf, "C": 3, "accept_same": same, "anchor": list(an), "threshold": 1.002,
                        "stated": st, "exact": v})
    for k in out:
        k["exact_str"] = mp.nstr(k.pop("exact"), 30)
        print("KAT", k["kind"], k["stated"], k["exact_str"], flush=True)
    return out


WFS = [("uniform", [3.0, 10.0]), ("kumaraswamy", [3.0, 10.0, 2.0, 5.0]), ("hyper_exp", [0.5, 0.5, 0.5, 1.0 / 3.0]),
       ("dagum", [2.0, 5.0, 1.0]), ("kumaraswamy", [0.0, 12.0, 1.5, 2.5]), ("hyper_exp", [3.0, 5.0, 2.0, 1 / 3.0, 0.2, 0.1])]
SDS = [("Hellinger", [2.0]), ("Hellinger", [3.5]), ("Kolmogorov-Smirnov", []), ("Kullback-Leibler", [0.01]),
       ("Renyi", [2.0, 0.05]), ("Renyi", [0.5, 0.001])]


def main():
    cases, arrays = [], {}
    k = 0
    # cases 36...: the special branches of the Renyi divergence (alpha = 1, +inf, 0); CPU checks only
    special = [(WFS[0], ("Renyi", [1.0, 0.02])), (WFS[3], ("Renyi", [1.0, 0.02])), (WFS[0], ("Renyi", [float("inf"), 0.02])),
               (WFS[2], ("Renyi", [float("inf"), 0.02])), (WFS[0], ("Renyi", [0.0, 0.0])), (WFS[1], ("Renyi", [0.0, 0.0]))]
    for wf, sd in [(wf, sd) for wf in WFS for sd in SDS] + special:
        for _ in (0,):
            # Hellinger-2 (the TMA-staged / tile kernels' arithmetic) gets the dense, protein-like cases
            rng = np.random.default_rng(7000 + k)
            C = int(rng.integers(3, 9))
            dense = sd == ("Hellinger", [2.0])
            n = 900 if dense else 220
            extent = 17.0 if dense else 13.0
            f32 = bool(k % 3 == 0)
            weights = np.ones(C) if k % 2 == 0 else np.round(rng.uniform(0.5, 2.0, C), 3)
            accept_same = [None, False, True][k % 3] if not dense else False
            structs = []
            base = rng.uniform(-extent, extent, (n, 3))
            for s in range(2):
                xyz = base + rng.normal(0, 0.8 if s else 0.0, (n, 3))
                if f32:
                    xyz = xyz.astype(np.float32).astype(np.float64)
                cat = rng.integers(0, C, n).astype(np.uint16) if s == 0 else structs[0][1].copy()
                if s == 1:   # a few substitutions, as between two models of one protein
                    m = rng.random(n) < 0.15
                    cat[m] = rng.integers(0, C, int(m.sum()))
                tag = (np.arange(n) // (60 if accept_same else 4)).astype(np.uint32)   # 'same tag only' needs big groups
                structs.append((xyz, cat, tag))
            anchors = np.stack([rng.choice(n, 10, replace=False)] * 2, axis=1).astype(np.uint32)
            thr = 10.0 if dense else 9.0
            vals, sizes = [], []
            for an in anchors:
                v, e = score(structs[0], structs[1], (int(an[0]), int(an[1])), thr, wf, sd, weights, accept_same)
                vals.append(v)
                sizes.append(e)
            for s in range(2):
                arrays[f"xyz_{k}_{s}"], arrays[f"cat_{k}_{s}"], arrays[f"tag_{k}_{s}"] = structs[s]
            arrays[f"anchors_{k}"] = anchors
            arrays[f"truth_{k}"] = np.array([float(v) for v in vals])
            cases.append({"id": k, "wf": wf, "sd": sd, "C": C, "weights": [float(x) for x in weights],
                          "accept_same": accept_same, "threshold": thr, "f32_exact": f32, "events": sizes,
                          "truth_str": [mp.nstr(v, 30) for v in vals]})
            print(k, wf[0], sd, "events", sizes, "score", mp.nstr(vals[0], 20), flush=True)
            k += 1
    meta = {"provenance": "60-digit mpmath evaluation of the definition (tests/golden/make_highprec.py); independent "
                          "of oracle/ and of the CUDA path", "dps": mp.mp.dps, "cases": cases,
            "reference_kats": reference_kats()}
    np.savez_compressed(OUT, meta=np.array(json.dumps(meta)), **arrays)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
