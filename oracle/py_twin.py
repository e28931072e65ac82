"""oracle/py_twin.py — TEST INFRASTRUCTURE.  A small pure-Python twin of the merge walk, written independently of
the C++ restatement, used only to cross-check it on small cases (tests/test_oracle_kat.py).

``stat_dist_integral`` follows /root/reference/src/locohd.rs:61-226 branch by branch; ``flat_scan`` is the
prefix-scan formulation the CUDA kernel uses: sum over all non-anchor members of both environments
sorted together of (W(t_k) - W(t_{k-1})) * H(state before event k), plus the tail to infinity.
"""
import math


def cdf(wf, x):
    name, p = wf
    if name == "hyper_exp":  # cdfs.rs:5-21
        h = len(p) // 2
        return 1.0 - sum(a * math.exp(-b * x) for a, b in zip(p[:h], p[h:])) / sum(p[:h])
    if name == "dagum":  # cdfs.rs:27-29
        inner = math.inf if x == 0 else (x / p[1]) ** (-p[0])
        return 0.0 if inner == math.inf else (1.0 + inner) ** (-p[2])
    if x < p[0]:
        return 0.0
    if x > p[1]:
        return 1.0
    z = (x - p[0]) / (p[1] - p[0])
    if name == "uniform":  # cdfs.rs:39-45
        return z
    return 1.0 - (1.0 - z ** p[2]) ** p[3]  # kumaraswamy, cdfs.rs:56-63


def distance(sd, p1, p2):
    name, q = sd
    if name == "Hellinger":  # statistical_distances.rs:4-10
        e = q[0]
        return (sum(abs(x ** (1 / e) - y ** (1 / e)) ** e for x, y in zip(p1, p2)) / 2.0) ** (1 / e)
    if name == "Kolmogorov-Smirnov":
        return max(abs(x - y) for x, y in zip(p1, p2))
    if name == "Kullback-Leibler":
        return sum(x * math.log((x + q[0]) / (y + q[0])) for x, y in zip(p1, p2))
    a, eps = q  # Renyi, general alpha only
    return math.log(sum(x * ((x + eps) / (y + eps)) ** (a - 1.0) for x, y in zip(p1, p2))) / (a - 1.0)


class _Pmf:
    def __init__(self, w):
        self.w, self.a, self.b = w, [0.0] * len(w), [0.0] * len(w)

    def h(self, sd):  # pmf.rs:65-88
        na, nb = sum(self.a), sum(self.b)
        return distance(sd, [x / na for x in self.a], [x / nb for x in self.b])


def stat_dist_integral(sa, sb, da, db, wf, sd, w):
    pm = _Pmf(w)
    pm.a[sa[0]] += w[sa[0]]
    pm.b[sb[0]] += w[sb[0]]
    ia = ib = 0
    total, buf = 0.0, 0.0
    rng = lambda lo, hi: cdf(wf, hi) - cdf(wf, lo)
    while ia < len(sa) - 1 and ib < len(sb) - 1:
        h = pm.h(sd)
        if da[ia + 1] < db[ib + 1]:
            ia += 1; pm.a[sa[ia]] += w[sa[ia]]; new = da[ia]
        elif da[ia + 1] > db[ib + 1]:
            ib += 1; pm.b[sb[ib]] += w[sb[ib]]; new = db[ib]
        else:
            ia += 1; ib += 1
            pm.a[sa[ia]] += w[sa[ia]]; pm.b[sb[ib]] += w[sb[ib]]; new = da[ia]
        total += rng(buf, new) * h
        buf = new
    if ib < len(sb) - 1:
        h = pm.h(sd); ib += 1
        total += rng(da[-1], db[ib]) * h
        pm.b[sb[ib]] += w[sb[ib]]
        while ib < len(sb) - 1:
            ib += 1
            total += rng(db[ib - 1], db[ib]) * pm.h(sd)
            pm.b[sb[ib]] += w[sb[ib]]
        total += rng(db[-1], math.inf) * pm.h(sd)
    elif ia < len(sa) - 1:
        h = pm.h(sd); ia += 1
        total += rng(db[-1], da[ia]) * h
        pm.a[sa[ia]] += w[sa[ia]]
        while ia < len(sa) - 1:
            ia += 1
            total += rng(da[ia - 1], da[ia]) * pm.h(sd)
            pm.a[sa[ia]] += w[sa[ia]]
        total += rng(da[-1], math.inf) * pm.h(sd)
    else:
        total += rng(da[-1], math.inf) * pm.h(sd)
    return total


def flat_scan(sa, sb, da, db, wf, sd, w):
    pm = _Pmf(w)
    pm.a[sa[0]] += w[sa[0]]
    pm.b[sb[0]] += w[sb[0]]
    events = [(d, 0, c) for d, c in zip(da[1:], sa[1:])] + [(d, 1, c) for d, c in zip(db[1:], sb[1:])]
    events.sort(key=lambda e: (e[0], e[1]))
    total, wprev = 0.0, cdf(wf, 0.0)
    for d, side, c in events:
        wv = cdf(wf, d)
        total += (wv - wprev) * pm.h(sd)
        wprev = wv
        (pm.a if side == 0 else pm.b)[c] += w[c]
    return total + (cdf(wf, math.inf) - wprev) * pm.h(sd)
