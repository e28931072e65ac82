// ============================================================================
// oracle/locohd_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17, optional OpenMP) of the reference's per-anchor
// LoCoHD scoring path. It exists to CHECK the CUDA path (tests/, smoke(),
// bench.py's cpu_baseline / --impl reference legs). Nothing under
// loco_hd_b200/ or loco_hd/ may import, link or call it.
//
// Parity status: the reference (Rust/PyO3) cannot be compiled in this image
// (no cargo/rustc), so this restatement is pinned against the reference's own
// known-answer tests only:
//   tests/test_locohd.py:27-52, tests/test_tag_pairing_rule.py:8-157,
//   tests/test_wfs.py:8-156 (see tests/test_oracle_kat.py here).
// The golden outputs of tests/test_locohd.py:75-133 are missing upstream
// (.MISSING_LARGE_BLOBS), and the `kd-tree` crate (Cargo.toml:18-19,
// version "0.6.0", source not vendored) decides neighbour membership, so the
// strict-vs-inclusive radius boundary is "parity unpinned" (see DESIGN.md).
//
// Every function cites the reference file:line it follows.
// All categories/tags are interned integer ids (the host interns strings):
//   category id in [0, C) ; LOCOHD_ORACLE_UNKNOWN_CAT marks an unknown name.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr uint16_t kUnknownCat = 0xFFFF;

enum Status : int {
    OK = 0,
    ERR_LEN_MISMATCH = 1,       // locohd.rs:70-73
    ERR_FIRST_NOT_ZERO = 2,     // locohd.rs:74-77
    ERR_UNKNOWN_CATEGORY = 3,   // pmf.rs:38-42
    ERR_ZERO_NORM = 4,          // pmf.rs:70-76
    ERR_NEGATIVE_POINT = 5,     // weight_function.rs:97-100
    ERR_NAN = 6,                // partial_cmp().unwrap() panic (utils.rs:28) / unreachable!() (locohd.rs:124)
    ERR_EMPTY_ENV = 7,          // dists_a[0] on an empty Vec panics (locohd.rs:74)
    ERR_INDEX = 8,              // prim_seq[anchor_idx] out of bounds panics (locohd.rs:521)
    ERR_DMX_SHAPE = 9,          // locohd.rs:420-428
    ERR_BAD_PARAM = 10
};

enum WfKind : int { WF_HYPER_EXP = 0, WF_DAGUM = 1, WF_UNIFORM = 2, WF_KUMARASWAMY = 3 };
enum SdKind : int { SD_HELLINGER = 0, SD_KS = 1, SD_KL = 2, SD_RENYI = 3 };
enum TprKind : int { TPR_WITHOUT_LIST = 0, TPR_WITH_LIST = 1 };

struct WeightFn {
    int kind;
    int n;
    const double* p;
};

// weight_function/cdfs.rs:5-21
double cdf_hyper_exp(const WeightFn& w, double x) {
    double norm = 0.0, sum = 0.0;
    const int half = w.n / 2;
    for (int i = 0; i < half; ++i) {
        sum += w.p[i] * std::exp(-w.p[half + i] * x);
        norm += w.p[i];
    }
    return 1.0 - sum / norm;
}
// weight_function/cdfs.rs:27-29
double cdf_dagum(const WeightFn& w, double x) {
    return std::pow(1.0 + std::pow(x / w.p[1], -w.p[0]), -w.p[2]);
}
// weight_function/cdfs.rs:39-45
double cdf_uniform(const WeightFn& w, double x) {
    if (x < w.p[0]) return 0.0;
    if (x > w.p[1]) return 1.0;
    return (x - w.p[0]) / (w.p[1] - w.p[0]);
}
// weight_function/cdfs.rs:56-63
double cdf_kumaraswamy(const WeightFn& w, double x) {
    if (x < w.p[0]) return 0.0;
    if (x > w.p[1]) return 1.0;
    const double z = (x - w.p[0]) / (w.p[1] - w.p[0]);
    return 1.0 - std::pow(1.0 - std::pow(z, w.p[2]), w.p[3]);
}

// weight_function.rs:95-103 (integral_point)
int integral_point(const WeightFn& w, double x, double* out) {
    if (x < 0.0) return ERR_NEGATIVE_POINT;
    switch (w.kind) {
        case WF_HYPER_EXP: *out = cdf_hyper_exp(w, x); break;
        case WF_DAGUM: *out = cdf_dagum(w, x); break;
        case WF_UNIFORM: *out = cdf_uniform(w, x); break;
        case WF_KUMARASWAMY: *out = cdf_kumaraswamy(w, x); break;
        default: return ERR_BAD_PARAM;
    }
    return OK;
}
// weight_function.rs:118-120 (integral_range = CDF(to) - CDF(from); `to` is evaluated first)
int integral_range(const WeightFn& w, double from, double to, double* out) {
    double a, b;
    int st = integral_point(w, to, &b);
    if (st) return st;
    st = integral_point(w, from, &a);
    if (st) return st;
    *out = b - a;
    return OK;
}

struct StatDist {
    int kind;
    double p[2];
};

// pmf/statistical_distances.rs:4-10
double sd_hellinger(const double* p1, const double* p2, int C, double e) {
    double dist = 0.0;
    for (int i = 0; i < C; ++i)
        dist += std::pow(std::fabs(std::pow(p1[i], 1.0 / e) - std::pow(p2[i], 1.0 / e)), e);
    return std::pow(dist / 2.0, 1.0 / e);
}
// pmf/statistical_distances.rs:12-21 (max_by partial_cmp; NaN would panic)
int sd_ks(const double* p1, const double* p2, int C, double* out) {
    double best = 0.0;
    bool first = true;
    for (int i = 0; i < C; ++i) {
        const double d = std::fabs(p1[i] - p2[i]);
        if (std::isnan(d)) return ERR_NAN;
        if (first || !(d < best)) { best = d; first = false; }  // max_by keeps the last maximum
    }
    *out = best;
    return OK;
}
// pmf/statistical_distances.rs:23-29
double sd_kl(const double* p1, const double* p2, int C, double eps) {
    double dist = 0.0;
    for (int i = 0; i < C; ++i) dist += p1[i] * std::log((p1[i] + eps) / (p2[i] + eps));
    return dist;
}
// pmf/statistical_distances.rs:31-78
int sd_renyi(const double* p1, const double* p2, int C, double alpha, double eps, double* out) {
    if (alpha == 1.0) { *out = sd_kl(p1, p2, C, eps); return OK; }
    if (alpha == std::numeric_limits<double>::infinity()) {
        double best = 0.0;
        bool first = true;
        for (int i = 0; i < C; ++i) {
            const double r = (p1[i] + eps) / (p2[i] + eps);
            if (std::isnan(r)) return ERR_NAN;
            if (first || !(r < best)) { best = r; first = false; }
        }
        *out = std::log(best);
        return OK;
    }
    if (alpha == 0.0) {
        double s = 0.0;
        for (int i = 0; i < C; ++i) if (p1[i] > 0.0) s += p2[i];
        *out = -std::log(s);
        return OK;
    }
    double s = 0.0;
    for (int i = 0; i < C; ++i) s += p1[i] * std::pow((p1[i] + eps) / (p2[i] + eps), alpha - 1.0);
    *out = std::log(s) / (alpha - 1.0);
    return OK;
}
// pmf/statistical_distances.rs:123-142 (run)
int sd_run(const StatDist& sd, const double* p1, const double* p2, int C, double* out) {
    switch (sd.kind) {
        case SD_HELLINGER: *out = sd_hellinger(p1, p2, C, sd.p[0]); return OK;
        case SD_KS: return sd_ks(p1, p2, C, out);
        case SD_KL: *out = sd_kl(p1, p2, C, sd.p[0]); return OK;
        case SD_RENYI: return sd_renyi(p1, p2, C, sd.p[0], sd.p[1], out);
        default: return ERR_BAD_PARAM;
    }
}

// pmf.rs:13-89 (PMFSystem)
struct PmfSystem {
    int C;
    const double* w;
    std::vector<double> pmf1, pmf2, n1, n2;
    PmfSystem(int C_, const double* w_) : C(C_), w(w_), pmf1(C_, 0.0), pmf2(C_, 0.0), n1(C_), n2(C_) {}
    int update1(uint16_t c) {  // pmf.rs:47-54
        if (c >= C) return ERR_UNKNOWN_CATEGORY;
        pmf1[c] += w[c];
        return OK;
    }
    int update2(uint16_t c) {  // pmf.rs:56-63
        if (c >= C) return ERR_UNKNOWN_CATEGORY;
        pmf2[c] += w[c];
        return OK;
    }
    int distance(const StatDist& sd, double* out) {  // pmf.rs:65-88
        double norm1 = 0.0, norm2 = 0.0;
        for (int i = 0; i < C; ++i) norm1 += pmf1[i];
        for (int i = 0; i < C; ++i) norm2 += pmf2[i];
        if (norm1 == 0.0 || norm2 == 0.0) return ERR_ZERO_NORM;
        for (int i = 0; i < C; ++i) n1[i] = pmf1[i] / norm1;
        for (int i = 0; i < C; ++i) n2[i] = pmf2[i] / norm2;
        return sd_run(sd, n1.data(), n2.data(), C, out);
    }
};

struct Model {
    int C;
    const double* cat_w;
    StatDist sd;
};

#define TRY(expr) do { int st__ = (expr); if (st__) return st__; } while (0)

// locohd.rs:61-226 (stat_dist_integral) — the three-way merge walk, statement for statement.
int stat_dist_integral(const Model& m, const uint16_t* seq_a, size_t len_a, const uint16_t* seq_b,
                       size_t len_b, const double* da, size_t dlen_a, const double* db, size_t dlen_b,
                       const WeightFn& wf, double* out, uint64_t* n_steps) {
    if (len_a != dlen_a || len_b != dlen_b) return ERR_LEN_MISMATCH;
    if (len_a == 0 || len_b == 0) return ERR_EMPTY_ENV;
    if (da[0] != 0.0 || db[0] != 0.0) return ERR_FIRST_NOT_ZERO;

    PmfSystem pmf(m.C, m.cat_w);
    TRY(pmf.update1(seq_a[0]));
    TRY(pmf.update2(seq_b[0]));

    size_t ia = 0, ib = 0;
    double integral = 0.0, buffer = 0.0, h, dw;
    uint64_t steps = 0;

    while (ia < len_a - 1 && ib < len_b - 1) {
        TRY(pmf.distance(m.sd, &h));
        double new_dist;
        if (da[ia + 1] < db[ib + 1]) {
            ++ia; TRY(pmf.update1(seq_a[ia])); new_dist = da[ia];
        } else if (da[ia + 1] > db[ib + 1]) {
            ++ib; TRY(pmf.update2(seq_b[ib])); new_dist = db[ib];
        } else if (da[ia + 1] == db[ib + 1]) {
            ++ia; ++ib;
            TRY(pmf.update1(seq_a[ia])); TRY(pmf.update2(seq_b[ib]));
            new_dist = da[ia];
        } else {
            return ERR_NAN;  // unreachable!() in the reference
        }
        TRY(integral_range(wf, buffer, new_dist, &dw));
        integral += dw * h;
        buffer = new_dist;
        ++steps;
    }

    if (ib < len_b - 1) {
        TRY(pmf.distance(m.sd, &h));
        ++ib;
        TRY(integral_range(wf, da[len_a - 1], db[ib], &dw));
        integral += dw * h; ++steps;
        TRY(pmf.update2(seq_b[ib]));
        while (ib < len_b - 1) {
            ++ib;
            TRY(pmf.distance(m.sd, &h));
            TRY(integral_range(wf, db[ib - 1], db[ib], &dw));
            integral += dw * h; ++steps;
            TRY(pmf.update2(seq_b[ib]));
        }
        TRY(pmf.distance(m.sd, &h));
        TRY(integral_range(wf, db[len_b - 1], std::numeric_limits<double>::infinity(), &dw));
        integral += dw * h; ++steps;
    } else if (ia < len_a - 1) {
        TRY(pmf.distance(m.sd, &h));
        ++ia;
        TRY(integral_range(wf, db[len_b - 1], da[ia], &dw));
        integral += dw * h; ++steps;
        TRY(pmf.update1(seq_a[ia]));
        while (ia < len_a - 1) {
            ++ia;
            TRY(pmf.distance(m.sd, &h));
            TRY(integral_range(wf, da[ia - 1], da[ia], &dw));
            integral += dw * h; ++steps;
            TRY(pmf.update1(seq_a[ia]));
        }
        TRY(pmf.distance(m.sd, &h));
        TRY(integral_range(wf, da[len_a - 1], std::numeric_limits<double>::infinity(), &dw));
        integral += dw * h; ++steps;
    } else {
        TRY(pmf.distance(m.sd, &h));
        TRY(integral_range(wf, da[len_a - 1], std::numeric_limits<double>::infinity(), &dw));
        integral += dw * h; ++steps;
    }
    *out = integral;
    if (n_steps) *n_steps = steps;
    return OK;
}

// utils.rs:1-8 (powf(2.) is x*x, powf(.5) is sqrt after LLVM's libcall simplification)
inline double euclidean_distance(const double* a, const double* b) {
    double distance = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double d = a[k] - b[k];
        distance += d * d;
    }
    return std::sqrt(distance);
}

// utils.rs:25-39 (stable index sort; NaN => panic)
int sort_together(const std::vector<double>& dists, const std::vector<uint16_t>& cats,
                  std::vector<double>& out_d, std::vector<uint16_t>& out_c, std::vector<uint32_t>* perm) {
    for (double d : dists) if (std::isnan(d)) return ERR_NAN;
    std::vector<uint32_t> mask(dists.size());
    std::iota(mask.begin(), mask.end(), 0u);
    std::stable_sort(mask.begin(), mask.end(), [&](uint32_t i, uint32_t j) { return dists[i] < dists[j]; });
    out_d.resize(dists.size());
    out_c.resize(dists.size());
    for (size_t k = 0; k < mask.size(); ++k) { out_d[k] = dists[mask[k]]; out_c[k] = cats[mask[k]]; }
    if (perm) *perm = std::move(mask);
    return OK;
}

struct TagRule {
    int kind;            // TprKind
    int accept_same;     // WithoutList
    int accepted_pairs;  // WithList
    int ordered;         // WithList
    size_t n_pairs;
    const uint64_t* pairs;  // sorted, (anchor_tag << 32) | neighbour_tag
    bool contains(uint32_t a, uint32_t b) const {
        const uint64_t key = (uint64_t(a) << 32) | b;
        return std::binary_search(pairs, pairs + n_pairs, key);
    }
    // tag_pairing_rule.rs:49-76
    bool accepted(uint32_t anchor_tag, uint32_t other_tag) const {
        if (kind == TPR_WITHOUT_LIST) {
            bool acc = anchor_tag == other_tag;
            if (!accept_same) acc = !acc;
            return acc;
        }
        bool acc = contains(anchor_tag, other_tag);
        if (!ordered) acc |= contains(other_tag, anchor_tag);
        if (!accepted_pairs) acc = !acc;
        return acc;
    }
};

// A 3-d tree with the layout and query semantics of the `kd-tree` crate 0.6
// (KdTree::build_by_ordered_float, locohd.rs:504-510; within_radius, locohd.rs:521):
// implicit balanced tree (median at len/2, axis cycling), box query
// `!(x_k < q_k - r) && !(x_k > q_k + r)` on every axis, then squared distance
// `((dx*dx + dy*dy) + dz*dz) < r*r` (strict). Restated from the crate's published
// algorithm; the crate source is not available offline (parity unpinned at the boundary).
struct KdTree3 {
    const double* xyz;
    std::vector<uint32_t> items;
    void build(const double* xyz_, size_t n) {
        xyz = xyz_;
        items.resize(n);
        std::iota(items.begin(), items.end(), 0u);
        sort_range(0, n, 0);
    }
    void sort_range(size_t lo, size_t hi, int axis) {
        if (hi - lo < 2) return;
        const size_t mid = lo + (hi - lo) / 2;
        std::nth_element(items.begin() + lo, items.begin() + mid, items.begin() + hi,
                         [&](uint32_t a, uint32_t b) { return xyz[3 * a + axis] < xyz[3 * b + axis]; });
        const int next = (axis + 1) % 3;
        sort_range(lo, mid, next);
        sort_range(mid + 1, hi, next);
    }
    static int cmp_axis(double coord, double q, double r) {
        if (coord < q - r) return -1;
        if (coord > q + r) return 1;
        return 0;
    }
    void within_box(size_t lo, size_t hi, int axis, const double* q, double r, std::vector<uint32_t>& out) const {
        if (lo >= hi) return;
        const size_t mid = lo + (hi - lo) / 2;
        const uint32_t it = items[mid];
        const int c = cmp_axis(xyz[3 * it + axis], q[axis], r);
        const int next = (axis + 1) % 3;
        if (c == 0) {
            if (cmp_axis(xyz[3 * it + next], q[next], r) == 0 &&
                cmp_axis(xyz[3 * it + (axis + 2) % 3], q[(axis + 2) % 3], r) == 0)
                out.push_back(it);
            within_box(lo, mid, next, q, r, out);
            within_box(mid + 1, hi, next, q, r, out);
        } else if (c < 0) {
            within_box(mid + 1, hi, next, q, r, out);
        } else {
            within_box(lo, mid, next, q, r, out);
        }
    }
    void within_radius(const double* q, double r, std::vector<uint32_t>& out) const {
        out.clear();
        within_box(0, items.size(), 0, q, r, out);
        size_t w = 0;
        for (uint32_t it : out) {
            double distance = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double diff = xyz[3 * it + k] - q[k];
                distance += diff * diff;
            }
            if (distance < r * r) out[w++] = it;
        }
        out.resize(w);
    }
};

// Membership by exhaustive scan with the same predicate (cross-check of the tree).
void brute_within_radius(const double* xyz, size_t n, const double* q, double r, std::vector<uint32_t>& out) {
    out.clear();
    for (size_t i = 0; i < n; ++i) {
        bool in_box = true;
        for (int k = 0; k < 3; ++k)
            if (KdTree3::cmp_axis(xyz[3 * i + k], q[k], r) != 0) in_box = false;
        if (!in_box) continue;
        double distance = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double diff = xyz[3 * i + k] - q[k];
            distance += diff * diff;
        }
        if (distance < r * r) out.push_back(uint32_t(i));
    }
}

struct Structure {
    size_t n;
    const double* xyz;
    const uint16_t* cat;
    const uint32_t* tag;
};

// locohd.rs:514-542 (env_from_idx closure)
int env_from_idx(const Structure& s, const KdTree3* tree, size_t anchor, double thr, const TagRule& rule,
                 std::vector<double>& out_d, std::vector<uint16_t>& out_c, std::vector<uint32_t>* out_idx) {
    if (anchor >= s.n) return ERR_INDEX;
    const double* q = s.xyz + 3 * anchor;
    std::vector<uint32_t> nb;
    if (tree) tree->within_radius(q, thr, nb);
    else brute_within_radius(s.xyz, s.n, q, thr, nb);
    std::vector<double> d;
    std::vector<uint16_t> c;
    std::vector<uint32_t> kept;
    d.reserve(nb.size()); c.reserve(nb.size()); kept.reserve(nb.size());
    for (uint32_t p : nb) {
        bool acc = (p == anchor);                           // ptr::eq, locohd.rs:525
        acc |= rule.accepted(s.tag[anchor], s.tag[p]);      // locohd.rs:526
        if (!acc) continue;
        c.push_back(s.cat[p]);
        d.push_back(euclidean_distance(q, s.xyz + 3 * p));  // locohd.rs:537
        kept.push_back(p);
    }
    std::vector<uint32_t> perm;
    TRY(sort_together(d, c, out_d, out_c, &perm));
    if (out_idx) {
        out_idx->resize(perm.size());
        for (size_t k = 0; k < perm.size(); ++k) (*out_idx)[k] = kept[perm[k]];
    }
    return OK;
}

struct ParamPack {
    Model model;
    std::vector<WeightFn> wfs;
    TagRule rule;
};

}  // namespace

extern "C" {

// Flat parameter block shared by every entry point (mirrors LoCoHD::build state, locohd.rs:42-55).
struct oracle_params {
    int32_t n_categories;
    const double* category_weights;  // [n_categories]
    int32_t sd_kind;
    double sd_params[2];
    int32_t n_wf;
    const int32_t* wf_kind;     // [n_wf]
    const int32_t* wf_nparams;  // [n_wf]
    const int32_t* wf_offset;   // [n_wf] offset into wf_params
    const double* wf_params;
    int32_t tpr_kind;
    int32_t tpr_accept_same;
    int32_t tpr_accepted_pairs;
    int32_t tpr_ordered;
    uint64_t n_tag_pairs;
    const uint64_t* tag_pairs;  // sorted
};

static ParamPack unpack(const oracle_params* p) {
    ParamPack pk;
    pk.model.C = p->n_categories;
    pk.model.cat_w = p->category_weights;
    pk.model.sd.kind = p->sd_kind;
    pk.model.sd.p[0] = p->sd_params[0];
    pk.model.sd.p[1] = p->sd_params[1];
    for (int i = 0; i < p->n_wf; ++i)
        pk.wfs.push_back(WeightFn{p->wf_kind[i], p->wf_nparams[i], p->wf_params + p->wf_offset[i]});
    pk.rule = TagRule{p->tpr_kind, p->tpr_accept_same, p->tpr_accepted_pairs, p->tpr_ordered,
                      size_t(p->n_tag_pairs), p->tag_pairs};
    return pk;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// WeightFunction::integral_point (weight_function.rs:95-103)
int oracle_wf_integral_point(int kind, int n, const double* params, double x, double* out) {
    return integral_point(WeightFn{kind, n, params}, x, out);
}
// WeightFunction::integral_range (weight_function.rs:118-120)
int oracle_wf_integral_range(int kind, int n, const double* params, double from, double to, double* out) {
    return integral_range(WeightFn{kind, n, params}, from, to, out);
}
// StatisticalDistance::run (statistical_distances.rs:123-142)
int oracle_sd_run(int kind, const double* sd_params, int C, const double* p1, const double* p2, double* out) {
    StatDist sd{kind, {sd_params[0], sd_params[1]}};
    return sd_run(sd, p1, p2, C, out);
}
// TagPairingRule::pair_accepted (tag_pairing_rule.rs:49-76) on interned ids
int oracle_tag_pair_accepted(const oracle_params* p, uint32_t anchor_tag, uint32_t other_tag) {
    return unpack(p).rule.accepted(anchor_tag, other_tag) ? 1 : 0;
}

// LoCoHD::from_anchors (locohd.rs:392-406)
int oracle_from_anchors(const oracle_params* p, const uint16_t* seq_a, uint64_t len_a, const uint16_t* seq_b,
                        uint64_t len_b, const double* da, uint64_t dlen_a, const double* db, uint64_t dlen_b,
                        int wf_idx, double* out, uint64_t* n_steps) {
    ParamPack pk = unpack(p);
    if (wf_idx < 0 || wf_idx >= int(pk.wfs.size())) return ERR_BAD_PARAM;
    return stat_dist_integral(pk.model, seq_a, len_a, seq_b, len_b, da, dlen_a, db, dlen_b, pk.wfs[wf_idx], out,
                              n_steps);
}

// Environment of one anchor (locohd.rs:514-542). Returns M (or -status). Buffers must hold n entries.
// use_tree != 0 -> kd-tree query, else exhaustive scan.
int64_t oracle_environment(const oracle_params* p, uint64_t n, const double* xyz, const uint16_t* cat,
                           const uint32_t* tag, uint64_t anchor, double threshold, int use_tree,
                           uint32_t* out_idx, double* out_dist, uint16_t* out_cat) {
    ParamPack pk = unpack(p);
    Structure s{size_t(n), xyz, cat, tag};
    KdTree3 tree;
    if (use_tree) tree.build(xyz, n);
    std::vector<double> d;
    std::vector<uint16_t> c;
    std::vector<uint32_t> idx;
    int st = env_from_idx(s, use_tree ? &tree : nullptr, anchor, threshold, pk.rule, d, c, &idx);
    if (st) return -int64_t(st);
    for (size_t k = 0; k < d.size(); ++k) {
        if (out_idx) out_idx[k] = idx[k];
        if (out_dist) out_dist[k] = d[k];
        if (out_cat) out_cat[k] = c[k];
    }
    return int64_t(d.size());
}

// LoCoHD::from_primitives (locohd.rs:479-567). wf_idx may be NULL (single weight function 0).
// Optional per-anchor debug outputs: env sizes [P][2], final integer category counts [P][2][C]
// (number of environment members per category, anchors included), and walk step counts [P].
int oracle_from_primitives(const oracle_params* p, uint64_t na, const double* xyz_a, const uint16_t* cat_a,
                           const uint32_t* tag_a, uint64_t nb, const double* xyz_b, const uint16_t* cat_b,
                           const uint32_t* tag_b, uint64_t n_pairs, const uint32_t* anchors /*[P][2]*/,
                           const int32_t* wf_idx, double threshold, int use_tree, int n_threads,
                           double* out_scores, uint32_t* out_env_sizes, uint32_t* out_counts,
                           uint64_t* out_steps) {
    ParamPack pk = unpack(p);
    Structure sa{size_t(na), xyz_a, cat_a, tag_a}, sb{size_t(nb), xyz_b, cat_b, tag_b};
    KdTree3 ta, tb;
    if (use_tree) { ta.build(xyz_a, na); tb.build(xyz_b, nb); }  // locohd.rs:504-510
    int first_err = OK;
    const int C = pk.model.C;
#ifdef _OPENMP
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
#endif
    for (int64_t i = 0; i < int64_t(n_pairs); ++i) {  // locohd.rs:545-554
        std::vector<double> da, db;
        std::vector<uint16_t> ca, cb;
        int st = env_from_idx(sa, use_tree ? &ta : nullptr, anchors[2 * i], threshold, pk.rule, da, ca, nullptr);
        if (!st) st = env_from_idx(sb, use_tree ? &tb : nullptr, anchors[2 * i + 1], threshold, pk.rule, db, cb, nullptr);
        const int w = wf_idx ? wf_idx[i] : 0;
        if (!st && (w < 0 || w >= int(pk.wfs.size()))) st = ERR_BAD_PARAM;
        double score = 0.0;
        uint64_t steps = 0;
        if (!st) st = stat_dist_integral(pk.model, ca.data(), ca.size(), cb.data(), cb.size(), da.data(), da.size(),
                                         db.data(), db.size(), pk.wfs[w], &score, &steps);
        if (st) {
#ifdef _OPENMP
#pragma omp critical
#endif
            { if (!first_err) first_err = st; }
            continue;
        }
        out_scores[i] = score;
        if (out_env_sizes) { out_env_sizes[2 * i] = uint32_t(ca.size()); out_env_sizes[2 * i + 1] = uint32_t(cb.size()); }
        if (out_counts) {
            uint32_t* cnt = out_counts + size_t(i) * 2 * C;
            std::fill(cnt, cnt + 2 * C, 0u);
            for (uint16_t c : ca) cnt[c]++;
            for (uint16_t c : cb) cnt[C + c]++;
        }
        if (out_steps) out_steps[i] = steps;
    }
    return first_err;
}

// LoCoHD::from_dmxs (locohd.rs:410-458): dmx_* are row-major [n_rows][len]; every row is an anchor.
int oracle_from_dmxs(const oracle_params* p, const uint16_t* seq_a, uint64_t len_a, const uint16_t* seq_b,
                     uint64_t len_b, const double* dmx_a, uint64_t rows_a, uint64_t cols_a, const double* dmx_b,
                     uint64_t rows_b, uint64_t cols_b, const int32_t* wf_idx, int n_threads, double* out_scores) {
    if (rows_a != rows_b) return ERR_DMX_SHAPE;
    ParamPack pk = unpack(p);
    int first_err = OK;
#ifdef _OPENMP
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
#endif
    for (int64_t i = 0; i < int64_t(rows_a); ++i) {  // locohd.rs:434-443
        std::vector<double> ra(dmx_a + i * cols_a, dmx_a + (i + 1) * cols_a);
        std::vector<double> rb(dmx_b + i * cols_b, dmx_b + (i + 1) * cols_b);
        std::vector<uint16_t> sa(seq_a, seq_a + len_a), sb(seq_b, seq_b + len_b);
        std::vector<double> da, db;
        std::vector<uint16_t> ca, cb;
        int st = OK;
        // sort_together indexes cats by the dists' indices (utils.rs:33-36): a short seq panics,
        // surplus seq entries are never read.
        if (sa.size() < ra.size() || sb.size() < rb.size()) st = ERR_INDEX;
        sa.resize(ra.size()); sb.resize(rb.size());
        if (!st) st = sort_together(ra, sa, da, ca, nullptr);
        if (!st) st = sort_together(rb, sb, db, cb, nullptr);
        const int w = wf_idx ? wf_idx[i] : 0;
        if (!st && (w < 0 || w >= int(pk.wfs.size()))) st = ERR_BAD_PARAM;
        double score = 0.0;
        if (!st) st = stat_dist_integral(pk.model, ca.data(), ca.size(), cb.data(), cb.size(), da.data(), da.size(),
                                         db.data(), db.size(), pk.wfs[w], &score, nullptr);
        if (st) {
#ifdef _OPENMP
#pragma omp critical
#endif
            { if (!first_err) first_err = st; }
            continue;
        }
        out_scores[i] = score;
    }
    return first_err;
}

// utils.rs:10-22 (calculate_distance_matrix) + LoCoHD::from_coords (locohd.rs:463-476)
int oracle_from_coords(const oracle_params* p, const uint16_t* seq_a, uint64_t len_a, const uint16_t* seq_b,
                       uint64_t len_b, const double* xyz_a, uint64_t na, const double* xyz_b, uint64_t nb,
                       const int32_t* wf_idx, int n_threads, double* out_scores) {
    auto dmx = [](const double* xyz, uint64_t n) {
        std::vector<double> m(n * n, 0.0);
        for (uint64_t i = 0; i < n; ++i)
            for (uint64_t j = i + 1; j < n; ++j) {
                const double d = euclidean_distance(xyz + 3 * i, xyz + 3 * j);
                m[i * n + j] = d;
                m[j * n + i] = d;
            }
        return m;
    };
    std::vector<double> ma = dmx(xyz_a, na), mb = dmx(xyz_b, nb);
    return oracle_from_dmxs(p, seq_a, len_a, seq_b, len_b, ma.data(), na, na, mb.data(), nb, nb, wf_idx, n_threads,
                            out_scores);
}

}  // extern "C"
