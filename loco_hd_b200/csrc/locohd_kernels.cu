// locohd_kernels.cu — hand-written sm_100a kernels of the LoCoHD per-anchor scoring path.
//
// Pipeline (all FP64 where the reference is FP64; no tensor cores: there is no dense contraction):
//   K0  build_cells_kernel   one CTA per structure: bounding box, cell grid, counting sort of the primitives
//                            into cell order (replaces KdTree::build_by_ordered_float, locohd.rs:504-510)
//   K1  env_count_kernel     one warp per anchor: float4 prefilter over the 27 neighbour cells + exact FP64
//                            membership test (kd-tree `within_radius` predicate) + tag rule (locohd.rs:521-528)
//   K1' env_fill_kernel      same gather, staged in shared memory, per-warp bucket sort by distance
//                            (utils::sort_together, utils.rs:25-39), written to the environment store
//   K2  score_kernel         one warp per anchor pair: merge-path split of the two sorted environments over the
//                            32 lanes, per-lane category counts by warp prefix sums, per-lane walk that
//                            accumulates dW * H (stat_dist_integral, locohd.rs:61-226, as a flat prefix scan)
// plus the small kernels around them (scan of counts, row sorting for from_dmxs/from_coords, the exact-order
// sequential walk for from_anchors, leaf-math probes).
#include "locohd_kernels.cuh"

#include <cfloat>
#include <cmath>

namespace locohd {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void raise(int* err, int code) { atomicCAS(err, 0, code); }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------
__global__ void convert_categories_kernel(const uint16_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t n,
                                          int C) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned c = in[i];
        out[i] = (c < (unsigned)C) ? (uint8_t)c : kUnknownCat8;
    }
}

__global__ void validate_xyz_kernel(const double* __restrict__ xyz, uint64_t n3, int* err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3 && !isfinite(xyz[i])) raise(err, LOCOHD_ERR_NAN);
}

// ------------------------------------------------------------------------------------------------
// K0: cell list.  One CTA (256 threads) per structure.
// ------------------------------------------------------------------------------------------------
constexpr int kCellThreads = 256;

__device__ __forceinline__ int cell_coord(double rel, double inv_cell, int n) {
    int c = (int)(rel * inv_cell);
    return min(max(c, 0), n - 1);
}

__global__ void __launch_bounds__(kCellThreads) build_cells_kernel(StructsView s, double threshold) {
    __shared__ double red[6][kCellThreads / 32];
    __shared__ StructMeta sm_meta;
    __shared__ uint32_t hist[kMaxCells];
    __shared__ uint32_t warp_tot[kCellThreads / 32];

    const uint64_t sid = blockIdx.x;
    const uint64_t base = s.prim_off[sid];
    const uint32_t n = (uint32_t)(s.prim_off[sid + 1] - base);
    const double* xyz = s.xyz + 3 * base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // ---- bounding box
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint32_t i = tid; i < n; i += kCellThreads) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = xyz[3 * (uint64_t)i + k];
            mn[k] = fmin(mn[k], v);
            mx[k] = fmax(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(kFull, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(kFull, mx[k], o));
        }
        if (lane == 0) { red[k][wid] = mn[k]; red[3 + k][wid] = mx[k]; }
    }
    __syncthreads();
    if (tid == 0) {
        double lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            lo[k] = red[k][0]; hi[k] = red[3 + k][0];
            for (int w = 1; w < kCellThreads / 32; ++w) { lo[k] = fmin(lo[k], red[k][w]); hi[k] = fmax(hi[k], red[3 + k][w]); }
        }
        if (n == 0) { lo[0] = lo[1] = lo[2] = 0.0; hi[0] = hi[1] = hi[2] = 0.0; }
        const double ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        const double emax = fmax(ext[0], fmax(ext[1], ext[2]));
        // cell edge slightly larger than the radius so that rounding can never push a neighbour two cells away
        double inv_cell = 1.0 / (threshold * (1.0 + 1e-6));
        if (!(inv_cell > 0.0) || !isfinite(inv_cell)) inv_cell = 0.0;  // infinite radius: a single cell
        if (emax * inv_cell > (double)kMaxCellsAxis - 0.001) inv_cell = ((double)kMaxCellsAxis - 0.001) / emax;
        StructMeta m;
        m.ox = lo[0]; m.oy = lo[1]; m.oz = lo[2];
        m.inv_cell = inv_cell;
        m.nx = min(kMaxCellsAxis, (int)(ext[0] * inv_cell) + 1);
        m.ny = min(kMaxCellsAxis, (int)(ext[1] * inv_cell) + 1);
        m.nz = min(kMaxCellsAxis, (int)(ext[2] * inv_cell) + 1);
        // FP32 prefilter: relative coordinates are rounded to f32 (error <= emax * 2^-24 each); the bound below
        // is generous (see DESIGN.md "prefilter margin").
        const double delta = 4.0 * emax * 5.9604644775390625e-8;
        const double tr = threshold + 2.0 * delta;
        const double t2 = tr * tr * (1.0 + 1e-6);
        float tf = (t2 < 3.0e38) ? (float)t2 : INFINITY;
        if (isfinite(tf)) tf = nextafterf(tf, INFINITY);
        m.thr2f = tf;
        sm_meta = m;
        s.meta[sid] = m;
    }
    for (int c = tid; c < kMaxCells; c += kCellThreads) hist[c] = 0;
    __syncthreads();
    const StructMeta m = sm_meta;
    const int ncell = m.nx * m.ny * m.nz;

    // ---- histogram
    for (uint32_t i = tid; i < n; i += kCellThreads) {
        const double x = xyz[3 * (uint64_t)i], y = xyz[3 * (uint64_t)i + 1], z = xyz[3 * (uint64_t)i + 2];
        const int cx = cell_coord(x - m.ox, m.inv_cell, m.nx);
        const int cy = cell_coord(y - m.oy, m.inv_cell, m.ny);
        const int cz = cell_coord(z - m.oz, m.inv_cell, m.nz);
        atomicAdd(&hist[(cz * m.ny + cy) * m.nx + cx], 1u);
    }
    __syncthreads();

    // ---- exclusive scan of hist[0..ncell) (16 entries per thread, kMaxCells = 4096)
    constexpr int kPer = kMaxCells / kCellThreads;
    uint32_t local[kPer];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int c = tid * kPer + q;
        local[q] = (c < ncell) ? hist[c] : 0u;
        sum += local[q];
    }
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += warp_tot[w];
    uint32_t run = wbase + incl - sum;
    __syncthreads();
    uint32_t* cell_start = s.cell_start + sid * kCellStride;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int c = tid * kPer + q;
        if (c < ncell) { hist[c] = run; cell_start[c] = run; }
        run += local[q];
    }
    if (tid == 0) cell_start[ncell] = n;
    __syncthreads();

    // ---- scatter into cell order (hist now holds the running cursor of each cell)
    for (uint32_t i = tid; i < n; i += kCellThreads) {
        const double x = xyz[3 * (uint64_t)i], y = xyz[3 * (uint64_t)i + 1], z = xyz[3 * (uint64_t)i + 2];
        const int cx = cell_coord(x - m.ox, m.inv_cell, m.nx);
        const int cy = cell_coord(y - m.oy, m.inv_cell, m.ny);
        const int cz = cell_coord(z - m.oz, m.inv_cell, m.nz);
        const uint32_t pos = atomicAdd(&hist[(cz * m.ny + cy) * m.nx + cx], 1u);
        PrimRec r;
        r.x = x; r.y = y; r.z = z;
        r.orig = i;
        r.cat = s.cat[base + i];
        s.pd[base + pos] = r;
        s.pf[base + pos] = make_float4((float)(x - m.ox), (float)(y - m.oy), (float)(z - m.oz),
                                       __uint_as_float(s.tag[base + i]));
        s.sorted_pos[base + i] = pos;
    }
}

// ------------------------------------------------------------------------------------------------
// K1: neighbour gather.  One warp per anchor.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tag_pair_in_table(const KParams& p, uint32_t a, uint32_t b) {
    const uint64_t key = ((uint64_t)a << 32) | b;
    uint64_t lo = 0, hi = p.n_tag_pairs;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint64_t v = __ldg(p.tag_pairs + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < p.n_tag_pairs && __ldg(p.tag_pairs + lo) == key;
}

// TagPairingRule::pair_accepted on interned ids (tag_pairing_rule.rs:49-76); pair = (anchor.tag, neighbour.tag)
__device__ __forceinline__ bool tag_rule_accepts(const KParams& p, uint32_t anchor_tag, uint32_t other_tag) {
    if (p.tpr_kind == LOCOHD_TPR_WITHOUT_LIST) {
        const bool same = anchor_tag == other_tag;
        return p.tpr_accept_same ? same : !same;
    }
    bool acc = tag_pair_in_table(p, anchor_tag, other_tag);
    if (!p.tpr_ordered) acc = acc || tag_pair_in_table(p, other_tag, anchor_tag);
    return p.tpr_accepted_pairs ? acc : !acc;
}

struct AnchorRef {
    bool ok;
    uint64_t base;       // first primitive of the structure
    uint32_t jpos;       // cell-sorted position of the anchor
    const uint32_t* cell_start;
    StructMeta m;
};

__device__ __forceinline__ AnchorRef resolve_anchor(const StructsView& s, const uint32_t* anchor_struct,
                                                    const uint32_t* anchor_prim, uint64_t e, int* err) {
    AnchorRef a;
    const uint64_t sid = anchor_struct ? anchor_struct[e] : 0;
    const uint32_t prim = anchor_prim[e];
    a.ok = false;
    if (sid >= s.n_structs) { raise(err, LOCOHD_ERR_INDEX); return a; }
    a.base = s.prim_off[sid];
    const uint64_t n = s.prim_off[sid + 1] - a.base;
    if (prim >= n) { raise(err, LOCOHD_ERR_INDEX); return a; }  // prim_seq[anchor_idx] panics upstream (locohd.rs:521)
    a.jpos = s.sorted_pos[a.base + prim];
    a.cell_start = s.cell_start + sid * kCellStride;
    a.m = s.meta[sid];
    a.ok = true;
    return a;
}

// Visits every member of the anchor's environment.  `emit(slot, d2, j)` is called by the lane that owns an
// accepted primitive (j = cell-sorted position), with slot = running index inside the environment.
// Membership = box test + d^2 < r^2 with unfused FP64 arithmetic in the kd-tree crate's operation order
// (neighbour minus anchor; ((dx^2 + dy^2) + dz^2)); the anchor itself is always kept, others must pass the tag rule
// (locohd.rs:521-528).  Returns the environment size (warp-uniform).
template <class Emit>
__device__ __forceinline__ uint32_t gather_environment(const StructsView& s, const KParams& p, const AnchorRef& a,
                                                       double threshold, int lane, Emit&& emit) {
    const PrimRec q = s.pd[a.base + a.jpos];
    const float4 qf = s.pf[a.base + a.jpos];
    const uint32_t qtag = __float_as_uint(qf.w);
    const double r2 = __dmul_rn(threshold, threshold);
    const double lox = q.x - threshold, hix = q.x + threshold;
    const double loy = q.y - threshold, hiy = q.y + threshold;
    const double loz = q.z - threshold, hiz = q.z + threshold;
    const StructMeta& m = a.m;
    const int cx = cell_coord(q.x - m.ox, m.inv_cell, m.nx);
    const int cy = cell_coord(q.y - m.oy, m.inv_cell, m.ny);
    const int cz = cell_coord(q.z - m.oz, m.inv_cell, m.nz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, m.nx - 1);
    const unsigned lt = lanemask_lt();
    uint32_t total = 0;
    for (int zz = max(cz - 1, 0); zz <= min(cz + 1, m.nz - 1); ++zz) {
        for (int yy = max(cy - 1, 0); yy <= min(cy + 1, m.ny - 1); ++yy) {
            const int row = (zz * m.ny + yy) * m.nx;
            const uint32_t beg = __ldg(a.cell_start + row + x0);
            const uint32_t end = __ldg(a.cell_start + row + x1 + 1);
            for (uint32_t j0 = beg; j0 < end; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool pass = false;
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < end) {
                    c = __ldg(s.pf + a.base + j);
                    const float dx = c.x - qf.x, dy = c.y - qf.y, dz = c.z - qf.z;
                    const float d2f = dx * dx + dy * dy + dz * dz;
                    pass = d2f <= m.thr2f;
                }
                bool acc = false;
                double d2 = 0.0;
                if (pass) {
                    const PrimRec r = s.pd[a.base + j];
                    const bool in_box = !(r.x < lox) && !(r.x > hix) && !(r.y < loy) && !(r.y > hiy) &&
                                        !(r.z < loz) && !(r.z > hiz);
                    const double ex = r.x - q.x, ey = r.y - q.y, ez = r.z - q.z;
                    d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                    acc = in_box && (d2 < r2);
                    if (acc && j != a.jpos) acc = tag_rule_accepts(p, qtag, __float_as_uint(c.w));
                }
                const unsigned bal = __ballot_sync(kFull, acc);
                if (acc) emit(total + __popc(bal & lt), d2, j);
                total += __popc(bal);
            }
        }
    }
    return total;
}

constexpr int kEnvWarps = 4;

__global__ void __launch_bounds__(kEnvWarps * 32) env_count_kernel(StructsView s, KParams p, uint64_t n_env,
                                                                   const uint32_t* __restrict__ anchor_struct,
                                                                   const uint32_t* __restrict__ anchor_prim,
                                                                   double threshold, uint32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const uint64_t e = (uint64_t)blockIdx.x * kEnvWarps + (threadIdx.x >> 5);
    if (e >= n_env) return;
    const AnchorRef a = resolve_anchor(s, anchor_struct, anchor_prim, e, p.err);
    uint32_t m = 0;
    if (a.ok) m = gather_environment(s, p, a, threshold, lane, [](uint32_t, double, uint32_t) {});
    if (lane == 0) count[e] = m;
}

// ------------------------------------------------------------------------------------------------
// Per-warp bucket sort in shared memory.
//   keys are non-NaN doubles; buckets are a monotone function of the key chosen so that spherical
//   environments fill them evenly (members within distance d grow like d^3), then every entry finds its exact
//   rank inside its bucket by comparison.  Expected O(M) work per environment instead of O(M log^2 M).
// ------------------------------------------------------------------------------------------------
template <int CAP>
struct WarpSortLayout {
    static constexpr int NB = CAP / 2;
    static constexpr int kKeyOff = 0;
    static constexpr int kPayOff = kKeyOff + 8 * CAP;
    static constexpr int kStartOff = kPayOff + 4 * CAP;
    static constexpr int kCursorOff = kStartOff + 4 * (NB + 4);
    static constexpr int kBktOff = kCursorOff + 4 * NB;
    static constexpr int kPermOff = kBktOff + 2 * CAP;
    static constexpr int kBytes = kPermOff + 2 * CAP;
};

// KEY_IS_SQUARED: keys are squared distances (bucket ~ x^1.5) else plain distances (bucket ~ x^3).
template <int CAP, bool KEY_IS_SQUARED, class Emit>
__device__ __forceinline__ void warp_bucket_sort(unsigned char* smem, uint32_t M, int lane, Emit&& emit) {
    using L = WarpSortLayout<CAP>;
    constexpr int NB = L::NB;
    double* key = reinterpret_cast<double*>(smem + L::kKeyOff);
    uint32_t* pay = reinterpret_cast<uint32_t*>(smem + L::kPayOff);
    uint32_t* start = reinterpret_cast<uint32_t*>(smem + L::kStartOff);
    uint32_t* cursor = reinterpret_cast<uint32_t*>(smem + L::kCursorOff);
    uint16_t* bkt = reinterpret_cast<uint16_t*>(smem + L::kBktOff);
    uint16_t* perm = reinterpret_cast<uint16_t*>(smem + L::kPermOff);

    // largest finite key
    double kmax = 0.0;
    for (uint32_t e = lane; e < M; e += 32) {
        const double k = key[e];
        if (isfinite(k)) kmax = fmax(kmax, k);
    }
    for (int o = 16; o; o >>= 1) kmax = fmax(kmax, __shfl_xor_sync(kFull, kmax, o));
    const double inv = (kmax > 0.0) ? 1.0 / kmax : 0.0;

    for (int b = lane; b < NB; b += 32) start[b] = 0;
    __syncwarp();
    for (uint32_t e = lane; e < M; e += 32) {
        const float x = (float)(key[e] * inv);
        const float f = KEY_IS_SQUARED ? x * sqrtf(fmaxf(x, 0.f)) : x * x * x;
        int b = (int)fminf(f * (float)NB, (float)(NB - 1));  // NaN/negative handled by the max below
        b = max(b, 0);
        bkt[e] = (uint16_t)b;
        atomicAdd(&start[b], 1u);
    }
    __syncwarp();
    // exclusive scan over NB buckets: NB/32 consecutive buckets per lane
    {
        constexpr int PER = NB / 32;
        uint32_t loc[PER];
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { loc[q] = start[lane * PER + q]; sum += loc[q]; }
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        uint32_t run = incl - sum;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            start[lane * PER + q] = run;
            cursor[lane * PER + q] = run;
            run += loc[q];
        }
        if (lane == 31) start[NB] = run;
    }
    __syncwarp();
    for (uint32_t e = lane; e < M; e += 32) {
        const uint32_t pos = atomicAdd(&cursor[bkt[e]], 1u);
        perm[pos] = (uint16_t)e;
    }
    __syncwarp();
    for (uint32_t e = lane; e < M; e += 32) {
        const uint32_t b = bkt[e];
        const uint32_t s0 = start[b], s1 = start[b + 1];
        const double k = key[e];
        uint32_t rank = 0;
        for (uint32_t t = s0; t < s1; ++t) {
            const uint32_t f = perm[t];
            const double kf = key[f];
            rank += (kf < k || (kf == k && f < e)) ? 1u : 0u;
        }
        emit(s0 + rank, k, pay[e]);
    }
}

// K1': gather + sort + store.  Processes environments whose size lies in (MIN_M, CAP].
template <int CAP>
__global__ void __launch_bounds__(kEnvWarps * 32) env_fill_kernel(StructsView s, KParams p,
                                                                  const uint32_t* __restrict__ anchor_struct,
                                                                  const uint32_t* __restrict__ anchor_prim,
                                                                  double threshold, EnvOut out, uint32_t min_m) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using L = WarpSortLayout<CAP>;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t e = (uint64_t)blockIdx.x * kEnvWarps + wib;
    if (e >= out.n_env) return;
    const uint32_t expect = out.count[e];
    if (expect <= min_m || expect > (uint32_t)CAP) return;
    unsigned char* smem = smem_raw + (size_t)wib * L::kBytes;
    double* key = reinterpret_cast<double*>(smem + L::kKeyOff);
    uint32_t* pay = reinterpret_cast<uint32_t*>(smem + L::kPayOff);

    const AnchorRef a = resolve_anchor(s, anchor_struct, anchor_prim, e, p.err);
    if (!a.ok) return;
    const uint32_t M = gather_environment(s, p, a, threshold, lane, [&](uint32_t slot, double d2, uint32_t j) {
        if (slot < (uint32_t)CAP) { key[slot] = d2; pay[slot] = j; }
    });
    if (M != expect) { raise(p.err, LOCOHD_ERR_CUDA); return; }  // count and fill passes must agree
    __syncwarp();
    const uint64_t off = out.off[e];
    const PrimRec* pd = s.pd + a.base;
    warp_bucket_sort<CAP, true>(smem, M, lane, [&](uint32_t pos, double d2, uint32_t j) {
        out.dist[off + pos] = sqrt(d2);  // utils.rs:1-8
        const PrimRec* r = pd + j;
        out.cat[off + pos] = (uint8_t)r->cat;
        if (out.idx) out.idx[off + pos] = r->orig;
    });
}

// Environments larger than every shared-memory class: written unsorted, then sorted in place by
// bitonic_sort_big_kernel.
__global__ void __launch_bounds__(kEnvWarps * 32) env_fill_unsorted_kernel(StructsView s, KParams p,
                                                                           const uint32_t* __restrict__ anchor_struct,
                                                                           const uint32_t* __restrict__ anchor_prim,
                                                                           double threshold, EnvOut out,
                                                                           uint32_t min_m) {
    const int lane = threadIdx.x & 31;
    const uint64_t e = (uint64_t)blockIdx.x * kEnvWarps + (threadIdx.x >> 5);
    if (e >= out.n_env) return;
    const uint32_t expect = out.count[e];
    if (expect <= min_m) return;
    const AnchorRef a = resolve_anchor(s, anchor_struct, anchor_prim, e, p.err);
    if (!a.ok) return;
    const uint64_t off = out.off[e];
    const PrimRec* pd = s.pd + a.base;
    const uint32_t M = gather_environment(s, p, a, threshold, lane, [&](uint32_t slot, double d2, uint32_t j) {
        if (slot < expect) {
            out.dist[off + slot] = sqrt(d2);
            const PrimRec* r = pd + j;
            out.cat[off + slot] = (uint8_t)r->cat;
            if (out.idx) out.idx[off + slot] = r->orig;
        }
    });
    if (M != expect) raise(p.err, LOCOHD_ERR_CUDA);
}

// In-place ascending bitonic network (min always to the lower index, so the virtual +inf padding above M
// never moves).  One CTA per environment with more than min_m members.
constexpr int kBigThreads = 256;
__global__ void __launch_bounds__(kBigThreads) bitonic_sort_big_kernel(EnvOut out, uint32_t min_m) {
    const uint64_t e = blockIdx.x;
    const uint32_t M = out.count[e];
    if (M <= min_m) return;
    const uint64_t off = out.off[e];
    double* d = out.dist + off;
    uint8_t* c = out.cat + off;
    uint32_t* ix = out.idx ? out.idx + off : nullptr;
    uint32_t n2 = 1;
    while (n2 < M) n2 <<= 1;
    auto cex = [&](uint32_t i, uint32_t q) {
        if (q < M) {
            const double di = d[i], dq = d[q];
            if (dq < di) {
                d[i] = dq; d[q] = di;
                const uint8_t t = c[i]; c[i] = c[q]; c[q] = t;
                if (ix) { const uint32_t u = ix[i]; ix[i] = ix[q]; ix[q] = u; }
            }
        }
    };
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        const uint32_t h = k >> 1;
        for (uint32_t t = threadIdx.x; t < n2 / 2; t += kBigThreads) {
            const uint32_t i = (t / h) * k + (t % h);
            cex(i, i ^ (k - 1));
        }
        __syncthreads();
        for (uint32_t j = h >> 1; j >= 1; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < n2 / 2; t += kBigThreads) {
                const uint32_t i = (t / j) * (2 * j) + (t % j);
                cex(i, i + j);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan of the environment sizes (u32 -> u64 offsets) + size-class statistics.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long wsum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w < wid) base += wsum[w];
        tot += wsum[w];
    }
    __syncthreads();
    if (total) *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ count, uint64_t n,
                                                                      uint64_t* __restrict__ block_sums,
                                                                      ScanResult* res) {
    const uint64_t t0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPer;
    unsigned long long sum = 0;
    unsigned mx = 0, ns = 0, nm = 0, nl = 0, nh = 0;
    for (int q = 0; q < kScanPer; ++q) {
        const uint64_t i = t0 + q;
        if (i < n) {
            const unsigned c = count[i];
            sum += c;
            mx = max(mx, c);
            if (c <= 256) ++ns; else if (c <= 512) ++nm; else if (c <= 2048) ++nl; else ++nh;
        }
    }
    unsigned long long tot;
    block_exclusive_scan(sum, &tot);
    for (int o = 16; o; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(kFull, mx, o));
        ns += __shfl_xor_sync(kFull, ns, o);
        nm += __shfl_xor_sync(kFull, nm, o);
        nl += __shfl_xor_sync(kFull, nl, o);
        nh += __shfl_xor_sync(kFull, nh, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&res->max_count, mx);
        if (ns) atomicAdd(&res->n_small, ns);
        if (nm) atomicAdd(&res->n_medium, nm);
        if (nl) atomicAdd(&res->n_large, nl);
        if (nh) atomicAdd(&res->n_huge, nh);
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(uint64_t* block_sums, uint64_t n_blocks,
                                                                       ScanResult* res) {
    unsigned long long carry = 0;
    for (uint64_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
        const uint64_t i = b0 + threadIdx.x;
        const unsigned long long v = (i < n_blocks) ? block_sums[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_exclusive_scan(v, &tot);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) res->total = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ count, uint64_t n,
                                                                  const uint64_t* __restrict__ block_sums,
                                                                  uint64_t* __restrict__ off) {
    const uint64_t t0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPer;
    unsigned loc[kScanPer];
    unsigned long long sum = 0;
    for (int q = 0; q < kScanPer; ++q) {
        const uint64_t i = t0 + q;
        loc[q] = (i < n) ? count[i] : 0u;
        sum += loc[q];
    }
    unsigned long long run = block_sums[blockIdx.x] + block_exclusive_scan(sum, nullptr);
    for (int q = 0; q < kScanPer; ++q) {
        const uint64_t i = t0 + q;
        if (i < n) off[i] = run;
        run += loc[q];
        if (i + 1 == n) off[n] = run;
    }
}

// ------------------------------------------------------------------------------------------------
// Rows as environments (from_dmxs / from_coords): every row is sorted by distance.
// ------------------------------------------------------------------------------------------------
__global__ void iota_rows_kernel(uint64_t* off, uint32_t* count, uint64_t n_rows, uint64_t row_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_rows) off[i] = i * row_len;
    if (i < n_rows) count[i] = (uint32_t)row_len;
}

// distance of point i to point j exactly as utils.rs:1-22 (squares are symmetric, the diagonal is 0)
__device__ __forceinline__ double row_distance(const double* xyz, uint64_t i, uint64_t j) {
    if (i == j) return 0.0;
    const double ex = xyz[3 * i] - xyz[3 * j], ey = xyz[3 * i + 1] - xyz[3 * j + 1], ez = xyz[3 * i + 2] - xyz[3 * j + 2];
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
}

template <int CAP>
__global__ void __launch_bounds__(kEnvWarps * 32) rows_fill_kernel(const double* __restrict__ dmx,
                                                                   const uint8_t* __restrict__ cat, uint64_t n_rows,
                                                                   uint64_t row_len, const double* __restrict__ xyz,
                                                                   KParams p, EnvOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using L = WarpSortLayout<CAP>;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t row = (uint64_t)blockIdx.x * kEnvWarps + wib;
    if (row >= n_rows) return;
    unsigned char* smem = smem_raw + (size_t)wib * L::kBytes;
    double* key = reinterpret_cast<double*>(smem + L::kKeyOff);
    uint32_t* pay = reinterpret_cast<uint32_t*>(smem + L::kPayOff);
    const uint32_t M = (uint32_t)row_len;
    bool bad = false;
    for (uint32_t j = lane; j < M; j += 32) {
        const double d = xyz ? row_distance(xyz, row, j) : dmx[row * row_len + j];
        bad |= isnan(d);
        key[j] = d;
        pay[j] = j;
    }
    if (__any_sync(kFull, bad)) { raise(p.err, LOCOHD_ERR_NAN); return; }  // partial_cmp().unwrap() panics (utils.rs:28)
    __syncwarp();
    const uint64_t off = out.off[row];
    warp_bucket_sort<CAP, false>(smem, M, lane, [&](uint32_t pos, double d, uint32_t j) {
        out.dist[off + pos] = d;
        out.cat[off + pos] = cat[j];
        if (out.idx) out.idx[off + pos] = j;
    });
}

__global__ void rows_copy_kernel(const double* __restrict__ dmx, const uint8_t* __restrict__ cat, uint64_t n_rows,
                                 uint64_t row_len, const double* __restrict__ xyz, KParams p, EnvOut out) {
    const uint64_t row = blockIdx.y;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < row_len;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const double d = xyz ? row_distance(xyz, row, j) : dmx[row * row_len + j];
        if (isnan(d)) raise(p.err, LOCOHD_ERR_NAN);
        out.dist[row * row_len + j] = d;
        out.cat[row * row_len + j] = cat[j];
        if (out.idx) out.idx[row * row_len + j] = (uint32_t)j;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: scoring.  One warp per anchor pair.
// ------------------------------------------------------------------------------------------------
constexpr int kScoreMaxWarps = 4;

struct ScoreSmem {
    int per_warp_bytes;
    int state_bytes;
};

__host__ __device__ inline int score_state_bytes(int C) { return ((2 * C * 32 * 8 + 2 * C * 32 * 4) + 15) & ~15; }
__host__ __device__ inline int score_stage_bytes(int cap) { return ((cap * 9 + 16) + 15) & ~15; }

template <bool HELL2>
__global__ void __launch_bounds__(kScoreMaxWarps * 32) score_kernel(ScoreArgs a, KParams P, int warps_per_block,
                                                                    int per_warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t pair = (uint64_t)blockIdx.x * warps_per_block + wib;
    if (wib >= warps_per_block || pair >= a.n_pairs) return;
    const int C = P.C;
    unsigned char* mine = smem_raw + (size_t)wib * per_warp_bytes;
    double* val = reinterpret_cast<double*>(mine);                         // [2C][32]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(mine + 2 * C * 32 * 8);    // [2C][32]
    unsigned char* stage = mine + score_state_bytes(C);

    // ---- which environments
    uint64_t ea, eb;
    if (a.pairs) {
        ea = a.pairs[2 * pair];
        eb = a.pairs[2 * pair + 1];
    } else {
        uint64_t job, within;
        if (a.uniform_n) {
            job = pair / a.uniform_n;
            within = pair - job * a.uniform_n;
        } else {
            uint64_t lo = 0, hi = a.n_jobs;  // last job with job_pair_off[job] <= pair
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (__ldg(a.job_pair_off + mid) <= pair) lo = mid; else hi = mid;
            }
            job = lo;
            within = pair - __ldg(a.job_pair_off + job);
        }
        ea = a.jobs[job].a_first + within;
        eb = a.jobs[job].b_first + within;
    }
    if (ea >= a.a.n_env || eb >= a.b.n_env) { raise(P.err, LOCOHD_ERR_INDEX); return; }
    const uint64_t oa = a.a.off[ea], ob = a.b.off[eb];
    const uint32_t Ma = (uint32_t)(a.a.off[ea + 1] - oa), Mb = (uint32_t)(a.b.off[eb + 1] - ob);
    if (Ma == 0 || Mb == 0) { raise(P.err, LOCOHD_ERR_EMPTY_ENV); return; }        // locohd.rs:74 panics upstream
    const double* gdA = a.a.dist + oa;
    const double* gdB = a.b.dist + ob;
    const uint8_t* gcA = a.a.cat + oa;
    const uint8_t* gcB = a.b.cat + ob;
    if (gdA[0] != 0.0 || gdB[0] != 0.0) { raise(P.err, LOCOHD_ERR_FIRST_NOT_ZERO); return; }  // locohd.rs:74-77

    // ---- stage both environments in shared memory when they fit
    const double* dA = gdA;
    const double* dB = gdB;
    const uint8_t* cA = gcA;
    const uint8_t* cB = gcB;
    if ((int)(Ma + Mb) <= a.stage_cap) {
        double* sd = reinterpret_cast<double*>(stage);
        uint8_t* sc = stage + (size_t)(Ma + Mb) * 8;
        for (uint32_t i = lane; i < Ma; i += 32) { sd[i] = gdA[i]; sc[i] = gcA[i]; }
        for (uint32_t i = lane; i < Mb; i += 32) { sd[Ma + i] = gdB[i]; sc[Ma + i] = gcB[i]; }
        dA = sd; dB = sd + Ma; cA = sc; cB = sc + Ma;
        __syncwarp();
    }
    const uint32_t catA0 = cA[0], catB0 = cB[0];
    // events = members after the anchor
    dA += 1; dB += 1; cA += 1; cB += 1;
    const uint32_t na = Ma - 1, nb = Mb - 1;
    const uint32_t E = na + nb;
    const uint32_t Q = (E + 31) / 32;

    // ---- merge-path split: lane l owns merged events [l*Q, (l+1)*Q); A precedes B on ties
    const uint32_t diag = min(E, (uint32_t)lane * Q);
    uint32_t lo = diag > nb ? diag - nb : 0, hi = min(diag, na);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (dA[mid] <= dB[diag - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    uint32_t i = lo, j = diag - lo;
    uint32_t i1 = __shfl_down_sync(kFull, i, 1), j1 = __shfl_down_sync(kFull, j, 1);
    if (lane == 31) { i1 = na; j1 = nb; }

    // ---- category counts before my chunk: per-lane histogram, then exclusive prefix over the lanes
    for (int r = 0; r < 2 * C; ++r) cnt[r * 32 + lane] = 0;
    bool unknown = (catA0 >= (uint32_t)C) || (catB0 >= (uint32_t)C);
    for (uint32_t x = i; x < i1; ++x) {
        const uint32_t c = cA[x];
        if (c < (uint32_t)C) cnt[c * 32 + lane] += 1; else unknown = true;
    }
    for (uint32_t x = j; x < j1; ++x) {
        const uint32_t c = cB[x];
        if (c < (uint32_t)C) cnt[(C + c) * 32 + lane] += 1; else unknown = true;
    }
    if (__any_sync(kFull, unknown)) { raise(P.err, LOCOHD_ERR_UNKNOWN_CATEGORY); return; }  // pmf.rs:38-42
    for (int r = 0; r < 2 * C; ++r) {
        const uint32_t v = cnt[r * 32 + lane];
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += u;
        }
        uint32_t ex = incl - v;
        if (r == (int)catA0 || r == C + (int)catB0) ex += 1;   // anchors (locohd.rs:82-84)
        cnt[r * 32 + lane] = ex;
    }

    // ---- state: HELL2 keeps sqrt(weighted count), otherwise the weighted count itself
    double normA = 0.0, normB = 0.0;
    uint32_t totA = 0, totB = 0;
    for (int r = 0; r < C; ++r) {
        const uint32_t ka = cnt[r * 32 + lane], kb = cnt[(C + r) * 32 + lane];
        const double w = P.cat_w[r];
        normA += (double)ka * w; normB += (double)kb * w;
        totA += ka; totB += kb;
        if (HELL2) {
            const double sw = P.cat_sw[r];
            val[r * 32 + lane] = (ka < (uint32_t)kSqrtTableSize ? __ldg(P.sqrt_tbl + ka) : sqrt((double)ka)) * sw;
            val[(C + r) * 32 + lane] = (kb < (uint32_t)kSqrtTableSize ? __ldg(P.sqrt_tbl + kb) : sqrt((double)kb)) * sw;
        } else {
            val[r * 32 + lane] = (double)ka * w;
            val[(C + r) * 32 + lane] = (double)kb * w;
        }
    }
    auto inv_sqrt_norm = [&](double norm, uint32_t tot) -> double {
        if (P.unit_w && tot < (uint32_t)kSqrtTableSize) return __ldg(P.rsqrt_tbl + tot);
        return 1.0 / sqrt(norm);
    };
    double rA = HELL2 ? inv_sqrt_norm(normA, totA) : 0.0;
    double rB = HELL2 ? inv_sqrt_norm(normB, totB) : 0.0;

    auto stat_dist = [&]() -> double {
        if (HELL2) {
            // (1/2 * sum (sqrt(p_i) - sqrt(q_i))^2)^(1/2) in difference form (statistical_distances.rs:4-10, e = 2)
            double acc = 0.0;
            for (int r = 0; r < C; ++r) {
                // both products are rounded before the subtraction (no FMA contraction): identical compositions
                // must give exactly 0, as they do upstream
                const double u = __dmul_rn(val[r * 32 + lane], rA) - __dmul_rn(val[(C + r) * 32 + lane], rB);
                acc = fma(u, u, acc);
            }
            return sqrt(0.5 * acc);
        } else {
            auto p1 = [&](int r) { return val[r * 32 + lane] / normA; };
            auto p2 = [&](int r) { return val[(C + r) * 32 + lane] / normB; };
            return sd_run(P.sd_kind, P.sd_p0, P.sd_p1, C, p1, p2);
        }
    };

    const WfDev& wf = P.wfs[a.wf_idx ? a.wf_idx[pair] : 0];
    double tprev = 0.0;
    if (i > 0) tprev = dA[i - 1];
    if (j > 0) tprev = fmax(tprev, dB[j - 1]);
    double wprev = wf_cdf(wf, tprev);
    double h = stat_dist();
    double acc = 0.0;

    // ---- walk my chunk
    double ta = (i < i1) ? dA[i] : 0.0, tb = (j < j1) ? dB[j] : 0.0;
    while (i < i1 || j < j1) {
        const bool takeA = (i < i1) && (!(j < j1) || ta <= tb);
        const double t = takeA ? ta : tb;
        const uint32_t c = takeA ? cA[i] : cB[j];
        const double w = wf_cdf(wf, t);
        acc = fma(w - wprev, h, acc);
        wprev = w;
        const int row = (takeA ? 0 : C) + (int)c;
        const uint32_t k = cnt[row * 32 + lane] + 1;
        cnt[row * 32 + lane] = k;
        const double wc = P.cat_w[c];
        if (HELL2) {
            val[row * 32 + lane] = (k < (uint32_t)kSqrtTableSize ? __ldg(P.sqrt_tbl + k) : sqrt((double)k)) * P.cat_sw[c];
        } else {
            val[row * 32 + lane] = (double)k * wc;
        }
        if (takeA) {
            normA += wc; totA += 1;
            if (HELL2) rA = inv_sqrt_norm(normA, totA);
            ++i;
            if (i < i1) ta = dA[i];
        } else {
            normB += wc; totB += 1;
            if (HELL2) rB = inv_sqrt_norm(normB, totB);
            ++j;
            if (j < j1) tb = dB[j];
        }
        h = stat_dist();
    }
    // ---- tail to infinity (locohd.rs:165-171, 204-221): owned by the last lane, whose state is final
    if (lane == 31) acc = fma(wf_cdf(wf, INFINITY) - wprev, h, acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (lane == 0) a.out[pair] = acc;
}

__global__ void job_means_kernel(const double* __restrict__ scores, const uint64_t* __restrict__ job_pair_off,
                                 uint64_t n_jobs, double* __restrict__ means) {
    const int lane = threadIdx.x & 31;
    const uint64_t job = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (job >= n_jobs) return;
    const uint64_t b = job_pair_off[job], e = job_pair_off[job + 1];
    double s = 0.0;
    for (uint64_t i = b + lane; i < e; i += 32) s += scores[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) means[job] = (e > b) ? s / (double)(e - b) : 0.0;
}

// ------------------------------------------------------------------------------------------------
// from_anchors: one caller-ordered pair, walked by a single thread in the reference's exact statement order
// (locohd.rs:61-226), so unsorted inputs behave as they do upstream.
// ------------------------------------------------------------------------------------------------
struct SeqWalk {
    const KParams& P;
    const WfDev& wf;
    double pmf1[LOCOHD_MAX_CATEGORIES], pmf2[LOCOHD_MAX_CATEGORIES];
    int status;
    __device__ SeqWalk(const KParams& P_, const WfDev& wf_) : P(P_), wf(wf_), status(0) {
        for (int r = 0; r < P.C; ++r) { pmf1[r] = 0.0; pmf2[r] = 0.0; }
    }
    __device__ void update(double* pmf, uint32_t c) {  // pmf.rs:47-63
        if (c >= (uint32_t)P.C) { if (!status) status = LOCOHD_ERR_UNKNOWN_CATEGORY; return; }
        pmf[c] += P.cat_w[c];
    }
    __device__ double distance() {  // pmf.rs:65-88
        double n1 = 0.0, n2 = 0.0;
        for (int r = 0; r < P.C; ++r) n1 += pmf1[r];
        for (int r = 0; r < P.C; ++r) n2 += pmf2[r];
        if (n1 == 0.0 || n2 == 0.0) { if (!status) status = LOCOHD_ERR_ZERO_NORM; return 0.0; }
        auto p1 = [&](int r) { return pmf1[r] / n1; };
        auto p2 = [&](int r) { return pmf2[r] / n2; };
        return sd_run(P.sd_kind, P.sd_p0, P.sd_p1, P.C, p1, p2);
    }
    __device__ double range(double from, double to) {  // weight_function.rs:95-120
        if (to < 0.0 || from < 0.0) { if (!status) status = LOCOHD_ERR_NEGATIVE_POINT; return 0.0; }
        return wf_cdf(wf, to) - wf_cdf(wf, from);
    }
};

__global__ void anchor_lists_kernel(KParams P, const uint8_t* __restrict__ sa, uint64_t len_a,
                                    const double* __restrict__ da, const uint8_t* __restrict__ sb, uint64_t len_b,
                                    const double* __restrict__ db, uint32_t wf_idx, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (da[0] != 0.0 || db[0] != 0.0) { raise(P.err, LOCOHD_ERR_FIRST_NOT_ZERO); return; }
    SeqWalk w(P, P.wfs[wf_idx]);
    w.update(w.pmf1, sa[0]);
    w.update(w.pmf2, sb[0]);
    uint64_t ia = 0, ib = 0;
    double integral = 0.0, buffer = 0.0;
    while (ia < len_a - 1 && ib < len_b - 1 && !w.status) {
        const double h = w.distance();
        double nd;
        if (da[ia + 1] < db[ib + 1]) { ++ia; w.update(w.pmf1, sa[ia]); nd = da[ia]; }
        else if (da[ia + 1] > db[ib + 1]) { ++ib; w.update(w.pmf2, sb[ib]); nd = db[ib]; }
        else if (da[ia + 1] == db[ib + 1]) { ++ia; ++ib; w.update(w.pmf1, sa[ia]); w.update(w.pmf2, sb[ib]); nd = da[ia]; }
        else { w.status = LOCOHD_ERR_NAN; break; }
        integral += w.range(buffer, nd) * h;
        buffer = nd;
    }
    if (!w.status) {
        if (ib < len_b - 1) {
            double h = w.distance();
            ++ib;
            integral += w.range(da[len_a - 1], db[ib]) * h;
            w.update(w.pmf2, sb[ib]);
            while (ib < len_b - 1 && !w.status) {
                ++ib;
                h = w.distance();
                integral += w.range(db[ib - 1], db[ib]) * h;
                w.update(w.pmf2, sb[ib]);
            }
            h = w.distance();
            integral += w.range(db[len_b - 1], INFINITY) * h;
        } else if (ia < len_a - 1) {
            double h = w.distance();
            ++ia;
            integral += w.range(db[len_b - 1], da[ia]) * h;
            w.update(w.pmf1, sa[ia]);
            while (ia < len_a - 1 && !w.status) {
                ++ia;
                h = w.distance();
                integral += w.range(da[ia - 1], da[ia]) * h;
                w.update(w.pmf1, sa[ia]);
            }
            h = w.distance();
            integral += w.range(da[len_a - 1], INFINITY) * h;
        } else {
            const double h = w.distance();
            integral += w.range(da[len_a - 1], INFINITY) * h;
        }
    }
    if (w.status) { raise(P.err, w.status); return; }
    *out = integral;
}

// ------------------------------------------------------------------------------------------------
// leaf-math probes
// ------------------------------------------------------------------------------------------------
__global__ void wf_points_kernel(const WfDev* wf, uint64_t n, const double* __restrict__ x, double* __restrict__ out,
                                 int* err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    if (v < 0.0) { raise(err, LOCOHD_ERR_NEGATIVE_POINT); return; }  // weight_function.rs:97-100
    out[i] = wf_cdf(*wf, v);
}

__global__ void sd_run_kernel(int kind, double q0, double q1, int C, uint64_t n, const double* __restrict__ p1,
                              const double* __restrict__ p2, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a = p1 + i * C;
    const double* b = p2 + i * C;
    out[i] = sd_run(kind, q0, q1, C, [&](int r) { return a[r]; }, [&](int r) { return b[r]; });
}

// FP64 FMA peak probe: 8 independent register-resident chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// ================================================================================================
// launchers
// ================================================================================================
int launch_convert_categories(const uint16_t* in, uint8_t* out, uint64_t n, int C, cudaStream_t st) {
    if (!n) return 0;
    convert_categories_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, out, n, C);
    return 1;
}

int launch_validate_xyz(const double* xyz, uint64_t n3, int* err, cudaStream_t st) {
    if (!n3) return 0;
    validate_xyz_kernel<<<blocks_for(n3, 256), 256, 0, st>>>(xyz, n3, err);
    return 1;
}

int launch_build_cells(const StructsView& s, double threshold, cudaStream_t st) {
    if (!s.n_structs) return 0;
    build_cells_kernel<<<(unsigned)s.n_structs, kCellThreads, 0, st>>>(s, threshold);
    return 1;
}

int launch_env_count(const StructsView& s, const KParams& p, uint64_t n_env, const uint32_t* anchor_struct,
                     const uint32_t* anchor_prim, double threshold, uint32_t* count, cudaStream_t st) {
    if (!n_env) return 0;
    env_count_kernel<<<blocks_for(n_env, kEnvWarps), kEnvWarps * 32, 0, st>>>(s, p, n_env, anchor_struct, anchor_prim,
                                                                              threshold, count);
    return 1;
}

uint64_t scan_scratch_entries(uint64_t n) { return (n + kScanTile - 1) / kScanTile + 1; }

int launch_scan_counts(const uint32_t* count, uint64_t n, uint64_t* off, uint64_t* block_sums, ScanResult* res,
                       cudaStream_t st) {
    cudaMemsetAsync(res, 0, sizeof(ScanResult), st);
    if (!n) { cudaMemsetAsync(off, 0, sizeof(uint64_t), st); return 0; }
    const unsigned nb = blocks_for(n, kScanTile);
    scan_tile_sums_kernel<<<nb, kScanThreads, 0, st>>>(count, n, block_sums, res);
    scan_block_sums_kernel<<<1, kScanThreads, 0, st>>>(block_sums, nb, res);
    scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(count, n, block_sums, off);
    return 3;
}

template <int CAP>
static int launch_env_fill_class(const StructsView& s, const KParams& p, const uint32_t* anchor_struct,
                                 const uint32_t* anchor_prim, double threshold, const EnvOut& out, uint32_t min_m,
                                 cudaStream_t st) {
    const int smem = WarpSortLayout<CAP>::kBytes * kEnvWarps;
    cudaFuncSetAttribute(env_fill_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    env_fill_kernel<CAP><<<blocks_for(out.n_env, kEnvWarps), kEnvWarps * 32, smem, st>>>(s, p, anchor_struct,
                                                                                        anchor_prim, threshold, out,
                                                                                        min_m);
    return 1;
}

int launch_env_fill(const StructsView& s, const KParams& p, const uint32_t* anchor_struct,
                    const uint32_t* anchor_prim, double threshold, const EnvOut& out, const ScanResult& cls,
                    cudaStream_t st) {
    if (!out.n_env) return 0;
    int launches = 0;
    if (cls.n_small) launches += launch_env_fill_class<256>(s, p, anchor_struct, anchor_prim, threshold, out, 0, st);
    if (cls.n_medium) launches += launch_env_fill_class<512>(s, p, anchor_struct, anchor_prim, threshold, out, 256, st);
    if (cls.n_large) launches += launch_env_fill_class<2048>(s, p, anchor_struct, anchor_prim, threshold, out, 512, st);
    if (cls.n_huge) {
        env_fill_unsorted_kernel<<<blocks_for(out.n_env, kEnvWarps), kEnvWarps * 32, 0, st>>>(
            s, p, anchor_struct, anchor_prim, threshold, out, 2048);
        bitonic_sort_big_kernel<<<(unsigned)out.n_env, kBigThreads, 0, st>>>(out, 2048);
        launches += 2;
    }
    return launches;
}

int launch_fill_u64_iota_rows(uint64_t* off, uint32_t* count, uint64_t n_rows, uint64_t row_len, cudaStream_t st) {
    iota_rows_kernel<<<blocks_for(n_rows + 1, 256), 256, 0, st>>>(off, count, n_rows, row_len);
    return 1;
}

template <int CAP>
static int launch_rows_class(const double* dmx, const uint8_t* cat, uint64_t n_rows, uint64_t row_len,
                             const double* xyz, const KParams& p, const EnvOut& out, cudaStream_t st) {
    const int smem = WarpSortLayout<CAP>::kBytes * kEnvWarps;
    cudaFuncSetAttribute(rows_fill_kernel<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rows_fill_kernel<CAP><<<blocks_for(n_rows, kEnvWarps), kEnvWarps * 32, smem, st>>>(dmx, cat, n_rows, row_len, xyz,
                                                                                      p, out);
    return 1;
}

int launch_rows_fill(const double* dmx, const uint8_t* cat, uint64_t n_rows, uint64_t row_len, const double* xyz,
                     const KParams& p, const EnvOut& out, cudaStream_t st) {
    if (!n_rows || !row_len) return 0;
    if (row_len <= 256) return launch_rows_class<256>(dmx, cat, n_rows, row_len, xyz, p, out, st);
    if (row_len <= 512) return launch_rows_class<512>(dmx, cat, n_rows, row_len, xyz, p, out, st);
    if (row_len <= 2048) return launch_rows_class<2048>(dmx, cat, n_rows, row_len, xyz, p, out, st);
    dim3 grid((unsigned)((row_len + 255) / 256), (unsigned)n_rows);
    if (grid.x > 64) grid.x = 64;
    rows_copy_kernel<<<grid, 256, 0, st>>>(dmx, cat, n_rows, row_len, xyz, p, out);
    bitonic_sort_big_kernel<<<(unsigned)n_rows, kBigThreads, 0, st>>>(out, 0);
    return 2;
}

int launch_score(const ScoreArgs& args, const KParams& p, unsigned max_members_a, unsigned max_members_b,
                 cudaStream_t st) {
    if (!args.n_pairs) return 0;
    ScoreArgs a = args;
    const int state = score_state_bytes(p.C);
    const int budget = 200 * 1024;
    // stage both environments when a warp's share of shared memory allows it
    uint64_t want = (uint64_t)max_members_a + max_members_b;
    int cap = (int)(want > 4096 ? 4096 : want);
    int per_warp = state + score_stage_bytes(cap);
    if (per_warp > budget) { cap = 0; per_warp = state + score_stage_bytes(0); }
    int warps = budget / per_warp;
    if (warps > kScoreMaxWarps) warps = kScoreMaxWarps;
    if (warps < 1) return -1;
    a.stage_cap = cap;
    const int smem = per_warp * warps;
    const unsigned grid = blocks_for(args.n_pairs, (unsigned)warps);
    if (p.hell2) {
        cudaFuncSetAttribute(score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        score_kernel<true><<<grid, warps * 32, smem, st>>>(a, p, warps, per_warp);
    } else {
        cudaFuncSetAttribute(score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        score_kernel<false><<<grid, warps * 32, smem, st>>>(a, p, warps, per_warp);
    }
    return 1;
}

int launch_job_means(const double* scores, const uint64_t* job_pair_off, uint64_t n_jobs, double* means,
                     cudaStream_t st) {
    if (!n_jobs) return 0;
    job_means_kernel<<<blocks_for(n_jobs * 32, 256), 256, 0, st>>>(scores, job_pair_off, n_jobs, means);
    return 1;
}

int launch_anchor_lists(const KParams& p, const uint8_t* seq_a, uint64_t len_a, const double* da, const uint8_t* seq_b,
                        uint64_t len_b, const double* db, uint32_t wf_idx, double* out, cudaStream_t st) {
    anchor_lists_kernel<<<1, 32, 0, st>>>(p, seq_a, len_a, da, seq_b, len_b, db, wf_idx, out);
    return 1;
}

int launch_wf_points(const WfDev* wf, uint64_t n, const double* x, double* out, int* err, cudaStream_t st) {
    if (!n) return 0;
    wf_points_kernel<<<blocks_for(n, 256), 256, 0, st>>>(wf, n, x, out, err);
    return 1;
}

int launch_fp64_peak(double* scratch, int blocks, int iters, cudaStream_t st) {
    fp64_peak_kernel<<<blocks, 256, 0, st>>>(scratch, iters, 0.999999, 1e-9);
    return 1;
}

int launch_sd_run(int kind, double q0, double q1, int C, uint64_t n, const double* p1, const double* p2, double* out,
                  cudaStream_t st) {
    if (!n) return 0;
    sd_run_kernel<<<blocks_for(n, 128), 128, 0, st>>>(kind, q0, q1, C, n, p1, p2, out);
    return 1;
}

}  // namespace locohd
