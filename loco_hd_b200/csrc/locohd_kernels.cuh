// locohd_kernels.cuh — internal interface between the C ABI (locohd_capi.cu) and the sm_100a kernels
// (locohd_kernels.cu).  Not installed; the public boundary is include/locohd_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "locohd_math.cuh"

namespace locohd {

constexpr int kMaxCellsAxis = 16;
constexpr int kMaxCells = kMaxCellsAxis * kMaxCellsAxis * kMaxCellsAxis;  // per structure
constexpr int kCellStride = kMaxCells + 1;                                 // cell_start entries per structure
constexpr uint8_t kUnknownCat8 = 0xFF;
constexpr int kSqrtTableSize = 4096;

// One primitive in cell-sorted order: exact coordinates for the FP64 membership test and distance.
struct __align__(32) PrimRec {
    double x, y, z;
    uint32_t orig;  // index inside its structure, original order
    uint32_t cat;   // category id (0..254) or kUnknownCat8
};

// Per-structure cell grid (built for one threshold).
struct __align__(16) StructMeta {
    double ox, oy, oz;   // grid origin = bounding-box minimum
    double inv_cell;     // 1 / cell edge (cell edge >= threshold * (1 + 1e-6))
    int nx, ny, nz;
    float thr2f;         // conservative squared radius for the FP32 prefilter
};

// Kernel-visible parameter block of one LoCoHD instance (LoCoHD struct, locohd.rs:42-55).
struct KParams {
    int C;
    int sd_kind;
    double sd_p0, sd_p1;
    int hell2;    // Hellinger with exponent exactly 2 -> sqrt-table fast path
    int unit_w;   // all category weights == 1
    int tpr_kind, tpr_accept_same, tpr_accepted_pairs, tpr_ordered;
    uint64_t n_tag_pairs;
    const uint64_t* tag_pairs;  // sorted
    const double* cat_w;        // [C]
    const double* cat_sw;       // [C] sqrt(w)
    const WfDev* wfs;
    int n_wf;
    const double* sqrt_tbl;     // sqrt(k), k < kSqrtTableSize
    const double* rsqrt_tbl;    // 1/sqrt(k)
    int* err;                   // device error word (first locohd_status raised by a kernel)
};

struct StructsView {
    uint64_t n_structs;
    const uint64_t* prim_off;     // [n_structs + 1]
    const double* xyz;            // raw [n_prims][3]
    const uint8_t* cat;           // raw [n_prims]
    const uint32_t* tag;          // raw [n_prims]
    StructMeta* meta;             // [n_structs]
    float4* pf;                   // cell-sorted: (x - ox, y - oy, z - oz) as f32, w = tag bits
    PrimRec* pd;                  // cell-sorted exact records
    uint32_t* sorted_pos;         // original index -> cell-sorted position (inside the structure)
    uint32_t* cell_start;         // [n_structs][kCellStride]
};

struct EnvView {
    uint64_t n_env;
    const uint64_t* off;    // [n_env] first member of every environment (any order, no overlap)
    const uint32_t* count;  // [n_env] members per environment
    const double* key;      // ascending per environment: W(distance) when key_is_w, else the distance
    const uint8_t* cat;
    int key_is_w;
};

struct FillStats {  // written by the fill kernels, read back by the host
    unsigned long long cursor;  // members allocated so far (== total when the launch is over)
    unsigned int max_count;
    unsigned int n_big;         // environments larger than the shared-memory class of the launch
    unsigned int overflow;      // some environment did not fit below `capacity`
    unsigned int pad;
};

struct EnvOut {
    uint64_t n_env;
    uint64_t* off;
    uint32_t* count;
    double* key;
    uint8_t* cat;
    double* dist;     // plain distances (debug/parity) or nullptr
    uint32_t* idx;    // primitive indices (debug/parity) or nullptr
    uint64_t capacity;
    FillStats* stats;
    int key_is_w;
};

struct ScoreArgs {
    EnvView a, b;
    uint64_t n_pairs;
    const uint32_t* pairs;         // explicit mode: [n_pairs][2]; nullptr -> job mode
    const locohd_job* jobs;        // job mode
    const uint64_t* job_pair_off;  // [n_jobs + 1]
    uint64_t n_jobs;
    uint64_t uniform_n;            // > 0: every job has this many pairs
    const uint32_t* wf_idx;        // per pair or nullptr
    double* out;
    int stage_cap;                 // members (A + B) staged in shared memory per warp
};

// ---- launchers (all asynchronous on `st`; each returns the number of kernel launches it made) ----
int launch_convert_categories(const uint16_t* in, uint8_t* out, uint64_t n, int C, cudaStream_t st);
int launch_validate_xyz(const double* xyz, uint64_t n3, int* err, cudaStream_t st);
int launch_build_cells(const StructsView& s, double threshold, cudaStream_t st);
// sampled size probe: environment sizes of anchors 0, stride, 2*stride, ... (n_sample of them) -> count[n_sample]
int launch_env_count_sample(const StructsView& s, const KParams& p, uint64_t n_sample, uint64_t stride,
                            const uint32_t* anchor_struct, const uint32_t* anchor_prim, double threshold,
                            uint32_t* count, cudaStream_t st);
// gather + sort + store with cursor allocation; cap_class in {256, 512, 1024, 2048}
int launch_env_fill(const StructsView& s, const KParams& p, const uint32_t* anchor_struct,
                    const uint32_t* anchor_prim, double threshold, const EnvOut& out, int cap_class, cudaStream_t st);
// second stage for environments larger than cap_class (after the host saw stats.n_big > 0)
int launch_env_fill_big(const StructsView& s, const KParams& p, const uint32_t* anchor_struct,
                        const uint32_t* anchor_prim, double threshold, const EnvOut& out, int cap_class,
                        cudaStream_t st);
int launch_rows_fill(const double* dmx, const uint8_t* cat, uint64_t n_rows, uint64_t row_len, const double* xyz,
                     const KParams& p, const EnvOut& out, cudaStream_t st);
int launch_score(const ScoreArgs& args, const KParams& p, unsigned max_members_a, unsigned max_members_b,
                 cudaStream_t st);
int launch_job_means(const double* scores, const uint64_t* job_pair_off, uint64_t n_jobs, double* means,
                     cudaStream_t st);
int launch_anchor_lists(const KParams& p, const uint8_t* seq_a, uint64_t len_a, const double* da, const uint8_t* seq_b,
                        uint64_t len_b, const double* db, uint32_t wf_idx, double* out, cudaStream_t st);
int launch_wf_points(const WfDev* wf, uint64_t n, const double* x, double* out, int* err, cudaStream_t st);
int launch_sd_run(int kind, double q0, double q1, int C, uint64_t n, const double* p1, const double* p2, double* out,
                  cudaStream_t st);
int launch_fp64_peak(double* scratch, int blocks, int iters, cudaStream_t st);
int launch_fill_u64_iota_rows(uint64_t* off, uint32_t* count, uint64_t n_rows, uint64_t row_len, cudaStream_t st);

}  // namespace locohd
