// Exhaustive host-side check of the tile kernel's unit order (TileOrder / tile_unit, locohd_kernels.cuh) at production
// sizes: every (tile, anchor) exactly once, tile < n_tiles, anchor < n, slice-major, 32-bit form == 64-bit form.
//   nvcc -O2 -std=c++17 -I include -o /tmp/tile_order_check tools/tile_order_check.cu && /tmp/tile_order_check
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../loco_hd_b200/csrc/locohd_kernels.cuh"
using namespace locohd;

static int check(uint64_t n_tiles, uint64_t n, uint64_t slice) {
    const TileOrder o = make_tile_order(n_tiles, n, slice);
    const uint64_t units = n_tiles * n;
    std::vector<uint8_t> seen(units, 0);
    uint64_t prev_slice = 0, prev_tile = 0;
    for (uint64_t u = 0; u < units; ++u) {
        uint64_t t, p;
        tile_unit<uint64_t>(o, u, &t, &p);
        if (t >= n_tiles || p >= n) { std::printf("out of range at %llu\n", (unsigned long long)u); return 1; }
        if ((units >> 32) == 0) {
            uint32_t t32, p32;
            tile_unit<uint32_t>(o, (uint32_t)u, &t32, &p32);
            if (t32 != t || p32 != p) { std::printf("32-bit form differs at %llu\n", (unsigned long long)u); return 1; }
        }
        if (seen[t * n + p]++) { std::printf("duplicate at %llu\n", (unsigned long long)u); return 1; }
        const uint64_t s = p / o.slice;
        if (s < prev_slice || (s == prev_slice && t < prev_tile)) { std::printf("order broken at %llu\n", (unsigned long long)u); return 1; }
        prev_slice = s; prev_tile = t;
    }
    std::printf("ok: %llu tiles x %llu anchors, slice %llu (%llu units, last slice %llu)\n", (unsigned long long)n_tiles,
                (unsigned long long)n, (unsigned long long)o.slice, (unsigned long long)units, (unsigned long long)o.last);
    return 0;
}

int main() {
    int bad = 0;
    bad |= check(31219, 5000, 16);    // the 1000-structure ensemble on one GPU
    bad |= check(31219, 5000, 24);
    bad |= check(3903, 5000, 16);     // one rank's share on 8 GPUs
    bad |= check(7813, 5000, 40);     // the 500-structure ensemble
    bad |= check(300, 5000, 216);     // 96 structures
    bad |= check(21, 203, 8);
    bad |= check(21, 203, 0);
    bad |= check(1, 5000, 16);
    bad |= check(900000, 5000, 16);   // > 2^32 units: 64-bit form only
    return bad;
}
