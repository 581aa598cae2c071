// lookback.cuh -- single-pass chained scan (decoupled look-back, Merrill & Garland 2016)
// over per-tile totals.  One 64-bit descriptor per tile: flag (2 bits) | value (62 bits),
// written and read with single relaxed 64-bit accesses, so flag and value travel together.
// Tiles must take their index from an atomic ticket so that every predecessor is already
// running (forward progress), and the descriptor array must be zeroed before the launch.
#pragma once
#include "common.cuh"

namespace gpuar {

constexpr uint64_t kFlagAgg = 1ull << 62;   // value = this tile's total
constexpr uint64_t kFlagPfx = 2ull << 62;   // value = inclusive prefix through this tile
constexpr uint64_t kValMask = (1ull << 62) - 1ull;

__device__ __forceinline__ uint64_t ld_desc(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp.  Publishes `total` for `tile`, returns the sum of the totals
// of all earlier tiles (valid in every lane).
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *desc, uint32_t tile, uint64_t total, uint32_t lane)
{
    if (tile == 0) {
        if (lane == 0) st_desc(&desc[0], kFlagPfx | total);
        return 0;
    }
    if (lane == 0) st_desc(&desc[tile], kFlagAgg | total);
    uint64_t base = 0;
    int64_t look = (int64_t)tile - 1;
    for (;;) {
        const int64_t idx = look - (int64_t)lane;                // lane 0 inspects the nearest tile
        uint64_t d = kFlagPfx;                                   // before tile 0: prefix 0
        if (idx >= 0) {
            do { d = ld_desc(&desc[idx]); } while ((d >> 62) == 0);
        }
        const uint32_t is_pfx = __ballot_sync(kFull, (d >> 62) == 2u);
        const uint32_t stop = is_pfx ? (uint32_t)__ffs(is_pfx) - 1u : 32u;
        uint64_t v = (lane <= stop) ? (d & kValMask) : 0ull;    // aggregates up to the first prefix
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        base += v;
        if (is_pfx) break;
        look -= 32;
    }
    if (lane == 0) st_desc(&desc[tile], kFlagPfx | (base + total));
    return base;
}

}  // namespace gpuar
