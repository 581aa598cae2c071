// encode_ws.cu -- warp-specialised variant of the encode kernel for inputs that cannot fill
// the machine (a 64 MiB input is 8192 packets = 256 lane=packet warps for 592 warp schedulers).
//
// Same arithmetic as encode_kernel (coder_math.h), same output, but the stages of a symbol
// step run in four different warps of a 128-thread CTA that owns 32 packets:
//     warp 0  MODEL-A  tree levels 0-1              -> ring A (partial cum[s])
//     warp 1  MODEL-B  tree levels 2-3              -> ring B (partial cum[s] | count[s] << 16)
//     warp 2  CODER    narrow_renorm (the chain)    -> ring F (k | u << 5 | U1 << 9)
//     warp 3  BITS     emit_field / bit sink        -> the packet's slot
// (the tree levels are independent of each other given the symbol, so the model splits in
// two).  Each warp keeps lane = packet, so per-packet state never crosses lanes; the rings
// are double buffered per round of 32 positions and handed over with named barriers
// (bar.arrive on the producer side, bar.sync on the consumer side).  Only CODER carries the
// serial dependence of arithmetic coding.
#include "common.cuh"
#include "kernels.h"

#ifndef GPUAR_WS_CODER_UNROLL
#define GPUAR_WS_CODER_UNROLL 8     // tuning knob: steps per unrolled block of the CODER warp
#endif

namespace gpuar {

constexpr uint32_t kRound = 32;
constexpr int kCoderUnroll = GPUAR_WS_CODER_UNROLL;

struct WsShared {
    uint64_t tree[kTreeStored][32];      // 21504 B: nodes 0-3 MODEL-A, nodes 4-83 MODEL-B
    uint32_t ring_a[2][kRound][32];      //  8192 B, MODEL-A -> CODER
    uint32_t ring_b[2][kRound][32];      //  8192 B, MODEL-B -> CODER
    uint32_t ring_f[2][kRound][32];      //  8192 B, CODER   -> BITS
    uint32_t final_l[32];                // CODER -> BITS at the end of the packet
};

// named barriers (0 is __syncthreads); each is shared by exactly two warps = 64 threads
enum : uint32_t { kAFull = 1, kAEmpty = 3, kBFull = 5, kBEmpty = 7, kFFull = 9, kFEmpty = 11, kDone = 13 };

__device__ __forceinline__ void bar_sync(uint32_t id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t id)
{
    __threadfence_block();                                       // ring writes visible before the hand-over
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

// the packet's input, 16 bytes per lane per half round, fetched one half round ahead
struct SymbolFeed {
    const uint4 *in16;
    uint32_t len;
    uint4 buf_a, buf_b;
    __device__ __forceinline__ uint4 fetch(uint32_t g) const
    {
        return (g * 16u < len) ? __ldg(in16 + g) : make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ void start(const uint8_t *p, uint32_t n)
    {
        in16 = reinterpret_cast<const uint4 *>(p);
        len = n;
        buf_a = fetch(0);
        buf_b = fetch(1);
    }
};

__device__ __forceinline__ uint32_t word_of(const uint4 &c, uint32_t q)
{
    return q == 0u ? c.x : q == 1u ? c.y : q == 2u ? c.z : c.w;
}

// MODEL-A (kRole 0: root + level 1, ring A) or MODEL-B (kRole 1: level 2 + leaves, ring B)
template <int kRole>
__device__ __forceinline__ void model_warp(WsShared &sm, uint64_t *tree, const uint8_t *in, uint32_t len,
                                           uint32_t min_len, uint32_t rounds, uint32_t lane)
{
    uint64_t root = tree_node_init(64);                          // MODEL-A only
    if (kRole == 0) {
        for (uint32_t nd = 0; nd < 4u; ++nd) tree[nd * 32u] = tree_node_init(16);
    } else {
        for (uint32_t nd = 4u; nd < 20u; ++nd) tree[nd * 32u] = tree_node_init(4);
        for (uint32_t nd = 20u; nd < kTreeStored; ++nd) tree[nd * 32u] = enc_leaf_init();
    }
    uint32_t(*ring)[kRound][32] = kRole == 0 ? sm.ring_a : sm.ring_b;
    constexpr uint32_t full_id = kRole == 0 ? kAFull : kBFull, empty_id = kRole == 0 ? kAEmpty : kBEmpty;
    auto model = [&](uint32_t s) -> uint32_t {
        if (kRole == 0) return tree_encode_upper(root, tree, 32u, s);
        uint32_t cnt;
        const uint32_t lo = tree_encode_lower(tree, 32u, s, cnt);
        return lo | (cnt << 16);
    };
    SymbolFeed feed;
    feed.start(in, len);
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t b = r & 1u, i0 = r * kRound;
        if (r >= 2u) bar_sync(empty_id + b);
        const bool full = i0 + kRound <= min_len;
        auto half = [&](const uint4 &c, uint32_t h) {
#pragma unroll 1
            for (uint32_t q = 0; q < 4u; ++q) {
                const uint32_t word = word_of(c, q);
                const uint32_t j0 = 16u * h + 4u * q;
                if (full) {
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) ring[b][j0 + j][lane] = model((word >> (8u * j)) & 0xFFu);
                } else {
#pragma unroll 1
                    for (uint32_t j = 0; j < 4u; ++j)
                        ring[b][j0 + j][lane] = (i0 + j0 + j < len) ? model((word >> (8u * j)) & 0xFFu) : 0u;
                }
            }
        };
        half(feed.buf_a, 0u);
        feed.buf_a = feed.fetch(2u * r + 2u);
        half(feed.buf_b, 1u);
        feed.buf_b = feed.fetch(2u * r + 3u);
        bar_arrive(full_id + b);
    }
}

__global__ void __launch_bounds__(128)
encode_ws_kernel(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ slots, uint32_t slot_stride,
                 uint32_t *__restrict__ sizes, uint32_t n_packets, uint32_t packet)
{
    __shared__ __align__(16) WsShared sm;
    const uint32_t lane = lane_id();
    const uint32_t role = threadIdx.x >> 5;
    const uint32_t my = blockIdx.x * 32u + lane;
    const bool mine = my < n_packets;
    const size_t off = (size_t)my * packet;
    uint32_t len = 0;
    if (mine) len = (n - off < packet) ? (uint32_t)(n - off) : packet;
    const uint32_t max_len = __reduce_max_sync(kFull, len);      // identical in the four warps
    const uint32_t min_len = __reduce_min_sync(kFull, mine ? len : packet);
    const uint32_t rounds = (max_len + kRound - 1u) / kRound;
    uint64_t *const tree = &sm.tree[0][lane];

    if (role == 0u) {
        model_warp<0>(sm, tree, src + off, len, min_len, rounds, lane);
    } else if (role == 1u) {
        model_warp<1>(sm, tree, src + off, len, min_len, rounds, lane);
    } else if (role == 2u) {
        // ------------------------------------------------------------ CODER
        uint32_t L = 0, V = 0;
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t b = r & 1u, i0 = r * kRound;
            uint32_t sh;
            const uint32_t m_l = magic_for(256u + i0 + lane, sh);
            sh = shift_for(256u + i0);
            const bool full = i0 + kRound <= min_len;
            bar_sync(kAFull + b);
            bar_sync(kBFull + b);
            if (r >= 2u) bar_sync(kFEmpty + b);
            if (full) {
#pragma unroll kCoderUnroll
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t m = __shfl_sync(kFull, m_l, j);
                    const uint32_t pa = sm.ring_a[b][j][lane], pb = sm.ring_b[b][j][lane];
                    uint32_t k, u, U1;
                    const uint32_t lo = pa + (pb & 0xFFFFu);
                    narrow_renorm(L, V, lo, lo + (pb >> 16), m, sh, k, u, U1);
                    sm.ring_f[b][j][lane] = k | (u << 5) | (U1 << 9) | 0x80000000u;
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t m = __shfl_sync(kFull, m_l, j);
                    const uint32_t pa = sm.ring_a[b][j][lane], pb = sm.ring_b[b][j][lane];
                    uint32_t f = 0;
                    if (i0 + j < len) {
                        uint32_t k, u, U1;
                        const uint32_t lo = pa + (pb & 0xFFFFu);
                        narrow_renorm(L, V, lo, lo + (pb >> 16), m, sh, k, u, U1);
                        f = k | (u << 5) | (U1 << 9) | 0x80000000u;
                    }
                    sm.ring_f[b][j][lane] = f;
                }
            }
            if (r + 2u < rounds) {
                bar_arrive(kAEmpty + b);
                bar_arrive(kBEmpty + b);
            }
            bar_arrive(kFFull + b);
        }
        sm.final_l[lane] = L;
        bar_arrive(kDone);
    } else {
        // ------------------------------------------------------------ BITS
        uint32_t pend = 0;
        uint8_t *const slot = slots + (size_t)my * slot_stride;
        BitSink out;
        out.acc = 0;
        out.nb = 0;
        out.widx = 0;
        out.wcap = mine ? ((slot_stride - kHdr) >> 2) : 0u;
        out.words = reinterpret_cast<uint32_t *>(slot + kHdr);
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t b = r & 1u, i0 = r * kRound;
            const bool full = i0 + kRound <= min_len;
            bar_sync(kFFull + b);
            if (full) {
                // four steps per block.  The rare long-underflow path needs pend > 16 at a step
                // with k != 0; pend grows by at most the u's of the group, so one warp-uniform
                // vote on (pend + sum of u) covers the whole group and the common block has no
                // branch at all: its four ring loads issue together.
#pragma unroll 1
                for (uint32_t j0 = 0; j0 < kRound; j0 += 4u) {
                    uint32_t f[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) f[j] = sm.ring_f[b][j0 + j][lane];
                    const uint32_t usum = ((f[0] >> 5) & 15u) + ((f[1] >> 5) & 15u) + ((f[2] >> 5) & 15u) +
                                          ((f[3] >> 5) & 15u);
                    if (__any_sync(kFull, pend + usum > 16u)) {
#pragma unroll 1
                        for (uint32_t j = 0; j < 4u; ++j)         // rare path: small code; f[] spills to 16 B of stack here only
                            emit_symbol(out, pend, f[j] & 31u, (f[j] >> 5) & 15u, (f[j] >> 9) & 0xFFFFu);
                    } else {
#pragma unroll
                        for (uint32_t j = 0; j < 4u; ++j)
                            emit_field(out, pend, f[j] & 31u, (f[j] >> 5) & 15u, (f[j] >> 9) & 0xFFFFu);
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t f = sm.ring_f[b][j][lane];
                    if (f >> 31) emit_symbol(out, pend, f & 31u, (f >> 5) & 15u, (f >> 9) & 0xFFFFu);
                }
            }
            if (r + 2u < rounds) bar_arrive(kFEmpty + b);
        }
        bar_sync(kDone);
        if (mine) {
            const uint32_t comp = finish_packet(out, sm.final_l[lane], pend, slot, len);
            if (sizes) sizes[my] = comp;
        }
    }
}

cudaError_t launch_encode_slots_ws(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                   uint32_t *d_sizes, uint32_t packet, cudaStream_t st)
{
    const uint32_t packets = (uint32_t)((n + packet - 1) / packet);
    if (!packets) return cudaSuccess;
    encode_ws_kernel<<<(packets + 31u) / 32u, 128, 0, st>>>(d_in, n, d_slots, slot_stride, d_sizes, packets, packet);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
