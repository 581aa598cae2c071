// decode.cu -- sm_100a decode path, lane = packet.
//
// Replaces garDecompress / arDecompress (reference src/gpuar_kernel.cu:848-892,
// 916-934).  Per symbol the reference does: getUnscaledCode (:703-716, a divide by
// the current range), getSymbolFromProbability (:727-763, binary search over Fenwick
// prefix sums, ~140 dependent shared loads), applySymbolRange (:256-288, two more
// Fenwick sums + update) and readEncodedBits (:787-836, bit-at-a-time).
//
// Here the adaptive model of each packet is a 4-ary cumulative-count tree in shared
// memory, interleaved so that lane l only ever touches bank pair (2l, 2l+1):
//   level 0: 1 node  (children span 64 symbols)      node = four u16 slots (0, t0, t1, t2):
//   level 1: 4 nodes (16)                            t0 = |child0|, t1 = t0+|child1|,
//   level 2: 16 nodes (4)                            t2 = t1+|child2|
//   level 3: 64 nodes (1)
// One branch-free descent (root in registers, then 3 dependent 8-byte shared loads) finds
// the symbol, yields cum[s] and count[s] and applies the model update (+1 on every slot
// right of the path) in the same pass (coder_math.h: tree_level).  Interval arithmetic and
// renormalisation are the single-normalisation step of coder_math.h (narrow_total: state =
// lower bound and range); bits come from a 64-bit reservoir fed by 32-bit words.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

#ifndef GPUAR_DEC_UNROLL
#define GPUAR_DEC_UNROLL 4          // steps per unrolled block of the full-round loop (tuning knob)
#endif

namespace gpuar {

constexpr int kDecUnroll = GPUAR_DEC_UNROLL;

__device__ __forceinline__ const uint32_t *clamp_ptr(const uint32_t *p, const uint32_t *last)
{
    return p < last ? p : last;          // reads past the buffer are pinned to its last word
}

constexpr uint32_t kRing = 8;        // per-lane ring of stream words in shared memory
constexpr uint32_t kAhead = 3;       // words requested ahead of the reader

template <bool kRingFeed>
struct DecShared {
    uint64_t tree[kTreeStored][32];          // 21504 B; lane l owns column l (banks 2l, 2l+1); root in registers
    uint32_t ring[kRingFeed ? kRing : 1][32];  // 1024 B stream ring (ring feed only: 21504 B keeps 10 CTAs per SM)
};

// 4-byte asynchronous global->shared copy (LDGSTS): no register, no scoreboard -- the stream
// prefetch never stalls the dependent chain of the decoder.
__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gmem_src)
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// kRingFeed selects how the bit window is refilled:
//   true   cp.async ring + predicated feed: nothing in the step ever waits on a global load and
//          there is no divergent branch -- shortest dependent chain; used while every warp has a
//          scheduler to itself (up to 4 warps per SM = 148 MiB: latency is everything);
//   false  one word prefetched in a register, fed under a (divergent) branch -- fewer
//          instructions per step; used when many warps per scheduler hide the latency.
template <bool kRingFeed>
__global__ void __launch_bounds__(32)
decode_kernel(const uint8_t *__restrict__ payload, size_t readable, const uint64_t *__restrict__ offsets,
              uint32_t stride, uint32_t n_packets, uint8_t *__restrict__ out, uint32_t packet,
              const uint64_t *__restrict__ count)
{
    __shared__ __align__(16) DecShared<kRingFeed> sm;
    const uint32_t lane = lane_id();
    const uint32_t my = blockIdx.x * 32u + lane;
    if (count) {                                    // sharded decode: the chain discovery left the count on the device
        n_packets = (uint32_t)min((uint64_t)n_packets, *count);
        if (blockIdx.x * 32u >= n_packets) return;
    }
    const bool mine = my < n_packets;

    // model init: every count 1 (gpuar_kernel.cu:403-419)
    uint64_t *const tree = &sm.tree[0][lane];
    uint64_t root;
    tree_init(root, tree, 32u);

    // bit source: aligned 32-bit words of the packet's bitstream.  The window is fed from a
    // per-lane ring in shared memory that cp.async keeps kAhead words ahead of the reader.
    const uint32_t *const wend = reinterpret_cast<const uint32_t *>(payload) + (readable >> 2) - 1;  // last readable word
    const uint32_t *gp = wend;      // next word to request (ring) / word held in `ahead` (register)
    uint32_t rd = 0;                // ring feed: words consumed
    uint32_t ahead = 0;             // register feed: prefetched word, byte-swapped only when fed
    BitSource in;
    in.start(0, 64u);
    uint32_t raw = 0;
    if (mine) {
        const size_t off = offsets ? (size_t)offsets[my] : (size_t)my * stride;
        const uint32_t hdr = (uint32_t)payload[off + 2] | ((uint32_t)payload[off + 3] << 8);   // rawLen, :859
        raw = min(hdr, packet);
        const size_t sp = off + kHdr;                              // first bitstream byte, any alignment
        gp = reinterpret_cast<const uint32_t *>(payload) + (sp >> 2);
        const uint32_t skip = 8u * (uint32_t)(sp & 3u);
        const uint64_t w0 = bswap32(*clamp_ptr(gp, wend));
        ++gp;
        const uint64_t w1 = bswap32(*clamp_ptr(gp, wend));
        ++gp;
        in.start(((w0 << 32) | w1) << skip, 64u - skip);           // 40..64 bits
    }
    // ring feed: ring position p holds stream word g0[min(p, p_max)] at shared address
    // ring_s + (p mod kRing) * 128; positions rd .. rd + kAhead - 1 are requested, so one counter
    // and 32-bit index arithmetic do (a 64-bit pointer to bump and clamp costs four instructions more)
    const uint32_t *const g0 = clamp_ptr(gp, wend);
    const uint32_t p_max = (uint32_t)min((ptrdiff_t)(wend - g0), (ptrdiff_t)0x3FFFFFFF);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(&sm.ring[0][lane]);
    auto request = [&](uint32_t p) {
        const uint32_t dst = ring_s + ((p & (kRing - 1u)) << 7);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g0 + min(p, p_max)) : "memory");
    };
    if (kRingFeed) {
        for (uint32_t p = 0; p < kAhead; ++p) request(p);
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        ahead = *clamp_ptr(gp, wend);
    }
    auto refill = [&]() {
        if (kRingFeed) {
            // One refill per step, predicated, never divergent: feed the ring's next word if the
            // window has room, and request one more word.  A word is read at least kAhead steps
            // after it was requested, so waiting for all but the kAhead-1 newest groups makes it
            // visible.
            cp_async_wait<kAhead - 1>();
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(ring_s + ((rd & (kRing - 1u)) << 7)) : "memory");
            const bool h = in.hungry();
            in.feed_if(h, bswap32(w));
            if (h) request(rd + kAhead);
            cp_async_commit();
            rd += h ? 1u : 0u;
        } else if (in.hungry()) {
            in.feed(bswap32(ahead));
            ++gp;
            ahead = *clamp_ptr(gp, wend);
        }
    };
    // initializeDecoder (:582-603): the first 16 bits
    uint32_t code = in.take(16u);
    refill();
    uint32_t L = 0, R = 65536u;                     // narrow_total state: lower bound and range

    const uint32_t max_raw = __reduce_max_sync(kFull, raw);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)my * packet);
    uint32_t packed = 0;

    // one symbol of this lane's packet; `slot` = position of the byte inside the 32-bit store word
    auto step = [&](uint32_t i, uint32_t m, uint32_t sh, uint32_t slot) {
        const uint32_t T = 256u + i;
        uint32_t lo, cnt;
        // latency variant: top levels decided by multiplication with speculative node loads; throughput variant: the
        // quotient first, then the plain tree (the other way round was measured on both, profiles/r2_kernel_experiments.md)
        const uint32_t s = kRingFeed ? tree_decode_early_range(root, tree, 32u, code, L, R, T, lo, cnt)
                                     : tree_decode(root, tree, 32u, unscale_range(code, L, R, T), T, lo, cnt);
        packed |= s << (8u * slot);
        uint32_t L1, S1, t, As;
        narrow_total(L, R, lo, lo + cnt, m, sh, L1, S1, t, As);
        code = advance_code_total(code, t, As, in);
        refill();
    };

    const uint32_t min_raw = __reduce_min_sync(kFull, mine ? raw : packet);
    const uint32_t rounds = (max_raw + 31u) >> 5;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i0 = r * 32u;
        uint32_t sh;
        const uint32_t m_l = magic_for(256u + i0 + lane, sh);      // lane j holds the multiplier of step j
        sh = shift_for(256u + i0);                                 // the shift is uniform over the round
        if (i0 + 32u <= min_raw) {
            // every lane of the warp has all 32 positions: no per-lane predicates
#pragma unroll kDecUnroll
            for (uint32_t j = 0; j < 32u; ++j) {
                step(i0 + j, __shfl_sync(kFull, m_l, j), sh, j & 3u);
                if ((j & 3u) == 3u) {
                    if (mine) dst[(i0 + j) >> 2] = packed;
                    packed = 0;
                }
            }
        } else {
            // ragged tail (only the warp holding the last packet of the stream gets here)
            for (uint32_t j = 0; j < 32u; ++j) {
                const uint32_t m = __shfl_sync(kFull, m_l, j);
                const uint32_t i = i0 + j;
                if (i < raw) step(i, m, sh, j & 3u);
                if ((j & 3u) == 3u) {
                    if (mine && (i & ~3u) < raw) {
                        if (i < raw) {
                            dst[i >> 2] = packed;
                        } else {
                            uint8_t *b = reinterpret_cast<uint8_t *>(dst) + (i & ~3u);
                            for (uint32_t q = 0; q < (raw & 3u); ++q) b[q] = (uint8_t)(packed >> (8u * q));
                        }
                    }
                    packed = 0;
                }
            }
        }
    }
}

cudaError_t launch_decode(const uint8_t *d_payload, size_t readable, const uint64_t *d_offsets, uint32_t stride,
                          uint32_t packets, uint8_t *d_out, uint32_t packet, cudaStream_t st, const uint64_t *d_count)
{
    if (!packets) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t grid = (packets + 31u) / 32u;
    // tuning aid: GPUAR_B200_DEC_RING_MAX=<CTAs> moves the switch between the two variants
    static const long forced = [] { const char *e = getenv("GPUAR_B200_DEC_RING_MAX"); return e && *e ? atol(e) : -1L; }();
    // The latency-optimised variant pays for its short chain with 246 instructions per step (the
    // other one: 177).  It wins while every warp has a scheduler to itself (4 per SM: 1.67 against
    // 2.55 ms at 128 MiB) and loses as soon as two warps share one (2.95 against 2.72 ms at 192 MiB,
    // 4.41 against 3.30 ms at 368 MiB; profiles/r1_s2_dec_switch.txt).
    const uint32_t ring_max = forced >= 0 ? (uint32_t)forced : (uint32_t)sms * 4u;
    if (grid <= ring_max)
        decode_kernel<true><<<grid, 32, 0, st>>>(d_payload, readable, d_offsets, stride, packets, d_out, packet, d_count);
    else
        decode_kernel<false><<<grid, 32, 0, st>>>(d_payload, readable, d_offsets, stride, packets, d_out, packet, d_count);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
