// decode.cu -- sm_100a decode path, lane = packet.
//
// Replaces garDecompress / arDecompress (reference src/gpuar_kernel.cu:848-892,
// 916-934).  Per symbol the reference does: getUnscaledCode (:703-716, a divide by
// the current range), getSymbolFromProbability (:727-763, binary search over Fenwick
// prefix sums, ~140 dependent shared loads), applySymbolRange (:256-288, two more
// Fenwick sums + update) and readEncodedBits (:787-836, bit-at-a-time).
//
// Here the adaptive model of each packet is a 4-ary cumulative-count tree in shared
// memory, interleaved so that lane l only ever touches its own banks:
//   level 0: 1 node  (children span 64 symbols)      node = four u16 slots (0, t0, t1, t2):
//   level 1: 4 nodes (16)                            t0 = |child0|, t1 = t0+|child1|,
//   level 2: 16 nodes (4)                            t2 = t1+|child2|
//   level 3: 64 leaves (1)                           leaf = inclusive prefix sums (s0, s1, s2, s3)
// One branch-free descent (root in registers, then 3 dependent 8-byte shared loads) finds
// the symbol, yields cum[s] and cum[s+1] and applies the model update (+1 on every slot
// right of the path) in the same pass.  The symbol step itself -- quotient, descent, interval
// narrowing with its single normalisation, next bits -- is decode_math.h (decode_step for the
// throughput variant, decode_step_latency for the latency variant; state = code - lower bound,
// lower bound, range); bits come from a 64-bit reservoir fed by 32-bit words.
#include <cstdlib>

#include "common.cuh"
#include "decode_math.h"
#include "kernels.h"

#ifndef GPUAR_DEC_UNROLL
#define GPUAR_DEC_UNROLL 4          // steps per unrolled block of the full-round loop (tuning knob)
#endif

namespace gpuar {

#ifndef GPUAR_DEC_UNROLL_LAT
#define GPUAR_DEC_UNROLL_LAT 8      // the same for the latency variant (a lone warp: 64 MiB decode 1.382 ms with 4, 1.350 with 8; 16 overflows the instruction cache)
#endif
constexpr int kDecUnroll = GPUAR_DEC_UNROLL;
constexpr int kDecUnrollLat = GPUAR_DEC_UNROLL_LAT;

__device__ __forceinline__ const uint32_t *clamp_ptr(const uint32_t *p, const uint32_t *last)
{
    return p < last ? p : last;          // reads past the buffer are pinned to its last word
}

constexpr uint32_t kRing = 8;        // per-lane ring of stream words in shared memory (latency variant)
constexpr uint32_t kRingSmall = 4;   // ... of the throughput variant: 512 B is what 10 CTAs per SM leave free
constexpr uint32_t kAhead = 3;       // words requested ahead of the reader

// kernel variants (template parameter kLatency): throughput (decode_step) and latency (decode_step_latency)
template <bool kLatency>
struct DecShared;
template <>
struct DecShared<false> {                    // throughput variant: 22016 B keeps 10 CTAs per SM
    uint64_t tree[kTreeStored][32];          // lane l owns column l (banks 2l, 2l+1); root in registers
    uint32_t ring[kRingSmall][32];           // 512 B stream ring
};
template <>
struct DecShared<true> {                     // latency variant: 23.5 KB, used up to 8 CTAs per SM (9 would fit)
    Quad l1[4][32];                          // level-1 thresholds as 32-bit words (16 B per lane and node)
    uint64_t l2[16][32];                     // packed (0, t0, t1, t2)
    uint64_t l3[64][32];                     // leaves: packed inclusive sums
    uint32_t ring[kRing][32];                // 1024 B stream ring
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

// The bit window of both variants is refilled from a per-lane ring of stream words in shared memory that cp.async
// (LDGSTS: no register, no scoreboard) keeps kAhead words ahead of the reader: nothing in a step ever waits on a
// global load and the refill is predicated, never divergent.  (Round 1 fed the throughput variant from one word
// prefetched in a register: the compiler had sunk that load to within a step of its use and 6 % of the warp time
// waited for it; a 512-byte ring is what 10 CTAs per SM leave free: 4 GiB decode 27.1 -> 26.5 ms.)
template <bool kLatency>
__global__ void __launch_bounds__(32)
decode_kernel(const uint8_t *__restrict__ payload, size_t readable, const uint64_t *__restrict__ offsets,
              uint32_t stride, uint32_t n_packets, uint8_t *__restrict__ out, uint32_t packet,
              const uint64_t *__restrict__ count)
{
    __shared__ __align__(16) DecShared<kLatency> sm;
    constexpr uint32_t kRingWords = kLatency ? kRing : kRingSmall;
    const uint32_t lane = lane_id();
    const uint32_t my = blockIdx.x * 32u + lane;
    if (count) {                                    // sharded decode: the chain discovery left the count on the device
        n_packets = (uint32_t)min((uint64_t)n_packets, *count);
        if (blockIdx.x * 32u >= n_packets) return;
    }
    const bool mine = my < n_packets;

    // model init: every count 1 (gpuar_kernel.cu:403-419)
    uint64_t root = 0;
    TopLevels top;                               // latency variants: the root's thresholds, level-1 copy
    uint64_t *tree = nullptr;
    LatTree lat{nullptr, nullptr, nullptr, 32u};
    if constexpr (kLatency) {
        lat = LatTree{&sm.l1[0][lane], &sm.l2[0][lane], &sm.l3[0][lane], 32u};
        lat_tree_init(top, lat);
    } else {
        tree = &sm.tree[0][lane];
        dec_tree_init(root, tree, 32u);
    }

    // bit source: aligned 32-bit words of the packet's bitstream.  The window is fed from a
    // per-lane ring in shared memory that cp.async keeps kAhead words ahead of the reader.
    const uint32_t *const wend = reinterpret_cast<const uint32_t *>(payload) + (readable >> 2) - 1;  // last readable word
    const uint32_t *gp = wend;      // next word to request
    uint32_t rd = 0;                // words consumed from the ring
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
        const uint32_t dst = ring_s + ((p & (kRingWords - 1u)) << 7);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g0 + min(p, p_max)) : "memory");
    };
    for (uint32_t p = 0; p < kAhead; ++p) request(p);
    cp_async_commit();
    cp_async_wait<0>();
    auto refill = [&]() {
        // One refill per two steps (below), predicated, never divergent: feed the ring's next word if the
        // window has room, and request one more word.  A word is read at least kAhead refills
        // after it was requested, so waiting for all but the kAhead-1 newest groups makes it
        // visible.
        cp_async_wait<kAhead - 1>();
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(ring_s + ((rd & (kRingWords - 1u)) << 7)) : "memory");
        const bool h = in.hungry();
        in.feed_if(h, bswap32(w));
        if (h) request(rd + kAhead);
        cp_async_commit();
        rd += h ? 1u : 0u;
    };
    // initializeDecoder (:582-603): the first 16 bits; lower bound 0, range 2^16
    DecState st;
    st.D = in.take(16u);
    st.L = 0;
    st.R = 65536u;
    refill();

    const uint32_t max_raw = __reduce_max_sync(kFull, raw);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)my * packet);
    uint32_t packed = 0;

    // one symbol of this lane's packet; `slot` = position of the byte inside the 32-bit store word
    auto step = [&](uint32_t i, uint32_t m, uint32_t sh, uint32_t slot) {
        const uint32_t T = 256u + i;
        // latency variant: every level decided by multiplication, speculative node loads; throughput variant: the
        // quotient first, then the plain tree (the other way round was measured on both, profiles/r2_decode_v2.md)
        uint32_t s;
        if constexpr (kLatency) s = decode_step_latency(st, top, lat, T, m, sh, in);
        else s = decode_step(st, root, tree, 32u, T, m, sh, in);
        packed = mad32(s, 1u << (8u * slot), packed);               // fields cannot overlap: a multiply-add, not shift + or
        // A step takes at most 16 bits and the window holds at least 33 after a refill: one refill (at most one
        // word) every SECOND step keeps at least 17 bits in front of every step and restores 33.
        if (slot & 1u) refill();
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
            constexpr int kUnroll = kLatency ? kDecUnrollLat : kDecUnroll;
#pragma unroll kUnroll
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

// Device self-check of the decoder's float-estimated quotient (decode_math.h: divide_floor): every range the
// coder can hold (2^14 < range <= 2^16), every quotient up to 2^14 - 1, numerators at both ends and in the
// middle of each quotient's interval.  The approximate reciprocal only exists on the device, so this is where
// the one-sided correction is pinned (the host model checks the same lattice with a divide in its place).
__global__ void __launch_bounds__(256)
selfcheck_divide_kernel(unsigned long long *__restrict__ bad)
{
    const uint32_t range = 16385u + blockIdx.x;
    unsigned long long n = 0;
    for (uint32_t q = threadIdx.x; q < 16384u; q += blockDim.x) {
        const uint32_t base = q * range;                            // (q + 1) * range - 1 < 2^30
        n += divide_floor(base, range) != q;
        n += divide_floor(base + (range >> 1), range) != q;
        n += divide_floor(base + range - 1u, range) != q;
        n += divide_exact(base + range - 1u, range) != q;
    }
    if (n) atomicAdd(bad, n);
}

cudaError_t launch_selfcheck(uint64_t *d_mismatches, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_mismatches, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    selfcheck_divide_kernel<<<65536u - 16384u, 256, 0, st>>>(reinterpret_cast<unsigned long long *>(d_mismatches));
    count_launch();
    return cudaGetLastError();
}

static int g_decode_path = 0;        // 0 auto, 1 latency variant, 2 throughput variant
bool set_decode_path(int path)
{
    if (path < 0 || path > 2) return false;
    g_decode_path = path;
    return true;
}

cudaError_t launch_decode(const uint8_t *d_payload, size_t readable, const uint64_t *d_offsets, uint32_t stride,
                          uint32_t packets, uint8_t *d_out, uint32_t packet, cudaStream_t st, const uint64_t *d_count)
{
    if (!packets) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t grid = (packets + 31u) / 32u;
    // The latency-optimised variant pays for its short chain with more instructions per step and more shared-memory
    // traffic.  Measured (profiles/r2_decode_v2.md): it wins up to two warps per scheduler (8 CTAs per SM: 296 MiB
    // 2.18 against 2.33 ms; 148 MiB 1.28 against 2.05 ms) and loses from 10 CTAs per SM on, which no longer fit in one
    // wave of its 23.5 KB of shared memory (370 MiB: 3.35 against 2.82 ms).  GPUAR_OPT_DECODE_PATH forces one of them
    // (tests run both on the same streams); GPUAR_B200_DEC_RING_MAX=<CTAs> moves the switch (tuning aid).
    static const long forced = [] { const char *e = getenv("GPUAR_B200_DEC_RING_MAX"); return e && *e ? atol(e) : -1L; }();
    uint32_t ring_max = forced >= 0 ? (uint32_t)forced : (uint32_t)sms * 8u;
    if (g_decode_path == 1) ring_max = 0xFFFFFFFFu;
    if (g_decode_path == 2) ring_max = 0u;
    if (grid <= ring_max)
        decode_kernel<true><<<grid, 32, 0, st>>>(d_payload, readable, d_offsets, stride, packets, d_out, packet, d_count);
    else
        decode_kernel<false><<<grid, 32, 0, st>>>(d_payload, readable, d_offsets, stride, packets, d_out, packet, d_count);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
