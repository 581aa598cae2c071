// encode_ws.cu -- warp-specialised variant of the encode kernel for inputs that cannot fill
// the machine (a 64 MiB input is 8192 packets = 256 lane=packet warps for 592 warp schedulers).
//
// Same arithmetic as encode_kernel (coder_math.h), same output, but the stages of a symbol
// step run in six different warps of a 192-thread CTA that owns 32 packets:
//     warp 0  MODEL-A  tree levels 0-1               -> ring A (partial cum[s])
//     warp 1  MODEL-B  tree level 2                  -> ring B (partial cum[s])
//     warp 2  MODEL-D  tree level 3 (the leaves)     -> ring D (partial cum[s] | count[s] << 16)
//     warp 3  FIELD    k, u and the bit field        -> ring F (pack_field descriptor)
//     warp 4  BITS     bit sink                      -> the packet's slot
//     warp 5  CODER    narrow_lazy (the chain)       -> ring C (L1 | U1 << 16)
// The tree levels are independent of each other given the symbol, so the model splits by
// level.  Only CODER carries the serial dependence of arithmetic coding, and it carries the
// minimum: the interval recurrence with its single normalisation (narrow_lazy); how the
// total shift splits into matching-MSB and underflow shifts -- two count-leading-zeros --
// is worked out by FIELD, which has no state at all, and everything of the emission that
// depends on the pending-underflow counter by BITS.  CODER is the highest warp of the CTA
// because the SM sub-partition arbiter favours the highest warp slot.
// Each warp keeps lane = packet, so per-packet state never crosses lanes; the rings are
// double buffered per round of 32 positions and handed over with named barriers
// (bar.arrive on the producer side, bar.sync on the consumer side).  The three model rings
// share one full barrier per buffer (128 threads: three producers + CODER).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "shard.cuh"

#ifndef GPUAR_WS_CODER_BLOCK
#define GPUAR_WS_CODER_BLOCK 4      // tuning knob: steps per straight-line block of the CODER warp
#endif

namespace gpuar {

constexpr uint32_t kRound = 32;
constexpr uint32_t kCoderBlock = GPUAR_WS_CODER_BLOCK;
// Role placement.  The CTA is launched with six or eight warps and a role map says what each one
// does: nibble w of the map = role of warp w, 0xF = none (the warp exits at once).  Which roles
// share a scheduler (SM sub-partition; warp index modulo 4 for a CTA alone on its SM) matters:
// with one CTA per SM (up to 148 x 32 packets = 37 MiB, and every chunk of the host-buffer
// pipeline) eight warps with the CODER chain alone on its scheduler take 118 instead of 128
// cycles per step; with two or three CTAs per SM every eight-warp placement measured was slower
// than the plain six-warp CTA (0.65-0.80 against 0.624 ms at 64 MiB; how the hardware spreads a
// second CTA's warps over the sub-partitions is not the simple rule above), so that one stays
// (profiles/r1_s2_ws_rolemap*.jsonl).  GPUAR_B200_WS_TUNE="warps,map_first,map_later" (hex maps)
// overrides the choice for experiments; the CTAs of the first wave (one per SM) take map_first,
// the others map_later.
constexpr uint32_t kWsMaxThreads = 256;
constexpr uint32_t kMapSix = 0xFF543210u;       // warps 0-5 = roles 0-5: A+BITS | B+CODER | D | FIELD
constexpr uint32_t kMapCoderAlone = 0x5F43F210u; // A+FIELD | B+BITS | D | CODER
#ifndef GPUAR_WS_SWAP_BD
#define GPUAR_WS_SWAP_BD 0          // tuning knob: which of warps 1 / 2 takes level 2 and which the leaves
#endif
constexpr uint32_t kRoleB = GPUAR_WS_SWAP_BD ? 2u : 1u, kRoleD = GPUAR_WS_SWAP_BD ? 1u : 2u;

struct WsShared {
    uint64_t tree[kTreeStored][32];      // 21504 B: nodes 0-3 MODEL-A, 4-19 MODEL-B, 20-83 MODEL-D
    uint32_t ring_a[2][kRound][32];      //  8192 B each (MODEL -> CODER)
    uint32_t ring_b[2][kRound][32];
    uint32_t ring_d[2][kRound][32];
    uint32_t ring_c[2][kRound][32];      // CODER -> FIELD
    uint2 ring_f[2][kRound][32];         // FIELD -> BITS (16384 B: two-word descriptors)
    uint32_t stage[3][8][32];            // the model warps' input words of the current round
    uint32_t final_l[32];                // CODER -> BITS at the end of the packet
};

// Named barriers, id + buffer index; all sixteen are in use (the kernel has no __syncthreads,
// so barrier 0 is free).  A "full" barrier is arrived at by the producer(s) and waited on by the
// one consumer; an "empty" barrier the other way round.  The three model rings share their full
// barrier (three producers + CODER = 128 threads) but are handed back one by one, so that
// every barrier has exactly one waiting warp (which is also what compute-sanitizer's synccheck
// expects: it reports warps that wait on one barrier from different instructions as divergent).
enum : uint32_t { kInFull = 0, kInEmpty = 2 /* + 2 * role */, kCFull = 8, kCEmpty = 10, kFFull = 12, kFEmpty = 14 };
constexpr uint32_t kInCount = 128;       // MODEL-A, MODEL-B, MODEL-D, CODER
constexpr uint32_t kPairCount = 64;      // one producer warp + one consumer warp

template <uint32_t kCount>
__device__ __forceinline__ void bar_sync(uint32_t id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kCount) : "memory");
}
template <uint32_t kCount>
__device__ __forceinline__ void bar_arrive_raw(uint32_t id)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kCount) : "memory");
}
template <uint32_t kCount>
__device__ __forceinline__ void bar_arrive(uint32_t id)
{
    __threadfence_block();                                       // ring accesses ordered before the hand-over
    bar_arrive_raw<kCount>(id);
}

// kRole 0: root + level 1 -> ring A;  1: level 2 -> ring B;  2: leaves -> ring D.
// The packet's input arrives as two 16-byte loads per lane and round, issued one round ahead;
// at the top of a round the eight words go to the warp's staging area (stage[word][lane]: the
// lane's own column, so no synchronisation) and the symbol loop picks them up by index --
// selecting among eight registers with a run-time index would cost more than the model step.
template <int kRole>
__device__ __forceinline__ void model_warp(WsShared &sm, uint64_t *tree, const uint8_t *in, uint32_t len,
                                           uint32_t min_len, uint32_t rounds, uint32_t lane)
{
    uint64_t root = tree_node_init(64);                          // MODEL-A only
    if (kRole == 0) {
        for (uint32_t nd = 0; nd < 4u; ++nd) tree[nd * 32u] = tree_node_init(16);
    } else if (kRole == 1) {
        for (uint32_t nd = 4u; nd < 20u; ++nd) tree[nd * 32u] = tree_node_init(4);
    } else {
        for (uint32_t nd = 20u; nd < kTreeStored; ++nd) tree[nd * 32u] = enc_leaf_init();
    }
    uint32_t(*ring)[kRound][32] = kRole == 0 ? sm.ring_a : kRole == 1 ? sm.ring_b : sm.ring_d;
    uint32_t(*stage)[32] = sm.stage[kRole];
    auto fields = [&](uint32_t word) -> WordFields {
        return kRole == 0 ? word_fields_upper(word) : kRole == 1 ? word_fields_mid(word) : word_fields_leaf(word);
    };
    auto model = [&](WordFields f, uint32_t j) -> uint32_t {    // symbol j of the word the fields come from
        if (kRole == 0) return tree_encode_upper_word(root, tree, 32u, f, j);
        if (kRole == 1) return tree_encode_mid_word(tree, 32u, f, j);
        uint32_t cnt;
        const uint32_t lo = tree_encode_leaf_word(tree, 32u, f, j, cnt);
        return cnt * 65536u + lo;                                // fields cannot overlap: a multiply-add, not shift + or
    };
    const uint4 *const in16 = reinterpret_cast<const uint4 *>(in);
    auto fetch = [&](uint32_t g) -> uint4 {
        return (g * 16u < len) ? __ldg(in16 + g) : make_uint4(0, 0, 0, 0);
    };
    uint4 buf_a = fetch(0), buf_b = fetch(1);
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t b = r & 1u, i0 = r * kRound;
        stage[0][lane] = buf_a.x, stage[1][lane] = buf_a.y, stage[2][lane] = buf_a.z, stage[3][lane] = buf_a.w;
        stage[4][lane] = buf_b.x, stage[5][lane] = buf_b.y, stage[6][lane] = buf_b.z, stage[7][lane] = buf_b.w;
        buf_a = fetch(2u * r + 2u);
        buf_b = fetch(2u * r + 3u);
        if (r >= 2u) bar_sync<kPairCount>(kInEmpty + 2u * kRole + b);
        if (i0 + kRound <= min_len) {
#pragma unroll 1
            for (uint32_t w = 0; w < 8u; ++w) {
                const WordFields f = fields(stage[w][lane]);
#pragma unroll
                for (uint32_t j = 0; j < 4u; ++j) ring[b][4u * w + j][lane] = model(f, j);
            }
        } else {
#pragma unroll 1
            for (uint32_t j = 0; j < kRound; ++j) {
                const WordFields f = fields(stage[j >> 2][lane]);
                ring[b][j][lane] = (i0 + j < len) ? model(f, j & 3u) : 0u;
            }
        }
        bar_arrive<kInCount>(kInFull + b);
    }
}

__global__ void __launch_bounds__(kWsMaxThreads)
encode_ws_kernel(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ slots, uint32_t slot_stride,
                 uint32_t *__restrict__ sizes, uint32_t n_packets, uint32_t packet, uint32_t map_first,
                 uint32_t map_later, uint32_t first_ctas, ShardTarget tg, uint64_t *__restrict__ acc)
{
    extern __shared__ __align__(16) uint8_t ws_smem[];           // 72 KB: above the static limit
    WsShared &sm = *reinterpret_cast<WsShared *>(ws_smem);
    const uint32_t lane = lane_id();
    const uint32_t role = ((blockIdx.x < first_ctas ? map_first : map_later) >> (4u * (threadIdx.x >> 5))) & 0xFu;
    const uint32_t my = blockIdx.x * 32u + lane;
    const bool mine = my < n_packets;
    const size_t off = (size_t)my * packet;
    uint32_t len = 0;
    if (mine) len = (n - off < packet) ? (uint32_t)(n - off) : packet;
    const uint32_t max_len = __reduce_max_sync(kFull, len);      // identical in all the warps
    const uint32_t min_len = __reduce_min_sync(kFull, mine ? len : packet);
    const uint32_t rounds = (max_len + kRound - 1u) / kRound;
    uint64_t *const tree = &sm.tree[0][lane];

    if (role == 0u) {
        model_warp<0>(sm, tree, src + off, len, min_len, rounds, lane);
    } else if (role == kRoleB) {
        model_warp<1>(sm, tree, src + off, len, min_len, rounds, lane);
    } else if (role == kRoleD) {
        model_warp<2>(sm, tree, src + off, len, min_len, rounds, lane);
    } else if (role == 5u) {
        // ------------------------------------------------------------ CODER
        uint32_t L = 0, R1 = 65536u, sx = 0;                      // narrow_lazy state: range = R1 >> sx
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t b = r & 1u, i0 = r * kRound;
            uint32_t sh;
            const uint32_t m_l = magic_for(256u + i0 + lane, sh);
            sh = shift_for(256u + i0);
            const bool full = i0 + kRound <= min_len;
            bar_sync<kInCount>(kInFull + b);
            if (r >= 2u) bar_sync<kPairCount>(kCEmpty + b);
            if (full) {
                // blocks of kCoderBlock steps: everything that does not depend on the coder state
                // (ring loads, the sums of the partial counts, the multipliers) is gathered for the
                // whole block first, so that the chain finds its operands in registers
#pragma unroll 1
                for (uint32_t j0 = 0; j0 < kRound; j0 += kCoderBlock) {
                    uint32_t lo[kCoderBlock], hi[kCoderBlock], m[kCoderBlock];
#pragma unroll
                    for (uint32_t j = 0; j < kCoderBlock; ++j) {
                        const uint32_t pd = sm.ring_d[b][j0 + j][lane];
                        lo[j] = sm.ring_a[b][j0 + j][lane] + sm.ring_b[b][j0 + j][lane] + (pd & 0xFFFFu);
                        hi[j] = lo[j] + (pd >> 16);
                        m[j] = __shfl_sync(kFull, m_l, j0 + j);
                    }
#pragma unroll
                    for (uint32_t j = 0; j < kCoderBlock; ++j) {
                        uint32_t L1, S1;
                        narrow_lazy(L, R1, sx, lo[j], hi[j], m[j], sh, L1, S1);
                        sm.ring_c[b][j0 + j][lane] = pack_bounds(L1, S1);
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t m = __shfl_sync(kFull, m_l, j);
                    const uint32_t pd = sm.ring_d[b][j][lane];
                    const uint32_t lo = sm.ring_a[b][j][lane] + sm.ring_b[b][j][lane] + (pd & 0xFFFFu);
                    uint32_t c = 0;
                    if (i0 + j < len) {
                        uint32_t L1, S1;
                        narrow_lazy(L, R1, sx, lo, lo + (pd >> 16), m, sh, L1, S1);
                        c = pack_bounds(L1, S1);
                    }
                    sm.ring_c[b][j][lane] = c;
                }
            }
            if (r + 2u < rounds) {                                 // the buffer goes back to the three model warps
                bar_arrive<kPairCount>(kInEmpty + b);
                bar_arrive_raw<kPairCount>(kInEmpty + 2u + b);
                bar_arrive_raw<kPairCount>(kInEmpty + 4u + b);
            }
            if (r + 1u == rounds) sm.final_l[lane] = L;            // reaches BITS through the two hand-overs below
            bar_arrive<kPairCount>(kCFull + b);
        }
    } else if (role == 3u) {
        // ------------------------------------------------------------ FIELD (stateless)
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t b = r & 1u, i0 = r * kRound;
            const bool full = i0 + kRound <= min_len;
            bar_sync<kPairCount>(kCFull + b);
            if (r >= 2u) bar_sync<kPairCount>(kFEmpty + b);
            if (full) {
                // straight-line blocks of eight independent steps: the count-leading-zeros latencies overlap
#pragma unroll 1
                for (uint32_t j0 = 0; j0 < kRound; j0 += 8u) {
                    uint32_t c[8];
#pragma unroll
                    for (uint32_t j = 0; j < 8u; ++j) c[j] = sm.ring_c[b][j0 + j][lane];
#pragma unroll
                    for (uint32_t j = 0; j < 8u; ++j) {
                        uint32_t k, u;
                        shifts_of(c[j] & 0xFFFFu, c[j] >> 16, k, u);
                        const FieldDesc d = pack_field(k, u, c[j]);
                        sm.ring_f[b][j0 + j][lane] = make_uint2(d.w0, d.w1);
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t c = sm.ring_c[b][j][lane];
                    uint32_t k, u;
                    shifts_of(c & 0xFFFFu, c >> 16, k, u);
                    const FieldDesc d = pack_field(k, u, c);
                    sm.ring_f[b][j][lane] = (i0 + j < len) ? make_uint2(d.w0, d.w1) : make_uint2(0u, 0u);
                }
            }
            if (r + 2u < rounds) bar_arrive<kPairCount>(kCEmpty + b);
            bar_arrive<kPairCount>(kFFull + b);
        }
    } else if (role == 4u) {
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
            bar_sync<kPairCount>(kFFull + b);
            if (full) {
                // four steps per block.  The rare long-underflow path needs pend > 16 at a step
                // with k != 0; pend grows by at most the u's of the group, so one warp-uniform
                // vote on (pend + sum of u) covers the whole group and the common block has no
                // branch at all: its four ring loads issue together.  The rare block reads the
                // ring again rather than indexing the four registers (which would put them on
                // the stack for the common block too).
#pragma unroll 1
                for (uint32_t j0 = 0; j0 < kRound; j0 += 4u) {
                    const uint2 f0 = sm.ring_f[b][j0][lane], f1 = sm.ring_f[b][j0 + 1u][lane],
                                f2 = sm.ring_f[b][j0 + 2u][lane], f3 = sm.ring_f[b][j0 + 3u][lane];
                    const uint32_t usum = (f0.x >> 28) + (f1.x >> 28) + (f2.x >> 28) + (f3.x >> 28);
                    if (__any_sync(kFull, pend + usum > 16u)) {
#pragma unroll 1
                        for (uint32_t j = 0; j < 4u; ++j) {
                            const uint2 f = sm.ring_f[b][j0 + j][lane];
                            emit_packed_any(out, pend, FieldDesc{f.x, f.y});
                        }
                    } else {
                        emit_packed(out, pend, FieldDesc{f0.x, f0.y});
                        emit_packed(out, pend, FieldDesc{f1.x, f1.y});
                        emit_packed(out, pend, FieldDesc{f2.x, f2.y});
                        emit_packed(out, pend, FieldDesc{f3.x, f3.y});
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint2 f = sm.ring_f[b][j][lane];
                    if (field_valid(FieldDesc{f.x, f.y})) emit_packed_any(out, pend, FieldDesc{f.x, f.y});
                }
            }
            if (r + 2u < rounds) bar_arrive<kPairCount>(kFEmpty + b);
        }
        // final_l: written by CODER before its last "C full", which FIELD waited for before its last
        // "F full", which this warp waited for in the last round (release/acquire chain at CTA scope)
        uint32_t comp = 0;
        if (mine) {
            comp = finish_packet(out, sm.final_l[lane], pend, slot, len);
            if (sizes) sizes[my] = comp;
        }
        if (tg.world > 1u) {                                      // sharded encode: this rank's total for the other ranks
            const uint32_t sum = __reduce_add_sync(kFull, comp);
            if (lane == 0) shard_publish_total(sum, acc, tg);
        }
    }
}

ShardTarget shard_target(const ShardPlace &where, uint64_t call);     // encode.cu

cudaError_t launch_encode_slots_ws(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                   uint32_t *d_sizes, uint32_t packet, cudaStream_t st, const ShardPlace *where,
                                   uint64_t call, uint64_t *d_acc)
{
    const uint32_t packets = (uint32_t)((n + packet - 1) / packet);
    if (!packets) return launch_encode_slots(d_in, n, d_slots, slot_stride, d_sizes, packet, st, where, call, d_acc);
    const ShardTarget tg = where ? shard_target(*where, call) : ShardTarget{};
    // per device: the attribute belongs to the function of the current context
    cudaError_t e = cudaFuncSetAttribute(encode_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(WsShared));
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t ctas = (packets + 31u) / 32u;
    struct Tune { uint32_t warps, first, later; };
    static const Tune forced = [] {
        Tune t{0, 0, 0};
        if (const char *e = getenv("GPUAR_B200_WS_TUNE")) {
            unsigned w = 0, a = 0, b = 0;
            if (sscanf(e, "%u,%x,%x", &w, &a, &b) == 3 && (w == 6 || w == 8)) t = Tune{w, a, b};
        }
        return t;
    }();
    Tune t = ctas <= (uint32_t)sms ? Tune{8, kMapCoderAlone, kMapCoderAlone} : Tune{6, kMapSix, kMapSix};
    if (forced.warps) t = forced;
    encode_ws_kernel<<<ctas, 32u * t.warps, sizeof(WsShared), st>>>(d_in, n, d_slots, slot_stride, d_sizes, packets,
                                                                     packet, t.first, t.later, (uint32_t)sms, tg, d_acc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
