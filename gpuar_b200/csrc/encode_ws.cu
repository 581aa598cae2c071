// encode_ws.cu -- warp-specialised variant of the encode kernel for inputs that cannot fill
// the machine (a 64 MiB input is 8192 packets = 256 lane=packet warps for 592 warp schedulers).
//
// Same arithmetic as encode_kernel (coder_math.h, encode_math.h), same output, but the stages of a symbol
// step run in five different warps of a CTA that owns 32 packets:
//     MODEL-A  tree levels 0-1               -> ring A (partial cum[s])
//     MODEL-B  tree level 2                  -> ring B (partial cum[s])
//     MODEL-D  tree level 3 (the leaves)     -> ring D (partial cum[s] | count[s] << 16)
//     CODER    narrow_plain_lazy (the chain) -> ring C (the bits that left the window + carry | their count << 20)
//     BITS     carry-propagating bit sink    -> the packet's slot
// The tree levels are independent of each other given the symbol, so the model splits by
// level.  Only CODER carries the serial dependence of arithmetic coding, and it carries the
// minimum: the interval recurrence with its single normalisation on the plain window of the lower
// bound (encode_math.h).  The bit stream is the lower bound written out as one long number, so BITS
// only shifts the step's bits into its accumulator and adds the carry -- there is no pending-underflow
// counter and no warp that works out how the shift splits into matching and underflow shifts (round 1
// had a sixth, stateless FIELD warp for that).  CODER is the highest warp of the CTA because the SM
// sub-partition arbiter favours the highest warp slot.
// Each warp keeps lane = packet, so per-packet state never crosses lanes; the rings are
// double buffered per round of 32 positions and handed over with named barriers
// (bar.arrive on the producer side, bar.sync on the consumer side).  The three model rings
// share one full barrier per buffer (128 threads: three producers + CODER).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "encode_math.h"
#include "kernels.h"
#include "shard.cuh"

#ifndef GPUAR_WS_CODER_BLOCK
#define GPUAR_WS_CODER_BLOCK 8      // tuning knob: steps per straight-line block of the CODER warp (2 / 4 / 8 / 16: 0.589 / 0.476 / 0.459 / 0.462 ms at 32 MiB)
#endif

namespace gpuar {

constexpr uint32_t kRound = 32;
constexpr uint32_t kCoderBlock = GPUAR_WS_CODER_BLOCK;
// Role placement.  The CTA is launched with five or eight warps and a role map says what each one
// does: nibble w of the map = role of warp w (0 MODEL-A, 1 MODEL-B, 2 MODEL-D, 4 BITS, 5 CODER), 0xF = none
// (the warp exits at once).  Which roles share a scheduler (SM sub-partition; warp index modulo 4 for a CTA
// alone on its SM) matters (profiles/r2_encode_v2.md, 32 / 64 / 128 / 192 MiB in ms):
//     eight warps, A+B | BITS | D | CODER      0.476  0.633  1.168  1.565    <- one CTA per SM (up to 37 MiB, and every
//     eight warps, A | B+BITS | D | CODER      0.508  0.574  1.146  1.553       chunk of the host-buffer pipeline)
//     five warps,  B+CODER | A | D | BITS      0.528  0.651  1.251  1.495    <- up to four CTAs per SM / above
// (how the hardware spreads a second CTA's warps over the sub-partitions is not the simple modulo rule).
// GPUAR_B200_WS_TUNE="warps,map_first,map_later" (hex maps) overrides the choice for experiments; the CTAs
// of the first wave (one per SM) take map_first, the others map_later.
constexpr uint32_t kWsMaxThreads = 256;
constexpr uint32_t kMapCompact = 0xFFF54201u;    // five warps: B+CODER | A | D | BITS
constexpr uint32_t kMapCoderAlone = 0x5F4FF210u; // eight warps: A | B+BITS | D | CODER
constexpr uint32_t kMapAllAlone = 0x5FF1F240u;   // eight warps: A+B | BITS | D | CODER
#ifndef GPUAR_WS_SWAP_BD
#define GPUAR_WS_SWAP_BD 0          // tuning knob: which of the roles 1 / 2 takes level 2 and which the leaves
#endif
constexpr uint32_t kRoleB = GPUAR_WS_SWAP_BD ? 2u : 1u, kRoleD = GPUAR_WS_SWAP_BD ? 1u : 2u;

struct WsShared {
    uint64_t tree[kTreeStored][32];      // 21504 B: nodes 0-3 MODEL-A, 4-19 MODEL-B, 20-83 MODEL-D
    uint32_t ring_a[2][kRound][32];      //  8192 B each (MODEL -> CODER)
    uint32_t ring_b[2][kRound][32];
    uint32_t ring_d[2][kRound][32];
    uint32_t ring_c[2][kRound][32];      // CODER -> BITS: one word per step (encode_math.h: narrow_plain_lazy)
    uint32_t stage[3][8][32];            // the model warps' input words of the current round
    uint32_t final_l[32];                // CODER -> BITS at the end of the packet
};

// Named barriers, id + buffer index (the kernel has no __syncthreads, so barrier 0 is free).  A "full"
// barrier is arrived at by the producer(s) and waited on by the one consumer; an "empty" barrier the other
// way round.  The three model rings share their full barrier (three producers + CODER = 128 threads) but are
// handed back one by one, so that every barrier has exactly one waiting warp (which is also what
// compute-sanitizer's synccheck expects: it reports warps that wait on one barrier from different
// instructions as divergent).
enum : uint32_t { kInFull = 0, kInEmpty = 2 /* + 2 * role */, kCFull = 8, kCEmpty = 10 };
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

__global__ void __launch_bounds__(kWsMaxThreads, 1)
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
        uint32_t Lp = 0, R1 = 65536u, sx = 0;                     // narrow_plain_lazy state: range = R1 >> sx
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
                    for (uint32_t j = 0; j < kCoderBlock; ++j)
                        sm.ring_c[b][j0 + j][lane] = narrow_plain_lazy(Lp, R1, sx, lo[j], hi[j], m[j], sh);
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < kRound; ++j) {
                    const uint32_t m = __shfl_sync(kFull, m_l, j);
                    const uint32_t pd = sm.ring_d[b][j][lane];
                    const uint32_t lo = sm.ring_a[b][j][lane] + sm.ring_b[b][j][lane] + (pd & 0xFFFFu);
                    uint32_t c = kStepNone;                        // no bits, no carry
                    if (i0 + j < len) c = narrow_plain_lazy(Lp, R1, sx, lo, lo + (pd >> 16), m, sh);
                    sm.ring_c[b][j][lane] = c;
                }
            }
            if (r + 2u < rounds) {                                 // the buffer goes back to the three model warps
                bar_arrive<kPairCount>(kInEmpty + b);
                bar_arrive_raw<kPairCount>(kInEmpty + 2u + b);
                bar_arrive_raw<kPairCount>(kInEmpty + 4u + b);
            }
            if (r + 1u == rounds) sm.final_l[lane] = Lp;           // reaches BITS through the hand-over below
            bar_arrive<kPairCount>(kCFull + b);
        }
    } else if (role == 4u) {
        // ------------------------------------------------------------ BITS
        uint8_t *const slot = slots + (size_t)my * slot_stride;
        CarrySink out;
        out.start(reinterpret_cast<uint32_t *>(slot + kHdr), mine ? ((slot_stride - kHdr) >> 2) : 0u);
        for (uint32_t r = 0; r < rounds; ++r) {
            const uint32_t b = r & 1u;
            bar_sync<kPairCount>(kCFull + b);
            // four steps per block: their ring loads issue together.  A step past the end of a lane's packet is a
            // kStepNone word (no bits, no carry), so the ragged rounds need no predicate.  A carry that leaves the sink
            // (through at least 16 pending one-bits) is added to the words already stored, under one vote.
#pragma unroll 1
            for (uint32_t j0 = 0; j0 < kRound; j0 += 4u) {
                uint32_t c[4];
#pragma unroll
                for (uint32_t j = 0; j < 4u; ++j) c[j] = sm.ring_c[b][j0 + j][lane];
#pragma unroll
                for (uint32_t j = 0; j < 4u; ++j) {
                    uint32_t inc, t;
                    step_unpack(c[j], inc, t);
                    const uint32_t stored = out.widx;
                    const bool carry = out.push(inc, t);
                    if (__any_sync(kFull, carry)) {
                        if (carry) out.carry_into_stored(stored);
                    }
                }
            }
            if (r + 2u < rounds) bar_arrive<kPairCount>(kCEmpty + b);
        }
        // final_l: written by CODER before its last "C full", which this warp waited for in the last round
        // (release/acquire chain at CTA scope)
        uint32_t comp = 0;
        if (mine) {
            comp = finish_packet_plain(out, sm.final_l[lane], slot, len);
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
            if (sscanf(e, "%u,%x,%x", &w, &a, &b) == 3 && w >= 5 && w <= 8) t = Tune{w, a, b};
        }
        return t;
    }();
    Tune t = ctas <= (uint32_t)sms        ? Tune{8, kMapAllAlone, kMapAllAlone}
             : ctas <= 4u * (uint32_t)sms ? Tune{8, kMapCoderAlone, kMapCoderAlone}
                                          : Tune{5, kMapCompact, kMapCompact};
    if (forced.warps) t = forced;
    encode_ws_kernel<<<ctas, 32u * t.warps, sizeof(WsShared), st>>>(d_in, n, d_slots, slot_stride, d_sizes, packets,
                                                                     packet, t.first, t.later, (uint32_t)sms, tg, d_acc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
