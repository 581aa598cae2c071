// encode.cu -- sm_100a encode path: one fused model+coder kernel with lane = packet,
// then a single-pass decoupled-look-back scan of the packet sizes fused with the
// compaction of the bitstreams into the .gip payload layout.
//
// Replaces garCompress / arCompress (reference src/gpuar_kernel.cu:487-531,
// 894-914) and the host-side compaction loop (src/gpu_compressor.cpp:136-169).
//
// Work decomposition: a warp owns 32 packets, lane = packet, and all lanes
// advance through their packets in lock step (position i of 32 different packets per
// step), so the running total T = 256 + i -- the divisor of both interval divisions -- is
// warp-uniform and becomes a multiply by a per-step reciprocal.  Per symbol:
//   MODEL  cum[s], count[s] and count[s]++ from the packet's 4-ary count tree in shared
//          memory (coder_math.h: tree_encode): root in registers, three independent 8-byte
//          loads/stores, byte permutes instead of branches.  Equals the reference's two
//          Fenwick prefix sums + update (gpuar_kernel.cu:215-238).
//   CODER  the interval recurrence (gpuar_kernel.cu:256-288) with the renormalisation loop
//          (:321-367) as a single normalisation on the plain window of the lower bound
//          (encode_math.h: narrow_plain).
//   BITS   the bits that left the window, plus the carry, added into a 64-bit accumulator
//          (the stream is the lower bound written out as one long number: no pending-underflow
//          counter), flushed as 32-bit words into the packet's slot (CarrySink).
// Only CODER is a loop-carried dependent chain.  The three stages are software pipelined
// inside each lane -- iteration i runs BITS(i-1), CODER(i), MODEL(i+1), which are mutually
// independent -- so the chain of one step overlaps the model and bit-packing work of its
// neighbours even when a scheduler holds a single warp (64 MiB = 256 warps on 592 schedulers).
#include <type_traits>

#include "common.cuh"
#include "encode_math.h"
#include "kernels.h"
#include "lookback.cuh"
#include "shard.cuh"

#ifndef GPUAR_ENC_Q_UNROLL
#define GPUAR_ENC_Q_UNROLL 1        // tuning knob: input words (4 symbols each) per pass of the steady-state loop
#endif

namespace gpuar {

constexpr int kEncWordsPerPass = GPUAR_ENC_Q_UNROLL;

// ------------------------------------------------------------------ encode
struct EncShared {
    uint64_t tree[kTreeStored][32];  // 21504 B; lane l owns column l (banks 2l, 2l+1); root in registers
};

__global__ void __launch_bounds__(32)
encode_kernel(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ slots, uint32_t slot_stride,
              uint32_t *__restrict__ sizes, uint32_t n_packets, uint32_t packet, ShardTarget tg, uint64_t *__restrict__ acc)
{
    __shared__ __align__(16) EncShared sm;
    const uint32_t lane = lane_id();
    const uint32_t my = blockIdx.x * 32u + lane;
    const bool mine = my < n_packets;

    uint64_t *const tree = &sm.tree[0][lane];
    uint64_t root;
    enc_tree_init(root, tree, 32u);                               // all counts 1 (:403-419)

    const size_t off = (size_t)my * packet;
    uint32_t len = 0;                                             // `packet` (8192) except for the last packet
    if (mine) len = (n - off < packet) ? (uint32_t)(n - off) : packet;
    const uint32_t max_len = __reduce_max_sync(kFull, len);
    const uint32_t min_len = __reduce_min_sync(kFull, mine ? len : packet);

    EncState st{0u, 65536u};                                      // plain window of the lower bound, range (encode_math.h)
    uint8_t *const slot = slots + (size_t)my * slot_stride;
    CarrySink out;
    out.start(reinterpret_cast<uint32_t *>(slot + kHdr), mine ? ((slot_stride - kHdr) >> 2) : 0u);

    // 16 input bytes per lane per half round, requested a whole round before their first use.  The read may
    // run up to 15 bytes past n inside the caller's 16-byte-rounded buffer (API contract).
    const uint4 *const in16 = reinterpret_cast<const uint4 *>(src + off);
    auto fetch = [&](uint32_t g) -> uint4 {
        return (g * 16u < len) ? __ldg(in16 + g) : make_uint4(0, 0, 0, 0);
    };
    uint4 bufA = fetch(0), bufB = fetch(1);                       // chunks 2r and 2r+1 at the top of round r

    // pipeline registers
    uint32_t lo_n = 0, cnt_n = 1;            // MODEL output for the next CODER step
    uint32_t p_inc = 0, p_t = 0;             // CODER output not yet in the bit sink: the bits that left the window (+ carry), their count
    bool pending = false;
    if (len) tree_encode(root, tree, 32u, bufA.x & 0xFFu, lo_n, cnt_n);          // MODEL(0)

    // BITS: append the step's bits; a carry that leaves the sink (through at least 16 pending one-bits: the
    // reference's pend >= 16) is added to the words already stored, under one warp-uniform vote
    auto bits = [&](auto fast_tag) {
        constexpr bool kFast = decltype(fast_tag)::value;
        const uint32_t stored = out.widx;
        const bool carry = out.push(p_inc, p_t);
        if (kFast ? __any_sync(kFull, carry) : carry) {
            if (carry) out.carry_into_stored(stored);
        }
    };

    // iteration i: BITS(i-1), CODER(i), MODEL(i+1).  kFast: every lane has positions i and
    // i+1 and a pending step, so nothing is predicated per lane.
    auto iter = [&](auto fast_tag, uint32_t i, uint32_t s_next, uint32_t m, uint32_t sh) {
        constexpr bool kFast = decltype(fast_tag)::value;
        if (kFast) {
            bits(fast_tag);
            narrow_plain(st, lo_n, lo_n + cnt_n, m, sh, p_inc, p_t);
            tree_encode(root, tree, 32u, s_next, lo_n, cnt_n);
        } else {
            if (pending) bits(fast_tag);
            pending = i < len;
            if (pending) narrow_plain(st, lo_n, lo_n + cnt_n, m, sh, p_inc, p_t);
            if (i + 1u < len) tree_encode(root, tree, 32u, s_next, lo_n, cnt_n);
        }
    };

    // 16 symbols held in `c`; `nx` = the word that follows them (first word of the next chunk)
    auto half = [&](const uint4 &c, uint32_t nx, bool fast, uint32_t i0, uint32_t h, uint32_t m_l, uint32_t sh) {
#pragma unroll kEncWordsPerPass
        for (uint32_t q = 0; q < 4u; ++q) {
            // 4 symbols per inner iteration keeps the loop body inside the instruction cache
            const uint32_t word = q == 0u ? c.x : q == 1u ? c.y : q == 2u ? c.z : c.w;
            const uint32_t next = q == 0u ? c.y : q == 1u ? c.z : q == 2u ? c.w : nx;
            const uint32_t ahead_syms = (word >> 8) | (next << 24);           // symbols i+1 .. i+4
            const uint32_t j0 = 16u * h + 4u * q;
            if (fast) {
#pragma unroll
                for (uint32_t j = 0; j < 4u; ++j)
                    iter(std::true_type{}, i0 + j0 + j, (ahead_syms >> (8u * j)) & 0xFFu,
                         __shfl_sync(kFull, m_l, j0 + j), sh);
            } else {
                // first round, last round, ragged tail: per-lane predicates
#pragma unroll 1
                for (uint32_t j = 0; j < 4u; ++j)
                    iter(std::false_type{}, i0 + j0 + j, (ahead_syms >> (8u * j)) & 0xFFu,
                         __shfl_sync(kFull, m_l, j0 + j), sh);
            }
        }
    };

    const uint32_t rounds = (max_len + 31u) >> 5;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i0 = r * 32u;
        uint32_t sh;
        const uint32_t m_l = magic_for(256u + i0 + lane, sh);     // lane j holds the multiplier of step j
        sh = shift_for(256u + i0);                                // the shift is uniform over the round
        // fast rounds: not the first (nothing pending yet at i = 0), and position i0+32 exists in every lane
        const bool fast = r != 0u && i0 + 33u <= min_len;
        // The next chunk of a buffer is requested at the START of the half round that still reads the buffer, into
        // other registers, and moved over at its end: every register read then waits for a load issued 16 steps
        // earlier.  (Requested at the end of the half round, the load shared its scoreboard with the other buffer's,
        // whose first read came right behind it: ncu showed 9 % of the warp time waiting there.)
        const uint4 nextA = fetch(2u * r + 2u);
        half(bufA, bufB.x, fast, i0, 0u, m_l, sh);
        bufA = nextA;
        const uint4 nextB = fetch(2u * r + 3u);
        half(bufB, bufA.x, fast, i0, 1u, m_l, sh);
        bufB = nextB;
    }
    if (pending) bits(std::false_type{});                         // BITS of the last step

    uint32_t comp = 0;
    if (mine) {
        comp = finish_packet_plain(out, st.Lp, slot, len);
        if (sizes) sizes[my] = comp;
    }
    if (tg.world > 1u) {                                          // sharded encode: this rank's total for the other ranks
        const uint32_t sum = __reduce_add_sync(kFull, comp);
        if (lane == 0) shard_publish_total(sum, acc, tg);
    }
}

// ------------------------------------------------- scan + compaction, one pass
// Tile = `tile` packets (the work unit of one CTA, 4..128, chosen by the launcher).  Decoupled
// look-back (Merrill & Garland) over the per-tile byte totals: descriptor = flag(2 bits) |
// value(62 bits) in one 64-bit word.
constexpr uint32_t kMaxTilePackets = 128;
constexpr uint32_t kCompactThreads = 256;

// Copy `len` bytes from src to dst (any alignment each) with one warp: 16-byte stores on the destination's
// alignment, the source read as aligned 16-byte words and shifted into place.  The source is read from
// the 16-byte boundary at or below src up to 32 bytes past src + len: inside the slots buffer (slots are
// 16-byte aligned and 8704 apart, a packet is at most 8281 bytes) -- also for the second piece of a
// packet that is split at a segment boundary, whose source starts anywhere inside the slot (a byte-wise
// copy of such a piece by a single warp took 70 us and held up the whole kernel, profiles/r2_shard_probe_trace_n8.txt).
__device__ __forceinline__ void warp_copy_unaligned(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src,
                                                    uint32_t len, uint32_t lane)
{
    // head: bring dst to 16-byte alignment
    const uint32_t head = min(len, (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u));
    if (lane < head) dst[lane] = src[lane];
    const uint32_t body = (len - head) >> 4;                       // whole 16-byte dst chunks
    const uint8_t *s = src + head;
    uint8_t *d = dst + head;
    const uint32_t mis = (uint32_t)((uintptr_t)s & 15u);           // warp-uniform
    const uint4 *sa = reinterpret_cast<const uint4 *>(s - mis);
    const uint32_t wsh = mis >> 2, bsh = (mis & 3u) * 8u;
    auto realign = [&](const uint4 &v0, const uint4 &v1) {
        uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t q[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) {                              // warp-uniform word shift
            q[t] = wsh == 0 ? w[t] : wsh == 1 ? w[t + 1] : wsh == 2 ? w[t + 2] : w[t + 3];
        }
        uint4 o;
        o.x = __funnelshift_r(q[0], q[1], bsh);
        o.y = __funnelshift_r(q[1], q[2], bsh);
        o.z = __funnelshift_r(q[2], q[3], bsh);
        o.w = __funnelshift_r(q[3], q[4], bsh);
        return o;
    };
    // two 512-byte rows per iteration: all the loads of the iteration are in flight together
    uint32_t c = lane;
    for (; c + 32u < body; c += 64u) {
        uint4 v0[2], v1[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            v0[r] = sa[c + 32u * r];
            v1[r] = mis ? sa[c + 32u * r + 1u] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) reinterpret_cast<uint4 *>(d)[c + 32u * r] = realign(v0[r], v1[r]);
    }
    for (; c < body; c += 32u) {
        const uint4 v0 = sa[c];
        const uint4 v1 = mis ? sa[c + 1] : make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4 *>(d)[c] = realign(v0, v1);
    }
    const uint32_t done = head + (body << 4);
    if (lane < len - done) dst[done + lane] = src[done + lane];
}

// ------------------------------------------------- multi-GPU: where a rank's stream lands
// The ranks' streams are concatenated in rank (= packet) order: rank r's stream starts at the
// exclusive scan of the W per-rank totals.  The concatenated stream lives in n_segments equal
// SEGMENTS of S = ceil(total / n_segments) bytes (rounded up to 256), segment g on GPU g: global
// offset o is byte o % S of segment o / S.  With balanced shards almost every byte stays on its
// own GPU and only the spill-over crosses NVLink; with skewed shards every GPU still receives
// exactly S bytes, so no GPU's ingress is the bottleneck.  One segment of the whole capacity is
// the gather-onto-one-GPU layout (ingress-limited there).  The segment pointers are local or
// peer-mapped (cudaIpcOpenMemHandle / cudaDeviceEnablePeerAccess); stores are 16 bytes wide.
//
// The only exchange between the ranks is the W totals.  They travel through MAILBOXES: every
// rank owns a small block of device memory that all ranks can write (peer-mapped); after its
// encode kernel a rank reduces its packet sizes and stores `tag | total` into word
// [parity][rank] of EVERY rank's mailbox (shard_total_kernel), and the compaction kernel of a
// rank reads the W words of its OWN mailbox, spinning until they carry the tag of this call.
// No host round trip, no NCCL, and the compaction writes each packet straight to its final
// place: there is no second pass over the payload.
//   tag    = call counter modulo 2^20 - 1, plus 1 (0 = never written), in bits 44..63
//   parity = call counter & 1: a rank can be at most one call ahead of the slowest one (it needs
//            that rank's total of call e+1, which is stream-ordered after its compaction of call
//            e), so two buffers suffice.
// Called by one full warp: the W totals of this call from the rank's own mailbox.  Returns false
// on a timeout.  base = bytes of the ranks before this one, all = bytes of all ranks.
__device__ __forceinline__ bool shard_totals(const ShardTarget &tg, uint32_t lane, uint64_t &base, uint64_t &all,
                                             uint64_t &mine)
{
    uint64_t v = 0;
    bool ok = true;
    if (lane < tg.world) {
        const uint64_t *p = tg.mailbox[tg.rank] + kMailTotals + tg.parity * kMaxRanks + lane;
        const uint64_t t0 = global_ns();
        for (;;) {
            v = ld_sys(p);
            if ((uint32_t)(v >> kTagShift) == tg.tag) break;
            if (global_ns() - t0 > kMailTimeoutNs) { ok = false; break; }
            __nanosleep(64);
        }
        v &= kValueMask;
    }
    ok = __all_sync(kFull, ok);
    uint64_t b = lane < tg.rank ? v : 0ull, a = v;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        b += __shfl_xor_sync(kFull, b, o);
        a += __shfl_xor_sync(kFull, a, o);
    }
    base = b;
    all = a;
    mine = __shfl_sync(kFull, v, (int)tg.rank);
    return ok;
}

__device__ __forceinline__ uint64_t shard_segment_bytes(const ShardTarget &tg, uint64_t all)
{
    if (tg.n_segments <= 1) return tg.seg_cap;
    const uint64_t seg = ((all + tg.n_segments - 1) / tg.n_segments + 255) & ~(uint64_t)255;
    return seg < kMinSegment ? kMinSegment : seg;            // a packet never spans more than two segments
}

// kGuard: the destination holds `cap` bytes and the sizes come from an untrusted stream (the
// rawLen fields of a foreign .gip, decode side): a packet that would end past `cap` is not
// copied; *total_out still reports the full sum, so the caller sees that it did not fit.
// kSharded: the destination is the concatenated stream of all ranks (ShardTarget); `total_out`
// = u64[5]: bytes of all ranks, segment size, this rank's base offset, this rank's bytes, status
// (0 ok, 1 a segment is too small, 2 a peer's total never arrived).
#ifdef GPUAR_SHARD_TRACE
__device__ uint64_t g_shard_trace[8192 * 4];      // per tile: start, after look-back, after the totals, end (globaltimer)
#define TRACE(slot) do { if (kSharded && threadIdx.x == 0 && tile < 8192u) g_shard_trace[tile * 4u + (slot)] = global_ns(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

template <bool kGuard, bool kSharded>
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const uint8_t *__restrict__ slots, uint32_t slot_stride, const uint32_t *__restrict__ sizes,
               uint32_t n_packets, uint32_t tile_packets, uint8_t *__restrict__ payload, uint64_t *__restrict__ desc,
               uint32_t *__restrict__ ticket, uint64_t *__restrict__ total_out, uint64_t cap, ShardTarget tg)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_off[kMaxTilePackets + 1];
    __shared__ uint64_t s_base, s_seg;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);          // tiles start in ticket order
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t first = tile * tile_packets;
    const uint32_t count = first < n_packets ? min(tile_packets, n_packets - first) : 0u;
    TRACE(0);

    if (warp == 0) {
        // exclusive scan of this tile's sizes (four per lane)
        uint32_t a[4], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) {
            a[i] = (4u * lane + i < count) ? sizes[first + 4u * lane + i] : 0u;
            sum += a[i];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        uint32_t ex = inc - sum;
#pragma unroll
        for (uint32_t i = 0; i < 4u; ++i) {
            s_off[4u * lane + i] = ex;
            ex += a[i];
        }
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        if (lane == 0) s_off[kMaxTilePackets] = total;              // entry 4*lane+i == count already holds it when count < 128

        // decoupled look-back for the bytes that precede this tile
        uint64_t base = lookback_exclusive(desc, tile, total, lane);
        TRACE(1);
        uint64_t seg = 0;
        if (kSharded) {
            uint64_t rank_base, all, mine;
            const uint64_t t_wait = global_ns();
            const bool ok = shard_totals(tg, lane, rank_base, all, mine);
            if (lane == 0 && tile == 0) total_out[5] = global_ns() - t_wait;   // diagnostics: ns the first tile waited for the totals
            seg = shard_segment_bytes(tg, all);
            const uint64_t status = !ok ? 2ull : (seg > tg.seg_cap || (tg.n_segments <= 1 && all > seg)) ? 1ull : 0ull;
            if (status) seg = 0;                                    // nothing is copied
            base += rank_base;
            if (lane == 0 && first + count >= n_packets) {
                total_out[0] = all;
                total_out[1] = status ? 0 : seg;
                total_out[2] = rank_base;
                total_out[3] = mine;
                total_out[4] = status;
            }
        } else if (lane == 0 && first + count == n_packets) {
            *total_out = base + total;
        }
        if (lane == 0) {
            s_base = base;
            s_seg = seg;
        }
        TRACE(2);
    }
    __syncthreads();

    const uint64_t base = s_base;
    if (kSharded) {
        const uint64_t seg = s_seg;
        if (seg == 0) return;
        for (uint32_t q = warp; q < count; q += kCompactThreads / 32u) {
            const uint64_t o = base + s_off[q];
            const uint32_t len = s_off[q + 1u] - s_off[q];
            const uint64_t g = o / seg, in = o - g * seg;
            const uint8_t *src = slots + (size_t)(first + q) * slot_stride;
            const uint32_t head = (uint32_t)min((uint64_t)len, seg - in);
            warp_copy_unaligned(tg.segment[g] + in, src, head, lane);
            uint32_t done = head;                                  // the rest straddles a boundary (or, with
            for (uint64_t h = g + 1u; done < len; ++h) {           // gather-sized segments, never gets here)
                const uint32_t part = (uint32_t)min((uint64_t)(len - done), seg);
                warp_copy_unaligned(tg.segment[h], src + done, part, lane);
                done += part;
            }
        }
#ifdef GPUAR_SHARD_TRACE
        __syncthreads();
        TRACE(3);
#endif
    } else {
        for (uint32_t q = warp; q < count; q += kCompactThreads / 32u) {
            if (kGuard && base + s_off[q + 1u] > cap) continue;
            warp_copy_unaligned(payload + base + s_off[q], slots + (size_t)(first + q) * slot_stride,
                                s_off[q + 1u] - s_off[q], lane);
        }
    }
}

// ------------------------------------------------------------------ launchers
const void *probe_kernel() { return reinterpret_cast<const void *>(&encode_kernel); }

ShardTarget shard_target(const ShardPlace &where, uint64_t call)
{
    ShardTarget tg{};
    for (uint32_t g = 0; g < where.n_segments; ++g) tg.segment[g] = where.segment[g];
    for (uint32_t r = 0; r < where.world; ++r) tg.mailbox[r] = where.mailbox[r];
    tg.seg_cap = where.seg_cap;
    tg.rank = where.rank;
    tg.world = where.world;
    tg.n_segments = where.n_segments;
    tg.tag = (uint32_t)(call % 0xFFFFFull) + 1u;
    tg.parity = (uint32_t)(call & 1u);
    return tg;
}

// one thread: a rank without packets still reports its total (zero)
__global__ void shard_publish_empty_kernel(uint64_t *acc, ShardTarget tg) { shard_publish_total(0, acc, tg); }

cudaError_t launch_encode_slots(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                uint32_t *d_sizes, uint32_t packet, cudaStream_t st, const ShardPlace *where,
                                uint64_t call, uint64_t *d_acc)
{
    const uint32_t packets = (uint32_t)((n + packet - 1) / packet);
    const ShardTarget tg = where ? shard_target(*where, call) : ShardTarget{};
    if (!packets) {
        if (where && where->world > 1u) {
            shard_publish_empty_kernel<<<1, 1, 0, st>>>(d_acc, tg);
            count_launch();
        }
        return cudaGetLastError();
    }
    encode_kernel<<<(packets + 31u) / 32u, 32, 0, st>>>(d_in, n, d_slots, slot_stride, d_sizes, packets, packet, tg, d_acc);
    count_launch();
    return cudaGetLastError();
}

// Work unit of the compaction.  0 = automatic: 64 packets (512 KiB of input) when that still
// gives every SM several CTAs, fewer for small inputs (a 64 MiB input is 8192 packets: 128
// CTAs of 64 packets would leave SMs idle and the copy latency-bound).
static uint32_t g_compact_tile = 0;
bool set_compact_tile(uint32_t packets_per_tile)
{
    if (packets_per_tile != 0 && (packets_per_tile < 4 || packets_per_tile > kMaxTilePackets ||
                                  (packets_per_tile & (packets_per_tile - 1)) != 0))
        return false;
    g_compact_tile = packets_per_tile;
    return true;
}
static uint32_t compact_tile_for(size_t packets)
{
    if (g_compact_tile) return g_compact_tile;
    uint32_t tile = 64;
    while (tile > 8 && packets / tile < 148u * 8u) tile >>= 1;
    return tile;
}

size_t compact_desc_bytes(size_t packets)
{
    const size_t tiles = (packets + 3) / 4;                         // the smallest tile there is
    return (tiles + 5) * sizeof(uint64_t);                          // descriptors + ticket word + totals accumulator
}

cudaError_t launch_compact(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                           uint32_t packets, uint8_t *d_payload, uint64_t *d_desc, uint64_t *d_total,
                           cudaStream_t st, uint64_t cap)
{
    const uint32_t tile = compact_tile_for(packets);
    const uint32_t tiles = (packets + tile - 1) / tile;
    cudaError_t e = cudaMemsetAsync(d_desc, 0, ((size_t)tiles + 2) * sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    if (!packets) return cudaMemsetAsync(d_total, 0, sizeof(uint64_t), st);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(d_desc + tiles);
    const ShardTarget none{};
    if (cap == kNoCap)
        compact_kernel<false, false><<<tiles, kCompactThreads, 0, st>>>(d_slots, slot_stride, d_sizes, packets, tile,
                                                                        d_payload, d_desc, ticket, d_total, cap, none);
    else
        compact_kernel<true, false><<<tiles, kCompactThreads, 0, st>>>(d_slots, slot_stride, d_sizes, packets, tile,
                                                                       d_payload, d_desc, ticket, d_total, cap, none);
    count_launch();
    return cudaGetLastError();
}

// The ranks' totals were published by the encode kernel (shard_publish_total); d_desc must have been
// zeroed by shard_desc_reset BEFORE that kernel (its accumulator words sit behind the descriptors).
uint64_t *shard_acc(uint64_t *d_desc, uint32_t packets)
{
    const uint32_t tile = compact_tile_for(packets);
    const uint32_t tiles = packets ? (packets + tile - 1) / tile : 1u;
    return d_desc + tiles + 1;
}
cudaError_t shard_desc_reset(uint64_t *d_desc, uint32_t packets, cudaStream_t st)
{
    const uint32_t tile = compact_tile_for(packets);
    const uint32_t tiles = packets ? (packets + tile - 1) / tile : 1u;
    return cudaMemsetAsync(d_desc, 0, ((size_t)tiles + 4) * sizeof(uint64_t), st);
}

cudaError_t launch_compact_sharded(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                                   uint32_t packets, uint64_t *d_desc, uint64_t *d_layout, const ShardPlace &where,
                                   uint64_t call, cudaStream_t st)
{
    if (where.world < 1 || where.world > kMaxRanks || where.rank >= where.world || where.n_segments < 1 ||
        where.n_segments > kMaxRanks)
        return cudaErrorInvalidValue;
    const ShardTarget tg = shard_target(where, call);
    const uint32_t tile = compact_tile_for(packets);
    const uint32_t tiles = packets ? (packets + tile - 1) / tile : 1u;   // a rank without packets still takes part
    uint32_t *ticket = reinterpret_cast<uint32_t *>(d_desc + tiles);
    compact_kernel<false, true><<<tiles, kCompactThreads, 0, st>>>(d_slots, slot_stride, d_sizes, packets, tile, nullptr,
                                                                   d_desc, ticket, d_layout, kNoCap, tg);
    count_launch();
    return cudaGetLastError();
}

#ifdef GPUAR_SHARD_TRACE
extern "C" int gpuar_b200_debug_shard_trace(uint64_t *host, size_t words)   // tools/shard_probe.py, tracing builds only
{
    return (int)cudaMemcpyFromSymbol(host, g_shard_trace, words * sizeof(uint64_t));
}
#endif

}  // namespace gpuar
