// encode.cu -- sm_100a encode path: model pass + coder pass (one kernel), then a
// single-pass decoupled-look-back scan of the packet sizes fused with the
// compaction of the bitstreams into the .gip payload layout.
//
// Replaces garCompress / arCompress (reference src/gpuar_kernel.cu:487-531,
// 894-914) and the host-side compaction loop (src/gpu_compressor.cpp:136-169).
//
// Work decomposition:
//   one warp owns 32 packets and alternates, 32 input positions at a time, between
//   (A) MODEL PASS, warp-cooperative, one packet at a time: lane j takes symbol
//       x[i0+j] and computes the cumulative-frequency pair the adaptive model
//       would hand the coder at that position,
//           lo = cum[s] = pre[s] + #{earlier symbols of this tile that are < s}
//           cnt = count[s] + #{earlier symbols of this tile that are == s}
//       from the running 256-bin histogram `cnt` and its exclusive scan `pre` in
//       shared memory (ballot bit-plane ranks inside the tile, packed warp-shuffle
//       scan to refresh `pre`).  This equals the reference's getRange() pair
//       (gpuar_kernel.cu:215-227,272,279) without the Fenwick tree, because the
//       model does not depend on the coder state.
//   (B) CODER PASS, lane = packet: the serial interval recurrence
//       (gpuar_kernel.cu:256-288, 321-367) in closed form, dividing by the
//       warp-uniform total with a multiply; bits go to a 64-bit accumulator that is
//       flushed as 32-bit words into the packet's slot.
#include "common.cuh"
#include "kernels.h"
#include "lookback.cuh"

namespace gpuar {

// ------------------------------------------------------------------ encode
struct EncShared {
    uint16_t cnt[32][256];   // running symbol counts per packet (all start at 1)
    uint16_t pre[32][256];   // exclusive scan of cnt
    uint32_t pair[32][33];   // (lo | cnt << 16) for the 32 positions of the round, padded
    uint32_t in[32][8];      // the round's 32 input bytes of each packet
};

__global__ void __launch_bounds__(32)
encode_kernel(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ slots, uint32_t slot_stride,
              uint32_t *__restrict__ sizes, uint32_t n_packets)
{
    __shared__ __align__(16) EncShared sm;
    const uint32_t lane = lane_id();
    const uint32_t p0 = blockIdx.x * 32u;
    const uint32_t ltmask = (1u << lane) - 1u;

    // model init: counts 1, prefix = symbol index (gpuar_kernel.cu:403-419)
    for (uint32_t p = 0; p < 32; ++p) {
        uint32_t *c = reinterpret_cast<uint32_t *>(sm.cnt[p]);
        uint32_t *q = reinterpret_cast<uint32_t *>(sm.pre[p]);
        for (uint32_t w = lane; w < 128; w += 32) {
            c[w] = 0x00010001u;
            q[w] = (2u * w) | ((2u * w + 1u) << 16);
        }
    }

    // this lane's packet (coder pass)
    const uint32_t my = p0 + lane;
    const bool mine = my < n_packets;
    uint32_t my_len = 0;
    if (mine) {
        const size_t off = (size_t)my * kPacket;
        my_len = (n - off < kPacket) ? (uint32_t)(n - off) : kPacket;
    }
    // lengths are 8192 except possibly for the last packet of the stream
    const uint32_t warp_packets = min(32u, n_packets - p0);
    const uint32_t max_len = __reduce_max_sync(kFull, my_len);

    uint32_t L = 0, V = 0, pend = 0;
    BitSink out;
    out.acc = 0;
    out.nb = 0;
    uint8_t *slot = slots + (size_t)my * slot_stride;
    out.wp = reinterpret_cast<uint32_t *>(slot + kHdr);
    out.end = reinterpret_cast<uint32_t *>(slot + (mine ? (slot_stride & ~3u) : 0u));

    // staging: per round the warp needs 32 B from each of its packets = 64 x 16 B;
    // lane l fetches chunks l and l+32 (packet = chunk>>1, half = chunk&1)
    auto fetch = [&](uint32_t round, uint32_t chunk) -> uint4 {
        const uint32_t p = chunk >> 1;
        const size_t a = (size_t)(p0 + p) * kPacket + round * 32u + (chunk & 1u) * 16u;
        uint4 v = make_uint4(0, 0, 0, 0);
        // the 16-byte read stays inside the caller's buffer rounded up to 16 (API contract)
        if (p < warp_packets && a < n) v = *reinterpret_cast<const uint4 *>(src + a);
        return v;
    };
    uint4 nx0 = fetch(0, lane), nx1 = fetch(0, lane + 32u);

    const uint32_t rounds = (max_len + 31u) >> 5;
    for (uint32_t r = 0; r < rounds; ++r) {
        __syncwarp();
        *reinterpret_cast<uint4 *>(&sm.in[lane >> 1][(lane & 1u) * 4u]) = nx0;
        *reinterpret_cast<uint4 *>(&sm.in[16u + (lane >> 1)][(lane & 1u) * 4u]) = nx1;
        if (r + 1 < rounds) {
            nx0 = fetch(r + 1, lane);
            nx1 = fetch(r + 1, lane + 32u);
        }
        __syncwarp();

        // ---------------- (A) model pass: one packet at a time, lane = position
        const uint32_t i0 = r * 32u;
        for (uint32_t p = 0; p < warp_packets; ++p) {
            const size_t off = (size_t)(p0 + p) * kPacket;
            const uint32_t plen = (n - off < kPacket) ? (uint32_t)(n - off) : kPacket;
            if (i0 >= plen) continue;                              // warp-uniform
            const uint32_t valid = min(32u, plen - i0);
            const bool act = lane < valid;
            const uint32_t amask = (valid == 32u) ? kFull : ((1u << valid) - 1u);

            const uint32_t s = reinterpret_cast<const uint8_t *>(sm.in[p])[lane];
            const uint32_t c_before = sm.cnt[p][s];
            const uint32_t p_before = sm.pre[p][s];

            // ranks inside the tile from 8 bit-plane ballots (MSB first):
            //   E  = lanes whose symbol equals mine on the planes seen so far
            //   LT = lanes whose symbol is smaller than mine
            uint32_t E = amask, LT = 0;
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                const bool bit = (s >> b) & 1u;
                const uint32_t B = __ballot_sync(kFull, bit && act);
                if (bit) {
                    LT |= E & ~B;
                    E &= B;
                } else {
                    E &= ~B;
                }
            }
            const uint32_t lo = p_before + __popc(LT & ltmask);
            const uint32_t c = c_before + __popc(E & ltmask);
            sm.pair[p][lane] = lo | (c << 16);

            // the last lane of each equal-symbol group adds the group to the histogram
            if (act && (E >> lane) == 1u) sm.cnt[p][s] = (uint16_t)(c_before + __popc(E));
            __syncwarp();

            // refresh pre = exclusive scan of cnt: 8 bins per lane, packed u16x2
            const uint4 c4 = *reinterpret_cast<const uint4 *>(&sm.cnt[p][8u * lane]);
            const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
            uint32_t ew[4];
            const uint32_t tot = prefix8_packed(cw, ew);
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, inc, d);
                if (lane >= (uint32_t)d) inc += t;
            }
            const uint32_t base = (inc - tot) * 0x10001u;
            const uint4 e = make_uint4(ew[0] + base, ew[1] + base, ew[2] + base, ew[3] + base);
            *reinterpret_cast<uint4 *>(&sm.pre[p][8u * lane]) = e;
        }
        __syncwarp();

        // ---------------- (B) coder pass: lane = packet, 32 positions
        uint32_t sh_l;
        const uint32_t m_l = magic_for(256u + i0 + lane, sh_l);    // lane j holds the divisor of step j
        const uint32_t steps = min(32u, max_len - i0);
        for (uint32_t j = 0; j < steps; ++j) {
            const uint32_t m = __shfl_sync(kFull, m_l, j);
            const uint32_t sh = __shfl_sync(kFull, sh_l, j);
            if (i0 + j < my_len) {
                const uint32_t pr = sm.pair[lane][j];
                const uint32_t lo = pr & 0xFFFFu;
                const uint32_t hi = lo + (pr >> 16);
                uint32_t k, u, U1;
                narrow_renorm(L, V, lo, hi, m, sh, k, u, U1);
                emit_symbol(out, pend, k, u, U1);
            }
        }
    }

    if (mine) {
        const uint32_t comp = finish_packet(out, L, pend, slot, my_len);
        if (sizes) sizes[my] = comp;
    }
}

// ------------------------------------------------- scan + compaction, one pass
// Tile = kTilePackets packets.  Decoupled look-back (Merrill & Garland) over the
// per-tile byte totals: descriptor = flag(2 bits) | value(62 bits) in one 64-bit word.
constexpr uint32_t kTilePackets = 64;
constexpr uint32_t kCompactThreads = 256;

// copy `len` bytes from src (16-byte aligned) to dst (any alignment) with one warp
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
    for (uint32_t c = lane; c < body; c += 32u) {
        const uint4 v0 = sa[c];
        uint4 v1 = make_uint4(0, 0, 0, 0);
        if (mis) v1 = sa[c + 1];
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
        reinterpret_cast<uint4 *>(d)[c] = o;
    }
    const uint32_t done = head + (body << 4);
    if (lane < len - done) dst[done + lane] = src[done + lane];
}

__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const uint8_t *__restrict__ slots, uint32_t slot_stride, const uint32_t *__restrict__ sizes,
               uint32_t n_packets, uint8_t *__restrict__ payload, uint64_t *__restrict__ desc,
               uint32_t *__restrict__ ticket, uint64_t *__restrict__ total_out)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_off[kTilePackets + 1];
    __shared__ uint64_t s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);          // tiles start in ticket order
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t first = tile * kTilePackets;
    const uint32_t count = min(kTilePackets, n_packets - first);

    if (warp == 0) {
        // exclusive scan of this tile's 64 sizes (two per lane)
        const uint32_t a = (2u * lane < count) ? sizes[first + 2u * lane] : 0u;
        const uint32_t b = (2u * lane + 1u < count) ? sizes[first + 2u * lane + 1u] : 0u;
        uint32_t inc = a + b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        const uint32_t ex = inc - (a + b);
        s_off[2u * lane] = ex;
        s_off[2u * lane + 1u] = ex + a;
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        if (lane == 0) s_off[kTilePackets] = total;

        // decoupled look-back for the bytes that precede this tile
        const uint64_t base = lookback_exclusive(desc, tile, total, lane);
        if (lane == 0) {
            s_base = base;
            if (first + count == n_packets) *total_out = base + total;
        }
    }
    __syncthreads();

    const uint64_t base = s_base;
    for (uint32_t q = warp; q < count; q += kCompactThreads / 32u) {
        const uint32_t len = s_off[q + 1u] - s_off[q];
        warp_copy_unaligned(payload + base + s_off[q], slots + (size_t)(first + q) * slot_stride, len, lane);
    }
}

// ------------------------------------------------------------------ launchers
const void *probe_kernel() { return reinterpret_cast<const void *>(&encode_kernel); }

cudaError_t launch_encode_slots(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                uint32_t *d_sizes, cudaStream_t st)
{
    const uint32_t packets = (uint32_t)((n + kPacket - 1) / kPacket);
    if (!packets) return cudaSuccess;
    encode_kernel<<<(packets + 31u) / 32u, 32, 0, st>>>(d_in, n, d_slots, slot_stride, d_sizes, packets);
    count_launch();
    return cudaGetLastError();
}

size_t compact_desc_bytes(size_t packets)
{
    const size_t tiles = (packets + kTilePackets - 1) / kTilePackets;
    return (tiles + 2) * sizeof(uint64_t);                          // descriptors + ticket word
}

cudaError_t launch_compact(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                           uint32_t packets, uint8_t *d_payload, uint64_t *d_desc, uint64_t *d_total,
                           cudaStream_t st)
{
    const uint32_t tiles = (packets + kTilePackets - 1) / kTilePackets;
    cudaError_t e = cudaMemsetAsync(d_desc, 0, compact_desc_bytes(packets), st);
    if (e != cudaSuccess) return e;
    if (!packets) return cudaMemsetAsync(d_total, 0, sizeof(uint64_t), st);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(d_desc + tiles);
    compact_kernel<<<tiles, kCompactThreads, 0, st>>>(d_slots, slot_stride, d_sizes, packets, d_payload,
                                                      d_desc, ticket, d_total);
    count_launch();
    return cudaGetLastError();
}

}  // namespace gpuar
