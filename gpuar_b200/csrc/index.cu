// index.cu -- packet-chain discovery on the device.
//
// The .gip payload stores no index: packet i+1 is found by adding compLen(i) to the
// offset of packet i (reference src/cpu_compressor.cpp:50-56; the reference GPU driver
// walks the chain on the host while reading the file, src/gpu_compressor.cpp:294-320).
// A serial walk costs one dependent memory access per packet, so here it is parallel:
//
//   1. mark   every byte offset o is tested for "could start a packet":
//             5 <= compLen <= packet+512, o + compLen <= C, and rawLen == packet (8192) -- or the packet
//             ends exactly at C (the last packet may be short).  Compressed data looks
//             random, so false candidates are ~C/500k; true starts always qualify.
//   2. emit   single-pass decoupled-look-back scan over the candidate bitmap writes the
//             sorted candidate list cand[0..N).
//   3. link   next[i] = position in cand[] of cand[i] + compLen(cand[i])   (galloping
//             search; TERMINAL if it equals C, DEAD if it is not a candidate).
//   4. lift   jump tables S_l = next^(32^l) for l = 1..top, each from the previous by a
//             32-hop walk.
//   5. finish one thread descends the tables from cand[0] = 0 to count the chain (P packets)
//             and to validate that it ends at C.
//   6. rank   thread r walks the tables along the base-32 digits of r: offsets[r].
//
// If the fast path cannot be used (too many candidates for the scratch, chain not
// reaching C through full packets) step 5 falls back to a serial walk, which also validates
// the stream and accepts short packets anywhere (flagged in result[3]).
//
// One stream over W GPUs (gpuar_b200_decode_sharded): the stream lies in W equal segments,
// segment g on GPU g, and a segment begins in the middle of a packet.  Every GPU runs steps 1-4
// over its own segment at the same time (offsets local to the segment; a packet that reaches into
// the next segment reads the head of it, which the GPU copied behind its own segment first) with
// "ends at or behind the end of the segment" as the terminal condition.  Only step 5 is a chain
// over the GPUs: GPU g learns from GPU g-1 where the first packet of its segment starts and how
// many packets precede it (two words written into its mailbox with peer stores), descends its
// jump tables from there, and hands the exit of its segment on to GPU g+1.
#include "common.cuh"
#include "kernels.h"
#include "lookback.cuh"
#include "shard.cuh"

namespace gpuar {

constexpr uint32_t kMarkThreads = 256;
constexpr uint32_t kEmitThreads = 256;
constexpr uint32_t kEmitWordsPerThread = 8;
constexpr uint32_t kEmitTileWords = kEmitThreads * kEmitWordsPerThread;   // 2048 words = 64 KiB of payload
constexpr uint32_t kMaxLevels = 7;                                        // 32^7 > 2^32 candidates

struct IndexPlan {
    size_t c;
    size_t words;        // bitmap words
    size_t tiles;        // emit tiles
    uint32_t cap;        // candidate capacity; sentinels live at cap (TERMINAL) and cap+1 (DEAD)
    uint32_t levels;     // jump tables S_0 .. S_{levels-1}
    // scratch layout (byte offsets)
    size_t off_bitmap, off_desc, off_cand, off_jump, off_ctl, total;
};

struct IndexCtl {        // device control block
    uint32_t n_cand;     // candidates found (may exceed cap: then the fast path is off)
    uint32_t serial;     // 1 = offsets were produced by the serial fallback
    uint32_t entry_pos;  // position in cand[] of the first packet of the chain (0 for a whole stream)
    uint32_t pad;
};

// The part of a stream one index run covers, in offsets local to `payload`: candidates are the
// offsets below `scan`; the stream itself ends at `end` >= scan (beyond what this run looks at when
// the segment is not the last one).  A whole stream: scan == end == c.
struct IndexRange { size_t scan, end; };

static IndexPlan make_plan(size_t c)
{
    IndexPlan p{};
    p.c = c;
    p.words = (c + 31) / 32;
    p.tiles = (p.words + kEmitTileWords - 1) / kEmitTileWords;
    // every-count-1 packets of zeros are 210 bytes: c/200 covers any stream this codec emits
    const size_t cap = c / 200 + 4096;
    p.cap = (uint32_t)(cap > 0xFFFFFFF0u ? 0xFFFFFFF0u : cap);
    p.levels = 1;
    for (uint64_t span = 32; span < (uint64_t)p.cap + 1 && p.levels < kMaxLevels; span *= 32) ++p.levels;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    p.off_bitmap = take((p.words + 8) * 4);
    p.off_desc = take((p.tiles + 2) * 8);
    p.off_cand = take(((size_t)p.cap + 2) * 8);
    p.off_jump = take((size_t)p.levels * ((size_t)p.cap + 2) * 4);
    p.off_ctl = take(sizeof(IndexCtl));
    p.total = o;
    return p;
}

size_t index_scratch_bytes(size_t c) { return make_plan(c).total; }

__device__ __forceinline__ uint32_t ld16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// ---- 1. mark: thread = 32 consecutive offsets = one bitmap word.
// Prefilter first, branch free: away from the end of the stream only rawLen == packet can start
// a packet, i.e. the two bytes at o+2 (one funnel shift + compare per offset, ~1 in 65536
// random offsets pass); the full test runs only on the survivors.  Within packet+512 bytes of
// the end every offset gets the full test (the last packet may be short).
__device__ __forceinline__ bool index_is_candidate(const uint8_t *__restrict__ payload, size_t o, size_t c, uint32_t packet)
{
    const uint32_t len = ld16(payload + o), raw = ld16(payload + o + 2);
    return len > kHdr && len <= packet + 512u && o + len <= c &&
           (raw == packet || (o + len == c && raw >= 1u && raw <= packet));
}

__global__ void __launch_bounds__(kMarkThreads)
index_mark_kernel(const uint8_t *__restrict__ payload, IndexRange rg, uint32_t *__restrict__ bitmap, size_t words,
                  uint32_t packet)
{
    const size_t w = (size_t)blockIdx.x * kMarkThreads + threadIdx.x;
    if (w >= words) return;
    const size_t o0 = w * 32;
    const size_t c = rg.end;
    uint32_t pre = 0xFFFFFFFFu;
    if (o0 + 32u + packet + 512u < c) {
        // bytes o0 .. o0+35 (payload is readable GPUAR_PAD_BYTES past c; base is 16-byte aligned)
        const uint4 *v = reinterpret_cast<const uint4 *>(payload + o0);
        const uint4 a = v[0], b = v[1];
        const uint32_t tail = *reinterpret_cast<const uint32_t *>(payload + o0 + 32);
        const uint32_t x[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, tail, 0u};   // x[9] only pads the funnel
        pre = 0;
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) {
            const uint32_t k = j + 2u;                              // rawLen lives at bytes o+2, o+3
            const uint32_t two = __funnelshift_r(x[k >> 2], x[(k >> 2) + 1u], (k & 3u) * 8u) & 0xFFFFu;
            pre |= (uint32_t)(two == packet) << j;
        }
    }
    uint32_t bits = 0;
    while (pre) {
        const uint32_t j = __ffs(pre) - 1u;
        pre &= pre - 1u;
        const size_t o = o0 + j;
        if (o < rg.scan && o + kHdr <= c && index_is_candidate(payload, o, c, packet)) bits |= 1u << j;
    }
    bitmap[w] = bits;
}

// ---- 2. emit: ordered compaction of the set bits into cand[]
__global__ void __launch_bounds__(kEmitThreads)
index_emit_kernel(const uint32_t *__restrict__ bitmap, size_t words, uint64_t *__restrict__ cand, uint32_t cap,
                  uint64_t *__restrict__ desc, uint32_t *__restrict__ ticket, IndexCtl *__restrict__ ctl,
                  size_t tiles)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[kEmitThreads / 32];
    __shared__ uint64_t s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;

    const size_t w0 = (size_t)tile * kEmitTileWords + (size_t)threadIdx.x * kEmitWordsPerThread;
    uint32_t bm[kEmitWordsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < kEmitWordsPerThread; ++k) {
        bm[k] = (w0 + k < words) ? bitmap[w0 + k] : 0u;
        mine += __popc(bm[k]);
    }
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < kEmitThreads / 32 ? s_warp[lane] : 0u;
        uint32_t wi = v;
#pragma unroll
        for (int d = 1; d < (int)(kEmitThreads / 32); d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        const uint32_t total = __shfl_sync(kFull, wi, kEmitThreads / 32 - 1);
        if (lane < kEmitThreads / 32) s_warp[lane] = wi - v;       // exclusive warp bases
        const uint64_t base = lookback_exclusive(desc, tile, total, lane);
        if (lane == 0) {
            s_base = base;
            if ((size_t)tile + 1 == tiles) {
                const uint64_t n = base + total;
                ctl->n_cand = n > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)n;
            }
        }
    }
    __syncthreads();
    uint64_t at = s_base + s_warp[warp] + (inc - mine);
#pragma unroll
    for (uint32_t k = 0; k < kEmitWordsPerThread; ++k) {
        uint32_t b = bm[k];
        while (b) {
            const uint32_t j = __ffs(b) - 1u;
            b &= b - 1u;
            if (at < cap) cand[at] = (uint64_t)(w0 + k) * 32u + j;
            ++at;
        }
    }
}

// ---- 3. link
__global__ void __launch_bounds__(256)
index_link_kernel(const uint8_t *__restrict__ payload, size_t scan, const uint64_t *__restrict__ cand, uint32_t cap,
                  const IndexCtl *__restrict__ ctl, uint32_t *__restrict__ jump, uint32_t levels)
{
    const uint32_t n = ctl->n_cand;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const uint32_t kTerminal = cap, kDead = cap + 1u;
    if (i < 2u) {                                                   // sentinels of every table: self loops
        for (uint32_t l = 0; l < levels; ++l) jump[(size_t)l * (cap + 2u) + cap + i] = cap + i;
    }
    if (n > cap || i >= n) return;
    const uint64_t o = cand[i];
    const uint64_t tgt = o + ld16(payload + o);
    uint32_t nx = kDead;
    if (tgt >= scan) {                                              // ends the stream, or leaves the segment
        nx = kTerminal;
    } else {
        // cand[] is sorted and tgt > o: gallop forward from i+1, then bisect
        uint32_t lo = i + 1u, step = 1u;
        uint32_t hi = lo;
        while (hi < n && cand[hi] < tgt) { lo = hi + 1u; hi += step; step <<= 1; }
        if (hi > n) hi = n;
        while (lo < hi) {                                           // first index with cand >= tgt in [lo, hi)
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (cand[mid] < tgt) lo = mid + 1u; else hi = mid;
        }
        if (lo < n && cand[lo] == tgt) nx = lo;
    }
    jump[i] = nx;
}

// ---- 4. lift: S_l[i] = S_{l-1} applied 32 times
__global__ void __launch_bounds__(256)
index_lift_kernel(const uint32_t *__restrict__ prev, uint32_t *__restrict__ next, uint32_t cap,
                  const IndexCtl *__restrict__ ctl)
{
    const uint32_t n = ctl->n_cand;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (n > cap || i >= n) return;
    uint32_t pos = i;
#pragma unroll 1
    for (int h = 0; h < 32 && pos < cap; ++h) pos = prev[pos];
    next[i] = pos;
}

// ---- 5. finish: chain length + validation, or the serial fallback
// ChainHop: how this run is chained to the GPUs before and after it (sharded decode); for a whole
// stream on one GPU `inbox` and `outbox` are null and the chain starts at offset 0.
struct ChainHop {
    const uint64_t *inbox;   // own mailbox words [4] of this call's parity, or null (first segment)
    uint64_t *outbox;        // the next rank's, or null (last rank)
    uint64_t seg_base;       // global offset of payload[0]
    uint64_t stream_bytes;   // length of the whole stream
    uint32_t tag;
    uint32_t last;           // 1: the stream must end inside this run (scan == end)
    uint32_t sharded;        // 1: result[3] / result[4] report what precedes this segment
};

__device__ __forceinline__ void hand_on_raw(const ChainHop &hop, uint64_t g_entry, uint64_t packets, uint64_t raw,
                                            uint64_t status)
{
    if (!hop.outbox) return;
    mail_post(hop.outbox + 0, hop.tag, g_entry);
    mail_post(hop.outbox + 1, hop.tag, packets);
    mail_post(hop.outbox + 2, hop.tag, raw);
    mail_post(hop.outbox + 3, hop.tag, status);
    __threadfence_system();
}

// result: [0] packets, [1] raw bytes, [2] status, [3] ragged flag (whole stream) / packets before this
// segment (sharded), [4] (sharded) raw bytes before this segment
__global__ void index_finish_kernel(const uint8_t *__restrict__ payload, IndexRange rg, const uint64_t *__restrict__ cand,
                                    uint32_t cap, IndexCtl *__restrict__ ctl, const uint32_t *__restrict__ jump,
                                    uint32_t levels, uint64_t *__restrict__ offsets, size_t max_packets,
                                    uint64_t *__restrict__ result, uint32_t packet, ChainHop hop)
{
    if (threadIdx.x || blockIdx.x) return;
    const bool sharded = hop.sharded != 0;
    const size_t c = rg.end;
    ctl->serial = 0;
    ctl->entry_pos = 0;
    // where the chain enters this run
    uint64_t entry = 0, before = 0, raw_before = 0, status_before = 0;
    if (hop.inbox) {
        uint64_t g_entry = 0;
        const bool ok = mail_wait(hop.inbox + 0, hop.tag, g_entry) && mail_wait(hop.inbox + 1, hop.tag, before) &&
                        mail_wait(hop.inbox + 2, hop.tag, raw_before) && mail_wait(hop.inbox + 3, hop.tag, status_before);
        if (!ok) status_before = 5;                                 // a rank before this one never reported
        if (!status_before && g_entry >= hop.stream_bytes) {        // the chain ended before this segment
            hand_on_raw(hop, g_entry, before, raw_before, g_entry == hop.stream_bytes ? 0 : 2);
            result[0] = 0; result[1] = 0; result[2] = g_entry == hop.stream_bytes ? 0 : (uint64_t)(int64_t)-2;
            result[3] = before; result[4] = raw_before;
            return;
        }
        entry = g_entry >= hop.seg_base ? g_entry - hop.seg_base : 0;
        if (g_entry < hop.seg_base && !status_before) status_before = 2;   // a packet cannot begin before its segment
    }
    auto hand_on = [&](uint64_t exit_local, uint64_t packets, uint64_t raw, uint64_t status) {
        hand_on_raw(hop, hop.seg_base + exit_local, before + packets, raw_before + raw, status);
    };
    auto report = [&](uint64_t packets, uint64_t raw, int64_t status, uint64_t ragged) {
        result[0] = packets;
        result[1] = raw;
        result[2] = (uint64_t)status;
        result[3] = sharded ? before : ragged;
        if (sharded) result[4] = raw_before;
    };
    if (status_before) {                                            // a segment before this one is broken
        hand_on(entry, 0, 0, status_before);
        report(0, 0, -2, 0);
        return;
    }
    if (c == 0 || entry >= rg.scan) {                               // nothing starts here
        const bool bad = hop.last && entry != c;                    // ... but the stream has to end here
        hand_on(entry, 0, 0, bad ? 2 : 0);
        report(0, 0, bad ? -2 : 0, 0);
        return;
    }
    const uint32_t n = ctl->n_cand;
    const size_t row = (size_t)cap + 2u;
    if (n && n <= cap) {
        // the entry has to be a candidate: first index with cand >= entry
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (cand[mid] < entry) lo = mid + 1u; else hi = mid;
        }
        if (lo < n && cand[lo] == entry) {
            uint32_t pos = lo;
            uint64_t hops = 0, weight = 1;
            for (uint32_t l = 1; l < levels; ++l) weight *= 32u;
            for (int l = (int)levels - 1; l >= 0; --l, weight /= 32u) {
                const uint32_t *S = jump + (size_t)l * row;
                for (;;) {
                    const uint32_t nx = S[pos];
                    if (nx >= cap) break;
                    pos = nx;
                    hops += weight;
                }
            }
            const uint64_t exit_at = cand[pos] + ld16(payload + cand[pos]);
            if (jump[pos] == cap && (!hop.last || exit_at == c)) {  // leaves the run through a well-formed packet
                const uint64_t packets = hops + 1u;
                const uint64_t last_raw = ld16(payload + cand[pos] + 2);   // short only if it ends the stream
                const uint64_t raw = (packets - 1u) * packet + last_raw;
                ctl->entry_pos = lo;
                hand_on(exit_at, packets, raw, 0);
                report(packets, raw, packets <= max_packets ? 0 : -1 /* GPUAR_E_ARG */, 0);
                return;
            }
        }
    }
    // serial fallback: walk and validate (cpu_compressor.cpp:47-78).  This is also the path of
    // streams with short packets before the last one (what the reference's CPU decoder accepts,
    // cpu_compressor.cpp:60-70, and its GPU decoder does not, gpuar_kernel.cu:924): they are
    // indexed here and flagged in result[3]; gpuar_b200_decode_packed writes them.
    ctl->serial = 1;
    uint64_t o = entry, k = 0, raw_total = 0, ragged = 0;
    int64_t status = 0;
    while (o < rg.scan) {
        if (c - o < kHdr) { status = -2; break; }
        const uint64_t len = ld16(payload + o), raw = ld16(payload + o + 2);
        if (len <= kHdr || len > c - o) { status = -2; break; }
        if (raw == 0 || raw > packet) { status = -4; break; }
        if (raw != packet && o + len != c) ragged = 1;
        if (k < max_packets) offsets[k] = o; else status = -1;
        ++k;
        raw_total += raw;
        o += len;
    }
    if (sharded && ragged && !status) status = -4;                  // the sharded decode writes packet p at p * packet
    if (hop.last && !status && o != c) status = -2;
    hand_on(o, k, raw_total, status ? 2 : 0);
    report(k, raw_total, status, ragged);
}

// ---- 6. rank
__global__ void __launch_bounds__(256)
index_rank_kernel(const uint64_t *__restrict__ cand, uint32_t cap, const IndexCtl *__restrict__ ctl,
                  const uint32_t *__restrict__ jump, uint32_t levels, uint64_t *__restrict__ offsets,
                  size_t max_packets, const uint64_t *__restrict__ result)
{
    if (ctl->serial || result[2] != 0) return;
    const uint64_t packets = result[0];
    const uint64_t r = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (r >= packets || r >= max_packets) return;
    const size_t row = (size_t)cap + 2u;
    uint32_t pos = ctl->entry_pos;
    for (int l = (int)levels - 1; l >= 0; --l) {
        const uint32_t *S = jump + (size_t)l * row;
        const uint32_t digit = (uint32_t)(r >> (5 * l)) & 31u;
        for (uint32_t h = 0; h < digit; ++h) pos = S[pos];
    }
    offsets[r] = cand[pos];
}

// ---- raw sizes of indexed packets (for the packed output layout of ragged streams)
__global__ void __launch_bounds__(256)
raw_sizes_kernel(const uint8_t *__restrict__ payload, size_t c, const uint64_t *__restrict__ offsets,
                 uint32_t n_packets, uint32_t packet, uint32_t *__restrict__ sizes)
{
    const uint32_t p = blockIdx.x * 256u + threadIdx.x;
    if (p >= n_packets) return;
    const uint64_t o = offsets[p];
    sizes[p] = o + kHdr <= c ? min(ld16(payload + o + 2), packet) : 0u;    // exactly what decode_kernel writes
}

cudaError_t launch_raw_sizes(const uint8_t *d_payload, size_t c, const uint64_t *d_offsets, uint32_t packets,
                             uint32_t packet, uint32_t *d_sizes, cudaStream_t st)
{
    if (!packets) return cudaSuccess;
    raw_sizes_kernel<<<(packets + 255u) / 256u, 256, 0, st>>>(d_payload, c, d_offsets, packets, packet, d_sizes);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t run_index(const uint8_t *d_payload, IndexRange rg, uint64_t *d_offsets, size_t max_packets,
                             uint64_t *d_result, void *d_scratch, size_t scratch_bytes, uint32_t packet, ChainHop hop,
                             cudaStream_t st)
{
    const IndexPlan p = make_plan(rg.scan);
    if (scratch_bytes < p.total) return cudaErrorInvalidValue;
    uint8_t *s = static_cast<uint8_t *>(d_scratch);
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(s + p.off_bitmap);
    uint64_t *desc = reinterpret_cast<uint64_t *>(s + p.off_desc);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(desc + p.tiles);
    uint64_t *cand = reinterpret_cast<uint64_t *>(s + p.off_cand);
    uint32_t *jump = reinterpret_cast<uint32_t *>(s + p.off_jump);
    IndexCtl *ctl = reinterpret_cast<IndexCtl *>(s + p.off_ctl);

    cudaError_t e = cudaMemsetAsync(desc, 0, (p.tiles + 2) * 8, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ctl, 0, sizeof(IndexCtl), st);
    if (e != cudaSuccess) return e;
    if (rg.scan) {
        index_mark_kernel<<<(unsigned)((p.words + kMarkThreads - 1) / kMarkThreads), kMarkThreads, 0, st>>>(
            d_payload, rg, bitmap, p.words, packet);
        index_emit_kernel<<<(unsigned)p.tiles, kEmitThreads, 0, st>>>(bitmap, p.words, cand, p.cap, desc, ticket,
                                                                      ctl, p.tiles);
        // candidates of a well-formed stream: one per packet plus ~c/500k false ones; the grid
        // covers the whole capacity and threads past n_cand exit
        const unsigned grid = (unsigned)(((size_t)p.cap + 255) / 256);
        index_link_kernel<<<grid, 256, 0, st>>>(d_payload, rg.scan, cand, p.cap, ctl, jump, p.levels);
        for (uint32_t l = 1; l < p.levels; ++l)
            index_lift_kernel<<<grid, 256, 0, st>>>(jump + (size_t)(l - 1) * ((size_t)p.cap + 2),
                                                    jump + (size_t)l * ((size_t)p.cap + 2), p.cap, ctl);
        count_launch(3 + (int)p.levels - 1);
    }
    index_finish_kernel<<<1, 32, 0, st>>>(d_payload, rg, cand, p.cap, ctl, jump, p.levels, d_offsets, max_packets,
                                          d_result, packet, hop);
    count_launch();
    if (rg.scan) {
        // at most one packet per 5 payload bytes; well-formed streams have far fewer
        const size_t upper = max_packets < rg.scan / 5 + 1 ? max_packets : rg.scan / 5 + 1;
        if (upper == 0) return cudaGetLastError();                 // no room for offsets: result[2] already says E_ARG
        index_rank_kernel<<<(unsigned)((upper + 255) / 256), 256, 0, st>>>(cand, p.cap, ctl, jump, p.levels,
                                                                           d_offsets, max_packets, d_result);
        count_launch();
    }
    return cudaGetLastError();
}

cudaError_t launch_index(const uint8_t *d_payload, size_t c, uint64_t *d_offsets, size_t max_packets,
                         uint64_t *d_result, void *d_scratch, size_t scratch_bytes, uint32_t packet, cudaStream_t st)
{
    ChainHop whole{};
    whole.last = 1;
    return run_index(d_payload, IndexRange{c, c}, d_offsets, max_packets, d_result, d_scratch, scratch_bytes, packet,
                     whole, st);
}

// copies the head of the next rank's segment behind this rank's own (peer loads, 16 bytes wide)
__global__ void __launch_bounds__(256)
shard_halo_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint32_t words16)
{
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < words16; i += gridDim.x * 256u) dst[i] = src[i];
}

cudaError_t launch_index_segment(const ShardPlace &where, uint64_t call, uint64_t stream_bytes, uint64_t seg_bytes,
                                 uint64_t *d_offsets, size_t max_packets, uint64_t *d_result, void *d_scratch,
                                 size_t scratch_bytes, cudaStream_t st)
{
    if (where.world < 1 || where.world > kMaxRanks || where.rank >= where.world || where.n_segments != where.world ||
        seg_bytes == 0 || (seg_bytes & 255u) || seg_bytes > where.seg_cap)
        return cudaErrorInvalidValue;
    const uint64_t base = seg_bytes * where.rank;
    const uint64_t here = base < stream_bytes ? (stream_bytes - base < seg_bytes ? stream_bytes - base : seg_bytes) : 0;
    uint8_t *seg = where.segment[where.rank];
    const bool has_next = where.rank + 1u < where.world && base + seg_bytes < stream_bytes;
    if (has_next) {
        shard_halo_kernel<<<4, 256, 0, st>>>(reinterpret_cast<uint4 *>(seg + seg_bytes),
                                             reinterpret_cast<const uint4 *>(where.segment[where.rank + 1u]),
                                             kShardHalo / 16u);
        count_launch();
    } else {
        cudaError_t e = cudaMemsetAsync(seg + here, 0, 64, st);    // the decoder reads a few bytes past the stream
        if (e != cudaSuccess) return e;
    }
    ChainHop hop{};
    const uint32_t parity = (uint32_t)(call & 1u);
    hop.inbox = where.rank ? where.mailbox[where.rank] + kMailChain + parity * 4u : nullptr;
    hop.outbox = where.rank + 1u < where.world ? where.mailbox[where.rank + 1u] + kMailChain + parity * 4u : nullptr;
    hop.seg_base = base;
    hop.stream_bytes = stream_bytes;
    hop.tag = (uint32_t)(call % 0xFFFFFull) + 1u;
    hop.last = base + seg_bytes >= stream_bytes ? 1u : 0u;
    hop.sharded = 1;
    // local coordinates: candidates below `here`, the stream ends at stream_bytes - base
    const IndexRange rg{(size_t)here, (size_t)(stream_bytes > base ? stream_bytes - base : 0)};
    return run_index(seg, rg, d_offsets, max_packets, d_result, d_scratch, scratch_bytes, kPacket, hop, st);
}

}  // namespace gpuar
