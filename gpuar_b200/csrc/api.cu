// api.cu -- the C ABI of libgpuar_b200.so (include/gpuar_b200.h): argument checking,
// scratch layout, the host-buffer pipelines and the reference-named shims.
#include "../../include/gpuar_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <atomic>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace gpuar {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline size_t packets_of(size_t n, size_t packet = kPacket) { return (n + packet - 1) / packet; }
static inline bool packet_ok(size_t packet) { return packet >= 16 && packet <= 16112 && packet % 16 == 0; }

// encode scratch: [slots: packets*8704 + 64][sizes: packets*4][descriptors][pad]
struct EncodePlan {
    size_t packets, off_slots, off_sizes, off_desc, total;
};
static EncodePlan encode_plan(size_t n, size_t packet = kPacket)
{
    EncodePlan p{};
    p.packets = packets_of(n, packet);
    size_t o = 0;
    p.off_slots = o;
    o += align_up(p.packets * (packet + 512) + 64, 256);
    p.off_sizes = o;
    o += align_up(p.packets * 4 + 4, 256);
    p.off_desc = o;
    o += align_up(compact_desc_bytes(p.packets), 256);
    p.total = o;
    return p;
}

static int ck(cudaError_t e) { return (int)e; }

// ---- tuning knobs (gpuar_b200_set_option)
static int g_encode_path = 0;                    // 0 auto, 1 fused lane=packet, 2 warp-specialised
// auto: the warp-specialised kernel up to two resident waves of its CTAs (three 74 KB CTAs fit one
// SM).  Measured crossover with the lane=packet kernel on B200: 192 MiB 1.74 vs 1.92 ms, 256 MiB
// 2.05 vs 1.94 ms (profiles/r1_end_work_unit_and_size_sweep.jsonl).
static size_t g_ws_max_packets = (size_t)148 * 3 * 32 * 2;

static cudaError_t encode_slots(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t stride, uint32_t *d_sizes,
                                uint32_t packet, cudaStream_t st, const ShardPlace *where = nullptr, uint64_t call = 0,
                                uint64_t *d_acc = nullptr)
{
    const size_t packets = packets_of(n, packet);
    const bool ws = g_encode_path == 2 || (g_encode_path == 0 && packets <= g_ws_max_packets);
    return ws ? launch_encode_slots_ws(d_in, n, d_slots, stride, d_sizes, packet, st, where, call, d_acc)
              : launch_encode_slots(d_in, n, d_slots, stride, d_sizes, packet, st, where, call, d_acc);
}

// ---- optional per-kernel timing (bench.py's roofline): CUDA events recorded on the
// caller's stream around each kernel of an entry point; off by default
struct Span { cudaEvent_t a, b; int what; };
static bool g_profile = false;
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event()
{
    cudaEvent_t e = nullptr;
    if (!g_event_pool.empty()) { e = g_event_pool.back(); g_event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
struct Scope {                                   // one timed span on stream st
    Span s{};
    cudaStream_t st;
    bool on;
    Scope(int what, cudaStream_t stream) : st(stream), on(g_profile)
    {
        if (!on) return;
        s.what = what; s.a = take_event(); s.b = take_event();
        cudaEventRecord(s.a, st);
    }
    ~Scope()
    {
        if (!on) return;
        cudaEventRecord(s.b, st);
        g_spans.push_back(s);
    }
};

}  // namespace gpuar

using namespace gpuar;

extern "C" {

int gpuar_b200_abi_version(void) { return 1; }

const char *gpuar_b200_strerror(int code)
{
    switch (code) {
    case 0: return "ok";
    case GPUAR_E_ARG: return "bad argument or buffer too small";
    case GPUAR_E_FORMAT: return "malformed .gip header or packet chain";
    case GPUAR_E_NODEVICE: return "no usable CUDA device";
    case GPUAR_E_UNSUPPORTED: return "stream layout not supported by the device path";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int gpuar_b200_init(void)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return GPUAR_E_NODEVICE;
    if ((e = cudaFree(nullptr)) != cudaSuccess) return ck(e);
    // the kernels are built for sm_100a only: fail here, loudly, on anything else
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, probe_kernel());
    if (e != cudaSuccess) { cudaGetLastError(); return GPUAR_E_NODEVICE; }
    return 0;
}

size_t gpuar_b200_packets(size_t n) { return packets_of(n); }
size_t gpuar_b200_payload_bound(size_t n) { return packets_of(n) * (size_t)kSlot + GPUAR_PAD_BYTES; }
size_t gpuar_b200_encode_scratch_bytes(size_t n) { return encode_plan(n).total; }
size_t gpuar_b200_index_scratch_bytes(size_t c) { return index_scratch_bytes(c); }
uint64_t gpuar_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t gpuar_b200_payload_bound_ex(size_t n, size_t packet_bytes)
{
    return packet_ok(packet_bytes) ? packets_of(n, packet_bytes) * (packet_bytes + 512) + GPUAR_PAD_BYTES : 0;
}
size_t gpuar_b200_encode_scratch_bytes_ex(size_t n, size_t packet_bytes)
{
    return packet_ok(packet_bytes) ? encode_plan(n, packet_bytes).total : 0;
}

int gpuar_b200_encode_ex(const uint8_t *d_in, size_t n, size_t packet_bytes, uint8_t *d_payload, size_t payload_cap,
                         uint64_t *d_payload_bytes, uint32_t *d_packet_sizes, void *d_scratch, size_t scratch_bytes,
                         void *stream)
{
    if (!packet_ok(packet_bytes)) return GPUAR_E_ARG;
    const EncodePlan p = encode_plan(n, packet_bytes);
    const uint32_t slot = (uint32_t)packet_bytes + 512u;
    if (!d_payload_bytes || (n && (!d_in || !d_payload || !d_scratch))) return GPUAR_E_ARG;
    if (((uintptr_t)d_in | (uintptr_t)d_payload | (uintptr_t)d_scratch) & 15u) return GPUAR_E_ARG;
    if (payload_cap < gpuar_b200_payload_bound_ex(n, packet_bytes) || scratch_bytes < p.total) return GPUAR_E_ARG;
    if (p.packets > 0xFFFFFFF0ull) return GPUAR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return ck(cudaMemsetAsync(d_payload_bytes, 0, sizeof(uint64_t), st));   // no packets: scratch may be NULL
    uint8_t *s = static_cast<uint8_t *>(d_scratch);
    uint32_t *sizes = d_packet_sizes ? d_packet_sizes : reinterpret_cast<uint32_t *>(s + p.off_sizes);
    cudaError_t e;
    {
        Scope t(GPUAR_SPAN_ENCODE, st);
        e = encode_slots(d_in, n, s + p.off_slots, slot, sizes, (uint32_t)packet_bytes, st);
    }
    if (e != cudaSuccess) return ck(e);
    {
        Scope t(GPUAR_SPAN_COMPACT, st);
        e = launch_compact(s + p.off_slots, slot, sizes, (uint32_t)p.packets, d_payload,
                           reinterpret_cast<uint64_t *>(s + p.off_desc), d_payload_bytes, st);
    }
    return ck(e);
}

int gpuar_b200_encode(const uint8_t *d_in, size_t n, uint8_t *d_payload, size_t payload_cap,
                      uint64_t *d_payload_bytes, uint32_t *d_packet_sizes, void *d_scratch,
                      size_t scratch_bytes, void *stream)
{
    return gpuar_b200_encode_ex(d_in, n, kPacket, d_payload, payload_cap, d_payload_bytes, d_packet_sizes, d_scratch,
                                scratch_bytes, stream);
}

int gpuar_b200_index_ex(const uint8_t *d_payload, size_t c, size_t packet_bytes, uint64_t *d_offsets,
                        size_t max_packets, uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream)
{
    if (!packet_ok(packet_bytes)) return GPUAR_E_ARG;
    if (!d_result || !d_scratch || (c && (!d_payload || !d_offsets))) return GPUAR_E_ARG;
    if (((uintptr_t)d_payload | (uintptr_t)d_scratch) & 15u) return GPUAR_E_ARG;
    if (scratch_bytes < index_scratch_bytes(c)) return GPUAR_E_ARG;
    Scope t(GPUAR_SPAN_INDEX, (cudaStream_t)stream);
    return ck(launch_index(d_payload, c, d_offsets, max_packets, d_result, d_scratch, scratch_bytes,
                           (uint32_t)packet_bytes, (cudaStream_t)stream));
}

int gpuar_b200_index(const uint8_t *d_payload, size_t c, uint64_t *d_offsets, size_t max_packets,
                     uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream)
{
    return gpuar_b200_index_ex(d_payload, c, kPacket, d_offsets, max_packets, d_result, d_scratch, scratch_bytes, stream);
}

int gpuar_b200_decode_ex(const uint8_t *d_payload, size_t c, size_t packet_bytes, const uint64_t *d_offsets,
                         size_t n_packets, uint8_t *d_out, size_t out_cap, void *stream)
{
    if (!packet_ok(packet_bytes)) return GPUAR_E_ARG;
    if (!n_packets) return 0;
    if (!d_payload || !d_offsets || !d_out) return GPUAR_E_ARG;
    if (((uintptr_t)d_payload | (uintptr_t)d_out) & 15u) return GPUAR_E_ARG;
    if (n_packets > 0xFFFFFFF0ull || out_cap < n_packets * packet_bytes) return GPUAR_E_ARG;
    Scope t(GPUAR_SPAN_DECODE, (cudaStream_t)stream);
    return ck(launch_decode(d_payload, c + GPUAR_PAD_BYTES, d_offsets, 0, (uint32_t)n_packets, d_out,
                            (uint32_t)packet_bytes, (cudaStream_t)stream));
}

int gpuar_b200_decode(const uint8_t *d_payload, size_t c, const uint64_t *d_offsets, size_t n_packets,
                      uint8_t *d_out, size_t out_cap, void *stream)
{
    return gpuar_b200_decode_ex(d_payload, c, kPacket, d_offsets, n_packets, d_out, out_cap, stream);
}

/* packed decode scratch: [strided output: packets*packet][sizes: packets*4][descriptors] */
struct PackedPlan {
    size_t off_strided, off_sizes, off_desc, total;
};
static PackedPlan packed_plan(size_t packets, size_t packet)
{
    PackedPlan p{};
    size_t o = 0;
    p.off_strided = o;
    o += align_up(packets * packet + 64, 256);
    p.off_sizes = o;
    o += align_up(packets * 4 + 4, 256);
    p.off_desc = o;
    o += align_up(compact_desc_bytes(packets), 256);
    p.total = o;
    return p;
}

size_t gpuar_b200_decode_packed_scratch_bytes(size_t n_packets, size_t packet_bytes)
{
    return packet_ok(packet_bytes) ? packed_plan(n_packets, packet_bytes).total : 0;
}

int gpuar_b200_decode_packed(const uint8_t *d_payload, size_t c, size_t packet_bytes, const uint64_t *d_offsets,
                             size_t n_packets, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bytes, void *d_scratch,
                             size_t scratch_bytes, void *stream)
{
    if (!packet_ok(packet_bytes) || !d_out_bytes) return GPUAR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!n_packets) return ck(cudaMemsetAsync(d_out_bytes, 0, sizeof(uint64_t), st));
    if (!d_payload || !d_offsets || !d_out || !d_scratch) return GPUAR_E_ARG;
    if (((uintptr_t)d_payload | (uintptr_t)d_out | (uintptr_t)d_scratch) & 15u) return GPUAR_E_ARG;
    const PackedPlan p = packed_plan(n_packets, packet_bytes);
    if (n_packets > 0xFFFFFFF0ull || scratch_bytes < p.total) return GPUAR_E_ARG;
    uint8_t *s = static_cast<uint8_t *>(d_scratch);
    uint32_t *sizes = reinterpret_cast<uint32_t *>(s + p.off_sizes);
    cudaError_t e;
    {
        Scope t(GPUAR_SPAN_DECODE, st);
        e = launch_decode(d_payload, c + GPUAR_PAD_BYTES, d_offsets, 0, (uint32_t)n_packets, s + p.off_strided,
                          (uint32_t)packet_bytes, st);
    }
    if (e != cudaSuccess) return ck(e);
    Scope t(GPUAR_SPAN_COMPACT, st);
    e = launch_raw_sizes(d_payload, c, d_offsets, (uint32_t)n_packets, (uint32_t)packet_bytes, sizes, st);
    if (e != cudaSuccess) return ck(e);
    return ck(launch_compact(s + p.off_strided, (uint32_t)packet_bytes, sizes, (uint32_t)n_packets, d_out,
                             reinterpret_cast<uint64_t *>(s + p.off_desc), d_out_bytes, st, (uint64_t)out_cap));
}

/* ------------------------------------------------------------------- options */
int gpuar_b200_set_option(int key, long long value)
{
    switch (key) {
    case GPUAR_OPT_ENCODE_PATH:
        if (value < 0 || value > 2) return GPUAR_E_ARG;
        g_encode_path = (int)value;
        return 0;
    case GPUAR_OPT_WS_MAX_PACKETS:
        if (value < 0) return GPUAR_E_ARG;
        g_ws_max_packets = (size_t)value;
        return 0;
    case GPUAR_OPT_COMPACT_TILE:
        if (value < 0 || value > 128 || !set_compact_tile((uint32_t)value)) return GPUAR_E_ARG;
        return 0;
    case GPUAR_OPT_DECODE_PATH:
        return set_decode_path((int)value) && value == (int)value ? 0 : GPUAR_E_ARG;
    default:
        return GPUAR_E_ARG;
    }
}

/* ----------------------------------------------------------------- profiling */
void gpuar_b200_profile(int enable) { g_profile = enable != 0; }

int gpuar_b200_profile_read(double ms[GPUAR_SPAN_COUNT], uint64_t calls[GPUAR_SPAN_COUNT])
{
    for (int k = 0; k < GPUAR_SPAN_COUNT; ++k) { ms[k] = 0; calls[k] = 0; }
    cudaError_t bad = cudaSuccess;
    for (const Span &s : g_spans) {
        float t = 0;
        cudaError_t e = cudaEventSynchronize(s.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, s.a, s.b);
        if (e != cudaSuccess) bad = e;
        else if (s.what >= 0 && s.what < GPUAR_SPAN_COUNT) { ms[s.what] += t; calls[s.what] += 1; }
        g_event_pool.push_back(s.a);
        g_event_pool.push_back(s.b);
    }
    g_spans.clear();
    return ck(bad);
}

int gpuar_b200_selfcheck(uint64_t *mismatches)
{
    if (!mismatches) return GPUAR_E_ARG;
    uint64_t *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(uint64_t));
    if (e != cudaSuccess) return ck(e);
    e = launch_selfcheck(d, 0);
    if (e == cudaSuccess) e = cudaMemcpy(mismatches, d, sizeof(uint64_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return ck(e);
}

/* ------------------------------------------------------------------ header */
void gpuar_b200_write_header(uint8_t hdr[20], uint64_t raw_bytes, uint64_t gip_bytes)
{
    memset(hdr, 0, GPUAR_FILE_HEADER);
    hdr[0] = 0; hdr[1] = 1; hdr[2] = 0;                 /* file_header.hpp:25-27,33-35 */
    hdr[3] = GPUAR_HEADER_WIDE_MARK;                    /* never written by the reference: marks 64-bit size fields */
    for (int k = 0; k < 8; ++k) {
        hdr[4 + k] = (uint8_t)(raw_bytes >> (8 * k));   /* :61-66 defines the low 4 bytes */
        hdr[12 + k] = (uint8_t)(gip_bytes >> (8 * k));  /* :67-72 */
    }
}

int gpuar_b200_check_header(const uint8_t hdr[20])
{
    return (hdr[0] == 0 && hdr[1] == 1 && hdr[2] == 0) ? 0 : GPUAR_E_FORMAT;   /* file_header.hpp:74-77 */
}

int gpuar_b200_gip_raw_size(const uint8_t *gip, size_t gip_bytes, uint64_t *raw_bytes)
{
    if (!gip || !raw_bytes || gip_bytes < GPUAR_FILE_HEADER) return GPUAR_E_FORMAT;
    if (gpuar_b200_check_header(gip)) return GPUAR_E_FORMAT;
    uint64_t lo = 0, hi = 0;
    for (int k = 0; k < 4; ++k) {
        lo |= (uint64_t)gip[4 + k] << (8 * k);
        hi |= (uint64_t)gip[8 + k] << (8 * k);
    }
    /* bytes 3 and 8-11 are uninitialised in reference-written files: the high half is trusted only
     * under this library's mark in byte 3 AND if the resulting size is plausible for this payload:
     * a stream of P full packets holds more than (P-1)*8192 raw bytes, P <= payload/210 + 1 (8192
     * equal bytes code into 210), and arithmetic coding with this model expands by < 7 % */
    const uint64_t payload = gip_bytes - GPUAR_FILE_HEADER;
    const uint64_t wide = lo | (hi << 32);
    const bool marked = gip[3] == GPUAR_HEADER_WIDE_MARK;
    const bool plausible = wide <= (payload / 210 + 1) * (uint64_t)kPacket && wide + wide / 14 + kSlot >= payload;
    *raw_bytes = (hi && marked && plausible) ? wide : lo;
    return 0;
}

int gpuar_b200_gip_walk(const uint8_t *gip, size_t gip_bytes, uint64_t *packets, uint64_t *raw_bytes)
{
    if (!gip || gip_bytes < GPUAR_FILE_HEADER || gpuar_b200_check_header(gip)) return GPUAR_E_FORMAT;
    const uint8_t *pay = gip + GPUAR_FILE_HEADER;
    const size_t c = gip_bytes - GPUAR_FILE_HEADER;
    uint64_t n = 0, raw = 0;
    for (size_t pos = 0; pos < c;) {                     /* cpu_compressor.cpp:47-78: the chain is the only index */
        if (c - pos < kHdr) return GPUAR_E_FORMAT;
        const size_t len = (size_t)pay[pos] | ((size_t)pay[pos + 1] << 8);
        const size_t r = (size_t)pay[pos + 2] | ((size_t)pay[pos + 3] << 8);
        if (len <= kHdr || len > c - pos) return GPUAR_E_FORMAT;
        if (r == 0 || r > kPacket) return GPUAR_E_UNSUPPORTED;
        ++n;
        raw += r;
        pos += len;
    }
    if (packets) *packets = n;
    if (raw_bytes) *raw_bytes = raw;
    return 0;
}

/* ------------------------------------------------------------ multi-GPU */
int gpuar_b200_peer_concat(uint8_t *d_dst, int dst_device, size_t dst_offset, const uint8_t *d_src,
                           int src_device, size_t bytes, void *stream)
{
    if (!bytes) return 0;
    if (!d_dst || !d_src) return GPUAR_E_ARG;
    return ck(cudaMemcpyPeerAsync(d_dst + dst_offset, dst_device, d_src, src_device, bytes, (cudaStream_t)stream));
}

static bool shard_ok(const gpuar_b200_shard *sh)
{
    if (!sh || sh->world < 1 || sh->world > GPUAR_MAX_RANKS || sh->rank < 0 || sh->rank >= sh->world) return false;
    if (sh->n_segments != 1 && sh->n_segments != sh->world) return false;
    for (int g = 0; g < sh->n_segments; ++g)
        if (!sh->segments[g] || ((uintptr_t)sh->segments[g] & 15u)) return false;
    for (int r = 0; r < sh->world; ++r)
        if (!sh->mailbox[r] || ((uintptr_t)sh->mailbox[r] & 7u)) return false;
    return true;
}

static ShardPlace shard_place(const gpuar_b200_shard *sh)
{
    ShardPlace w{};
    for (int g = 0; g < sh->n_segments; ++g) w.segment[g] = sh->segments[g];
    for (int r = 0; r < sh->world; ++r) w.mailbox[r] = sh->mailbox[r];
    w.seg_cap = sh->seg_cap;
    w.rank = (uint32_t)sh->rank;
    w.world = (uint32_t)sh->world;
    w.n_segments = (uint32_t)sh->n_segments;
    return w;
}

int gpuar_b200_encode_sharded(gpuar_b200_shard *shard, const uint8_t *d_in, size_t n, uint64_t *d_layout,
                              uint32_t *d_packet_sizes, void *d_scratch, size_t scratch_bytes, void *stream)
{
    if (!shard_ok(shard) || !d_layout || !d_scratch || (n && !d_in)) return GPUAR_E_ARG;
    if (((uintptr_t)d_in | (uintptr_t)d_scratch) & 15u) return GPUAR_E_ARG;
    const EncodePlan p = encode_plan(n);
    if (scratch_bytes < p.total || p.packets > 0xFFFFFFF0ull) return GPUAR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *s = static_cast<uint8_t *>(d_scratch);
    uint32_t *sizes = d_packet_sizes ? d_packet_sizes : reinterpret_cast<uint32_t *>(s + p.off_sizes);
    const ShardPlace where = shard_place(shard);
    const uint64_t call = shard->calls[0]++;
    uint64_t *desc = reinterpret_cast<uint64_t *>(s + p.off_desc);
    cudaError_t e = shard_desc_reset(desc, (uint32_t)p.packets, st);
    if (e != cudaSuccess) return ck(e);
    {
        // the encode kernel also sums the packet sizes; its last CTA stores the total into every rank's mailbox
        Scope t(GPUAR_SPAN_ENCODE, st);
        e = encode_slots(d_in, n, s + p.off_slots, kSlot, sizes, kPacket, st, &where, call, shard_acc(desc, (uint32_t)p.packets));
    }
    if (e != cudaSuccess) return ck(e);
    Scope t(GPUAR_SPAN_COMPACT, st);
    return ck(launch_compact_sharded(s + p.off_slots, kSlot, sizes, (uint32_t)p.packets, desc, d_layout, where, call, st));
}

uint64_t gpuar_b200_shard_segment_bytes(uint64_t stream_bytes, int n_segments)
{
    if (n_segments < 1) return 0;
    const uint64_t seg = ((stream_bytes + (uint64_t)n_segments - 1) / (uint64_t)n_segments + 255) & ~(uint64_t)255;
    return seg < GPUAR_SHARD_MIN_SEGMENT ? GPUAR_SHARD_MIN_SEGMENT : seg;   /* what the compaction kernel computes (shard_segment_bytes) */
}

/* sharded decode scratch: [offsets: max_packets * 8][index scratch of one segment] */
size_t gpuar_b200_decode_sharded_scratch_bytes(uint64_t seg_bytes, size_t max_packets)
{
    return align_up(max_packets * 8 + 8, 256) + index_scratch_bytes((size_t)seg_bytes);
}

int gpuar_b200_decode_sharded(gpuar_b200_shard *shard, uint64_t stream_bytes, uint8_t *d_out, size_t out_cap,
                              uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream)
{
    if (!shard_ok(shard) || shard->n_segments != shard->world || !d_result || !d_scratch) return GPUAR_E_ARG;
    if (((uintptr_t)d_out | (uintptr_t)d_scratch) & 15u) return GPUAR_E_ARG;
    const uint64_t seg = gpuar_b200_shard_segment_bytes(stream_bytes, shard->world);
    const size_t max_packets = out_cap / kPacket;
    if (seg > shard->seg_cap || max_packets > 0xFFFFFFF0ull || (max_packets && !d_out)) return GPUAR_E_ARG;
    if (scratch_bytes < gpuar_b200_decode_sharded_scratch_bytes(seg, max_packets)) return GPUAR_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *s = static_cast<uint8_t *>(d_scratch);
    uint64_t *offsets = reinterpret_cast<uint64_t *>(s);
    const size_t off_index = align_up(max_packets * 8 + 8, 256);
    const ShardPlace where = shard_place(shard);
    const uint64_t call = shard->calls[1]++;
    cudaError_t e = cudaMemsetAsync(d_result, 0, 8 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return ck(e);
    {
        Scope t(GPUAR_SPAN_INDEX, st);
        e = launch_index_segment(where, call, stream_bytes, seg, offsets, max_packets, d_result, s + off_index,
                                 scratch_bytes - off_index, st);
    }
    if (e != cudaSuccess || !max_packets) return ck(e);
    const uint64_t base = seg * (uint64_t)shard->rank;
    const uint64_t here = base < stream_bytes ? (stream_bytes - base < seg ? stream_bytes - base : seg) : 0;
    Scope t(GPUAR_SPAN_DECODE, st);
    return ck(launch_decode(shard->segments[shard->rank], (size_t)here + (base + seg < stream_bytes ? kShardHalo : 64u),
                            offsets, 0, (uint32_t)max_packets, d_out, kPacket, st, d_result));
}

int gpuar_b200_enable_peer(int peer_device)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ck(e);
    if (dev == peer_device) return 0;
    int can = 0;
    if ((e = cudaDeviceCanAccessPeer(&can, dev, peer_device)) != cudaSuccess) return ck(e);
    if (!can) return GPUAR_E_UNSUPPORTED;
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    return ck(e);
}

int gpuar_b200_device_alloc(size_t bytes, void **d_ptr)
{
    if (!d_ptr) return GPUAR_E_ARG;
    return ck(cudaMalloc(d_ptr, bytes));
}

int gpuar_b200_device_free(void *d_ptr) { return ck(cudaFree(d_ptr)); }

int gpuar_b200_host_alloc(size_t bytes, void **h_ptr)
{
    if (!h_ptr) return GPUAR_E_ARG;
    return ck(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocPortable));   // usable from every device's context
}

int gpuar_b200_host_free(void *h_ptr) { return ck(cudaFreeHost(h_ptr)); }

int gpuar_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gpuar_b200_set_device(int device) { return ck(cudaSetDevice(device)); }

int gpuar_b200_ipc_export(const void *d_ptr, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, const_cast<void *>(d_ptr));
    if (e != cudaSuccess) return ck(e);
    memcpy(handle, &hd, 64);
    return 0;
}

int gpuar_b200_ipc_open(const uint8_t handle[64], void **d_ptr)
{
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    return ck(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
}

int gpuar_b200_ipc_close(void *d_ptr) { return ck(cudaIpcCloseMemHandle(d_ptr)); }

/* ------------------------------------------- reference-named shims (gpuar.h:74,77-78) */
void initConstantRange(void) { (void)gpuar_b200_init(); }

void garCompressExecutor(const uint8_t *source, size_t size, uint8_t *destination, uint32_t numBlocks)
{
    (void)numBlocks;
    (void)encode_slots(source, size, destination, kSlot, nullptr, kPacket, (cudaStream_t)0);
}

void garDecompressExecutor(const uint8_t *source, size_t size, uint8_t *destination, uint32_t numBlocks)
{
    (void)numBlocks;
    (void)launch_decode(source, size, nullptr, kSlot, (uint32_t)(size / kSlot), destination, kPacket, (cudaStream_t)0);
}

}  // extern "C"
