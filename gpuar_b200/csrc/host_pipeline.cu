// host_pipeline.cu -- the host-buffer entry points of the C ABI: a whole .gip image in host memory
// <-> raw bytes in host memory, over one or several GPUs of the box, from one host thread.
//
// Replaces the reference's GPUCompressor::compress / decompress staging (src/gpu_compressor.cpp:84-395:
// one 8 KiB cudaMemcpyAsync + stream synchronisation per packet, host-side compaction, host-side
// chain walk into 8704-byte slots).  Here the input is cut into CHUNKS of whole packets; chunk k goes
// to device k mod G (G = devices in use), where it occupies one of a few LANES: a set of device
// buffers and a stream.  Per device all host->device copies travel on one `up` stream and all
// device->host copies on one `down` stream, in chunk order (copies issued on different streams
// share the link in time slices, so that every chunk arrives late; in order, chunk k is complete
// after (k+1)/chunks of the transfer time), and the kernels of a chunk run on the lane's own
// stream: H2D | kernels | D2H overlap within a device, and the devices' PCIe links run side by side.
// The host thread only issues work and waits for the 8-byte payload total of a chunk (written by
// the compaction kernel straight into pinned memory) to know where the next one lands in the image.
#include "../../include/gpuar_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace gpuar {

static inline size_t align_up_h(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DeviceBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes)                 // on the current device
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
};

constexpr int kLanes = 16;                         // lanes per device (a single device uses all of them)
struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr, arrived = nullptr, drained = nullptr;
    DeviceBuf in, out, scratch, offsets;
    uint64_t *h_off = nullptr;                     // pinned: packet offsets of the lane's chunk (decode)
    size_t h_off_cap = 0;
    bool used = false;
};
struct HostPath {                                  // everything one device needs for the pipelines
    int device = -1;
    std::mutex mu;                                 // one pipeline at a time per device
    // All host->device copies go through `up` and all device->host copies through `down`, in chunk order
    cudaStream_t up = nullptr, down = nullptr;
    Lane lane[kLanes];
    uint64_t *h_total = nullptr;                   // pinned: per-lane payload totals
};
static std::mutex g_paths_mu;
static std::vector<HostPath *> g_paths;

// tuning aid: GPUAR_B200_HOST_CHUNKS=<n> overrides the chunk count of the host-buffer pipelines
static size_t host_chunks(size_t dflt)
{
    static const long v = [] { const char *e = getenv("GPUAR_B200_HOST_CHUNKS"); return e ? atol(e) : 0L; }();
    return v > 0 ? (size_t)v : dflt;
}

static int ck(cudaError_t e) { return (int)e; }

// the path of `device` (created on first use); leaves `device` current
static int host_path(int device, HostPath **out)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return GPUAR_E_NODEVICE; }
    if (device < 0 || device >= count) return GPUAR_E_ARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { cudaGetLastError(); return GPUAR_E_NODEVICE; }   // do not leave the error for a later call to find
    std::lock_guard<std::mutex> lock(g_paths_mu);
    for (HostPath *h : g_paths)
        if (h->device == device) { *out = h; return 0; }
    HostPath *h = new HostPath();
    h->device = device;
    for (int i = 0; i < kLanes; ++i) {
        Lane &l = h->lane[i];
        if ((e = cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking)) != cudaSuccess) return ck(e);
        if ((e = cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming)) != cudaSuccess) return ck(e);
        if ((e = cudaEventCreateWithFlags(&l.arrived, cudaEventDisableTiming)) != cudaSuccess) return ck(e);
        if ((e = cudaEventCreateWithFlags(&l.drained, cudaEventDisableTiming)) != cudaSuccess) return ck(e);
    }
    if ((e = cudaStreamCreateWithFlags(&h->up, cudaStreamNonBlocking)) != cudaSuccess) return ck(e);
    if ((e = cudaStreamCreateWithFlags(&h->down, cudaStreamNonBlocking)) != cudaSuccess) return ck(e);
    if ((e = cudaHostAlloc((void **)&h->h_total, (kLanes + 8) * sizeof(uint64_t), cudaHostAllocPortable)) != cudaSuccess)
        return ck(e);
    g_paths.push_back(h);
    *out = h;
    return 0;
}

// The devices of one call: paths locked in device order (no deadlock between overlapping calls),
// the caller's current device restored at the end.
struct Crew {
    std::vector<HostPath *> path;                  // in the caller's order: chunk k -> path[k % G]
    std::vector<HostPath *> locked;
    int home = 0;
    int lanes = kLanes;                            // lanes per device in use
    ~Crew()
    {
        for (auto it = locked.rbegin(); it != locked.rend(); ++it) (*it)->mu.unlock();
        cudaSetDevice(home);
    }
    int open(const int *devices, int n)
    {
        if (cudaGetDevice(&home) != cudaSuccess) return GPUAR_E_NODEVICE;
        int cur = home;
        if (!devices) { devices = &cur; n = 1; }
        if (n < 1 || n > GPUAR_MAX_RANKS) return GPUAR_E_ARG;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j)
                if (devices[i] == devices[j]) return GPUAR_E_ARG;
        for (int i = 0; i < n; ++i) {
            HostPath *h = nullptr;
            const int rc = host_path(devices[i], &h);
            if (rc) return rc;
            path.push_back(h);
        }
        std::vector<HostPath *> order = path;
        for (size_t i = 1; i < order.size(); ++i)
            for (size_t j = i; j > 0 && order[j]->device < order[j - 1]->device; --j) std::swap(order[j], order[j - 1]);
        for (HostPath *h : order) { h->mu.lock(); locked.push_back(h); }
        // a lane per chunk for inputs of up to 16 chunks per device (nothing waits for a lane then);
        // measured on one device: flat from 8 to 16 chunks in flight at 64 MiB, worse outside
        lanes = kLanes;
        return 0;
    }
    size_t slots() const { return path.size() * (size_t)lanes; }
    HostPath &dev(size_t k) const { return *path[(k % slots()) % path.size()]; }
    int lane_index(size_t k) const { return (int)((k % slots()) / path.size()); }
    Lane &lane(size_t k) const { return dev(k).lane[lane_index(k)]; }
    // nothing may still be reading or writing the caller's buffers (or ours) on return
    cudaError_t drain_all() const
    {
        cudaError_t first = cudaSuccess;
        auto keep = [&](cudaError_t e) { if (first == cudaSuccess) first = e; };
        for (HostPath *h : path) {
            keep(cudaSetDevice(h->device));
            for (int l = 0; l < kLanes; ++l) keep(cudaStreamSynchronize(h->lane[l].st));
            keep(cudaStreamSynchronize(h->up));
            keep(cudaStreamSynchronize(h->down));
        }
        return first;
    }
};

static int compress_on(Crew &crew, const uint8_t *in, size_t n, uint8_t *gip, size_t *gip_bytes)
{
    // A chunk's kernels take about the same time from 1 to ~10 000 packets (a packet is a serial
    // chain of 8192 steps), so the end-to-end time is roughly H2D(everything) + one kernel latency
    // + D2H(last chunk): small inputs are cut into as many chunks as there are lanes to shorten that
    // tail (2..32 measured at 64 MiB with tools/e2e_timeline.cu: flat from 8 to 16, worse outside),
    // large ones into 64 MiB chunks to keep enough packets in flight.
    const size_t G = crew.path.size(), L = crew.slots();
    size_t chunk = align_up_h(n / host_chunks(G > 1 ? G * 8 : (size_t)kLanes) + 1, kPacket);
    chunk = chunk < ((size_t)2 << 20) ? ((size_t)2 << 20) : chunk > ((size_t)64 << 20) ? ((size_t)64 << 20) : chunk;
    const size_t chunks = (n + chunk - 1) / chunk;
    size_t pos = GPUAR_FILE_HEADER, next_drain = 0;
    cudaError_t e = cudaSuccess;
    int rc = 0;
    // the payload of chunk j leaves for its place in the image as soon as the host knows its size
    auto drain = [&](size_t j) -> cudaError_t {
        HostPath &h = crew.dev(j);
        Lane &l = crew.lane(j);
        cudaError_t er = cudaSetDevice(h.device);
        if (er == cudaSuccess) er = cudaEventSynchronize(l.done);
        if (er != cudaSuccess) return er;
        const size_t bytes = (size_t)h.h_total[crew.lane_index(j)];
        er = cudaMemcpyAsync(gip + pos, l.out.p, bytes, cudaMemcpyDeviceToHost, h.down);
        if (er == cudaSuccess) er = cudaEventRecord(l.drained, h.down);
        pos += bytes;
        return er;
    };
    for (size_t k = 0; k < chunks && e == cudaSuccess; ++k) {
        // lane reuse: its input buffer is free once the host has seen `done` of the previous
        // occupant, its payload and scratch once that occupant's copy has left
        while (e == cudaSuccess && next_drain + L <= k) e = drain(next_drain++);
        if (e != cudaSuccess) break;
        HostPath &h = crew.dev(k);
        Lane &l = crew.lane(k);
        const size_t off = k * chunk, m = (n - off < chunk) ? n - off : chunk;
        if ((e = cudaSetDevice(h.device)) != cudaSuccess) break;
        if (l.used) e = cudaStreamWaitEvent(l.st, l.drained, 0);
        if (e == cudaSuccess) e = l.in.need(align_up_h(m, 16) + 16);
        if (e == cudaSuccess) e = l.out.need(gpuar_b200_payload_bound(m) + 16);
        if (e == cudaSuccess) e = l.scratch.need(gpuar_b200_encode_scratch_bytes(m));
        if (e != cudaSuccess) break;
        e = cudaMemcpyAsync(l.in.p, in + off, m, cudaMemcpyHostToDevice, h.up);
        if (e == cudaSuccess) e = cudaEventRecord(l.arrived, h.up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(l.st, l.arrived, 0);
        if (e != cudaSuccess) break;
        rc = gpuar_b200_encode((const uint8_t *)l.in.p, m, (uint8_t *)l.out.p, l.out.cap,
                               &h.h_total[crew.lane_index(k)], nullptr, l.scratch.p, l.scratch.cap, l.st);
        if (rc) break;
        e = cudaEventRecord(l.done, l.st);
        l.used = true;
        // chunks that have finished meanwhile start their way down now, not when their lane is needed
        while (e == cudaSuccess && next_drain < k) {
            if (cudaSetDevice(crew.dev(next_drain).device) != cudaSuccess) break;
            const cudaError_t q = cudaEventQuery(crew.lane(next_drain).done);
            if (q == cudaErrorNotReady) break;
            e = q == cudaSuccess ? drain(next_drain++) : q;
        }
    }
    while (e == cudaSuccess && rc == 0 && next_drain < chunks) e = drain(next_drain++);
    {
        const cudaError_t es = crew.drain_all();
        if (e == cudaSuccess) e = es;
    }
    for (HostPath *h : crew.path)
        for (int l = 0; l < kLanes; ++l) h->lane[l].used = false;
    if (rc) return rc;
    if (e != cudaSuccess) return ck(e);
    gpuar_b200_write_header(gip, n, pos);
    *gip_bytes = pos;
    return 0;
}

// The transfers of compress_on without its kernels: the same chunks go up on the same streams and
// out_bytes * (chunk / n) bytes come down per chunk.  What the links (and the host memory behind
// them) can do for this access pattern: the ceiling the end-to-end numbers are read against.
static int link_probe_on(Crew &crew, const uint8_t *in, size_t n, uint8_t *out, size_t out_bytes)
{
    const size_t G = crew.path.size();
    size_t chunk = align_up_h(n / host_chunks(G > 1 ? G * 8 : (size_t)kLanes) + 1, kPacket);
    chunk = chunk < ((size_t)2 << 20) ? ((size_t)2 << 20) : chunk > ((size_t)64 << 20) ? ((size_t)64 << 20) : chunk;
    const size_t chunks = (n + chunk - 1) / chunk;
    cudaError_t e = cudaSuccess;
    size_t pos = 0;
    for (size_t k = 0; k < chunks && e == cudaSuccess; ++k) {
        HostPath &h = crew.dev(k);
        Lane &l = crew.lane(k);
        const size_t off = k * chunk, m = (n - off < chunk) ? n - off : chunk;
        const size_t down = k + 1 == chunks ? out_bytes - pos : (size_t)((double)out_bytes * ((double)m / (double)n));
        if ((e = cudaSetDevice(h.device)) != cudaSuccess) break;
        if (l.used) e = cudaStreamWaitEvent(h.up, l.drained, 0);       // the lane's buffers are still on their way down
        if (e == cudaSuccess) e = l.in.need(align_up_h(m, 16) + 16);
        if (e == cudaSuccess) e = l.out.need((down > gpuar_b200_payload_bound(m) ? down : gpuar_b200_payload_bound(m)) + 16);
        if (e != cudaSuccess) break;
        e = cudaMemcpyAsync(l.in.p, in + off, m, cudaMemcpyHostToDevice, h.up);
        if (e == cudaSuccess) e = cudaEventRecord(l.arrived, h.up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h.down, l.arrived, 0);
        if (e == cudaSuccess && down) e = cudaMemcpyAsync(out + pos, l.out.p, down, cudaMemcpyDeviceToHost, h.down);
        if (e == cudaSuccess) e = cudaEventRecord(l.drained, h.down);
        l.used = true;
        pos += down;
    }
    const cudaError_t es = crew.drain_all();
    for (HostPath *h : crew.path)
        for (int l = 0; l < kLanes; ++l) h->lane[l].used = false;
    return ck(e == cudaSuccess ? es : e);
}

static int decompress_on(Crew &crew, const uint8_t *gip, size_t gip_bytes, uint64_t raw, uint8_t *out, size_t out_cap,
                         size_t *out_bytes)
{
    // The payload is in host memory, so the packet chain is walked here, one u32 per packet, as
    // the reference's host driver does while it reads the file (gpu_compressor.cpp:294-320) -- but
    // only to cut the stream into chunks of whole packets that flow through the lanes:
    // H2D(chunk k+1) | decode(chunk k) | D2H(chunk k-1).  (Device-resident callers use
    // gpuar_b200_index, the parallel chain discovery on the device.)  A decode launch takes about
    // the same time from 1 packet to a full wave, so the end-to-end time is roughly
    // H2D(everything) + one decode latency + D2H(last chunk).
    const size_t c = gip_bytes - GPUAR_FILE_HEADER;
    const uint8_t *pay = gip + GPUAR_FILE_HEADER;
    const size_t G = crew.path.size();
    size_t chunk_bytes = c / host_chunks(G > 1 ? G * 8 : (size_t)kLanes) + 1;   // payload bytes per chunk, before rounding to packets
    if (chunk_bytes < ((size_t)2 << 20)) chunk_bytes = (size_t)2 << 20;
    if (chunk_bytes > ((size_t)64 << 20)) chunk_bytes = (size_t)64 << 20;
    // a chunk of full packets holds at most chunk_bytes / 210 of them (8192 equal bytes code into
    // 210); the bound only cuts chunks of short packets, whose buffers are sized by the packet count
    const size_t chunk_packets = chunk_bytes / 128 + 1024;

    size_t pos = 0, total = 0, k = 0;
    int status = 0, rc = 0;
    cudaError_t e = cudaSuccess;
    while (pos < c && status == 0) {
        HostPath &h = crew.dev(k);
        Lane &l = crew.lane(k);
        if ((e = cudaSetDevice(h.device)) != cudaSuccess) break;
        // the lane's pinned offsets are rewritten below: the previous occupant's copy up must have left
        if (l.used && (e = cudaEventSynchronize(l.arrived)) != cudaSuccess) break;
        if (l.h_off_cap < chunk_packets) {
            if (l.h_off) cudaFreeHost(l.h_off);
            l.h_off = nullptr;
            l.h_off_cap = 0;
            if ((e = cudaHostAlloc((void **)&l.h_off, chunk_packets * sizeof(uint64_t), cudaHostAllocPortable)) != cudaSuccess) break;
            l.h_off_cap = chunk_packets;
        }
        // one chunk: whole packets until chunk_bytes of payload; offsets relative to the 16-byte
        // aligned start of what is copied up
        const size_t a = pos, a16 = a & ~(size_t)15, raw0 = total;
        size_t m = 0;
        bool ragged = false;                                             // a short packet that is not the chunk's last
        size_t last_raw = kPacket;
        while (pos < c && pos - a < chunk_bytes && m < chunk_packets) {
            if (c - pos < kHdr) { status = GPUAR_E_FORMAT; break; }
            const size_t len = (size_t)pay[pos] | ((size_t)pay[pos + 1] << 8);
            const size_t r = (size_t)pay[pos + 2] | ((size_t)pay[pos + 3] << 8);
            if (len <= kHdr || len > c - pos) { status = GPUAR_E_FORMAT; break; }
            if (r == 0 || r > kPacket) { status = GPUAR_E_UNSUPPORTED; break; }
            if (total + r > out_cap || !out) { status = GPUAR_E_ARG; break; }
            ragged = ragged || last_raw != kPacket;
            last_raw = r;
            l.h_off[m++] = pos - a16;
            total += r;
            pos += len;
        }
        if (status || !m) break;
        const size_t bytes = pos - a16, chunk_raw = total - raw0;
        if (l.used) {
            // the previous occupant: its decode has read the input and the offsets, its copy down the output
            e = cudaStreamWaitEvent(h.up, l.done, 0);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(l.st, l.drained, 0);
        }
        if (e == cudaSuccess) e = l.in.need(align_up_h(bytes, 16) + GPUAR_PAD_BYTES + 16);
        if (e == cudaSuccess) e = l.offsets.need(chunk_packets * sizeof(uint64_t));
        if (e == cudaSuccess) e = l.out.need((m + 1) * (size_t)kPacket);
        if (e != cudaSuccess) break;
        uint8_t *d_pay = (uint8_t *)l.in.p;
        // H2D of this chunk's bytes and of its offsets on the `up` stream (in chunk order), decode on
        // the lane's stream, D2H on the `down` stream
        e = cudaMemcpyAsync(d_pay, pay + a16, bytes, cudaMemcpyHostToDevice, h.up);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_pay + bytes, 0, GPUAR_PAD_BYTES, h.up);   // the decoder reads a few bytes past the chunk
        if (e == cudaSuccess) e = cudaMemcpyAsync(l.offsets.p, l.h_off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, h.up);
        if (e == cudaSuccess) e = cudaEventRecord(l.arrived, h.up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(l.st, l.arrived, 0);
        if (e != cudaSuccess) break;
        l.used = true;
        if (!ragged) {
            rc = gpuar_b200_decode(d_pay, bytes, (const uint64_t *)l.offsets.p, m, (uint8_t *)l.out.p, l.out.cap, l.st);
        } else {
            // short packets inside the chunk (never written by the reference, legal for its CPU
            // decoder): decode at the 8192-byte stride into the lane's scratch, then close the gaps
            e = l.scratch.need(gpuar_b200_decode_packed_scratch_bytes(m, kPacket));
            if (e != cudaSuccess) break;
            rc = gpuar_b200_decode_packed(d_pay, bytes, kPacket, (const uint64_t *)l.offsets.p, m, (uint8_t *)l.out.p,
                                          chunk_raw, &h.h_total[crew.lane_index(k)], l.scratch.p, l.scratch.cap, l.st);
        }
        if (rc) break;
        e = cudaEventRecord(l.done, l.st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h.down, l.done, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + raw0, l.out.p, chunk_raw, cudaMemcpyDeviceToHost, h.down);
        if (e == cudaSuccess) e = cudaEventRecord(l.drained, h.down);
        if (e != cudaSuccess) break;
        ++k;
    }
    {
        const cudaError_t es = crew.drain_all();
        if (e == cudaSuccess) e = es;
    }
    for (HostPath *h : crew.path)
        for (int l = 0; l < kLanes; ++l) h->lane[l].used = false;
    if (rc) return rc;
    if (status) return status;
    if (e != cudaSuccess) return ck(e);
    if ((uint32_t)raw != (uint32_t)total) return GPUAR_E_FORMAT;    // header field, file_header.hpp:61-66
    *out_bytes = total;
    return 0;
}

}  // namespace gpuar

using namespace gpuar;

extern "C" {

int gpuar_b200_compress_host_multi(const int *devices, int n_devices, const uint8_t *in, size_t n, uint8_t *gip,
                                   size_t gip_cap, size_t *gip_bytes)
{
    if (!gip || !gip_bytes || (n && !in)) return GPUAR_E_ARG;
    if (gip_cap < GPUAR_FILE_HEADER + gpuar_b200_payload_bound(n)) return GPUAR_E_ARG;
    Crew crew;
    const int rc = crew.open(devices, n_devices);
    return rc ? rc : compress_on(crew, in, n, gip, gip_bytes);
}

int gpuar_b200_decompress_host_multi(const int *devices, int n_devices, const uint8_t *gip, size_t gip_bytes,
                                     uint8_t *out, size_t out_cap, size_t *out_bytes)
{
    if (!gip || !out_bytes) return GPUAR_E_ARG;
    uint64_t raw = 0;
    int rc = gpuar_b200_gip_raw_size(gip, gip_bytes, &raw);
    if (rc) return rc;
    if (gip_bytes == GPUAR_FILE_HEADER) { *out_bytes = 0; return 0; }
    Crew crew;
    if ((rc = crew.open(devices, n_devices))) return rc;
    return decompress_on(crew, gip, gip_bytes, raw, out, out_cap, out_bytes);
}

int gpuar_b200_host_link_probe(const int *devices, int n_devices, const uint8_t *in, size_t n, uint8_t *out,
                               size_t out_bytes)
{
    if (!in || !out || !n) return GPUAR_E_ARG;
    Crew crew;
    const int rc = crew.open(devices, n_devices);
    return rc ? rc : link_probe_on(crew, in, n, out, out_bytes);
}

int gpuar_b200_compress_host(const uint8_t *in, size_t n, uint8_t *gip, size_t gip_cap, size_t *gip_bytes)
{
    return gpuar_b200_compress_host_multi(nullptr, 1, in, n, gip, gip_cap, gip_bytes);
}

int gpuar_b200_decompress_host(const uint8_t *gip, size_t gip_bytes, uint8_t *out, size_t out_cap, size_t *out_bytes)
{
    return gpuar_b200_decompress_host_multi(nullptr, 1, gip, gip_bytes, out, out_cap, out_bytes);
}

}  // extern "C"
