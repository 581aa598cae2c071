/*
 * ref_harness.cu -- thin driver around the UNMODIFIED reference codec.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/gpuar_oracle.c header).  This file is
 * ours; it is compiled TOGETHER WITH /root/reference/src/gpuar_kernel.cu, read
 * in place (never copied into this repo), into oracle/_ref/libgpuar_ref.so by
 * `make -C oracle ref`.  It exposes the reference's own arCompress /
 * arDecompress (gpuar_kernel.cu:487, :848) and its GPU launchers
 * (gpuar_kernel.cu:936-949) behind a plain C ABI so that tests can pin the
 * oracle against them and bench.py can time them.
 *
 * The packet loop mirrors the reference's CPU driver, cpu_compressor.cpp:144-173
 * (encode) and :47-78 (decode).  The reference is single-threaded; `threads`
 * > 1 only partitions the independent packets over host threads (each packet is
 * still coded by the reference's own function, untouched).
 */
#include "gpuar.h"

#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Scratch {
    std::vector<uint8_t> in, out;
    AdaptiveProbabilityRange model;
    Scratch() : in(UNCOMPRESSED_PACKET_SIZE + 64), out(COMPRESSED_PACKET_SIZE + 64) {}
};

template <class F>
void fan_out(size_t items, int threads, F fn)
{
    if (threads < 1) threads = 1;
    if ((size_t)threads > items) threads = items ? (int)items : 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        size_t a = items * t / threads, b = items * (t + 1) / threads;
        pool.emplace_back([=] { fn(a, b); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

/* reference constants, so callers need no reference headers */
int gpuar_ref_packet_bytes(void) { return UNCOMPRESSED_PACKET_SIZE; }
int gpuar_ref_slot_bytes(void) { return COMPRESSED_PACKET_SIZE; }

/* One packet through the reference encoder.  Returns compLen. */
size_t gpuar_ref_encode_packet(const uint8_t *in, size_t n, uint8_t *out)
{
    Scratch s;
    probability_t total;
    memcpy(s.in.data(), in, n);                 /* arCompress reads 16-byte elements */
    initializeAdaptiveProbabilityRangeList(&s.model, total);
    size_t len = arCompress(s.in.data(), (uint16_t)n, s.out.data(), s.model, total);
    memcpy(out, s.out.data(), len);
    return len;
}

/* One packet through the reference decoder.  Returns bytes produced. */
size_t gpuar_ref_decode_packet(const uint8_t *pkt, uint8_t *out)
{
    Scratch s;
    probability_t total;
    size_t len = getCompressedSize(pkt);
    memcpy(s.out.data(), pkt, len);
    memset(s.out.data() + len, 0, 8);           /* bytes the decoder over-reads */
    initializeAdaptiveProbabilityRangeList(&s.model, total);
    size_t n = arDecompress(s.out.data(), (uint16_t)len, s.in.data(), s.model, total);
    memcpy(out, s.in.data(), n);
    return n;
}

/*
 * Whole stream -> fixed-stride slots (COMPRESSED_PACKET_SIZE each), packets
 * partitioned over `threads`.  sizes[p] receives compLen of packet p.
 */
void gpuar_ref_encode_slots(const uint8_t *in, size_t n, uint8_t *slots, uint32_t *sizes, int threads)
{
    size_t packets = (n + UNCOMPRESSED_PACKET_SIZE - 1) / UNCOMPRESSED_PACKET_SIZE;
    fan_out(packets, threads, [=](size_t a, size_t b) {
        Scratch s;
        probability_t total;
        for (size_t p = a; p < b; p++) {
            size_t off = p * UNCOMPRESSED_PACKET_SIZE;
            size_t m = n - off < UNCOMPRESSED_PACKET_SIZE ? n - off : UNCOMPRESSED_PACKET_SIZE;
            memcpy(s.in.data(), in + off, m);
            initializeAdaptiveProbabilityRangeList(&s.model, total);
            size_t len = arCompress(s.in.data(), (uint16_t)m, s.out.data(), s.model, total);
            memcpy(slots + p * COMPRESSED_PACKET_SIZE, s.out.data(), len);
            sizes[p] = (uint32_t)len;
        }
    });
}

/* Whole stream -> .gip payload (bytes from offset 20).  Returns payload length. */
size_t gpuar_ref_encode_stream(const uint8_t *in, size_t n, uint8_t *payload, int threads)
{
    size_t packets = (n + UNCOMPRESSED_PACKET_SIZE - 1) / UNCOMPRESSED_PACKET_SIZE;
    std::vector<uint8_t> slots(packets * COMPRESSED_PACKET_SIZE);
    std::vector<uint32_t> sizes(packets);
    gpuar_ref_encode_slots(in, n, slots.data(), sizes.data(), threads);
    size_t pos = 0;
    for (size_t p = 0; p < packets; p++) {
        memcpy(payload + pos, slots.data() + p * COMPRESSED_PACKET_SIZE, sizes[p]);
        pos += sizes[p];
    }
    return pos;
}

/* .gip payload -> bytes.  Chain walk as in cpu_compressor.cpp:47-78.  Returns bytes produced. */
size_t gpuar_ref_decode_stream(const uint8_t *payload, size_t c, uint8_t *out, int threads)
{
    std::vector<size_t> at, dst;
    size_t pos = 0, produced = 0;
    while (pos + PACKET_HEADER_LENGTH <= c) {
        size_t len = getCompressedSize(payload + pos);
        if (len <= PACKET_HEADER_LENGTH || pos + len > c) break;
        at.push_back(pos);
        dst.push_back(produced);
        produced += getUncompressedSize(payload + pos);
        pos += len;
    }
    const size_t *pat = at.data(), *pdst = dst.data();
    fan_out(at.size(), threads, [=](size_t a, size_t b) {
        Scratch s;
        probability_t total;
        for (size_t p = a; p < b; p++) {
            size_t len = getCompressedSize(payload + pat[p]);
            memcpy(s.out.data(), payload + pat[p], len);
            memset(s.out.data() + len, 0, 8);
            initializeAdaptiveProbabilityRangeList(&s.model, total);
            size_t n = arDecompress(s.out.data(), (uint16_t)len, s.in.data(), s.model, total);
            memcpy(out + pdst[p], s.in.data(), n);
        }
    });
    return produced;
}

/* ---- the reference's original GPU kernels (for the same-box GPU baseline) ---- */
void gpuar_ref_gpu_init(void) { initConstantRange(); }

/* d_src: n input bytes; d_slots: ceil(n/8192)*8704 bytes.  Launch shape as
 * gpu_compressor.cpp:183.  Asynchronous on the legacy default stream. */
void gpuar_ref_gpu_encode(const uint8_t *d_src, size_t n, uint8_t *d_slots)
{
    uint32_t blocks = (uint32_t)((n + (size_t)UNCOMPRESSED_PACKET_SIZE * NUM_THREADS - 1) /
                                 ((size_t)UNCOMPRESSED_PACKET_SIZE * NUM_THREADS));
    garCompressExecutor(d_src, n, d_slots, blocks);
}

/* d_slots: packets*8704 bytes; d_dst: packets*8192 bytes.  gpu_compressor.cpp:355-357. */
void gpuar_ref_gpu_decode(const uint8_t *d_slots, size_t packets, uint8_t *d_dst)
{
    uint32_t blocks = (uint32_t)((packets + NUM_THREADS - 1) / NUM_THREADS);
    garDecompressExecutor(d_slots, packets * (size_t)COMPRESSED_PACKET_SIZE, d_dst, blocks);
}

int gpuar_ref_gpu_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
