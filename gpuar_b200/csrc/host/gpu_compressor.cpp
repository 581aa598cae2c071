// gpu_compressor.cpp -- the GPU path of the CLI: large page-locked segments in, one library call
// per segment (the library overlaps H2D, kernels and D2H internally), header written last.
//
// Replaces the reference's per-packet pipeline (src/gpu_compressor.cpp:84-395: one 8 KiB
// cudaMemcpyAsync + stream sync + fwrite per packet, host-side compaction, host-side chain walk).
#include "gpu_compressor.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>

#include "../../../include/gpuar_b200.h"

namespace gip {

// GPUAR_B200_TRACE=1: start-up costs (context creation, page-locking) on stderr; they are in
// neither "Compute time" nor "I/O time" (the reference does not count its cudaMallocHost either)
static void trace(const char *what, std::chrono::steady_clock::time_point since)
{
    static const bool on = std::getenv("GPUAR_B200_TRACE") != nullptr;
    if (on)
        std::fprintf(stderr, "[gpuar] %s: %.1f ms\n", what,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - since).count());
}

static void check(int rc, const char *what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + gpuar_b200_strerror(rc));
}

GpuCompressor::GpuCompressor(std::size_t segmentBytes) : segmentBytes_(std::max<std::size_t>(kPacketBytes, segmentBytes / kPacketBytes * kPacketBytes))
{
    const auto t0 = std::chrono::steady_clock::now();
    check(gpuar_b200_init(), "gpuar_b200_init");                // replaces initConstantRange(), gpu_compressor.cpp:19
    trace("context + constants", t0);
}

GpuCompressor::~GpuCompressor()
{
    for (int b = 0; b < 2; ++b) {
        if (in_[b]) gpuar_b200_host_free(in_[b]);
        if (out_[b]) gpuar_b200_host_free(out_[b]);
    }
}

void GpuCompressor::chooseDevice(int id)
{
    check(gpuar_b200_set_device(id), "gpuar_b200_set_device");
    check(gpuar_b200_init(), "gpuar_b200_init");
    device_ = id;
}

void GpuCompressor::useDevices(const std::vector<int> &ids)
{
    if (ids.empty()) throw std::runtime_error("useDevices: empty device list");
    const int count = gpuar_b200_device_count();
    for (int id : ids)
        if (id < 0 || id >= count) throw std::runtime_error("no such CUDA device: " + std::to_string(id));
    chooseDevice(ids[0]);
    devices_ = ids;
}

void GpuCompressor::reserve(std::size_t inBytes, std::size_t outBytes, bool secondInput)
{
    // Page-locking is the expensive part of start-up (the kernel faults in and pins every page,
    // ~0.7 s per GiB on the B200 hosts), and the four buffers pin independently: one helper
    // thread each.  A helper selects the device first so that it does not create a context on
    // device 0.
    const auto t0 = std::chrono::steady_clock::now();
    std::uint8_t **slot[4] = {&in_[0], &in_[1], &out_[0], &out_[1]};
    // the decode path reads into in_[0] only (one sliding window): in_[1] is not pinned for it
    const std::size_t want[4] = {inBytes > inCap_ ? inBytes : 0, secondInput && inBytes > inCap1_ ? inBytes : 0,
                                 outBytes > outCap_ ? outBytes : 0, outBytes > outCap_ ? outBytes : 0};
    std::future<int> pinned[4];
    const int device = device_;
    for (int k = 0; k < 4; ++k) {
        if (!want[k]) continue;
        if (*slot[k]) gpuar_b200_host_free(*slot[k]);
        *slot[k] = nullptr;
        std::uint8_t **dst = slot[k];
        const std::size_t bytes = want[k];
        pinned[k] = std::async(std::launch::async, [dst, bytes, device] {
            if (device >= 0) {
                const int rc = gpuar_b200_set_device(device);
                if (rc) return rc;
            }
            return gpuar_b200_host_alloc(bytes, (void **)dst);
        });
    }
    int rc = 0;
    for (int k = 0; k < 4; ++k)
        if (pinned[k].valid()) {
            const int r = pinned[k].get();
            if (r && !rc) rc = r;
        }
    if (want[0]) inCap_ = rc ? 0 : inBytes;
    if (want[1]) inCap1_ = rc ? 0 : inBytes;
    if (want[2]) outCap_ = rc ? 0 : outBytes;
    check(rc, "gpuar_b200_host_alloc");
    trace("page-locked staging", t0);
}

CompressionInfo GpuCompressor::compress(ProgressMonitor *monitor)
{
    CompressionInfo info;
    StopWatch io, proc;
    monitor->reset();
    File in(openFileName, "rb"), out(saveFileName, "wb");

    io.start();
    info.uncompressedFileSize = in.size();
    std::uint8_t header[kFileHeader] = {0};
    if (std::fwrite(header, kFileHeader, 1, out.get()) != 1) throw std::runtime_error("Write data to file failed");
    io.stop();
    info.compressedFileSize = kFileHeader;

    const std::size_t seg = std::min<std::size_t>(segmentBytes_, std::max<std::size_t>(info.uncompressedFileSize, 1));
    reserve(seg + 16, kFileHeader + gpuar_b200_payload_bound(seg), true);
    // three-stage pipeline over segments: read(i+1) | device(i) | write(i-1)
    auto readSegment = [&](int b) { return std::fread(in_[b], 1, seg, in.get()); };
    auto writeSegment = [&](int b, std::size_t payload) {
        // the per-segment header is dropped: packets are self-delimiting
        return payload == 0 || std::fwrite(out_[b] + kFileHeader, payload, 1, out.get()) == 1;
    };
    io.start();
    std::size_t got = readSegment(0);
    io.stop();
    std::future<bool> writing;
    for (int i = 0; got; ++i) {
        const int b = i & 1;
        std::future<std::size_t> reading;
        if (got == seg) reading = std::async(std::launch::async, readSegment, b ^ 1);
        proc.start();
        std::size_t image = 0;
        // write(i-1) may still be running on out_[b^1]; out_[b] is free (write(i-2) finished before write(i-1) started)
        check(gpuar_b200_compress_host_multi(devices_.empty() ? nullptr : devices_.data(), (int)std::max<std::size_t>(devices_.size(), 1),
                                             in_[b], got, out_[b], outCap_, &image), "gpuar_b200_compress_host");
        proc.stop();
        io.start();
        if (writing.valid() && !writing.get()) throw std::runtime_error("Write compressed data to output file failed");
        const std::size_t payload = image - kFileHeader;
        writing = std::async(std::launch::async, writeSegment, b, payload);
        io.stop();
        info.processedUncompressedSize += got;
        info.compressedFileSize += payload;
        monitor->updateProgress(&info);
        io.start();
        got = reading.valid() ? reading.get() : 0;
        io.stop();
    }
    io.start();
    if (writing.valid() && !writing.get()) throw std::runtime_error("Write compressed data to output file failed");
    io.stop();

    io.start();                                                  // header last, as gpu_compressor.cpp:199-208
    gpuar_b200_write_header(header, info.uncompressedFileSize, info.compressedFileSize);
    if (std::fseek(out.get(), 0, SEEK_SET) != 0 || std::fwrite(header, kFileHeader, 1, out.get()) != 1)
        throw std::runtime_error("Write data to file failed");
    out.close("Write data to file failed");                      // a full disk shows up here at the latest
    io.stop();
    info.processTime = proc.ms();
    info.ioTime = io.ms();
    return info;
}

CompressionInfo GpuCompressor::decompress(ProgressMonitor *monitor)
{
    CompressionInfo info;
    StopWatch io, proc;
    monitor->reset();
    File in(openFileName, "rb"), out(saveFileName, "wb");

    io.start();
    const std::uint64_t fileBytes = in.size();
    std::uint8_t header[kFileHeader];
    if (fileBytes < kFileHeader || std::fread(header, kFileHeader, 1, in.get()) != 1)
        throw std::runtime_error("Incorrect file format");
    io.stop();
    if (gpuar_b200_check_header(header) != 0) throw std::runtime_error("Incorrect file format");
    info.compressedFileSize = fileBytes;
    std::uint64_t announced = 0;
    for (int k = 0; k < 4; ++k) announced |= (std::uint64_t)header[4 + k] << (8 * k);   // low 32 bits are what the reference defines
    info.uncompressedFileSize = announced;

    // Segments of whole packets: the host only hops over compLen fields to find where to cut
    // (one u16 per packet); the packets themselves are indexed and decoded on the device.
    // The staging is page-locked and pinning costs ~1 s per GiB: size it for this file, not for
    // the largest segment.  The announced raw size only bounds the segment (a header that lies just
    // means more or fewer library calls), and never below the payload size: arithmetic coding with
    // this model expands by at most 6 %, so a file whose header says less than that is lying.
    const std::uint64_t payloadBytes = fileBytes - kFileHeader;
    std::uint64_t hint = 0;
    if (gpuar_b200_gip_raw_size(header, (std::size_t)fileBytes, &hint) != 0) hint = announced;
    hint = std::max<std::uint64_t>(std::max(hint, payloadBytes), kPacketBytes);
    const std::size_t segRaw = (std::size_t)std::min<std::uint64_t>(
        segmentBytes_, (hint + kPacketBytes - 1) / kPacketBytes * kPacketBytes);
    const std::size_t segPayload = (std::size_t)std::min<std::uint64_t>(segRaw + segRaw / 16, payloadBytes + kSlotBytes);
    reserve(kFileHeader + segPayload + kSlotBytes + 64, segRaw + 4 * kPacketBytes, false);
    // window [begin, end) of the staging buffer holds payload bytes not yet decoded
    std::size_t begin = 0, end = 0;
    std::uint64_t remaining = fileBytes - kFileHeader;
    std::uint64_t produced = 0;
    std::uint8_t *const pay = in_[0] + kFileHeader;
    std::future<bool> writing;
    int ob = 0;
    while (remaining || end > begin) {
        if (begin && (end - begin < kSlotBytes || begin > segPayload / 2)) {      // make room at the tail
            std::memmove(pay, pay + begin, end - begin);
            end -= begin;
            begin = 0;
        }
        io.start();
        const std::size_t want = (std::size_t)std::min<std::uint64_t>(remaining, segPayload - end);
        if (want && std::fread(pay + end, 1, want, in.get()) != want) throw std::runtime_error("Invalid file length");
        io.stop();
        remaining -= want;
        end += want;
        // cut after the last complete packet; bound the raw size so the output staging cannot overflow
        std::size_t cut = begin, raw = 0;
        while (cut + 4 <= end) {
            const std::size_t len = (std::size_t)pay[cut] | ((std::size_t)pay[cut + 1] << 8);
            const std::size_t r = (std::size_t)pay[cut + 2] | ((std::size_t)pay[cut + 3] << 8);
            if (len <= 4 || r > kPacketBytes) throw std::runtime_error("Incorrect file format");
            if (cut + len > end || raw + r > segRaw + kPacketBytes) break;
            cut += len;
            raw += r;
        }
        if (cut == begin) throw std::runtime_error(remaining ? "Incorrect file format" : "Invalid file length");
        proc.start();
        std::uint8_t *image = pay + begin - kFileHeader;         // a .gip image of just this segment
        gpuar_b200_write_header(image, raw, kFileHeader + (cut - begin));
        std::size_t got = 0;
        check(gpuar_b200_decompress_host_multi(devices_.empty() ? nullptr : devices_.data(), (int)std::max<std::size_t>(devices_.size(), 1),
                                               image, kFileHeader + (cut - begin), out_[ob], outCap_, &got),
              "gpuar_b200_decompress_host");
        proc.stop();
        if (got != raw) throw std::runtime_error("Incorrect file format");
        io.start();                                              // write(i) overlaps read + device work of i+1
        if (writing.valid() && !writing.get()) throw std::runtime_error("Write uncompressed data to output file failed");
        {
            std::uint8_t *src = out_[ob];
            std::FILE *f = out.get();
            writing = std::async(std::launch::async, [src, got, f] { return got == 0 || std::fwrite(src, got, 1, f) == 1; });
        }
        ob ^= 1;
        io.stop();
        begin = cut;
        produced += got;
        info.processedUncompressedSize = produced;
        if (produced > info.uncompressedFileSize) info.uncompressedFileSize = produced;
        monitor->updateProgress(&info);
    }
    io.start();
    if (writing.valid() && !writing.get()) throw std::runtime_error("Write uncompressed data to output file failed");
    out.close("Write uncompressed data to output file failed");
    io.stop();
    info.uncompressedFileSize = produced;
    info.processTime = proc.ms();
    info.ioTime = io.ms();
    return info;
}

}  // namespace gip
