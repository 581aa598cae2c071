#include "cpu_compressor.hpp"

#include <cstring>
#include <vector>

#include "cpu_codec.hpp"

namespace gip {

static void putHeader(std::uint8_t h[kFileHeader], std::uint64_t raw, std::uint64_t total)
{
    std::memset(h, 0, kFileHeader);
    h[1] = 1;                                              // version 0.1.0, file_header.hpp:25-27
    h[3] = 0xB2;                                           // GPUAR_HEADER_WIDE_MARK: bytes 8-11 / 16-19 hold the high halves
    for (int k = 0; k < 8; ++k) {
        h[4 + k] = (std::uint8_t)(raw >> (8 * k));
        h[12 + k] = (std::uint8_t)(total >> (8 * k));
    }
}

CompressionInfo CpuCompressor::compress(ProgressMonitor *monitor)
{
    CompressionInfo info;
    StopWatch io, proc;
    monitor->reset();
    File in(openFileName, "rb"), out(saveFileName, "wb");
    info.uncompressedFileSize = in.size();
    std::uint8_t header[kFileHeader] = {0};
    if (std::fwrite(header, kFileHeader, 1, out.get()) != 1) throw std::runtime_error("Write data to file failed");
    info.compressedFileSize = kFileHeader;
    std::vector<std::uint8_t> data(kPacketBytes + 16);
    alignas(16) std::uint8_t slot[kSlotBytes + 16];
    for (;;) {                                             // cpu_compressor.cpp:144-173
        io.start();
        const std::size_t got = std::fread(data.data(), 1, kPacketBytes, in.get());
        io.stop();
        if (!got) break;
        proc.start();
        const std::uint32_t len = cpuEncodePacket(data.data(), (std::uint32_t)got, slot);
        proc.stop();
        io.start();
        if (std::fwrite(slot, len, 1, out.get()) != 1) throw std::runtime_error("Write data to file failed");
        io.stop();
        info.processedUncompressedSize += got;
        info.compressedFileSize += len;
        monitor->updateProgress(&info);
    }
    putHeader(header, info.uncompressedFileSize, info.compressedFileSize);
    if (std::fseek(out.get(), 0, SEEK_SET) != 0 || std::fwrite(header, kFileHeader, 1, out.get()) != 1)
        throw std::runtime_error("Write data to file failed");
    out.close("Write data to file failed");
    info.processTime = proc.ms();
    info.ioTime = io.ms();
    return info;
}

CompressionInfo CpuCompressor::decompress(ProgressMonitor *monitor)
{
    CompressionInfo info;
    StopWatch io, proc;
    monitor->reset();
    File in(openFileName, "rb"), out(saveFileName, "wb");
    info.compressedFileSize = in.size();
    std::uint8_t header[kFileHeader];
    if (std::fread(header, kFileHeader, 1, in.get()) != 1 || header[0] != 0 || header[1] != 1 || header[2] != 0)
        throw std::runtime_error("Incorrect file format");
    for (int k = 0; k < 4; ++k) info.uncompressedFileSize |= (std::size_t)header[4 + k] << (8 * k);
    std::vector<std::uint8_t> pkt(65536 + 16), data(kPacketBytes + 16);
    for (;;) {                                             // cpu_compressor.cpp:47-78
        io.start();
        const std::size_t h = std::fread(pkt.data(), 1, 4, in.get());
        io.stop();
        if (h == 0) break;
        const std::size_t len = (std::size_t)pkt[0] | ((std::size_t)pkt[1] << 8);
        if (h != 4 || len <= 4) throw std::runtime_error("Incorrect file format");
        io.start();
        if (std::fread(pkt.data() + 4, 1, len - 4, in.get()) != len - 4) throw std::runtime_error("Incorrect file format");
        io.stop();
        std::memset(pkt.data() + len, 0, 8);
        proc.start();
        const std::uint32_t n = cpuDecodePacket(pkt.data(), data.data());
        proc.stop();
        io.start();
        if (n && std::fwrite(data.data(), n, 1, out.get()) != 1) throw std::runtime_error("Write raw data to file failed");
        io.stop();
        info.processedUncompressedSize += n;
        monitor->updateProgress(&info);
    }
    out.close("Write raw data to file failed");
    info.uncompressedFileSize = info.processedUncompressedSize;
    info.processTime = proc.ms();
    info.ioTime = io.ms();
    return info;
}

}  // namespace gip
