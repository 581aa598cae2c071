// compressor.hpp -- the abstract driver the CLI talks to (mirrors gip::Compressor, reference
// src/compressor.hpp:10-65: file names in, CompressionInfo out, compress/decompress virtual).
// Differences, all deliberate: a virtual destructor (the reference's is not, hpp:50); errors are
// std::runtime_error (the reference throws string literals that main() cannot catch);
// no CUDA allocation in the base class (the reference's constructor needs a device even for
// --host, compressor.cpp:23-25).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "compress_info.hpp"
#include "progress_monitor.hpp"

namespace gip {

constexpr std::size_t kPacketBytes = 8192;     // UNCOMPRESSED_PACKET_SIZE, gpu.h:13
constexpr std::size_t kSlotBytes = 8704;       // COMPRESSED_PACKET_SIZE,   gpu.h:12
constexpr std::size_t kFileHeader = 20;        // FileHeader::HEADER_LENGTH, file_header.hpp:19-22

class StopWatch {                              // cumulative milliseconds, like the reference's timers
    double total_ = 0;
    std::chrono::steady_clock::time_point t0_;

  public:
    void start() { t0_ = std::chrono::steady_clock::now(); }
    void stop() { total_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count(); }
    double ms() const { return total_; }
};

class File {                                   // RAII FILE*
    std::FILE *f_ = nullptr;

  public:
    File(const std::string &name, const char *mode) : f_(std::fopen(name.c_str(), mode))
    {
        if (!f_) throw std::runtime_error("Can not open file: " + name);
    }
    ~File() { if (f_) std::fclose(f_); }
    File(const File &) = delete;
    File &operator=(const File &) = delete;
    std::FILE *get() const { return f_; }
    void close(const char *what)                   // flush + close with the error reported (ENOSPC shows up here)
    {
        std::FILE *f = f_;
        f_ = nullptr;
        if (f && (std::fflush(f) != 0 || std::fclose(f) != 0)) throw std::runtime_error(what);
    }
    std::uint64_t size()
    {
        const long at = std::ftell(f_);
        std::fseek(f_, 0, SEEK_END);
        const long n = std::ftell(f_);
        std::fseek(f_, at, SEEK_SET);
        return (std::uint64_t)n;
    }
};

class Compressor {
  protected:
    std::string openFileName;
    std::string saveFileName;

  public:
    virtual ~Compressor() = default;
    void setOpenFileName(const std::string &name) { openFileName = name; }
    void setSaveFileName(const std::string &name) { saveFileName = name; }
    virtual CompressionInfo compress(ProgressMonitor *monitor) = 0;
    virtual CompressionInfo decompress(ProgressMonitor *monitor) = 0;
};

}  // namespace gip
