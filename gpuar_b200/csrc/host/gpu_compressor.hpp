// gpu_compressor.hpp -- file <-> device driver over the C ABI (mirrors gip::GPUCompressor,
// reference src/gpu_compressor.hpp:8-39).
#pragma once
#include "compressor.hpp"

namespace gip {

class GpuCompressor : public Compressor {
    std::uint8_t *in_ = nullptr;       // page-locked staging: one segment of input
    std::uint8_t *out_ = nullptr;      // page-locked staging: one segment of output
    std::size_t inCap_ = 0, outCap_ = 0;
    std::size_t segmentBytes_;         // raw bytes handled per library call (multiple of 8192)

    void reserve(std::size_t inBytes, std::size_t outBytes);

  public:
    explicit GpuCompressor(std::size_t segmentBytes = (std::size_t)1 << 30);
    ~GpuCompressor() override;
    void chooseDevice(int id);                                  // gpu_compressor.cpp:67-82, but really selects it
    CompressionInfo compress(ProgressMonitor *monitor) override;
    CompressionInfo decompress(ProgressMonitor *monitor) override;
};

}  // namespace gip
