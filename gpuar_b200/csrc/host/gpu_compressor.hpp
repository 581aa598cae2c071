// gpu_compressor.hpp -- file <-> device driver over the C ABI (mirrors gip::GPUCompressor,
// reference src/gpu_compressor.hpp:8-39).
#pragma once
#include <vector>

#include "compressor.hpp"

namespace gip {

class GpuCompressor : public Compressor {
    // page-locked staging, double buffered: while the device works on segment i, segment i+1 is
    // read from the input file and segment i-1 is written to the output file by helper threads
    std::uint8_t *in_[2] = {nullptr, nullptr};
    std::uint8_t *out_[2] = {nullptr, nullptr};
    std::size_t inCap_ = 0, inCap1_ = 0, outCap_ = 0;
    std::size_t segmentBytes_;         // raw bytes handled per library call (multiple of 8192)
    int device_ = -1;                  // chooseDevice() argument; -1 = the process default (device 0)
    std::vector<int> devices_;         // useDevices(): the segments' chunks rotate over these GPUs (empty = device_ alone)

    void reserve(std::size_t inBytes, std::size_t outBytes, bool secondInput);

  public:
    explicit GpuCompressor(std::size_t segmentBytes = (std::size_t)128 << 20);
    ~GpuCompressor() override;
    void chooseDevice(int id);                                  // gpu_compressor.cpp:67-82, but really selects it
    void useDevices(const std::vector<int> &ids);               // --gpus=N: several GPUs of the box for one file
    CompressionInfo compress(ProgressMonitor *monitor) override;
    CompressionInfo decompress(ProgressMonitor *monitor) override;
};

}  // namespace gip
