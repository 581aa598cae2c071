// cpu_compressor.hpp -- the CLI's --host mode (mirrors gip::CPUCompressor, reference
// src/cpu_compressor.hpp / .cpp: single-threaded packet loop, header written last).
#pragma once
#include "compressor.hpp"

namespace gip {

class CpuCompressor : public Compressor {
  public:
    CompressionInfo compress(ProgressMonitor *monitor) override;
    CompressionInfo decompress(ProgressMonitor *monitor) override;
};

}  // namespace gip
