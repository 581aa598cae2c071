// progress_monitor.hpp -- "NN%.." each time progress crosses a 10 % boundary, as the reference
// prints it (src/progress_monitor.cpp:16-33).
#pragma once
#include <cstdio>

#include "compress_info.hpp"

namespace gip {

class ProgressMonitor {
    int lastDecile_ = 0;

  public:
    void reset() { lastDecile_ = 0; }
    void updateProgress(const CompressionInfo *info)
    {
        if (!info->uncompressedFileSize) return;
        const int percent = (int)(100.0 * (double)info->processedUncompressedSize / (double)info->uncompressedFileSize);
        if (percent / 10 != lastDecile_) {
            lastDecile_ = percent / 10;
            std::printf("%d%%..", percent);
            if (percent >= 100) std::printf("Closing file..");
            std::fflush(stdout);
        }
    }
};

}  // namespace gip
