// compress_info.hpp -- in-memory statistics of one run (mirrors gip::CompressionInfo,
// reference src/compress_info.hpp:9-26; it is not an on-disk structure).
#pragma once
#include <cstddef>

namespace gip {

struct CompressionInfo {
    double ratio = 0;                        // compressed / uncompressed
    double processTime = 0;                  // milliseconds inside the codec (device work + its transfers)
    double ioTime = 0;                       // milliseconds of file I/O
    std::size_t processedUncompressedSize = 0;
    std::size_t compressedFileSize = 0;
    std::size_t uncompressedFileSize = 0;
};

}  // namespace gip
