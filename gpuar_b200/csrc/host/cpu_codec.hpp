// cpu_codec.hpp -- scalar packet codec for the CLI's explicit --host mode (the reference's
// CPUCompressor path, src/cpu_compressor.cpp).  It is built from the same closed forms the CUDA
// kernels use (../coder_math.h compiles for the host), one packet at a time.  It is never chosen
// automatically: without --host and without a device the CLI fails.
#pragma once
#include <cstddef>
#include <cstdint>

namespace gip {

// in[n] (n <= 8192) -> slot (>= 8704 bytes, 4-byte aligned); returns compLen incl. the 4-byte packet header
std::uint32_t cpuEncodePacket(const std::uint8_t *in, std::uint32_t n, std::uint8_t *slot);
// packet at pkt (readable 8 bytes past compLen) -> out; returns bytes produced (rawLen)
std::uint32_t cpuDecodePacket(const std::uint8_t *pkt, std::uint8_t *out);

}  // namespace gip
