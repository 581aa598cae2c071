// kernels.h -- host-side launchers of the codec kernels (internal to libgpuar_b200.so)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace gpuar {

// encode.cu
const void *probe_kernel();   // address of a kernel of this library, for image-loadability checks
// `packet` = raw bytes per packet: 8192 in the reference format (gpu.h:13); any multiple of 16 up to
// 16112 is coded correctly (the reference's own limit, compressor.cpp:13)
cudaError_t launch_encode_slots(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                uint32_t *d_sizes, uint32_t packet, cudaStream_t st);
// encode_ws.cu: same contract, six specialised warps per 32 packets (for inputs that cannot fill the GPU)
cudaError_t launch_encode_slots_ws(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                   uint32_t *d_sizes, uint32_t packet, cudaStream_t st);
size_t compact_desc_bytes(size_t packets);
bool set_compact_tile(uint32_t packets_per_tile);   // 0 = automatic, else a power of two 4..128
// `cap` != kNoCap: d_payload holds cap bytes and the sizes are untrusted (decode side): packets that
// would end past cap are skipped, *d_total still reports the full sum
constexpr uint64_t kNoCap = ~0ull;
cudaError_t launch_compact(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                           uint32_t packets, uint8_t *d_payload, uint64_t *d_desc, uint64_t *d_total,
                           cudaStream_t st, uint64_t cap = kNoCap);

cudaError_t launch_shard_concat(const uint8_t *d_payload, const uint64_t *d_totals, uint32_t rank, uint32_t world,
                                uint8_t *const *segments, uint32_t n_segments, uint64_t seg_cap, uint64_t *d_layout,
                                cudaStream_t st);

// decode.cu: d_offsets == nullptr means packet p starts at p * stride (reference slot layout)
cudaError_t launch_decode(const uint8_t *d_payload, size_t readable, const uint64_t *d_offsets, uint32_t stride,
                          uint32_t packets, uint8_t *d_out, uint32_t packet, cudaStream_t st);

// index.cu
// sizes[p] = min(rawLen of the packet at d_offsets[p], packet): what decode writes for packet p
cudaError_t launch_raw_sizes(const uint8_t *d_payload, size_t c, const uint64_t *d_offsets, uint32_t packets,
                             uint32_t packet, uint32_t *d_sizes, cudaStream_t st);
size_t index_scratch_bytes(size_t c);
cudaError_t launch_index(const uint8_t *d_payload, size_t c, uint64_t *d_offsets, size_t max_packets,
                         uint64_t *d_result, void *d_scratch, size_t scratch_bytes, uint32_t packet, cudaStream_t st);

}  // namespace gpuar
