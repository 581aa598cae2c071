// kernels.h -- host-side launchers of the codec kernels (internal to libgpuar_b200.so)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace gpuar {

// encode.cu
const void *probe_kernel();   // address of a kernel of this library, for image-loadability checks
// Multi-GPU: where a rank's stream lands (encode.cu) -- the segments of the concatenated stream and
// the mailboxes through which the ranks exchange their totals, all local or peer-mapped.
struct ShardPlace {
    uint8_t *segment[16];
    uint64_t *mailbox[16];
    uint64_t seg_cap;
    uint32_t rank, world, n_segments;
};
// `packet` = raw bytes per packet: 8192 in the reference format (gpu.h:13); any multiple of 16 up to
// 16112 is coded correctly (the reference's own limit, compressor.cpp:13).
// where != nullptr (sharded encode): the kernel also adds up the rank's packet sizes and its last
// CTA stores the total into every rank's mailbox; d_acc = shard_acc(...), zeroed by shard_desc_reset.
cudaError_t launch_encode_slots(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                uint32_t *d_sizes, uint32_t packet, cudaStream_t st, const ShardPlace *where = nullptr,
                                uint64_t call = 0, uint64_t *d_acc = nullptr);
// encode_ws.cu: same contract, six specialised warps per 32 packets (for inputs that cannot fill the GPU)
cudaError_t launch_encode_slots_ws(const uint8_t *d_in, size_t n, uint8_t *d_slots, uint32_t slot_stride,
                                   uint32_t *d_sizes, uint32_t packet, cudaStream_t st, const ShardPlace *where = nullptr,
                                   uint64_t call = 0, uint64_t *d_acc = nullptr);
size_t compact_desc_bytes(size_t packets);
bool set_compact_tile(uint32_t packets_per_tile);   // 0 = automatic, else a power of two 4..128
// `cap` != kNoCap: d_payload holds cap bytes and the sizes are untrusted (decode side): packets that
// would end past cap are skipped, *d_total still reports the full sum
constexpr uint64_t kNoCap = ~0ull;
cudaError_t launch_compact(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                           uint32_t packets, uint8_t *d_payload, uint64_t *d_desc, uint64_t *d_total,
                           cudaStream_t st, uint64_t cap = kNoCap);

// Multi-GPU: the rank's packets go straight from the slots to their final place in the stream
// concatenated over all ranks (segments, local or peer-mapped); the ranks' totals arrive through the
// mailboxes (published by the encode kernel).  `call` = per-context counter of sharded encode calls,
// equal on every rank.  d_layout = u64[5]: bytes of all ranks, segment size, this rank's base offset,
// this rank's bytes, status.  Order on the stream: shard_desc_reset, encode (with shard_acc), compact.
uint64_t *shard_acc(uint64_t *d_desc, uint32_t packets);
cudaError_t shard_desc_reset(uint64_t *d_desc, uint32_t packets, cudaStream_t st);
cudaError_t launch_compact_sharded(const uint8_t *d_slots, uint32_t slot_stride, const uint32_t *d_sizes,
                                   uint32_t packets, uint64_t *d_desc, uint64_t *d_layout, const ShardPlace &where,
                                   uint64_t call, cudaStream_t st);

// decode.cu: d_offsets == nullptr means packet p starts at p * stride (reference slot layout);
// d_count != nullptr: the packet count is read on the device (min(*d_count, packets); `packets` sizes the grid)
cudaError_t launch_decode(const uint8_t *d_payload, size_t readable, const uint64_t *d_offsets, uint32_t stride,
                          uint32_t packets, uint8_t *d_out, uint32_t packet, cudaStream_t st,
                          const uint64_t *d_count = nullptr);

bool set_decode_path(int path);    // 0 auto (by packet count), 1 latency variant, 2 throughput variant

// device self-check of the decoder's closed forms (decode.cu): *d_mismatches = 0 when they all hold
cudaError_t launch_selfcheck(uint64_t *d_mismatches, cudaStream_t st);

// index.cu
// sizes[p] = min(rawLen of the packet at d_offsets[p], packet): what decode writes for packet p
cudaError_t launch_raw_sizes(const uint8_t *d_payload, size_t c, const uint64_t *d_offsets, uint32_t packets,
                             uint32_t packet, uint32_t *d_sizes, cudaStream_t st);
size_t index_scratch_bytes(size_t c);
cudaError_t launch_index(const uint8_t *d_payload, size_t c, uint64_t *d_offsets, size_t max_packets,
                         uint64_t *d_result, void *d_scratch, size_t scratch_bytes, uint32_t packet, cudaStream_t st);

// Sharded decode: packet chain of this rank's segment of a stream of stream_bytes bytes laid out in
// `world` segments of seg_bytes (index.cu).  Offsets are local to the segment; d_result as
// index_finish_kernel documents.  Segments must be readable kShardHalo bytes past seg_cap.
constexpr uint32_t kShardHalo = 8704 + 512;           // GPUAR_SHARD_HALO: a packet plus the decoder's over-read
cudaError_t launch_index_segment(const ShardPlace &where, uint64_t call, uint64_t stream_bytes, uint64_t seg_bytes,
                                 uint64_t *d_offsets, size_t max_packets, uint64_t *d_result, void *d_scratch,
                                 size_t scratch_bytes, cudaStream_t st);

}  // namespace gpuar
