/*
 * gpuar_b200.h -- C ABI of the B200 (sm_100a) GPUAR codec library, libgpuar_b200.so.
 *
 * This is the drop-in boundary for the reference's device codec seam:
 *   - the `extern "C"` block of the reference,           src/gpuar.h:59-86
 *   - its constants,                                      src/gpu.h:8-14
 *   - and the device work the reference's host driver does around it
 *     (host-side compaction of fixed-stride slots, gpu_compressor.cpp:136-169, and
 *     the host-side packet-chain walk, gpu_compressor.cpp:294-320), which here
 *     run on the device.
 *
 * Plain pointers and sizes only.  Every `d_` pointer is device memory on the
 * CURRENT CUDA device (cudaSetDevice is the caller's business, as in the
 * reference); `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  All device entry points are asynchronous with respect to the host
 * unless stated otherwise; results written to device memory are valid after
 * the stream has been synchronised.  Return value: 0 = ok, > 0 = a cudaError_t,
 * < 0 = one of the GPUAR_E_* codes below.  There is no CPU fallback anywhere in
 * this library: with no usable device every entry point fails.
 *
 * Wire format (identical to the reference):
 *   .gip    = 20-byte header | payload
 *   payload = packets tightly concatenated in input order
 *   packet  = u16 LE compLen (incl. these 4 bytes) | u16 LE rawLen | bitstream
 *   rawLen  = 8192 for every packet except possibly the last.
 */
#ifndef GPUAR_B200_H
#define GPUAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* format constants -- src/gpu.h:8-14, src/file_header.hpp:19-22 */
#define GPUAR_PACKET_BYTES 8192u  /* UNCOMPRESSED_PACKET_SIZE */
#define GPUAR_SLOT_BYTES 8704u    /* COMPRESSED_PACKET_SIZE   */
#define GPUAR_PACKET_HEADER 4u    /* PACKET_HEADER_LENGTH     */
#define GPUAR_FILE_HEADER 20u     /* FileHeader::HEADER_LENGTH */
#define GPUAR_PAD_BYTES 64u       /* readable slack required past a payload (decoder over-read) */
#define GPUAR_HEADER_WIDE_MARK 0xB2u /* header byte 3 (never written by the reference, file_header.hpp:28-36):
                                        set by this library when bytes 8-11 / 16-19 carry the high halves of
                                        64-bit sizes */

#define GPUAR_E_ARG (-1)          /* bad argument / buffer too small            */
#define GPUAR_E_FORMAT (-2)       /* malformed .gip header or packet chain      */
#define GPUAR_E_NODEVICE (-3)     /* no CUDA device / kernel image not loadable */
#define GPUAR_E_UNSUPPORTED (-4)  /* valid but outside what the device path handles */

int gpuar_b200_abi_version(void);
const char *gpuar_b200_strerror(int code);

/* Once per device before any other call (replaces initConstantRange, gpuar.h:74):
 * uploads the per-position reciprocal table.  Synchronous. */
int gpuar_b200_init(void);

/* ---------------------------------------------------------------- sizing */
size_t gpuar_b200_packets(size_t n);              /* ceil(n / 8192)                              */
size_t gpuar_b200_payload_bound(size_t n);        /* packets * 8704 + GPUAR_PAD_BYTES            */
size_t gpuar_b200_encode_scratch_bytes(size_t n); /* device scratch for gpuar_b200_encode        */
size_t gpuar_b200_index_scratch_bytes(size_t c);  /* device scratch for gpuar_b200_index         */

/* ---------------------------------------------------------------- encode
 * d_in[n] -> d_payload (packets compacted on the device, no 20-byte header).
 * Replaces garCompressExecutor (gpuar.h:77) + the per-packet D2H/fwrite loop
 * (gpu_compressor.cpp:136-169).
 *   d_in            16-byte aligned and READABLE up to n rounded up to 16: the kernels fetch the
 *                   input as 16-byte words, so the last fetch may read up to 15 bytes past n (their
 *                   values are ignored).  The same holds for garCompressExecutor below, as it does
 *                   for the reference's own kernel (gpuar_kernel.cu:496-517 reads ulonglong2 words).
 *   d_payload       16-byte aligned, capacity >= gpuar_b200_payload_bound(n).
 *   d_payload_bytes device u64: total payload length C.
 *   d_packet_sizes  optional device u32[packets]: compLen of each packet (may be NULL).
 */
int gpuar_b200_encode(const uint8_t *d_in, size_t n, uint8_t *d_payload, size_t payload_cap,
                      uint64_t *d_payload_bytes, uint32_t *d_packet_sizes, void *d_scratch,
                      size_t scratch_bytes, void *stream);

/* ---------------------------------------------------------------- index
 * Packet-chain discovery on the device (the format stores no index).  Replaces
 * the host chain walk of gpu_compressor.cpp:294-320.
 *   d_payload     16-byte aligned, c payload bytes, readable to c + GPUAR_PAD_BYTES.
 *   d_offsets     device u64[max_packets]: byte offset of each packet in d_payload.
 *   d_result      device u64[4]: [0] packet count, [1] total rawLen, [2] status
 *                 (0 ok, else a GPUAR_E_* code as two's complement), [3] 1 if a packet other
 *                 than the last is short (rawLen < 8192; never written by the reference, but
 *                 legal for its CPU decoder): decode such a stream with gpuar_b200_decode_packed.
 * max_packets >= c/5 + 1 is always enough; >= ceil(raw/8192) suffices for well-formed input.
 */
int gpuar_b200_index(const uint8_t *d_payload, size_t c, uint64_t *d_offsets, size_t max_packets,
                     uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream);

/* ---------------------------------------------------------------- decode
 * Packets at d_offsets[0..n_packets) -> d_out; packet p is written at p * 8192
 * (gpuar_kernel.cu:924), rawLen bytes of it.  Replaces garDecompressExecutor (gpuar.h:78).
 *   d_out  16-byte aligned, capacity >= n_packets * 8192.
 */
int gpuar_b200_decode(const uint8_t *d_payload, size_t c, const uint64_t *d_offsets, size_t n_packets,
                      uint8_t *d_out, size_t out_cap, void *stream);

/* ----------------------------------------------------- decode, packed output
 * The same packets written back to back: packet p lands at the sum of the rawLen fields of the
 * packets before it.  This is the layout of the reference's CPU decoder (cpu_compressor.cpp:60-70:
 * fwrite of rawLen bytes per packet), which -- unlike its GPU decoder -- accepts short packets
 * anywhere in the stream; gpuar_b200_index flags such streams in d_result[3].  For streams whose
 * packets are all full except the last, it equals gpuar_b200_decode (which is one kernel cheaper).
 *   d_out        16-byte aligned, out_cap bytes; packets that would end past out_cap are not written.
 *   d_out_bytes  device u64: total raw bytes of the n_packets packets (compare with out_cap).
 *   d_scratch    16-byte aligned, gpuar_b200_decode_packed_scratch_bytes(n_packets, packet_bytes).
 */
size_t gpuar_b200_decode_packed_scratch_bytes(size_t n_packets, size_t packet_bytes);
int gpuar_b200_decode_packed(const uint8_t *d_payload, size_t c, size_t packet_bytes, const uint64_t *d_offsets,
                             size_t n_packets, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bytes, void *d_scratch,
                             size_t scratch_bytes, void *stream);

/* ------------------------------------------------ other packet sizes (config sweep)
 * The packet size is a compile-time constant of the reference (gpu.h:12-13) and is not recorded
 * in the .gip; a reference rebuilt with another UNCOMPRESSED_PACKET_SIZE produces and reads a
 * different dialect.  The _ex entry points take it as a run-time argument: any multiple of 16
 * from 16 to 16112 (the reference's own limit 2^14 - 256, compressor.cpp:13); slots are
 * packet_bytes + 512.  packet_bytes = 8192 is exactly the plain entry points. */
size_t gpuar_b200_payload_bound_ex(size_t n, size_t packet_bytes);
size_t gpuar_b200_encode_scratch_bytes_ex(size_t n, size_t packet_bytes);
int gpuar_b200_encode_ex(const uint8_t *d_in, size_t n, size_t packet_bytes, uint8_t *d_payload, size_t payload_cap,
                         uint64_t *d_payload_bytes, uint32_t *d_packet_sizes, void *d_scratch, size_t scratch_bytes,
                         void *stream);
int gpuar_b200_index_ex(const uint8_t *d_payload, size_t c, size_t packet_bytes, uint64_t *d_offsets,
                        size_t max_packets, uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream);
int gpuar_b200_decode_ex(const uint8_t *d_payload, size_t c, size_t packet_bytes, const uint64_t *d_offsets,
                         size_t n_packets, uint8_t *d_out, size_t out_cap, void *stream);

/* ------------------------------------------------- host-buffer entry points
 * Whole .gip image in host memory <-> raw bytes in host memory on the current
 * device: staging, H2D, kernels, D2H, header, pipelined in chunks over several
 * streams.  Synchronous.  What a caller of the reference's
 * GPUCompressor::compress/decompress gets, minus the file I/O.  decompress_host
 * hops over the compLen fields on the host (the image is in host memory anyway,
 * cf. gpu_compressor.cpp:294-320) only to cut it into chunks of whole packets.
 * `gip_cap` >= 20 + gpuar_b200_payload_bound(n).
 */
int gpuar_b200_compress_host(const uint8_t *in, size_t n, uint8_t *gip, size_t gip_cap, size_t *gip_bytes);
int gpuar_b200_decompress_host(const uint8_t *gip, size_t gip_bytes, uint8_t *out, size_t out_cap,
                               size_t *out_bytes);
/* The same over several GPUs of the box, from ONE host thread of one process: the chunks rotate
 * over the devices (chunk k on devices[k mod n_devices]), every device's PCIe link carries its own
 * chunks up and down, and the image is assembled in place (a chunk's payload is copied to its final
 * position as soon as the totals of the chunks before it are known).  devices = NULL means the
 * current device alone (exactly the plain entry points).  Host buffers should be page-locked
 * (gpuar_b200_host_alloc) for the copies to run at link speed.  Calls that share a device are
 * serialised per device; calls on disjoint devices run concurrently. */
int gpuar_b200_compress_host_multi(const int *devices, int n_devices, const uint8_t *in, size_t n, uint8_t *gip,
                                   size_t gip_cap, size_t *gip_bytes);
int gpuar_b200_decompress_host_multi(const int *devices, int n_devices, const uint8_t *gip, size_t gip_bytes,
                                     uint8_t *out, size_t out_cap, size_t *out_bytes);
/* Measurement hook: the transfers of gpuar_b200_compress_host_multi without its kernels -- the same
 * chunks of in[n] go up on the same streams and out_bytes come down, spread over the chunks.
 * Synchronous; the caller times it.  What the PCIe links and the host memory behind them can do
 * for this access pattern: the ceiling of the end-to-end numbers. */
int gpuar_b200_host_link_probe(const int *devices, int n_devices, const uint8_t *in, size_t n, uint8_t *out,
                               size_t out_bytes);
/* raw size announced by a .gip image (walks nothing: header field, 32-bit in the
 * reference's layout, 64-bit when written by this library -- see DESIGN.md). */
int gpuar_b200_gip_raw_size(const uint8_t *gip, size_t gip_bytes, uint64_t *raw_bytes);

/* The same by hopping over the compLen / rawLen fields of the whole image in host memory, as the
 * reference's CPU decoder finds its packets (cpu_compressor.cpp:47-78): exact for any stream,
 * including reference-written images of 4 GiB and more whose 32-bit header field has wrapped.
 * One u32 read per packet; validates the chain (GPUAR_E_FORMAT / GPUAR_E_UNSUPPORTED). */
int gpuar_b200_gip_walk(const uint8_t *gip, size_t gip_bytes, uint64_t *packets, uint64_t *raw_bytes);

/* 20-byte header, byte-compatible with file_header.hpp:28-36,61-72 in every byte the
 * reference defines; the bytes it leaves uninitialised carry the high halves of the
 * 64-bit sizes (zero below 4 GiB) and, in byte 3, GPUAR_HEADER_WIDE_MARK, without which a
 * reader must not trust them. */
void gpuar_b200_write_header(uint8_t hdr[20], uint64_t raw_bytes, uint64_t gip_bytes);
int gpuar_b200_check_header(const uint8_t hdr[20]);

/* --------------------------------------------- multi-GPU stream concatenation
 * Peer copy of one rank's compacted payload into the gathered payload that lives
 * on another device of the same box (NVLink): dst_device/dst may be a
 * cudaIpcOpenMemHandle mapping.  Asynchronous on `stream`. */
int gpuar_b200_peer_concat(uint8_t *d_dst, int dst_device, size_t dst_offset, const uint8_t *d_src,
                           int src_device, size_t bytes, void *stream);
/* ---------------------------------------------------- sharded encode (one input, W GPUs)
 * Packets are independent (gpuar_kernel.cu:901-907), so rank r of W encodes a contiguous packet
 * range of the input and the .gip payload is the concatenation of the ranks' streams in rank
 * order.  gpuar_b200_encode_sharded does that without a host round trip, without NCCL and without
 * a second pass over the payload: after its encode kernel a rank sums its packet sizes and stores
 * the total into every rank's MAILBOX (peer stores over NVLink); its size scan + compaction kernel
 * waits for the W totals of this call in its own mailbox, takes their exclusive scan as its
 * landing offset, and writes every packet straight to its final place in the concatenated stream.
 *
 * The stream is laid out in n_segments equal segments of S = ceil(total / n_segments) bytes
 * (rounded up to 256, at least GPUAR_SHARD_MIN_SEGMENT): global offset o is byte o % S of segments[o / S].  n_segments = world with
 * segment g on GPU g keeps every GPU's ingress at S bytes (with balanced shards almost nothing
 * crosses NVLink); n_segments = 1 gathers the whole stream into one buffer.
 *
 * gpuar_b200_shard describes the group from ONE rank's point of view; every pointer in it is a
 * device pointer valid on that rank's device: its own allocation, or a peer's mapped with
 * gpuar_b200_ipc_open (one process per GPU) or made accessible with gpuar_b200_enable_peer (one
 * process, several GPUs).  Mailboxes are GPUAR_MAILBOX_BYTES each, zeroed once before the first
 * call.  `calls` is the library's per-context call counter: zero it once, never touch it again;
 * all ranks must make the same sequence of sharded calls (they are collective: a rank that does
 * not call leaves the others waiting -- for at most a few seconds, then status 2). */
#define GPUAR_MAX_RANKS 16
#define GPUAR_MAILBOX_BYTES 1024u
typedef struct gpuar_b200_shard {
    int32_t rank, world;                     /* this rank; number of ranks, <= GPUAR_MAX_RANKS          */
    int32_t n_segments;                      /* world, or 1                                             */
    int32_t reserved;
    uint64_t seg_cap;                        /* capacity of every segment, bytes                        */
    uint8_t *segments[GPUAR_MAX_RANKS];      /* [n_segments]                                            */
    uint64_t *mailbox[GPUAR_MAX_RANKS];      /* [world]; mailbox[rank] is this rank's own               */
    uint64_t calls[4];                       /* library state: zero once                                */
} gpuar_b200_shard;

/* d_in[n] = this rank's packet range (n a multiple of 8192 on every rank but the last).
 *   d_layout   device u64[8]: [0] bytes of the whole concatenated stream, [1] segment size S,
 *              [2] this rank's landing offset, [3] this rank's bytes, [4] status: 0 ok, 1 a segment
 *              is smaller than S (nothing was written), 2 a peer's total did not arrive;
 *              [5] diagnostics: nanoseconds the first compaction tile waited for the ranks' totals.
 *   d_scratch  gpuar_b200_encode_scratch_bytes(n); the rank's own payload buffer is not needed. */
int gpuar_b200_encode_sharded(gpuar_b200_shard *shard, const uint8_t *d_in, size_t n, uint64_t *d_layout,
                              uint32_t *d_packet_sizes, void *d_scratch, size_t scratch_bytes, void *stream);

/* ---------------------------------------------------- sharded decode (one stream, W GPUs)
 * The stream of stream_bytes bytes lies in world segments of S = gpuar_b200_shard_segment_bytes
 * bytes, segment g on GPU g -- the layout gpuar_b200_encode_sharded writes with n_segments = world
 * (or a .gip payload cut into equal pieces by any other means).  Every segment must be allocated
 * GPUAR_SHARD_HALO bytes larger than seg_cap.  The stream must be complete and visible on every
 * GPU before the call (synchronise the ranks after the encode).
 * Every rank discovers the packet chain of its own segment in parallel (a segment starts in the
 * middle of a packet: where its first packet begins, and how many packets precede it, arrives
 * from the rank before through the mailbox -- the only serial step, a few microseconds per rank)
 * and decodes the packets that START in its segment; the tail of the last one is read from the
 * head of the next segment, copied behind the rank's own first.  The result stays sharded: the
 * rank's k-th packet is written at d_out + k * 8192.
 *   d_result  device u64[8]: [0] packets decoded by this rank, [1] their raw bytes, [2] status
 *             (0, or a GPUAR_E_* code as two's complement; GPUAR_E_ARG: out_cap too small),
 *             [3] packets before this rank's first = index of its first packet in the stream,
 *             [4] raw bytes before it = where d_out belongs in the decoded file.
 *   d_scratch gpuar_b200_decode_sharded_scratch_bytes(S, out_cap / 8192). */
#define GPUAR_SHARD_HALO (8704u + 512u)
#define GPUAR_SHARD_MIN_SEGMENT 16384u   /* S never gets smaller: a packet spans at most two segments */
uint64_t gpuar_b200_shard_segment_bytes(uint64_t stream_bytes, int n_segments);
size_t gpuar_b200_decode_sharded_scratch_bytes(uint64_t seg_bytes, size_t max_packets);
int gpuar_b200_decode_sharded(gpuar_b200_shard *shard, uint64_t stream_bytes, uint8_t *d_out, size_t out_cap,
                              uint64_t *d_result, void *d_scratch, size_t scratch_bytes, void *stream);
/* one process, several GPUs: lets kernels on the current device reach memory of `peer_device` */
int gpuar_b200_enable_peer(int peer_device);
/* plain cudaMalloc/cudaFree on the current device: IPC-exportable allocations for the gather buffer */
int gpuar_b200_device_alloc(size_t bytes, void **d_ptr);
int gpuar_b200_device_free(void *d_ptr);
/* page-locked host staging buffers (replaces the cudaMallocHost calls of compressor.cpp:23-25 and
 * gpu_compressor.cpp:62-63) and device selection (gpu_compressor.cpp:67-82), so that host code
 * above this ABI needs no CUDA headers */
int gpuar_b200_host_alloc(size_t bytes, void **h_ptr);
int gpuar_b200_host_free(void *h_ptr);
int gpuar_b200_device_count(void);
int gpuar_b200_set_device(int device);
int gpuar_b200_ipc_export(const void *d_ptr, uint8_t handle[64]);
int gpuar_b200_ipc_open(const uint8_t handle[64], void **d_ptr);
int gpuar_b200_ipc_close(void *d_ptr);

/* --------------------------------------------------- reference-named shims
 * Same names, signatures and slot layout as src/gpuar.h:74,77-78, so the
 * reference's gpu_compressor.cpp links against this library unchanged:
 * packet t is read at d_src + t*8192 and written to d_dst + t*8704 (encode),
 * or read at d_src + t*8704 and written to d_dst + t*8192 (decode); launched on
 * the legacy default stream; numBlocks is accepted and ignored (the grid is
 * derived from size). */
void initConstantRange(void);
void garCompressExecutor(const uint8_t *source, size_t size, uint8_t *destination, uint32_t numBlocks);
void garDecompressExecutor(const uint8_t *source, size_t size, uint8_t *destination, uint32_t numBlocks);

/* ------------------------------------------------------------------- options
 * The encoder has two kernels with identical output: lane = packet with the three stages of a
 * step software-pipelined in one warp (best when packets >> warp schedulers), and a
 * warp-specialised one (model / coder / bit-packing warps per 32 packets; best for small
 * inputs).  Default: chosen by packet count. */
#define GPUAR_OPT_ENCODE_PATH 1      /* 0 auto (default), 1 fused, 2 warp-specialised */
#define GPUAR_OPT_WS_MAX_PACKETS 2   /* auto: use the warp-specialised kernel up to this many packets */
#define GPUAR_OPT_COMPACT_TILE 3     /* packets per work unit of the scan + compaction kernel: 0 auto (default),
                                        or a power of two 4..128 (32 KiB .. 1 MiB of input per CTA) */
#define GPUAR_OPT_DECODE_PATH 4      /* decode kernel: 0 auto (default: by packet count), 1 latency variant
                                        (speculative node loads, for inputs that leave warp schedulers idle),
                                        2 throughput variant */
int gpuar_b200_set_option(int key, long long value);

/* ----------------------------------------------------------- measurement hooks
 * gpuar_b200_profile(1) makes encode/index/decode record CUDA events on the caller's
 * stream around each kernel group; gpuar_b200_profile_read waits for the recorded spans,
 * returns accumulated milliseconds and span counts per group, and clears them. */
#define GPUAR_SPAN_ENCODE 0   /* model + coder kernel                     */
#define GPUAR_SPAN_COMPACT 1  /* size scan + stream compaction kernel     */
#define GPUAR_SPAN_INDEX 2    /* packet-chain discovery kernels           */
#define GPUAR_SPAN_DECODE 3   /* decode kernel                            */
#define GPUAR_SPAN_COUNT 4
void gpuar_b200_profile(int enable);
int gpuar_b200_profile_read(double ms[GPUAR_SPAN_COUNT], uint64_t calls[GPUAR_SPAN_COUNT]);

/* Device self-check of the decoder's float-estimated quotient (the decoder's replacement for the division
 * of getUnscaledCode, src/gpuar_kernel.cu:703-716): every range the coder can hold, every quotient, both
 * ends of each quotient's interval, on the current device (the approximate reciprocal it is built on only
 * exists there).  *mismatches = 0 when it holds everywhere.  A test hook; ~20 ms. */
int gpuar_b200_selfcheck(uint64_t *mismatches);

/* launch counter: number of this library's kernels launched since load (bench.py's gpu_launches) */
uint64_t gpuar_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUAR_B200_H */
