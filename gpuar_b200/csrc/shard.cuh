// shard.cuh -- what the GPUs of one sharded stream share: the mailbox words and how they are read
// and written (peer memory, system scope).
#pragma once
#include "common.cuh"

namespace gpuar {

// Mailbox of a rank: GPUAR_MAILBOX_BYTES of device memory every rank can write.  Words (u64):
//   [ 0..31]  kMailTotals: per-rank payload totals of a sharded encode, [parity][rank]
//   [32..47]  kMailChain : hand-over of the packet chain of a sharded decode, [parity][4]:
//             global offset of the first packet that starts in this rank's segment | packets
//             before it | raw bytes before it | status of the ranks before
// Every word carries the tag of the call in bits 44..63 (0 = never written) and a 44-bit value.
constexpr uint32_t kMaxRanks = 16;                    // GPUAR_MAX_RANKS
constexpr uint32_t kMailTotals = 0;                   // u64[2][16]: per-rank totals, by parity
constexpr uint32_t kMailChain = 32;                   // u64[2][4]: chain hand-over, by parity
constexpr uint64_t kMinSegment = 16384;               // GPUAR_SHARD_MIN_SEGMENT: more than a packet with its halo
constexpr uint64_t kTagShift = 44;
constexpr uint64_t kValueMask = (1ull << kTagShift) - 1ull;

struct ShardTarget {
    uint8_t *segment[kMaxRanks];
    uint64_t *mailbox[kMaxRanks];
    uint64_t seg_cap;
    uint32_t rank, world, n_segments;
    uint32_t tag, parity;
};

__device__ __forceinline__ uint64_t ld_sys(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr uint64_t kMailTimeoutNs = 4000000000ull;    // a peer that never publishes must not hang the GPU


__device__ __forceinline__ uint32_t call_tag(uint64_t call) { return (uint32_t)(call % 0xFFFFFull) + 1u; }

// waits for a word of the own mailbox to carry `tag`; false on a timeout
__device__ __forceinline__ bool mail_wait(const uint64_t *p, uint32_t tag, uint64_t &value)
{
    const uint64_t t0 = global_ns();
    for (;;) {
        const uint64_t v = ld_sys(p);
        if ((uint32_t)(v >> kTagShift) == tag) { value = v & kValueMask; return true; }
        if (global_ns() - t0 > kMailTimeoutNs) { value = 0; return false; }
        __nanosleep(64);
    }
}
__device__ __forceinline__ void mail_post(uint64_t *p, uint32_t tag, uint64_t value)
{
    st_sys(p, ((uint64_t)tag << kTagShift) | (value & kValueMask));
}


// A kernel's CTAs add up this rank's payload bytes and the last one to finish stores the total into
// every rank's mailbox.  Called by one thread per CTA with the CTA's sum; `acc` = u64[2] (sum, CTAs
// done), zeroed before the launch.
__device__ __forceinline__ void shard_publish_total(uint64_t cta_sum, uint64_t *acc, const ShardTarget &tg)
{
    atomicAdd(reinterpret_cast<unsigned long long *>(&acc[0]), (unsigned long long)cta_sum);
    __threadfence();
    if (atomicAdd(reinterpret_cast<unsigned long long *>(&acc[1]), 1ull) + 1ull != gridDim.x) return;
    __threadfence();
    uint64_t total;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(total) : "l"(&acc[0]) : "memory");
    const uint64_t word = ((uint64_t)tg.tag << kTagShift) | (total & kValueMask);
    for (uint32_t r = 0; r < tg.world; ++r) st_sys(tg.mailbox[r] + kMailTotals + tg.parity * kMaxRanks + tg.rank, word);
    __threadfence_system();
}

}  // namespace gpuar
