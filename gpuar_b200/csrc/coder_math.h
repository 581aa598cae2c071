// coder_math.h -- the per-symbol arithmetic of the codec, shared verbatim by the CUDA
// kernels (device) and by tests/host_model.cpp (host, g++), so the closed forms can be
// checked against the oracle on a machine without a GPU.  No CUDA headers needed.
//
// Coder arithmetic follows the reference exactly (src/gpuar_kernel.cu:256-388,
// 703-716, 787-836) but in closed form; derivation in DESIGN.md §3 / SURVEY.md App. B.
// State per packet:
//   L  = lower bound                                            (16 bit)
//   V  = ~upper & 0xFFFF   ("inverted upper": it shifts in zeros exactly like L)
//   range = upper - lower + 1 = 65536 - V - L                   (> 2^14 between symbols)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GPUAR_HD __host__ __device__ __forceinline__
#else
#define GPUAR_HD inline
#endif

namespace gpuar {

constexpr uint32_t kPacket = 8192;   // UNCOMPRESSED_PACKET_SIZE, gpu.h:13
constexpr uint32_t kSlot = 8704;     // COMPRESSED_PACKET_SIZE,   gpu.h:12
constexpr uint32_t kHdr = 4;         // PACKET_HEADER_LENGTH,     gpu.h:14

// ---- intrinsics with host fall-backs
GPUAR_HD uint32_t clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
GPUAR_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
GPUAR_HD uint32_t bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// Division by the running total T = 256 + i (gpuar_kernel.cu:273,280) is a division by
// a warp-uniform constant: floor(n / T) = mulhi(n, m) >> sh for every n < 2^30, with
//   sh = ceil(log2 T) - 2,   m = ceil(2^(32+sh) / T) < 2^31.
// (e = m*T - 2^(32+sh) < T, and n*e < 2^30 * 2^(sh+2) = 2^(32+sh), so the floor is exact.)
GPUAR_HD uint32_t magic_for(uint32_t T, uint32_t &sh)
{
    sh = 30u - clz32(T - 1u);                        // ceil(log2 T) - 2 for T >= 256
    const uint64_t two_p = 1ull << (32u + sh);
    return (uint32_t)((two_p + T - 1u) / T);
}
GPUAR_HD uint32_t div_total(uint32_t n, uint32_t m, uint32_t sh) { return mulhi32(n, m) >> sh; }

// One interval-narrowing + renormalisation step: gpuar_kernel.cu:256-288, then the closed
// form of the loops at :321-367 (encoder) / :787-836 (decoder).
//   in : L, V; lo = cum[s], hi = cum[s+1]; (m, sh) for the current total
//   out: L, V renormalised; k = equal MSBs shifted out (0..16); u = underflow shifts (0..15);
//        U1 = upper bound before renormalisation (its top k bits are the output bits)
GPUAR_HD void narrow_renorm(uint32_t &L, uint32_t &V, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh,
                            uint32_t &k, uint32_t &u, uint32_t &U1)
{
    const uint32_t range = 65536u - V - L;
    const uint32_t qa = div_total(hi * range, m, sh);
    const uint32_t qb = div_total(lo * range, m, sh);
    const uint32_t V1 = 65536u - L - qa;             // 0xFFFF - (L + qa - 1)
    const uint32_t L1 = L + qb;
    U1 = V1 ^ 0xFFFFu;
    k = clz32((L1 ^ U1) & 0xFFFFu) - 16u;
    const uint32_t g = (L1 & V1) << k;               // positions with L=1, U=0 after the k shifts
    u = clz32(~g & 0x7FFFu) - 17u;                   // run of them starting at bit 14
    const uint32_t t = k + u;
    L = (L1 << t) & 0x7FFFu;
    V = (V1 << t) & 0x7FFFu;
}

// ---- encoder model pass: exclusive prefix of 8 consecutive u16 counts held as 4 packed
// u16x2 words (c[0] = c0 | c1 << 16, ...).  Sums stay below 2^16 (<= 8448), so the packed
// adds never carry across halves.  e[] receives the lane-local exclusive prefixes in the
// same packing; the caller adds the warp-level base (replicated in both halves).
GPUAR_HD uint32_t prefix8_packed(const uint32_t c[4], uint32_t e[4])
{
    const uint32_t a0 = c[0] * 0x10001u;             // (c0, c0+c1)
    const uint32_t a1 = c[1] * 0x10001u;
    const uint32_t a2 = c[2] * 0x10001u;
    const uint32_t a3 = c[3] * 0x10001u;
    const uint32_t s0 = a0 >> 16;                    // c0+c1
    const uint32_t s1 = s0 + (a1 >> 16);             // c0..c3
    const uint32_t s2 = s1 + (a2 >> 16);             // c0..c5
    e[0] = a0 - c[0];
    e[1] = a1 - c[1] + s0 * 0x10001u;
    e[2] = a2 - c[2] + s1 * 0x10001u;
    e[3] = a3 - c[3] + s2 * 0x10001u;
    return s2 + (a3 >> 16);                          // c0..c7
}

// ---- encoder bit sink: MSB-first stream (gpuar_kernel.cu:128-151), flushed as 32-bit words
struct BitSink {
    uint64_t acc;
    uint32_t nb;       // valid low bits of acc, < 32 between calls
    uint32_t *wp;      // next word
    uint32_t *end;     // one past the last writable whole word

    GPUAR_HD void put(uint32_t val, uint32_t len)    // len <= 32, val < 2^len
    {
        acc = (acc << len) | val;
        nb += len;
        if (nb >= 32u) {
            nb -= 32u;
            if (wp < end) *wp = bswap32((uint32_t)(acc >> nb));
            ++wp;
        }
    }
    GPUAR_HD void put_run(uint32_t bit, uint32_t n)  // n copies of bit
    {
        const uint32_t ones = bit ? 0xFFFFu : 0u;
        while (n > 16u) { put(ones, 16u); n -= 16u; }
        put(ones & ((1u << n) - 1u), n);
    }
};

// Bits of one symbol: the top k bits of U1 with, right after the first of them, `pend`
// inverted copies of it (gpuar_kernel.cu:325-336); then the underflow count carries on.
GPUAR_HD void emit_symbol(BitSink &out, uint32_t &pend, uint32_t k, uint32_t u, uint32_t U1)
{
    if (k) {
        const uint32_t b = U1 >> 15;
        const uint32_t rest = (U1 >> (16u - k)) & ((1u << (k - 1u)) - 1u);
        if (pend <= 16u) {
            const uint32_t head = (1u << pend) - (b ^ 1u);          // b, then pend x !b
            out.put((head << (k - 1u)) | rest, k + pend);
        } else {
            out.put(b, 1u);
            out.put_run(b ^ 1u, pend);
            out.put(rest, k - 1u);
        }
        pend = u;
    } else {
        pend += u;
    }
}

// End of packet: bit 14 of L, then pend+1 inverted copies (gpuar_kernel.cu:379-388); zero
// padding to a byte (:430-439).  Returns the number of bitstream bytes; writes the tail
// bytes and the 4-byte packet header (:525-528) at `slot`.
GPUAR_HD uint32_t finish_packet(BitSink &out, uint32_t L, uint32_t pend, uint8_t *slot, uint32_t raw_len)
{
    uint32_t *const first = reinterpret_cast<uint32_t *>(slot + kHdr);
    const uint32_t b = (L >> 14) & 1u;
    out.put(b, 1u);
    out.put_run(b ^ 1u, pend + 1u);
    uint32_t bytes = (uint32_t)(out.wp - first) * 4u;
    if (out.nb) {
        const uint32_t tail = (out.nb + 7u) >> 3;
        const uint32_t w = (uint32_t)(out.acc << (32u - out.nb));  // left-aligned, zero padded
        uint8_t *bp = reinterpret_cast<uint8_t *>(out.wp);
        for (uint32_t t = 0; t < tail; ++t)
            if (bp + t < reinterpret_cast<uint8_t *>(out.end)) bp[t] = (uint8_t)(w >> (24u - 8u * t));
        bytes += tail;
    }
    const uint32_t comp = bytes + kHdr;
    *reinterpret_cast<uint32_t *>(slot) = (comp & 0xFFFFu) | (raw_len << 16);   // u16 compLen | u16 rawLen, LE
    return comp;
}

// ---- decoder: target = ((code - L + 1) * T - 1) / range  (getUnscaledCode, :703-716)
// num < 2^30 and 2^14 < range <= 2^16: a float estimate is within 1 of the quotient, one
// correction step makes it exact.
GPUAR_HD uint32_t unscale(uint32_t code, uint32_t L, uint32_t V, uint32_t T)
{
    const uint32_t range = 65536u - V - L;
    const uint32_t num = (((code - L) & 0xFFFFu) + 1u) * T - 1u;
#if defined(__CUDA_ARCH__)
    uint32_t q = (uint32_t)__fmul_rz(__uint2float_rn(num), __frcp_rn(__uint2float_rn(range)));
#else
    uint32_t q = (uint32_t)((double)(float)num * (double)(1.0f / (float)range));
#endif
    const int32_t r = (int32_t)(num - q * range);
    if (r < 0) --q;
    else if (r >= (int32_t)range) ++q;
    return q;
}

// ---- decoder model: 4-ary cumulative-count tree, 85 nodes of three u16 thresholds
//   t0 = |child0|, t1 = t0 + |child1|, t2 = t1 + |child2|; node n lives at base[n * stride]
struct TreeNode { uint32_t x, y; };                   // x = t0 | t1 << 16, y = t2
constexpr uint32_t kTreeNodes = 1 + 4 + 16 + 64;

GPUAR_HD void tree_init(TreeNode *base, uint32_t stride)
{
    uint32_t node = 0;
    for (uint32_t lvl = 0, span = 64; lvl < 4; ++lvl, span >>= 2)
        for (uint32_t q = 0; q < (1u << (2u * lvl)); ++q, ++node) {
            base[node * stride].x = span | ((2u * span) << 16);
            base[node * stride].y = 3u * span;
        }
}

// Finds the symbol whose cumulative interval holds `target` (getSymbolFromProbability,
// :727-763), returns it with lo = cum[s], cnt = count[s], and bumps count[s] (:288).
GPUAR_HD uint32_t tree_decode(TreeNode *base, uint32_t stride, uint32_t target, uint32_t total, uint32_t &lo,
                              uint32_t &cnt)
{
    uint32_t rem = target, tot = total, idx = 0, acc = 0, first = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t lvl = 0; lvl < 4; ++lvl) {
        TreeNode *n = base + (first + idx) * stride;
        TreeNode t = *n;
        const uint32_t t0 = t.x & 0xFFFFu, t1 = t.x >> 16, t2 = t.y;
        const uint32_t c = (uint32_t)(rem >= t0) + (uint32_t)(rem >= t1) + (uint32_t)(rem >= t2);
        const uint32_t below = c == 0 ? 0u : c == 1 ? t0 : c == 2 ? t1 : t2;
        const uint32_t above = c == 0 ? t0 : c == 1 ? t1 : c == 2 ? t2 : tot;
        t.x += c == 0 ? 0x00010001u : c == 1 ? 0x00010000u : 0u;
        t.y += c <= 2 ? 1u : 0u;
        *n = t;
        rem -= below;
        acc += below;
        tot = above - below;
        idx = idx * 4u + c;
        first += 1u << (2u * lvl);
    }
    lo = acc;
    cnt = tot;
    return idx;
}

// ---- decoder bit source: 64-bit reservoir, next bit = MSB, fed one 32-bit word at a time
struct BitSource {
    uint64_t buf;
    uint32_t have;     // valid bits in buf; >= 33 at the top of every step
    GPUAR_HD uint32_t take(uint32_t t)               // t <= 31
    {
        const uint32_t bits = (uint32_t)((buf >> 1) >> (63u - t));
        buf <<= t;
        have -= t;
        return bits;
    }
    GPUAR_HD bool hungry() const { return have <= 32u; }
    GPUAR_HD void feed(uint32_t be_word)
    {
        buf |= (uint64_t)be_word << (32u - have);
        have += 32u;
    }
};

// readEncodedBits in closed form (:787-836): shift in k+u bits; an underflow run leaves
// the MSB flipped.
GPUAR_HD uint32_t advance_code(uint32_t code, uint32_t k, uint32_t u, BitSource &in)
{
    const uint32_t t = k + u;
    return (((code << t) | in.take(t)) & 0xFFFFu) ^ (u ? 0x8000u : 0u);
}

}  // namespace gpuar
