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
#include <string.h>

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
//   sh = floor(log2 T) - 1,   m = ceil(2^(32+sh) / T) <= 2^31.
// (e = m*T - 2^(32+sh) < T <= 2^(sh+2), and n*e < 2^30 * 2^(sh+2) = 2^(32+sh), so the floor
// is exact; for T a power of two e = 0.)  sh only changes where T crosses a power of two,
// i.e. at i = 256, 768, 1792, 3840, 7936 -- all multiples of 32, so it is constant over
// any aligned run of 32 positions.
GPUAR_HD uint32_t shift_for(uint32_t T) { return 30u - clz32(T); }
GPUAR_HD uint32_t magic_for(uint32_t T, uint32_t &sh)
{
    sh = shift_for(T);
    const uint64_t two_p = 1ull << (32u + sh);
    return (uint32_t)((two_p + T - 1u) / T);
}
GPUAR_HD uint32_t div_total(uint32_t n, uint32_t m, uint32_t sh) { return mulhi32(n, m) >> sh; }

// One interval-narrowing + renormalisation step: gpuar_kernel.cu:256-288, then the closed
// form of the loops at :321-367 (encoder) / :787-836 (decoder).
//   in : L, V; lo = cum[s], hi = cum[s+1]; (m, sh) for the current total
//   out: L, V renormalised; k = equal MSBs shifted out (0..16); u = underflow shifts (0..15);
//        U1 = upper bound before renormalisation (its top k bits are the output bits)
// Reference loop in closed form: always k matching-MSB shifts, then u underflow shifts, and
// L = (L1 << (k+u)) & 0x7FFF, V = (V1 << (k+u)) & 0x7FFF.
GPUAR_HD void shifts_of(uint32_t L1, uint32_t U1, uint32_t &k, uint32_t &u)
{
    k = clz32((L1 ^ U1) & 0xFFFFu) - 16u;
    const uint32_t g = (L1 & ~U1) << k;              // positions with L=1, U=0 after the k shifts
    u = clz32(~g & 0x7FFFu) - 17u;                   // run of them starting at bit 14
}

GPUAR_HD void narrow_renorm(uint32_t &L, uint32_t &V, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh,
                            uint32_t &k, uint32_t &u, uint32_t &U1)
{
    const uint32_t range = 65536u - V - L;
    const uint32_t qa = div_total(hi * range, m, sh);
    const uint32_t qb = div_total(lo * range, m, sh);
    const uint32_t V1 = 65536u - L - qa;             // 0xFFFF - (L + qa - 1)
    const uint32_t L1 = L + qb;
    U1 = V1 ^ 0xFFFFu;
    shifts_of(L1, U1, k, u);
    const uint32_t t = k + u;
    L = (L1 << t) & 0x7FFFu;
    V = (V1 << t) & 0x7FFFu;
}

// ---- the same step with a single normalisation on the dependent chain.
// State here is (L, R) with R = range = upper - lower + 1 carried explicitly.  The recurrence
// only needs the TOTAL shift t = k + u, and t follows from the WIDTH W = U1 - L1 + 1 of the
// narrowed interval up to one position: the renormalised width W * 2^t lies in (2^14, 2^16],
// so with e = floor(log2(W - 1)) (e = -1 for W = 1) t is 14 - e or 15 - e.  Every shift
// doubles the width exactly, so the new range is W << t and needs neither bound.
//
// Which of the two: shift by s1 = 15 - e.  A = L1 << s1 is the lower bound in a 17-bit frame
// (bits above 16 are the bits already gone), R1 = W << s1 lies in (2^15, 2^16], and the
// upper bound is A + R1 - 1.  R1 - 1 lies in [2^15, 2^16), so going from A to the upper bound
// advances the two-bit number formed by frame bits 16 and 15 by one, plus the carry out of
// the low 15 bits.  Advancing by one always leaves top bits that ask for the last shift --
// (0,1) and (2,3) have equal MSBs, (1,2) and (3,0) show the underflow pattern L = 1, U = 0 in
// the second bits -- and advancing by two never does: (0,2), (1,3), (2,0), (3,1) have
// different MSBs and equal second bits.  So "one shift less" IS that carry:
//     sx = ((A & 0x7FFF) + R1 - 0x8001) >> 15;   t = s1 - sx;   R = R1 >> sx;
//     L = (A >> sx) & 0x7FFF;   and bit 15 of A >> sx tells whether u != 0.
//
// e comes from the exponent of a float, built with one integer add and one float subtract,
// so no count-leading-zeros or conversion instruction -- both variable-latency on the XU
// pipe -- sits on the chain: 0x4BFFFFFF + W are the bits of 2^25 + 4(W - 1) (exact, ulp 4),
// and subtracting 2^25 - 2 leaves 4W - 2 = 4 (W - 0.5) exactly, whose biased exponent is
// E = 129 + e.  E is congruent to e + 1 = 16 - s1 modulo 32, so "<< s1" is a funnel shift
// RIGHT of the operand pre-shifted by 16 with E itself as the (wrapping) shift amount.
// k and u individually matter only for the emitted bits; the encoder derives them off the
// chain (shifts_of).
GPUAR_HD uint32_t width_exponent(uint32_t qa, uint32_t qb)    // 129 + floor(log2(qa - qb - 1)); 128 for qa - qb = 1
{
    float f;
#if defined(__CUDA_ARCH__)
    uint32_t b;                                               // one three-input add
    asm("{\n\t.reg .u32 t;\n\tsub.u32 t, %1, %2;\n\tadd.u32 %0, t, 0x4BFFFFFF;\n\t}" : "=r"(b) : "r"(qa), "r"(qb));
    f = __uint_as_float(b) - 33554430.0f;
    return __float_as_uint(f) >> 23;
#else
    const uint32_t b = 0x4BFFFFFFu + (qa - qb);
    memcpy(&f, &b, 4);
    f -= 33554430.0f;
    uint32_t r;
    memcpy(&r, &f, 4);
    return r >> 23;
#endif
}
GPUAR_HD uint32_t funnel_r_wrap(uint32_t lo, uint32_t hi, uint32_t s)    // low 32 bits of {hi:lo} >> (s mod 32)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u));
#endif
}

//   in : L, R; lo = cum[s], hi = cum[s+1]; (m, sh) for the current total
//   out: L, R renormalised; L1 = lower bound and S1 = upper bound + 1 before renormalisation;
//        t = total shift (0..16); As = L1 << t with bit 15 kept (bit 15 set <=> u != 0)
GPUAR_HD void narrow_total(uint32_t &L, uint32_t &R, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh,
                           uint32_t &L1, uint32_t &S1, uint32_t &t, uint32_t &As)
{
    const uint32_t qa = div_total(hi * R, m, sh);
    const uint32_t qb = div_total(lo * R, m, sh);
    const uint32_t E = width_exponent(qa, qb);                // = 16 - s1 (mod 32)
    L1 = L + qb;
    S1 = L + qa;
    const uint32_t A = funnel_r_wrap(L1 << 16, 0u, E);        // L1 << s1
    const uint32_t R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    const uint32_t sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15; // carry out of the low 15 bits (sum < 2^16)
    R = R1 >> sx;
    As = A >> sx;
    L = As & 0x7FFFu;
    t = 16u - (E & 31u) - sx;
}

// The encoder's CODER warp goes one step further and never applies the decision to the
// range at all: it carries R1 together with the pending halving sx, and the next step
// divides with sx added to the shift of the reciprocal division:
// floor(c * (R1 >> sx) / T) = floor(floor(c * R1 / T) >> sx), R1 being even whenever sx = 1
// (sx = 1 needs s1 >= 1).  The products and the multiply-high of the next step then start
// from R1 directly, in parallel with the decision.
//   state: L (15 bit), R1, sx;  true range = R1 >> sx.   Start: L = 0, R1 = 65536, sx = 0.
//   out  : L1 = lower bound, S1 = upper bound + 1, both before renormalisation
GPUAR_HD void narrow_lazy(uint32_t &L, uint32_t &R1, uint32_t &sx, uint32_t lo, uint32_t hi, uint32_t m,
                          uint32_t sh, uint32_t &L1, uint32_t &S1)
{
    const uint32_t qa = (mulhi32(hi * R1, m) >> sh) >> sx;
    const uint32_t qb = (mulhi32(lo * R1, m) >> sh) >> sx;
    const uint32_t E = width_exponent(qa, qb);
    L1 = L + qb;
    S1 = L + qa;
    const uint32_t A = funnel_r_wrap(L1 << 16, 0u, E);        // L1 << s1
    R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15;                // carry out of the low 15 bits (sum < 2^16)
    L = (A >> sx) & 0x7FFFu;
}

// L1 | U1 << 16 from L1 and S1 = U1 + 1, as two multiply-adds
GPUAR_HD uint32_t pack_bounds(uint32_t L1, uint32_t S1) { return S1 * 0x10000u + (L1 - 0x10000u); }

// ---- encoder bit sink: MSB-first stream (gpuar_kernel.cu:128-151), flushed as 32-bit words.
// Branch free: after appending, at most one whole word is ready; it is stored under a
// predicate and the counters are advanced arithmetically.
struct BitSink {
    uint64_t acc;
    uint32_t nb;        // valid low bits of acc, < 32 between calls
    uint32_t widx;      // next word of the slot's bitstream
    uint32_t wcap;      // writable words
    uint32_t *words;    // first bitstream word of the slot (4-byte aligned)

    GPUAR_HD void put(uint32_t val, uint32_t len)    // len <= 32, val < 2^len
    {
        acc = (acc << len) | val;
        nb += len;
        const uint32_t w = (uint32_t)(acc >> (nb & 31u));          // the oldest 32 bits when nb >= 32
        if (nb >= 32u && widx < wcap) words[widx] = bswap32(w);
        widx += nb >> 5;
        nb &= 31u;
    }
};

// n copies of `bit`, any n (the rare long-underflow path and the end-of-packet flush)
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
BitSink put_run(BitSink out, uint32_t bit, uint32_t n)    // by value: keeps the sink in registers
{
    const uint32_t ones = bit ? 0xFFFFu : 0u;
    while (n > 16u) { out.put(ones, 16u); n -= 16u; }
    out.put(ones & ((1u << n) - 1u), n);
    return out;
}

// Bits of one symbol: the top k bits of U1 with, right after the first of them, `pend`
// inverted copies of it (gpuar_kernel.cu:325-336); then the underflow count carries on.
// emit_field builds the common case (pend <= 16) as one field of k + pend <= 32 bits without
// branches; emit_long is the rare remainder.
GPUAR_HD bool emit_is_long(uint32_t pend, uint32_t k) { return k != 0u && pend > 16u; }

GPUAR_HD void emit_field(BitSink &out, uint32_t &pend, uint32_t k, uint32_t u, uint32_t U1)
{
    const uint32_t b = U1 >> 15;
    const uint32_t km1 = k ? k - 1u : 0u;
    const uint32_t rest = (U1 >> (16u - k)) & ((1u << km1) - 1u);
    const uint32_t head = (1u << (pend & 31u)) - (b ^ 1u);        // b, then pend x !b
    const uint32_t val = k ? ((head << km1) | rest) : 0u;
    const uint32_t len = k ? k + pend : 0u;
    out.put(val, len);
    pend = k ? u : pend + u;
}

GPUAR_HD void emit_long(BitSink &out, uint32_t &pend, uint32_t k, uint32_t u, uint32_t U1)   // k != 0, pend > 16
{
    const uint32_t b = U1 >> 15;
    out.put(b, 1u);
    out = put_run(out, b ^ 1u, pend);
    out.put((U1 >> (16u - k)) & ((1u << (k - 1u)) - 1u), k - 1u);
    pend = u;
}

GPUAR_HD void emit_symbol(BitSink &out, uint32_t &pend, uint32_t k, uint32_t u, uint32_t U1)
{
    if (emit_is_long(pend, k)) emit_long(out, pend, k, u, U1);
    else emit_field(out, pend, k, u, U1);
}

// The same emission from a two-word descriptor, so that a warp other than the one that owns
// the bit sink can do everything that does not depend on `pend` (encode_ws.cu, FIELD -> BITS).
// The fields sit where the consumer gets each of them with one instruction (mask of the low
// bits, plain shift from the top, or a wrapping shift amount that ignores the bits above it):
//   w0: bits 0-4 k | 5-9 max(k,1)-1 | 10 valid | 28-31 u
//   w1: bits 0-14 the k-1 bits after the first | 31 the first bit, inverted
// The producer builds them with multiply-adds where the fields cannot overlap (FMA pipe; the
// integer ALU pipe is the one these kernels saturate).
struct FieldDesc { uint32_t w0, w1; };

GPUAR_HD FieldDesc pack_field(uint32_t k, uint32_t u, uint32_t c)       // c = L1 | U1 << 16
{
    const uint32_t U1 = c >> 16;
    const uint32_t km1 = k ? k - 1u : 0u;
    const uint32_t rest = (U1 >> (16u - k)) & ~(0xFFFFFFFFu << km1);
    FieldDesc d;
    d.w0 = (u << 28) + (km1 * 32u + k) + 1024u;
    d.w1 = rest | (~c & 0x80000000u);
    return d;
}

GPUAR_HD uint32_t shl_wrap(uint32_t x, uint32_t s)            // x << (s mod 32)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(0u, x, s);
#else
    return x << (s & 31u);
#endif
}

GPUAR_HD void emit_packed(BitSink &out, uint32_t &pend, FieldDesc d)        // requires !emit_is_long(pend, k)
{
    const uint32_t k = d.w0 & 31u, u = d.w0 >> 28;
    const uint32_t head = (1u << (pend & 31u)) - (d.w1 >> 31);
    const uint32_t val = shl_wrap(head, d.w0 >> 5) | (d.w1 & 0x7FFFu);
    out.put(k ? val : 0u, k ? k + pend : 0u);
    pend = k ? u : pend + u;
}

GPUAR_HD void emit_packed_any(BitSink &out, uint32_t &pend, FieldDesc d)
{
    const uint32_t k = d.w0 & 31u;
    if (emit_is_long(pend, k)) {
        const uint32_t nb = d.w1 >> 31;
        out.put(nb ^ 1u, 1u);
        out = put_run(out, nb, pend);
        out.put(d.w1 & 0x7FFFu, k - 1u);
        pend = d.w0 >> 28;
    } else {
        emit_packed(out, pend, d);
    }
}
GPUAR_HD bool field_valid(FieldDesc d) { return (d.w0 & 1024u) != 0u; }

// End of packet: bit 14 of L, then pend+1 inverted copies (gpuar_kernel.cu:379-388); zero
// padding to a byte (:430-439).  Writes the tail bytes and the 4-byte packet header (:525-528)
// at `slot` (out.words == slot + 4).  Returns compLen.
GPUAR_HD uint32_t finish_packet(BitSink &out, uint32_t L, uint32_t pend, uint8_t *slot, uint32_t raw_len)
{
    const uint32_t b = (L >> 14) & 1u;
    out.put(b, 1u);
    out = put_run(out, b ^ 1u, pend + 1u);
    uint32_t bytes = out.widx * 4u;
    if (out.nb) {
        const uint32_t tail = (out.nb + 7u) >> 3;
        const uint32_t w = (uint32_t)(out.acc << (32u - out.nb));  // left-aligned, zero padded
        uint8_t *bp = reinterpret_cast<uint8_t *>(out.words + out.widx);
        if (out.widx < out.wcap)
            for (uint32_t t = 0; t < tail; ++t) bp[t] = (uint8_t)(w >> (24u - 8u * t));
        bytes += tail;
    }
    const uint32_t comp = bytes + kHdr;
    *reinterpret_cast<uint32_t *>(slot) = (comp & 0xFFFFu) | (raw_len << 16);   // u16 compLen | u16 rawLen, LE
    return comp;
}

// ---- decoder: target = ((code - L + 1) * T - 1) / range  (getUnscaledCode, :703-716)
// num < 2^30 and 2^14 < range <= 2^16: a float estimate (approximate reciprocal, a few ulp)
// is within 1 of the quotient; one correction step makes it exact.
GPUAR_HD uint32_t divide_exact(uint32_t num, uint32_t range)      // num < 2^30, 2^14 < range <= 2^16
{
#if defined(__CUDA_ARCH__)
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(__uint2float_rz(range)));
    uint32_t q = (uint32_t)(__uint2float_rz(num) * rcp);
#else
    uint32_t q = (uint32_t)((double)(float)num * (double)(1.0f / (float)range));
#endif
    const int32_t r = (int32_t)(num - q * range);
    if (r < 0) --q;
    else if (r >= (int32_t)range) ++q;
    return q;
}

GPUAR_HD uint32_t unscale_range(uint32_t code, uint32_t L, uint32_t range, uint32_t T)
{
    const uint32_t num = (((code - L) & 0xFFFFu) + 1u) * T - 1u;
    return divide_exact(num, range);
}
GPUAR_HD uint32_t unscale(uint32_t code, uint32_t L, uint32_t V, uint32_t T)
{
    return unscale_range(code, L, 65536u - V - L, T);
}

// byte permute of the 8 bytes {b:a} (prmt.b32, default mode).  Selector nibbles 0..7 pick a
// byte; nibbles with bit 3 set (which the hardware turns into sign replication) only ever
// occur in result bytes the callers ignore, so the host emulation may differ there.
GPUAR_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
    return r;
#endif
}

// ---- decoder model: 4-ary cumulative-count tree over the 256 symbols, 85 nodes.
// A node is four u16 slots in one 64-bit word: (0, t0, t1, t2) with t0 = |child0|,
// t1 = t0 + |child1|, t2 = t1 + |child2|.  Slot 0 is the constant 0 so that "the running
// sum just below child c" is simply slot c.  Levels: 1, 4, 16, 64 nodes whose children span
// 64, 16, 4, 1 symbols.  The root lives in registers; node n >= 1 at nodes[(n - 1) * stride].
constexpr uint32_t kTreeNodes = 1 + 4 + 16 + 64;
constexpr uint32_t kTreeStored = kTreeNodes - 1;

GPUAR_HD uint64_t tree_node_init(uint32_t span)    // every symbol count 1
{
    return ((uint64_t)span << 16) | ((uint64_t)(2u * span) << 32) | ((uint64_t)(3u * span) << 48);
}

GPUAR_HD void tree_init(uint64_t &root, uint64_t *nodes, uint32_t stride)
{
    root = tree_node_init(64);
    uint32_t n = 0;
    for (uint32_t q = 0; q < 4; ++q, ++n) nodes[n * stride] = tree_node_init(16);
    for (uint32_t q = 0; q < 16; ++q, ++n) nodes[n * stride] = tree_node_init(4);
    for (uint32_t q = 0; q < 64; ++q, ++n) nodes[n * stride] = tree_node_init(1);
}

// One level, branch free.  rem = target relative to the node (0 <= rem < node total),
// room = node total - rem.  All three thresholds are compared at once: with the guard
// bit 0x8000 in every 16-bit slot, slot j of  (rem,rem,rem,rem) + guards - node  is
// 0x8000 + rem - slot_j, whose bit 15 says rem >= slot_j, and whose low bits are already
// the child-relative remainder.  Returns the child index c and bumps the node.
GPUAR_HD uint32_t tree_level(uint64_t &node, uint32_t &rem, uint32_t &room)
{
    const uint32_t lo = (uint32_t)node, hi = (uint32_t)(node >> 32);
    const uint32_t rr = rem * 0x10001u + 0x80008000u;
    const uint32_t dlo = rr - lo, dhi = rr - hi;
    const uint32_t c = (dlo >> 31) + ((dhi >> 15) & 1u) + (dhi >> 31);
    const uint32_t p = prmt(dlo, dhi, 0x3210u + 0x2222u * c);     // slot c | slot c+1 << 16
    rem = p & 0x7FFFu;
    room = c == 3u ? room : 0x8000u - (p >> 16);
    node += 0x0001000100010000ull << (16u * c);                    // +1 on every slot above c
    return c;
}

// The same level for the THROUGHPUT decoder (many warps per scheduler).  The three comparison bits are the
// child index AND the update: "+1 on every slot above c" is slot1 += 1 - b0, slot2 += 1 - b1, slot3 += 1 - b2
// (b_j = rem >= t_j; slots never carry into each other), i.e. two multiply-adds on the halves instead of a
// 64-bit variable shift and a carry chain: 1 GiB decode 8.18 -> 8.00 ms.  (The `room` select as a
// multiply-add too: 8.35 ms; measured, not kept -- profiles/r2_kernel_experiments.md.)
GPUAR_HD uint32_t mad32(uint32_t a, uint32_t b, uint32_t c)       // a * b + c, kept a multiply-add (FMA pipe)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a * b + c;
#endif
}
GPUAR_HD uint32_t tree_level_fma(uint64_t &node, uint32_t &rem, uint32_t &room)
{
    const uint32_t lo = (uint32_t)node, hi = (uint32_t)(node >> 32);
    const uint32_t rr = rem * 0x10001u + 0x80008000u;
    const uint32_t dlo = rr - lo, dhi = rr - hi;
    const uint32_t b0 = dlo >> 31, b1 = (dhi >> 15) & 1u, b2 = dhi >> 31;
    const uint32_t c = b0 + b1 + b2;
    const uint32_t p = prmt(dlo, dhi, 0x3210u + 0x2222u * c);     // slot c | slot c+1 << 16
    rem = p & 0x7FFFu;
    room = c == 3u ? room : 0x8000u - (p >> 16);
    const uint32_t nlo = mad32(b0, 0xFFFF0000u, lo + 0x10000u);               // slot1 += 1 - b0
    const uint32_t nhi = mad32(b2, 0xFFFF0000u, hi + 0x10001u) - b1;          // slot2 += 1 - b1, slot3 += 1 - b2
    node = ((uint64_t)nhi << 32) | nlo;
    return c;
}

// Finds the symbol whose cumulative interval holds `target` (getSymbolFromProbability,
// :727-763), returns it with lo = cum[s], cnt = count[s], and bumps count[s] (:288).
GPUAR_HD uint32_t tree_decode(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t target, uint32_t total,
                              uint32_t &lo, uint32_t &cnt)
{
    auto level = [](uint64_t &node, uint32_t &rem, uint32_t &room) { return tree_level_fma(node, rem, room); };
    uint32_t rem = target, room = total - target;
    uint32_t idx = level(root, rem, room);
    {
        uint64_t *n = nodes + idx * stride;
        uint64_t v = *n;
        idx = idx * 4u + level(v, rem, room);
        *n = v;
    }
    {
        uint64_t *n = nodes + (4u + idx) * stride;
        uint64_t v = *n;
        idx = idx * 4u + level(v, rem, room);
        *n = v;
    }
    {
        uint64_t *n = nodes + (20u + idx) * stride;
        uint64_t v = *n;
        idx = idx * 4u + level(v, rem, room);
        *n = v;
    }
    lo = target - rem;
    cnt = rem + room;
    return idx;
}

// Latency-oriented variant of tree_decode.  target = floor(num / range), so for an integer
// threshold t:  t <= target  <=>  t * range <= num.  The two top levels are therefore decided
// with multiplications only, while the divide is still in flight, and the level-1 and
// level-2 node loads are issued before `target` exists; the two lower levels then run on
// `target` as in tree_decode.  Same result, the dependent chain is ~2 shared-memory round
// trips shorter.  (t * range <= 8448 * 65536 < 2^30, num < 2^30: 32-bit compares are exact.)
#ifndef GPUAR_DEC_SPEC
#define GPUAR_DEC_SPEC 7            // tuning knob: bits 0 / 1 / 2 = speculative loads for tree levels 1 / 2 / 3
#endif

// bitwise select: m = all ones picks x, m = 0 picks y (one LOP3)
GPUAR_HD uint32_t bsel(uint32_t m, uint32_t x, uint32_t y) { return (x & m) | (y & ~m); }
// one of four by the three "threshold above target" masks of a level (m0 implies m1 implies m2):
// child 0 iff m0, child 1 iff m1 & ~m0, child 2 iff m2 & ~m1, child 3 iff ~m2; two selects deep
GPUAR_HD uint32_t pick4(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return bsel(m1, bsel(m0, a, b), bsel(m2, c, d));
}

// Variant of tree_decode_early_range whose node loads do not wait for the child index: the
// four candidates of the next level are loaded as soon as their parent is known (level 1: at
// the top of the step, level 2: once the level-0 child is known) and the right one is picked
// with two bitwise selects on the sign masks of the threshold tests -- a shared-memory round
// trip (~30 cycles) on the dependent chain becomes ~10 cycles of logic, for three more loads
// and six selects per level.  Same result as tree_decode_early_range.
template <int kSpec>
GPUAR_HD uint32_t tree_decode_spec_range(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t code,
                                         uint32_t L, uint32_t range, uint32_t T, uint32_t &lo, uint32_t &cnt)
{
    const uint32_t num = (((code - L) & 0xFFFFu) + 1u) * T - 1u;
    const uint32_t nr = 0u - range;
    auto mask = [](uint32_t d) { return (uint32_t)((int32_t)d >> 31); };   // all ones if the threshold lies above the target
    // level-1 candidates: independent of everything in this step
    const uint64_t a1 = nodes[0], b1 = nodes[stride], c1n = nodes[2u * stride], d1n = nodes[3u * stride];
    // level 0
    const uint32_t r0 = (uint32_t)root, r1 = (uint32_t)(root >> 32);
    const uint32_t t0 = r0 >> 16, t1 = r1 & 0xFFFFu, t2 = r1 >> 16;
    const uint32_t e0 = t0 * nr + num, e1 = t1 * nr + num, e2 = t2 * nr + num;     // num - threshold * range
    const uint32_t m0 = mask(e0), m1 = mask(e1), m2 = mask(e2);
    const uint32_t c0 = 3u + m0 + m1 + m2;
    uint64_t *const p1 = nodes + c0 * stride;
    const uint32_t q0 = pick4(m0, m1, m2, (uint32_t)a1, (uint32_t)b1, (uint32_t)c1n, (uint32_t)d1n);
    const uint32_t q1 = pick4(m0, m1, m2, (uint32_t)(a1 >> 32), (uint32_t)(b1 >> 32), (uint32_t)(c1n >> 32),
                              (uint32_t)(d1n >> 32));
    const uint32_t num1 = pick4(m0, m1, m2, num, e0, e1, e2);       // what is left of num below the child
    const uint32_t below0 = pick4(m0, m1, m2, 0u, t0, t1, t2);
    const uint32_t above0 = pick4(m0, m1, m2, t0, t1, t2, T);
    root += 0x0001000100010000ull << (16u * c0);
    // level-2 candidates of that child
    uint64_t *const g2 = nodes + (4u + c0 * 4u) * stride;
    uint64_t a2 = 0, b2 = 0, c2n = 0, d2n = 0;
    if (kSpec & 2) {
        a2 = g2[0];
        b2 = g2[stride];
        c2n = g2[2u * stride];
        d2n = g2[3u * stride];
    }
    // level 1, absolute thresholds
    const uint32_t u0 = q0 >> 16, u1 = q1 & 0xFFFFu, u2 = q1 >> 16;
    const uint32_t f0 = u0 * nr + num1, f1 = u1 * nr + num1, f2 = u2 * nr + num1;
    const uint32_t k0 = mask(f0), k1 = mask(f1), k2 = mask(f2);
    const uint32_t c1 = 3u + k0 + k1 + k2;
    uint32_t idx = c0 * 4u + c1;
    uint64_t *const p2 = g2 + c1 * stride;
    uint64_t *const g3 = nodes + (20u + idx * 4u) * stride;
    uint64_t a3 = 0, b3 = 0, c3n = 0, d3n = 0;
    if (kSpec & 4) {
        a3 = g3[0];
        b3 = g3[stride];
        c3n = g3[2u * stride];
        d3n = g3[3u * stride];
    }
    uint64_t n2;
    if (kSpec & 2) {
        const uint32_t lo2 = pick4(k0, k1, k2, (uint32_t)a2, (uint32_t)b2, (uint32_t)c2n, (uint32_t)d2n);
        const uint32_t hi2 = pick4(k0, k1, k2, (uint32_t)(a2 >> 32), (uint32_t)(b2 >> 32), (uint32_t)(c2n >> 32),
                                   (uint32_t)(d2n >> 32));
        n2 = ((uint64_t)hi2 << 32) | lo2;
    } else {
        n2 = *p2;
    }
    const uint32_t below1 = below0 + pick4(k0, k1, k2, 0u, u0, u1, u2);
    const uint32_t above1 = pick4(k0, k1, k2, below0 + u0, below0 + u1, below0 + u2, above0);
    *p1 = (((uint64_t)q1 << 32) | q0) + (0x0001000100010000ull << (16u * c1));
    // levels 2 and 3 on the quotient
    const uint32_t target = divide_exact(num, range);
    uint32_t rem = target - below1, room = above1 - target;
    if (kSpec & 4) {
        // the four leaf nodes below (c0, c1) were requested before the quotient existed
        const uint32_t z0 = (uint32_t)n2, z1 = (uint32_t)(n2 >> 32);
        const uint32_t rr = rem * 0x10001u + 0x80008000u;
        const uint32_t dlo = rr - z0, dhi = rr - z1;               // bit 31 / 15: rem >= slot (tree_level)
        const uint32_t j0 = ~mask(dlo), j1 = ~mask(dhi << 16), j2 = ~mask(dhi);
        const uint32_t lo3 = pick4(j0, j1, j2, (uint32_t)a3, (uint32_t)b3, (uint32_t)c3n, (uint32_t)d3n);
        const uint32_t hi3 = pick4(j0, j1, j2, (uint32_t)(a3 >> 32), (uint32_t)(b3 >> 32), (uint32_t)(c3n >> 32),
                                   (uint32_t)(d3n >> 32));
        const uint32_t c2 = tree_level(n2, rem, room);
        *p2 = n2;
        uint64_t v = ((uint64_t)hi3 << 32) | lo3;
        idx = idx * 4u + c2;
        uint64_t *n = nodes + (20u + idx) * stride;
        idx = idx * 4u + tree_level(v, rem, room);
        *n = v;
    } else {
        idx = idx * 4u + tree_level(n2, rem, room);
        *p2 = n2;
        uint64_t *n = nodes + (20u + idx) * stride;
        uint64_t v = *n;
        idx = idx * 4u + tree_level(v, rem, room);
        *n = v;
    }
    lo = target - rem;
    cnt = rem + room;
    return idx;
}

template <int kSpec = GPUAR_DEC_SPEC>
GPUAR_HD uint32_t tree_decode_early_range(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t code,
                                          uint32_t L, uint32_t range, uint32_t T, uint32_t &lo, uint32_t &cnt)
{
    if (kSpec) return tree_decode_spec_range<kSpec>(root, nodes, stride, code, L, range, T, lo, cnt);
    const uint32_t num = (((code - L) & 0xFFFFu) + 1u) * T - 1u;
    // "threshold * range <= num" as the sign of num - threshold * range (everything < 2^30): one
    // multiply-add and one shift per threshold, no predicates (their write-to-use latency is
    // several times that of a register, and turning them into numbers costs a select each)
    const uint32_t nr = 0u - range;
    auto above = [](uint32_t d) { return d >> 31; };               // 1 if the threshold lies above the target
    // level 0
    const uint32_t r0 = (uint32_t)root, r1 = (uint32_t)(root >> 32);
    const uint32_t c0 = 3u - (above((r0 >> 16) * nr + num) + above((r1 & 0xFFFFu) * nr + num) +
                              above((r1 >> 16) * nr + num));
    uint64_t *const p1 = nodes + c0 * stride;
    uint64_t n1 = *p1;
    const uint32_t w0 = prmt(r0, r1, 0x3210u + 0x2222u * c0);        // slot c0 | slot c0+1 << 16
    const uint32_t below0 = w0 & 0xFFFFu;
    const uint32_t above0 = c0 == 3u ? T : (w0 >> 16);
    root += 0x0001000100010000ull << (16u * c0);
    // level 1, absolute thresholds
    const uint32_t q0 = (uint32_t)n1, q1 = (uint32_t)(n1 >> 32);
    const uint32_t num1 = below0 * nr + num;                       // what is left of num below this node
    const uint32_t c1 = 3u - (above((q0 >> 16) * nr + num1) + above((q1 & 0xFFFFu) * nr + num1) +
                              above((q1 >> 16) * nr + num1));
    uint32_t idx = c0 * 4u + c1;
    uint64_t *const p2 = nodes + (4u + idx) * stride;
    uint64_t n2 = *p2;
    const uint32_t w1 = prmt(q0, q1, 0x3210u + 0x2222u * c1);
    const uint32_t below1 = below0 + (w1 & 0xFFFFu);
    const uint32_t above1 = c1 == 3u ? above0 : below0 + (w1 >> 16);
    n1 += 0x0001000100010000ull << (16u * c1);
    *p1 = n1;
    // levels 2 and 3 on the quotient
    const uint32_t target = divide_exact(num, range);
    uint32_t rem = target - below1, room = above1 - target;
    idx = idx * 4u + tree_level(n2, rem, room);
    *p2 = n2;
    {
        uint64_t *n = nodes + (20u + idx) * stride;
        uint64_t v = *n;
        idx = idx * 4u + tree_level(v, rem, room);
        *n = v;
    }
    lo = target - rem;
    cnt = rem + room;
    return idx;
}

GPUAR_HD uint32_t tree_decode_early(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t code, uint32_t L,
                                    uint32_t V, uint32_t T, uint32_t &lo, uint32_t &cnt)
{
    return tree_decode_early_range(root, nodes, stride, code, L, 65536u - V - L, T, lo, cnt);
}

// Encoder side of the model (own tree, own leaf format).  The symbol is known, so the child
// indices are its bit pairs and the node loads are independent of each other.
//   levels 0-2: W-form nodes (0, t0, t1, t2) as in the decoder: slot c = count below child c;
//               slot 0 is constant zero, which also provides the zero bytes for the permute.
//   level 3   : four plain counts (n0, n1, n2, n3) of the node's four symbols.
// Returns lo = cum[s], cnt = count[s] before this symbol and bumps count[s]
// (getRange(LOWER/UPPER) + update, gpuar_kernel.cu:215-238,272,279,288).
GPUAR_HD uint64_t enc_leaf_init() { return 0x0001000100010001ull; }

GPUAR_HD void enc_tree_init(uint64_t &root, uint64_t *nodes, uint32_t stride)
{
    root = tree_node_init(64);
    uint32_t n = 0;
    for (uint32_t q = 0; q < 4; ++q, ++n) nodes[n * stride] = tree_node_init(16);
    for (uint32_t q = 0; q < 16; ++q, ++n) nodes[n * stride] = tree_node_init(4);
    for (uint32_t q = 0; q < 64; ++q, ++n) nodes[n * stride] = enc_leaf_init();
}

// Node increments of the encoder (a constant-memory table of the four increments instead of the
// variable 64-bit shift was measured: the per-lane index replays the LDC, slower everywhere).
GPUAR_HD void bump_above(uint64_t &node, uint32_t c) { node += 0x0001000100010000ull << (16u * c); }   // +1 on every slot above c
GPUAR_HD void bump_at(uint64_t &node, uint32_t c) { node += 1ull << (16u * c); }                       // +1 on slot c

GPUAR_HD uint32_t enc_upper(uint64_t &node, uint32_t c)   // slot c, then +1 above c
{
    const uint32_t below = prmt((uint32_t)node, (uint32_t)(node >> 32), 0x0010u + 0x0022u * c);
    bump_above(node, c);
    return below;
}

// The levels are independent of each other given the symbol, so the model can be split:
// levels 0-1 (root + 4 nodes) and levels 2-3 (16 nodes + 64 leaves).  cum[s] is the sum
// of the two partial results.
GPUAR_HD uint32_t tree_encode_upper_at(uint64_t &root, uint64_t *node1, uint32_t c0, uint32_t c1)
{
    uint64_t n1 = *node1;
    uint32_t acc = enc_upper(root, c0);
    acc += enc_upper(n1, c1);
    *node1 = n1;
    return acc;
}
GPUAR_HD uint32_t tree_encode_upper(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t s)
{
    return tree_encode_upper_at(root, nodes + (s >> 6) * stride, s >> 6, (s >> 4) & 3u);
}

GPUAR_HD uint32_t tree_encode_mid_at(uint64_t *node2, uint32_t c)                      // level 2 alone
{
    uint64_t n2 = *node2;
    const uint32_t acc = enc_upper(n2, c);
    *node2 = n2;
    return acc;
}
GPUAR_HD uint32_t tree_encode_mid(uint64_t *nodes, uint32_t stride, uint32_t s)
{
    return tree_encode_mid_at(nodes + (4u + (s >> 4)) * stride, (s >> 2) & 3u);
}

GPUAR_HD uint32_t tree_encode_leaf_at(uint64_t *node3, uint32_t c, uint32_t &cnt)      // level 3 alone
{
    uint64_t n3 = *node3;
    // prefix sums of the four plain counts in W-form, then the same permute
    const uint32_t l = (uint32_t)n3, h = (uint32_t)(n3 >> 32);
    const uint32_t x = l * 0x10001u;                               // (n0, n0+n1)
    const uint32_t s01 = x >> 16;
    const uint32_t s012 = s01 + (h & 0xFFFFu);
    const uint32_t acc = prmt(x << 16, s01 | (s012 << 16), 0x0010u + 0x0022u * c);
    cnt = prmt(l, h, 0x3210u + 0x2222u * c) & 0xFFFFu;
    bump_at(n3, c);
    *node3 = n3;
    return acc;
}
GPUAR_HD uint32_t tree_encode_leaf(uint64_t *nodes, uint32_t stride, uint32_t s, uint32_t &cnt)
{
    return tree_encode_leaf_at(nodes + (20u + (s >> 2)) * stride, s & 3u, cnt);
}

// Four symbols at a time (encode_ws.cu): the bit fields a level needs are masked for all four
// bytes of an input word at once and picked per symbol with one byte permute, instead of a
// shift-and-mask pair per field and symbol (the integer ALU pipe is what that kernel saturates).
//   level 0-1: a = c0 of each byte, b = c1        level 2: a = (s >> 4) << 4, b = c2
//   level 3  : a = (s >> 2) << 2,    b = c3
struct WordFields { uint32_t a, b; };
GPUAR_HD WordFields word_fields_upper(uint32_t w) { return {(w >> 6) & 0x03030303u, (w >> 4) & 0x03030303u}; }
GPUAR_HD WordFields word_fields_mid(uint32_t w) { return {w & 0xF0F0F0F0u, (w >> 2) & 0x03030303u}; }
GPUAR_HD WordFields word_fields_leaf(uint32_t w) { return {w & 0xFCFCFCFCu, w & 0x03030303u}; }
GPUAR_HD uint32_t byte_of(uint32_t w, uint32_t j) { return prmt(w, 0u, 0x4440u + j); }   // byte j, zero extended

GPUAR_HD uint64_t *byte_offset(uint64_t *p, uint32_t bytes)
{
    return reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(p) + bytes);
}
GPUAR_HD uint32_t tree_encode_upper_word(uint64_t &root, uint64_t *nodes, uint32_t stride, WordFields f, uint32_t j)
{
    const uint32_t c0 = byte_of(f.a, j);
    return tree_encode_upper_at(root, byte_offset(nodes, c0 * stride * 8u), c0, byte_of(f.b, j));
}
GPUAR_HD uint32_t tree_encode_mid_word(uint64_t *nodes, uint32_t stride, WordFields f, uint32_t j)
{
    return tree_encode_mid_at(byte_offset(nodes + 4u * stride, (byte_of(f.a, j) * stride) >> 1), byte_of(f.b, j));
}
GPUAR_HD uint32_t tree_encode_leaf_word(uint64_t *nodes, uint32_t stride, WordFields f, uint32_t j, uint32_t &cnt)
{
    return tree_encode_leaf_at(byte_offset(nodes + 20u * stride, byte_of(f.a, j) * stride * 2u), byte_of(f.b, j), cnt);
}

GPUAR_HD uint32_t tree_encode_lower(uint64_t *nodes, uint32_t stride, uint32_t s, uint32_t &cnt)
{
    uint64_t *const p2 = nodes + (4u + (s >> 4)) * stride;
    uint64_t *const p3 = nodes + (20u + (s >> 2)) * stride;
    uint64_t n2 = *p2, n3 = *p3;
    uint32_t acc = enc_upper(n2, (s >> 2) & 3u);
    // leaf: prefix sums of the four plain counts in W-form, then the same permute
    const uint32_t c = s & 3u, l = (uint32_t)n3, h = (uint32_t)(n3 >> 32);
    const uint32_t x = l * 0x10001u;                               // (n0, n0+n1)
    const uint32_t s01 = x >> 16;
    const uint32_t s012 = s01 + (h & 0xFFFFu);
    acc += prmt(x << 16, s01 | (s012 << 16), 0x0010u + 0x0022u * c);
    cnt = prmt(l, h, 0x3210u + 0x2222u * c) & 0xFFFFu;
    bump_at(n3, c);
    *p2 = n2;
    *p3 = n3;
    return acc;
}

GPUAR_HD void tree_encode(uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t s, uint32_t &lo, uint32_t &cnt)
{
    const uint32_t up = tree_encode_upper(root, nodes, stride, s);
    lo = up + tree_encode_lower(nodes, stride, s, cnt);
}

// ---- funnel shifts (shf.l / shf.r.clamp) with host fall-backs
GPUAR_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s)      // upper 32 bits of {hi:lo} << s, s in 0..31
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (s & 31u)) >> 32);
#endif
}
GPUAR_HD uint32_t shr_clamp(uint32_t x, uint32_t s)                    // x >> s with s in 0..32 (32 gives 0)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(x, 0u, s);
#else
    return s >= 32u ? 0u : x >> s;
#endif
}

// ---- decoder bit source: 64-bit window {hi:lo}, next stream bit = MSB of hi, fed one 32-bit
// word at a time.  `have` = valid bits; >= 33 at the top of every step, so hi is always whole.
struct BitSource {
    uint32_t hi, lo;
    uint32_t have;
    GPUAR_HD void start(uint64_t window, uint32_t valid)
    {
        hi = (uint32_t)(window >> 32);
        lo = (uint32_t)window;
        have = valid;
    }
    GPUAR_HD void skip(uint32_t t)                   // t <= 31
    {
        hi = funnel_l(lo, hi, t);
        lo <<= t;
        have -= t;
    }
    GPUAR_HD uint32_t take(uint32_t t)               // t <= 31
    {
        const uint32_t bits = funnel_l(hi, 0u, t);
        skip(t);
        return bits;
    }
    GPUAR_HD bool hungry() const { return have <= 32u; }
    GPUAR_HD void feed_if(bool on, uint32_t be_word)  // on == hungry(): have <= 32, lo is empty then
    {
        hi |= shr_clamp(be_word, have);              // have >= 33 when not hungry: the clamped shift gives 0
        lo = on ? be_word << ((32u - have) & 31u) : lo;
        have += on ? 32u : 0u;
    }
    GPUAR_HD void feed(uint32_t be_word) { feed_if(true, be_word); }
};

// readEncodedBits in closed form (:787-836): shift in k+u bits; an underflow run leaves
// the MSB flipped.  One funnel shift moves the bits from the window into the code register.
GPUAR_HD uint32_t advance_code(uint32_t code, uint32_t k, uint32_t u, BitSource &in)
{
    const uint32_t t = k + u;
    const uint32_t next = (funnel_l(in.hi, code, t) & 0xFFFFu) ^ (u ? 0x8000u : 0u);
    in.skip(t);
    return next;
}

// the same from narrow_total's outputs: t = k + u, and bit 15 of As says whether u != 0
GPUAR_HD uint32_t advance_code_total(uint32_t code, uint32_t t, uint32_t As, BitSource &in)
{
    const uint32_t next = (funnel_l(in.hi, code, t) & 0xFFFFu) ^ (As & 0x8000u);
    in.skip(t);
    return next;
}

}  // namespace gpuar
