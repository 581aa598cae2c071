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
// m in 32-bit arithmetic (a 64-bit division is a subroutine call on the device -- once per 32 steps in every
// kernel, and it waited for the warp's outstanding loads): with A = 2^(16+sh) = q1 * T + r1,
// 2^(32+sh) = A * 2^16, so m = q1 * 2^16 + ceil(r1 * 2^16 / T); r1 < T < 2^15 keeps everything below 2^32.
GPUAR_HD uint32_t magic_for(uint32_t T, uint32_t &sh)
{
    sh = shift_for(T);
    const uint32_t A = 1u << (16u + sh);
    const uint32_t q1 = A / T, r1 = A - q1 * T;
    return (q1 << 16) + ((r1 << 16) + T - 1u) / T;
}
GPUAR_HD uint32_t div_total(uint32_t n, uint32_t m, uint32_t sh) { return mulhi32(n, m) >> sh; }

// One interval-narrowing + renormalisation step: gpuar_kernel.cu:256-288, then the closed
// form of the loops at :321-367 (encoder) / :787-836 (decoder).  This is the first-generation step (k and u
// individually, two count-leading-zeros); the kernels run the single-normalisation steps below, and this
// one stays as the independent formulation the host checks compare them with (tests/host_model.cpp).
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
// k and u individually do not matter to anybody: the encoders write the stream as the lower bound's binary
// expansion with carries (encode_math.h), the decoders only shift (decode_math.h).
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

// The steps built on it: narrow_track / narrow_track_products (decode_math.h), narrow_plain / narrow_plain_lazy
// (encode_math.h).

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

// ---- the model: 4-ary cumulative-count tree over the 256 symbols, 85 nodes.
// A packed node is four u16 slots in one 64-bit word: (0, t0, t1, t2) with t0 = |child0|,
// t1 = t0 + |child1|, t2 = t1 + |child2|.  Slot 0 is the constant 0 so that "the running
// sum just below child c" is simply slot c.  Levels: 1, 4, 16, 64 nodes whose children span
// 64, 16, 4, 1 symbols.  The root lives in registers; node n >= 1 at nodes[(n - 1) * stride].
// (The decoders' trees and steps: decode_math.h.)
constexpr uint32_t kTreeNodes = 1 + 4 + 16 + 64;
constexpr uint32_t kTreeStored = kTreeNodes - 1;

GPUAR_HD uint64_t tree_node_init(uint32_t span)    // every symbol count 1
{
    return ((uint64_t)span << 16) | ((uint64_t)(2u * span) << 32) | ((uint64_t)(3u * span) << 48);
}

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

}  // namespace gpuar
