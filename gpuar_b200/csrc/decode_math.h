// decode_math.h -- the decoder's symbol step, second generation (round 2), shared verbatim by
// decode.cu (device) and tests/host_model.cpp (host, g++) like coder_math.h.
//
// Same arithmetic as the reference (getUnscaledCode src/gpuar_kernel.cu:703-716,
// getSymbolFromProbability :727-763, applySymbolRange :256-288, readEncodedBits :787-836); what
// changed against the first-generation step (round 1: quotient, packed levels that hand the room above the
// target down, code register with its underflow flip) is the bookkeeping around it -- the instruction count and, for
// a lone warp, the length of the dependent chain are what bound the decoder (profiles/r2_decode_v2.md):
//   * the decoder carries D = code - lower bound (mod 2^16) instead of `code`: the numerator of the
//     quotient is D * T + T - 1 directly, and an underflow step flips bit 15 of the code AND drops
//     bit 15 of the shifted lower bound, so D' = ((D - qb) << t | next t bits) mod 2^16 needs
//     neither the flip nor the subtraction;
//   * the leaves hold INCLUSIVE prefix sums (s0, s1, s2, s3) of their four symbol counts -- s3 is
//     the leaf's own total -- so cum[s] and cum[s+1] both come out of the leaf and no level has to
//     hand "what is left above the target" down to the next one;
//   * throughput variant (decode_step): quotient with a one-sided correction, then four packed levels;
//   * latency variant (decode_step_latency): no quotient at all -- every level is decided by the sign
//     of num - threshold * range, the next level's four candidate nodes are loaded before the child
//     is known and picked with selects on predicates, and the interval narrowing takes the PRODUCTS
//     cum * range straight from the sign tests.
#pragma once
#include "coder_math.h"

namespace gpuar {

// ---- model of the v2 decoders.  Levels: root (children span 64 symbols), 4 level-1 nodes (16),
// 16 level-2 nodes (4), 64 leaves (1).
//   packed node: (0, t0, t1, t2) u16 x 4, t_j = cumulative size of children 0..j  (as coder_math.h)
//   leaf       : (s0, s1, s2, s3) u16 x 4, s_j = cumulative count of symbols 0..j of the leaf
constexpr uint64_t kLeafInit = 0x0004000300020001ull;             // every count 1 (gpuar_kernel.cu:403-419)

struct alignas(16) Quad { uint32_t x, y, z, w; };                 // a level-1 node of the latency variant: thresholds x, y, z

// byte permute with the hardware's sign-replication mode (selector nibble bit 3): nibble 9 = the MSB of
// byte 1 of `a` replicated.  Byte 1 is the high byte of a 14-bit count here, so nibble 9 IS a zero byte.
GPUAR_HD uint32_t prmt_s(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return prmt(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t nib = (sel >> (4 * k)) & 15u;
        uint32_t byte = (uint32_t)(v >> (8 * (nib & 7u))) & 0xFFu;
        if (nib & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * k);
    }
    return r;
#endif
}

GPUAR_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)  // low 32 bits of {hi:lo} >> s, s in 0..31
{
    return funnel_r_wrap(lo, hi, s);
}

// The quotient of the decoder, target = floor(num / range) with num < 2^30 and 2^14 < range <= 2^16 (so target <
// 2^14), with a ONE-sided correction: the float estimate is scaled down by 1 - 2^-20, which outweighs every
// rounding on the way (conversion toward zero, 1 ulp of the approximate reciprocal, two roundings of the
// products: 3 * 2^-23 in all), so it never exceeds the true quotient and falls short of it by less than
// 0.02 -- its floor is the quotient or one less, and one compare settles it.  (divide_exact of coder_math.h
// corrects both ways: three instructions more.)  Checked on the device over every range
// (gpuar_b200_selfcheck) and on the host lattice (tests/test_host_model.py).
GPUAR_HD uint32_t divide_floor(uint32_t num, uint32_t range)
{
#if defined(__CUDA_ARCH__)
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(__uint2float_rz(range)));
    uint32_t q = (uint32_t)(__uint2float_rz(num) * (rcp * 0.99999904632568359375f));
#else
    const float rcp = 1.0f / (float)range;
    const float numf = num < (1u << 24) ? (float)num : (float)(num & ~63u);    // exact, at or below the value rounded toward zero
    uint32_t q = (uint32_t)(numf * (rcp * 0.99999904632568359375f));
#endif
    if (num - q * range >= range) ++q;
    return q;
}

// One packed level: rem = target relative to the node.  Returns the child index, leaves the
// child-relative remainder in rem and bumps the node (+1 on every slot above the child).
GPUAR_HD uint32_t packed_level(uint32_t &lo, uint32_t &hi, uint32_t &rem)
{
    const uint32_t rr = rem * 0x10001u + 0x80008000u;
    const uint32_t dlo = rr - lo, dhi = rr - hi;                  // slot j: 0x8000 + rem - slot_j (bit 15: rem >= slot_j)
    const uint32_t b0 = dlo >> 31, b1 = (dhi >> 15) & 1u, b2 = dhi >> 31;
    const uint32_t c = b0 + b1 + b2;
    rem = prmt(dlo, dhi, 0x3210u + 0x2222u * c) & 0x7FFFu;
    lo = mad32(b0, 0xFFFF0000u, lo + 0x10000u);                   // slot1 += 1 - b0
    hi = mad32(b2, 0xFFFF0000u, hi + 0x10001u) - b1;              // slot2 += 1 - b1, slot3 += 1 - b2
    return c;
}

// The leaf: returns the child index c, below = s_{c-1} (0 for c = 0), upto = s_c, and bumps s_c .. s_3.
// The pair (s_{c-1}, s_c) is one byte permute whose selector is a shifted constant: 0x1099 (zero, zero,
// byte 0, byte 1), 0x3210, 0x5432, 0x7654 for c = 0 .. 3.
GPUAR_HD uint32_t leaf_level(uint32_t &lo, uint32_t &hi, uint32_t rem, uint32_t &below, uint32_t &upto)
{
    const uint32_t rr = rem * 0x10001u + 0x80008000u;
    const uint32_t dlo = rr - lo, dhi = rr - hi;                  // slot j: 0x8000 + rem - s_j
    const uint32_t b0 = (dlo >> 15) & 1u, b1 = dlo >> 31, b2 = (dhi >> 15) & 1u;
    const uint32_t c = b0 + b1 + b2;
    const uint32_t w = prmt_s(lo, hi, funnel_r(0x54321099u, 0x76u, 8u * c));
    below = w & 0xFFFFu;
    upto = w >> 16;
    lo = mad32(b1, 0xFFFF0000u, lo + 0x10001u) - b0;              // s0 += 1 - b0, s1 += 1 - b1
    hi = hi + 0x10001u - b2;                                      // s2 += 1 - b2, s3 += 1
    return c;
}

// State of one packet's decoder between symbols.
struct DecState {
    uint32_t D;          // (code - lower bound) mod 2^16
    uint32_t L, R;       // lower bound (15 bits) and range
};

// Interval narrowing with the single normalisation (coder_math.h) + the next bits, on (D, L, R): the code
// register is folded into D.
GPUAR_HD void narrow_track(DecState &st, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh, BitSource &in)
{
    const uint32_t qa = div_total(hi * st.R, m, sh);
    const uint32_t qb = div_total(lo * st.R, m, sh);
    const uint32_t E = width_exponent(qa, qb);                    // = 16 - s1 (mod 32)
    const uint32_t A = funnel_r_wrap((st.L + qb) << 16, 0u, E);   // L1 << s1
    const uint32_t R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    const uint32_t sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15;     // one shift less (coder_math.h)
    st.R = R1 >> sx;
    st.L = (A >> sx) & 0x7FFFu;
    const uint32_t t = 16u - (E & 31u) - sx;
    st.D = funnel_l(in.hi, st.D - qb, t) & 0xFFFFu;
    in.skip(t);
}

// ---------------------------------------------------------------- throughput variant
// Quotient first, then four dependent levels.  Tree: packed root in registers, stored nodes
// nodes[n * stride]: n = 0..3 level 1, 4..19 level 2 (packed), 20..83 leaves.
GPUAR_HD void dec_tree_init(uint64_t &root, uint64_t *nodes, uint32_t stride)
{
    root = tree_node_init(64);
    uint32_t n = 0;
    for (uint32_t q = 0; q < 4; ++q, ++n) nodes[n * stride] = tree_node_init(16);
    for (uint32_t q = 0; q < 16; ++q, ++n) nodes[n * stride] = tree_node_init(4);
    for (uint32_t q = 0; q < 64; ++q, ++n) nodes[n * stride] = kLeafInit;
}

GPUAR_HD uint32_t decode_step(DecState &st, uint64_t &root, uint64_t *nodes, uint32_t stride, uint32_t T, uint32_t m,
                              uint32_t sh, BitSource &in)
{
    const uint32_t num = st.D * T + (T - 1u);
    const uint32_t target = divide_floor(num, st.R);
    uint32_t rem = target;
    uint32_t lo = (uint32_t)root, hi = (uint32_t)(root >> 32);
    uint32_t idx = packed_level(lo, hi, rem);
    root = ((uint64_t)hi << 32) | lo;
    {
        uint64_t *const n = nodes + idx * stride;
        const uint64_t v = *n;
        lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        idx = idx * 4u + packed_level(lo, hi, rem);
        *n = ((uint64_t)hi << 32) | lo;
    }
    {
        uint64_t *const n = nodes + (4u + idx) * stride;
        const uint64_t v = *n;
        lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        idx = idx * 4u + packed_level(lo, hi, rem);
        *n = ((uint64_t)hi << 32) | lo;
    }
    uint32_t below, upto;
    {
        uint64_t *const n = nodes + (20u + idx) * stride;
        const uint64_t v = *n;
        lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
        idx = idx * 4u + leaf_level(lo, hi, rem, below, upto);
        *n = ((uint64_t)hi << 32) | lo;
    }
    const uint32_t base = target - rem;                           // cum of everything left of the leaf
    narrow_track(st, base + below, base + upto, m, sh, in);
    return idx;
}

// ---------------------------------------------------------------- latency variant
// target = floor(num / range), so for an integer threshold t:  t <= target  <=>  t * range <= num.  Every level is
// decided by the sign of num - t * range (t * range <= 16368 * 65536 < 2^30, num < 2^30: signs are exact), so there
// is no divide, no conversion and no byte permute on the chain, and no node load waits for its child index: the four
// candidates of a level are requested as soon as their parent is known and the right one is picked with two
// predicated selects per word.  Neither cum[s] nor cum[s+1] is ever formed: the interval narrowing needs the
// PRODUCTS cum * range, and num - cum * range is exactly what the sign test of the chosen threshold computed.
// The root is three registers and the four level-1 nodes are three 32-bit words each (nothing to extract, one add
// per threshold to update); level 2 and the leaves stay packed (8 bytes: a third of the shared-memory traffic of
// unpacked nodes, which was measured too -- faster only with a single warp per SM).
struct LatTree {
    Quad *l1;            // [4][lanes]   thresholds of the level-1 nodes
    uint64_t *l2;        // [16][lanes]  packed
    uint64_t *l3;        // [64][lanes]  leaves
    uint32_t lanes;      // 32 on the device (lane-interleaved), 1 on the host
};

// The root's thresholds and a register copy of the four level-1 nodes: the copy is reloaded right after the
// step's level-1 store, a whole step before it is needed (the loads may not move above that store, and
// issued at the top of the next step they were not back when the level-0 decision wanted them).
struct TopLevels {
    uint32_t T0, T1, T2;
    Quad c[4];
};

GPUAR_HD void lat_tree_init(TopLevels &top, const LatTree &tr)
{
    top.T0 = 64, top.T1 = 128, top.T2 = 192;
    for (uint32_t q = 0; q < 4; ++q) top.c[q] = Quad{16u, 32u, 48u, 0u};
    for (uint32_t q = 0; q < 4; ++q) tr.l1[q * tr.lanes] = Quad{16u, 32u, 48u, 0u};
    for (uint32_t q = 0; q < 16; ++q) tr.l2[q * tr.lanes] = tree_node_init(4);
    for (uint32_t q = 0; q < 64; ++q) tr.l3[q * tr.lanes] = kLeafInit;
}

// one of four by the three "threshold above the target" predicates of a level (p0 implies p1 implies p2):
// child 0 iff p0, child 1 iff p1 & !p0, child 2 iff p2 & !p1, child 3 iff !p2
GPUAR_HD uint32_t pick4p(bool p0, bool p1, bool p2, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return p1 ? (p0 ? a : b) : (p2 ? c : d);
}
GPUAR_HD uint32_t neg_bit(uint32_t d) { return d >> 31; }

GPUAR_HD void narrow_track_products(DecState &st, uint32_t lo_r, uint32_t hi_r, uint32_t m, uint32_t sh, BitSource &in)
{
    const uint32_t qa = div_total(hi_r, m, sh);
    const uint32_t qb = div_total(lo_r, m, sh);
    const uint32_t E = width_exponent(qa, qb);
    const uint32_t A = funnel_r_wrap((st.L + qb) << 16, 0u, E);
    const uint32_t R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    const uint32_t sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15;
    st.R = R1 >> sx;
    st.L = (A >> sx) & 0x7FFFu;
    const uint32_t t = 16u - (E & 31u) - sx;
    st.D = funnel_l(in.hi, st.D - qb, t) & 0xFFFFu;
    in.skip(t);
}

GPUAR_HD uint32_t decode_step_latency(DecState &st, TopLevels &top, const LatTree &tr, uint32_t T, uint32_t m, uint32_t sh,
                                   BitSource &in)
{
    const uint32_t lanes = tr.lanes;
    uint32_t &T0 = top.T0, &T1 = top.T1, &T2 = top.T2;
    const uint32_t num = st.D * T + (T - 1u);
    const uint32_t nr = 0u - st.R;
    const Quad qa1 = top.c[0], qb1 = top.c[1], qc1 = top.c[2], qd1 = top.c[3];
    // level 0
    const uint32_t e0 = T0 * nr + num, e1 = T1 * nr + num, e2 = T2 * nr + num;      // num - threshold * range
    const bool p0 = (int32_t)e0 < 0, p1 = (int32_t)e1 < 0, p2 = (int32_t)e2 < 0;
    const uint32_t c0 = 3u - (neg_bit(e0) + neg_bit(e1) + neg_bit(e2));
    uint64_t *const g2 = tr.l2 + c0 * 4u * lanes;
    const uint64_t a2 = g2[0], b2 = g2[lanes], c2n = g2[2u * lanes], d2 = g2[3u * lanes];
    Quad n1;
    n1.x = pick4p(p0, p1, p2, qa1.x, qb1.x, qc1.x, qd1.x);
    n1.y = pick4p(p0, p1, p2, qa1.y, qb1.y, qc1.y, qd1.y);
    n1.z = pick4p(p0, p1, p2, qa1.z, qb1.z, qc1.z, qd1.z);
    n1.w = 0u;
    const uint32_t num1 = pick4p(p0, p1, p2, num, e0, e1, e2);
    T0 += neg_bit(e0), T1 += neg_bit(e1), T2 += neg_bit(e2);
    // level 1
    const uint32_t f0 = n1.x * nr + num1, f1 = n1.y * nr + num1, f2 = n1.z * nr + num1;
    const bool k0 = (int32_t)f0 < 0, k1 = (int32_t)f1 < 0, k2 = (int32_t)f2 < 0;
    const uint32_t c1 = 3u - (neg_bit(f0) + neg_bit(f1) + neg_bit(f2));
    uint32_t idx = c0 * 4u + c1;
    uint64_t *const g3 = tr.l3 + idx * 4u * lanes;
    const uint64_t a3 = g3[0], b3 = g3[lanes], c3n = g3[2u * lanes], d3 = g3[3u * lanes];
    uint32_t lo2 = pick4p(k0, k1, k2, (uint32_t)a2, (uint32_t)b2, (uint32_t)c2n, (uint32_t)d2);
    uint32_t hi2 = pick4p(k0, k1, k2, (uint32_t)(a2 >> 32), (uint32_t)(b2 >> 32), (uint32_t)(c2n >> 32), (uint32_t)(d2 >> 32));
    const uint32_t num2 = pick4p(k0, k1, k2, num1, f0, f1, f2);
    n1.x += neg_bit(f0), n1.y += neg_bit(f1), n1.z += neg_bit(f2);
    tr.l1[c0 * lanes] = n1;
    top.c[0] = tr.l1[0], top.c[1] = tr.l1[lanes], top.c[2] = tr.l1[2u * lanes], top.c[3] = tr.l1[3u * lanes];
    // level 2: packed (0, t0, t1, t2)
    const uint32_t g0 = (lo2 >> 16) * nr + num2, g1 = (hi2 & 0xFFFFu) * nr + num2, g2v = (hi2 >> 16) * nr + num2;
    const bool j0 = (int32_t)g0 < 0, j1 = (int32_t)g1 < 0, j2 = (int32_t)g2v < 0;
    const uint32_t c2 = 3u - (neg_bit(g0) + neg_bit(g1) + neg_bit(g2v));
    uint32_t lo3 = pick4p(j0, j1, j2, (uint32_t)a3, (uint32_t)b3, (uint32_t)c3n, (uint32_t)d3);
    uint32_t hi3 = pick4p(j0, j1, j2, (uint32_t)(a3 >> 32), (uint32_t)(b3 >> 32), (uint32_t)(c3n >> 32), (uint32_t)(d3 >> 32));
    const uint32_t num3 = pick4p(j0, j1, j2, num2, g0, g1, g2v);
    lo2 = mad32(neg_bit(g0), 0x10000u, lo2);
    hi2 = mad32(neg_bit(g2v), 0x10000u, hi2) + neg_bit(g1);
    g2[c1 * lanes] = ((uint64_t)hi2 << 32) | lo2;
    idx = idx * 4u + c2;
    // the leaf: packed inclusive sums (s0, s1, s2, s3)
    const uint32_t h0 = (lo3 & 0xFFFFu) * nr + num3, h1 = (lo3 >> 16) * nr + num3, h2 = (hi3 & 0xFFFFu) * nr + num3,
                   h3 = (hi3 >> 16) * nr + num3;
    const bool r0 = (int32_t)h0 < 0, r1 = (int32_t)h1 < 0, r2 = (int32_t)h2 < 0;
    const uint32_t c3 = 3u - (neg_bit(h0) + neg_bit(h1) + neg_bit(h2));
    const uint32_t num_lo = pick4p(r0, r1, r2, num3, h0, h1, h2);     // num - cum[s] * range
    const uint32_t num_hi = pick4p(r0, r1, r2, h0, h1, h2, h3);       // num - cum[s+1] * range
    lo3 = mad32(neg_bit(h1), 0x10000u, lo3) + neg_bit(h0);
    hi3 = hi3 + 0x10000u + neg_bit(h2);
    g3[c2 * lanes] = ((uint64_t)hi3 << 32) | lo3;
    narrow_track_products(st, num - num_lo, num - num_hi, m, sh, in);
    return idx * 4u + c3;
}

}  // namespace gpuar
