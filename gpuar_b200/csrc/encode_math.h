// encode_math.h -- the encoder's coder + bit-output step, second generation (round 2), shared verbatim by
// encode.cu / encode_ws.cu (device) and tests/host_model.cpp (host, g++) like coder_math.h.
//
// Same interval arithmetic as the reference (applySymbolRange src/gpuar_kernel.cu:256-288) and the same bit
// stream as its writeEncodedBits / writeRemaining (:321-388), but the stream is produced the way a
// carry-propagating range coder does it instead of with the reference's pending-underflow counter:
//
//   The reference keeps a 16-bit window (lower, upper) of the interval.  An "underflow" shift (lower = 01..,
//   upper = 10..) drops the SECOND bit of both and counts it (`pend`); when the first bits finally agree on b it
//   writes b followed by `pend` copies of !b.  Read as integers that is: the bits that left the window were
//   tentatively 0 1 1 .. 1, and a decision b = 1 adds one to them (0 1 1 1 -> 1 0 0 0), a decision b = 0 leaves
//   them.  So the emitted stream IS the binary expansion of the lower bound kept as one long integer: shift the
//   window plainly (no bit dropped) and let the addition of a symbol's offset carry out of the window into the
//   bits already shifted out.  The plain window Lp is the reference's `lower` with bit 15 set while pend > 0;
//   the interval width -- all that the subdivision depends on -- is the same.
//
//   Per symbol:  X = Lp + qb  (17 bits: bit 16 is the carry),  t = total shift (single normalisation, as the
//   decoder),  out = (out << t) + (X >> (16 - t)),  Lp = (X << t) mod 2^16.
//   Neither k (matching-MSB shifts) nor u (underflow shifts) nor pend exists any more: no count-leading-zeros,
//   no field assembly, no "long underflow" path.  End of packet (:379-388: bit 14 of lower, then pend + 1
//   inverted copies) = add 0x4000 to the window and shift out two bits.
//
//   The carry can run through every bit that is still pending.  The sink keeps at least 16 bits below each word
//   it stores, so a carry reaches an already stored word only through 16 one-bits (what the reference calls
//   pend >= 16): the kernels handle that under one warp-uniform vote by incrementing the stored words in place.
#pragma once
#include "coder_math.h"

namespace gpuar {

// a store the compiler must keep as ONE predicated instruction (a branch here would diverge on most steps)
GPUAR_HD void store_word_if(bool on, uint32_t *dst, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.u32 [%1], %2;\n\t}" ::"r"((uint32_t)on), "l"(dst), "r"(v)
                 : "memory");
#else
    if (on) *dst = v;
#endif
}

struct EncState {
    uint32_t Lp;         // plain 16-bit window of the lower bound (the reference's lower | 0x8000 while underflow bits are pending)
    uint32_t R;          // range = upper - lower + 1
};

// Interval narrowing with a single normalisation (derivation in coder_math.h) on the plain window.
//   out: inc = the t bits that leave the window plus the carry (t + 1 bits), t = total shift (0..16)
GPUAR_HD void narrow_plain(EncState &st, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh, uint32_t &inc, uint32_t &t)
{
    const uint32_t qa = div_total(hi * st.R, m, sh);
    const uint32_t qb = div_total(lo * st.R, m, sh);
    const uint32_t E = width_exponent(qa, qb);                    // = 16 - s1 (mod 32)
    const uint32_t X = st.Lp + qb;
    const uint32_t A = funnel_r_wrap(X << 16, 0u, E);             // (X mod 2^16) << s1
    const uint32_t R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    const uint32_t sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15;     // one shift less (coder_math.h)
    st.R = R1 >> sx;
    st.Lp = (A >> sx) & 0xFFFFu;
    const uint32_t so = (E & 31u) + sx;                           // 16 - t
    t = 16u - so;
    inc = X >> so;
}

// The same step for the CODER warp of encode_ws.cu, which carries the range unnormalised by one halving
// (state R1, sx with range = R1 >> sx: floor(c * (R1 >> sx) / T) = floor(c * R1 / T) >> sx, R1 being even whenever
// sx = 1, so the next step's multiplies start from R1 while the decision sx is still being computed).  The step's output travels to the BITS warp as one word:
// inc in bits 0..16, t in bits 20..24 (handing X, E and sx over raw and letting BITS work out the shift was
// measured too: BITS then paces the kernel, profiles/r2_encode_v2.md).  Zero is a step that moves nothing.
//   state: Lp, R1, sx.   Start: Lp = 0, R1 = 65536, sx = 0.
constexpr uint32_t kStepNone = 0u;
GPUAR_HD uint32_t narrow_plain_lazy(uint32_t &Lp, uint32_t &R1, uint32_t &sx, uint32_t lo, uint32_t hi, uint32_t m, uint32_t sh)
{
    const uint32_t qa = (mulhi32(hi * R1, m) >> sh) >> sx;
    const uint32_t qb = (mulhi32(lo * R1, m) >> sh) >> sx;
    const uint32_t E = width_exponent(qa, qb);
    const uint32_t X = Lp + qb;
    const uint32_t A = funnel_r_wrap(X << 16, 0u, E);             // (X mod 2^16) << s1
    R1 = funnel_r_wrap((qa - qb) << 16, 0u, E);
    sx = ((A & 0x7FFFu) + R1 - 0x8001u) >> 15;
    Lp = (A >> sx) & 0xFFFFu;
    const uint32_t so = (E & 31u) + sx;                           // 16 - t
    return (X >> so) + ((16u - so) << 20);
}
GPUAR_HD void step_unpack(uint32_t w, uint32_t &inc, uint32_t &t)
{
    inc = w & 0xFFFFFu;
    t = w >> 20;
}

// Bit sink with carry: `acc` holds the nb newest bits of the stream as a number, plus possibly one carry bit
// above them.  A 32-bit word is stored as soon as 48 bits are pending, so 16..47 stay behind.
struct CarrySink {
    uint64_t acc;
    uint32_t nb;
    uint32_t widx;       // next word of the slot's bitstream
    uint32_t wcap;       // writable words
    uint32_t *words;     // first bitstream word of the slot (4-byte aligned)

    GPUAR_HD void start(uint32_t *w, uint32_t cap)
    {
        acc = 0;
        nb = 0;
        widx = 0;
        wcap = cap;
        words = w;
    }
    // append t bits (inc < 2^(t+1): its top bit is a carry into the bits before).  Returns true if a carry left
    // the sink: the caller must then call carry_into_stored(widx_before) -- rare, see the header.  Branch free: the
    // lanes of a warp complete their words at different steps, so the store is a predicated instruction and the
    // counters move by selects.
    GPUAR_HD bool push(uint32_t inc, uint32_t t)
    {
        acc = (acc << t) + inc;
        nb += t;
        const bool full = nb >= 48u;
        const uint32_t keep = nb & 31u;                           // bits that stay behind when a word goes out (16..31)
        const uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
        const uint32_t w = funnel_r_wrap(lo, hi, keep);           // the oldest 32 bits ...
        const bool carry = full && (hi >> keep) != 0u;            // ... and a carry above them
        store_word_if(full && widx < wcap, words + widx, bswap32(w));
        const uint32_t kept = lo & ~(0xFFFFFFFFu << keep);
        acc = full ? (uint64_t)kept : acc;
        widx += full ? 1u : 0u;
        nb = full ? keep : nb;
        return carry;
    }
    // +1 on the stream that ends just before word `upto` (words 0 .. upto-1 are stored, big-endian)
    GPUAR_HD void carry_into_stored(uint32_t upto)
    {
        uint32_t p = upto < wcap ? upto : wcap;
        while (p > 0u) {
            --p;
            const uint32_t v = bswap32(words[p]) + 1u;
            words[p] = bswap32(v);
            if (v != 0u) break;
        }
    }
};

// End of packet: the two closing bits (gpuar_kernel.cu:379-388), zero padding to a byte (:430-439), the tail
// bytes and the 4-byte packet header (:525-528) at `slot` (out.words == slot + 4).  Returns compLen.
GPUAR_HD uint32_t finish_packet_plain(CarrySink &out, uint32_t Lp, uint8_t *slot, uint32_t raw_len)
{
    const uint32_t X = Lp + 0x4000u;
    out.acc = (out.acc << 2) + (X >> 14);
    out.nb += 2u;                                                 // <= 49
    if (out.acc >> out.nb) {                                      // the last carry
        out.carry_into_stored(out.widx);
        out.acc &= ~(~0ull << out.nb);
    }
    uint32_t bytes = out.widx * 4u;
    const uint32_t tail = (out.nb + 7u) >> 3;                     // <= 7
    const uint64_t left = out.acc << (64u - out.nb);              // left-aligned, zero padded (nb >= 2)
    uint8_t *bp = reinterpret_cast<uint8_t *>(out.words + out.widx);
    const uint32_t room = out.widx < out.wcap ? (out.wcap - out.widx) * 4u : 0u;
    for (uint32_t b = 0; b < tail && b < room; ++b) bp[b] = (uint8_t)(left >> (56u - 8u * b));
    bytes += tail;
    const uint32_t comp = bytes + kHdr;
    *reinterpret_cast<uint32_t *>(slot) = (comp & 0xFFFFu) | (raw_len << 16);   // u16 compLen | u16 rawLen, LE
    return comp;
}

}  // namespace gpuar
