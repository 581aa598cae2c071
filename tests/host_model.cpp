// host_model.cpp -- CPU emulation of the CUDA kernels' data flow, built with g++ from the
// SAME arithmetic source the kernels use (gpuar_b200/csrc/coder_math.h, encode_math.h,
// decode_math.h).  Test infrastructure: it lets `pytest -m "not gpu"` check the closed forms
// (reciprocal division, single-normalisation renormalisation, the carry-propagating bit output,
// the 4-ary model tree with its packed compares and byte permutes, the float-estimated divide,
// the multiplicative tree descent) against the oracle without a GPU.  All kernels are
// lane = packet, so one lane is emulated at a time.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../gpuar_b200/csrc/coder_math.h"
#include "../gpuar_b200/csrc/decode_math.h"
#include "../gpuar_b200/csrc/encode_math.h"

using namespace gpuar;

extern "C" {

// the same packet through the stages of encode_ws_kernel: three model warps (levels 0-1, 2, 3), CODER on the
// plain window with the lazily normalised range, one word per step to BITS
uint32_t host_model_encode_packet_ws(const uint8_t *x, uint32_t n, uint8_t *slot, uint32_t slot_bytes)
{
    std::vector<uint64_t> tree(kTreeStored);
    uint64_t root;
    enc_tree_init(root, tree.data(), 1);
    uint32_t Lp = 0, R = 65536, sx = 0;
    CarrySink out;
    out.start(reinterpret_cast<uint32_t *>(slot + kHdr), (slot_bytes - kHdr) >> 2);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t sh;
        const uint32_t m = magic_for(256u + i, sh);
        uint32_t cnt;
        // the model warps take their symbols four at a time from one input word
        uint32_t word = 0;
        memcpy(&word, x + (i & ~3u), (n - (i & ~3u)) < 4u ? (n - (i & ~3u)) : 4u);
        const uint32_t pa = tree_encode_upper_word(root, tree.data(), 1, word_fields_upper(word), i & 3u);
        const uint32_t pb = tree_encode_mid_word(tree.data(), 1, word_fields_mid(word), i & 3u);
        const uint32_t lo_d = tree_encode_leaf_word(tree.data(), 1, word_fields_leaf(word), i & 3u, cnt);
        const uint32_t pd = cnt * 65536u + lo_d;
        const uint32_t lo = pa + pb + (pd & 0xFFFFu);
        const uint32_t c = narrow_plain_lazy(Lp, R, sx, lo, lo + (pd >> 16), m, sh);   // ring C entry
        uint32_t inc, t;
        step_unpack(c, inc, t);
        const uint32_t before = out.widx;
        if (out.push(inc, t)) out.carry_into_stored(before);
    }
    return finish_packet_plain(out, Lp, slot, n);
}

// how often host_model_encode_packet had to carry into words it had already stored (the rare path of the
// kernels): lets the tests assert that their long-underflow inputs really exercise it
static uint64_t g_carry_events = 0;
uint64_t host_model_carry_events(void) { return g_carry_events; }

// encode one packet exactly as a lane of encode_kernel does (encode_math.h: plain window, carry into the
// pending bits); returns compLen
uint32_t host_model_encode_packet(const uint8_t *x, uint32_t n, uint8_t *slot, uint32_t slot_bytes)
{
    std::vector<uint64_t> tree(kTreeStored);
    uint64_t root;
    enc_tree_init(root, tree.data(), 1);
    EncState st{0u, 65536u};
    CarrySink out;
    out.start(reinterpret_cast<uint32_t *>(slot + kHdr), (slot_bytes - kHdr) >> 2);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t sh;
        const uint32_t m = magic_for(256u + i, sh);
        uint32_t lo, cnt, inc, t;
        tree_encode(root, tree.data(), 1, x[i], lo, cnt);
        narrow_plain(st, lo, lo + cnt, m, sh, inc, t);
        const uint32_t before = out.widx;
        if (out.push(inc, t)) {
            out.carry_into_stored(before);
            ++g_carry_events;
        }
    }
    return finish_packet_plain(out, st.Lp, slot, n);
}

static size_t encode_stream_with(uint32_t (*enc)(const uint8_t *, uint32_t, uint8_t *, uint32_t), const uint8_t *in,
                                 size_t n, uint8_t *payload, uint32_t packet)
{
    std::vector<uint8_t> slot(packet + 512 + 16);
    size_t pos = 0;
    for (size_t off = 0; off < n; off += packet) {
        const uint32_t m = (uint32_t)(n - off < packet ? n - off : packet);
        const uint32_t len = enc(in + off, m, slot.data(), packet + 512);
        memcpy(payload + pos, slot.data(), len);
        pos += len;
    }
    return pos;
}

size_t host_model_encode_stream_ws(const uint8_t *in, size_t n, uint8_t *payload, uint32_t packet)
{
    return encode_stream_with(host_model_encode_packet_ws, in, n, payload, packet);
}

size_t host_model_encode_stream(const uint8_t *in, size_t n, uint8_t *payload, uint32_t packet)
{
    return encode_stream_with(host_model_encode_packet, in, n, payload, packet);
}

// decode the packet at byte offset `off` of a padded, 4-byte aligned payload exactly as a lane of
// decode_kernel does (decode_math.h); variant 0 = throughput (decode_step), 1 = latency
// (decode_step_latency); returns bytes produced
uint32_t host_model_decode_packet(const uint8_t *payload, size_t readable, size_t off, uint8_t *out, int variant)
{
    std::vector<uint64_t> tree(kTreeStored);
    std::vector<Quad> l1(4);
    uint64_t root = 0;
    TopLevels top;
    LatTree tr{l1.data(), tree.data() + 4, tree.data() + 20, 1u};
    if (variant == 0) dec_tree_init(root, tree.data(), 1);
    else lat_tree_init(top, tr);
    const uint32_t *const words = reinterpret_cast<const uint32_t *>(payload);
    const uint32_t *const wend = words + (readable >> 2) - 1;
    auto word = [&](const uint32_t *p) { return bswap32(*(p < wend ? p : wend)); };
    const uint32_t raw = (uint32_t)payload[off + 2] | ((uint32_t)payload[off + 3] << 8);
    const size_t sp = off + kHdr;
    const uint32_t *wp = words + (sp >> 2);
    const uint32_t skip = 8u * (uint32_t)(sp & 3u);
    BitSource in;
    const uint64_t w0 = word(wp);
    ++wp;
    const uint64_t w1 = word(wp);
    ++wp;
    in.start(((w0 << 32) | w1) << skip, 64u - skip);
    uint32_t ahead = word(wp);
    DecState st;
    st.D = in.take(16u);
    st.L = 0;
    st.R = 65536u;
    if (in.hungry()) { in.feed(ahead); ++wp; ahead = word(wp); }
    for (uint32_t i = 0; i < raw; ++i) {
        const uint32_t T = 256u + i;
        uint32_t sh;
        const uint32_t m = magic_for(T, sh);
        const uint32_t s = variant == 0 ? decode_step(st, root, tree.data(), 1, T, m, sh, in)
                                        : decode_step_latency(st, top, tr, T, m, sh, in);
        out[i] = (uint8_t)s;
        // the kernels refill every second step (a step takes at most 16 bits, a refill restores 33)
        if ((i & 1u) && in.hungry()) { in.feed(ahead); ++wp; ahead = word(wp); }
    }
    return raw;
}

// exhaustive-ish check of the reciprocal division: for every total T and for numerators
// around every multiple of T up to max_n (plus random ones): returns the number of mismatches
uint64_t host_model_check_division(uint32_t max_n, uint32_t packet)
{
    uint64_t bad = 0;
    for (uint32_t T = 256; T < 256 + packet; ++T) {
        uint32_t sh;
        const uint32_t m = magic_for(T, sh);
        for (uint64_t q = 0; q * T <= max_n; q += 1 + q / 64) {
            for (int d = -1; d <= 1; ++d) {
                const int64_t n = (int64_t)(q * T) + d;
                if (n < 0 || n > max_n) continue;
                bad += div_total((uint32_t)n, m, sh) != (uint32_t)n / T;
            }
        }
        bad += div_total(max_n, m, sh) != max_n / T;
    }
    return bad;
}

// The steps of the kernels against the reference's bit-at-a-time loops (gpuar_kernel.cu:256-288, 321-367,
// 787-836) on random reachable states: returns the number of mismatches.  Checked per state:
//   narrow_renorm (first generation: k, u, both bounds), narrow_plain and narrow_plain_lazy (encoders: plain
//   window with and without pending underflow bits, total shift, new range, the bits that leave the window),
//   narrow_track and narrow_track_products (decoders: code - lower bound, lower bound, range, bit window).
uint64_t host_model_check_renorm(uint64_t seed, uint32_t count, uint32_t packet)
{
    uint64_t bad = 0, x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (uint32_t n = 0; n < count; ++n) {
        // a renormalised state: MSBs differ and not the 01../10.. pattern
        uint32_t lo16, hi16;
        do {
            lo16 = (uint32_t)rnd() & 0x7FFFu;
            hi16 = ((uint32_t)rnd() & 0x7FFFu) | 0x8000u;
        } while ((lo16 & 0x4000u) && !(hi16 & 0x4000u));
        const uint32_t T = 256u + (uint32_t)(rnd() % packet);
        uint32_t cl = (uint32_t)(rnd() % T), ch = cl + 1u + (uint32_t)(rnd() % (T - cl));
        if (n & 1u) ch = cl + 1u;                                  // narrow symbols stress the shifts
        if (ch > T) ch = T;
        // reference arithmetic, encoder and decoder side (the decoder's code lies inside the symbol's interval)
        const uint32_t range = hi16 - lo16 + 1u;
        uint16_t U = (uint16_t)(lo16 + (uint16_t)(ch * range / T) - 1u), Lr = (uint16_t)(lo16 + (uint16_t)(cl * range / T));
        const uint16_t U1ref = U, L1ref = Lr;
        const uint32_t code0 = L1ref + (uint32_t)(rnd() % ((uint32_t)U1ref - L1ref + 1u));
        const uint64_t window = rnd();                             // the next 64 stream bits
        uint16_t code = (uint16_t)code0;
        uint32_t kk = 0, uu = 0, taken = 0;
        for (;;) {
            if (((U ^ Lr) & 0x8000u) == 0) { ++kk; }
            else if ((Lr & 0x4000u) && !(U & 0x4000u)) { ++uu; Lr &= 0x3FFFu; U |= 0x4000u; code ^= 0x4000u; }
            else break;
            Lr = (uint16_t)(Lr << 1);
            U = (uint16_t)((U << 1) | 1u);
            code = (uint16_t)((code << 1) | (uint32_t)((window >> (63u - taken)) & 1u));
            ++taken;
        }
        const uint32_t new_range = (uint32_t)U - Lr + 1u;
        uint32_t sh;
        const uint32_t m = magic_for(T, sh);
        {   // first generation
            uint32_t L = lo16, V = (~hi16) & 0xFFFFu, k, u, U1;
            narrow_renorm(L, V, cl, ch, m, sh, k, u, U1);
            bad += (L != Lr) || ((V ^ 0xFFFFu) != U) || (k != kk) || (u != uu) || (U1 != U1ref);
        }
        for (uint32_t pf = 0; pf < 2u; ++pf) {                     // encoders: without / with pending underflow bits
            const uint32_t X = (pf << 15) + L1ref;
            const uint32_t want_top = uu ? 1u : kk ? 0u : pf;
            EncState st{lo16 | (pf << 15), range};
            uint32_t inc, t;
            narrow_plain(st, cl, ch, m, sh, inc, t);
            bad += (t != kk + uu) || (st.R != new_range) || ((st.Lp & 0x7FFFu) != Lr) || ((st.Lp >> 15) != want_top) ||
                   (inc != X >> (16u - t));
            for (uint32_t pre = 0; pre < 2u; ++pre) {              // the range held as is (sx = 0) or doubled (sx = 1)
                uint32_t Lp = lo16 | (pf << 15), R1 = range << pre, sx = pre, inc2, t2;
                step_unpack(narrow_plain_lazy(Lp, R1, sx, cl, ch, m, sh), inc2, t2);
                bad += (inc2 != inc) || (t2 != t) || (Lp != st.Lp) || ((R1 >> sx) != new_range) || (sx && (R1 & 1u)) ||
                       R1 <= 32768u || R1 > 65536u;
            }
        }
        for (uint32_t products = 0; products < 2u; ++products) {   // decoders
            DecState st{(code0 - lo16) & 0xFFFFu, lo16, range};
            BitSource in;
            in.start(window, 64u);
            if (products) narrow_track_products(st, cl * range, ch * range, m, sh, in);
            else narrow_track(st, cl, ch, m, sh, in);
            const uint64_t left = taken ? window << taken : window;
            bad += (st.D != (((uint32_t)code - Lr) & 0xFFFFu)) || (st.L != Lr) || (st.R != new_range) ||
                   (in.have != 64u - taken) || (in.hi != (uint32_t)(left >> 32)) || (in.lo != (uint32_t)left);
        }
    }
    return bad;
}

// the float-estimated divide of the decoder against integer division, on a lattice of states
uint64_t host_model_check_unscale(uint32_t stride, uint32_t packet)
{
    uint64_t bad = 0;
    for (uint32_t T = 256; T < 256 + packet; T += 37)
        for (uint32_t range = 16385; range <= 65536; range += stride)
            for (uint32_t cl = 0; cl < range; cl += 1 + range / 97) {
                const uint32_t L = 0, V = 65536u - range, code = cl;
                const uint32_t want = ((cl + 1u) * T - 1u) / range;
                bad += unscale(code, L, V, T) != want;
                bad += divide_floor((cl + 1u) * T - 1u, range) != want;
            }
    return bad;
}

}  // extern "C"
