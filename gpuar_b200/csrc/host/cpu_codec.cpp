#include "cpu_codec.hpp"

#include <cstring>
#include <vector>

#include "../coder_math.h"

namespace gip {

using namespace gpuar;

std::uint32_t cpuEncodePacket(const std::uint8_t *in, std::uint32_t n, std::uint8_t *slot)
{
    std::uint64_t tree[kTreeStored], root;
    enc_tree_init(root, tree, 1);
    std::uint32_t L = 0, V = 0, pend = 0;
    BitSink out;
    out.acc = 0;
    out.nb = 0;
    out.widx = 0;
    out.wcap = (kSlot - kHdr) >> 2;
    out.words = reinterpret_cast<std::uint32_t *>(slot + kHdr);
    for (std::uint32_t i = 0; i < n; ++i) {
        std::uint32_t sh, lo, cnt, k, u, U1;
        const std::uint32_t m = magic_for(256u + i, sh);
        tree_encode(root, tree, 1, in[i], lo, cnt);
        narrow_renorm(L, V, lo, lo + cnt, m, sh, k, u, U1);
        emit_symbol(out, pend, k, u, U1);
    }
    return finish_packet(out, L, pend, slot, n);
}

std::uint32_t cpuDecodePacket(const std::uint8_t *pkt, std::uint8_t *out)
{
    std::uint64_t tree[kTreeStored], root;
    tree_init(root, tree, 1);
    const std::uint32_t raw = (std::uint32_t)pkt[2] | ((std::uint32_t)pkt[3] << 8);
    const std::uint8_t *p = pkt + kHdr;
    auto word = [&]() {                                   // next 4 stream bytes, first byte = most significant
        const std::uint32_t w = ((std::uint32_t)p[0] << 24) | ((std::uint32_t)p[1] << 16) | ((std::uint32_t)p[2] << 8) | p[3];
        p += 4;
        return w;
    };
    BitSource in;
    const std::uint64_t w0 = word(), w1 = word();
    in.start((w0 << 32) | w1, 64u);
    std::uint32_t code = in.take(16u);
    if (in.hungry()) in.feed(word());
    std::uint32_t L = 0, V = 0;
    for (std::uint32_t i = 0; i < raw && i < kPacket; ++i) {
        const std::uint32_t T = 256u + i;
        std::uint32_t sh, lo, cnt, k, u, U1;
        const std::uint32_t m = magic_for(T, sh);
        const std::uint32_t s = tree_decode(root, tree, 1, unscale(code, L, V, T), T, lo, cnt);
        out[i] = (std::uint8_t)s;
        narrow_renorm(L, V, lo, lo + cnt, m, sh, k, u, U1);
        code = advance_code(code, k, u, in);
        if (in.hungry()) in.feed(word());
    }
    return raw < kPacket ? raw : kPacket;
}

}  // namespace gip
