#include "cpu_codec.hpp"

#include <cstring>
#include <vector>

#include "../coder_math.h"
#include "../decode_math.h"
#include "../encode_math.h"

namespace gip {

using namespace gpuar;

// The explicit CPU mode of the command line (--host; the reference's CPUCompressor, src/cpu_compressor.cpp): the
// same closed-form steps the kernels run (encode_math.h / decode_math.h), one packet at a time.
std::uint32_t cpuEncodePacket(const std::uint8_t *in, std::uint32_t n, std::uint8_t *slot)
{
    std::uint64_t tree[kTreeStored], root;
    enc_tree_init(root, tree, 1);
    EncState st{0u, 65536u};
    CarrySink out;
    out.start(reinterpret_cast<std::uint32_t *>(slot + kHdr), (kSlot - kHdr) >> 2);
    for (std::uint32_t i = 0; i < n; ++i) {
        std::uint32_t sh, lo, cnt, inc, t;
        const std::uint32_t m = magic_for(256u + i, sh);
        tree_encode(root, tree, 1, in[i], lo, cnt);
        narrow_plain(st, lo, lo + cnt, m, sh, inc, t);
        const std::uint32_t stored = out.widx;
        if (out.push(inc, t)) out.carry_into_stored(stored);
    }
    return finish_packet_plain(out, st.Lp, slot, n);
}

std::uint32_t cpuDecodePacket(const std::uint8_t *pkt, std::uint8_t *out)
{
    std::uint64_t tree[kTreeStored], root;
    dec_tree_init(root, tree, 1);
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
    DecState st;
    st.D = in.take(16u);
    st.L = 0;
    st.R = 65536u;
    if (in.hungry()) in.feed(word());
    for (std::uint32_t i = 0; i < raw && i < kPacket; ++i) {
        const std::uint32_t T = 256u + i;
        std::uint32_t sh;
        const std::uint32_t m = magic_for(T, sh);
        out[i] = (std::uint8_t)decode_step(st, root, tree, 1, T, m, sh, in);
        if (in.hungry()) in.feed(word());
    }
    return raw < kPacket ? raw : kPacket;
}

}  // namespace gip
