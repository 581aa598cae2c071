/*
 * gpuar_oracle.c -- CPU restatement of the GPUAR packet codec and .gip framing.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gpuar_b200/ (the product) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker.  Parity status: PINNED -- validated bit-for-bit against the
 * reference's own arCompress/arDecompress compiled from /root/reference (see
 * oracle/Makefile target `ref`, tests/golden/make_golden.py) and against the
 * golden vectors committed under tests/golden/.
 *
 * The algorithm is an adaptive order-0 arithmetic coder with 16-bit bounds
 * (reference: src/gpuar_kernel.cu).  This file follows the reference's
 * loop structure (bit-at-a-time renormalisation) on purpose: the CUDA product
 * uses closed forms, so agreement between the two is a real cross-check.
 * The only liberty taken is the model container: the reference keeps a
 * Fenwick tree (gpuar_kernel.cu:205-238); here the same cumulative counts live
 * in a flat array cum[0..256], cum[k] = sum of counts of symbols < k.
 *
 * Build: gcc -O3 -shared -fPIC oracle/gpuar_oracle.c -o oracle/libgpuar_oracle.so
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_HDR 4u            /* packet header bytes, gpu.h:14 */
#define ORC_NSYM 256u
#define ORC_MSB 0x8000u       /* MASK_BIT(0), gpuar.h:26 */
#define ORC_MSB2 0x4000u      /* MASK_BIT(1) */

/* ---- MSB-first bit writer (gpuar_kernel.cu:76-84,128-151,430-439) ---- */
typedef struct {
    uint8_t *p;
    unsigned acc;   /* pending bits, left-aligned as they arrive */
    unsigned n;     /* number of pending bits, 0..7 */
} orc_bw;

static void bw_put(orc_bw *w, unsigned bit)
{
    w->acc = (w->acc << 1) | (bit & 1u);
    if (++w->n == 8) {
        *w->p++ = (uint8_t)w->acc;
        w->acc = 0;
        w->n = 0;
    }
}

static void bw_close(orc_bw *w)   /* zero-pad the last byte, :430-439 */
{
    if (w->n) {
        *w->p++ = (uint8_t)(w->acc << (8 - w->n));
        w->acc = 0;
        w->n = 0;
    }
}

/* ---- MSB-first bit reader (gpuar_kernel.cu:533-569) ---- */
typedef struct {
    const uint8_t *p;
    unsigned cur;
    unsigned n;
} orc_br;

static unsigned br_get(orc_br *r)
{
    if (r->n == 0) {
        r->cur = *r->p++;
        r->n = 8;
    }
    r->n--;
    return (r->cur >> r->n) & 1u;
}

/* ---- model: all counts 1, total 256 (gpuar_kernel.cu:403-419) ---- */
static void model_init(uint16_t cum[ORC_NSYM + 1])
{
    for (unsigned k = 0; k <= ORC_NSYM; k++) cum[k] = (uint16_t)k;
}

/* count[s]++ (gpuar_kernel.cu:229-238 via :288) */
static void model_bump(uint16_t cum[ORC_NSYM + 1], unsigned s)
{
    for (unsigned k = s + 1; k <= ORC_NSYM; k++) cum[k]++;
}

/* interval narrowing, gpuar_kernel.cu:256-288.  total is the running
 * cumulativeProb (256 + symbols coded so far). */
static void narrow(uint16_t *lo, uint16_t *hi, const uint16_t cum[ORC_NSYM + 1],
                   unsigned s, unsigned total)
{
    uint32_t range = (uint32_t)(uint16_t)(*hi - *lo) + 1u;              /* :269 */
    uint32_t up = ((uint32_t)cum[s + 1] * range) / total;               /* :272-273 */
    uint32_t dn = ((uint32_t)cum[s] * range) / total;                   /* :279-280 */
    *hi = (uint16_t)(*lo + (uint16_t)up - 1u);                          /* :276 */
    *lo = (uint16_t)(*lo + (uint16_t)dn);                               /* :283 */
}

/*
 * Encode one packet of n (<= 16127) bytes.  out must hold n + 512 bytes.
 * Returns compLen (header included).  Mirrors arCompress, gpuar_kernel.cu:487-531.
 */
size_t gpuar_oracle_encode_packet(const uint8_t *in, size_t n, uint8_t *out)
{
    uint16_t cum[ORC_NSYM + 1];
    uint16_t lo = 0, hi = 0xFFFFu;
    unsigned pending = 0;
    unsigned total = ORC_NSYM;
    orc_bw w = { out + ORC_HDR, 0, 0 };

    model_init(cum);
    for (size_t i = 0; i < n; i++) {
        unsigned s = in[i];
        narrow(&lo, &hi, cum, s, total);
        total++;                                                        /* :286 */
        model_bump(cum, s);
        for (;;) {                                                      /* :321-367 */
            if (((hi ^ lo) & ORC_MSB) == 0) {
                unsigned b = (hi & ORC_MSB) != 0;
                bw_put(&w, b);
                while (pending) { bw_put(&w, !b); pending--; }
            } else if ((lo & ORC_MSB2) && !(hi & ORC_MSB2)) {
                pending++;
                lo &= (uint16_t)~(ORC_MSB | ORC_MSB2);
                hi |= ORC_MSB2;
            } else {
                break;
            }
            lo = (uint16_t)(lo << 1);
            hi = (uint16_t)((hi << 1) | 1u);
        }
    }
    {                                                                   /* :379-388 */
        unsigned b = (lo & ORC_MSB2) != 0;
        bw_put(&w, b);
        for (pending++; pending; pending--) bw_put(&w, !b);
    }
    bw_close(&w);
    size_t len = (size_t)(w.p - out);
    out[0] = (uint8_t)len;          out[1] = (uint8_t)(len >> 8);       /* :527 */
    out[2] = (uint8_t)n;            out[3] = (uint8_t)(n >> 8);         /* :528 */
    return len;
}

/*
 * Decode one packet.  pkt must be readable 2 bytes past compLen (the decoder
 * pre-reads 16 bits and shifts in one bit per bit the encoder shifted out).
 * Returns the number of bytes produced.  Mirrors arDecompress, :848-892.
 */
size_t gpuar_oracle_decode_packet(const uint8_t *pkt, uint8_t *out)
{
    uint16_t cum[ORC_NSYM + 1];
    uint16_t lo = 0, hi = 0xFFFFu, code = 0;
    unsigned total = ORC_NSYM;
    size_t n = (size_t)pkt[2] | ((size_t)pkt[3] << 8);                  /* :859 */
    orc_br r = { pkt + ORC_HDR, 0, 0 };

    model_init(cum);
    for (int i = 0; i < 16; i++) code = (uint16_t)((code << 1) | br_get(&r));  /* :582-603 */

    for (size_t i = 0; i < n; i++) {
        uint32_t range = (uint32_t)(uint16_t)(hi - lo) + 1u;            /* :708 */
        uint32_t t = (uint32_t)(uint16_t)(code - lo) + 1u;              /* :711 */
        uint16_t target = (uint16_t)((t * total - 1u) / range);         /* :712-715 */
        /* symbol with cum[s] <= target < cum[s+1] (:727-763) */
        if (target >= cum[ORC_NSYM]) return i;                          /* :875-879 */
        unsigned s = 0, b = ORC_NSYM;       /* counts >= 1, so cum[] is strictly increasing */
        while (b - s > 1) {
            unsigned m = (s + b) >> 1;
            if (cum[m] <= target) s = m; else b = m;
        }
        out[i] = (uint8_t)s;
        narrow(&lo, &hi, cum, s, total);
        total++;
        model_bump(cum, s);
        for (;;) {                                                      /* :787-836 */
            if (((hi ^ lo) & ORC_MSB) == 0) {
                /* shift the matching MSB out */
            } else if ((lo & ORC_MSB2) && !(hi & ORC_MSB2)) {
                lo &= (uint16_t)~(ORC_MSB | ORC_MSB2);
                hi |= ORC_MSB2;
                code ^= ORC_MSB2;
            } else {
                break;
            }
            lo = (uint16_t)(lo << 1);
            hi = (uint16_t)((hi << 1) | 1u);
            code = (uint16_t)((code << 1) | br_get(&r));
        }
    }
    return n;
}

/*
 * Payload (= .gip bytes from offset 20): packets of packet_bytes input bytes,
 * tightly concatenated (cpu_compressor.cpp:144-173).  payload must hold
 * ceil(n/packet_bytes) * (packet_bytes + 512) bytes.  Returns payload length.
 * packet_bytes = 8192 is the reference (gpu.h:12-13).
 */
size_t gpuar_oracle_encode_stream(const uint8_t *in, size_t n, uint8_t *payload,
                                  size_t packet_bytes)
{
    size_t pos = 0;
    for (size_t off = 0; off < n; off += packet_bytes) {
        size_t m = n - off < packet_bytes ? n - off : packet_bytes;
        pos += gpuar_oracle_encode_packet(in + off, m, payload + pos);
    }
    return pos;
}

/*
 * Walk the packet chain (cpu_compressor.cpp:47-78).  payload must be padded
 * with >= 2 readable bytes past c.  Returns bytes produced, or (size_t)-1 on a
 * malformed chain or if out_cap would be exceeded.
 */
size_t gpuar_oracle_decode_stream(const uint8_t *payload, size_t c, uint8_t *out,
                                  size_t out_cap)
{
    size_t pos = 0, produced = 0;
    while (pos < c) {
        if (c - pos < ORC_HDR) return (size_t)-1;
        size_t len = (size_t)payload[pos] | ((size_t)payload[pos + 1] << 8);
        size_t raw = (size_t)payload[pos + 2] | ((size_t)payload[pos + 3] << 8);
        if (len <= ORC_HDR || len > c - pos || raw > out_cap - produced) return (size_t)-1;
        size_t got = gpuar_oracle_decode_packet(payload + pos, out + produced);
        if (got != raw) return (size_t)-1;
        produced += got;
        pos += len;
    }
    return produced;
}

/* Packet offsets by chain walk; returns packet count or (size_t)-1. */
size_t gpuar_oracle_index(const uint8_t *payload, size_t c, uint64_t *offsets, size_t cap)
{
    size_t pos = 0, k = 0;
    while (pos < c) {
        if (c - pos < ORC_HDR) return (size_t)-1;
        size_t len = (size_t)payload[pos] | ((size_t)payload[pos + 1] << 8);
        if (len <= ORC_HDR || len > c - pos) return (size_t)-1;
        if (k < cap) offsets[k] = pos;
        k++;
        pos += len;
    }
    return k;
}

/*
 * 20-byte .gip header as the reference DEFINES it (file_header.hpp:19-36,61-72):
 * bytes 0-2 version 0.1.0, bytes 4-7 uncompressed size u32 LE, bytes 12-15
 * total .gip size u32 LE.  Bytes 3, 8-11, 16-19 are never written by the
 * reference (stack garbage); this writer zeroes them.
 */
void gpuar_oracle_write_header(uint8_t hdr[20], uint64_t raw_bytes, uint64_t gip_bytes)
{
    memset(hdr, 0, 20);
    hdr[0] = 0; hdr[1] = 1; hdr[2] = 0;
    for (int k = 0; k < 4; k++) {
        hdr[4 + k] = (uint8_t)(raw_bytes >> (8 * k));
        hdr[12 + k] = (uint8_t)(gip_bytes >> (8 * k));
    }
}
