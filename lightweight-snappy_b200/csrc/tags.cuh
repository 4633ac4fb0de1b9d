// tags.cuh -- Snappy element headers, shared by the decode kernels.
//
// reference: tag dispatch of decompressor src/snappy_decompression.c:290-333, literal length
// forms of do_literal :193-224.
#pragma once

#include "common.cuh"

namespace sb200 {

struct Header {
    uint32_t hdr;  // header bytes (tag + extra)
    uint32_t len;  // output bytes
    uint32_t info; // literal: stream position of its bytes; copy: offset
    bool is_lit;
    bool slow; // header does not fit in the 4 bytes of v
};

// Decodes the element whose first 4 stream bytes are v and whose tag sits at stream position pos.
__device__ __forceinline__ Header decode_header(uint32_t v, uint32_t pos)
{
    Header h;
    const uint32_t tag = v & 0xffu;
    const uint32_t type = tag & 3u;
    h.slow = false;
    h.is_lit = type == 0;
    if (type == 0) {
        const uint32_t m = tag >> 2;
        if (m < 60) {
            h.hdr = 1;
            h.len = m + 1;
        } else {
            const uint32_t k = m - 59; // 1..4 length bytes
            h.hdr = 1 + k;
            h.slow = k == 4;
            h.len = ((v >> 8) & (0xffffffu >> (8 * (3 - min(k, 3u))))) + 1;
        }
        h.info = pos + h.hdr;
    } else if (type == 1) {
        h.hdr = 2;
        h.len = ((tag >> 2) & 7u) + 4;
        h.info = ((tag >> 5) << 8) | ((v >> 8) & 0xffu);
    } else if (type == 2) {
        h.hdr = 3;
        h.len = (tag >> 2) + 1;
        h.info = (v >> 8) & 0xffffu;
    } else {
        h.hdr = 5;
        h.len = (tag >> 2) + 1;
        h.info = 0;
        h.slow = true;
    }
    return h;
}

// Tag byte -> header facts, looked up instead of branched over:
//   bits 0-2 header bytes, bits 3-9 output length when the tag alone gives it (0: a literal whose
//   length follows in 1..4 bytes), bit 10 literal, bit 11 header longer than 4 bytes (copy-4, 4-byte
//   literal length), bit 12 copy with a 1-byte offset.
constexpr uint32_t kTagLit = 1u << 10, kTagSlow = 1u << 11, kTagCopy1 = 1u << 12;

__device__ __forceinline__ uint32_t tag_facts(uint32_t tag)
{
    const uint32_t type = tag & 3u, m = tag >> 2;
    if (type == 0) {
        if (m < 60)
            return 1u | ((m + 1u) << 3) | kTagLit;
        const uint32_t k = m - 59u; // length bytes
        return (1u + k) | kTagLit | (k == 4 ? kTagSlow : 0u);
    }
    if (type == 1)
        return 2u | (((m & 7u) + 4u) << 3) | kTagCopy1;
    if (type == 2)
        return 3u | ((m + 1u) << 3);
    return 5u | ((m + 1u) << 3) | kTagSlow;
}

// Element facts from the table (v = the 4 stream bytes at pos).  When `slow` is set the fifth
// header byte is missing from v: the caller completes len (literal) or info (copy-4).
__device__ __forceinline__ Header decode_header_lut(const uint16_t *__restrict__ lut, uint32_t v, uint32_t pos)
{
    Header h;
    const uint32_t f = lut[v & 0xffu];
    h.hdr = f & 7u;
    h.len = (f >> 3) & 127u;
    h.is_lit = f & kTagLit;
    h.slow = f & kTagSlow;
    if (h.len == 0) // literal with 1..4 length bytes (the fourth one does not fit v: the caller's business)
        h.len = ((v >> 8) & (0xffffffu >> (8u * (4u - min(h.hdr, 4u))))) + 1u;
    const uint32_t c1 = ((v << 3) & 0x700u) | ((v >> 8) & 0xffu), c2 = (v >> 8) & 0xffffu;
    h.info = h.is_lit ? pos + h.hdr : ((f & kTagCopy1) ? c1 : c2);
    return h;
}

// Is the block [c0, c1) of the stream exactly one literal of olen bytes (incompressible data)?
__device__ __forceinline__ bool one_literal_block(const uint8_t *__restrict__ stream, uint64_t c0, uint64_t c1, uint32_t olen)
{
    const uint8_t *p = stream + c0;
    const uint32_t tag = __ldg(p);
    if ((tag & 3u) != 0 || (tag >> 2) < 60u)
        return false;
    const uint32_t k = (tag >> 2) - 59u; // length bytes
    if (c1 - c0 <= 1 + k)
        return false;
    uint64_t raw = 0;
    for (uint32_t i = 0; i < k; ++i)
        raw |= (uint64_t)__ldg(p + 1 + i) << (8u * i);
    return raw + 1 == olen && c1 - c0 == 1ull + k + olen;
}

} // namespace sb200
