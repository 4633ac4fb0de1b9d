// decode_lane.cu -- K4/K5: block decode, ONE LANE per 64 KiB output block.
//
// reference: decompressor src/snappy_decompression.c:290-333 (tag dispatch), do_literal :193-224,
// write_literal :232-239, do_copy :253-265, write_copy :273-280.
//
// The copy graph of a Snappy block is deep (a low-entropy block is one chain of ~2000 dependent levels,
// tools/dep_model.py), so spreading the elements of a block over the lanes of a warp leaves most lanes
// idle most of the time: the warp-per-block decoder (decode.cu), the CTA-per-block tile decoder
// (decode_tile.cu) and the group-per-block window decoder (decode_win.cu) all spend 3 G warp instructions
// or more per GiB (profiles/r02_*).  Here the parallelism is across blocks only: every lane walks its own
// block from the first element to the last, the way the reference does, and the 32 lanes of a warp are
// 32 independent blocks.  All lanes run the same straight-line step, so nothing diverges:
//
//   front     decodes one element header per step from a 64-byte window of the stream kept in shared
//             memory (refilled with 16-byte loads), pushes {length, offset | literal position} into a
//             per-lane queue of 8 entries and, for a copy that reaches back beyond the lane's output
//             ring, prefetches the source line into L1.  Headers depend only on the stream, so the
//             front runs ahead of the byte moves and the L2 latency of far sources is hidden.
//   back      moves up to 8 bytes of the current element per step into the lane's output ring (512
//             bytes of shared memory): from the ring itself (near copies; a self-overlapping copy,
//             write_copy :273-280, is replicated with a distance that doubles as the periodic region
//             grows), from the stream (literals) or from output already written to HBM (far copies),
//             both fetched with two aligned 8-byte loads.
//   flush     every finished 16 bytes of the ring go to HBM with one 16-byte store.
// All byte-granular work happens in shared memory; global memory only sees 8/16-byte accesses.  Only
// the block offsets are needed (from K0 or from the caller's side index), not K0's element maps.
// Blocks that are a single literal (incompressible data) are moved by k_copy_literal_blocks
// (decode_win.cu) with plain 16-byte vector copies and skipped here.
// Unlike the reference, malformed input is detected and reported in *status instead of being undefined
// behaviour.
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tags.cuh"

namespace sb200 {

namespace {

constexpr uint32_t kRing = 512;            // output ring per lane (bytes, power of two)
constexpr uint32_t kNear = kRing - 80;     // copies with offset <= kNear read the ring
constexpr uint32_t kQueue = 8;             // decoded-but-not-executed elements per lane
constexpr uint32_t kSbuf = 64;             // stream window per lane: four 16-byte chunks
constexpr uint32_t kLaneWords = (kRing + kSbuf + kQueue * 8) / 4 + 1; // +1: odd stride, lanes start in different banks
constexpr uint32_t kLutBytes = 512;

__global__ void __launch_bounds__(256) k_copy_literal_blocks(const uint8_t *__restrict__ stream, uint64_t body_offset,
                                                             const uint64_t *__restrict__ offsets, uint64_t total_out,
                                                             uint8_t *out_base, const uint32_t *__restrict__ status,
                                                             uint64_t n_blocks)
{
    const uint64_t blk = blockIdx.x;
    if (*reinterpret_cast<const volatile uint32_t *>(status) != 0)
        return;
    const uint64_t c0 = offsets[blk], c1 = offsets[blk + 1];
    if (c1 <= c0 || c0 < body_offset || c1 > offsets[n_blocks] || c1 - c0 > 2u * kBlock)
        return; // k_decode_lane reports it
    const uint64_t oleft = total_out - blk * (uint64_t)kBlock;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    if (!one_literal_block(stream, c0, c1, olen))
        return;
    const uint32_t hdr = ((uint32_t)__ldg(stream + c0) >> 2) - 58u; // tag + length bytes
    coop_copy_ro(out_base + blk * (uint64_t)kBlock, stream + c0 + hdr, olen, threadIdx.x, 256);
}

__device__ __forceinline__ uint64_t shr64_bytes(uint64_t lo, uint64_t hi, uint32_t bytes)
{
    // bytes [bytes, bytes + 8) of the 16-byte little-endian value hi:lo (bytes in 0..7)
    const uint32_t s = bytes * 8u;
    return s ? (lo >> s) | (hi << (64u - s)) : lo;
}

__global__ void __launch_bounds__(32) k_decode_lane(const uint8_t *__restrict__ stream, uint64_t body_offset,
                                                    const uint64_t *__restrict__ offsets, uint64_t total_out,
                                                    uint8_t *out_base, uint32_t *__restrict__ status,
                                                    uint64_t n_blocks, uint64_t blk_base, int skip_literal_blocks)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem);
    const uint32_t lane = threadIdx.x;
    uint32_t *mine = reinterpret_cast<uint32_t *>(smem + kLutBytes) + lane * kLaneWords;
    uint32_t *ringw = mine;                                             // kRing / 4 words
    uint8_t *ring = reinterpret_cast<uint8_t *>(ringw);
    uint32_t *sbufw = mine + kRing / 4;                                 // 16 words
    uint8_t *sbuf = reinterpret_cast<uint8_t *>(sbufw);
    uint32_t *queuew = mine + kRing / 4 + kSbuf / 4;                     // kQueue entries of two words
    for (uint32_t i = lane; i < 256; i += 32)
        lut[i] = (uint16_t)tag_facts(i);
    __syncwarp();
    if (*reinterpret_cast<volatile uint32_t *>(status) != 0)
        return; // an earlier stage rejected the stream: the offsets are not trustworthy

    const uint64_t blk = blockIdx.x * 32ull + lane;
    bool live = blk < n_blocks;
    uint32_t err = 0;
    uint64_t c0 = body_offset, c1 = body_offset + 1;
    const uint64_t stream_bytes = offsets[n_blocks];
    if (live) {
        c0 = offsets[blk], c1 = offsets[blk + 1];
        if (c1 <= c0 || c0 < body_offset || c1 > stream_bytes || c1 - c0 > 2u * kBlock) {
            err = SNAPPY_B200_ST_CORRUPT;
            live = false;
            c0 = body_offset, c1 = body_offset + 1;
        }
    }
    const uint64_t oleft = live ? total_out - blk * (uint64_t)kBlock : 0;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    if (live && skip_literal_blocks && one_literal_block(stream, c0, c1, olen))
        live = false; // k_copy_literal_blocks moves it
    const uint32_t clen = (uint32_t)(c1 - c0);
    const uint8_t *__restrict__ in = stream + c0; // stream positions below are relative to the block
    uint8_t *out = out_base + blk * (uint64_t)kBlock;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out_base) & 15u) == 0;
    // the stream window: 16-byte chunks of the aligned address space; position p lives at q = p + mis
    const uint32_t mis = (uint32_t)reinterpret_cast<uintptr_t>(in) & 15u;
    const uint4 *gchunk = reinterpret_cast<const uint4 *>(in - mis);
    const uint32_t last_chunk = (mis + clen - 1) >> 4; // last chunk that holds a byte of the block

    // front
    uint32_t ip = 0, fop = 0, qhead = 0, nchunk = 0;
    bool fdone = !live;
    // back
    uint32_t qtail = 0, op = 0, flushed = 0, rem = 0, cinfo = 0, cdist = 0, cdone = 0;
    bool clit = false;

    for (;;) {
        const bool busy = !fdone || qtail != qhead || rem != 0 || op - flushed >= 16u;
        if (!__any_sync(kFull, busy))
            break;
        // ------------------------------------------------------------------ front: one header
        {
            const uint32_t q = ip + mis;
            if (!fdone && (q >> 4) > nchunk)
                nchunk = q >> 4; // the payload of a literal was skipped: those chunks are never needed
            if (!fdone && nchunk * 16u < q + 8u && nchunk <= last_chunk) {
                const uint4 cv = __ldg(gchunk + nchunk);
                uint32_t *slot = sbufw + (nchunk & 3u) * 4u;
                slot[0] = cv.x, slot[1] = cv.y, slot[2] = cv.z, slot[3] = cv.w;
                ++nchunk;
            }
            const bool can = !fdone && qhead - qtail < kQueue && (nchunk * 16u >= q + 8u || nchunk > last_chunk);
            if (can) {
                const uint32_t wq = q >> 2;
                const uint32_t v = __funnelshift_r(sbufw[wq & 15u], sbufw[(wq + 1u) & 15u], (q & 3u) * 8u);
                const Header h = decode_header_lut(lut, v, ip);
                uint32_t len = h.len, info = h.info;
                if (h.slow) { // the fifth header byte: top of a 4-byte literal length / of a copy-4 offset
                    const uint32_t top = (uint32_t)sbuf[(q + 4u) & 63u] << 24;
                    if (h.is_lit)
                        len = ((v >> 8) | top) + 1u; // (0xffffffff + 1 wraps to 0: rejected below)
                    else
                        info = (v >> 8) | top;
                }
                uint32_t e = 0;
                if (len == 0 || ip + h.hdr > clen || (h.is_lit && (uint64_t)ip + h.hdr + len > clen) ||
                    (!h.is_lit && info == 0))
                    e = SNAPPY_B200_ST_CORRUPT;
                else if (len > olen - fop)
                    e = SNAPPY_B200_ST_FRAMING;
                else if (!h.is_lit && info > fop) // reaches back into an earlier block
                    e = (uint64_t)info > (blk_base + blk) * (uint64_t)kBlock + fop ? SNAPPY_B200_ST_CORRUPT
                                                                                   : SNAPPY_B200_ST_FRAMING;
                if (e) {
                    err |= e;
                    fdone = true;
                } else {
                    uint32_t *qe = queuew + (qhead & (kQueue - 1u)) * 2u;
                    qe[0] = info;
                    qe[1] = len | (h.is_lit ? 0x80000000u : 0u);
                    ++qhead;
                    if (!h.is_lit && info > kNear)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(out + (fop - info)));
                    ip += h.hdr + (h.is_lit ? len : 0u);
                    fop += len;
                    if (ip >= clen) {
                        fdone = true;
                        if (fop != olen)
                            err |= SNAPPY_B200_ST_CORRUPT;
                    }
                }
            }
        }
        // ------------------------------------------------------------------ back: up to 8 bytes
        if (rem == 0 && qtail != qhead) {
            const uint32_t *qe = queuew + (qtail & (kQueue - 1u)) * 2u;
            cinfo = qe[0];
            const uint32_t l = qe[1];
            rem = l & 0x7fffffffu;
            clit = l >> 31;
            cdist = cinfo;
            cdone = 0;
            ++qtail;
        }
        if (rem != 0) {
            const bool near = !clit && cinfo <= kNear;
            uint32_t n = min(8u, rem);
            if (near)
                n = min(n, cdist); // a step never outruns the distance to its source
            uint64_t v;
            if (near) {
                const uint32_t r = (op - cdist) & (kRing - 1u);
                const uint32_t w = r >> 2;
                const uint32_t w0 = ringw[w], w1 = ringw[(w + 1u) & (kRing / 4u - 1u)], w2 = ringw[(w + 2u) & (kRing / 4u - 1u)];
                v = shr64_bytes((uint64_t)w0 | ((uint64_t)w1 << 32), (uint64_t)w2, r & 3u);
            } else {
                // the stream (literal) or output that is already in HBM (far copy): two aligned 8-byte loads
                const uint8_t *a = clit ? in + cinfo : out + (op - cinfo);
                const uint32_t sa = (uint32_t)reinterpret_cast<uintptr_t>(a) & 7u;
                const uint64_t *aw = reinterpret_cast<const uint64_t *>(a - sa);
                const uint64_t lo = aw[0];
                const uint64_t hi = sa + n > 8u ? aw[1] : 0ull;
                v = shr64_bytes(lo, hi, sa);
            }
            const uint32_t vl = (uint32_t)v, vh = (uint32_t)(v >> 32);
            const uint32_t d0 = op;
            if (n > 0)
                ring[(d0 + 0u) & (kRing - 1u)] = (uint8_t)vl;
            if (n > 1)
                ring[(d0 + 1u) & (kRing - 1u)] = (uint8_t)(vl >> 8);
            if (n > 2)
                ring[(d0 + 2u) & (kRing - 1u)] = (uint8_t)(vl >> 16);
            if (n > 3)
                ring[(d0 + 3u) & (kRing - 1u)] = (uint8_t)(vl >> 24);
            if (n > 4)
                ring[(d0 + 4u) & (kRing - 1u)] = (uint8_t)vh;
            if (n > 5)
                ring[(d0 + 5u) & (kRing - 1u)] = (uint8_t)(vh >> 8);
            if (n > 6)
                ring[(d0 + 6u) & (kRing - 1u)] = (uint8_t)(vh >> 16);
            if (n > 7)
                ring[(d0 + 7u) & (kRing - 1u)] = (uint8_t)(vh >> 24);
            op += n;
            rem -= n;
            cdone += n;
            if (clit)
                cinfo += n;
            else if (near && 2u * cdist <= cdone + cinfo)
                cdist <<= 1; // the periodic region has doubled (a multiple of the offset: same bytes)
        }
        // ------------------------------------------------------------------ flush: one finished 16 bytes
        if (op - flushed >= 16u) {
            const uint32_t w = (flushed & (kRing - 1u)) >> 2;
            const uint4 cv = make_uint4(ringw[w], ringw[w + 1], ringw[w + 2], ringw[w + 3]);
            if (vec_ok) {
                *reinterpret_cast<uint4 *>(out + flushed) = cv;
            } else {
                const uint32_t x[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    out[flushed + i] = (uint8_t)(x[i >> 2] >> (8 * (i & 3)));
            }
            flushed += 16u;
        }
    }
    // the tail of the block (fewer than 16 bytes)
    for (uint32_t i = flushed; i < op; ++i)
        out[i] = ring[i & (kRing - 1u)];
    if (live && !err && op != olen)
        err = SNAPPY_B200_ST_CORRUPT;
    if (err)
        atomicOr(status, err);
}

} // namespace

cudaError_t launch_copy_literal_blocks(const uint8_t *d_stream, uint64_t body_offset, const uint64_t *d_offsets,
                                       uint64_t n_blocks, uint64_t total_out, uint8_t *d_out, const uint32_t *d_status,
                                       cudaStream_t st, uint64_t *launches)
{
    if (n_blocks == 0)
        return cudaSuccess;
    k_copy_literal_blocks<<<(unsigned)n_blocks, 256, 0, st>>>(d_stream, body_offset, d_offsets, total_out, d_out, d_status,
                                                             n_blocks);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_decode_lane(const uint8_t *d_stream, uint64_t body_offset, const uint64_t *d_offsets, uint64_t n_blocks,
                               uint64_t total_out, uint8_t *d_out, uint32_t *d_status, uint64_t blk_base,
                               bool skip_literal_blocks, cudaStream_t st, uint64_t *launches)
{
    if (n_blocks == 0)
        return cudaSuccess;
    if (n_blocks > 0x7fffffffull)
        return cudaErrorInvalidValue;
    const size_t smem = kLutBytes + 32u * kLaneWords * 4u;
    if (skip_literal_blocks) {
        k_copy_literal_blocks<<<(unsigned)n_blocks, 256, 0, st>>>(d_stream, body_offset, d_offsets, total_out, d_out,
                                                                 d_status, n_blocks);
        *launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    k_decode_lane<<<(unsigned)((n_blocks + 31) / 32), 32, smem, st>>>(d_stream, body_offset, d_offsets, total_out, d_out,
                                                                     d_status, n_blocks, blk_base,
                                                                     skip_literal_blocks ? 1 : 0);
    *launches += 1;
    return cudaGetLastError();
}

} // namespace sb200
