// decode.cu -- K4/K5 (first version): one warp per 64 KiB output block, elements in order.
//
// reference: decompressor src/snappy_decompression.c:290-333 (tag dispatch), do_literal
// :193-224, write_literal :232-239, do_copy :253-265, write_copy :273-280.
//
// The warp keeps a 32-byte window of the compressed block in registers (one byte per lane),
// decodes the element at its head with shuffles, and moves the bytes cooperatively: literal
// bytes straight from the stream, copy bytes from the output written earlier.  A copy whose
// offset is smaller than its length repeats the pattern (byte i comes from i mod offset), so
// every source byte lies before the element and one __syncwarp() per element orders the
// read-after-write through global memory.  Unlike the reference, malformed input is detected
// and reported in *status instead of being undefined behaviour.
#include "common.cuh"

namespace sb200 {

__global__ void __launch_bounds__(32) k_decode_warp(const uint8_t *__restrict__ stream,
                                                    const uint64_t *__restrict__ offsets, uint64_t total_out,
                                                    uint8_t *out_base, uint32_t *__restrict__ status)
{
    const uint32_t lane = threadIdx.x;
    const uint64_t blk = blockIdx.x;
    if (*reinterpret_cast<volatile uint32_t *>(status) != 0)
        return; // an earlier stage (K0) rejected the stream: the offsets are not trustworthy
    const uint64_t c0 = offsets[blk], c1 = offsets[blk + 1];
    const uint8_t *__restrict__ in = stream + c0;
    const uint64_t clen64 = c1 - c0;
    uint8_t *out = out_base + blk * (uint64_t)kBlock;
    const uint64_t oleft = total_out - blk * (uint64_t)kBlock;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    if (c1 < c0 || clen64 > 2u * kBlock) { // a 64 KiB block never needs more than 65536+1010 bytes
        if (lane == 0)
            atomicOr(status, SNAPPY_B200_ST_CORRUPT);
        return;
    }
    const uint32_t clen = (uint32_t)clen64;

    uint32_t ip = 0, op = 0;
    uint32_t err = 0;
    while (op < olen) {
        if (ip >= clen) {
            err = SNAPPY_B200_ST_CORRUPT;
            break;
        }
        const uint32_t wb = (ip + lane < clen) ? (uint32_t)__ldg(in + ip + lane) : 0u;
        const uint32_t tag = __shfl_sync(kFull, wb, 0);
        const uint32_t b1 = __shfl_sync(kFull, wb, 1), b2 = __shfl_sync(kFull, wb, 2);
        const uint32_t b3 = __shfl_sync(kFull, wb, 3), b4 = __shfl_sync(kFull, wb, 4);
        const uint32_t type = tag & 3u;
        if (type == 0) { // literal
            uint32_t m = tag >> 2, hdr = 1;
            if (m >= 60) {
                const uint32_t k = m - 59;
                hdr = 1 + k;
                const uint32_t raw = b1 | (b2 << 8) | (b3 << 16) | (b4 << 24);
                m = k == 4 ? raw : raw & ((1u << (8 * k)) - 1u);
            }
            if (m >= olen - op || (uint64_t)ip + hdr + m + 1 > clen) {
                // m + 1 > room left in the block, or the literal runs past the compressed block
                err = (m >= olen - op && (uint64_t)ip + hdr + m + 1 <= clen) ? SNAPPY_B200_ST_FRAMING
                                                                              : SNAPPY_B200_ST_CORRUPT;
                break;
            }
            const uint32_t len = m + 1;
            if (hdr + len <= 32) { // whole literal already sits in the window
                if (lane >= hdr && lane < hdr + len)
                    out[op + lane - hdr] = (uint8_t)wb;
            } else {
                coop_copy_ro(out + op, in + ip + hdr, len, lane, 32);
            }
            ip += hdr + len;
            op += len;
        } else {
            uint32_t len, off, hdr;
            if (type == 1) {
                len = ((tag >> 2) & 7u) + 4;
                off = ((tag >> 5) << 8) | b1;
                hdr = 2;
            } else if (type == 2) {
                len = (tag >> 2) + 1;
                off = b1 | (b2 << 8);
                hdr = 3;
            } else {
                len = (tag >> 2) + 1;
                off = b1 | (b2 << 8) | (b3 << 16) | (b4 << 24);
                hdr = 5;
            }
            if (ip + hdr > clen || off == 0) {
                err = SNAPPY_B200_ST_CORRUPT;
                break;
            }
            if (off > op || len > olen - op) {
                // source before this block, or the copy crosses the block end: legal raw Snappy
                // only if the stream was not framed in independent 64 KiB blocks
                err = (off > op && (uint64_t)off > blk * (uint64_t)kBlock + op) ? SNAPPY_B200_ST_CORRUPT
                                                                               : SNAPPY_B200_ST_FRAMING;
                break;
            }
            __syncwarp(); // earlier elements' stores are visible to every lane from here
            const uint8_t *src = out + op - off;
            for (uint32_t i = lane; i < len; i += 32) {
                const uint32_t s = off >= len ? i : i % off;
                out[op + i] = src[s];
            }
            __syncwarp();
            ip += hdr;
            op += len;
        }
    }
    if (!err && ip != clen)
        err = SNAPPY_B200_ST_CORRUPT; // the index said this block ends at c1
    if (err && lane == 0)
        atomicOr(status, err);
}

cudaError_t launch_decode(const uint8_t *d_stream, const uint64_t *d_offsets, uint64_t n_blocks, uint64_t total_out,
                          uint8_t *d_out, uint32_t *d_status, cudaStream_t st, uint64_t *launches)
{
    if (n_blocks == 0)
        return cudaSuccess;
    if (n_blocks > 0x7fffffffull)
        return cudaErrorInvalidValue;
    k_decode_warp<<<(unsigned)n_blocks, 32, 0, st>>>(d_stream, d_offsets, total_out, d_out, d_status);
    *launches += 1;
    return cudaGetLastError();
}

} // namespace sb200
