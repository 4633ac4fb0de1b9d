// decode.cu -- K4/K5: one warp per 64 KiB output block, 32 stream bytes per step.
//
// reference: decompressor src/snappy_decompression.c:290-333 (tag dispatch), do_literal
// :193-224, write_literal :232-239, do_copy :253-265, write_copy :273-280.
//
// The reference walks one element at a time.  Here a warp looks at a 32-byte window of the
// compressed block per step:
//   1. every lane decodes the byte at its offset as if an element started there (tag, header
//      length, output length, literal source / copy offset);
//   2. the lanes that really are element starts are the orbit of lane 0 under
//      "next = lane + element size"; it is found with pointer doubling over shuffles and five
//      warp-wide OR reductions instead of a serial walk;
//   3. a warp prefix sum of the output lengths of those lanes gives every element its output
//      offset;
//   4. the output bytes of the whole window are then produced 32 at a time, one byte per lane:
//      a lane finds its element from a bit map of the element starts inside the round, then
//      reads either the stream (literal) or the output written earlier (copy; offset < length
//      repeats the pattern, write_copy :273-280).  A source byte that is produced by the very
//      same round is followed back through the round's elements until it leaves the round or
//      lands in a literal.
// Literals of 64 bytes or more are moved by the 16-byte copy loop, and the rare element kinds
// whose header does not fit the 4 bytes a lane holds (copy-4, 4-byte literal length) take a
// one-element path.  Copies read the output through global memory; one __syncwarp() per step
// orders the read-after-write.  Unlike the reference, malformed input is detected and
// reported in *status instead of being undefined behaviour.
#include "common.cuh"

namespace sb200 {

constexpr uint32_t kLongLiteral = 64;

// Little-endian 32 bits at an arbitrarily aligned address; the aligned words touched are
// clamped to `last_word`, the last aligned word that still holds a byte of the stream.
__device__ __forceinline__ uint32_t ld_le32_any(const uint8_t *__restrict__ p, const uint32_t *__restrict__ last_word)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const uint32_t *w1 = w0 + 1;
    w0 = w0 > last_word ? last_word : w0;
    w1 = w1 > last_word ? last_word : w1;
    return __funnelshift_r(__ldg(w0), __ldg(w1), (uint32_t)(a & 3u) * 8u);
}

__global__ void __launch_bounds__(32) k_decode_warp(const uint8_t *__restrict__ stream,
                                                    const uint64_t *__restrict__ offsets, uint64_t total_out,
                                                    uint8_t *out_base, uint32_t *__restrict__ status)
{
    const uint32_t lane = threadIdx.x;
    const uint64_t blk = blockIdx.x;
    if (*reinterpret_cast<volatile uint32_t *>(status) != 0)
        return; // an earlier stage (K0) rejected the stream: the offsets are not trustworthy
    const uint64_t c0 = offsets[blk], c1 = offsets[blk + 1];
    const uint64_t stream_bytes = offsets[gridDim.x]; // entry n_blocks of the index = end of the stream
    const uint64_t clen64 = c1 - c0;
    if (c1 < c0 || c1 > stream_bytes || clen64 > 2u * kBlock) { // a block never needs more than 65536+1010
        if (lane == 0)
            atomicOr(status, SNAPPY_B200_ST_CORRUPT);
        return;
    }
    const uint8_t *__restrict__ in = stream + c0;
    const uint32_t *last_word = reinterpret_cast<const uint32_t *>(
        reinterpret_cast<uintptr_t>(stream + stream_bytes - 1) & ~uintptr_t(3));
    uint8_t *out = out_base + blk * (uint64_t)kBlock;
    const uint64_t oleft = total_out - blk * (uint64_t)kBlock;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    const uint32_t clen = (uint32_t)clen64;

    __shared__ uint4 elems[32]; // the elements of the current window: {start, len, literal|info, 2^16/offset}

    uint32_t ip = 0, op = 0;
    uint32_t err = 0;
    while (op < olen) {
        if (ip >= clen) {
            err = SNAPPY_B200_ST_CORRUPT;
            break;
        }
        // ---- 1. speculative decode of the element that would start at ip + lane
        const uint32_t pos = ip + lane;
        const bool inside = pos < clen;
        const uint32_t v = ld_le32_any(in + (inside ? pos : clen - 1), last_word);
        const uint32_t tag = v & 0xffu;
        const uint32_t type = tag & 3u;
        uint32_t hdr, len, info; // info: literal -> stream position of its bytes, copy -> offset
        bool slow = false;       // header does not fit in v (needs the one-element path)
        if (type == 0) {
            const uint32_t m = tag >> 2;
            if (m < 60) {
                hdr = 1;
                len = m + 1;
            } else {
                const uint32_t k = m - 59; // 1..4 length bytes
                hdr = 1 + k;
                slow = k == 4;
                len = ((v >> 8) & (0xffffffu >> (8 * (3 - min(k, 3u))))) + 1;
            }
            info = pos + hdr;
        } else if (type == 1) {
            hdr = 2;
            len = ((tag >> 2) & 7u) + 4;
            info = ((tag >> 5) << 8) | ((v >> 8) & 0xffu);
        } else if (type == 2) {
            hdr = 3;
            len = (tag >> 2) + 1;
            info = (v >> 8) & 0xffffu;
        } else {
            hdr = 5;
            len = (tag >> 2) + 1;
            info = 0;
            slow = true;
        }
        const bool is_lit = type == 0;
        const uint32_t size = hdr + (is_lit ? len : 0u); // stream bytes of the element

        // ---- 2. which lanes are element starts: orbit of lane 0 under lane -> lane + size
        // (an element that ends at or past the end of the compressed block leaves the window)
        uint32_t j0 = (inside && pos + size < clen) ? min(lane + size, 32u) : 32u;
        uint32_t j1 = __shfl_sync(kFull, j0, j0 & 31);
        j1 = j0 < 32 ? j1 : 32u;
        uint32_t j2 = __shfl_sync(kFull, j1, j1 & 31);
        j2 = j1 < 32 ? j2 : 32u;
        uint32_t j3 = __shfl_sync(kFull, j2, j2 & 31);
        j3 = j2 < 32 ? j3 : 32u;
        uint32_t j4 = __shfl_sync(kFull, j3, j3 & 31);
        j4 = j3 < 32 ? j4 : 32u;
        unsigned M = 1u;
        M |= __reduce_or_sync(kFull, ((M >> lane) & 1u) && j4 < 32 ? 1u << j4 : 0u);
        M |= __reduce_or_sync(kFull, ((M >> lane) & 1u) && j3 < 32 ? 1u << j3 : 0u);
        M |= __reduce_or_sync(kFull, ((M >> lane) & 1u) && j2 < 32 ? 1u << j2 : 0u);
        M |= __reduce_or_sync(kFull, ((M >> lane) & 1u) && j1 < 32 ? 1u << j1 : 0u);
        M |= __reduce_or_sync(kFull, ((M >> lane) & 1u) && j0 < 32 ? 1u << j0 : 0u);

        // cut the step before the first element that needs special handling
        const bool special = slow || (is_lit && len >= kLongLiteral);
        const unsigned S = __ballot_sync(kFull, ((M >> lane) & 1u) && special);
        if (S & 1u) {
            // ---- one-element path for the element at ip (uniform: every lane recomputes it)
            const uint32_t t0 = __shfl_sync(kFull, v, 0);
            const uint32_t tg = t0 & 0xffu;
            const uint32_t ty = tg & 3u;
            const uint32_t b4 = (ip + 4 < clen) ? (uint32_t)__ldg(in + ip + 4) : 0u;
            uint32_t h, l, off = 0;
            if (ty == 0) {
                const uint32_t m = tg >> 2;
                const uint32_t k = m >= 60 ? m - 59 : 0;
                h = 1 + k;
                const uint32_t raw = (t0 >> 8) | (b4 << 24);
                l = k == 0 ? m : (k == 4 ? raw : raw & ((1u << (8 * k)) - 1u));
                if (ip + h > clen || l >= olen - op || (uint64_t)ip + h + l + 1 > clen) {
                    err = (ip + h <= clen && (uint64_t)ip + h + l + 1 <= clen) ? SNAPPY_B200_ST_FRAMING
                                                                             : SNAPPY_B200_ST_CORRUPT;
                    break;
                }
                l += 1;
                coop_copy_ro(out + op, in + ip + h, l, lane, 32);
                ip += h + l;
            } else {
                // copy-4 (src/snappy_decompression.c:323-327)
                h = 5;
                l = (tg >> 2) + 1;
                off = (t0 >> 8) | (b4 << 24);
                if (ip + h > clen || off == 0) {
                    err = SNAPPY_B200_ST_CORRUPT;
                    break;
                }
                if (off > op || l > olen - op) {
                    err = ((uint64_t)off > blk * (uint64_t)kBlock + op) ? SNAPPY_B200_ST_CORRUPT
                                                                        : SNAPPY_B200_ST_FRAMING;
                    break;
                }
                __syncwarp();
                for (uint32_t i = lane; i < l; i += 32)
                    out[op + i] = out[op - off + (off >= l ? i : i % off)];
                ip += h;
            }
            op += l;
            __syncwarp();
            continue;
        }
        if (S)
            M &= (1u << (__ffs((int)S) - 1)) - 1u;
        const bool mine = (M >> lane) & 1u;

        // ---- 3. output offsets (inclusive prefix sum; non-element lanes contribute nothing)
        const uint32_t mylen = mine ? len : 0u;
        uint32_t end = mylen;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, end, d);
            if ((int)lane >= d)
                end += t;
        }
        const uint32_t T = __shfl_sync(kFull, end, 31);
        const uint32_t start = end - mylen; // window-relative output offset of my element
        // validation
        bool bad_corrupt = mine && (!inside || pos + hdr > clen || (is_lit && pos + size > clen) ||
                                    (!is_lit && info == 0));
        bool bad_framing = mine && !is_lit && info > op + start;
        if (T > olen - op)
            bad_framing = true;
        const unsigned BC = __ballot_sync(kFull, bad_corrupt), BF = __ballot_sync(kFull, bad_framing);
        if (BC | BF) {
            err = BC ? SNAPPY_B200_ST_CORRUPT : SNAPPY_B200_ST_FRAMING;
            break;
        }
        const int last = 31 - __clz((int)M);
        const uint32_t consumed = __shfl_sync(kFull, lane + size, last);

        // ---- 4. produce T output bytes, 32 per round
        // The elements of the window are compacted into a small shared-memory table (rank =
        // number of element lanes below mine).  In every round the lane that produces output
        // byte k finds its element from a 32-bit map of the element starts inside the round:
        // rank = (#elements starting before the round) + popc(map up to k) - 1.
        const uint32_t rank = __popc(M & ((1u << lane) - 1u));
        if (mine) {
            // small offsets repeat their pattern (offset < length): keep 2^16/offset for "i mod offset"
            const uint32_t inv = (!is_lit && info < len) ? 65536u / info + 1u : 0u;
            elems[rank] = make_uint4(start, len, info | (is_lit ? 0x80000000u : 0u), inv);
        }
        __syncwarp(); // also: stores of earlier steps are visible to every lane from here
        const uint32_t ne = __popc(M);
        const uint32_t my_start = lane < ne ? elems[lane].x : 0xffffffffu;
        for (uint32_t c = 0; c < T; c += 32) {
            const uint32_t before = __popc(__ballot_sync(kFull, my_start < c));
            const uint32_t rel = my_start - c;
            const unsigned B = __reduce_or_sync(kFull, rel < 32u ? 1u << rel : 0u);
            uint32_t k = c + lane; // window-relative output byte
            bool pending = k < T;
            uint32_t val = 0;
            do {
                if (pending) {
                    const uint32_t r = before + __popc(B & (0xffffffffu >> (31u - (k - c)))) - 1u;
                    const uint4 e = elems[r];
                    const uint32_t i = k - e.x;
                    if (e.z & 0x80000000u) {
                        val = __ldg(in + (e.z & 0x7fffffffu) + i); // write_literal :232-239
                        pending = false;
                    } else {
                        const uint32_t off = e.z;
                        // write_copy :273-280: byte i comes from i mod offset when the copy overlaps itself
                        const uint32_t s = e.w ? i - off * ((i * e.w) >> 16) : i;
                        const uint32_t src = op + e.x + s - off; // block-relative, >= 0 (checked above)
                        if (src < op + c) {
                            val = out[src]; // written by an earlier step or an earlier round
                            pending = false;
                        } else {
                            k = src - op; // produced by this very round: follow it back
                        }
                    }
                }
            } while (__any_sync(kFull, pending));
            if (c + lane < T)
                out[op + c + lane] = (uint8_t)val;
            __syncwarp(); // the next round may read what this one wrote
        }
        ip += consumed;
        op += T;
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
