// decode_win.cu -- K4/K5: decompression of an index-less stream.  A sub-warp GROUP of GW lanes decodes
// one 64 KiB output block, every lane one element at a time, through a sliding window of the output
// that lives in shared memory.
//
// reference: decompressor src/snappy_decompression.c:290-333 (tag dispatch), do_literal :193-224,
// write_literal :232-239, do_copy :253-265, write_copy :273-280.
//
// Why this shape (measurements: profiles/r02_*.txt, tools/dep_model.py):
//   * The copy graph of a block is deep: a low-entropy block is one chain of ~2000 dependent levels
//     (each run copies the bytes just before it), so a block is latency bound and the SM only fills
//     with MANY blocks in flight.  A whole 64 KiB output tile per block in shared memory allows 3 per
//     SM (decode_tile.cu: 8x slower on such data); a window of a few KiB allows 48-64.
//   * In a dependent chain only ~2 elements are ready per step, so a 32-lane warp per block wastes its
//     instruction slots.  Here a warp carries 32/GW blocks side by side (GW = 8: four), which divides
//     the instructions per block by the same factor.
//
// K0 (index.cu) has left the exact bit map of the element starts of every 128-byte stream segment.
// One pass of a group:
//   select    lane i takes the i-th next element start from the map (up to GW elements per pass),
//   headers   decodes its header through a tag table; a prefix sum over the group gives every element
//             its output offset,
//   rounds    every lane moves its own element into the window, 16 bytes per step with word loads and
//             stores (lanecopy.cuh).  Literals (source: the stream) and copies that reach back beyond
//             the window (source: output already flushed to HBM) are ready at once; a copy inside the
//             window waits until its whole source range lies below the first pending element of the
//             group (multi-round resolution; elements are in output order, so that is one ballot).
//             Copies that overlap themselves (offset < length, write_copy :273-280) are replicated
//             with a distance that doubles as the periodic region grows.
// When the window is full its finished part is flushed to HBM with 16-byte coalesced stores (a group
// writes 128 contiguous bytes per instruction) and the most recent bytes slide to the front.
// Literals longer than 64 bytes are moved by the whole group (into the window, or straight
// stream -> HBM when longer than a pass can be).  Blocks that are a single literal (incompressible
// data) are moved by k_copy_literal_blocks, a plain 16-byte-vector copy, and skipped here.
// Unlike the reference, malformed input is detected and reported in *status instead of being
// undefined behaviour.
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "lanecopy.cuh"
#include "tags.cuh"

namespace sb200 {

namespace {

constexpr uint32_t kSegB = 128;   // K0 segment size (csrc/index.cu)
constexpr uint32_t kLaneLit = 64; // literals up to this long are moved by one lane
constexpr int kWinThreads = 32; // one warp per CTA: the finest scheduling grain (a CTA ends with its slowest block)
constexpr uint32_t kLutBytes = 512;

// Is the block [c0, c1) of the stream exactly one literal of olen bytes?  hdr = its header length.
__device__ __forceinline__ bool single_literal_block(const uint8_t *__restrict__ stream, uint64_t c0, uint64_t c1,
                                                     uint32_t olen, uint32_t &hdr)
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
    hdr = 1 + k;
    return raw + 1 == olen && c1 - c0 == 1ull + k + olen;
}

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
        return; // k_decode_win reports it
    const uint64_t oleft = total_out - blk * (uint64_t)kBlock;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    uint32_t hdr;
    if (!single_literal_block(stream, c0, c1, olen, hdr))
        return;
    coop_copy_ro(out_base + blk * (uint64_t)kBlock, stream + c0 + hdr, olen, threadIdx.x, 256);
}

template <int GW>
__global__ void __launch_bounds__(kWinThreads)
    k_decode_win(const uint8_t *__restrict__ stream, uint64_t body_offset, const uint64_t *__restrict__ offsets,
                 const uint4 *__restrict__ starts, uint64_t total_out, uint8_t *out_base, uint32_t *__restrict__ status,
                 uint64_t n_blocks, uint64_t blk_base, uint32_t kbuf)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem);
    constexpr uint32_t kGroups = kWinThreads / GW;
    constexpr uint32_t kPassMax = GW * 64u; // output bytes of one pass of ordinary elements
    const uint32_t tid = threadIdx.x, lane = tid & 31u, glane = tid & (GW - 1u), gbase = lane & ~(GW - 1u);
    constexpr unsigned gbits = GW == 32 ? kFull : ((1u << (GW & 31)) - 1u);
    const unsigned gmask = gbits << gbase;
    uint8_t *buf = smem + kLutBytes + (tid / GW) * kbuf; // the window: output bytes [B, B + kbuf)
    // bytes that stay in the window when it slides: half of it (every output byte is moved once on average)
    const uint32_t wkeep = min(kbuf - 2u * kPassMax, (kbuf / 2u) & ~15u);
    for (uint32_t i = tid; i < 256; i += kWinThreads)
        lut[i] = (uint16_t)tag_facts(i);
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t *>(status) != 0)
        return; // K0 rejected the stream: its maps are not trustworthy

    const uint64_t blk = blockIdx.x * (uint64_t)kGroups + tid / GW;
    bool active = blk < n_blocks;
    uint32_t err = 0;
    const uint64_t stream_bytes = offsets[n_blocks];
    uint64_t c0 = body_offset, c1 = body_offset + 1; // (harmless geometry for idle groups)
    if (active) {
        c0 = offsets[blk], c1 = offsets[blk + 1];
        if (c1 <= c0 || c0 < body_offset || c1 > stream_bytes || c1 - c0 > 2u * kBlock) {
            err = SNAPPY_B200_ST_CORRUPT;
            active = false;
            c0 = body_offset, c1 = body_offset + 1;
        }
    }
    const uint64_t oleft = active ? total_out - blk * (uint64_t)kBlock : 0;
    const uint32_t olen = oleft < kBlock ? (uint32_t)oleft : kBlock;
    if (active) {
        uint32_t h;
        if (single_literal_block(stream, c0, c1, olen, h))
            active = false; // k_copy_literal_blocks moves it
    }
    // ---- stream geometry.  Positions P below are relative to the first segment the block touches.
    const uint64_t b0 = c0 - body_offset, b1 = c1 - body_offset;
    const uint64_t t0 = b0 / kSegB;
    const uint32_t nseg = (uint32_t)((b1 - 1) / kSegB - t0) + 1;
    const uint8_t *__restrict__ in = stream + body_offset + t0 * kSegB;
    const uint32_t first = (uint32_t)(b0 - t0 * kSegB); // where the block starts
    const uint32_t lim = (uint32_t)(b1 - t0 * kSegB);   // where it ends
    const uint32_t *last_word = reinterpret_cast<const uint32_t *>(
        reinterpret_cast<uintptr_t>(stream + stream_bytes - 1) & ~uintptr_t(3));
    uint8_t *out = out_base + blk * (uint64_t)kBlock;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out_base) & 15u) == 0;

    uint32_t sidx = 0, seg_lo = 0;          // next segment to load / P of the current one
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0; // element starts of the current segment not consumed yet
    uint32_t cur = 0;                       // output bytes produced so far
    uint32_t B = 0;                         // output position of buf[0] (multiple of 16)
    uint32_t Fl = 0;                        // output bytes [0, Fl) are in HBM (Fl >= B; Fl - B >= 64 once B > 0)

    // window bytes [from, to) -> HBM
    auto flush = [&](uint32_t from, uint32_t to) {
        if (!vec_ok) {
            for (uint32_t i = from + glane; i < to; i += GW)
                out[i] = buf[i - B];
            return;
        }
        uint32_t p = from;
        const uint32_t head = min(to - p, (16u - (p & 15u)) & 15u);
        for (uint32_t i = glane; i < head; i += GW)
            out[p + i] = buf[p - B + i];
        p += head;
        const uint32_t nvec = (to - p) >> 4;
        for (uint32_t i = glane; i < nvec; i += GW)
            *reinterpret_cast<uint4 *>(out + p + 16u * i) = *reinterpret_cast<const uint4 *>(buf + (p - B) + 16u * i);
        p += nvec << 4;
        for (uint32_t i = glane; i < to - p; i += GW)
            out[p + i] = buf[p - B + i];
    };
    // make room: flush what is finished, keep the last wkeep bytes
    auto slide = [&]() {
        __syncwarp(gmask);
        const uint32_t upto = cur & ~15u;
        if (upto > Fl) {
            flush(Fl, upto);
            Fl = upto;
        }
        const uint32_t nb = (cur - min(cur, wkeep)) & ~15u;
        if (nb > B) {
            const uint32_t shift = nb - B, nvec = (cur - nb + 15u) >> 4;
            for (uint32_t i0 = 0; i0 < nvec; i0 += GW) {
                const uint32_t i = i0 + glane;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (i < nvec)
                    v = *reinterpret_cast<const uint4 *>(buf + shift + 16u * i);
                __syncwarp(gmask); // every lane has read before any lane overwrites
                if (i < nvec)
                    *reinterpret_cast<uint4 *>(buf + 16u * i) = v;
                __syncwarp(gmask);
            }
            B = nb;
        }
    };

    // The groups of a warp run in lockstep: every step below is reached by all 32 lanes together, loops run
    // until no group needs another turn, and a group that has nothing to do in a step just idles through it
    // (divergence between groups would serialise them: measured 4x the instructions).
    for (;;) {
        if (!__any_sync(kFull, active))
            break;
        // ---- (1) the next segment that holds element starts of this block
        for (;;) {
            const bool adv = active && (m0 | m1 | m2 | m3) == 0 && sidx < nseg;
            if (!__any_sync(kFull, adv))
                break;
            if (adv) {
                const uint4 sv = __ldg(starts + t0 + sidx);
                uint32_t R[4] = {sv.x, sv.y, sv.z, sv.w};
                seg_lo = sidx * kSegB;
                if (sidx == 0 || sidx + 1 == nseg) // keep only the starts that belong to this block
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t lo = seg_lo + 32 * q; // stream position of bit 0 of this word
                        if (lo + 32 <= first || lo >= lim) {
                            R[q] = 0;
                        } else {
                            if (lo < first)
                                R[q] &= ~((1u << (first - lo)) - 1u);
                            if (lo + 32 > lim)
                                R[q] &= (1u << (lim - lo)) - 1u;
                        }
                    }
                m0 = R[0], m1 = R[1], m2 = R[2], m3 = R[3];
                // the block is a latency-bound chain: have the stream (and its maps) on their way before they are needed
                if (glane == 0 && sidx + 16 < nseg) {
                    prefetch_l2(in + seg_lo + 16 * kSegB);
                    if ((sidx & 7u) == 0)
                        prefetch_l2(starts + t0 + sidx + 16);
                }
                ++sidx;
            }
        }
        {
            const bool fin = active && (m0 | m1 | m2 | m3) == 0; // ---- the block is done
            if (__any_sync(kFull, fin)) {
                if (fin) {
                    __syncwarp(gmask);
                    flush(Fl, cur);
                    if (cur != olen)
                        err |= SNAPPY_B200_ST_CORRUPT;
                    active = false;
                }
            }
        }
        // ---- (2) room for one pass
        {
            const bool sl = active && cur - B + kPassMax > kbuf;
            if (__any_sync(kFull, sl)) {
                if (sl)
                    slide();
            }
        }
        // ---- (3) lane i takes the i-th next element start
        const uint32_t n0 = __popc(m0), n1 = n0 + __popc(m1), n2 = n1 + __popc(m2), n3 = n2 + __popc(m3);
        const uint32_t cnt = active ? min(n3, (uint32_t)GW) : 0u;
        const bool has = glane < cnt;
        uint32_t w, r, wbase;
        if (glane < n0)
            w = m0, r = glane, wbase = 0;
        else if (glane < n1)
            w = m1, r = glane - n0, wbase = 32;
        else if (glane < n2)
            w = m2, r = glane - n1, wbase = 64;
        else
            w = m3, r = glane - n2, wbase = 96;
#pragma unroll
        for (int k = 0; k < GW - 1; ++k)
            if (k < (int)r)
                w &= w - 1u;
        const uint32_t pos = has ? seg_lo + wbase + (uint32_t)(__ffs((int)w) - 1) : seg_lo;
        const uint32_t v = ld_le32_any(in + pos, last_word);
        const Header h = decode_header_lut(lut, v, pos);
        uint32_t len = h.len, info = h.info;
        if (has && h.slow) { // the fifth header byte: top of a 4-byte literal length / of a copy-4 offset
            const uint32_t top = (pos + 4 < lim ? (uint32_t)__ldg(in + pos + 4) : 0u) << 24;
            if (h.is_lit)
                len = ((v >> 8) | top) + 1u; // (0xffffffff + 1 wraps to 0: caught below as an empty element)
            else
                info = (v >> 8) | top;
        }
        const bool special = has && h.is_lit && (len > kLaneLit || len == 0);
        const unsigned S = (__ballot_sync(kFull, special) >> gbase) & gbits;
        uint32_t ntake = S ? (uint32_t)(__ffs((int)S) - 1) : cnt;

        // ---- (4) a literal longer than one lane moves: the whole group
        {
            const bool sp = cnt > 0 && ntake == 0;
            if (__any_sync(kFull, sp)) {
                if (sp) {
                    const uint32_t spos = __shfl_sync(gmask, pos, gbase), shdr = __shfl_sync(gmask, h.hdr, gbase);
                    const uint32_t slen = __shfl_sync(gmask, len, gbase);
                    if (slen == 0 || spos + shdr > lim || (uint64_t)spos + shdr + slen > lim) {
                        err |= SNAPPY_B200_ST_CORRUPT;
                        active = false;
                    } else if (slen > olen - cur) {
                        err |= SNAPPY_B200_ST_FRAMING;
                        active = false;
                    } else {
                        __syncwarp(gmask);
                        if (slen <= kPassMax) { // into the window
                            if (cur - B + slen > kbuf)
                                slide();
                            coop_copy_ro(buf + (cur - B), in + spos + shdr, slen, glane, GW);
                            cur += slen;
                        } else { // straight to HBM, then the window is refilled from there
                            flush(Fl, cur);
                            coop_copy_ro(out + cur, in + spos + shdr, slen, glane, GW);
                            cur += slen;
                            Fl = cur;
                            const uint32_t nb = (cur - min(cur, wkeep)) & ~15u;
                            __syncwarp(gmask); // the group's stores are visible to all of its lanes
                            if (vec_ok) {
                                const uint32_t nvec = (cur - nb + 15u) >> 4;
                                for (uint32_t i = glane; i < nvec; i += GW)
                                    *reinterpret_cast<uint4 *>(buf + 16u * i) =
                                        *reinterpret_cast<const uint4 *>(out + nb + 16u * i);
                            } else {
                                for (uint32_t i = glane; i < cur - nb; i += GW)
                                    buf[i] = out[nb + i];
                            }
                            B = nb;
                        }
                        __syncwarp(gmask);
                        if (m0)
                            m0 &= m0 - 1u;
                        else if (m1)
                            m1 &= m1 - 1u;
                        else if (m2)
                            m2 &= m2 - 1u;
                        else
                            m3 &= m3 - 1u;
                    }
                }
            }
            if (sp)
                ntake = 0; // (nothing else for this group in this turn)
        }

        // ---- (5) output offsets and checks
        bool take = active && glane < ntake;
        const uint32_t mylen = take ? len : 0u;
        uint32_t end = mylen;
#pragma unroll
        for (int d = 1; d < GW; d <<= 1) {
            const uint32_t u = __shfl_up_sync(kFull, end, d, GW);
            if ((int)glane >= d)
                end += u;
        }
        const uint32_t total = __shfl_sync(kFull, end, GW - 1, GW);
        const uint32_t dst = cur + end - mylen;
        const bool bad_c = take && (pos + h.hdr > lim || (h.is_lit && (uint64_t)pos + h.hdr + len > lim) ||
                                    (!h.is_lit && info == 0));
        const bool back = take && !h.is_lit && info > dst; // reaches back into an earlier block
        const bool back_c = back && (uint64_t)info > (blk_base + blk) * (uint64_t)kBlock + dst;
        const unsigned BC = (__ballot_sync(kFull, bad_c || back_c) >> gbase) & gbits;
        const unsigned BF = (__ballot_sync(kFull, back && !back_c) >> gbase) & gbits;
        if (active && ntake > 0 && (BC | BF | (unsigned)(total > olen - cur))) {
            err |= BC ? SNAPPY_B200_ST_CORRUPT : SNAPPY_B200_ST_FRAMING;
            active = false;
            take = false;
        }
        // where my element comes from, and below which output position everything must be final first
        uint8_t *d = buf + (dst - B);
        const uint8_t *s = in + info; // literal: the stream
        uint32_t need = 0, dist = 0xffffu;
        bool near = false;
        if (take && !h.is_lit) {
            const uint32_t a = dst - info;
            if (a >= B) { // inside the window
                s = buf + (a - B);
                near = true;
                dist = info;
                need = info >= len ? a + len : dst; // a self-overlapping copy: everything below its own start
            } else { // beyond the window: those bytes are final and in HBM (a + len <= B + 64 <= Fl)
                s = out + a;
            }
        }
        // ---- (6) rounds
        bool pend = take;
        for (;;) {
            const unsigned Pw = __ballot_sync(kFull, pend);
            if (!Pw)
                break;
            const unsigned P = (Pw >> gbase) & gbits;
            const uint32_t G = __shfl_sync(kFull, dst, P ? __ffs((int)P) - 1 : 0, GW); // elements are in output order
            if (pend && need <= G) {
                const uint8_t *sb = s;
                uint32_t done = 0;
                while (done < len) {
                    const uint32_t n = min(min(16u, len - done), dist); // a step never outruns the distance
                    store16(d + done, load16(sb + done, n), n);
                    done += n;
                    if (near && 2u * dist <= done + info) { // the periodic region has doubled
                        dist <<= 1;
                        sb = d - dist;
                    }
                }
                pend = false;
            }
            __syncwarp();
        }
        // ---- (7) advance
        if (active && ntake > 0) {
            cur += total;
            const uint32_t rel = __shfl_sync(gmask, pos, gbase + ntake - 1) - seg_lo; // last element taken
            m0 = rel >= 31 ? 0u : m0 & ~((2u << rel) - 1u);
            m1 = rel >= 63 ? 0u : (rel >= 32 ? m1 & ~((2u << (rel - 32)) - 1u) : m1);
            m2 = rel >= 95 ? 0u : (rel >= 64 ? m2 & ~((2u << (rel - 64)) - 1u) : m2);
            m3 = rel >= 127 ? 0u : (rel >= 96 ? m3 & ~((2u << (rel - 96)) - 1u) : m3);
        }
    }
    if (err && glane == 0)
        atomicOr(status, err);
}

struct WinConfig {
    int gw;
    uint32_t kbuf;
};

WinConfig win_config()
{
    static const WinConfig cfg = [] {
        WinConfig c{8, 4480};
        if (const char *v = getenv("SNAPPY_B200_WIN_GW"))
            c.gw = atoi(v);
        if (c.gw != 8 && c.gw != 16 && c.gw != 32)
            c.gw = 8;
        if (const char *v = getenv("SNAPPY_B200_WIN_KBUF"))
            c.kbuf = (uint32_t)atoi(v);
        const uint32_t lo = 2u * c.gw * 64u + 2048u; // two passes of room + at least 2 KiB kept
        const uint32_t hi = (227u * 1024u - kLutBytes) / (kWinThreads / c.gw);
        c.kbuf = (c.kbuf < lo ? lo : (c.kbuf > hi ? hi : c.kbuf)) & ~15u;
        return c;
    }();
    return cfg;
}

template <int GW>
cudaError_t launch_win(const uint8_t *d_stream, uint64_t body_offset, const uint64_t *d_offsets, const uint4 *d_starts,
                       uint64_t n_blocks, uint64_t total_out, uint8_t *d_out, uint32_t *d_status, uint64_t blk_base,
                       uint32_t kbuf, cudaStream_t st)
{
    constexpr uint32_t kGroups = kWinThreads / GW;
    const size_t smem = kLutBytes + (size_t)kGroups * kbuf;
    static std::once_flag once[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    std::call_once(once[dev & 63], [&] { // the opt-in to > 48 KiB of dynamic shared memory is per device
        e = cudaFuncSetAttribute(k_decode_win<GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (e != cudaSuccess)
        return e;
    k_decode_win<GW><<<(unsigned)((n_blocks + kGroups - 1) / kGroups), kWinThreads, smem, st>>>(
        d_stream, body_offset, d_offsets, d_starts, total_out, d_out, d_status, n_blocks, blk_base, kbuf);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_decode_win(const uint8_t *d_stream, uint64_t body_offset, const uint64_t *d_offsets,
                              const uint4 *d_starts, uint64_t n_blocks, uint64_t total_out, uint8_t *d_out,
                              uint32_t *d_status, uint64_t blk_base, cudaStream_t st, uint64_t *launches)
{
    if (n_blocks == 0)
        return cudaSuccess;
    if (n_blocks > 0x7fffffffull)
        return cudaErrorInvalidValue;
    const WinConfig cfg = win_config();
    k_copy_literal_blocks<<<(unsigned)n_blocks, 256, 0, st>>>(d_stream, body_offset, d_offsets, total_out, d_out, d_status,
                                                             n_blocks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    if (cfg.gw == 8)
        e = launch_win<8>(d_stream, body_offset, d_offsets, d_starts, n_blocks, total_out, d_out, d_status, blk_base,
                          cfg.kbuf, st);
    else if (cfg.gw == 16)
        e = launch_win<16>(d_stream, body_offset, d_offsets, d_starts, n_blocks, total_out, d_out, d_status, blk_base,
                           cfg.kbuf, st);
    else
        e = launch_win<32>(d_stream, body_offset, d_offsets, d_starts, n_blocks, total_out, d_out, d_status, blk_base,
                           cfg.kbuf, st);
    *launches += 2;
    return e;
}

} // namespace sb200
