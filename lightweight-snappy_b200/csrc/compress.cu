// compress.cu -- K1 (hash-table parse) and K2 (exact-key parse): one warp per 64 KiB block.
//
// What is computed is exactly the greedy parse of the reference
//   hash mode : src/snappy_compression.c:384-403 (compress_next_block)
//   exact mode: src/snappy_compression_tree.c:269-288 with the dictionary of src/BST.c
// and the element encodings of src/snappy_compression.c:95-165.  How it is computed is not:
//
// The parse is serial in the table state, so a warp speculates.  Each step lays the next 16
// probe positions out as 32 "events" (lane 2j = the p-1 insertion that probe j would make on
// a miss, lane 2j+1 = probe j itself), under the assumption that every earlier probe of the
// step misses.  Lanes look their key up in the shared-memory table AND in the earlier lanes
// of the same step (__match_any_sync), which reproduces what the table would hold after
// those misses.  The first probe that hits (or reaches the end-of-block test) cuts the step:
// the misses before it are committed to the table in one go, the hit is extended with a
// warp-wide 4-byte-per-lane compare, and literal + copy elements are emitted.  Everything the
// speculation assumed about lanes after the cut is simply discarded, so the result is the
// reference's parse, bit for bit.
//
// Hash-mode table entry: (16-bit key fingerprint << 16) | u16 position.  The fingerprint is
// bits 19..4 of key*0x1e35a7bd -- together with the 12-bit slot index it pins 28 of the 32
// bits of an injective function of the key, so the "does the candidate's 4 bytes equal mine"
// test of the reference (found_match, :259-265) is answered from shared memory; the candidate
// bytes are only fetched to confirm the one lane that wins.  The reference's zero-filled
// table ("candidate = position 0") becomes (fingerprint of the block's first 4 bytes, 0).
//
// Exact mode keeps an open-addressing table of u16 positions keyed by the exact 4 bytes
// (0xffff = empty, which no insertable position can be: positions >= n-15 are never probed).
// Keys are compared through the block itself.  Tables come in three sizes; a block whose
// dictionary outgrows the small table is marked and redone by the next tier.
#include "common.cuh"

namespace sb200 {

// ------------------------------------------------------------------------------- emission
// reference: write_literal, src/snappy_compression.c:95-120
__device__ __forceinline__ void emit_literal(const uint8_t *__restrict__ b, uint32_t src, uint32_t len,
                                             uint8_t *__restrict__ out, uint32_t &o, uint32_t lane)
{
    const uint32_t m = len - 1;
    uint32_t hdr;
    if (m < 60) {
        hdr = 1;
        if (lane == 0)
            out[o] = (uint8_t)(m << 2);
    } else if (m < 256) {
        hdr = 2;
        if (lane == 0) {
            out[o] = 60u << 2;
            out[o + 1] = (uint8_t)m;
        }
    } else { // m <= 65535 inside a 64 KiB block
        hdr = 3;
        if (lane == 0) {
            out[o] = 61u << 2;
            out[o + 1] = (uint8_t)m;
            out[o + 2] = (uint8_t)(m >> 8);
        }
    }
    coop_copy_ro(out + o + hdr, b + src, len, lane, 32);
    o += hdr + len;
}

// reference: write_copy :153-165 and write_single_copy :131-145
__device__ __forceinline__ void emit_copy(uint8_t *__restrict__ out, uint32_t &o, uint32_t len, uint32_t off,
                                          uint32_t lane)
{
    const uint32_t n64 = len > 68 ? (len - 5) / 64 : 0; // "while (len > 68) emit 64"
    uint32_t rem = len - 64 * n64;
    for (uint32_t k = lane; k < n64; k += 32) {
        uint8_t *p = out + o + 3 * k;
        p[0] = 0xfe; // ((64-1) << 2) | 2
        p[1] = (uint8_t)off;
        p[2] = (uint8_t)(off >> 8);
    }
    o += 3 * n64;
    if (rem > 64) { // 64 < rem <= 68: emit 60 so that at least 4 remain
        if (lane == 0) {
            out[o] = 0xee; // ((60-1) << 2) | 2
            out[o + 1] = (uint8_t)off;
            out[o + 2] = (uint8_t)(off >> 8);
        }
        o += 3;
        rem -= 60;
    }
    if (rem < 12 && off < 2048) {
        if (lane == 0) {
            out[o] = (uint8_t)(((off >> 8) << 5) + ((rem - 4) << 2) + 1);
            out[o + 1] = (uint8_t)off;
        }
        o += 2;
    } else {
        if (lane == 0) {
            out[o] = (uint8_t)(((rem - 1) << 2) | 2);
            out[o + 1] = (uint8_t)off;
            out[o + 2] = (uint8_t)(off >> 8);
        }
        o += 3;
    }
}

// reference: find_copy_length :61-72 (+4 for the bytes the probe already matched).
// Lane l compares bytes [base+4l, base+4l+4) of the two strings; the first lane that sees a
// difference (or the end of the block) decides.
__device__ __forceinline__ uint32_t match_extend(const uint8_t *__restrict__ b, uint32_t p, uint32_t c, uint32_t n,
                                                 uint32_t last_word, uint32_t lane)
{
    uint32_t base = 4;
    uint32_t width = 8; // most matches are short: first look at 32 bytes, then 128 at a time
    for (;;) {
        uint32_t t = 4; // number of equal bytes in this lane's word; 4 = keep going
        const uint32_t pp = p + base + 4 * lane;
        if (lane < width) {
            if (pp >= n) {
                t = 0;
            } else {
                const uint32_t nvalid = min(4u, n - pp);
                uint32_t x = ld_le32(b, pp, last_word) ^ ld_le32(b, c + base + 4 * lane, last_word);
                if (nvalid < 4)
                    x = (x & ((1u << (8 * nvalid)) - 1u)) | (1u << (8 * nvalid));
                if (x)
                    t = (uint32_t)(__ffs((int)x) - 1) >> 3;
            }
        }
        const unsigned stop = __ballot_sync(kFull, t < 4);
        if (stop) {
            const int l = __ffs((int)stop) - 1;
            return base + 4 * (uint32_t)l + __shfl_sync(kFull, t, l);
        }
        base += 4 * width;
        width = 32;
    }
}

// ------------------------------------------------------------------------------- exact table
template <int LOG_SLOTS> struct ExactTable {
    static constexpr uint32_t kSlots = 1u << LOG_SLOTS;
    static constexpr uint32_t kEmpty = 0xffffu;
    uint16_t *tab;

    __device__ __forceinline__ uint32_t home(uint32_t key) const { return (key * 0x9e3779b1u) >> (32 - LOG_SLOTS); }

    // Returns the slot holding `key`, or the first empty slot of its probe sequence.
    __device__ __forceinline__ uint32_t find(const uint8_t *__restrict__ b, uint32_t last_word, uint32_t key,
                                             bool &found, uint32_t &pos) const
    {
        uint32_t s = home(key);
        for (;;) {
            const uint32_t v = tab[s];
            if (v == kEmpty) {
                found = false;
                pos = 0;
                return s;
            }
            if (ld_be32(b, v, last_word) == key) {
                found = true;
                pos = v;
                return s;
            }
            s = (s + 1) & (kSlots - 1);
        }
    }

    // Inserts a key known to be absent (several lanes may insert different keys at once).
    __device__ __forceinline__ void insert_absent(uint32_t key, uint32_t pos)
    {
        uint32_t s = home(key);
        for (;;) {
            const unsigned short old = atomicCAS(reinterpret_cast<unsigned short *>(tab + s),
                                                 (unsigned short)kEmpty, (unsigned short)pos);
            if (old == kEmpty)
                return;
            s = (s + 1) & (kSlots - 1);
        }
    }
};

// ------------------------------------------------------------------------------- the kernel
// MODE 0 = hash table (LOG_SLOTS ignored: the table has up to 4096 u32 entries)
// MODE 1 = exact dictionary with 2^LOG_SLOTS u16 slots; a block whose dictionary would grow
//          past 3/4 of the table gives up (sizes[blk] = kAbortMark) unless FINAL.
template <int MODE, int LOG_SLOTS, bool FINAL>
__global__ void __launch_bounds__(32) k_compress(const uint8_t *__restrict__ in, uint64_t n_bytes,
                                                 uint8_t *__restrict__ scratch, uint32_t *__restrict__ sizes,
                                                 int only_marked)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x;
    const uint64_t blk = blockIdx.x;
    if (only_marked && sizes[blk] != kAbortMark)
        return;

    const uint8_t *__restrict__ b = in + blk * (uint64_t)kBlock;
    const uint64_t left = n_bytes - blk * (uint64_t)kBlock;
    const uint32_t n = left < kBlock ? (uint32_t)left : kBlock;
    const uint32_t last_word = (n - 1) >> 2;
    uint8_t *__restrict__ out = scratch + blk * (uint64_t)kSlot;

    // per-miss skip bookkeeping: hash mode skip += 1 per miss and the step uses the value the
    // end test saw (:229-232, :283-287); BST mode post-increments inside the end test as well
    // (tree.c:154-157), so skip += 2 per miss and the step sees skip+1.
    constexpr uint32_t C = MODE == 0 ? 1 : 2;
    constexpr uint32_t D = MODE == 0 ? 0 : 1;

    uint32_t *htab = reinterpret_cast<uint32_t *>(smem_raw);
    ExactTable<LOG_SLOTS> et{reinterpret_cast<uint16_t *>(smem_raw)};
    uint32_t shift = 20;
    uint32_t n_keys = 0; // exact mode: dictionary population (warp-uniform)

    if (MODE == 0) {
        // set_htable_size, :198-204
        uint32_t lg = 8;
        while ((1u << lg) < SNAPPY_B200_HTABLE_SIZE && (1u << lg) < n)
            ++lg;
        shift = 32 - lg;
        const uint32_t fp0 = n >= 4 ? ((ld_be32(b, 0, last_word) * kHashMul) >> 4) & 0xffffu : 0;
        const uint4 init = make_uint4(fp0 << 16, fp0 << 16, fp0 << 16, fp0 << 16);
        uint4 *t4 = reinterpret_cast<uint4 *>(htab);
        for (uint32_t i = lane; i < (1u << lg) / 4; i += 32)
            t4[i] = init;
    } else {
        uint4 *t4 = reinterpret_cast<uint4 *>(smem_raw);
        const uint4 init = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        for (uint32_t i = lane; i < ExactTable<LOG_SLOTS>::kSlots / 8; i += 32)
            t4[i] = init;
    }
    __syncwarp();

    uint32_t pos = 1, skip = 33, lit_start = 0, o = 0; // :386-387: the first byte is a literal

    const uint32_t j = lane >> 1;
    const bool is_probe = lane & 1u;
    const unsigned vis_mask = lane >= 1 ? (1u << (lane - 1)) - 1u : 0u; // events a probe may see

    for (;;) {
        // ---- lay out 16 probes under the all-miss assumption
        const uint32_t a = skip + D;
        const uint32_t q = a >> 5, r = a & 31u;
        const uint32_t k1 = (32u - r + C - 1u) / C; // first probe index whose step is q+1
        const uint32_t pj = pos + j * q + (j > k1 ? j - k1 : 0u);
        const bool end_j = pj + ((skip + C * j) >> 5) + 15u > n; // is_block_end
        const uint32_t ev_pos = is_probe ? pj : pj - 1;
        const uint32_t key = ld_be32(b, end_j ? 0u : ev_pos, last_word);

        bool cand_hit;   // this lane's probe would hit (hash mode: still to be confirmed)
        bool forwarded;  // ... through an earlier event of the same step
        uint32_t cand;   // candidate position
        uint32_t idx = 0, fpv = 0;
        unsigned grp;
        bool in_table = false;

        if (MODE == 0) {
            const uint32_t prod = key * kHashMul; // hash_bytes :81-84
            idx = prod >> shift;
            fpv = (prod >> 4) & 0xffffu;
            grp = __match_any_sync(kFull, idx);
            const unsigned vis = grp & vis_mask;
            const int src = vis ? 31 - __clz((int)vis) : 0; // latest earlier writer of the slot
            const uint32_t skey = __shfl_sync(kFull, key, src);
            const uint32_t spos = __shfl_sync(kFull, ev_pos, src);
            const uint32_t entry = htab[idx];
            forwarded = vis != 0;
            cand_hit = forwarded ? skey == key : (entry >> 16) == fpv;
            cand = forwarded ? spos : entry & 0xffffu;
        } else {
            grp = __match_any_sync(kFull, key);
            const unsigned vis = grp & vis_mask;
            const int src = vis ? __ffs((int)vis) - 1 : 0; // first occurrence wins (insert-if-absent)
            const uint32_t spos = __shfl_sync(kFull, ev_pos, src);
            uint32_t tpos = 0;
            if (!end_j)
                (void)et.find(b, last_word, key, in_table, tpos);
            forwarded = !in_table && vis != 0;
            cand_hit = in_table || forwarded;
            cand = in_table ? tpos : spos;
        }

        unsigned H = __ballot_sync(kFull, is_probe && !end_j && cand_hit);
        const unsigned E = __ballot_sync(kFull, is_probe && end_j);
        const int first_end = E ? __ffs((int)E) - 1 : 32;
        if (first_end < 32)
            H &= (1u << first_end) - 1u;

        // ---- first real hit
        int f = -1;
        while (H) {
            const int c = __ffs((int)H) - 1;
            bool ok = true;
            if (MODE == 0) {
                const uint32_t ccand = __shfl_sync(kFull, cand, c);
                const uint32_t ckey = __shfl_sync(kFull, key, c);
                const bool cfwd = __shfl_sync(kFull, (int)forwarded, c);
                ok = cfwd || ld_be32(b, ccand, last_word) == ckey; // found_match :259-265
            }
            if (ok) {
                f = c;
                break;
            }
            H &= H - 1; // fingerprint collision: that probe is a miss after all
        }

        // ---- commit the misses before the cut (the cut probe's own p-1 event is not one)
        const int L = f >= 0 ? f - 1 : (first_end < 32 ? first_end - 1 : 32);
        const unsigned cm = L >= 32 ? kFull : (1u << L) - 1u;
        const unsigned g = grp & cm;
        if (MODE == 0) {
            // update_hash_table :303-307: in program order the last writer of a slot wins
            if ((int)lane < L && 31 - __clz((int)g) == (int)lane)
                htab[idx] = (fpv << 16) | ev_pos;
        } else {
            // insert-if-absent, src/BST.c:30-43: the first occurrence of a new key is kept
            const bool ins = (int)lane < L && !in_table && __ffs((int)g) - 1 == (int)lane;
            if (ins)
                et.insert_absent(key, ev_pos);
            n_keys += __popc(__ballot_sync(kFull, ins));
        }
        __syncwarp();

        if (f >= 0) {
            const uint32_t p = __shfl_sync(kFull, ev_pos, f);
            const uint32_t c = __shfl_sync(kFull, cand, f);
            if (MODE == 0) {
                if ((int)lane == f)
                    htab[idx] = (fpv << 16) | ev_pos; // emit_copy :327
            } else {
                if ((int)lane == f) { // tree.c:221: the found node now points at this position
                    bool fnd;
                    uint32_t tp;
                    const uint32_t s = et.find(b, last_word, key, fnd, tp);
                    et.tab[s] = (uint16_t)ev_pos;
                }
            }
            if (p > lit_start)
                emit_literal(b, lit_start, p - lit_start, out, o, lane); // emit_literal :313-316
            const uint32_t len = match_extend(b, p, c, n, last_word, lane);
            emit_copy(out, o, len, p - c, lane);
            pos = p + len;
            lit_start = pos;
            skip = 32; // start_new_literal :271-274
            __syncwarp();
        } else if (first_end < 32) {
            break;
        } else {
            pos += 16 * q + (16 > k1 ? 16 - k1 : 0u); // 16 x append_literal :283-287
            skip += 16 * C;
        }

        if (MODE == 1 && !FINAL && n_keys > (ExactTable<LOG_SLOTS>::kSlots * 3) / 4) {
            if (lane == 0)
                sizes[blk] = kAbortMark;
            return;
        }
    }

    if (n > lit_start)
        emit_literal(b, lit_start, n - lit_start, out, o, lane); // exhaust_input :292-297, :401-402
    if (lane == 0)
        sizes[blk] = o;
}

// ------------------------------------------------------------------------------- launchers
static uint64_t g_launches_compress = 0;

cudaError_t launch_compress(const uint8_t *d_in, uint64_t n_bytes, int mode, uint8_t *d_scratch, uint32_t *d_sizes,
                            cudaStream_t st, uint64_t *launches)
{
    const uint64_t nb = (n_bytes + kBlock - 1) / kBlock;
    if (nb == 0)
        return cudaSuccess;
    if (nb > 0x7fffffffull)
        return cudaErrorInvalidValue;
    const dim3 grid((unsigned)nb), cta(32);
    if (mode == SNAPPY_B200_MODE_HASH) {
        k_compress<0, 12, true><<<grid, cta, 4096 * 4, st>>>(d_in, n_bytes, d_scratch, d_sizes, 0);
        *launches += 1;
        return cudaGetLastError();
    }
    // exact mode: 8 Ki slots (16 KiB), then 32 Ki slots (64 KiB), then 64 Ki slots (128 KiB)
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_compress<1, 15, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (1 << 15) * 2);
        if (e != cudaSuccess)
            return e;
        e = cudaFuncSetAttribute(k_compress<1, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 16) * 2);
        if (e != cudaSuccess)
            return e;
        attr_done = true;
    }
    k_compress<1, 13, false><<<grid, cta, (1 << 13) * 2, st>>>(d_in, n_bytes, d_scratch, d_sizes, 0);
    k_compress<1, 15, false><<<grid, cta, (1 << 15) * 2, st>>>(d_in, n_bytes, d_scratch, d_sizes, 1);
    k_compress<1, 16, true><<<grid, cta, (1 << 16) * 2, st>>>(d_in, n_bytes, d_scratch, d_sizes, 1);
    *launches += 3;
    (void)g_launches_compress;
    return cudaGetLastError();
}

} // namespace sb200
