/*
 * snappy_oracle.c -- CPU restatement of the lightweight-snappy codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (lightweight-snappy_b200/)
 * may link, import or execute this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or the timed CPU baseline.
 *
 * Parity status: PINNED.  This restatement is checked byte-for-byte against
 * the reference itself compiled unmodified from /root/reference/src into
 * oracle/_ref/ (see oracle/Makefile, tests/test_oracle_vs_ref.py) and against
 * the committed golden vectors under tests/golden/ that were captured from
 * that build (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  The code is written from the behavioural spec in
 * SURVEY.md Appendix A, not copied from the reference sources.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "snappy_oracle.h"

/* ------------------------------------------------------------------ varint */

/* src/varint.c:12-20 (parse_to_varint): plain LEB128 of a u64. */
unsigned oracle_varint_encode(uint64_t v, uint8_t *dst)
{
    unsigned k = 0;
    while (v >= 0x80u) {
        dst[k++] = (uint8_t)(v | 0x80u);
        v >>= 7;
    }
    dst[k++] = (uint8_t)v;
    return k;
}

/* src/varint.c:44-58 (str_varint_to_dim_) / :28-42 (varint_to_dim): LEB128
 * decode.  The reference accumulates into an `int`; this restatement keeps a
 * u64 and reports how many bytes were consumed (0 = truncated / too long).  */
unsigned oracle_varint_decode(const uint8_t *src, size_t avail, uint64_t *out)
{
    uint64_t v = 0;
    unsigned shift = 0;
    for (unsigned k = 0; k < avail && k < 10; ++k) {
        v |= (uint64_t)(src[k] & 0x7fu) << shift;
        shift += 7;
        if (!(src[k] & 0x80u)) {
            *out = v;
            return k + 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------- emit helpers */

/* src/snappy_compression.c:95-120 (write_literal). */
static uint8_t *put_literal(uint8_t *o, const uint8_t *src, uint32_t len)
{
    uint32_t m = len - 1;
    if (m < 60) {
        *o++ = (uint8_t)(m << 2);
    } else {
        uint8_t *tag = o++;
        unsigned code = 59;
        while (m > 0) {
            *o++ = (uint8_t)m;
            m >>= 8;
            ++code;
        }
        *tag = (uint8_t)(code << 2);
    }
    memcpy(o, src, len);
    return o + len;
}

/* src/snappy_compression.c:131-145 (write_single_copy). */
static uint8_t *put_copy1(uint8_t *o, uint32_t len, uint32_t off)
{
    if (len < 12 && off < 2048) {
        *o++ = (uint8_t)(((off >> 8) << 5) + ((len - 4) << 2) + 1);
        *o++ = (uint8_t)off;
    } else {
        *o++ = (uint8_t)(((len - 1) << 2) | 2);
        *o++ = (uint8_t)off;
        *o++ = (uint8_t)(off >> 8);
    }
    return o;
}

/* src/snappy_compression.c:153-165 (write_copy). */
static uint8_t *put_copy(uint8_t *o, uint32_t len, uint32_t off)
{
    while (len > 68) {
        o = put_copy1(o, 64, off);
        len -= 64;
    }
    if (len > 64) {
        o = put_copy1(o, 60, off);
        len -= 60;
    }
    return put_copy1(o, len, off);
}

static inline uint32_t be32(const uint8_t *p)
{
    /* src/snappy_compression.c:239-241 (get_next_u32): big-endian load. */
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

/* ------------------------------------------------- exact-key dictionary (BST) */

/* Stand-in for src/BST.c: insert-if-absent (insert_node :30-43), exact find
 * (find_node :66-78).  The bucket index of the reference only partitions the
 * key space, so a flat exact-key map gives the same answers.                 */
#define DICT_SLOTS (1u << 18)
typedef struct {
    uint32_t key[DICT_SLOTS];
    int32_t pos[DICT_SLOTS]; /* -1 = empty */
    uint32_t used[1u << 17];
    uint32_t n_used;
} dict_t;

static void dict_reset(dict_t *d)
{
    for (uint32_t i = 0; i < d->n_used; ++i)
        d->pos[d->used[i]] = -1;
    d->n_used = 0;
}

static uint32_t dict_slot(const dict_t *d, uint32_t key)
{
    uint32_t s = (key * 0x9e3779b1u) >> 14;
    while (d->pos[s] >= 0 && d->key[s] != key)
        s = (s + 1) & (DICT_SLOTS - 1);
    return s;
}

static void dict_insert_if_absent(dict_t *d, uint32_t key, uint32_t pos)
{
    uint32_t s = dict_slot(d, key);
    if (d->pos[s] < 0) {
        d->key[s] = key;
        d->pos[s] = (int32_t)pos;
        d->used[d->n_used++] = s;
    }
}

/* ------------------------------------------------------------ block compress */

/* src/snappy_compression.c:384-403 (compress_next_block) in hash mode,
 * src/snappy_compression_tree.c:269-288 in BST mode.
 * Returns the number of bytes written to out.                                */
static uint32_t compress_block(const uint8_t *b, uint32_t n, uint8_t *out, int mode, dict_t *dict)
{
    uint16_t table[ORACLE_MAX_HTABLE];
    uint8_t *o = out;

    /* set_htable_size, src/snappy_compression.c:198-204 */
    uint32_t ts = 256;
    unsigned lg = 8;
    while (ts < ORACLE_MAX_HTABLE && ts < n) {
        ts <<= 1;
        ++lg;
    }
    const unsigned shift = 32 - lg;

    if (mode == ORACLE_MODE_HASH)
        memset(table, 0, sizeof(table)); /* zero == "position 0", :29, :342-344 */
    else
        dict_reset(dict);

    /* start_new_literal + append_literal, :386-387 */
    uint32_t pos = 1, lit = 1, skip = 33;

    for (;;) {
        /* is_block_end: :229-232 (hash) / tree.c:154-157 (BST post-increments skip) */
        int end = (n - pos) < (skip >> 5) + 15;
        if (mode == ORACLE_MODE_BST)
            ++skip;
        if (end)
            break;

        const uint32_t cur = be32(b + pos); /* generate_hash_index :247-252 */
        uint32_t idx = 0, cand = 0, dslot = 0;
        int hit;
        if (mode == ORACLE_MODE_HASH) {
            idx = (cur * 0x1e35a7bdu) >> shift; /* hash_bytes :81-84 */
            cand = table[idx];
            hit = be32(b + cand) == cur; /* found_match :259-265 */
        } else {
            dslot = dict_slot(dict, cur); /* found_match_tree tree.c:174-180 */
            hit = dict->pos[dslot] >= 0;
            if (hit)
                cand = (uint32_t)dict->pos[dslot];
        }

        if (hit) {
            if (lit > 0) /* emit_literal :313-316 */
                o = put_literal(o, b + pos - lit, lit);
            lit = 0; /* start_new_literal :271-274 */
            skip = 32;
            /* emit_copy :323-329, find_copy_length :61-72 */
            uint32_t len = 4;
            while (pos + len < n && b[pos + len] == b[cand + len])
                ++len;
            o = put_copy(o, len, pos - cand);
            if (mode == ORACLE_MODE_HASH)
                table[idx] = (uint16_t)pos; /* :327 */
            else
                dict->pos[dslot] = (int32_t)pos; /* tree.c:221 */
            pos += len;
        } else {
            const uint32_t prev = be32(b + pos - 1);
            if (mode == ORACLE_MODE_HASH) { /* update_hash_table :303-307 */
                table[(prev * 0x1e35a7bdu) >> shift] = (uint16_t)(pos - 1);
                table[idx] = (uint16_t)pos;
            } else { /* update_hash_table_tree tree.c:204-208 */
                dict_insert_if_absent(dict, prev, pos - 1);
                dict_insert_if_absent(dict, cur, pos);
            }
            const uint32_t step = skip >> 5; /* append_literal :283-287 */
            ++skip;
            lit += step;
            pos += step;
        }
    }
    lit += n - pos; /* exhaust_input :292-297 */
    if (lit > 0)
        o = put_literal(o, b + n - lit, lit);
    return (uint32_t)(o - out);
}

/* src/snappy_compression.c:180-190: worst case 65536 + 1010 per block (+ varint). */
uint64_t oracle_max_compressed_size(uint64_t n)
{
    uint64_t blocks = (n + ORACLE_BLOCK_SIZE - 1) / ORACLE_BLOCK_SIZE;
    return 10 + n + blocks * 1010 + 16;
}

/* src/snappy_compression.c:414-428 (snappy_compress) and
 * src/snappy_compression_tree.c:291-306 (snappy_compress_bst):
 * stream = varint(total) || block_0 || block_1 ...; an empty input gives an
 * empty stream (the varint is never flushed).
 * block_sizes (optional) receives the compressed size of every block, the
 * first one NOT including the varint.                                        */
uint64_t oracle_compress(const uint8_t *in, uint64_t n, uint8_t *out, int mode, uint32_t *block_sizes)
{
    if (n == 0)
        return 0;
    dict_t *dict = NULL;
    if (mode == ORACLE_MODE_BST) {
        dict = (dict_t *)malloc(sizeof(dict_t));
        if (!dict)
            return 0;
        memset(dict->pos, 0xff, sizeof(dict->pos));
        dict->n_used = 0;
    }
    uint8_t *o = out;
    o += oracle_varint_encode(n, o);
    uint64_t done = 0, bi = 0;
    while (done < n) {
        uint32_t len = (uint32_t)((n - done) < ORACLE_BLOCK_SIZE ? (n - done) : ORACLE_BLOCK_SIZE);
        uint32_t c = compress_block(in + done, len, o, mode, dict);
        if (block_sizes)
            block_sizes[bi] = c;
        ++bi;
        o += c;
        done += len;
    }
    free(dict);
    return (uint64_t)(o - out);
}

/* ---------------------------------------------------------------- decompress */

/* src/snappy_decompression.c:345-363 (snappy_decompress), :290-333
 * (decompressor), :193-224 (do_literal), :253-280 (do_copy / write_copy).
 * Restated as a memory-to-memory decoder with the bounds checks the
 * reference lacks.  Returns 0 on success and stores the produced size.       */
int oracle_decompress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_len)
{
    uint64_t total = 0;
    unsigned hdr = oracle_varint_decode(in, n, &total);
    if (hdr == 0)
        return ORACLE_ERR_VARINT;
    if (total > cap)
        return ORACLE_ERR_CAPACITY;
    uint64_t ip = hdr, op = 0;
    while (op < total) {
        if (ip >= n)
            return ORACLE_ERR_TRUNCATED;
        const uint8_t tag = in[ip++];
        uint64_t len, off;
        switch (tag & 3) {
        case 0: { /* literal, do_literal :193-224 */
            len = (uint64_t)(tag >> 2);
            if (len >= 60) {
                unsigned k = (unsigned)len - 59;
                if (ip + k > n)
                    return ORACLE_ERR_TRUNCATED;
                len = 0;
                for (unsigned j = 0; j < k; ++j)
                    len |= (uint64_t)in[ip + j] << (8 * j);
                ip += k;
            }
            len += 1;
            if (ip + len > n)
                return ORACLE_ERR_TRUNCATED;
            if (op + len > total)
                return ORACLE_ERR_OVERRUN;
            memcpy(out + op, in + ip, len); /* write_literal :232-239 */
            ip += len;
            op += len;
            continue;
        }
        case 1: /* copy-1, :311-317 */
            if (ip + 1 > n)
                return ORACLE_ERR_TRUNCATED;
            len = ((tag >> 2) & 7) + 4;
            off = ((uint64_t)(tag >> 5) << 8) | in[ip];
            ip += 1;
            break;
        case 2: /* copy-2, :318-322 */
            if (ip + 2 > n)
                return ORACLE_ERR_TRUNCATED;
            len = (tag >> 2) + 1;
            off = (uint64_t)in[ip] | ((uint64_t)in[ip + 1] << 8);
            ip += 2;
            break;
        default: /* copy-4, :323-327 */
            if (ip + 4 > n)
                return ORACLE_ERR_TRUNCATED;
            len = (tag >> 2) + 1;
            off = (uint64_t)in[ip] | ((uint64_t)in[ip + 1] << 8) | ((uint64_t)in[ip + 2] << 16) |
                  ((uint64_t)in[ip + 3] << 24);
            ip += 4;
            break;
        }
        if (off == 0 || off > op)
            return ORACLE_ERR_OFFSET;
        if (op + len > total)
            return ORACLE_ERR_OVERRUN;
        /* write_copy :273-280: forward byte copy, overlap repeats the pattern */
        for (uint64_t j = 0; j < len; ++j)
            out[op + j] = out[op + j - off];
        op += len;
    }
    *out_len = op;
    return 0;
}

/* Tag walk only: finds, for an index-less stream, the compressed offset at
 * which every 64 KiB output block starts (what the sequential loop of
 * src/snappy_decompression.c:353-356 discovers implicitly).  offsets must hold
 * ceil(total/65536)+1 entries; the last one is the end of the stream.
 * Returns the number of blocks, or a negative ORACLE_ERR_*.                  */
int64_t oracle_block_index(const uint8_t *in, uint64_t n, uint64_t *offsets, uint64_t *total_out)
{
    uint64_t total = 0;
    unsigned hdr = oracle_varint_decode(in, n, &total);
    if (hdr == 0)
        return -ORACLE_ERR_VARINT;
    *total_out = total;
    uint64_t ip = hdr, op = 0, nb = 0;
    while (op < total) {
        if ((op % ORACLE_BLOCK_SIZE) == 0)
            offsets[nb++] = ip;
        if (ip >= n)
            return -ORACLE_ERR_TRUNCATED;
        const uint8_t tag = in[ip++];
        uint64_t len;
        switch (tag & 3) {
        case 0:
            len = (uint64_t)(tag >> 2);
            if (len >= 60) {
                unsigned k = (unsigned)len - 59;
                if (ip + k > n)
                    return -ORACLE_ERR_TRUNCATED;
                len = 0;
                for (unsigned j = 0; j < k; ++j)
                    len |= (uint64_t)in[ip + j] << (8 * j);
                ip += k;
            }
            len += 1;
            ip += len;
            break;
        case 1:
            len = ((tag >> 2) & 7) + 4;
            ip += 1;
            break;
        case 2:
            len = (tag >> 2) + 1;
            ip += 2;
            break;
        default:
            len = (tag >> 2) + 1;
            ip += 4;
            break;
        }
        /* an element may not straddle a 64 KiB output block in this framing */
        if ((op / ORACLE_BLOCK_SIZE) != ((op + len - 1) / ORACLE_BLOCK_SIZE))
            return -ORACLE_ERR_FRAMING;
        op += len;
    }
    if (ip > n)
        return -ORACLE_ERR_TRUNCATED;
    offsets[nb] = ip;
    return (int64_t)nb;
}
