// dropin.cpp -- the reference's own entry points (include/snappy_compression.h,
// snappy_compression_tree.h, snappy_decompression.h, varint.h, buffer_compression.h) on top
// of the host-buffer API.  Same symbol names, signatures, FILE* ownership rules and stream
// bytes as the reference; the work happens on the GPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "buffer_compression.h"
#include "snappy_b200.h"
#include "snappy_compression.h"
#include "snappy_compression_tree.h"
#include "snappy_decompression.h"
#include "varint.h"

namespace sb200 {
int fail_msg(int code, const char *msg); // abi.cu: records the thread's last error
void clear_error();
} // namespace sb200

namespace {

int io_fail(const char *what) { return sb200::fail_msg(SNAPPY_B200_ERR_IO, what); }

// A growable byte buffer in page-locked host memory (plain malloc when no device is usable, so
// that error paths still work): what the reference's Buffer / IO_utils allocations become.
struct PinnedBuf {
    uint8_t *p = nullptr;
    size_t cap = 0, n = 0;
    bool pinned = false;
    ~PinnedBuf() { drop(); }
    void drop()
    {
        if (p) {
            if (pinned)
                snappy_b200_host_free(p);
            else
                free(p);
        }
        p = nullptr;
        cap = n = 0;
    }
    bool reserve(size_t want)
    {
        if (want <= cap)
            return true;
        void *q = snappy_b200_host_alloc(want);
        bool qp = q != nullptr;
        if (!q)
            q = malloc(want);
        if (!q)
            return false;
        if (n)
            memcpy(q, p, n);
        const size_t keep = n;
        drop();
        p = static_cast<uint8_t *>(q);
        cap = want;
        n = keep;
        pinned = qp;
        return true;
    }
    uint8_t *data() { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
};

// Reads from the current position to EOF (the reference's fread loop, src/snappy_compression.c:210-213).
bool read_all(FILE *f, uint64_t hint, PinnedBuf &buf)
{
    size_t cap = hint ? hint + 1 : (1u << 20);
    buf.n = 0;
    for (;;) {
        if (!buf.reserve(cap))
            return false;
        const size_t got = fread(buf.p + buf.n, 1, cap - buf.n, f);
        buf.n += got;
        if (buf.n < cap) {
            if (ferror(f))
                return false;
            break;
        }
        cap *= 2;
    }
    return true;
}

const char kIndexMagic[8] = {'S', 'N', 'P', 'I', 'D', 'X', '1', 0};

int compress_file(FILE *in, unsigned long long declared, FILE *out, int mode, FILE *index_out = nullptr)
{
    PinnedBuf data, stream;
    sb200::clear_error();
    if (!in || !out)
        return io_fail("null FILE*");
    if (!read_all(in, declared, data))
        return io_fail("reading the input failed (read error or out of host memory)");
    if (data.empty())
        return SNAPPY_B200_OK; // reference: empty input -> empty output
    if (!stream.reserve(snappy_b200_max_compressed_bytes(data.size())))
        return io_fail("out of host memory for the compressed stream");
    uint64_t n = 0;
    const uint64_t nb = snappy_b200_block_count(data.size());
    uint64_t *offsets = index_out ? static_cast<uint64_t *>(malloc((nb + 1) * 8)) : nullptr;
    if (index_out && !offsets)
        return io_fail("out of host memory for the block index");
    struct Free {
        void *p;
        ~Free() { free(p); }
    } free_offsets{offsets};
    const int rc =
        snappy_b200_compress_host_indexed(data.data(), data.size(), mode, stream.data(), stream.cap, &n, offsets);
    if (rc != SNAPPY_B200_OK)
        return rc; // (the message is in snappy_b200_last_error())
    // The reference writes the DECLARED size into the preamble (src/snappy_compression.c:417)
    // and then whatever the file really held; keep that even when the two disagree.
    unsigned char hdr[10];
    const unsigned hdr_real = parse_to_varint(data.size(), hdr);
    const unsigned hdr_decl = parse_to_varint(declared, hdr);
    if (fwrite(hdr, 1, hdr_decl, out) != hdr_decl)
        return io_fail("short write of the compressed stream");
    if (fwrite(stream.data() + hdr_real, 1, n - hdr_real, out) != n - hdr_real)
        return io_fail("short write of the compressed stream");
    if (index_out) {
        // offsets as they are in the file just written (the declared-size preamble may be longer or shorter)
        for (uint64_t b = 0; b <= nb; ++b)
            offsets[b] = offsets[b] - hdr_real + hdr_decl;
        const uint64_t head[2] = {data.size(), nb};
        if (fwrite(kIndexMagic, 1, 8, index_out) != 8 || fwrite(head, 8, 2, index_out) != 2 ||
            fwrite(offsets, 8, nb + 1, index_out) != nb + 1)
            return io_fail("short write of the block index");
    }
    return SNAPPY_B200_OK;
}

} // namespace

extern "C" {

int snappy_b200_compress_file_indexed(FILE *in, unsigned long long input_size, int mode, FILE *out, FILE *index_out)
{
    return compress_file(in, input_size, out, mode, index_out);
}

int snappy_b200_decompress_file_indexed(FILE *in, FILE *index_in, FILE *out)
{
    PinnedBuf stream, data;
    sb200::clear_error();
    if (!in || !index_in || !out || !read_all(in, 0, stream))
        return io_fail("reading the compressed stream failed");
    char magic[8];
    uint64_t head[2];
    if (fread(magic, 1, 8, index_in) != 8 || memcmp(magic, kIndexMagic, 8) != 0 || fread(head, 8, 2, index_in) != 2 ||
        head[1] != snappy_b200_block_count(head[0]) || head[1] >= (1ull << 31)) {
        return sb200::fail_msg(SNAPPY_B200_ERR_CORRUPT, "not a block index file");
    }
    uint64_t *offsets = static_cast<uint64_t *>(malloc((head[1] + 1) * 8));
    if (!offsets)
        return io_fail("out of host memory for the block index");
    int rc = SNAPPY_B200_OK;
    if (fread(offsets, 8, head[1] + 1, index_in) != head[1] + 1)
        rc = sb200::fail_msg(SNAPPY_B200_ERR_CORRUPT, "truncated block index file");
    uint64_t n = 0;
    if (rc == SNAPPY_B200_OK && head[0] && !data.reserve(head[0]))
        rc = io_fail("out of host memory for the output");
    if (rc == SNAPPY_B200_OK && head[0]) {
        rc = snappy_b200_decompress_host_indexed(stream.data(), stream.size(), offsets, head[1], data.data(), data.cap, &n);
        if (rc == SNAPPY_B200_OK && n != head[0])
            rc = sb200::fail_msg(SNAPPY_B200_ERR_CORRUPT, "the index file declares another length than the stream");
        if (rc == SNAPPY_B200_OK && fwrite(data.data(), 1, n, out) != n)
            rc = io_fail("short write of the output");
    }
    free(offsets);
    return rc;
}

void snappy_compress(FILE *file_input, unsigned long long input_size, FILE *file_compressed)
{
    (void)compress_file(file_input, input_size, file_compressed, SNAPPY_B200_MODE_HASH);
}

int snappy_compress_bst(FILE *file_input, unsigned long long input_size, FILE *file_compressed)
{
    return compress_file(file_input, input_size, file_compressed, SNAPPY_B200_MODE_BST);
}

int snappy_decompress(FILE *file_input, FILE *file_decompressed)
{
    PinnedBuf stream, out;
    sb200::clear_error();
    if (!file_input || !file_decompressed || !read_all(file_input, 0, stream))
        return io_fail("reading the compressed stream failed");
    uint64_t total = 0;
    int rc = snappy_b200_uncompressed_length(stream.data(), stream.size(), &total);
    if (rc == SNAPPY_B200_OK && total) {
        if (!out.reserve(total))
            return io_fail("out of host memory for the output");
        uint64_t n = 0;
        rc = snappy_b200_decompress_host(stream.data(), stream.size(), out.data(), out.cap, &n);
        if (rc == SNAPPY_B200_OK && fwrite(out.data(), 1, n, file_decompressed) != n)
            rc = io_fail("short write of the output");
    }
    return rc;
}

// ---- varint.h (reference src/varint.c) ---------------------------------------------------
unsigned int parse_to_varint(unsigned long long n, unsigned char *varint)
{
    unsigned int k = 0;
    while (n >= 0x80u) {
        varint[k++] = (unsigned char)(n | 0x80u);
        n >>= 7;
    }
    varint[k++] = (unsigned char)n;
    return k;
}

int str_varint_to_dim_(unsigned char *varint)
{
    // the reference accumulates in an int (src/varint.c:44-58): keep its wrap-around
    unsigned int result = 0, mult = 1;
    unsigned char b;
    do {
        b = *varint++;
        result += (unsigned int)(b & 0x7fu) * mult;
        mult *= 128u;
    } while (b & 0x80u);
    return (int)result;
}

int varint_to_dim(FILE *source)
{
    unsigned int result = 0, mult = 1;
    unsigned char b = 0;
    do {
        if (fread(&b, 1, 1, source) != 1)
            break;
        result += (unsigned int)(b & 0x7fu) * mult;
        mult *= 128u;
    } while (b & 0x80u);
    return (int)result;
}

// ---- buffer_compression.h (reference src/buffer_compression.c) ----------------------------
void init_Buffer(Buffer *bf, unsigned int buffer_size)
{
    bf->current = (char *)calloc(buffer_size, 1);
    bf->beginning = bf->current;
    bf->bytes_left = buffer_size;
}

void move_current(Buffer *bf, unsigned int offset)
{
    bf->current += offset;
    bf->bytes_left -= offset;
}

void reset(Buffer *bf) { bf->current = bf->beginning; }

} // extern "C"
