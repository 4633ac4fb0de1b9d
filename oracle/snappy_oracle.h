/*
 * snappy_oracle.h -- interface of the CPU oracle (TEST INFRASTRUCTURE ONLY;
 * see the header of snappy_oracle.c for who may use it and how it is pinned).
 */
#ifndef SNAPPY_ORACLE_H
#define SNAPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_BLOCK_SIZE 65536u /* MAX_BLOCK_SIZE, src/snappy_compression.c:9 */
#define ORACLE_MAX_HTABLE 4096u  /* MAX_HTABLE_SIZE, src/snappy_compression.c:10 */

#define ORACLE_MODE_HASH 0 /* src/snappy_compression.c */
#define ORACLE_MODE_BST 1  /* src/snappy_compression_tree.c */

#define ORACLE_ERR_VARINT 1
#define ORACLE_ERR_CAPACITY 2
#define ORACLE_ERR_TRUNCATED 3
#define ORACLE_ERR_OFFSET 4
#define ORACLE_ERR_OVERRUN 5
#define ORACLE_ERR_FRAMING 6

unsigned oracle_varint_encode(uint64_t v, uint8_t *dst);
unsigned oracle_varint_decode(const uint8_t *src, size_t avail, uint64_t *out);
uint64_t oracle_max_compressed_size(uint64_t n);
uint64_t oracle_compress(const uint8_t *in, uint64_t n, uint8_t *out, int mode, uint32_t *block_sizes);
int oracle_decompress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *out_len);
int64_t oracle_block_index(const uint8_t *in, uint64_t n, uint64_t *offsets, uint64_t *total_out);

#ifdef __cplusplus
}
#endif
#endif
