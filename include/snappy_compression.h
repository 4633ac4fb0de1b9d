/* Drop-in for the reference's src/snappy_compression.h:8 -- same symbol, same signature,
 * same stream bytes; the per-block work runs on the GPU (see snappy_b200.h).            */
#ifndef SNAPPY_B200_DROPIN_COMPRESSION_H
#define SNAPPY_B200_DROPIN_COMPRESSION_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Reads file_input from its current position to EOF, appends the hash-table-path stream
 * to file_compressed.  input_size is trusted and only used for the varint preamble
 * (reference: src/snappy_compression.c:414-428).  Errors: snappy_b200_last_error().    */
void snappy_compress(FILE *file_input, unsigned long long input_size, FILE *file_compressed);
#ifdef __cplusplus
}
#endif
#endif
