/* Drop-in for the reference's src/snappy_decompression.h:15.                            */
#ifndef SNAPPY_B200_DROPIN_DECOMPRESSION_H
#define SNAPPY_B200_DROPIN_DECOMPRESSION_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Reads a whole stream from file_input and writes the original bytes to
 * file_decompressed (reference: src/snappy_decompression.c:345-363, which always
 * returns 0).  Returns 0 on success; on a malformed stream nothing is written and a
 * negative SNAPPY_B200_ERR_* is returned (the reference has undefined behaviour there). */
int snappy_decompress(FILE *file_input, FILE *file_decompressed);
#ifdef __cplusplus
}
#endif
#endif
