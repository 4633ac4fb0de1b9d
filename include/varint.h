/* Drop-in for the reference's src/varint.h:4-6 (plain LEB128).            */
#ifndef SNAPPY_B200_DROPIN_VARINT_H
#define SNAPPY_B200_DROPIN_VARINT_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
/* src/varint.c:12-20: writes n as LEB128 at varint, returns the byte count (1..10). */
unsigned int parse_to_varint(unsigned long long n, unsigned char *varint);
/* src/varint.c:28-42: reads one varint from the file (the reference's int return type is
 * kept, so values above 2^31-1 wrap exactly as in the reference).                     */
int varint_to_dim(FILE *source);
/* src/varint.c:44-58: same, from memory. */
int str_varint_to_dim_(unsigned char *varint);
#ifdef __cplusplus
}
#endif
#endif
