/*
 * ref_shim.c -- memory-to-memory driver around the UNMODIFIED reference
 * (TEST INFRASTRUCTURE ONLY).  It is compiled together with the reference's
 * own sources, taken where they lie under /root/reference/src, into
 * oracle/_ref/libsnappy_ref.so by oracle/Makefile.  No reference source is
 * copied into this repository; this file only calls the reference's public
 * FILE*-based entry points (src/snappy_compression.h:8,
 * src/snappy_compression_tree.h:10, src/snappy_decompression.h:15) on
 * fmemopen()/open_memstream() streams so that tests and the CPU-baseline leg
 * of bench.py can drive it on buffers.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void snappy_compress(FILE *file_input, unsigned long long input_size, FILE *file_compressed);
int snappy_compress_bst(FILE *file_input, unsigned long long input_size, FILE *file_compressed);
int snappy_decompress(FILE *file_input, FILE *file_decompressed);

static FILE *open_input(const uint8_t *in, uint64_t n)
{
    if (n == 0)
        return tmpfile(); /* fmemopen rejects zero-length buffers */
    return fmemopen((void *)in, (size_t)n, "rb");
}

/* mode 0: snappy_compress (hash table); mode 1: snappy_compress_bst.
 * Returns the stream length, or (uint64_t)-1 if it does not fit in cap.      */
uint64_t ref_compress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, int mode)
{
    FILE *fi = open_input(in, n);
    char *mem = NULL;
    size_t mem_len = 0;
    FILE *fo = open_memstream(&mem, &mem_len);
    if (!fi || !fo)
        return (uint64_t)-1;
    if (mode == 0)
        snappy_compress(fi, n, fo);
    else
        snappy_compress_bst(fi, n, fo);
    fclose(fi);
    fclose(fo);
    uint64_t r = mem_len;
    if (mem_len <= cap)
        memcpy(out, mem, mem_len);
    else
        r = (uint64_t)-1;
    free(mem);
    return r;
}

/* Runs the reference decoder.  Returns the number of bytes it wrote.         */
uint64_t ref_decompress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap)
{
    FILE *fi = open_input(in, n);
    char *mem = NULL;
    size_t mem_len = 0;
    FILE *fo = open_memstream(&mem, &mem_len);
    if (!fi || !fo)
        return (uint64_t)-1;
    snappy_decompress(fi, fo);
    fclose(fi);
    fclose(fo);
    uint64_t r = mem_len;
    if (mem_len <= cap)
        memcpy(out, mem, mem_len);
    else
        r = (uint64_t)-1;
    free(mem);
    return r;
}
