/* Drop-in for the reference's src/buffer_compression.h:8-16: the three-field cursor the
 * reference's compressors keep over a byte array.  Kept for callers that use it; the batched block
 * scheduler of this build (csrc/host_pipeline.cu) keeps its own page-locked staging buffers.   */
#ifndef SNAPPY_B200_DROPIN_BUFFER_COMPRESSION_H
#define SNAPPY_B200_DROPIN_BUFFER_COMPRESSION_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct buffer {
    char *current;
    char *beginning;
    unsigned int bytes_left;
} Buffer;

/* src/buffer_compression.c:10-14: zero-filled array of buffer_size bytes (calloc, like the
 * reference; release it with free(bf->beginning)). */
void init_Buffer(Buffer *bf, unsigned int buffer_size);
/* src/buffer_compression.c:22-25 */
void move_current(Buffer *bf, unsigned int offset);
/* src/buffer_compression.c:32-34 (does not restore bytes_left, like the reference) */
void reset(Buffer *bf);
#ifdef __cplusplus
}
#endif
#endif
