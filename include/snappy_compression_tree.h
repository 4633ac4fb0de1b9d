/* Drop-in for the reference's src/snappy_compression_tree.h:10 (BST match finder).      */
#ifndef SNAPPY_B200_DROPIN_COMPRESSION_TREE_H
#define SNAPPY_B200_DROPIN_COMPRESSION_TREE_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Same contract as snappy_compress, exact-key dictionary path
 * (reference: src/snappy_compression_tree.c:291-306; its body has no return statement,
 * this one returns 0 on success and a negative SNAPPY_B200_ERR_* otherwise).            */
int snappy_compress_bst(FILE *file_input, unsigned long long input_size, FILE *file_compressed);
#ifdef __cplusplus
}
#endif
#endif
