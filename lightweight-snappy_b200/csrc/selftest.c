/* selftest.c -- round-trip driver in the shape of the reference's src/snappy_test.c:66-104
 * (compress, decompress, check integrity, print a report per file), over GENERATED fixtures:
 * the reference's Files_test/ directory is not in its repository.
 *
 *     snappy_b200_test [workdir]        (default: a fresh directory under $TMPDIR or /tmp)
 *
 * Fixtures: the six named ones the reference lists (src/snappy_test.c:8-13), rebuilt from their
 * names -- 32 KiB of 0xff, 32 KiB of random bytes, an English-like text, an empty file, three
 * 0xff bytes, a synthetic "image" (smooth gradients + noise) -- and, like the reference's
 * `dim[]` loop (:7, :92-103), five files for each of thirteen sizes from 500 B to 1 MB, cycling
 * through text / runs / random / mixed / periodic content.  Every file goes through the
 * FILE*-based drop-in API (snappy_compress and snappy_compress_bst, then snappy_decompress);
 * integrity is a memcmp of the whole file (the reference's compare_files mis-handles 0xff and
 * length mismatches).  Exit status 0 only if every file round-trips.           */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "snappy_b200.h"
#include "snappy_compression.h"
#include "snappy_compression_tree.h"
#include "snappy_decompression.h"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd(void) /* splitmix64, top 32 bits */
{
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

enum { TEXT, RUNS, RANDOM, MIXED, PERIODIC, IMAGE, FILL_FF };

static void fill(unsigned char *p, size_t n, int kind)
{
    static const char *words[] = {"the", "of", "and", "a", "to", "in", "is", "you", "that", "it", "he", "was", "for",
                                  "on", "are", "as", "with", "his", "they", "I", "at", "be", "this", "have", "from",
                                  "alice", "rabbit", "queen", "said", "little", "curious", "wonderland"};
    size_t i = 0;
    switch (kind) {
    case TEXT:
        while (i < n) {
            const char *w = words[(rnd() % 32) * (rnd() % 3 ? 1 : 0) % 32];
            for (size_t k = 0; w[k] && i < n; ++k)
                p[i++] = (unsigned char)w[k];
            if (i < n)
                p[i++] = rnd() % 12 ? ' ' : '\n';
        }
        break;
    case RUNS:
        while (i < n) {
            const unsigned char sym = (unsigned char)("\0\1\2\377"[rnd() % 4]);
            for (uint32_t r = 1 + rnd() % 15; r && i < n; --r)
                p[i++] = sym;
        }
        break;
    case RANDOM:
        for (; i < n; ++i)
            p[i] = (unsigned char)rnd();
        break;
    case MIXED:
        while (i < n) {
            const size_t len = 1 + rnd() % 4096, m = len < n - i ? len : n - i;
            fill(p + i, m, (int)(rnd() % 3));
            i += m;
        }
        break;
    case PERIODIC: {
        const uint32_t period = 1 + rnd() % 300;
        for (; i < n; ++i)
            p[i] = i < period ? (unsigned char)rnd() : p[i - period];
        break;
    }
    case IMAGE:
        for (; i < n; ++i)
            p[i] = (unsigned char)((i / 3 % 640) / 3 + (i / 1920) / 2 + rnd() % 3);
        break;
    default:
        memset(p, 0xff, n);
    }
}

static int write_file(const char *name, const unsigned char *p, size_t n)
{
    FILE *f = fopen(name, "wb");
    if (!f)
        return -1;
    const size_t w = n ? fwrite(p, 1, n, f) : 0;
    return fclose(f) == 0 && w == n ? 0 : -1;
}

static unsigned char *read_file(const char *name, size_t *n)
{
    FILE *f = fopen(name, "rb");
    if (!f)
        return NULL;
    fseek(f, 0, SEEK_END);
    const long s = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *p = malloc(s > 0 ? (size_t)s : 1);
    *n = s > 0 ? fread(p, 1, (size_t)s, f) : 0;
    fclose(f);
    return p;
}

/* one direction, FILE* in / FILE* out, like run_test_mode (src/snappy_test.c:24-57) */
static int run_mode(const char *in_name, const char *out_name, int mode, double *seconds)
{
    FILE *in = fopen(in_name, "rb"), *out = fopen(out_name, "wb");
    if (!in || !out) {
        fprintf(stderr, "cannot open %s / %s: %s\n", in_name, out_name, strerror(errno));
        exit(EXIT_FAILURE);
    }
    fseek(in, 0, SEEK_END);
    const unsigned long long size = (unsigned long long)ftell(in);
    fseek(in, 0, SEEK_SET);
    int rc = 0;
    const double t0 = now();
    if (mode == 0)
        snappy_compress(in, size, out);
    else if (mode == 1)
        rc = snappy_compress_bst(in, size, out);
    else
        rc = snappy_decompress(in, out);
    *seconds = now() - t0;
    fclose(in);
    fclose(out);
    return rc != 0 || snappy_b200_last_error()[0];
}

static int n_run = 0, n_failed = 0;

static void run_test(const char *dir, const char *name, size_t n, int kind)
{
    char fin[512], fcomp[512], fdec[512];
    snprintf(fin, sizeof fin, "%s/%s", dir, name);
    unsigned char *data = malloc(n ? n : 1);
    fill(data, n, kind);
    if (write_file(fin, data, n)) {
        fprintf(stderr, "cannot write %s\n", fin);
        exit(EXIT_FAILURE);
    }
    for (int mode = 0; mode < 2; ++mode) {
        snprintf(fcomp, sizeof fcomp, "%s/%s.%s", dir, name, mode ? "bsnp" : "snp");
        snprintf(fdec, sizeof fdec, "%s/%s.%s.dec", dir, name, mode ? "bsnp" : "snp");
        double tc = 0, td = 0;
        int bad = run_mode(fin, fcomp, mode, &tc);
        size_t nc = 0, nd = 0;
        unsigned char *comp = read_file(fcomp, &nc);
        /* an empty input gives an empty stream, which is not decodable */
        unsigned char *dec = NULL;
        if (!bad && n) {
            bad = run_mode(fcomp, fdec, 2, &td);
            dec = read_file(fdec, &nd);
            if (!bad)
                bad = nd != n || memcmp(dec, data, n) != 0;
        } else if (!bad) {
            bad = nc != 0;
        }
        printf("%-16s %-4s %9zu -> %9zu B  ratio %6.3f  comp %8.1f MB/s  decomp %8.1f MB/s  %s\n", name,
               mode ? "bst" : "hash", n, nc, nc ? (double)n / (double)nc : 0.0, tc > 0 ? n / tc / 1e6 : 0.0,
               td > 0 ? n / td / 1e6 : 0.0, bad ? "FAILED" : "ok");
        if (bad && snappy_b200_last_error()[0])
            printf("    last error: %s\n", snappy_b200_last_error());
        ++n_run;
        n_failed += bad;
        free(comp);
        free(dec);
        remove(fcomp);
        remove(fdec);
    }
    free(data);
    remove(fin);
}

int main(int argc, char **argv)
{
    char dir[400];
    if (argc > 1) {
        snprintf(dir, sizeof dir, "%s", argv[1]);
        mkdir(dir, 0777);
    } else {
        const char *tmp = getenv("TMPDIR");
        snprintf(dir, sizeof dir, "%s/snappy_b200_test.XXXXXX", tmp ? tmp : "/tmp");
        if (!mkdtemp(dir)) {
            fprintf(stderr, "cannot create %s: %s\n", dir, strerror(errno));
            return EXIT_FAILURE;
        }
    }
    if (snappy_b200_device_count() <= 0) {
        fprintf(stderr, "snappy_b200_test: no CUDA device (there is no CPU fallback)\n");
        return EXIT_FAILURE;
    }
    /* the reference's named fixtures (src/snappy_test.c:8-13) */
    run_test(dir, "32k_ff", 32768, FILL_FF);
    run_test(dir, "32k_random", 32768, RANDOM);
    run_test(dir, "alice.txt", 152089, TEXT);
    run_test(dir, "empty", 0, RANDOM);
    run_test(dir, "ff_ff_ff", 3, FILL_FF);
    run_test(dir, "immagine.tiff", 921654, IMAGE);
    /* five files for each size (src/snappy_test.c:7, :92-103) */
    static const unsigned dim[] = {500, 1000, 2000, 5000, 10000, 20000, 50000, 80000, 100000, 200000, 500000, 800000, 1000000};
    for (unsigned i = 0; i < sizeof dim / sizeof dim[0]; ++i)
        for (int j = 1; j <= 5; ++j) {
            char name[64];
            snprintf(name, sizeof name, "%ub%d", dim[i], j);
            run_test(dir, name, dim[i], j - 1);
        }
    if (argc <= 1)
        rmdir(dir);
    printf("%d round trips, %d failed\n", n_run, n_failed);
    return n_failed ? EXIT_FAILURE : EXIT_SUCCESS;
}
