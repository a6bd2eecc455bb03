/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Command-line front end: generate a SURVEY App. D input, run the saca.rs
 * restatement + emission, print one JSON line with origin, CRC-32s, timings,
 * the SA-IS recursion trace and (with --profile) the §8(d) LCP profile/B_alg.
 * Used by tests/golden/make_golden.py to produce the committed fixtures.
 *
 *   oracle_cli <dna|rep17|text|mixed> <seed> <n> [--profile] [--dump-bwt FILE]
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static uint32_t crc_table[256];
static void crc_init(void)
{
    uint32_t i, j, c;
    for (i = 0; i < 256; i++) {
        c = i;
        for (j = 0; j < 8; j++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[i] = c;
    }
}
static uint32_t crc32_buf(const void *p, uint64_t len)
{
    const uint8_t *b = (const uint8_t *)p;
    uint32_t c = 0xFFFFFFFFu;
    uint64_t i;
    for (i = 0; i < len; i++) c = crc_table[(c ^ b[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    const char *kind, *dump = NULL;
    uint64_t seed, n, words, origin = 0;
    int profile = 0, i, rc;
    uint8_t *text, *bwt;
    uint32_t *arena;
    oracle_trace tr;
    double t0, t1, t2;

    if (argc < 4) {
        fprintf(stderr, "usage: %s <dna|rep17|text|mixed> <seed> <n> [--profile] [--dump-bwt FILE]\n", argv[0]);
        return 2;
    }
    kind = argv[1];
    seed = strtoull(argv[2], NULL, 0);
    n = strtoull(argv[3], NULL, 0);
    for (i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--profile")) profile = 1;
        else if (!strcmp(argv[i], "--dump-bwt") && i + 1 < argc) dump = argv[++i];
    }
    crc_init();
    text = (uint8_t *)malloc(n);
    bwt = (uint8_t *)malloc(n);
    words = oracle_arena_words(n);
    arena = (uint32_t *)calloc(words, 4);
    if (!text || !bwt || !arena) { fprintf(stderr, "out of memory\n"); return 1; }
    if (oracle_gen(kind, seed, text, n)) { fprintf(stderr, "unknown kind %s\n", kind); return 2; }

    t0 = now_s();
    rc = oracle_saca_arena(text, n, arena, words, &tr);
    t1 = now_s();
    if (rc) { fprintf(stderr, "oracle_saca failed: %d\n", rc); return 1; }
    oracle_bwt_emit(text, n, arena, bwt, &origin);
    t2 = now_s();

    printf("{\"kind\": \"%s\", \"seed\": %llu, \"n\": %llu, \"origin\": %llu, ", kind, (unsigned long long)seed,
           (unsigned long long)n, (unsigned long long)origin);
    printf("\"text_crc32\": \"%08x\", \"bwt_crc32\": \"%08x\", \"sa_crc32\": \"%08x\", ", crc32_buf(text, n),
           crc32_buf(bwt, n), crc32_buf(arena, n * 4));
    printf("\"saca_s\": %.4f, \"emit_s\": %.4f, \"mb_per_s\": %.3f, ", t1 - t0, t2 - t1, n / 1e6 / (t2 - t0));
    printf("\"sais_levels\": [");
    for (i = 0; i < tr.depth; i++)
        printf("%s[%llu, %llu]", i ? ", " : "", (unsigned long long)tr.n1[i], (unsigned long long)tr.names[i]);
    printf("]");
    if (profile) {
        oracle_profile pf;
        rc = oracle_profile_lcp(text, n, arena, &pf);
        if (rc) { fprintf(stderr, "profile failed: %d\n", rc); return 1; }
        printf(", \"profile\": {\"b\": %u, \"P\": %u, \"R\": %u, \"max_lcp\": %llu, \"mean_lcp\": %.3f, \"sum_m\": %llu, "
               "\"b_alg\": %.0f, \"b_alg_per_byte\": %.3f, \"m\": [",
               pf.b, pf.P, pf.R, (unsigned long long)pf.max_lcp, pf.mean_lcp, (unsigned long long)pf.sum_m, pf.b_alg,
               pf.b_alg / (double)n);
        for (i = 0; i < (int)pf.R; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)pf.m[i]);
        printf("]}");
    }
    printf("}\n");
    if (dump) {
        FILE *f = fopen(dump, "wb");
        if (f) { fwrite(bwt, 1, n, f); fclose(f); }
    }
    free(text); free(bwt); free(arena);
    return 0;
}
