/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * The synthetic input generators of SURVEY.md Appendix D (normative spec,
 * integer-only, counter-based splitmix64).  The shapes stand in for the five
 * BASELINE.json configs: text (book1-sized), dna, rep17, mixed.
 * Regression values (CRC-32 of prefixes) are checked in tests/test_oracle.py.
 */
#include "oracle.h"

#include <string.h>

uint64_t oracle_sm64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* hb(seed,i) = byte (i & 7), little-endian, of sm64(seed * 0x100000001B3 + (i >> 3)) */
static inline uint8_t hb(uint64_t seed, uint64_t i)
{
    uint64_t w = oracle_sm64(seed * 0x100000001B3ull + (i >> 3));
    return (uint8_t)(w >> (8 * (i & 7)));
}

void oracle_gen_dna(uint64_t seed, uint8_t *out, uint64_t n)
{
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    uint64_t i = 0;
    while (i < n) {
        uint64_t w = oracle_sm64(seed * 0x100000001B3ull + (i >> 3));
        int k;
        for (k = 0; k < 8 && i < n; k++, i++) out[i] = (uint8_t)acgt[(w >> (8 * k)) & 3];
    }
}

void oracle_gen_rep17(uint64_t seed, uint8_t *out, uint64_t n)
{
    uint8_t pat[17];
    uint64_t i;
    int k;
    for (k = 0; k < 17; k++) pat[k] = (uint8_t)('a' + hb(seed ^ 0xABCDull, (uint64_t)k) % 26);
    for (i = 0; i < n; i++) {
        uint64_t h = oracle_sm64(seed + i * 0x9E37ull);
        out[i] = ((h & 0xFFF) == 0) ? (uint8_t)((h >> 40) & 0xFF) : pat[i % 17];
    }
}

static uint8_t g_vocab[4096][12];
static uint8_t g_vocab_len[4096];
static int g_vocab_ready = 0;
static void vocab_init(void)
{
    uint64_t w;
    if (g_vocab_ready) return;
    for (w = 0; w < 4096; w++) {
        unsigned len = 2 + (unsigned)(oracle_sm64(77 + w) % 9), k;
        g_vocab_len[w] = (uint8_t)len;
        for (k = 0; k < len; k++) g_vocab[w][k] = (uint8_t)('a' + oracle_sm64(1000003ull * w + k) % 26);
    }
    g_vocab_ready = 1;
}

static void text_seg(uint8_t *out, uint64_t len, uint64_t seed)
{
    uint64_t pos = 0, c = 0, col = 0;
    vocab_init();
#define EMIT(ch) do { if (pos < len) out[pos++] = (uint8_t)(ch); } while (0)
    while (pos < len) {
        uint64_t r = oracle_sm64(seed * 7919ull + c++);
        unsigned e = (unsigned)((r >> 8) % 12);
        uint64_t w = (1ull << e) - 1 + ((r >> 16) & ((1ull << e) - 1));
        unsigned k, wl = g_vocab_len[w];
        for (k = 0; k < wl; k++) EMIT(g_vocab[w][k]);
        col += wl + 1;
        if ((r & 63) == 0) {
            EMIT('.');
            EMIT(' ');
        } else if (col > 70) {
            EMIT('\n');
            col = 0;
        } else {
            EMIT(' ');
        }
    }
#undef EMIT
}

void oracle_gen_text(uint64_t seed, uint8_t *out, uint64_t n) { text_seg(out, n, seed); }

void oracle_gen_mixed(uint64_t seed, uint8_t *out, uint64_t n)
{
    uint64_t s, off;
    for (s = 0, off = 0; off < n; s++, off += 65536) {
        uint64_t len = n - off < 65536 ? n - off : 65536, i;
        uint8_t *seg = out + off;
        unsigned kind = (unsigned)(oracle_sm64(seed + s) % 3);
        if (kind == 0) {
            text_seg(seg, len, seed + s);
        } else if (kind == 1) {
            for (i = 0; i < len; i++) seg[i] = hb(seed + s, i);
        } else {
            for (i = 0; i < len; i++) {
                unsigned r = (unsigned)(i % 16);
                uint32_t rec = (uint32_t)((off + i) / 16);
                if (r < 4) seg[i] = (uint8_t)(rec >> (8 * r));
                else if (r < 8) seg[i] = 0;
                else seg[i] = hb(seed, ((i / 16) % 64) * 16 + r);
            }
        }
    }
}

int oracle_gen(const char *kind, uint64_t seed, uint8_t *out, uint64_t n)
{
    if (!strcmp(kind, "dna")) oracle_gen_dna(seed, out, n);
    else if (!strcmp(kind, "rep17")) oracle_gen_rep17(seed, out, n);
    else if (!strcmp(kind, "text")) oracle_gen_text(seed, out, n);
    else if (!strcmp(kind, "mixed")) oracle_gen_mixed(seed, out, n);
    else return 1;
    return 0;
}
