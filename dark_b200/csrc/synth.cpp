// synth.cpp — synthetic block generators for benches and tests.
//
// The reference ships no data (its Makefile expects a git-ignored data/book1,
// /root/reference/Makefile:14-19), so the five BASELINE.json shapes are produced by these
// counter-based, integer-only generators: "text" (book1-like), "dna", "rep17", "mixed".
// Independent of oracle/gen.c; tests compare the two byte for byte.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dark_bwt.h"

namespace {

inline uint64_t mix(uint64_t x) {  // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline uint8_t hash_byte(uint64_t seed, uint64_t i) {
    return (uint8_t)(mix(seed * 0x100000001B3ull + (i >> 3)) >> ((i & 7) * 8));
}

struct Vocabulary {
    std::vector<std::string> words;
    Vocabulary() : words(4096) {
        for (uint64_t w = 0; w < 4096; ++w) {
            const unsigned len = 2 + (unsigned)(mix(77 + w) % 9);
            std::string& s = words[w];
            for (unsigned k = 0; k < len; ++k) s.push_back((char)('a' + mix(1000003ull * w + k) % 26));
        }
    }
};
const Vocabulary& vocabulary() {
    static const Vocabulary v;
    return v;
}

// Word stream: Zipf-like (log-uniform) word choice, ". " every ~64 words, newline past column 70.
void fill_text(uint8_t* out, uint64_t len, uint64_t seed) {
    const Vocabulary& voc = vocabulary();
    uint64_t pos = 0, counter = 0, column = 0;
    auto put = [&](char c) {
        if (pos < len) out[pos++] = (uint8_t)c;
    };
    while (pos < len) {
        const uint64_t r = mix(seed * 7919ull + counter++);
        const unsigned e = (unsigned)((r >> 8) % 12);
        const uint64_t w = ((uint64_t)1 << e) - 1 + ((r >> 16) & (((uint64_t)1 << e) - 1));
        const std::string& word = voc.words[w];
        for (char c : word) put(c);
        column += word.size() + 1;
        if ((r & 63) == 0) {
            put('.');
            put(' ');
        } else if (column > 70) {
            put('\n');
            column = 0;
        } else {
            put(' ');
        }
    }
}

void fill_dna(uint8_t* out, uint64_t n, uint64_t seed) {
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)"ACGT"[hash_byte(seed, i) & 3];
}

void fill_rep17(uint8_t* out, uint64_t n, uint64_t seed) {
    uint8_t pattern[17];
    for (int k = 0; k < 17; ++k) pattern[k] = (uint8_t)('a' + hash_byte(seed ^ 0xABCDull, (uint64_t)k) % 26);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t h = mix(seed + i * 0x9E37ull);
        out[i] = (h & 0xFFF) ? pattern[i % 17] : (uint8_t)(h >> 40);
    }
}

void fill_mixed(uint8_t* out, uint64_t n, uint64_t seed) {
    const uint64_t kSeg = 65536;
    for (uint64_t s = 0; s * kSeg < n; ++s) {
        const uint64_t off = s * kSeg;
        const uint64_t len = (n - off < kSeg) ? n - off : kSeg;
        uint8_t* seg = out + off;
        switch (mix(seed + s) % 3) {
            case 0: fill_text(seg, len, seed + s); break;
            case 1:
                for (uint64_t i = 0; i < len; ++i) seg[i] = hash_byte(seed + s, i);
                break;
            default:  // 16-byte records: LE32 record number, 4 zero bytes, 8 bytes with period 64 records
                for (uint64_t i = 0; i < len; ++i) {
                    const unsigned r = (unsigned)(i & 15);
                    const uint32_t rec = (uint32_t)((off + i) >> 4);
                    seg[i] = r < 4 ? (uint8_t)(rec >> (8 * r)) : r < 8 ? (uint8_t)0 : hash_byte(seed, ((i >> 4) & 63) * 16 + r);
                }
        }
    }
}

}  // namespace

extern "C" int dark_bwt_synth(const char* kind, uint64_t seed, uint8_t* out, uint64_t n) {
    if (!kind || (!out && n)) return DARK_BWT_E_INVALID_ARG;
    if (!strcmp(kind, "dna")) fill_dna(out, n, seed);
    else if (!strcmp(kind, "rep17")) fill_rep17(out, n, seed);
    else if (!strcmp(kind, "text")) fill_text(out, n, seed);
    else if (!strcmp(kind, "mixed")) fill_mixed(out, n, seed);
    else return DARK_BWT_E_INVALID_ARG;
    return DARK_BWT_OK;
}
