/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
 *
 * CPU restatement of distance coding + move-to-front as the reference calls it:
 *     bwt::dc::encode(&output, suf, &mut self.mtf)     (/root/reference/src/block/dc.rs:52)
 *     dc_iter.get_init(), for (d, ctx) in dc_iter      (/root/reference/src/block/dc.rs:54-85)
 *     bwt::dc::decode(init, &mut out, &mut mtf, |ctx|) (/root/reference/src/block/dc.rs:146-150)
 * The code lives in the third-party crate `compress = "0.1"` (Cargo.toml:18; upstream rust-compress, modules
 * bwt::dc and bwt::mtf), which is NOT in /root/reference and cannot be fetched here, and the reference holds no
 * known answer for it (its DC tests are round trips only, block/dc.rs:187-192).  What follows restates the upstream
 * algorithm as recalled:
 *   encode: distances[] starts as the filler n.  Scanning left to right with an MTF list, at every occurrence i of
 *           a symbol whose previous occurrence is `base`: rank = mtf.encode(sym) = number of distinct symbols seen
 *           since `base`; if rank > 0, distances[base] = i - base - rank - 1.  A first occurrence records
 *           init[sym] = i.  At the end every symbol's last occurrence gets n - base - rank - 1 with its final rank.
 *   iterate: every position whose distance is not the filler yields (distance, Context{symbol, last_rank,
 *           distance_limit = n - i}), last_rank = last_active - pos[symbol] (the rank a decoder would know).
 * The exact constants (the "- 1"s) are from memory; the tests therefore check self-consistency (encode -> decode
 * restores the input) and the GPU against THIS restatement, not against the crate.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct {
    uint8_t symbols[256];
} mtf_t;

/* compress::bwt::mtf::MTF::encode: rank of sym, sym moved to the front */
static unsigned mtf_encode(mtf_t *m, uint8_t sym) {
    uint8_t next = m->symbols[0];
    if (next == sym) return 0;
    unsigned rank = 1;
    for (;;) {
        uint8_t t = m->symbols[rank];
        m->symbols[rank] = next;
        next = t;
        if (next == sym) break;
        ++rank;
    }
    m->symbols[0] = sym;
    return rank;
}

int oracle_dc_encode(const uint8_t *input, uint64_t n, uint32_t *distances, uint64_t init[256], uint8_t mtf_symbols[256],
                     uint32_t *num_unique_out) {
    if (n >= 0xFFFFFFFFull) return ORACLE_E_LENGTH;
    mtf_t mtf;
    memset(&mtf, 0, sizeof(mtf));
    uint64_t last[256];
    unsigned num_unique = 0;
    for (int c = 0; c < 256; ++c) last[c] = n, init[c] = n;
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t sym = input[i];
        distances[i] = (uint32_t)n; /* filler */
        const uint64_t base = last[sym];
        last[sym] = i;
        if (base == n) {
            const unsigned rank = num_unique;
            mtf.symbols[rank] = sym;
            mtf_encode(&mtf, sym); /* == rank */
            init[sym] = i;
            ++num_unique;
        } else {
            const unsigned rank = mtf_encode(&mtf, sym);
            if (rank > 0) {
                if (i < base + rank + 1) return ORACLE_E_ASSERT;
                distances[base] = (uint32_t)(i - base - rank - 1);
            }
        }
    }
    for (unsigned rank = 0; rank < num_unique; ++rank) {
        const uint8_t sym = mtf.symbols[rank];
        const uint64_t base = last[sym];
        if (n < base + rank + 1) return ORACLE_E_ASSERT;
        distances[base] = (uint32_t)(n - base - rank - 1);
    }
    memcpy(mtf_symbols, mtf.symbols, 256);
    *num_unique_out = num_unique;
    return ORACLE_OK;
}

/* EncodeIterator: the (distance, Context) stream.  Outputs have room for n entries; returns the count. */
uint64_t oracle_dc_stream(const uint8_t *input, const uint32_t *distances, uint64_t n, const uint64_t init[256], uint32_t *out_pos,
                          uint32_t *out_dist, uint8_t *out_sym, uint8_t *out_rank) {
    uint64_t pos[256];
    memcpy(pos, init, sizeof(pos));
    uint64_t last_active = 0, cnt = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (distances[i] == (uint32_t)n) continue;
        const uint8_t sym = input[i];
        const uint64_t rank = last_active - pos[sym];
        last_active = i + 1;
        pos[sym] = i + 1 + distances[i];
        out_pos[cnt] = (uint32_t)i; /* distance_limit = n - i */
        out_dist[cnt] = distances[i];
        out_sym[cnt] = sym;
        out_rank[cnt] = (uint8_t)rank;
        ++cnt;
    }
    return cnt;
}

/* dc::decode restated to match the encoder above: rebuilds the block from init[] and the distance stream (in order).
 * next[sym] = position of the next occurrence of sym (n + rank at the end = none left). */
int oracle_dc_decode(const uint64_t init[256], const uint32_t *stream_dist, uint64_t count, uint8_t *output, uint64_t n) {
    uint64_t next[256];
    uint8_t order[256];
    unsigned alphabet = 0;
    memcpy(next, init, sizeof(next));
    for (int c = 0; c < 256; ++c) {
        if (next[c] < n) { /* insertion sort by first occurrence */
            unsigned j = alphabet;
            while (j > 0 && next[order[j - 1]] > next[c]) {
                order[j] = order[j - 1];
                --j;
            }
            order[j] = (uint8_t)c;
            ++alphabet;
        }
    }
    if (alphabet == 0) return n == 0 ? ORACLE_OK : ORACLE_E_ASSERT;
    uint64_t i = 0, used = 0;
    while (i < n) {
        const uint8_t sym = order[0];
        const uint64_t stop = alphabet > 1 ? (next[order[1]] < n ? next[order[1]] : n) : n;
        while (i < stop) output[i++] = sym;
        /* the run of sym ended at i-1: its distance tells where sym comes next */
        if (used >= count) return ORACLE_E_ASSERT;
        const uint64_t d = stream_dist[used++];
        /* encoder: d = i_next - base - r - 1 with base = i-1 and r = the number of symbols that come before i_next.  Walking
         * the order list (sorted by next occurrence), every symbol met before the target pushes the target one further;
         * the walk stops at index r+1, so i_next = future + rank - 1 (as upstream's `next[sym] = future+rank-1`). */
        const uint64_t future = i + d;
        unsigned rank = 1;
        while (rank < alphabet && future + rank > next[order[rank]]) {
            order[rank - 1] = order[rank];
            ++rank;
        }
        order[rank - 1] = sym;
        next[sym] = future + rank - 1; /* >= n: no further occurrence */
    }
    return used == count ? ORACLE_OK : ORACLE_E_ASSERT;
}
