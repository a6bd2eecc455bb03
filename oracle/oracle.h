/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the forward-BWT path of kvark/dark:
 *   - `saca::Constructor::{new,compute}`            (/root/reference/src/saca.rs:344-378)
 *   - `compress::bwt::TransformIterator` as called  (/root/reference/src/block/dc.rs:45-50,
 *                                                    /root/reference/src/block/raw.rs:39-44)
 * plus the synthetic generators of SURVEY.md App. D and the LCP profiler that
 * yields the algorithmic byte count B_alg of SURVEY.md §8(d).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or as
 * the timed CPU baseline.  Nothing under dark_b200/ links, imports or calls it.
 *
 * Parity status: PINNED for SA / BWT bytes / origin by the reference's own
 * known-answer test (saca.rs:409-413) and cross-checked against the reference's
 * built-in specification `sort_direct` (saca.rs:25-35); see tests/test_oracle.py.
 * The reference itself (Rust) cannot be compiled in this image: no rustc/cargo,
 * and its BWT emission lives in the un-vendored crate `compress = "0.1"`
 * (Cargo.toml:18) — upstream rust-compress; the emission loop below is that
 * crate's published behaviour, pinned by the same known-answer test.
 */
#ifndef DARK_ORACLE_H
#define DARK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_OK 0
#define ORACLE_E_ASSERT 1 /* one of the reference's assert!s would fire   */
#define ORACLE_E_ARENA 2  /* saca.rs:272 — arena too small                */
#define ORACLE_E_LENGTH 3 /* saca.rs:369 — input.len() != capacity; n < 2 */
#define ORACLE_E_NOMEM 4

#define ORACLE_MAX_DEPTH 32
typedef struct {
    int depth;                       /* recursion levels entered        */
    uint64_t n1[ORACLE_MAX_DEPTH];    /* LMS count per level             */
    uint64_t names[ORACLE_MAX_DEPTH]; /* distinct LMS names per level    */
} oracle_trace;

/* saca.rs:353-354: extra = 0x100 + max(n/4, min(2^15+2^7, n/2)) words. */
uint64_t oracle_arena_words(uint64_t max_n);

/* Constructor::new + compute: SA of text[0..n) into sa[0..n).
 * n == 0 and n == 1 return an error, as the reference panics. */
int oracle_saca(const uint8_t *text, uint64_t n, uint32_t *sa, oracle_trace *trace /* nullable */);

/* Same, on a caller-provided arena of oracle_arena_words(n) u32 (the SA is
 * arena[0..n) on return) — the form bench.py times, allocation excluded just as
 * Constructor::new is outside Encoder::encode. */
int oracle_saca_arena(const uint8_t *text, uint64_t n, uint32_t *arena, uint64_t arena_words,
                      oracle_trace *trace /* nullable */);

/* TransformIterator: bwt[i] = SA[i]==0 ? (origin=i, T[n-1]) : T[SA[i]-1]. */
void oracle_bwt_emit(const uint8_t *text, uint64_t n, const uint32_t *sa, uint8_t *bwt, uint64_t *origin);

/* compute + emission in one call (what block/dc.rs:45-50 does). */
int oracle_bwt_forward(const uint8_t *text, uint64_t n, uint8_t *bwt, uint64_t *origin,
                       uint32_t *sa_out /* nullable */);

/* saca.rs:25-35 `sort_direct`: the reference's own specification of the order
 * (slice comparison, shorter-is-smaller).  O(n^2 log n): small inputs only. */
void oracle_sort_direct(const uint8_t *text, uint64_t n, uint32_t *sa);

/* Inverse BWT (compress::bwt::decode as used in saca.rs:405): radix count ->
 * inversion table -> walk from origin.  Used for round-trip tests. */
int oracle_bwt_decode(const uint8_t *bwt, uint64_t n, uint64_t origin, uint8_t *text_out);

/* O(n) check that sa is the suffix array of text (permutation + order via ISA). */
int oracle_verify_sa(const uint8_t *text, uint64_t n, const uint32_t *sa);

/* ---- SURVEY.md App. D generators (integer-only, counter-based) ---- */
uint64_t oracle_sm64(uint64_t x);
void oracle_gen_dna(uint64_t seed, uint8_t *out, uint64_t n);
void oracle_gen_rep17(uint64_t seed, uint8_t *out, uint64_t n);
void oracle_gen_text(uint64_t seed, uint8_t *out, uint64_t n);
void oracle_gen_mixed(uint64_t seed, uint8_t *out, uint64_t n);
/* kind: "dna" | "rep17" | "text" | "mixed"; returns non-zero for an unknown kind */
int oracle_gen(const char *kind, uint64_t seed, uint8_t *out, uint64_t n);

/* ---- SURVEY.md §8(d) profiler ---- */
typedef struct {
    uint32_t b;        /* ceil(log2(N+1))                                   */
    uint32_t P;        /* ceil(2b/8) sort passes per doubling round         */
    uint32_t R;        /* rounds r >= 1 with m_r > 0 (h_r = 8 * 2^(r-1))    */
    uint64_t m[64];    /* m[r-1] = #suffixes not unique by their first h_r  */
    uint64_t sum_m;    /* sum over r                                        */
    uint64_t max_lcp;
    double mean_lcp;
    double b_alg;      /* 243 N + (48 + 24 P) * sum_m  (bytes)              */
} oracle_profile;
int oracle_profile_lcp(const uint8_t *text, uint64_t n, const uint32_t *sa, oracle_profile *out);

/* ---- distance coding + MTF (dc_oracle.c; third-party `compress::bwt::dc`, PARITY UNPINNED: restated from memory of
 * upstream rust-compress, no reference-held vector exists; call sites /root/reference/src/block/dc.rs:52,54-85,146) ---- */
int oracle_dc_encode(const uint8_t *input, uint64_t n, uint32_t *distances, uint64_t init[256], uint8_t mtf_symbols[256],
                     uint32_t *num_unique_out);
uint64_t oracle_dc_stream(const uint8_t *input, const uint32_t *distances, uint64_t n, const uint64_t init[256], uint32_t *out_pos,
                          uint32_t *out_dist, uint8_t *out_sym, uint8_t *out_rank);
int oracle_dc_decode(const uint64_t init[256], const uint32_t *stream_dist, uint64_t count, uint8_t *output, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
