/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h for the rules).
 *
 * CPU restatement of kvark/dark's forward BWT: `saca::Constructor` (SA-IS variant,
 * /root/reference/src/saca.rs) + the `compress::bwt::TransformIterator` emission loop
 * (call sites /root/reference/src/block/dc.rs:45-50, block/raw.rs:39-44).
 * Parity: PINNED by the reference's known-answer test saca.rs:409-413
 * (tests/test_oracle.py::test_reference_known_answers).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

typedef uint32_t suf_t;               /* saca.rs:20  type Suffix = u32 */
#define SUF_INVALID ((suf_t)0xFFFFFFFFu) /* saca.rs:22  SUF_INVALID = !0 */

static int saca_u32(const uint32_t *input, size_t n, size_t k, suf_t *storage, size_t storage_len, int depth,
                    oracle_trace *tr);

#define SYM uint32_t
#define FN(x) x##_u32
#include "saca_impl.inc"
#undef SYM
#undef FN

#define SYM uint8_t
#define FN(x) x##_u8
#include "saca_impl.inc"
#undef SYM
#undef FN

/* saca.rs:351-360  Constructor::new sizing. */
uint64_t oracle_arena_words(uint64_t max_n)
{
    uint64_t extra_2s = (1ull << 15) + (1ull << 7);
    uint64_t half = max_n / 2, quarter = max_n / 4;
    uint64_t mn = extra_2s < half ? extra_2s : half;
    uint64_t extra = 0x100 + (quarter > mn ? quarter : mn);
    return max_n + extra;
}

/* saca.rs:368-378  Constructor::compute on an existing arena. */
int oracle_saca_arena(const uint8_t *text, uint64_t n, uint32_t *arena, uint64_t arena_words, oracle_trace *trace)
{
    if (trace) memset(trace, 0, sizeof(*trace));
    if (n < 2) return ORACLE_E_LENGTH;          /* the reference panics: SURVEY §0.7 */
    if (n > 0xFFFFFFFEull) return ORACLE_E_LENGTH; /* Suffix = u32, SUF_INVALID = !0    */
    return saca_u8(text, (size_t)n, 0x100, arena, (size_t)arena_words, 0, trace);
}

int oracle_saca(const uint8_t *text, uint64_t n, uint32_t *sa, oracle_trace *trace)
{
    uint64_t words;
    uint32_t *arena;
    int rc;
    if (n < 2 || n > 0xFFFFFFFEull) return ORACLE_E_LENGTH;
    words = oracle_arena_words(n);
    arena = (uint32_t *)calloc(words, sizeof(uint32_t)); /* zeroed like iter::repeat(0) */
    if (!arena) return ORACLE_E_NOMEM;
    rc = oracle_saca_arena(text, n, arena, words, trace);
    if (rc == 0) memcpy(sa, arena, n * sizeof(uint32_t));
    free(arena);
    return rc;
}

/* compress::bwt::TransformIterator (third-party, un-vendored; pinned by saca.rs:411-412):
 * walks the suffix array; a zero suffix yields the LAST input byte and records origin. */
void oracle_bwt_emit(const uint8_t *text, uint64_t n, const uint32_t *sa, uint8_t *bwt, uint64_t *origin)
{
    uint64_t i;
    for (i = 0; i < n; i++) {
        uint32_t p = sa[i];
        if (p == 0) {
            *origin = i;
            bwt[i] = text[n - 1];
        } else {
            bwt[i] = text[p - 1];
        }
    }
}

/* block/dc.rs:45-50: compute, then collect the iterator, then get_origin. */
int oracle_bwt_forward(const uint8_t *text, uint64_t n, uint8_t *bwt, uint64_t *origin, uint32_t *sa_out)
{
    uint64_t words;
    uint32_t *arena;
    int rc;
    if (n < 2 || n > 0xFFFFFFFEull) return ORACLE_E_LENGTH;
    words = oracle_arena_words(n);
    arena = (uint32_t *)calloc(words, sizeof(uint32_t));
    if (!arena) return ORACLE_E_NOMEM;
    rc = oracle_saca_arena(text, n, arena, words, NULL);
    if (rc == 0) {
        oracle_bwt_emit(text, n, arena, bwt, origin);
        if (sa_out) memcpy(sa_out, arena, n * sizeof(uint32_t));
    }
    free(arena);
    return rc;
}

/* ---- saca.rs:25-35 sort_direct ---- */
static const uint8_t *g_sd_text;
static uint64_t g_sd_n;
static int sd_cmp(const void *pa, const void *pb)
{
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    uint64_t la = g_sd_n - a, lb = g_sd_n - b;
    uint64_t l = la < lb ? la : lb;
    int c = memcmp(g_sd_text + a, g_sd_text + b, l); /* unsigned bytes, like [u8]::cmp */
    if (c) return c;
    return la < lb ? -1 : (la > lb ? 1 : 0); /* a proper prefix sorts first */
}
void oracle_sort_direct(const uint8_t *text, uint64_t n, uint32_t *sa)
{
    uint64_t i;
    for (i = 0; i < n; i++) sa[i] = (uint32_t)i;
    g_sd_text = text;
    g_sd_n = n;
    qsort(sa, n, sizeof(uint32_t), sd_cmp); /* suffixes are pairwise distinct: stability irrelevant */
}

/* Inverse of the emission above.  The rows are the suffixes with end-of-text
 * lowest, i.e. the rotations of T$ minus the "$T" row; re-insert that row
 * (its last column is T[n-1] = bwt[origin]) and LF-walk from it. */
int oracle_bwt_decode(const uint8_t *bwt, uint64_t n, uint64_t origin, uint8_t *text_out)
{
    uint64_t count[257], i, row;
    uint32_t *lf;
    if (n == 0 || origin >= n) return ORACLE_E_LENGTH;
    lf = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    if (!lf) return ORACLE_E_NOMEM;
    memset(count, 0, sizeof(count));
    for (i = 0; i < n; i++) count[bwt[i] + 1]++;
    count[0] = 1; /* rows whose first column is '$' */
    for (i = 1; i < 257; i++) count[i] += count[i - 1];
    /* virtual last column: L[0] = bwt[origin]; L[k+1] = bwt[k] (k != origin); L[origin+1] = '$' */
    lf[0] = (uint32_t)count[bwt[origin]]++;
    for (i = 0; i < n; i++) {
        if (i == origin) lf[i + 1] = 0; /* '$' maps to row 0 */
        else lf[i + 1] = (uint32_t)count[bwt[i]]++;
    }
    row = 0;
    for (i = n; i-- > 0;) {
        text_out[i] = (row == 0) ? bwt[origin] : bwt[row - 1];
        row = lf[row];
    }
    free(lf);
    return row == origin + 1 ? 0 : ORACLE_E_ASSERT;
}

int oracle_verify_sa(const uint8_t *text, uint64_t n, const uint32_t *sa)
{
    uint32_t *isa;
    uint64_t j;
    int rc = 0;
    if (n == 0) return ORACLE_E_LENGTH;
    isa = (uint32_t *)malloc(n * sizeof(uint32_t));
    if (!isa) return ORACLE_E_NOMEM;
    memset(isa, 0xFF, n * sizeof(uint32_t));
    for (j = 0; j < n; j++) {
        if (sa[j] >= n || isa[sa[j]] != 0xFFFFFFFFu) { rc = ORACLE_E_ASSERT; goto done; }
        isa[sa[j]] = (uint32_t)j;
    }
    for (j = 1; j < n; j++) {
        uint64_t a = sa[j - 1], b = sa[j];
        if (text[a] < text[b]) continue;
        if (text[a] > text[b]) { rc = ORACLE_E_ASSERT; goto done; }
        {
            int64_t ra = (a + 1 < n) ? (int64_t)isa[a + 1] : -1;
            int64_t rb = (b + 1 < n) ? (int64_t)isa[b + 1] : -1;
            if (!(ra < rb)) { rc = ORACLE_E_ASSERT; goto done; }
        }
    }
done:
    free(isa);
    return rc;
}

/* ---- LCP profile and B_alg ---- */
int oracle_profile_lcp(const uint8_t *text, uint64_t n, const uint32_t *sa, oracle_profile *out)
{
    uint32_t *plcp; /* holds phi, then PLCP in place (Karkkainen-Manzini-Puglisi) */
    uint64_t i, j, l = 0, cnt[64], sum = 0;
    uint32_t prev_lcp = 0;
    double tot = 0.0;
    int r, top = 0;
    memset(out, 0, sizeof(*out));
    memset(cnt, 0, sizeof(cnt));
    if (n < 2) return ORACLE_E_LENGTH;
    plcp = (uint32_t *)malloc(n * sizeof(uint32_t));
    if (!plcp) return ORACLE_E_NOMEM;
    plcp[sa[0]] = 0xFFFFFFFFu;
    for (j = 1; j < n; j++) plcp[sa[j]] = sa[j - 1];
    for (i = 0; i < n; i++) {
        uint32_t phi = plcp[i];
        if (phi == 0xFFFFFFFFu) { plcp[i] = 0; l = 0; continue; }
        while (i + l < n && (uint64_t)phi + l < n && text[i + l] == text[phi + l]) l++;
        plcp[i] = (uint32_t)l;
        if (l > 0) l--;
    }
    /* v_j = max(LCP[j], LCP[j+1]) with LCP[0] = LCP[n] = 0; bucket k counts 8*2^k <= v < 8*2^(k+1) */
    prev_lcp = 0;
    for (j = 0; j < n; j++) {
        uint32_t cur = plcp[sa[j]]; /* LCP[j] = lcp(SA[j-1], SA[j]) */
        uint32_t nxt = (j + 1 < n) ? plcp[sa[j + 1]] : 0;
        uint32_t v = cur > nxt ? cur : nxt;
        (void)prev_lcp;
        tot += cur;
        if (cur > out->max_lcp) out->max_lcp = cur;
        if (v >= 8) {
            int k = 0;
            uint32_t t = v >> 3;
            while (t > 1) { t >>= 1; k++; }
            cnt[k]++;
            if (k + 1 > top) top = k + 1;
        }
    }
    free(plcp);
    /* m_r = #{v >= 8*2^(r-1)} = suffix sum of the buckets */
    for (r = top; r-- > 0;) {
        sum += cnt[r];
        out->m[r] = sum;
    }
    out->R = (uint32_t)top;
    out->sum_m = 0;
    for (r = 0; r < top; r++) out->sum_m += out->m[r];
    {
        uint32_t b = 0;
        while ((1ull << b) < n + 1) b++;
        out->b = b;
        out->P = (2 * b + 7) / 8;
    }
    out->mean_lcp = tot / (double)n;
    out->b_alg = 243.0 * (double)n + (48.0 + 24.0 * out->P) * (double)out->sum_m;
    return 0;
}
