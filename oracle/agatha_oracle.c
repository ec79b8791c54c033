/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See agatha_oracle.h for the rules on who may call this.
 *
 * Scalar CPU restatement of readwrite112/AGAThA's alignment kernel. Every function cites the reference
 * lines (relative to /root/reference/AGAThA/src/) it restates. It is written anti-diagonal-major
 * (so it can stop exactly where the reference stops) instead of in the reference's block/slice order;
 * DP values do not depend on evaluation order, only the scan schedule does, and that is replayed
 * literally in run_pair().
 *
 * Parity pin: checked against the reference header compiled as host code (oracle/_ref, see
 * oracle/Makefile and tests/test_oracle_vs_ref.py) and against tests/golden/ vectors made from it.
 *
 * Semantics, exact for band widths W == 7 (mod 8); for other W this file
 * implements the strict band |r-q| <= W, which the reference does not (A.3).
 *  - arithmetic is int32 with the reference's -16384 "minus infinity"; the reference's int16
 *    storage wrap-around (Appendix C) is NOT modelled: outside its valid domain the oracle is the
 *    arbiter.
 */
#include "agatha_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG16 (-16384)   /* MINUS_INF2 = SHRT_MIN/2, gasal_kernels.h:38-39 */
#define N_NIBBLE 14      /* N_CODE 0x4E & 0xF, Makefile:4, gasal_kernels.h:41 */
#define N_PENALTY 1      /* Makefile:5 */

static int g_model_phantom = 1;

void agatha_oracle_set_model(int32_t flags) { g_model_phantom = (flags & 1) ? 1 : 0; }

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* DEV_GET_SUB_SCORE_GLOBAL with N_PENALTY defined, gasal_kernels.h:48-50.
 * Base code is ascii & 15 (pack_rc_seqs.h:24-31). */
static inline int sub_score(int a, int b, const agatha_oracle_params_t *p)
{
    int s = (a == b) ? p->match : -p->mismatch;
    if (a == N_NIBBLE || b == N_NIBBLE) s = -N_PENALTY;
    return s;
}

/* Initial contents of the horizontal/vertical strips, agatha_kernel.h:126-148.
 * edge_h(j) is H(-1,j) == H(j,-1); edge_g(j) is F(0,j) == E(j,0). */
static inline int edge_h(int j, int goe, int ge, int W) { return j <= W ? -(goe + ge * j) : NEG16; }
static inline int edge_g(int j, int goe, int ge, int W) { return j <= W ? -(goe + ge * j) - goe : NEG16; }

typedef struct {
    int *H[3];   /* H on anti-diagonals d, d-1, d-2, indexed by target column r */
    int *En[2];  /* E produced at (q,r) for (q,r+1), agatha_kernel.h:27 */
    int *Fn[2];  /* F produced at (q,r) for (q+1,r), agatha_kernel.h:26 */
    uint8_t *qc, *tc;
    int cap;
} scratch_t;

static int scratch_reserve(scratch_t *s, int n)
{
    if (n <= s->cap) return 0;
    int cap = n + n / 4 + 64;
    for (int i = 0; i < 3; i++) { free(s->H[i]); s->H[i] = (int *)malloc(sizeof(int) * (size_t)cap); if (!s->H[i]) return -1; }
    for (int i = 0; i < 2; i++) {
        free(s->En[i]); s->En[i] = (int *)malloc(sizeof(int) * (size_t)cap);
        free(s->Fn[i]); s->Fn[i] = (int *)malloc(sizeof(int) * (size_t)cap);
        if (!s->En[i] || !s->Fn[i]) return -1;
    }
    free(s->qc); s->qc = (uint8_t *)malloc((size_t)cap);
    free(s->tc); s->tc = (uint8_t *)malloc((size_t)cap);
    if (!s->qc || !s->tc) return -1;
    s->cap = cap;
    return 0;
}

static void scratch_free(scratch_t *s)
{
    for (int i = 0; i < 3; i++) free(s->H[i]);
    for (int i = 0; i < 2; i++) { free(s->En[i]); free(s->Fn[i]); }
    free(s->qc); free(s->tc);
    memset(s, 0, sizeof(*s));
}

typedef struct { int max, mt, mq; } scan_state_t;

/* Termination Condition & Score Update, agatha_kernel.h:292-314 (same code at :337-356).
 * best_h/best_r are the anti-diagonal maximum with ties resolved to the largest target index,
 * which is what max() over (h<<16)+ref_idx gives (agatha_kernel.h:30,296-299).
 * Returns 1 when the Z-drop condition fires. */
static inline int scan_diag(scan_state_t *s, int d, int best_h, int best_r, int Z, int ge)
{
    if (best_h > s->max) {
        s->max = best_h; s->mt = best_r; s->mq = d - best_r;
    } else if (best_r >= s->mt && d - best_r >= s->mq) {
        int tl = best_r - s->mt, ql = (d - best_r) - s->mq;
        int l = tl > ql ? tl - ql : ql - tl;
        if (Z >= 0 && s->max - best_h > Z + l * ge) return 1;
    }
    return 0;
}

/* All cells of cell-anti-diagonal d that the reference computes: rows q < qlen (agatha_kernel.h:236),
 * columns r < 8*ceil(tlen/8) including the N padding ("phantom" columns, SURVEY A.7: CORE_COMPUTE has
 * no r < ref_len guard), band |r-q| <= W (chunk bounds :224-225 plus CORE_COMPUTE_BOUNDARY :33).
 * Recurrence: CORE_COMPUTE, agatha_kernel.h:20-30. Returns number of real (r < tlen) cells. */
static int compute_diag(scratch_t *s, int d, int qlen, int tlen, int pt, const agatha_oracle_params_t *p,
                        int *best_h, int *best_r)
{
    const int W = p->band_width, sw = p->slice_width;
    const int ge = p->gap_extend, goe = p->gap_open + p->gap_extend;
    const int tcols = g_model_phantom ? 8 * pt : tlen;
    int *Hc = s->H[d % 3], *H2 = s->H[(d + 1) % 3];           /* (d-2) % 3 == (d+1) % 3 */
    int *Enc = s->En[d & 1], *En1 = s->En[(d + 1) & 1];
    int *Fnc = s->Fn[d & 1], *Fn1 = s->Fn[(d + 1) & 1];

    int rlo = imax(0, d - (qlen - 1));
    rlo = imax(rlo, (d - W + 1) >> 1);          /* ceil((d-W)/2): 2r-d >= -W */
    int rhi = imin(d, tcols - 1);
    rhi = imin(rhi, (d + W) >> 1);              /* floor((d+W)/2): 2r-d <= W */

    int bh = -32768, br = 0;                    /* empty ring slot INT_MIN: (INT_MIN>>16, INT_MIN&65535), :152,:296-299 */
    int real = 0;
    for (int r = rlo; r <= rhi; r++) {
        const int q = d - r, k = r - q;
        const int phantom = r >= tlen;
        /* h[m]/f[m] of a padding column are never stored, and are re-initialised to MINUS_INF2 each time a
         * lane (re)loads its column block: at the first query block of every slice chunk, i.e. when
         * (q/8 + R) % sw == 0 for the last target block R = pt-1, and at q == 0 (agatha_kernel.h:206-221,272-279). */
        const int reset = phantom && (q & 7) == 0 && (q == 0 || ((q >> 3) + pt - 1) % sw == 0);
        int diag, ein, fin;
        if (q == 0 && r == 0) diag = 0;                                /* topleft[0] = 0, :146 */
        else if (q == 0) diag = edge_h(r - 1, goe, ge, W);             /* left strip .x / topleft, :138,:146 */
        else if (r == 0) diag = edge_h(q - 1, goe, ge, W);             /* top strip .x, :130 */
        else diag = H2[r - 1];
        if (reset && r - 1 >= tlen) diag = NEG16;                      /* p[m] = h[m-1] of a padding column, :213,:219-221 */

        if (r == 0) ein = edge_g(q, goe, ge, W);                       /* top strip .y, :130 */
        else ein = (k - 1 >= -W) ? En1[r - 1] : NEG16;                 /* untouched strip entry is initHD, :130 */

        if (q == 0) fin = edge_g(r, goe, ge, W);                       /* left strip .y, :138 */
        else fin = (k + 1 <= W) ? Fn1[r] : NEG16;                      /* untouched strip entry is initHD, :138 */
        if (reset) fin = NEG16;                                        /* f[m] = MINUS_INF2, :214 */

        const int a = s->qc[q];
        const int b = phantom ? N_NIBBLE : s->tc[r];                   /* host_batch.cpp:143-146 pads with 'N' */
        const int m = diag + sub_score(a, b, p);                       /* temp_score += p[m], :22-23 */
        const int h = imax(imax(m, fin), ein);                         /* :24-25 */
        Fnc[r] = imax(m - goe, fin - ge);                              /* :26 */
        Enc[r] = imax(m - goe, ein - ge);                              /* :27 */
        Hc[r] = h;
        if (h >= bh) { bh = h; br = r; }                               /* max of (h<<16)+r, r ascending, :30 */
        real += !phantom;
    }
    *best_h = bh; *best_r = br;
    return real;
}

/* The job loop of agatha_kernel for one pair, agatha_kernel.h:157-363. */
static void run_pair(scratch_t *s, const uint8_t *q, int qlen, const uint8_t *t, int tlen,
                     const agatha_oracle_params_t *p, agatha_oracle_result_t *out)
{
    memset(out, 0, sizeof(*out));
    if (qlen <= 0 || tlen <= 0) { out->stop = AGATHA_ORACLE_STOP_END; return; }

    const int W = p->band_width, sw = p->slice_width, Z = p->z_threshold, ge = p->gap_extend;
    const int pq = (qlen + 7) >> 3, pt = (tlen + 7) >> 3;             /* :120-121 */
    const int total = pq + pt - 1;                                    /* total_anti_diags, :165 */
    const int L = qlen + tlen - 1;                                    /* prev_max_score, :289 */

    for (int i = 0; i < qlen; i++) s->qc[i] = q[i] & 15;              /* pack_rc_seqs.h:24-31 */
    for (int i = 0; i < tlen; i++) s->tc[i] = t[i] & 15;

    scan_state_t st = {0, 0, 0};                                      /* :158-161 */
    int stop = AGATHA_ORACLE_STOP_END, d_stop = L;
    int64_t cells = 0;
    int d = 0, i = 0, done = 0;
    int bh, br;

    for (i = 0; i < total && !done; i += sw) {                        /* while (i < total_anti_diags), :180 ... i += sw, :330 */
        /* slice bounds, :183-186 (C integer division, truncating) */
        int ss = imax(0, i - pq + 1);
        ss = imax(ss, (i * 8 + 8 - W) / 2 / 8);
        int se = imin(pt - 1, i + sw - 1);
        se = imin(se, ((i + sw - 1) * 8 + 7 + W) / 2 / 8);
        if (ss > se) {                                                /* :189-191 -> terminated, nothing of this slice is scanned */
            stop = AGATHA_ORACLE_STOP_BANDEXIT; d_stop = imin(8 * i, L); done = 1; break;
        }
        for (; d < 8 * (i + sw); d++) {                               /* diag_idx in [i<<3, (i+sw)<<3), :293 */
            int real = compute_diag(s, d, qlen, tlen, pt, p, &bh, &br);
            if (d < L) {                                              /* :294 */
                cells += real;
                if (scan_diag(&st, d, bh, br, Z, ge)) { stop = AGATHA_ORACLE_STOP_ZDROP; d_stop = d + 1; done = 1; break; }
            }
        }
    }
    /* Job wrap-up, :334-356: taken when i >= total_anti_diags; scans 8 more ring slots starting at i*8
     * with no d < L guard. Only when i == total can those slots hold cells (blocks on block-anti-diagonal
     * total-1 reach cell anti-diagonal 8*total+6); for i > total they are already reset or alias
     * scanned slots and cannot change the result. */
    if (!done && i == total) {
        for (; d < 8 * total + 8; d++) {
            int real = compute_diag(s, d, qlen, tlen, pt, p, &bh, &br);
            if (d < L) cells += real;
            if (scan_diag(&st, d, bh, br, Z, ge)) {
                if (d < L) { stop = AGATHA_ORACLE_STOP_ZDROP; d_stop = d + 1; }
                break;
            }
        }
    }
    out->score = st.max; out->query_end = st.mq; out->target_end = st.mt;   /* :359-363 */
    out->stop = stop; out->d_stop = d_stop; out->cells = cells;
}

int agatha_oracle_align(const uint8_t *q, int32_t qlen, const uint8_t *t, int32_t tlen,
                        const agatha_oracle_params_t *p, agatha_oracle_result_t *out)
{
    scratch_t s; memset(&s, 0, sizeof(s));
    int n = imax(qlen, ((tlen + 7) & ~7)) + 16;
    if (scratch_reserve(&s, n)) { scratch_free(&s); return -1; }
    run_pair(&s, q, qlen, t, tlen, p, out);
    scratch_free(&s);
    return 0;
}

int agatha_oracle_align_batch(const uint8_t *qbuf, const uint32_t *qoff, const uint32_t *qlen,
                              const uint8_t *tbuf, const uint32_t *toff, const uint32_t *tlen,
                              int32_t n, const agatha_oracle_params_t *p,
                              agatha_oracle_result_t *out, int32_t nthreads)
{
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#else
    (void)nthreads;
#endif
    int err = 0;
#pragma omp parallel num_threads(used)
    {
        scratch_t s; memset(&s, 0, sizeof(s));
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < n; i++) {
            int need = imax((int)qlen[i], (((int)tlen[i] + 7) & ~7)) + 16;
            if (scratch_reserve(&s, need)) { err = 1; continue; }
            run_pair(&s, qbuf + qoff[i], (int)qlen[i], tbuf + toff[i], (int)tlen[i], p, &out[i]);
        }
        scratch_free(&s);
    }
    return err ? -1 : used;
}

int64_t agatha_oracle_band_cells(int32_t qlen, int32_t tlen, int32_t W)
{
    int64_t c = 0;
    for (int q = 0; q < qlen; q++) {
        int lo = imax(0, q - W), hi = imin(tlen - 1, q + W);
        if (hi >= lo) c += hi - lo + 1;
    }
    return c;
}

/* pack_rc_seqs.h:111-169 (reverse) and :171-205 (complement) */
void agatha_oracle_apply_op(uint8_t *seq, int32_t len, int32_t op)
{
    if (op & 1)
        for (int32_t i = 0, j = len - 1; i < j; i++, j--) { uint8_t c = seq[i]; seq[i] = seq[j]; seq[j] = c; }
    if (op & 2)
        for (int32_t i = 0; i < len; i++) {
            uint8_t lo = seq[i] & 15u;
            lo = lo == 1 ? 4 : lo == 4 ? 1 : lo == 3 ? 7 : lo == 7 ? 3 : lo;
            seq[i] = (uint8_t)((seq[i] & 0xF0u) | lo);
        }
}
