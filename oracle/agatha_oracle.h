/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the guided-alignment hot path of readwrite112/AGAThA:
 * banded affine-gap extension with Z-drop, score + end coordinates, no traceback.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the shipped engine (agatha_b200/) never links or calls it.
 *
 * Parity pin: the reference has no golden vectors or tests, so this restatement
 * is pinned against the reference kernel header itself compiled as single-lane host code
 * (oracle/_ref/libagatha_ref_host.so, built by oracle/Makefile from /root/reference in place) and
 * against vectors generated from that build and committed under tests/golden/.
 */
#ifndef AGATHA_ORACLE_H
#define AGATHA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same fields and meaning as the reference's gasal_subst_scores (AGAThA/src/gasal.h:165-173). */
typedef struct {
    int32_t match;       /* -m, added on equal bases            */
    int32_t mismatch;    /* -x, subtracted on unequal bases     */
    int32_t gap_open;    /* -q                                   */
    int32_t gap_extend;  /* -r                                   */
    int32_t slice_width; /* -s, block-anti-diagonals per slice   */
    int32_t z_threshold; /* -z, < 0 disables Z-drop              */
    int32_t band_width;  /* -w                                   */
} agatha_oracle_params_t;

enum {
    AGATHA_ORACLE_STOP_END = 0,      /* every anti-diagonal was scanned                     */
    AGATHA_ORACLE_STOP_ZDROP = 1,    /* Z-drop fired on a real anti-diagonal                */
    AGATHA_ORACLE_STOP_BANDEXIT = 2  /* band left the DP matrix at a slice boundary (A.6)   */
};

typedef struct {
    int32_t score;       /* gasal_res_t.aln_score          */
    int32_t query_end;   /* gasal_res_t.query_batch_end    */
    int32_t target_end;  /* gasal_res_t.target_batch_end   */
    int32_t stop;        /* AGATHA_ORACLE_STOP_*           */
    int32_t d_stop;      /* number of leading cell anti-diagonals whose cells were needed */
    int32_t reserved;
    int64_t cells;       /* in-band real cells on anti-diagonals < d_stop (GCUPS numerator) */
} agatha_oracle_result_t;

/* One pair. q/t are ASCII bases (only the low nibble is used, as in pack_rc_seqs.h:24-31). */
int agatha_oracle_align(const uint8_t *q, int32_t qlen, const uint8_t *t, int32_t tlen,
                        const agatha_oracle_params_t *p, agatha_oracle_result_t *out);

/* A batch laid out like the reference's host batch: offsets in bytes into qbuf/tbuf, lengths in bases.
 * nthreads <= 0 uses every core OpenMP offers. Returns the number of threads used. */
int agatha_oracle_align_batch(const uint8_t *qbuf, const uint32_t *qoff, const uint32_t *qlen,
                              const uint8_t *tbuf, const uint32_t *toff, const uint32_t *tlen,
                              int32_t n, const agatha_oracle_params_t *p,
                              agatha_oracle_result_t *out, int32_t nthreads);

/* In-band real cells of the whole matrix (closed form of SURVEY.md section 8d), no stopping. */
int64_t agatha_oracle_band_cells(int32_t qlen, int32_t tlen, int32_t band_width);

/* Per-sequence reverse / complement op (bit 0 = reverse, bit 1 = complement), in place on the ASCII bases.
 * Restates the intent of gasal_reversecomplement_kernel (kernels/pack_rc_seqs.h:56-212): reverse the real bases (the
 * padding is not part of the sequence) and swap A<->T, C<->G by 4-bit code (low nibbles 1<->4, 3<->7), every other
 * symbol unchanged. The high nibble of each byte is kept. See oracle/README.md for where the reference binary, as
 * compiled with -DN_CODE=0x4E, departs from this for lengths that are not multiples of 8. */
void agatha_oracle_apply_op(uint8_t *seq, int32_t len, int32_t op);

/* Model switches for experiments (tests only). Bit 0: model phantom (padding) target columns
 * exactly like the reference (default on). */
void agatha_oracle_set_model(int32_t flags);

#ifdef __cplusplus
}
#endif
#endif
