/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Host shim that lets the reference's own alignment kernel header
 * (/root/reference/AGAThA/src/kernels/agatha_kernel.h) compile and run as single-lane CPU code
 * (SURVEY.md Appendix B). oracle/Makefile pipes that header through three sed edits into a temporary
 * directory (subwarp length 8 -> 1, rejoin loop bound -> 1, extern __shared__ line dropped), extracts
 * the reference's scoring macros from gasal_kernels.h the same way, and compiles this file against
 * them. Nothing of the reference is copied into the repository; only the resulting
 * oracle/_ref/libagatha_ref_host.so is kept (git-ignored).
 *
 * This is the real reference arithmetic (int16 strips, (h<<16)+ref_idx packing, slice schedule), so
 * it pins the restatement in agatha_oracle.c inside the reference's valid domain (Appendix C).
 */
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

struct short2 { short x, y; };
static inline short2 make_short2(int x, int y) { short2 r; r.x = (short)x; r.y = (short)y; return r; }
struct uint4 { unsigned x, y, z, w; };

/* Field order of gasal_res_t, AGAThA/src/gasal.h:85-94. */
struct gasal_res_t {
    int32_t *aln_score;
    int32_t *query_batch_end;
    int32_t *target_batch_end;
    int32_t *query_batch_start;
    int32_t *target_batch_start;
    uint8_t *cigar;
    uint32_t *n_cigar_ops;
};

struct shim_dim3 { unsigned x, y, z; };
static const shim_dim3 blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1}, threadIdx = {0, 0, 0};

#define __global__ static
#define __syncwarp() do { } while (0)
#define __activemask() 1u
#define __match_any_sync(mask, v) 1u
#define __reduce_max_sync(mask, v) (v)
#define __popc(x) __builtin_popcount(x)
using std::max;
using std::min;

/* The __constant__ scoring symbols of gasal_kernels.h:29-36, one copy per host thread. */
static thread_local int32_t _cudaGapO, _cudaGapOE, _cudaGapExtend, _cudaMatchScore, _cudaMismatchScore,
    _cudaSliceWidth, _cudaZThreshold, _cudaBandWidth;

/* Dynamic shared memory of one 32-thread block: ring 32*8*(sw+1) ints + 28-int job board. */
static thread_local int32_t shared_maxHH[1 << 16];

#define N_CODE 0x4E
#define N_PENALTY 1
#include "ref_scoring_macros.h"   /* generated: MINUS_INF2, N_VALUE, DEV_GET_SUB_SCORE_GLOBAL from gasal_kernels.h */
#include "agatha_kernel_host.h"   /* generated: sed-edited agatha_kernel.h */

/* 8 ASCII bases -> one word, first base in bits 31..28, restating gasal_pack_kernel (pack_rc_seqs.h:24-36),
 * after padding to a multiple of 8 with 'N' like gasal_host_batch_fill (host_batch.cpp:143-146). */
static void pack_like_reference(const uint8_t *s, int len, std::vector<uint32_t> &out)
{
    int words = (len + 7) / 8;
    out.assign((size_t)words + 1, 0u);
    for (int w = 0; w < words; w++) {
        uint32_t v = 0;
        for (int b = 0; b < 8; b++) {
            int i = w * 8 + b;
            uint32_t c = (i < len ? s[i] : (uint8_t)N_CODE) & 15u;
            v |= c << (28 - 4 * b);
        }
        out[(size_t)w] = v;
    }
}

extern "C" int ref_host_align(const uint8_t *q, int32_t qlen, const uint8_t *t, int32_t tlen,
                              const int32_t params[7], int32_t out[3])
{
    /* params: match, mismatch, gap_open, gap_extend, slice_width, z_threshold, band_width (gasal.h:165-173) */
    _cudaMatchScore = params[0]; _cudaMismatchScore = params[1];
    _cudaGapO = params[2]; _cudaGapExtend = params[3]; _cudaGapOE = params[2] + params[3];   /* gasal_align.cu:298-302 */
    _cudaSliceWidth = params[4]; _cudaZThreshold = params[5]; _cudaBandWidth = params[6];

    std::vector<uint32_t> pq, pt;
    pack_like_reference(q, qlen, pq);
    pack_like_reference(t, tlen, pt);
    uint32_t qlens = (uint32_t)qlen, tlens = (uint32_t)tlen, qoff = 0, toff = 0;
    int32_t score = 0, qend = 0, tend = 0;
    gasal_res_t res; std::memset(&res, 0, sizeof(res));
    res.aln_score = &score; res.query_batch_end = &qend; res.target_batch_end = &tend;

    uint32_t maxlen = (uint32_t)std::max(qlen, tlen);                 /* maximum_sequence_length, test_prog.cpp:149-152 */
    if (maxlen < 8) maxlen = 8;
    /* global_buffer: 3 strips x (blockDim/8 = 4 subwarps) x maxlen + n_tasks sort slots, ctors.cpp:89 */
    std::vector<short2> gb((size_t)maxlen * 12 + 16);
    std::memset(gb.data(), 0, gb.size() * sizeof(short2));
    int npq = (qlen + 7) / 8, npt = (tlen + 7) / 8;
    gb[(size_t)maxlen * 12] = make_short2(npq + npt - 1, 0);          /* what agatha_sort writes, agatha_kernel.h:450 */
    std::memset(shared_maxHH, 0, sizeof(int32_t) * (size_t)(32 * 8 * (params[4] + 1) + 64));

    agatha_kernel(pq.data(), pt.data(), &qlens, &tlens, &qoff, &toff, &res, nullptr, nullptr, 1, maxlen, gb.data());
    out[0] = score; out[1] = qend; out[2] = tend;
    return 0;
}

extern "C" int ref_host_align_batch(const uint8_t *qbuf, const uint32_t *qoff, const uint32_t *qlen,
                                    const uint8_t *tbuf, const uint32_t *toff, const uint32_t *tlen,
                                    int32_t n, const int32_t params[7], int32_t *out3, int32_t nthreads)
{
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(used)
    for (int i = 0; i < n; i++)
        ref_host_align(qbuf + qoff[i], (int32_t)qlen[i], tbuf + toff[i], (int32_t)tlen[i], params, out3 + 3 * (size_t)i);
    return used;
}
