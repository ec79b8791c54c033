// Banded affine-gap extension with Z-drop for sm_100a: the GENERAL kernel -- every band width, every symbol, int32 state.
// The hot path of the shipped workloads is the packed kernel in extend16_kernel.cuh; this kernel aligns the pairs that one
// marks (redo pass) and everything the packed kernel does not take (W not 7 mod 8, W < 135, scoring outside its table, ...).
//
// Replaces the reference's agatha_kernel (AGAThA/src/kernels/agatha_kernel.h:49-431) + agatha_sort (:434-458);
// same results (score, query end, target end), different algorithm layout. Parity spec: SURVEY.md Appendix A,
// restated and pinned in oracle/agatha_oracle.c.
//
// Formulation: cells are addressed by anti-diagonal d = q + r and diagonal offset k = r - q.
// In (d,k) space the band |r-q| <= W is a FIXED stripe, so a lane owns a fixed, contiguous k-range for the whole
// alignment and its H/E/F state never leaves registers. One group of NW warps handles one alignment:
//   global lane gl = warp_in_group*32 + lane owns k = -W + 2*(C*gl + j) + u,  j in [0,C), u in {0,1}
// and step d updates the C cells of parity u = (d+W)&1 of every lane (they are mutually independent). A cell reads
//   E from (d-1,k-1), F from (d-1,k+1), H from (d-2,k); only the first/last cell of a lane needs a neighbour lane
// (one warp shuffle per step). The per-anti-diagonal maximum needed by the Z-drop test is a lane-local max of
// (H*32+j) keys, one warp REDUX per step and a warp-uniform scalar update: the scan is exact per anti-diagonal,
// so a terminated alignment stops at once and the warp pulls the next job (work redistribution on termination).
// Per cell of the 32-bit loops, ALU pipe: IDP.4A (score byte select + add), VIMNMX3 (H), 2x VIADDMNMX (E, F), 1/2 VIMNMX3
// (tracking); FMA pipe: IMAD for t = M - goe and for the tracking key. One-warp shapes with W = 7 (mod 8) may run their
// prologue and steady state on 16-bit packed state (run_fast16, AGATHA_S16=1/2) -- round 1's fast path, superseded by
// extend16_kernel.cuh and kept for A/B measurements.
#pragma once

#include <cstdint>
#include <climits>
#include <type_traits>
#include <cuda_runtime.h>

#include "agatha_b200.h"

#ifndef AGATHA_TAIL16
#define AGATHA_TAIL16 0
#endif

namespace agatha {

constexpr int NEG16 = -16384;        // the reference's MINUS_INF2 (gasal_kernels.h:39); exact value matters for parity
// Packed kernel (extend16_kernel.cuh): its range monitor looks at the live values every RANGE16_PERIOD anti-diagonals and its
// margins cover RANGE16_PAIRS pairs of anti-diagonals (the period plus two: a multi-warp group evaluates the check two steps
// late). It used to run every 32: a source-level profile of the 1-8 kb workload put 8 % of the kernel's time there (the check
// is cold code, every visit costs instruction fetch on top of its 250 instructions). Measured, C1 / C2 / C4 GCUPS on one box:
// 64 -> 2,914 / 4,322 / 2,741; 128 -> 3,083 / 4,392 / 2,932; 256 -> 3,186 / 4,489 / 2,942; on another box 256 -> 3,220 / 4,509 /
// 2,966; 512 -> 3,336 / 4,538 / 2,974; 1024 -> 3,366 / 4,549 / 3,121. The price is the margin: parameter sets whose scores
// are so large that 257 * (match + mismatch) no longer fits the window go to the general kernel; 512 it is.
#ifndef AGATHA_RANGE16_PERIOD
#define AGATHA_RANGE16_PERIOD 512
#endif
constexpr int RANGE16_PERIOD = AGATHA_RANGE16_PERIOD;
constexpr int RANGE16_PAIRS = RANGE16_PERIOD / 2 + 1;
constexpr int NEGBIG = -(1 << 25);   // "never a real score"; NEGBIG*32 still fits int32 (tracking keys)
constexpr unsigned FULL = 0xffffffffu;

// canonical 4-bit codes written by the pack kernel (pack_kernel.cuh)
constexpr int QCODE_N = 4;           // query  'N'
constexpr int TCODE_N = 13;          // target 'N'   (4 ^ 13 = 9, (0..3) ^ 13 = 12..15: all hit sign-replicated LUT entries)
constexpr int QCODE_Y = 13;          // the one non-ACGTN symbol whose query code collides with TCODE_N ...
constexpr int TCODE_Y = 4;           // ... and its target code

struct KernelParams {
    int match, mismatch, goe, ge, sw, Z, W;
    unsigned tab_lo, tab_hi;         // PRMT lookup table: byte x -> score for code XOR x (fast alphabet)
    int LW, JW;                      // lane / cell index of k = +W  (global cell index g = W, u = 0)
    int one, k32;                    // the constants 1 and 32, opaque to ptxas: keep a*1+c and h*32+j on IMAD (FMA pipe)
    int m16;                         // 0xffff, opaque too: (bit pair) * 0xffff widens a valid-cell bit to a 16-bit mask on the FMA pipe
    int k65536;                      // 65536, opaque: x * 65536 is a shift by 16 on the FMA pipe
    // packed kernel (extend16_kernel.cuh): table of the biased scores s + mismatch, initial offset of the stored values
    unsigned tabb_lo, tabb_hi;
    int bias16;
    int p16_ok;                      // the packed kernel's preconditions hold for these parameters
    int s16;                         // bit 0: steady state, bit 1: prologue, bit 2: tail may run on 16-bit packed state (engine.cu)
    int force_generic;               // match/mismatch do not fit the byte table: score every pair with compare/select
};

struct JobArrays {
    const uint32_t* qpk;             // query  bases, 8 per word, first base in bits 31..28, codes cq()
    const uint32_t* tpk;             // target bases, 8 per word, first base in bits  3..0,  codes ct()
    const uint32_t* qoff_w;          // word offset of each query in qpk
    const uint32_t* toff_w;
    const uint32_t* qlen;            // in bases
    const uint32_t* tlen;
    const uint32_t* order;           // job -> pair index, longest first (host-side bucketing), may be null
    int32_t* score;
    int32_t* qend;
    int32_t* tend;
    int32_t* stop;                   // AGATHA_STOP_*
    int32_t* dstop;
    unsigned* counter;               // work queue head
    int n;
    int redo;                        // 0: align every pair; 1: only the pairs marked REDO_MARK by the packed kernel
};

// Integer multiply-add that ptxas cannot strength-reduce to IADD3/LEA (b comes from the constant bank at run time).
// ncu (profiles/extend_kernel_r01_ncu_full.csv): ALU pipe 81 % busy, FMA pipe 14 % -- every op moved here is free.
// (AGATHA_HOST_EMU: tests/emu compiles this header for the CPU, where inline PTX has to be spelled in C.)
__device__ __forceinline__ int imad(int a, int b, int c)
{
#ifdef AGATHA_HOST_EMU
    return (int)((unsigned)a * (unsigned)b + (unsigned)c);
#else
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel)
{
#ifdef AGATHA_HOST_EMU
    return __byte_perm(a, b, sel);
#else
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// One anti-diagonal step for the C cells of parity U owned by this lane.
//   H   : H of the parity-U cells (value from anti-diagonal d-2 on entry, d on exit)
//   E,F : on entry outputs of the previous step (other parity); on exit outputs of this step
//   Qw/Rw: nibble j of the windows holds the query/target code of cell j
// Returns the lane's best tracking key max_j(H_j*32 + j) (ties -> largest j == largest target index); TAIL: only over the
// cells whose bit is set in vmask (the cells inside the matrix).
// Recurrence: CORE_COMPUTE, agatha_kernel.h:20-30 (gap opens from M = diag + s, not from H).
// ---------------------------------------------------------------------------------------------------------------
template <int C, int U, bool TAIL, bool GENERIC>
__device__ __forceinline__ int step_cells(int (&H)[C], int (&E)[C], int (&F)[C],
                                          const uint32_t (&Qw)[(C + 7) / 8], const uint32_t (&Rw)[(C + 7) / 8],
                                          int edge_in, const KernelParams& p, unsigned vmask)
{
    constexpr int NWORD = (C + 7) / 8;
    unsigned sc[2 * NWORD];
    if (!GENERIC) {
#pragma unroll
        for (int w = 0; w < NWORD; w++) {
            unsigned x = Qw[w] ^ Rw[w];
            sc[2 * w] = prmt(p.tab_lo, p.tab_hi, x);
            sc[2 * w + 1] = prmt(p.tab_lo, p.tab_hi, x >> 16);
        }
    }
    const int mge = -p.ge, mgoe = -p.goe;
    int best = INT_MIN;
    int pend = INT_MIN;

#pragma unroll
    for (int jj = 0; jj < C; jj++) {
        // U == 0 reads E[j-1] (old) -> walk j downwards; U == 1 reads F[j+1] (old) -> walk j upwards
        const int j = (U == 0) ? (C - 1 - jj) : jj;
        int ein, fin;
        if (U == 0) { ein = (j == 0) ? edge_in : E[j - 1]; fin = F[j]; }
        else        { ein = E[j]; fin = (j == C - 1) ? edge_in : F[j + 1]; }
        int m;
        if (!GENERIC) {
            m = __dp4a((int)sc[j >> 2], 1 << (8 * (j & 3)), H[j]);                 // H(d-2,k) + s
        } else {
            const unsigned a = (Qw[j >> 3] >> (4 * (j & 7))) & 15u, b = (Rw[j >> 3] >> (4 * (j & 7))) & 15u;
            const bool isn = (a == (unsigned)QCODE_N) | (b == (unsigned)TCODE_N);
            const bool eq = (a == b) | ((a == (unsigned)QCODE_Y) & (b == (unsigned)TCODE_Y));
            const int s = isn ? -1 : (eq ? p.match : -p.mismatch);                // DEV_GET_SUB_SCORE_GLOBAL, N_PENALTY=1
            m = H[j] + s;
        }
        const int h = __vimax3_s32(m, ein, fin);
        const int t = imad(m, p.one, mgoe);
        E[j] = __viaddmax_s32(ein, mge, t);
        F[j] = __viaddmax_s32(fin, mge, t);
        H[j] = h;
        int key = imad(h, p.k32, j);
        if (TAIL) key = ((vmask >> j) & 1u) ? key : INT_MIN;                       // one bit test per cell instead of two compares
        if (jj & 1) best = __vimax3_s32(best, pend, key); else pend = key;
    }
    if (C & 1) best = max(best, pend);
    return best;
}

// ---------------------------------------------------------------------------------------------------------------
// The same step on 16-bit packed state (steady state only): register jj of each array holds cell jj in its low half
// and cell jj + C/2 in its high half, values relative to a per-alignment base, so one VIADDMNMX.S16x2 / VIMNMX3.S16x2
// updates two cells (measured: same issue rate as the 32-bit forms, profiles/int_peak_r01.json).
//   M is clamped at FLOOR16 by the same VIADDMNMX that computes it, which bounds every value from below (t >= FLOOR - goe,
//   E,F >= t) and keeps dead (out-of-band) cells near the floor; run_pair's range monitor guarantees that no LIVE value
//   ever gets near the clamp or the top of the range, so inside this loop the arithmetic is identical to the 32-bit one
//   (DESIGN.md section 2).
// Returns the packed running maximum of H (no index: the argmax is recovered lazily from a snapshot).
// ---------------------------------------------------------------------------------------------------------------
// Representation: stored half = (value - base) + BIAS16 as an UNSIGNED 16-bit number, FLOOR16 <= value - base <= TOP16.
// Unsigned, because then every half is >= FLOORU16 > goe and t = M - goe can be ONE 32-bit IMAD on the (idle) FMA pipe:
// no half can borrow from its neighbour. VIADDMNMX.U16x2 adds modulo 2^16 (checked on a B200), so the sign-extended
// negative scores and -ge work as two's complement addends.
constexpr int FLOOR16 = -30000;
constexpr int BIAS16 = 34000;
constexpr int TOP16 = 65535 - BIAS16;                    // 31535
constexpr unsigned FLOORU16 = (unsigned)(FLOOR16 + BIAS16);   // 4000

__device__ __forceinline__ unsigned pack16(int lo, int hi) { return ((unsigned)(lo + BIAS16) & 0xffffu) | ((unsigned)(hi + BIAS16) << 16); }
__device__ __forceinline__ unsigned pack16raw(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }   // two's complement halves
__device__ __forceinline__ int lo16(unsigned x) { return (int)(x & 0xffffu) - BIAS16; }
__device__ __forceinline__ int hi16(unsigned x) { return (int)(x >> 16) - BIAS16; }

// TAILM: `vm2` holds one bit per cell of this anti-diagonal that lies inside the matrix (cells 0..P-1 in bits 0.., cells
// P..C-1 in bits 16..); the others still take part in the recurrence (nothing inside the matrix ever reads them) but are
// kept out of the maximum: their halves are ANDed to 0, the smallest packed value.
template <int C, int U, bool TAILM = false>
__device__ __forceinline__ unsigned step_cells16(unsigned (&H)[C / 2], unsigned (&E)[C / 2], unsigned (&F)[C / 2],
                                                 const uint32_t (&Qw)[(C + 7) / 8], const uint32_t (&Rw)[(C + 7) / 8],
                                                 unsigned edge_in, const KernelParams& p, unsigned mge2, int mgoe32, unsigned floor2,
                                                 unsigned vm2 = 0u)
{
    constexpr int P = C / 2, NWORD = (C + 7) / 8;
    unsigned sc[2 * NWORD];
#pragma unroll
    for (int w = 0; w < NWORD; w++) {
        unsigned x = Qw[w] ^ Rw[w];
        sc[2 * w] = prmt(p.tab_lo, p.tab_hi, x);
        sc[2 * w + 1] = prmt(p.tab_lo, p.tab_hi, x >> 16);
    }
    unsigned best = 0u, pend = 0u;
#pragma unroll
    for (int t_ = 0; t_ < P; t_++) {
        const int jj = (U == 0) ? (P - 1 - t_) : t_;          // same read-before-write order as step_cells
        unsigned ein, fin;
        if (U == 0) { ein = (jj == 0) ? edge_in : E[jj - 1]; fin = F[jj]; }
        else        { ein = E[jj]; fin = (jj == P - 1) ? edge_in : F[jj + 1]; }
        // sign-extended score pair: byte (a&3) of sc[a>>2] for cell a = jj, byte (b&3) of sc[b>>2] for cell b = jj + P
        const int a = jj, b = jj + P;
        const unsigned sel = (unsigned)(a & 3) | ((unsigned)((a & 3) | 8) << 4) | ((unsigned)(4 + (b & 3)) << 8) | ((unsigned)((4 + (b & 3)) | 8) << 12);
        const unsigned s2 = prmt(sc[a >> 2], sc[b >> 2], sel);
        const unsigned m = __viaddmax_u16x2(H[jj], s2, floor2);           // max(H(d-2,k) + s, FLOOR)
        const unsigned h = __vimax3_u16x2(m, ein, fin);
        const unsigned t = (unsigned)imad((int)m, p.one, mgoe32);          // M - goe in both halves: no borrow, m >= FLOORU16 > goe
        E[jj] = __viaddmax_u16x2(ein, mge2, t);
        F[jj] = __viaddmax_u16x2(fin, mge2, t);
        H[jj] = h;
        unsigned hm = h;
        if (TAILM) hm = h & (unsigned)imad((int)((vm2 >> jj) & 0x00010001u), p.m16, 0);
        if (t_ & 1) best = __vimax3_u16x2(best, pend, hm); else pend = hm;
    }
    if (P & 1) best = __vimax3_u16x2(best, pend, pend);
    return best;
}

// Write `val` to arr[idx] with a run-time idx while keeping arr in registers: a compare + select per element (every index
// is static). idx < 0 writes nothing, so a lane that does not own the target cell passes -1 and no branch is needed.
// (A switch looks cheaper but ptxas materialises a copy of the whole array per case: ~120 moves per poke, measured.)
template <int C>
__device__ __forceinline__ void poke(int (&arr)[C], int idx, int val)
{
#pragma unroll
    for (int j = 0; j < C; j++) arr[j] = (j == idx) ? val : arr[j];
}
template <int C>
__device__ __forceinline__ void poke2(int (&a)[C], int va, int (&b)[C], int vb, int idx)
{
#pragma unroll
    for (int j = 0; j < C; j++) { const bool hit = (j == idx); a[j] = hit ? va : a[j]; b[j] = hit ? vb : b[j]; }
}

// ---------------------------------------------------------------------------------------------------------------
// Everything about one alignment that is uniform across the warp.
// ---------------------------------------------------------------------------------------------------------------
struct Pair {
    const uint32_t* q;   // packed query words of this pair
    const uint32_t* t;
    int qlen, tlen, qwords, twords, pq, pt, tcols, total, L;
};

__device__ __forceinline__ uint32_t load_qword(const Pair& pr, int wi)
{
    return (wi >= 0 && wi < pr.qwords) ? __ldg(pr.q + wi) : 0u;
}
__device__ __forceinline__ uint32_t load_tword(const Pair& pr, int wi)
{
    return (wi >= 0 && wi < pr.twords) ? __ldg(pr.t + wi) : 0xddddddddu;   // TCODE_N everywhere
}
__device__ __forceinline__ unsigned qbase(const Pair& pr, int i)   // code of query[i], 0 outside
{
    if (i < 0 || i >= pr.qwords * 8) return 0u;
    return (__ldg(pr.q + (i >> 3)) >> (28 - 4 * (i & 7))) & 15u;
}
__device__ __forceinline__ unsigned tbase(const Pair& pr, int i)
{
    if (i < 0 || i >= pr.twords * 8) return (unsigned)TCODE_N;
    return (__ldg(pr.t + (i >> 3)) >> (4 * (i & 7))) & 15u;
}

// true when the pair holds a symbol outside {A,C,G,T,N}: the PRMT lookup table cannot score those
__device__ __forceinline__ bool has_rare_symbols(const Pair& pr, int lane)
{
    bool rare = false;
    for (int i = lane; i < pr.qwords; i += 32) {
        const uint32_t w = __ldg(pr.q + i);
        // nibble > 4  <=>  bit3 | (bit2 & (bit1|bit0))
        const uint32_t b3 = w & 0x88888888u, b2 = w & 0x44444444u, lo = ((w << 1) | (w << 2)) & 0x44444444u;
        rare |= (b3 | (b2 & lo)) != 0u;
    }
    for (int i = lane; i < pr.twords; i += 32) {
        const uint32_t w = __ldg(pr.t + i);
        // nibble in {0,1,2,3} or == 13 is fine
        const uint32_t hi = (w | (w >> 1)) & 0x44444444u;                 // bit2 set <=> nibble >= 4 (bit3|bit2)
        const uint32_t x = w ^ 0xddddddddu;                               // nibble == 0 <=> code 13
        const uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;   // 1 <=> nibble != 13
        rare |= ((hi >> 2) & nz) != 0u;
    }
    return __any_sync(FULL, rare);
}

// thr: an anti-diagonal whose maximum is in [thr, max] can neither raise the maximum nor fire Z-drop (l*ge >= 0)
struct ScanState { int max, mt, mq, thr; };

__device__ __forceinline__ int scan_threshold(int mx, const KernelParams& p)
{
    if (p.ge < 0) return INT_MAX;            // no shortcut: the l*ge term could lower the bar
    if (p.Z < 0) return INT_MIN;             // Z-drop disabled
    return mx - p.Z;
}

// Shared memory of one multi-warp group (NW > 1): lane-edge hand-over between neighbouring warps and the per-warp
// anti-diagonal maxima. One __syncthreads per anti-diagonal; every slot is written before and read after it, and is
// not written again before the next barrier, so no double buffering is needed for the edges; the scan slots
// alternate by anti-diagonal parity.
template <int NW>
struct GroupShared {
    int edgeE[NW];          // E[C-1] of lane 31 of each warp (published after steps of parity 1)
    int edgeF[NW];          // F[0]   of lane 0  of each warp (published after steps of parity 0)
    int scan_h[2][NW];      // per-warp maximum H of the anti-diagonal
    int scan_g[2][NW];      // global cell index (C*gl + j) of that maximum, ties -> largest
    unsigned job;
};

// Termination Condition & Score Update for one anti-diagonal (agatha_kernel.h:292-314), uniform across the group.
// (hmax, g) is the anti-diagonal maximum and the global cell index of its right-most occurrence. Returns true when
// Z-drop fires.
__device__ __forceinline__ bool scan_update(ScanState& st, int hmax, int g, int d, int u, const KernelParams& p)
{
    int r;
    if (hmax < -32768) {                 // empty ring slot reads as INT_MIN -> (h,r) = (-32768, 0), :152,:296-299
        hmax = -32768; r = 0;
    } else {
        const int k = -p.W + 2 * g + u;
        r = (d + k) >> 1;
    }
    if (hmax > st.max) { st.max = hmax; st.mt = r; st.mq = d - r; st.thr = scan_threshold(hmax, p); return false; }
    if (r >= st.mt && d - r >= st.mq) {
        const int tl = r - st.mt, ql = (d - r) - st.mq;
        const int l = tl > ql ? tl - ql : ql - tl;
        if (p.Z >= 0 && st.max - hmax > p.Z + l * p.ge) return true;
    }
    return false;
}

// Single-warp group: `best` is this lane's key from step_cells (H*32 + j).
template <int C>
__device__ __forceinline__ bool scan_diag(ScanState& st, int best, int d, int u, int lane, const KernelParams& p)
{
    const int hmax = __reduce_max_sync(FULL, best) >> 5;          // H dominates the key, so this is the maximum H
    if (hmax <= st.max && hmax >= st.thr) return false;           // common case: nothing can happen on this anti-diagonal
    int g = 0;
    if (hmax >= -32768) {
        const unsigned who = __ballot_sync(FULL, (best >> 5) == hmax);
        const int src = 31 - __clz((int)who);                                        // ties -> largest target index
        g = C * src + __shfl_sync(FULL, best & 31, src);
    }
    return scan_update(st, hmax, g, d, u, p);
}

// Multi-warp group: publish this warp's maximum, meet the other warps, combine. Contains the group's only barrier of
// the anti-diagonal, so it must be called by every warp on every step (scan == false just skips the update).
template <int C, int NW>
__device__ __forceinline__ bool scan_diag_group(ScanState& st, int best, int d, int u, int lane, int warp, bool scan,
                                                GroupShared<NW>* sm, const KernelParams& p)
{
    const int v = best >> 5;
    const int hw = __reduce_max_sync(FULL, v);
    const unsigned who = __ballot_sync(FULL, v == hw);
    const int src = 31 - __clz((int)who);
    const int gw = C * (32 * warp + src) + __shfl_sync(FULL, best & 31, src);
    if (lane == 0) { sm->scan_h[d & 1][warp] = hw; sm->scan_g[d & 1][warp] = gw; }
    __syncthreads();
    if (!scan) return false;
    const int hv = lane < NW ? sm->scan_h[d & 1][lane] : INT_MIN;
    const int hmax = __reduce_max_sync(FULL, hv);
    if (hmax <= st.max && hmax >= st.thr) return false;
    const unsigned whow = __ballot_sync(FULL, lane < NW && hv == hmax);
    const int srcw = 31 - __clz((int)whow);                                          // ties -> largest target index
    const int g = sm->scan_g[d & 1][srcw];
    return scan_update(st, hmax, g, d, u, p);
}

// ---------------------------------------------------------------------------------------------------------------
// One alignment, one group of NW warps (NW == 1: a single warp, no shared memory, no barrier).
// WODD = (W & 1): fixes which parity class the even anti-diagonals use.
// JWS >= 0: the cell index of k = +W inside its lane is known at compile time (W % C == JWS); -1: run-time.
//
// Two loop bodies only, so that the steady state stays small and branch-free:
//   FAST  anti-diagonals W < d < d_tail: every lane's cells are inside the matrix, nothing to inject or mask
//   SLOW  everything else: matrix-edge injection (d < W), far-edge masking, padding-column patch, wrap-up
// ---------------------------------------------------------------------------------------------------------------
template <int C, int NW, bool WODD, int JWS, bool GENERIC>
__device__ __forceinline__ void run_pair(const Pair& pr, const KernelParams& p, int lane, int warp, GroupShared<NW>* sm,
                                         int& out_score, int& out_qend, int& out_tend, int& out_stop, int& out_dstop)
{
    const int gl = 32 * warp + lane;                 // lane index inside the group: owns cells g in [C*gl, C*gl + C)
    static_assert((C % 8 == 0 || C == 2 || C == 4) && C <= 32, "cells per lane: 2, 4 or a multiple of 8");
    constexpr int NWORD = (C + 7) / 8;
    using M0 = std::integral_constant<int, 0>;
    using M1 = std::integral_constant<int, 1>;
    using M2 = std::integral_constant<int, 2>;
    using M3 = std::integral_constant<int, 3>;
    using M4 = std::integral_constant<int, 4>;
    using S0 = std::integral_constant<int, 0>;
    using S1 = std::integral_constant<int, 1>;
    using S2 = std::integral_constant<int, 2>;
    using S3 = std::integral_constant<int, 3>;
    using UA = std::integral_constant<int, WODD ? 1 : 0>;   // parity class of even anti-diagonals
    using UB = std::integral_constant<int, WODD ? 0 : 1>;   // ... of odd ones
    const int W = p.W;

    int H0[C], H1[C], E[C], F[C];
#pragma unroll
    for (int j = 0; j < C; j++) { H0[j] = NEGBIG; H1[j] = NEGBIG; E[j] = NEGBIG; F[j] = NEGBIG; }

    // --- sequence windows at d = 0: nibble j <-> query[qtop - j], target[rbot + j] ---------------------------
    int qtop = (W >> 1) - C * gl;                   // ((d+W)>>1) - C*gl at d = 0
    int rbot = ((1 - W) >> 1) + C * gl;             // ((d-W+1)>>1) + C*gl at d = 0
    uint32_t Qw[NWORD], Rw[NWORD];
#pragma unroll
    for (int w = 0; w < NWORD; w++) { Qw[w] = 0u; Rw[w] = 0u; }
#pragma unroll 1
    for (int j = 0; j < C; j++) {
        const unsigned qb = qbase(pr, qtop - j) << (4 * (j & 7)), tb = tbase(pr, rbot + j) << (4 * (j & 7));
#pragma unroll
        for (int w = 0; w < NWORD; w++) if (w == (j >> 3)) { Qw[w] |= qb; Rw[w] |= tb; }
    }
    // feeds: next query base in the top nibble of qfeed, next target base in the bottom nibble of rfeed
    uint32_t qfeed, rfeed;
    {
        const int nq = qtop + 1, nb = rbot + C;
        qfeed = load_qword(pr, nq >> 3) << (4 * (nq & 7));
        rfeed = load_tword(pr, nb >> 3) >> (4 * (nb & 7));
    }

    // --- boundary: H(-1,-1) = 0, then the virtual cells of "anti-diagonal -1" (agatha_kernel.h:126-148) -------
    // virtual cells of anti-diagonal d: (q=-1, r=d+1) at k = d+2 and (q=d+1, r=-1) at k = -(d+2);
    // values H(-1,j) = H(j,-1) = -(goe + ge*j), F(0,j) = E(j,0) = that - goe, for j <= W.  (cell (g,u): k = -W + 2g + u)
    auto inject = [&](int d, auto u_tag) {
        constexpr int U = decltype(u_tag)::value;
        const int jv = d + 1;
        if (jv > W) return;
        const int hv = -(p.goe + p.ge * jv), gv = hv - p.goe;
        {   // top: consumer (0, d+1) reads F; (0, d+2) reads H as its diagonal
            const int g = (d + 2 + W) >> 1;
            const int j = g - C * gl;                                   // in [0,C) only in the owning lane
            const int jj = (j >= 0 && j < C) ? j : -1;
            if ((d + 2) <= W) { if (U == 0) poke2<C>(H0, hv, F, gv, jj); else poke2<C>(H1, hv, F, gv, jj); }
            else poke<C>(F, jj, gv);
        }
        if (W - d - 2 >= 0) {   // left: consumer (d+1, 0) reads E; (d+2, 0) reads H as its diagonal
            const int g = (W - d - 2) >> 1;
            const int j = g - C * gl;
            const int jj = (j >= 0 && j < C) ? j : -1;
            if (U == 0) poke2<C>(H0, hv, E, gv, jj); else poke2<C>(H1, hv, E, gv, jj);
        }
    };
    {
        const int g = W >> 1;                       // k = 0: 2g + u = W, u = W & 1
        const int j = g - C * gl;
        const int jj = (j >= 0 && j < C) ? j : -1;
        if (WODD) poke<C>(H1, jj, 0); else poke<C>(H0, jj, 0);
        inject(-1, UB{});                           // u(-1) = (W-1) & 1
    }

    // The same injection for a block of 8 prologue anti-diagonals d0 .. d0+7 with d0 % 8 == 0, when W % 8 == 7 (JWS >= 0)
    // and C % 4 == 0. The two virtual cells move by one cell every second anti-diagonal, in opposite directions, and with
    // this alignment their position INSIDE a group of four cells is a compile-time function of the step's place in the
    // block (S = (d - d0) / 2 and its parity); only the group is a run-time value. A poke is then C/4 selects per array
    // instead of C. blk_qt / blk_ql: index of the block's top / left group relative to this lane's first group (outside
    // [0, C/4) = another lane's); blk_last: the block that ends at d = W.
    //   top : g = (d+2+W)>>1 = G0 + S (even d), G0 + S + 1 (odd d),  G0 = (d0+W+1)/2 = 0 (mod 4)
    //   left: g = (W-d-2)>>1 = GL0 - S (both parities),              GL0 = (W-3-d0)/2 = 2 (mod 4)
    constexpr bool STATIC_PRO = JWS >= 0 && WODD && C % 4 == 0 && !GENERIC;
    constexpr int NG = C / 4;
    int blk_qt = 0, blk_ql = 0;
    bool blk_last = false;
    auto inject_static = [&](int d, auto u_tag, auto s_tag) {
        constexpr int U = decltype(u_tag)::value, S = decltype(s_tag)::value;
        constexpr bool EVEN = (U == UA::value);                            // the step on the even anti-diagonal of the pair
        constexpr int PT = EVEN ? S : ((S + 1) & 3), PL = (2 - S) & 3;     // positions inside the group
        const int hv = -(p.goe + p.ge * (d + 1)), gv = hv - p.goe;
        const int qt = (!EVEN && S == 3) ? blk_qt + 1 : blk_qt;           // the odd step of S == 3 is already in the next group
        const int ql = (S == 3) ? blk_ql - 1 : blk_ql;                    // ... and the left cell in the previous one
        const bool h_top = !(EVEN && S == 3 && blk_last);                 // d + 2 > W: (0, d+2) does not exist, only F is needed
#pragma unroll
        for (int Q = 0; Q < NG; Q++) {
            const bool ht = (qt == Q), hl = (ql == Q);
            F[4 * Q + PT] = ht ? gv : F[4 * Q + PT];
            E[4 * Q + PL] = hl ? gv : E[4 * Q + PL];
            if (U == 0) {
                H0[4 * Q + PT] = (ht && h_top) ? hv : H0[4 * Q + PT];
                H0[4 * Q + PL] = hl ? hv : H0[4 * Q + PL];
            } else {
                H1[4 * Q + PT] = (ht && h_top) ? hv : H1[4 * Q + PT];
                H1[4 * Q + PL] = hl ? hv : H1[4 * Q + PL];
            }
        }
    };

    // The step's place in its block selects one of four small bodies through a warp-uniform switch: the loop stays two
    // anti-diagonals long (small enough for the instruction cache) and still pokes compile-time positions.
    int blk_S = 0;
    auto inject_static_sw = [&](int d, auto u_tag) {
        switch (blk_S) {
        case 0: inject_static(d, u_tag, S0{}); break;
        case 1: inject_static(d, u_tag, S1{}); break;
        case 2: inject_static(d, u_tag, S2{}); break;
        default: inject_static(d, u_tag, S3{}); break;
        }
    };

    ScanState st = {0, 0, 0, scan_threshold(0, p)};   // agatha_kernel.h:158-161
    int stop = AGATHA_STOP_END, d_stop = pr.L;
    const bool has_phantom = pr.tcols > pr.tlen;
    const bool edge_lane = (gl == p.LW);
    const int jw_dyn = edge_lane ? p.JW : -1;

    // phantom (padding) target columns: their F and diagonal inputs restart from MINUS_INF2 at the first row of every
    // slice chunk of the last target block (agatha_kernel.h:206-221 reload, :272-279 never stored); see oracle.
    auto phantom_patch = [&](int d, auto u_tag) {
        constexpr int U = decltype(u_tag)::value;
        const int qc = (d - pr.tlen) & ~7;           // the only multiple of 8 in (d - tcols, d - tlen]
        if (d - pr.tlen < 0 || d - qc >= pr.tcols || qc >= pr.qlen) return;
        if (!(qc == 0 || ((qc >> 3) + pr.pt - 1) % p.sw == 0)) return;
        const int r = d - qc, k = r - qc;
        if (k > W || k < -W) return;
        const int g = (k + W - U) >> 1;              // cell (g,U) itself
        const int gf = (U == 0) ? g : g + 1;         // its F input: U==0 reads F[j], U==1 reads F[j+1] / next lane's F[0]
        const int jf = gf - C * gl;
        poke<C>(F, (jf >= 0 && jf < C) ? jf : -1, NEG16);
        if (r - 1 >= pr.tlen) {
            const int j = g - C * gl;
            const int jj = (j >= 0 && j < C) ? j : -1;
            if (U == 0) poke<C>(H0, jj, NEG16); else poke<C>(H1, jj, NEG16);
        }
    };

    auto shift_query = [&]() {                       // qtop -> qtop + 1
        const int nq = qtop + 1;
        if ((nq & 7) == 0) qfeed = load_qword(pr, nq >> 3);
#pragma unroll
        for (int w = NWORD - 1; w > 0; w--) Qw[w] = __funnelshift_l(Qw[w - 1], Qw[w], 4);
        Qw[0] = __funnelshift_l(qfeed, Qw[0], 4);
        qfeed <<= 4;
        qtop = nq;
    };
    auto shift_ref = [&]() {                         // rbot -> rbot + 1
        const int nb = rbot + C;
        if ((nb & 7) == 0) rfeed = load_tword(pr, nb >> 3);
#pragma unroll
        for (int w = 0; w < NWORD - 1; w++) Rw[w] = __funnelshift_r(Rw[w], Rw[w + 1], 4);
        if (C % 8 == 0) Rw[NWORD - 1] = __funnelshift_r(Rw[NWORD - 1], rfeed, 4);
        else Rw[NWORD - 1] = (Rw[NWORD - 1] >> 4) | ((rfeed & 15u) << (4 * ((C - 1) & 7)));   // nibbles >= C stay zero
        rfeed >>= 4;
        rbot++;
    };

    // one anti-diagonal; returns true when Z-drop fires on it
    // MODE 0 FAST: every in-band cell is inside the matrix. MODE 1 PRO: near edges only -- inject the boundary cells; what
    // lies beyond them is dead (NEGBIG), so the maximum needs no mask. MODE 2 TAIL: far edges only -- mask the maximum,
    // patch the padding columns. MODE 3: both (pairs shorter than the band).
    // MODE 4: PRO with inject_static (blk_* describe the aligned block of 8 the step belongs to).
    auto do_step = [&](int d, bool scan, auto u_tag, auto mode_tag) -> bool {
        constexpr int U = decltype(u_tag)::value;
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr bool FAST = MODE == 0, MASK = MODE == 2 || MODE == 3, INJECT = MODE == 1 || MODE == 3;
        using UN = std::integral_constant<int, 1 - U>;
        unsigned vmask = 0xffffffffu;                // cells of this lane inside the matrix, bit j <-> cell j
        if (MASK) {
            const int klo = max(-W, max(-d, d - 2 * (pr.qlen - 1)));
            const int khi = min(W, min(d, 2 * (pr.tcols - 1) - d));
            const int k0 = -W + 2 * C * gl + U;
            const int jlo = max((klo - k0 + 1) >> 1, 0);               // ceil((klo-k0)/2)
            const int jhi = min((khi - k0) >> 1, C - 1);
            vmask = (jhi >= jlo) ? ((0xffffffffu >> (31 - jhi)) & (0xffffffffu << jlo)) : 0u;
        }
        int best;
        if (U == 0) {
            int ein = __shfl_up_sync(FULL, E[C - 1], 1);
            if (lane == 0) {
                // k = -W: left of it is outside the band (MINUS_INF2), except on the matrix edge where E(W,0) is a boundary
                // value (agatha_kernel.h:130); before the band edge enters the matrix the cell is not real: keep it dead
                if (NW == 1 || warp == 0) ein = (FAST || d > W) ? NEG16 : ((d == W) ? (-(p.goe + p.ge * W) - p.goe) : NEGBIG);
                else ein = sm->edgeE[warp - 1];
            }
            best = step_cells<C, 0, MASK, GENERIC>(H0, E, F, Qw, Rw, ein, p, vmask);
            if (JWS >= 0) { if (edge_lane) E[JWS >= 0 ? JWS : 0] = NEGBIG; } else poke<C>(E, jw_dyn, NEGBIG);   // nothing may leak into k = W+1
            shift_ref();
        } else {
            int fin = __shfl_down_sync(FULL, F[0], 1);
            if (lane == 31) {
                if (NW == 1 || warp == NW - 1) fin = NEGBIG;
                else fin = sm->edgeF[warp + 1];
            }
            best = step_cells<C, 1, MASK, GENERIC>(H1, E, F, Qw, Rw, fin, p, vmask);
            // k = W reads MINUS_INF2 from outside the band (agatha_kernel.h:138); F(0,W) is injected below at d = W-1
            { const int v = (FAST || d + 1 > W) ? NEG16 : NEGBIG; if (JWS >= 0) { if (edge_lane) F[JWS >= 0 ? JWS : 0] = v; } else poke<C>(F, jw_dyn, v); }
            shift_query();
        }
        if (INJECT) { if (d < W) inject(d, u_tag); }
        if (MODE == 4) { if (d < W) inject_static_sw(d, u_tag); }
        // padding columns enter the band only after d_tail, i.e. never right after a PRO step
        if (MASK) { if (has_phantom) phantom_patch(d + 1, UN{}); }    // inputs of the next anti-diagonal, before they are handed over
        if (NW == 1) {
            if (!scan) return false;
            return scan_diag<C>(st, best, d, U, lane, p);
        } else {
            if (U == 1) { if (lane == 31) sm->edgeE[warp] = E[C - 1]; }
            else        { if (lane == 0) sm->edgeF[warp] = F[0]; }
            return scan_diag_group<C, NW>(st, best, d, U, lane, warp, scan, sm, p);
        }
    };

    // --- the slice schedule of the reference (agatha_kernel.h:180-330), replayed per anti-diagonal ----------------
    // first d whose valid k-range is clipped by the far matrix edges (q = qlen-1 or r = tcols-1)
    // (tlen, not tcols: the padding columns need the SLOW body's patches)
    const int d_tail = min(2 * pr.qlen - 2 - W, 2 * pr.tlen - 2 - W) + 1;
    const int d_fast_lo = (W + 2) & ~1;                                   // even, > W: no injection, band edges are real
    const int d_fast_hi = min(d_tail - 1, pr.L - 1) & ~1;                 // FAST pairs (d, d+1) need d+1 < d_tail and d+1 < L

    // --- steady state on 16-bit packed state ---------------------------------------------------------------------
    // Runs anti-diagonals [d, d_fast_hi) in one go (no slice of the reference can end the alignment in this range: the
    // band-exit rule only fires after the band has left the matrix, i.e. beyond d_tail). State is converted at entry and
    // exit. Returns 0: reached d_fast_hi; 1: Z-drop fired on anti-diagonal d; 2: values left the safe 16-bit range at d,
    // state is back in the 32-bit arrays and the caller continues with the 32-bit loop.
    constexpr bool CAN16 = !GENERIC && NW == 1 && JWS >= 0 && C % 8 == 0;
    // The packed tail is compiled only with -DAGATHA_TAIL16=1 (and then still needs AGATHA_S16=7 at run time): bit-exact and
    // 12-14 % faster on equal-length pairs, but its mere presence costs the steady-state loop 15 instructions per two
    // anti-diagonals (uniform-register pressure: loop invariants get recomputed inside the loop) and on mixed-length batches
    // the extra loop costs more in instruction fetch than it saves.
    constexpr bool TAIL16 = CAN16 && (AGATHA_TAIL16 != 0);
    auto run_fast16 = [&](int& d) -> int {
        constexpr int P = C / 2;
        constexpr int JP = (JWS >= 0 ? JWS : 0) % P, JH = (JWS >= 0 ? JWS : 0) / P;   // register / half holding the cell k = +W
        int base = st.max;                                                 // packed value = true value - base
        unsigned A0[P], A1[P], AE[P], AF[P];
        {
            auto sat = [&](int x) { return max(min(x - base, TOP16), FLOOR16); };     // dead cells (NEGBIG) land on FLOOR16
#pragma unroll
            for (int jj = 0; jj < P; jj++) {
                A0[jj] = pack16(sat(H0[jj]), sat(H0[jj + P])); A1[jj] = pack16(sat(H1[jj]), sat(H1[jj + P]));
                AE[jj] = pack16(sat(E[jj]), sat(E[jj + P]));   AF[jj] = pack16(sat(F[jj]), sat(F[jj + P]));
            }
        }
        const unsigned floor2 = pack16(FLOOR16, FLOOR16);
        const unsigned mge2 = pack16raw(-p.ge, -p.ge);                    // two's complement halves
        const int mgoe32 = -(int)((unsigned)p.goe | ((unsigned)p.goe << 16));   // subtracts goe from both halves at once
        // Between two range checks (32 anti-diagonals) the smallest live H falls by at most 16*mismatch and the largest rises
        // by at most 16*match; M = H + s and t = M - goe must stay above the clamp, H + match below 32767.
        // Dead (out-of-band) cells creep upwards by `match` whenever their bases happen to be equal (nothing else feeds them):
        // they are pushed back to the floor at every range check, so they stay below FLOOR16 + 16*match < low_ok.
        const int low_ok = FLOOR16 + 17 * max(max(p.mismatch, p.match), 1) + p.goe + 64;
        const int high_ok = TOP16 - 17 * max(p.match, 0) - 64;
        int neg16 = max(NEG16 - base, FLOOR16);                            // MINUS_INF2 as seen from `base`
        int maxrel = st.max - base;
        auto rel_thr = [&]() { return (st.thr == INT_MIN || st.thr == INT_MAX) ? st.thr : st.thr - base; };
        int thrrel = rel_thr();
        // the position of a new maximum is only needed by a later Z-drop test or at the end: keep a snapshot of the
        // anti-diagonal and search it lazily instead of carrying an index through every cell update
        unsigned S[P];
#pragma unroll
        for (int jj = 0; jj < P; jj++) S[jj] = 0u;
        int snap_d = -1, snap_u = 0, snap_src = 0, snap_h = 0;

        auto search = [&](const unsigned (&A)[P], int h) -> int {          // largest cell index whose value is h, -1 if none
            int jl = -1, jh = -1;
#pragma unroll
            for (int jj = 0; jj < P; jj++) { if (lo16(A[jj]) == h) jl = jj; if (hi16(A[jj]) == h) jh = jj + P; }
            return jh >= 0 ? jh : jl;
        };
        auto resolve = [&]() {
            if (snap_d < 0) return;
            const int jb = __shfl_sync(FULL, search(S, snap_h), snap_src);
            const int k = -W + 2 * (C * snap_src + jb) + snap_u;
            const int r = (snap_d + k) >> 1;
            st.mt = r; st.mq = snap_d - r;
            snap_d = -1;
        };
        // which halves of register jj are live in the band-edge lane (cell j live iff j <= JW for parity 0 / E, j < JW for
        // parity 1 / F ... E is produced by parity-1 cells for parity-0 consumers and vice versa; being conservative costs
        // nothing: a half is treated as dead only if its cell index is beyond JW for BOTH parities)
        auto keep_mask = [&](int jj, bool strict) -> unsigned {            // strict: live iff j < JW, else j <= JW
            const bool lo = strict ? (jj < JWS) : (jj <= JWS), hi = strict ? (jj + P < JWS) : (jj + P <= JWS);
            return (lo ? 0xffffu : 0u) | (hi ? 0xffff0000u : 0u);
        };
        auto unpack = [&]() {
            const bool dead_lane = gl > p.LW;
            auto unp = [&](unsigned x, bool hi, unsigned keep) {
                const int v = (hi ? hi16(x) : lo16(x)) + base;
                const bool live = !dead_lane && (!edge_lane || (keep & (hi ? 0xffff0000u : 0xffffu)));
                return live ? v : NEGBIG;
            };
#pragma unroll
            for (int jj = 0; jj < P; jj++) {
                const unsigned k0 = keep_mask(jj, false), k1 = keep_mask(jj, true);
                H0[jj] = unp(A0[jj], false, k0); H0[jj + P] = unp(A0[jj], true, k0);
                H1[jj] = unp(A1[jj], false, k1); H1[jj + P] = unp(A1[jj], true, k1);
                // The loop always leaves after a UB step. E then feeds the other parity's cells and is live for j < JW only
                // (E[JW] is the dead hand-over into k = W+1); F is live for j <= JW.
                E[jj] = unp(AE[jj], false, k1); E[jj + P] = unp(AE[jj], true, k1);
                F[jj] = unp(AF[jj], false, k0); F[jj + P] = unp(AF[jj], true, k0);
            }
            // after a parity-1 step F[JW] is not a DP value but the MINUS_INF2 that k = +W reads from outside the band
            if (!WODD && edge_lane) F[JWS >= 0 ? JWS : 0] = NEG16;
        };
        // range monitor over the live H values (both parities) + rebasing; false = leave the packed loop
        auto check_range = [&]() -> bool {
            // push the dead positions back to the floor (lanes beyond the band; in the band-edge lane the cells beyond k = +W)
            const bool dead_lane = gl > p.LW;
#pragma unroll
            for (int jj = 0; jj < P; jj++) {
                const unsigned k0 = keep_mask(jj, false), k1 = keep_mask(jj, true), kd = k0 & k1;
                if (dead_lane) { A0[jj] = floor2; A1[jj] = floor2; AE[jj] = floor2; AF[jj] = floor2; }
                else if (edge_lane) {
                    A0[jj] = (A0[jj] & k0) | (floor2 & ~k0); A1[jj] = (A1[jj] & k1) | (floor2 & ~k1);
                    AE[jj] = (AE[jj] & k0) | (floor2 & ~k0); AF[jj] = (AF[jj] & k0) | (floor2 & ~k0);
                    (void)kd;
                }
            }
            unsigned mn2 = 0xffffffffu, mx2 = 0u;
#pragma unroll
            for (int jj = 0; jj < P; jj++) {
                unsigned x0 = A0[jj], x1 = A1[jj];
                mx2 = __vimax3_u16x2(mx2, x0, x1);
                // dead cells are out of the minimum: force their halves to the largest value
                const unsigned k0 = keep_mask(jj, false), k1 = keep_mask(jj, true);
                const unsigned e0 = x0 | ~k0, e1 = x1 | ~k1;
                mn2 = __vimin3_u16x2(mn2, edge_lane ? e0 : x0, edge_lane ? e1 : x1);
            }
            int mn = min(lo16(mn2), hi16(mn2)), mx = max(lo16(mx2), hi16(mx2));
            if (gl > p.LW) mn = TOP16;                                       // lanes beyond the band hold nothing live
            mn = __reduce_min_sync(FULL, mn);
            mx = __reduce_max_sync(FULL, mx);
            if (mn < low_ok || mx > high_ok) return false;
            if (mx > 8192) {                                                 // re-centre: rare (every ~8192 score units)
                const int delta = min(mx, mn - low_ok);
                if (delta > 0) {
                    // the packed add does not saturate: lift everything to FLOOR16 + delta first, then subtract
                    const unsigned md2 = pack16raw(-delta, -delta), lift2 = pack16(FLOOR16 + delta, FLOOR16 + delta);
                    auto shift_down = [&](unsigned x) { return __viaddmax_u16x2(__vimax3_u16x2(x, lift2, lift2), md2, floor2); };
#pragma unroll
                    for (int jj = 0; jj < P; jj++) {
                        A0[jj] = shift_down(A0[jj]); A1[jj] = shift_down(A1[jj]); AE[jj] = shift_down(AE[jj]); AF[jj] = shift_down(AF[jj]);
                    }
                    base += delta;
                    neg16 = max(NEG16 - base, FLOOR16);
                    maxrel = st.max - base; thrrel = rel_thr();
                }
            }
            return true;
        };
        // Termination Condition & Score Update (agatha_kernel.h:292-314) on the packed anti-diagonal, split in two so that
        // the hot loop only carries the common cases. scan_fast: nothing to do, or a new maximum (snapshot for the lazy
        // argmax); returns true when the anti-diagonal might fire Z-drop -> scan_slow, outside the hot loop.
        int ev_lane_h = 0, ev_hrel = 0;
        // tail: valid-cell bits of the anti-diagonal being computed (layout as in step_cells16) and the matching 16-bit masks
        unsigned vm2 = 0u;
        auto half_mask = [&](int jj) -> unsigned { return (unsigned)imad((int)((vm2 >> jj) & 0x00010001u), p.m16, 0); };
        auto scan_fast = [&](unsigned best2, const unsigned (&A)[P], int dd, int u, bool TAILM) -> bool {   // TAILM is a constant at every call site (a generic lambda here crashes cicc 12.9)
            const int lane_h = max(lo16(best2), hi16(best2));
            const int hrel = __reduce_max_sync(FULL, lane_h);
            if (hrel <= maxrel && hrel >= thrrel) return false;
            if (hrel > maxrel) {
                const unsigned who = __ballot_sync(FULL, lane_h == hrel);
#pragma unroll
                for (int jj = 0; jj < P; jj++) S[jj] = TAILM ? (A[jj] & half_mask(jj)) : A[jj];   // cells outside the matrix never match
                snap_d = dd; snap_u = u; snap_src = 31 - __clz((int)who); snap_h = hrel;   // ties -> largest target index
                st.max = hrel + base; st.thr = scan_threshold(st.max, p);
                maxrel = hrel; thrrel = rel_thr();
                return false;
            }
            ev_lane_h = lane_h; ev_hrel = hrel;
            return true;
        };
        auto scan_slow = [&](const unsigned (&A)[P], int dd, int u, bool tailm = false) -> bool {   // tailm: a constant at every call site
            resolve();                                                       // the test needs (mt, mq)
            const unsigned who = __ballot_sync(FULL, ev_lane_h == ev_hrel);
            const int src = 31 - __clz((int)who);
            unsigned B[P];
#pragma unroll
            for (int jj = 0; jj < P; jj++) B[jj] = tailm ? (A[jj] & half_mask(jj)) : A[jj];
            const int jb = __shfl_sync(FULL, search(B, ev_hrel), src);
            return scan_update(st, ev_hrel + base, C * src + jb, dd, u, p);
        };
        // phantom_patch on the packed arrays (tail only; rare: a padding column at the first row of a slice chunk). Run-time
        // cell index -> compare + select of a PRMT selector per register; j < 0 (another lane's cell) changes nothing.
        auto poke16 = [&](unsigned (&A)[P], int j, unsigned v2) {
#pragma unroll
            for (int jj = 0; jj < P; jj++) {
                const unsigned sel = (j == jj) ? 0x3254u : ((j == jj + P) ? 0x7610u : 0x3210u);
                A[jj] = prmt(A[jj], v2, sel);
            }
        };
        auto phantom_patch16 = [&](int dn, auto u_tag) {
            constexpr int U = decltype(u_tag)::value;
            const int qc = (dn - pr.tlen) & ~7;
            if (dn - pr.tlen < 0 || dn - qc >= pr.tcols || qc >= pr.qlen) return;
            if (!(qc == 0 || ((qc >> 3) + pr.pt - 1) % p.sw == 0)) return;
            const int r = dn - qc, k = r - qc;
            if (k > W || k < -W) return;
            const unsigned v2 = pack16(neg16, neg16);
            const int g = (k + W - U) >> 1;
            const int gf = (U == 0) ? g : g + 1;
            const int jf = gf - C * gl;
            poke16(AF, (jf >= 0 && jf < C) ? jf : -1, v2);
            if (r - 1 >= pr.tlen) {
                const int j = g - C * gl;
                const int jj = (j >= 0 && j < C) ? j : -1;
                if (U == 0) poke16(A0, jj, v2); else poke16(A1, jj, v2);
            }
        };
        // one packed anti-diagonal; true = scan_slow must look at it. PRO: an anti-diagonal of the prologue (d <= W): what
        // lies beyond the matrix edges is dead, the caller injects the edge cells after the scan.
        auto step16 = [&](int dd, auto u_tag, auto pro_tag) -> bool {
            constexpr int U = decltype(u_tag)::value;
            constexpr int MODE16 = decltype(pro_tag)::value;                // 0 steady state, 1 prologue, 2 tail
            constexpr bool PRO = MODE16 == 1, TAILM = MODE16 == 2;
            using UN = std::integral_constant<int, 1 - U>;
            if (TAILM) {
                // cells of this anti-diagonal inside the matrix (do_step's jlo / jhi) as one bit per cell
                const int klo = max(-W, max(-dd, dd - 2 * (pr.qlen - 1)));
                const int khi = min(W, min(dd, 2 * (pr.tcols - 1) - dd));
                const int k0 = -W + 2 * C * gl + U;
                const int a = max((klo - k0 + 1) >> 1, 0), b = min((khi - k0) >> 1, C - 1);
                const unsigned fm = (b >= a) ? ((0xffffffffu >> (31 - b)) & (0xffffffffu << a)) : 0u;
                vm2 = (fm & ((1u << P) - 1u)) | ((fm >> P) << 16);
            }
            if (U == 0) {
                unsigned x = __shfl_up_sync(FULL, AE[P - 1], 1);             // neighbour's (E[P-1], E[C-1])
                if (lane == 0) {
                    // left of k = -W: MINUS_INF2; in the prologue the cell is not real yet (dead), and on d == W its left
                    // neighbour is the matrix-edge value E(W,0) (agatha_kernel.h:130), exactly as in do_step
                    if (!PRO) x = (unsigned)(neg16 + BIAS16) << 16;
                    else x = (unsigned)(((dd == W) ? (-(p.goe + p.ge * W) - p.goe - base) : FLOOR16) + BIAS16) << 16;
                }
                const unsigned ein = prmt(x, AE[P - 1], 0x5432);             // lo: neighbour's E[C-1], hi: own E[P-1]
                const unsigned best2 = step_cells16<C, 0, TAILM>(A0, AE, AF, Qw, Rw, ein, p, mge2, mgoe32, floor2, vm2);
                if (edge_lane) AE[JP] = JH ? ((AE[JP] & 0xffffu) | (FLOORU16 << 16)) : ((AE[JP] & 0xffff0000u) | FLOORU16);
                shift_ref();
                if (TAILM) { if (has_phantom) phantom_patch16(dd + 1, UN{}); }   // inputs of the next anti-diagonal
                return scan_fast(best2, A0, dd, 0, TAILM);
            } else {
                unsigned y = __shfl_down_sync(FULL, AF[0], 1);               // neighbour's (F[0], F[P])
                if (lane == 31) y = FLOORU16;                                // right of the last lane: dead
                const unsigned fin = prmt(AF[0], y, 0x5432);                 // lo: own F[P], hi: neighbour's F[0]
                const unsigned best2 = step_cells16<C, 1, TAILM>(A1, AE, AF, Qw, Rw, fin, p, mge2, mgoe32, floor2, vm2);
                // k = +W reads MINUS_INF2 from outside the band; in the prologue that cell is dead until F(0,W) is injected
                if (!PRO) { if (edge_lane) AF[JP] = JH ? ((AF[JP] & 0xffffu) | ((unsigned)(neg16 + BIAS16) << 16)) : ((AF[JP] & 0xffff0000u) | (unsigned)(neg16 + BIAS16)); }
                shift_query();
                if (TAILM) { if (has_phantom) phantom_patch16(dd + 1, UN{}); }
                return scan_fast(best2, A1, dd, 1, TAILM);
            }
        };
        // inject_static on the packed arrays: cell j = 4Q + pos lives in register (4Q + pos) % P, half Q / (NG/2); the value
        // is dropped into that half with a PRMT whose selector is chosen per candidate register
        // Selectors for the block: keep / replace the low half / replace the high half of candidate register 4R + pos, for the
        // block's top group, the group after it, the left group and the group before it. Computed once per block.
        constexpr int NGH = NG / 2, NSEL = NGH > 0 ? NGH : 1;
        unsigned selT[NSEL], selTn[NSEL], selL[NSEL], selLp[NSEL];
        auto block_selectors = [&]() {
#pragma unroll
            for (int R = 0; R < NGH; R++) {
                auto pick = [&](int q) { return (q == R) ? 0x3254u : ((q == R + NGH) ? 0x7610u : 0x3210u); };
                selT[R] = pick(blk_qt); selTn[R] = pick(blk_qt + 1);
                selL[R] = pick(blk_ql); selLp[R] = pick(blk_ql - 1);
            }
        };
        auto inject16 = [&](int dd, auto u_tag, auto s_tag) {
            constexpr int U = decltype(u_tag)::value, S = decltype(s_tag)::value;
            constexpr bool EVEN = (U == UA::value);
            constexpr int PT = EVEN ? S : ((S + 1) & 3), PL = (2 - S) & 3;
            const int hv = -(p.goe + p.ge * (dd + 1)) - base, gv = hv - p.goe;
            const unsigned hv2 = pack16(hv, hv), gv2 = pack16(gv, gv);
            const bool h_top = !(EVEN && S == 3 && blk_last);                 // d + 2 > W: only F is needed
#pragma unroll
            for (int R = 0; R < NGH; R++) {
                const unsigned st_ = (!EVEN && S == 3) ? selTn[R] : selT[R];  // the odd step of S == 3 is already in the next group
                const unsigned sl_ = (S == 3) ? selLp[R] : selL[R];           // ... and the left cell in the previous one
                const unsigned sh_ = h_top ? st_ : 0x3210u;
                AF[4 * R + PT] = prmt(AF[4 * R + PT], gv2, st_);
                AE[4 * R + PL] = prmt(AE[4 * R + PL], gv2, sl_);
                if (U == 0) { A0[4 * R + PT] = prmt(A0[4 * R + PT], hv2, sh_); A0[4 * R + PL] = prmt(A0[4 * R + PL], hv2, sl_); }
                else        { A1[4 * R + PT] = prmt(A1[4 * R + PT], hv2, sh_); A1[4 * R + PL] = prmt(A1[4 * R + PL], hv2, sl_); }
            }
        };
        // The prologue loop: two anti-diagonals per iteration; the place of the pair inside its block of 8 (S) only selects
        // which of four small injection bodies runs (a warp-uniform switch). The executed path of an iteration is ~400
        // instructions -- it has to be small: an earlier version with the 8 anti-diagonals of a block as straight-line code
        // (1,700 instructions, 26 KB) lost 3 of 4 issue slots to instruction fetch (ncu: no_instruction 2.6 warps per issue,
        // I-cache hit rate 62 %). The scan -- including the search for the position of a low maximum -- runs BEFORE the
        // injection, so an injected edge value can never be mistaken for the anti-diagonal's maximum.
        using PRO1 = std::integral_constant<int, 1>;
        using PRO0 = std::integral_constant<int, 0>;
        auto inject16_sw = [&](int dd, auto u_tag, int S) {
            switch (S) {
            case 0: inject16(dd, u_tag, S0{}); break;
            case 1: inject16(dd, u_tag, S1{}); break;
            case 2: inject16(dd, u_tag, S2{}); break;
            default: inject16(dd, u_tag, S3{}); break;
            }
        };
        if (d == 0) {
            // the whole prologue, d = 0 .. W (the caller guarantees STATIC_PRO, p.s16 & 2 and W + 1 < d_tail). Inside it no value
            // can leave the 16-bit range (host-side bound, engine.cu), so there is no range check and no bail-out.
            for (;;) {
                int ev = 0;
#pragma unroll 1
                for (; d < W; d += 2) {
                    const int S = (d >> 1) & 3;
                    if (S == 0) {                                            // a new block of 8
                        blk_qt = ((d + W + 1) >> 3) - NG * gl;
                        blk_ql = ((W + 1 - d) >> 3) - 1 - NG * gl;
                        blk_last = (d + 7 == W);
                        block_selectors();
                    }
                    if (step16(d, UA{}, PRO1{})) { ev = 1; break; }
                    inject16_sw(d, UA{}, S);                                 // d <= W - 1: always an injection
                    if (step16(d + 1, UB{}, PRO1{})) { ev = 2; break; }
                    if (d + 1 < W) inject16_sw(d + 1, UB{}, S);              // nothing after the anti-diagonal d == W
                }
                if (!ev) break;
                // cold: a closer look at the anti-diagonal that might fire, then finish the pair
                const int S = (d >> 1) & 3;
                if (ev == 1) {
                    if (WODD ? scan_slow(A1, d, 1) : scan_slow(A0, d, 0)) { resolve(); return 1; }
                    inject16_sw(d, UA{}, S);
                    if (step16(d + 1, UB{}, PRO1{})) ev = 2;
                }
                if (ev == 2) {
                    if (WODD ? scan_slow(A0, d + 1, 0) : scan_slow(A1, d + 1, 1)) { resolve(); d++; return 1; }
                }
                if (d + 1 < W) inject16_sw(d + 1, UB{}, S);
                d += 2;
            }
            if (!check_range()) { resolve(); unpack(); return 2; }
        } else
        if (!check_range()) return 2;                                        // 32-bit arrays untouched so far
        using TAIL2 = std::integral_constant<int, 2>;
        // the tail's loop (opt-in, p.s16 & 4): same structure as the steady state below plus the slice schedule
        auto run_chunks = [&](int d_hi, auto mode_tag) -> int {             // 0: reached d_hi, 1: Z-drop on d, 2: range, 3: band exit at d
            constexpr bool TAILM = decltype(mode_tag)::value == 2;
            int next_slice = 0;
            if (TAILM) { const int span = 8 * p.sw; next_slice = ((d + span - 1) / span) * span; }
            for (;;) {
                if (TAILM) {
                    // leave well before an anti-diagonal could have no cell inside matrix and band (the reference then reads an
                    // empty ring slot, scan_update's special case): the valid range shrinks by at most 2 per anti-diagonal
                    const int klo = max(-W, max(-d, d - 2 * (pr.qlen - 1))), khi = min(W, min(d, 2 * (pr.tcols - 1) - d));
                    if (khi - klo < 80) return 0;
                }
                const int dchunk = min(d_hi, d + 32);
                int ev = 0;
#pragma unroll 1
                for (; d < dchunk; d += 2) {
                    if (TAILM) {
                        if (d == next_slice) {
                            // slice bounds, agatha_kernel.h:183-191 (same arithmetic as the 32-bit driver below)
                            const int i = d >> 3;
                            int ss = max(0, i - pr.pq + 1);
                            ss = max(ss, (i * 8 + 8 - W) / 2 / 8);
                            int se = min(pr.pt - 1, i + p.sw - 1);
                            se = min(se, ((i + p.sw - 1) * 8 + 7 + W) / 2 / 8);
                            if (ss > se) { ev = 3; break; }
                            next_slice += 8 * p.sw;
                        }
                    }
                    if (step16(d, UA{}, mode_tag)) { ev = 1; break; }
                    if (step16(d + 1, UB{}, mode_tag)) { ev = 2; break; }
                }
                if (ev == 3) return 3;
                if (ev == 1) {
                    if (WODD ? scan_slow(A1, d, 1, TAILM) : scan_slow(A0, d, 0, TAILM)) return 1;
                    if (step16(d + 1, UB{}, mode_tag)) ev = 2;                   // finish the pair (cold copy of the second step)
                    else d += 2;
                }
                if (ev == 2) {
                    if (WODD ? scan_slow(A0, d + 1, 0, TAILM) : scan_slow(A1, d + 1, 1, TAILM)) { d++; return 1; }
                    d += 2;
                }
                if (d >= d_hi) return 0;
                if (!check_range()) return 2;
            }
        };
        // The steady state is written out instead of going through run_chunks: the shared lambda cost its hot loop 15 extra
        // (uniform-datapath) instructions per two anti-diagonals.
        while (d < d_fast_hi) {
            // hot loop: up to 32 anti-diagonals between two range checks, left early only for a possible Z-drop
            const int dchunk = min(d_fast_hi, d + 32);
            int ev = 0;
#pragma unroll 1
            for (; d < dchunk; d += 2) {
                if (step16(d, UA{}, PRO0{})) { ev = 1; break; }
                if (step16(d + 1, UB{}, PRO0{})) { ev = 2; break; }
            }
            if (ev == 1) {
                if (WODD ? scan_slow(A1, d, 1) : scan_slow(A0, d, 0)) { resolve(); return 1; }
                if (step16(d + 1, UB{}, PRO0{})) ev = 2;                     // finish the pair (cold copy of the second step)
                else d += 2;
            }
            if (ev == 2) {
                if (WODD ? scan_slow(A0, d + 1, 0) : scan_slow(A1, d + 1, 1)) { resolve(); d++; return 1; }
                d += 2;
            }
            if (d >= d_fast_hi) break;
            if (!check_range()) { resolve(); unpack(); return 2; }
        }
        // the tail: cells beyond the far matrix edges are masked out of the maximum, padding columns are patched, the slice
        // schedule is checked for band exit. The last anti-diagonals (and the wrap-up scan) are left to the 32-bit driver.
        const int d_end16 = (min(pr.L, 8 * pr.total) - 2) & ~1;
        if constexpr (TAIL16) if ((p.s16 & 4) && d == d_fast_hi && d > W && d < d_end16) {
            if (has_phantom) phantom_patch16(d, UA{});                        // first tail step after the steady state
            const int rc = run_chunks(d_end16, TAIL2{});
            if (rc == 1) { resolve(); return 1; }
            if (rc == 2) { resolve(); unpack(); return 2; }
            if (rc == 3) { resolve(); return 3; }
        }
        resolve();
        unpack();
        return 0;
    };

    int d = 0;
    if (has_phantom) phantom_patch(0, UA{});
    if (NW > 1) {
        // hand-over slots must describe THIS alignment's initial state before the first step reads them
        if (lane == 31) sm->edgeE[warp] = E[C - 1];
        if (lane == 0) sm->edgeF[warp] = F[0];
        __syncthreads();
    }
    bool allow16 = CAN16 && p.s16 != 0;
    for (;;) {
        const int i = ((d >> 3) / p.sw) * p.sw;          // the slice that contains anti-diagonal d
        bool wrap = false;
        int dend;
        if (i >= pr.total) {
            // job wrap-up, agatha_kernel.h:334-356: 8 more anti-diagonals are scanned (without the d < L guard) when the
            // slice loop ended exactly on total_anti_diags; otherwise those ring slots are empty (see oracle).
            if (i != pr.total) break;
            wrap = true; dend = 8 * pr.total + 8;
        } else {
            if (d == 8 * i) {
                // slice bounds, agatha_kernel.h:183-191 (truncating division as in the reference)
                int ss = max(0, i - pr.pq + 1);
                ss = max(ss, (i * 8 + 8 - W) / 2 / 8);
                int se = min(pr.pt - 1, i + p.sw - 1);
                se = min(se, ((i + p.sw - 1) * 8 + 7 + W) / 2 / 8);
                if (ss > se) { stop = AGATHA_STOP_BANDEXIT; d_stop = min(8 * i, pr.L); break; }
            }
            dend = 8 * (i + p.sw);
        }
        bool fired = false, reslice = false, band_exit = false;
        while (d < dend) {
            if ((d >= d_fast_lo && d < d_fast_hi) || (CAN16 && STATIC_PRO && d == 0 && allow16 && (p.s16 & 2) && d_fast_lo < d_fast_hi)) {
                if (CAN16 && allow16) {
                    const int rc = run_fast16(d);          // may cross slice boundaries: re-derive the slice afterwards
                    if (rc == 1) { fired = true; break; }
                    if (rc == 3) { stop = AGATHA_STOP_BANDEXIT; d_stop = min(d, pr.L); band_exit = true; break; }   // at a slice start of the tail
                    if (rc == 2) allow16 = false;
                    reslice = true;
                    break;
                }
                const int dlim = min(dend, d_fast_hi);
#pragma unroll 1
                for (; d < dlim; d += 2) {
                    if (do_step(d, true, UA{}, M0{})) { fired = true; break; }
                    if (do_step(d + 1, true, UB{}, M0{})) { fired = true; d++; break; }
                }
                if (fired) break;
            } else {
                if (has_phantom && d == d_fast_hi && d > 0) phantom_patch(d, UA{});   // first SLOW step after the FAST run
                const bool s0 = wrap || d < pr.L, s1 = wrap || d + 1 < pr.L;
                const bool far = d + 1 >= d_tail, near = d < W;                      // which matrix edges touch this pair of steps
                bool f0, f1 = false;
                if (!far) {
                    if constexpr (STATIC_PRO) {
                        const int d0 = d & ~7;                                   // the aligned block of 8 this pair belongs to
                        blk_qt = ((d0 + W + 1) >> 3) - NG * gl;
                        blk_ql = ((W + 1 - d0) >> 3) - 1 - NG * gl;
                        blk_last = (d0 + 7 == W);
                        blk_S = (d >> 1) & 3;
                        f0 = do_step(d, s0, UA{}, M4{}); if (!f0) f1 = do_step(d + 1, s1, UB{}, M4{});
                    } else {
                        f0 = do_step(d, s0, UA{}, M1{}); if (!f0) f1 = do_step(d + 1, s1, UB{}, M1{});
                    }
                }
                else if (!near) { f0 = do_step(d, s0, UA{}, M2{}); if (!f0) f1 = do_step(d + 1, s1, UB{}, M2{}); }
                else            { f0 = do_step(d, s0, UA{}, M3{}); if (!f0) f1 = do_step(d + 1, s1, UB{}, M3{}); }
                if (f0) { fired = true; break; }
                if (f1) { fired = true; d++; break; }
                d += 2;
            }
        }
        if (fired) { if (d < pr.L) { stop = AGATHA_STOP_ZDROP; d_stop = d + 1; } break; }
        if (band_exit) break;
        if (reslice) continue;
        if (wrap) break;
    }
    out_score = st.max; out_qend = st.mq; out_tend = st.mt; out_stop = stop; out_dstop = d_stop;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent kernel: every group of NW warps pulls alignments from a queue ordered longest-first by the host
// scheduler. NW == 1: 4 independent warps per CTA. NW > 1: one group per CTA (wide bands).
// ---------------------------------------------------------------------------------------------------------------
template <int C, int NW>
struct KernelShape {
    static constexpr int threads = NW == 1 ? 128 : 32 * NW;
    static constexpr int min_blocks = NW == 1 ? (C <= 4 ? 8 : (C <= 24 ? 3 : 2)) : (32 * NW * (C <= 16 ? 128 : (C <= 24 ? 168 : 255)) <= 32768 ? 2 : 1);
};

// Redo mode (template parameter REDO, JobArrays::redo says which instance the host launches): the packed kernel (extend16_kernel.cuh) ran first and marked the pairs it could not finish with
// REDO_MARK in their query-end slot; this kernel then looks at every pair, 32 candidates per queue access, and aligns the
// marked ones. The 32 candidates of one access are spread over the whole (longest-first) order -- positions c, c + n/32,
// c + 2n/32, ... -- because marked pairs come in runs (every pair not longer than the band sits at the end of the order) and a
// run must not end up in one group. A launch without marked pairs costs a few microseconds.
constexpr int REDO_MARK = INT_MIN;

template <int C, int NW, bool WODD, int JWS, bool REDO = false>
__global__ void __launch_bounds__(KernelShape<C, NW>::threads, KernelShape<C, NW>::min_blocks) extend_kernel(JobArrays ja, KernelParams p)
{
    const int lane = threadIdx.x & 31;
    const int warp = NW == 1 ? 0 : (int)(threadIdx.x >> 5);
    __shared__ GroupShared<(NW > 1 ? NW : 1)> smem;
    GroupShared<NW>* sm = reinterpret_cast<GroupShared<NW>*>(&smem);
    // (REDO is a template parameter, not a run-time mode: the plain kernel keeps exactly the loop it always had)
    unsigned todo = 0, cand_idx = 0;
    const unsigned n_chunks = ((unsigned)ja.n + 31u) / 32u;            // redo mode: queue positions
    for (;;) {
        const unsigned step = 1u;
        unsigned job = 0;
        if (!REDO || !todo) {                          // next queue position, uniform across the group
            if (NW == 1) {
                if (lane == 0) job = atomicAdd(ja.counter, step);
                job = __shfl_sync(FULL, job, 0);
            } else {
                __syncthreads();                      // everybody is done with the previous job's shared state
                if (threadIdx.x == 0) sm->job = atomicAdd(ja.counter, step);
                __syncthreads();
                job = sm->job;
            }
            if (job >= (REDO ? n_chunks : (unsigned)ja.n)) break;
        } else if (NW > 1) __syncthreads();           // the next alignment re-initialises the group's shared slots
        unsigned idx;
        if (!REDO) {
            idx = ja.order ? __ldg(ja.order + job) : job;
        } else {
            if (!todo) {
                const unsigned cand = job + (unsigned)lane * n_chunks;
                bool marked = false;
                if (cand < (unsigned)ja.n) {
                    cand_idx = ja.order ? __ldg(ja.order + cand) : cand;
                    marked = ja.qend[cand_idx] == REDO_MARK;       // written by the previous kernel on this stream: plain load
                }
                todo = __ballot_sync(FULL, marked);
                if (!todo) continue;
            }
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            idx = __shfl_sync(FULL, cand_idx, src);
        }

        Pair pr;
        pr.qlen = (int)__ldg(ja.qlen + idx);
        pr.tlen = (int)__ldg(ja.tlen + idx);
        pr.q = ja.qpk + (__ldg(ja.qoff_w + idx) >> 3);     // offsets are in bases, multiples of 8 (agatha_kernel.h:116-117)
        pr.t = ja.tpk + (__ldg(ja.toff_w + idx) >> 3);
        pr.pq = (pr.qlen + 7) >> 3; pr.pt = (pr.tlen + 7) >> 3;        // agatha_kernel.h:120-121
        pr.qwords = pr.pq; pr.twords = pr.pt;
        pr.tcols = 8 * pr.pt;
        pr.total = pr.pq + pr.pt - 1;                                  // :165
        pr.L = pr.qlen + pr.tlen - 1;                                  // :289

        int score = 0, qend = 0, tend = 0, stop = AGATHA_STOP_END, dstop = 0;
        if (pr.qlen > 0 && pr.tlen > 0) {
            if (p.force_generic || has_rare_symbols(pr, lane)) run_pair<C, NW, WODD, JWS, true>(pr, p, lane, warp, sm, score, qend, tend, stop, dstop);
            else run_pair<C, NW, WODD, JWS, false>(pr, p, lane, warp, sm, score, qend, tend, stop, dstop);
        }
        if (lane == 0 && warp == 0) {
            ja.score[idx] = score; ja.qend[idx] = qend; ja.tend[idx] = tend;   // agatha_kernel.h:359-363
            if (ja.stop) ja.stop[idx] = stop;
            if (ja.dstop) ja.dstop[idx] = dstop;
        }
    }
}

}  // namespace agatha
