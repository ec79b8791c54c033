// The 16-bit packed extension kernel: the whole alignment -- prologue, steady state, tail, wrap-up -- on packed state, for
// one-warp AND multi-warp (wide band) shapes. Same results as extend_kernel (extend_kernel.cuh), which stays the general
// kernel: pairs this kernel cannot finish exactly (symbols outside {A,C,G,T,N}, pairs shorter than the band, values that
// leave the safe 16-bit window) are marked in their result slot (REDO_MARK) and extend_kernel, launched right after on the same stream, aligns them.
//
// Replaces agatha_kernel (AGAThA/src/kernels/agatha_kernel.h:49-431); parity spec: SURVEY.md Appendix A, oracle/agatha_oracle.c.
//
// Layout: cells are addressed by anti-diagonal d = q + r and diagonal k = r - q; global lane gl of a
// group of NW warps owns the cells k = -W + 2*(C*gl + j) + u, j in [0,C), u in {0,1}. Register jj of each state array holds
// cell jj in its low half and cell jj + C/2 in its high half, as UNSIGNED biased 16-bit numbers relative to a per-alignment
// base, so VIADDMNMX.U16x2 / VIMNMX3.U16x2 update two cells per instruction.
//
// Drifting representation. ncu shows the ALU pipe at 92 % of its peak in this kernel and the FMA pipe at 14 %, so every
// operation that can be an IMAD should be one. M = H(d-2) + s is an IMAD when no half can borrow, i.e. when the score is
// never negative: the table holds s + X (X = mismatch penalty: 0 for a mismatch, match + X for a match, X - 1 against N)
// and the surplus is absorbed by the representation itself -- a value of anti-diagonal d is stored with the offset
//     off(d) = bias - base + D(d),   D(d) = X * (d >> 1) + (d odd ? X/2 : 0)
// so that H(d-2) + (s + X) is already in the units of anti-diagonal d. E and F are produced in the units of the NEXT
// anti-diagonal (their VIADDMNMX adds delta(d) - ge, the IMAD for t adds delta(d) - goe, delta(d) = D(d+1) - D(d)), and the
// thresholds of the Z-drop scan advance by delta per step. Stored H never decreases along a diagonal, so the low end of the
// 16-bit window needs no monitoring between re-basings; the drift (X/2 per anti-diagonal) is taken out by the re-basing
// that the growing score needs anyway. MINUS_INF2 outside the band is the stored floor: the range monitor hands a pair to
// the general kernel before a live value could get as low as -16384 + goe, so the exact value of the sentinel never matters.
// What differs from the packed loops inside extend_kernel (run_fast16):
//   * no 32-bit state at all: half the live registers, no conversion code, no spills;
//   * the anti-diagonal that holds the running maximum is snapshotted to SHARED memory (3-4 STS.128 per lane) instead of
//     P registers; its position is searched only when a Z-drop test or the end of the alignment needs it;
//   * the steady state runs in blocks of 16 anti-diagonals: the sequence feeds of a block are loaded once at its top (one
//     coalesced word per lane and sequence), the inner two-anti-diagonal body carries no load, no bounds test;
//   * multi-warp groups (W > 1023): lane-edge hand-over and per-warp maxima through shared memory. Prologue and tail run in
//     lock step (one barrier per anti-diagonal; warps without a cell inside the matrix skip the cell update); the steady
//     state is pipelined over two arrive / wait barriers -- a warp publishes its edge value before it waits for its
//     neighbours', maxima are tested two anti-diagonals late, the range monitor runs inside the pipeline (see "pipelined
//     steady state" below). Re-basing and events are decided group-wide;
//   * the tail runs packed to the very end, including the anti-diagonals without any cell and the wrap-up scan
//     (agatha_kernel.h:334-356).
#pragma once

#include "extend_kernel.cuh"

#ifndef AGATHA_MB24
#define AGATHA_MB24 4            // CTAs per SM of the one-warp shapes with C <= 24 (measured on C2: 2 -> 63.4 ms, 3 -> 58.8, 4 -> 57.1)
#endif
#ifndef AGATHA_MBW4
#define AGATHA_MBW4 3
#endif
#ifndef AGATHA_DEADSKIP
#define AGATHA_DEADSKIP 1        // multi-warp groups: warps without a cell inside the matrix skip the cell update (prologue, tail)
#endif
#ifndef AGATHA_ASYNC_RANGE
#define AGATHA_ASYNC_RANGE 1     // multi-warp groups: range monitor inside the pipeline (0: drain the pipeline for every check)
#endif
#ifndef AGATHA_INLINE_EVENTS
#define AGATHA_INLINE_EVENTS 0
#endif

namespace agatha {

template <int C, int NW>
struct Shape16 {
    static constexpr int P = C / 2;
    static constexpr int groups = NW == 1 ? 4 : 1;                 // alignments in flight per CTA
    static constexpr int warps = NW * groups;
    static constexpr int threads = 32 * warps;
    // register budget: C = 8/16/24 run with 128 registers (4 CTAs of 4 warps per SM; C = 24 spills six registers outside its
    // steady-state loop, which costs less than the fourth CTA brings), C = 32 gets 168 (3 CTAs; at 128 it spills in the loop)
    static constexpr int min_blocks = NW == 1 ? (C <= 24 ? AGATHA_MB24 : 3) : (NW == 2 ? 6 : (NW == 4 ? AGATHA_MBW4 : 1));
};

// Shared-memory barrier with split arrive / wait (mbarrier): one elected lane per warp arrives, every lane waits on the phase
// parity. Used by the pipelined steady state of multi-warp groups, where a warp signals "my edge value of this anti-diagonal
// is published" long before it needs its neighbours' (a __syncthreads would make it wait right there).
#ifdef AGATHA_HOST_EMU
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) { emu::mbar_init(b, count); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { emu::mbar_arrive(b); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) { emu::mbar_wait(b, parity); }
__device__ __forceinline__ bool mbar_test(uint64_t* b, unsigned parity) { return emu::mbar_test(b, parity); }
#else
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)                 // release: orders this thread's earlier shared-memory writes
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity)  // acquire: returns once the phase of that parity is complete
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AGATHA_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AGATHA_MBAR_DONE;\n"
        "bra AGATHA_MBAR_WAIT;\n"
        "AGATHA_MBAR_DONE:\n"
        "}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, unsigned parity)  // the same question without waiting (acquire when true)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
    return ok != 0u;
}
#endif

// Scoring and recurrence constants of one alignment, held in plain registers: ptxas otherwise re-materialises them from the
// constant bank inside the hot loop (12 moves per two anti-diagonals for the PRMT table alone).
struct Consts16 {
    unsigned tab_lo, tab_hi;   // PRMT lookup table: byte x -> score + X for code XOR x (never negative)
    unsigned ce[2];            // by parity U: delta - ge in both halves (two's complement), the addend of the E/F extension
    int ct[2];                 // by parity U: (delta - goe) * 0x10001, the addend of t = M - goe
    int one;                   // 1, opaque to ptxas: keeps a*1 + c on the FMA pipe (IMAD) instead of the ALU pipe (IADD3)
    int m16;                   // 0xffff, opaque too: (bit pair) * 0xffff widens a valid-cell bit to a 16-bit mask on the FMA pipe
};

// One anti-diagonal step for the C cells of parity U owned by this lane, on packed state in the drifting representation
// (file header). Recurrence: CORE_COMPUTE, agatha_kernel.h:20-30 (gap opens from M = diag + s). Per register (two cells):
// ALU pipe PRMT (score pair), VIMNMX3 (H), 2 x VIADDMNMX (E, F), 1/2 VIMNMX3 (maximum); FMA pipe 2 x IMAD (M, t).
// TAILM: `vm2` holds one bit per cell of this anti-diagonal that lies inside the matrix (cells 0..P-1 in bits 0.., cells
// P..C-1 in bits 16..); the others still take part in the recurrence (nothing inside the matrix ever reads them) but are
// kept out of the maximum. Returns the packed maximum of H.
template <int C, int U, bool TAILM>
__device__ __forceinline__ unsigned cells16(unsigned (&H)[C / 2], unsigned (&E)[C / 2], unsigned (&F)[C / 2],
                                            const uint32_t (&Qw)[C / 8], const uint32_t (&Rw)[C / 8],
                                            unsigned edge_in, const Consts16& k, unsigned vm2)
{
    constexpr int P = C / 2, NWORD = C / 8;
    unsigned sc[2 * NWORD];
#pragma unroll
    for (int w = 0; w < NWORD; w++) {
        const unsigned x = Qw[w] ^ Rw[w];
        sc[2 * w] = prmt(k.tab_lo, k.tab_hi, x);
        sc[2 * w + 1] = prmt(k.tab_lo, k.tab_hi, x >> 16);
    }
    unsigned best = 0u, pend = 0u;
    unsigned oE[P], oF[P];                                     // inputs of this step (the outputs of the previous one)
#pragma unroll
    for (int jj = 0; jj < P; jj++) { oE[jj] = E[jj]; oF[jj] = F[jj]; }
#pragma unroll
    for (int t_ = 0; t_ < P; t_++) {
        const int jj = t_;
        unsigned ein, fin;
        if (U == 0) { ein = (jj == 0) ? edge_in : oE[jj - 1]; fin = oF[jj]; }
        else        { ein = oE[jj]; fin = (jj == P - 1) ? edge_in : oF[jj + 1]; }
        // zero-extended score pair: byte (a&3) of sc[a>>2] for cell a = jj, byte (b&3) of sc[b>>2] for cell b = jj + P
        // (selector nibbles 1 and 3 replicate the sign of a byte that is never negative: zero)
        const int a = jj, b = jj + P;
        const unsigned sel = (unsigned)(a & 3) | ((unsigned)((a & 3) | 8) << 4) | ((unsigned)(4 + (b & 3)) << 8) | ((unsigned)((4 + (b & 3)) | 8) << 12);
        const unsigned s2 = prmt(sc[a >> 2], sc[b >> 2], sel);
        const unsigned m = (unsigned)imad((int)H[jj], k.one, (int)s2);       // H(d-2,k) + s + X: no half can carry or borrow
        const unsigned t = (unsigned)imad((int)m, k.one, k.ct[U]);           // M - goe, in the units of the next anti-diagonal
        const unsigned h = __vimax3_u16x2(m, ein, fin);
        E[jj] = __viaddmax_u16x2(ein, k.ce[U], t);
        F[jj] = __viaddmax_u16x2(fin, k.ce[U], t);
        H[jj] = h;
        unsigned hm = h;
        if (TAILM) hm = h & (unsigned)imad((int)((vm2 >> jj) & 0x00010001u), k.m16, 0);
        if (t_ == 0) pend = hm;
        else if (t_ == 1) best = __vimax3_u16x2(pend, hm, hm);
        else if (t_ & 1) best = __vimax3_u16x2(best, pend, hm);
        else pend = hm;
    }
    if (P & 1) best = __vimax3_u16x2(best, pend, pend);
    return best;
}

// The edge value a step of parity U hands to the neighbouring lane, computed on its own: (F[0], F[P]) for U == 0, (E[P-1],
// E[C-1]) for U == 1. It depends on this lane's state only -- not on the edge value coming in during the same step -- which is
// what lets a warp of a multi-warp group publish it before it waits for its neighbours (pipelined steady state). Same
// operations on the same inputs as the corresponding register in cells16.
template <int C, int U>
__device__ __forceinline__ unsigned out_edge16(const unsigned (&H)[C / 2], const unsigned (&E)[C / 2], const unsigned (&F)[C / 2],
                                               const uint32_t (&Qw)[C / 8], const uint32_t (&Rw)[C / 8], const Consts16& k)
{
    constexpr int P = C / 2;
    constexpr int jj = (U == 0) ? 0 : P - 1, a = jj, b = jj + P;
    const unsigned xa = Qw[a >> 3] ^ Rw[a >> 3], xb = Qw[b >> 3] ^ Rw[b >> 3];
    const unsigned sa = prmt(k.tab_lo, k.tab_hi, ((a >> 2) & 1) ? (xa >> 16) : xa);
    const unsigned sb = prmt(k.tab_lo, k.tab_hi, ((b >> 2) & 1) ? (xb >> 16) : xb);
    const unsigned sel = (unsigned)(a & 3) | ((unsigned)((a & 3) | 8) << 4) | ((unsigned)(4 + (b & 3)) << 8) | ((unsigned)((4 + (b & 3)) | 8) << 12);
    const unsigned s2 = prmt(sa, sb, sel);
    const unsigned m = (unsigned)imad((int)H[jj], k.one, (int)s2);
    const unsigned t = (unsigned)imad((int)m, k.one, k.ct[U]);
    return __viaddmax_u16x2(U == 0 ? F[0] : E[P - 1], k.ce[U], t);
}

// Symbols the packed kernel scores: query {A,C,G,T}, target {A,C,G,T,N} (codes 0..3 and TCODE_N, which the windows hold as
// TCODE_N & 7 = 5 so that every code XOR stays below 8 and indexes the table directly). Anything else -- including an N in
// the read -- goes to the general kernel.
__device__ __forceinline__ bool outside_packed_alphabet(const Pair& pr, int lane)
{
    bool rare = false;
    for (int i = lane; i < pr.qwords; i += 32) {
        const uint32_t w = __ldg(pr.q + i);
        const bool last = (i == pr.qwords - 1);                            // the last word carries the 'N' padding (QCODE_N = 4)
        const uint32_t keep = last ? (0xffffffffu << (4 * ((8 - (pr.qlen & 7)) & 7))) : 0xffffffffu;   // first base in the top nibble
        rare |= (w & keep & 0xccccccccu) != 0u;
    }
    for (int i = lane; i < pr.twords; i += 32) {
        const uint32_t w = __ldg(pr.t + i);
        const uint32_t hi = (w | (w >> 1)) & 0x44444444u;                 // bit2 set <=> nibble >= 4 (bit3|bit2)
        const uint32_t x = w ^ 0xddddddddu;                               // nibble == 0 <=> code 13
        const uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;   // 1 <=> nibble != 13
        rare |= ((hi >> 2) & nz) != 0u;
    }
    return __any_sync(FULL, rare);
}

template <int C, int NW>
struct Shared16 {
    unsigned snap[Shape16<C, NW>::warps][C / 2][32];   // snapshot of the anti-diagonal holding the running maximum: [warp][register][lane]
    // hand-over between the warps of a group: the step of anti-diagonal d writes slot (d >> 1) & 1, the step of d + 1 reads it
    // (two slots: in the pipelined steady state a warp publishes the value of d + 2 while a neighbour may still read that of d)
    unsigned edgeE[2][NW];    // (E[P-1], E[C-1]) of lane 31 of each warp, published by steps of parity 1 (even d)
    unsigned edgeF[2][NW];    // (F[0], F[P]) of lane 0 of each warp, published by steps of parity 0 (odd d)
    unsigned evsnap[NW > 1 ? Shape16<C, NW>::warps : 1][NW > 1 ? C / 2 : 1][32];   // pipelined steady state: H of an anti-diagonal that may fire Z-drop
    int scan_h[4][NW];        // per-warp maximum of the anti-diagonal (stored domain), ring of four slots indexed by d & 3
    int rng[2][NW];           // per-warp live minimum / maximum for the range monitor
    int bcast;                // cell index found by the owner warp
    unsigned job;
    // NW > 1: arrive / wait barriers of the pipelined steady state, one arrival per warp and anti-diagonal; anti-diagonal d uses
    // barrier d & 1 (a warp arrives for d before it waits for d-1, and a barrier counts arrivals, not warps: on a single barrier
    // the early arrival would be taken for the phase that is still open)
    alignas(8) uint64_t mbar[2];
};

// One alignment on packed state. Returns false when the pair has to be redone by the general kernel (nothing written).
template <int C, int NW, int JWS>
__device__ __forceinline__ bool run_pair16(const Pair& pr, const KernelParams& p, int lane, int warp, int cta_warp, Shared16<C, NW>* sm, unsigned& mphase,
                                           int& out_score, int& out_qend, int& out_tend, int& out_stop, int& out_dstop)
{
    static_assert(C % 8 == 0 && C <= 32 && JWS >= 0 && JWS < C && (JWS & 7) == 7, "packed kernel: C multiple of 8, W = 7 (mod 8)");
    constexpr int P = C / 2, NWORD = C / 8, NG = C / 4, NGH = NG / 2;
    constexpr int JP = JWS % P, JH = JWS / P;          // register / half holding the cell k = +W
    using U1 = std::integral_constant<int, 1>;          // parity class of even anti-diagonals (W is odd)
    using U0 = std::integral_constant<int, 0>;
    using S0 = std::integral_constant<int, 0>;
    using S1 = std::integral_constant<int, 1>;
    using S2 = std::integral_constant<int, 2>;
    using S3 = std::integral_constant<int, 3>;
    using MSTEADY = std::integral_constant<int, 0>;
    using MPRO = std::integral_constant<int, 1>;
    using MTAIL = std::integral_constant<int, 2>;
    const int W = p.W;
    const int gl = 32 * warp + lane;
    const bool edge_lane = (gl == p.LW), dead_lane = gl > p.LW;
    // PRMT selector that replaces the half of register JP holding the cell k = +W (band-edge lane) or nothing (other lanes)
    const unsigned edge_sel = edge_lane ? (JH ? 0x7610u : 0x3254u) : 0x3210u;

    // ---- what this kernel does not handle --------------------------------------------------------------------------
    // first d whose valid k-range is clipped by the far matrix edges (tlen, not tcols: padding columns need the tail's patches)
    const int d_tail = min(2 * pr.qlen - 2 - W, 2 * pr.tlen - 2 - W) + 1;
    const int d_fast_hi = min(d_tail - 1, pr.L - 1) & ~1;
    if (W + 1 >= d_fast_hi) return false;                               // pair not longer than the band: general kernel
    if (outside_packed_alphabet(pr, lane)) return false;

    // ---- state -----------------------------------------------------------------------------------------------------------
    const unsigned floor2 = FLOORU16 * 0x10001u;
    const int X = p.mismatch;                                            // surplus of the biased scores, absorbed by the drift
    const int de = X >> 1, dod = X - de;                                 // delta(d) for even / odd d
    auto D = [&](int dd) { return X * (dd >> 1) + ((dd & 1) ? de : 0); };   // cumulative drift (arithmetic shift: D(-2) = -X)
    const int bias = p.bias16;                                           // stored = true - base + bias + D(d)
    auto pk = [&](int v) { return ((unsigned)v & 0xffffu) * 0x10001u; };   // a stored value in both halves
    Consts16 k;
    k.tab_lo = p.tabb_lo; k.tab_hi = p.tabb_hi;
    k.ce[1] = pack16raw(de - p.ge, de - p.ge);   k.ct[1] = (de - p.goe) * 0x10001;      // parity 1 <-> even d
    k.ce[0] = pack16raw(dod - p.ge, dod - p.ge); k.ct[0] = (dod - p.goe) * 0x10001;
    k.one = p.one; k.m16 = p.m16;
    {
        // ptxas keeps warp-uniform values in uniform registers and copies them into a vector register in front of every PRMT
        // (the table operand cannot be uniform): 12 moves per two anti-diagonals. A value that went through a shuffle is not
        // uniform in its eyes, so the table stays in two ordinary registers for the whole alignment.
        const unsigned z = (NW > 1) ? (unsigned)(lane >= p.k32) : __shfl_sync(FULL, 0u, lane);   // 0; k32 = 32 is opaque to ptxas
        k.tab_lo ^= z; k.tab_hi ^= z;
    }
    unsigned A0[P], A1[P], AE[P], AF[P];
#pragma unroll
    for (int jj = 0; jj < P; jj++) { A0[jj] = floor2; A1[jj] = floor2; AE[jj] = floor2; AF[jj] = floor2; }
    int base = 0;                                                        // re-basing so far (stored = true - base + bias + D(d))

    // sequence windows at d = 0: nibble j <-> query[qtop - j], target[rbot + j]
    int qtop = (W >> 1) - C * gl;
    int rbot = ((1 - W) >> 1) + C * gl;
    uint32_t Qw[NWORD], Rw[NWORD];
#pragma unroll
    for (int w = 0; w < NWORD; w++) { Qw[w] = 0u; Rw[w] = 0u; }
#pragma unroll 1
    for (int j = 0; j < C; j++) {
        const unsigned qb = qbase(pr, qtop - j) << (4 * (j & 7)), tb = (tbase(pr, rbot + j) & 7u) << (4 * (j & 7));
#pragma unroll
        for (int w = 0; w < NWORD; w++) if (w == (j >> 3)) { Qw[w] |= qb; Rw[w] |= tb; }
    }
    uint32_t qfeed, rfeed;
    auto refeed = [&]() {                                // feeds for the generic (one base at a time) window shifts
        const int nq = qtop + 1, nb = rbot + C;
        qfeed = load_qword(pr, nq >> 3) << (4 * (nq & 7));
        rfeed = (load_tword(pr, nb >> 3) & 0x77777777u) >> (4 * (nb & 7));
    };
    refeed();
    auto shift_query = [&]() {                           // qtop -> qtop + 1
        const int nq = qtop + 1;
        if ((nq & 7) == 0) qfeed = load_qword(pr, nq >> 3);
#pragma unroll
        for (int w = NWORD - 1; w > 0; w--) Qw[w] = __funnelshift_l(Qw[w - 1], Qw[w], 4);
        Qw[0] = __funnelshift_l(qfeed, Qw[0], 4);
        qfeed <<= 4;
        qtop = nq;
    };
    auto shift_ref = [&]() {                             // rbot -> rbot + 1
        const int nb = rbot + C;
        if ((nb & 7) == 0) rfeed = load_tword(pr, nb >> 3) & 0x77777777u;
#pragma unroll
        for (int w = 0; w < NWORD - 1; w++) Rw[w] = __funnelshift_r(Rw[w], Rw[w + 1], 4);
        Rw[NWORD - 1] = __funnelshift_r(Rw[NWORD - 1], rfeed, 4);
        rfeed >>= 4;
        rbot++;
    };
    // the same shifts inside a steady-state block: the feed word was prepared at the top of the block
    auto shift_query_blk = [&]() {
#pragma unroll
        for (int w = NWORD - 1; w > 0; w--) Qw[w] = __funnelshift_l(Qw[w - 1], Qw[w], 4);
        Qw[0] = __funnelshift_l(qfeed, Qw[0], 4);
        qfeed <<= 4;
    };
    auto shift_ref_blk = [&]() {
#pragma unroll
        for (int w = 0; w < NWORD - 1; w++) Rw[w] = __funnelshift_r(Rw[w], Rw[w + 1], 4);
        Rw[NWORD - 1] = __funnelshift_r(Rw[NWORD - 1], rfeed, 4);
        rfeed >>= 4;
    };
    // window positions as a function of the anti-diagonal about to be computed (even d)
    auto window_pos = [&](int dd) { qtop = ((dd + W) >> 1) - C * gl; rbot = ((dd - W + 1) >> 1) + C * gl; };

    // Cell g of the group (g = C * lane + cell; the same number in every lane) takes the value v2: the register is a
    // warp-uniform choice, which lane and which half goes into one PRMT selector. g outside the group changes nothing.
    auto poke16 = [&](unsigned (&A)[P], int g, unsigned v2) {
        const int owner = g >= 0 ? g / C : -1, j = g - owner * C;
        const int jr = j >= P ? j - P : j;
        const unsigned sel = (gl == owner) ? (j >= P ? 0x7610u : 0x3254u) : 0x3210u;
#pragma unroll
        for (int jj = 0; jj < P; jj++) if (jj == jr) A[jj] = prmt(A[jj], v2, sel);
    };

    // ---- boundary: H(-1,-1) = 0 and the virtual cells of "anti-diagonal -1" (agatha_kernel.h:126-148) -------------------
    {
        // H(-1,-1) = 0 is read as the diagonal of step 0 (units of anti-diagonal -2); H(-1,0) = H(0,-1) = -goe as diagonals of
        // step 1 (units of -1); F(0,0) = E(0,0) = -2 goe as gap inputs of step 0 (units of 0)
        const int hv = -p.goe + bias + D(-1), gv = -2 * p.goe + bias + D(0);
        poke16(A1, W >> 1, pk(0 + bias + D(-2)));                      // k = 0 belongs to parity 1 (W odd)
        const int jt = (W + 1) >> 1;                                   // top: (q=-1, r=0) at k = 1
        poke16(A0, jt, pk(hv)); poke16(AF, jt, pk(gv));
        const int jl = (W - 1) >> 1;                                   // left: (q=0, r=-1) at k = -1
        poke16(A0, jl, pk(hv)); poke16(AE, jl, pk(gv));
    }

    ScanState st = {0, 0, 0, scan_threshold(0, p)};                    // agatha_kernel.h:158-161
    int stop = AGATHA_STOP_END, d_stop = pr.L;
    const bool has_phantom = pr.tcols > pr.tlen;

    // Z-drop scan state of the hot loops, in the STORED units of the anti-diagonal being computed: thrS = (running maximum)
    // - Z, advanced by delta at the start of every step. "Nothing can happen on this anti-diagonal" (thr <= h <= max) is ONE
    // unsigned comparison, h - thrS <= Z (Z-drop off: Z = 2^30, so only a new maximum can fail it). The maximum itself is
    // kept as (mx_h, mx_d): its stored value and the anti-diagonal whose units that value is in; st.max / st.thr are brought
    // up to date only where the cold paths need them.
    const int Zeff = p.Z < 0 ? (1 << 30) : p.Z;
    int mx_h = bias + D(0), mx_d = 0;                                    // true 0 (agatha_kernel.h:158)
    int thrS = mx_h - Zeff - dod;                                        // step 0 adds delta(-1) = dod first
    auto sync_state = [&]() { st.max = mx_h - bias + base - D(mx_d); st.thr = scan_threshold(st.max, p); };   // cold: before scan_update / output
    // Range monitor (every RANGE16_PERIOD = 512 anti-diagonals; RANGE16_PAIRS = 257). Stored H never decreases along a diagonal, so the low end only moves when the
    // state is re-based; low_ok keeps t = M - goe + delta free of borrows and the floor below every live candidate. At the top
    // a live value rises by at most (match + X) per two anti-diagonals. neg_ok: the lowest TRUE live value for which
    // MINUS_INF2 (-16384) outside the band still loses every maximum it enters during the next 2 * RANGE16_PAIRS anti-diagonals.
    // Dead (out-of-band) cells creep upwards from the floor by at most (match + X) per two anti-diagonals between two checks
    // (they are pushed back at every check) and must stay below every live value: hence the RANGE16_PAIRS * (match + X).
    const int low_ok = (int)FLOORU16 + RANGE16_PAIRS * (max(p.match, 0) + X) + p.goe + 2 * p.ge + 64;
    const int high_ok = 65535 - RANGE16_PAIRS * (max(p.match, 0) + X) - 64;
    const int neg_ok = NEG16 + p.goe + p.ge + RANGE16_PAIRS * X + 64;

    // ---- snapshot of the anti-diagonal that holds the running maximum (shared memory) ------------------------------------
    int snap_d = -1, snap_u = 0, snap_w = 0, snap_h = 0;
    // (scalar stores: a vector store would need the state registers in aligned quadruples, which costs moves in the hot path)
    auto snapshot = [&](const unsigned (&A)[P], unsigned keep_bits, bool masked) {
#pragma unroll
        for (int jj = 0; jj < P; jj++) {
            unsigned x = A[jj];
            if (masked) x &= (unsigned)imad((int)((keep_bits >> jj) & 0x00010001u), p.m16, 0);   // cells outside the matrix never match
            sm->snap[cta_warp][jj][lane] = x;
        }
    };
    auto snapshot_ev = [&](const unsigned (&A)[P]) {                    // NW > 1, cold: an anti-diagonal that may fire, for pipe_drain
#pragma unroll
        for (int jj = 0; jj < (NW > 1 ? P : 0); jj++) sm->evsnap[cta_warp][jj][lane] = A[jj];
    };
    auto search = [&](const unsigned (&A)[P], int h) -> int {          // largest cell index whose stored value is h, -1 if none
        int jl = -1, jh = -1;
#pragma unroll
        for (int jj = 0; jj < P; jj++) { if ((int)(A[jj] & 0xffffu) == h) jl = jj; if ((int)(A[jj] >> 16) == h) jh = jj + P; }
        return jh >= 0 ? jh : jl;
    };
    // group-wide broadcast of a value computed by warp `w` (NW > 1 only; cold paths)
    auto group_bcast = [&](int v, int w) -> int {
        if (NW == 1) return v;
        if (warp == w && lane == 0) sm->bcast = v;
        __syncthreads();
        const int r = sm->bcast;
        __syncthreads();
        return r;
    };
    auto resolve = [&]() {                                              // (mt, mq) of the running maximum
        if (snap_d < 0) return;
        int g = 0;
        if (NW == 1 || warp == snap_w) {
            unsigned S[P];
#pragma unroll
            for (int jj = 0; jj < P; jj++) S[jj] = sm->snap[cta_warp][jj][lane];
            // ties go to the largest target index: the highest lane that holds the value, its highest cell
            const int jb = search(S, snap_h);
            const int src = 31 - __clz((int)__ballot_sync(FULL, jb >= 0));
            g = C * (32 * warp + src) + __shfl_sync(FULL, jb, src);
        }
        g = group_bcast(g, snap_w);
        const int k = -W + 2 * g + snap_u;
        const int r = (snap_d + k) >> 1;
        st.mt = r; st.mq = snap_d - r;
        snap_d = -1;
    };

    // which halves of register jj are live in the band-edge lane: cell j is in band iff j <= JW (parity 0) / j < JW (parity 1)
    auto keep_mask = [&](int jj, bool strict) -> unsigned {
        const bool lo = strict ? (jj < JWS) : (jj <= JWS), hi = strict ? (jj + P < JWS) : (jj + P <= JWS);
        return (lo ? 0xffffu : 0u) | (hi ? 0xffff0000u : 0u);
    };

    // ---- range monitor over the live H values (both parities) + re-basing; false = values leave the safe window --------
    // the warp's part: dead positions pushed back to the floor, minimum and maximum of the warp's live H values
    // floor_tag: whether cells at the stored floor can occur among the live ones (tail, multi-warp groups) -- see below
    auto range_local = [&](int& mn, int& mx, auto floor_tag) {
        constexpr bool FLOORS = decltype(floor_tag)::value != 0;
        // push the dead positions back to the floor (lanes beyond the band; in the band-edge lane the cells beyond k = +W)
#pragma unroll
        for (int jj = 0; jj < P; jj++) {
            const unsigned k0 = keep_mask(jj, false), k1 = keep_mask(jj, true);
            if (dead_lane) { A0[jj] = floor2; A1[jj] = floor2; AE[jj] = floor2; AF[jj] = floor2; }
            else if (edge_lane) {
                A0[jj] = (A0[jj] & k0) | (floor2 & ~k0); A1[jj] = (A1[jj] & k1) | (floor2 & ~k1);
                AE[jj] = (AE[jj] & k0) | (floor2 & ~k0); AF[jj] = (AF[jj] & k0) | (floor2 & ~k0);
            }
        }
        // The minimum is taken over the values ABOVE the floor. A cell that holds exactly the floor is not a live value on its
        // way down (those are caught between the floor and low_ok, the margins see to that): it is the stored MINUS_INF2 -- an
        // input that the padding-column patch of the tail has just reset, or a cell of a warp that skipped the prologue. Such a
        // cell must not hand the pair to the general kernel. (x - floor - 1 wraps the floor to 0xffff.) The steady state of a
        // one-warp group has neither (FLOORS false): plain minimum.
        const unsigned off2 = pack16raw(-((int)FLOORU16 + 1), -((int)FLOORU16 + 1));
        unsigned mn2 = 0xffffffffu, mx2 = 0u;
#pragma unroll
        for (int jj = 0; jj < P; jj++) {
            const unsigned x0 = A0[jj], x1 = A1[jj];
            mx2 = __vimax3_u16x2(mx2, x0, x1);
            const unsigned k0 = keep_mask(jj, false), k1 = keep_mask(jj, true);   // dead cells are out of the minimum
            const unsigned y0 = edge_lane ? (x0 | ~k0) : x0, y1 = edge_lane ? (x1 | ~k1) : x1;
            if (FLOORS) mn2 = __vimin3_u16x2(mn2, __viaddmax_u16x2(y0, off2, 0u), __viaddmax_u16x2(y1, off2, 0u));
            else mn2 = __vimin3_u16x2(mn2, y0, y1);
        }
        mn = (int)min(mn2 & 0xffffu, mn2 >> 16) + (FLOORS ? (int)FLOORU16 + 1 : 0); mx = (int)max(mx2 & 0xffffu, mx2 >> 16);
        if (dead_lane) mn = 65535 + (int)FLOORU16;
        mn = __reduce_min_sync(FULL, mn);
        mx = __reduce_max_sync(FULL, mx);
    };
    // what the monitor looks for: true = the full check has something to do (hand-over to the general kernel or re-basing)
    auto range_alarm = [&](int mn, int mx, int d_now) -> bool {
        return mn < low_ok || mn - bias + base - D(d_now) < neg_ok || mx > high_ok - 1024;
    };
    auto check_range = [&](int d_now, auto floor_tag) -> bool {         // d_now: the anti-diagonal about to be computed
        int mn, mx;
        range_local(mn, mx, floor_tag);
        if (NW > 1) {
            if (lane == 0) { sm->rng[0][warp] = mn; sm->rng[1][warp] = mx; }
            __syncthreads();
            mn = __reduce_min_sync(FULL, lane < NW ? sm->rng[0][lane] : INT_MAX);
            mx = __reduce_max_sync(FULL, lane < NW ? sm->rng[1][lane] : 0);
            __syncthreads();
        }
        if (mn < low_ok) return false;
        if (mn - bias + base - D(d_now) < neg_ok) return false;          // MINUS_INF2 could matter soon: general kernel
        if (mx > high_ok - 1024) {                                       // re-base: bring the lowest live value down to low_ok
            const int delta = mn - low_ok;
            if (mx - delta > high_ok) return false;                      // live values span more than the window
            if (delta > 0) {
                // the packed add does not saturate: lift everything to the floor + delta first, then subtract
                const unsigned md2 = pack16raw(-delta, -delta), lift2 = pk((int)FLOORU16 + delta);
                auto shift_down = [&](unsigned x) { return __viaddmax_u16x2(__vimax3_u16x2(x, lift2, lift2), md2, floor2); };
#pragma unroll
                for (int jj = 0; jj < P; jj++) {
                    A0[jj] = shift_down(A0[jj]); A1[jj] = shift_down(A1[jj]); AE[jj] = shift_down(AE[jj]); AF[jj] = shift_down(AF[jj]);
                }
                // (the snapshot and snap_h stay in the units they were taken in: they are only ever compared with each other,
                // and with the drift a snapshot that followed every re-basing would sink below the floor)
                base += delta;
                mx_h -= delta;
                thrS -= delta;
            }
        }
        if (NW > 1) {
            // the hand-over slots hold values of the old base (and possibly dead cells that were just reset): publish again
            if (lane == 31) { sm->edgeE[0][warp] = AE[P - 1]; sm->edgeE[1][warp] = AE[P - 1]; }
            if (lane == 0) { sm->edgeF[0][warp] = AF[0]; sm->edgeF[1][warp] = AF[0]; }
            __syncthreads();
        }
        return true;
    };

    // ---- Termination Condition & Score Update (agatha_kernel.h:292-314) on a packed anti-diagonal -------------------------
    // scan_fast runs in the hot loops: nothing to do, or a new maximum (snapshot for the lazy argmax); returns true when the
    // anti-diagonal might fire Z-drop -> scan_slow, outside the hot loops.
    int ev_h = 0;
    bool ev_empty = false;
    unsigned vm2 = 0u;                                                   // tail: valid-cell bits of the anti-diagonal being computed
    auto owner_warp = [&](int h, int slot) -> int {                     // highest warp whose maximum is h (ties -> largest target index)
        if (NW == 1) return 0;
        const unsigned who = __ballot_sync(FULL, lane < NW && sm->scan_h[slot][lane < NW ? lane : 0] == h);
        return 31 - __clz((int)who);
    };
    // The maximum over the warp of both halves of best2 without touching the ALU pipe (which bounds this kernel): one warp
    // reduction of the packed word yields the largest high half, one of the word shifted left by 16 (FMA pipe) the largest low
    // half. Which lane and cell hold the maximum is found by searching the anti-diagonal, and only when that is needed.
    auto scan_fast = [&](unsigned best2, const unsigned (&A)[P], int dd, int u, bool tailm) -> bool {   // tailm: constant at every call site
        const unsigned rhi = __reduce_max_sync(FULL, best2);
        const unsigned rlo = __reduce_max_sync(FULL, (unsigned)imad((int)best2, p.k65536, 0));
        int h = (int)(max(rhi, rlo) >> 16);
        if (NW > 1) {
            if (lane == 0) sm->scan_h[dd & 3][warp] = h;
            __syncthreads();
            h = __reduce_max_sync(FULL, lane < NW ? sm->scan_h[dd & 3][lane] : INT_MIN);
        }
        if ((unsigned)(h - thrS) <= (unsigned)Zeff) return false;
        if (h > thrS) {                                                  // above the window: a new maximum
            const int ow = owner_warp(h, dd & 3);
            if (NW == 1 || warp == ow) snapshot(A, vm2, tailm);
            snap_d = dd; snap_u = u; snap_w = ow; snap_h = h;
            mx_h = h; mx_d = dd;
            thrS = h - Zeff;
            return false;
        }
        ev_h = h;
        return true;
    };
    // mask: 0 none, 1 the tail's valid-cell bits (vm2), 2 prologue -- the cells beyond the near matrix edges (|k| > dd) hold
    // matrix-edge values injected after the step, which are not cells of this anti-diagonal (constant at every call site)
    auto scan_slow = [&](const unsigned (&A)[P], int dd, int u, int mask) -> bool {
        resolve();                                                       // the test needs (mt, mq) ...
        sync_state();                                                    // ... and the maximum as a true score
        if (ev_empty) { ev_empty = false; return scan_update(st, INT_MIN, 0, dd, u, p); }   // no cell on this anti-diagonal
        const int ow = owner_warp(ev_h, dd & 3);
        int g = 0;
        if (NW == 1 || warp == ow) {
            unsigned m2 = vm2;
            if (mask == 2) {
                const int k0 = -W + 2 * C * gl + u;
                const int a_ = max((-dd - k0 + 1) >> 1, 0), b_ = min((dd - k0) >> 1, C - 1);
                const unsigned fm = (b_ >= a_) ? ((0xffffffffu >> (31 - b_)) & (0xffffffffu << a_)) : 0u;
                m2 = (fm & ((1u << P) - 1u)) | ((fm >> P) << 16);
            }
            unsigned B[P];
#pragma unroll
            for (int jj = 0; jj < P; jj++) B[jj] = mask ? (A[jj] & (unsigned)imad((int)((m2 >> jj) & 0x00010001u), p.m16, 0)) : A[jj];
            const int jb = search(B, ev_h);
            const int src = 31 - __clz((int)__ballot_sync(FULL, jb >= 0));
            g = C * (32 * warp + src) + __shfl_sync(FULL, jb, src);
        }
        g = group_bcast(g, ow);
        return scan_update(st, ev_h - bias + base - D(dd), g, dd, u, p);
    };

    // phantom (padding) target columns: their F and diagonal inputs restart from MINUS_INF2 at the first row of every slice
    // chunk of the last target block (agatha_kernel.h:206-221 reload, :272-279 never stored); see oracle.
    auto phantom_patch16 = [&](int dn, auto u_tag) {
        constexpr int U = decltype(u_tag)::value;
        const int qc = (dn - pr.tlen) & ~7;                              // the only multiple of 8 in (dn - tcols, dn - tlen]
        if (dn - pr.tlen < 0 || dn - qc >= pr.tcols || qc >= pr.qlen) return;
        if (!(qc == 0 || ((qc >> 3) + pr.pt - 1) % p.sw == 0)) return;
        const int r = dn - qc, k = r - qc;
        if (k > W || k < -W) return;
        const unsigned v2 = floor2;                                      // MINUS_INF2 (file header)
        const int g = (k + W - U) >> 1;
        const int gf = (U == 0) ? g : g + 1;                             // its F input: U==0 reads F[j], U==1 reads F[j+1]
        poke16(AF, gf, v2);
        if (r - 1 >= pr.tlen) { if (U == 0) poke16(A0, g, v2); else poke16(A1, g, v2); }
    };


    // ---- one packed anti-diagonal; true = scan_slow must look at it ---------------------------------------------------------
    // MODE 0 steady state (every in-band cell inside the matrix); 1 prologue (d <= W: what lies beyond the matrix edges is
    // dead, the caller injects the edge cells after the scan); 2 tail (far edges: cells outside the matrix are masked out of
    // the maximum, padding columns are patched). BLK: inside a steady-state block (feeds prepared by the caller).
    auto step16 = [&](int dd, bool scan, auto u_tag, auto mode_tag, auto blk_tag, auto& inject) -> bool {
        constexpr int U = decltype(u_tag)::value;
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr bool PRO = MODE == 1, TAILM = MODE == 2, BLK = decltype(blk_tag)::value != 0;
        using UN = std::integral_constant<int, 1 - U>;
        thrS += (U == 1) ? dod : de;                                     // into the units of this anti-diagonal: delta(dd - 1)
        bool empty = false;
        // Multi-warp groups: a warp none of whose cells lies inside the matrix on this anti-diagonal (prologue: beyond the near
        // edges, with a margin for the injected matrix-edge cells; tail: beyond the far edges) skips the cell update. Nothing
        // inside the matrix ever reads such cells; the warp keeps shifting its sequence windows and stays in lock step.
        bool wdead = false;
        const int wk_lo = -W + 2 * C * 32 * warp, wk_hi = wk_lo + 2 * C * 32 - 1;      // diagonals owned by this warp
        if (AGATHA_DEADSKIP && NW > 1 && PRO) wdead = (wk_lo > dd + 3) || (wk_hi < -(dd + 3));
        if (TAILM) {
            // cells of this anti-diagonal inside the matrix (padding columns count: agatha_kernel.h CORE_COMPUTE has no r < tlen guard)
            const int klo = max(-W, max(-dd, dd - 2 * (pr.qlen - 1)));
            const int khi = min(W, min(dd, 2 * (pr.tcols - 1) - dd));
            const int k0 = -W + 2 * C * gl + U;
            const int a = max((klo - k0 + 1) >> 1, 0), b = min((khi - k0) >> 1, C - 1);
            const unsigned fm = (b >= a) ? ((0xffffffffu >> (31 - b)) & (0xffffffffu << a)) : 0u;
            vm2 = (fm & ((1u << P) - 1u)) | ((fm >> P) << 16);
            // no cell at all: parity of the valid range included (k has the parity of dd)
            empty = ((khi - ((khi ^ dd) & 1)) < (klo + ((klo ^ dd) & 1)));
            if (AGATHA_DEADSKIP && NW > 1) wdead = (wk_lo > khi) || (wk_hi < klo);
        }
        unsigned best2 = 0u;
        if (U == 0) {
            unsigned x = __shfl_up_sync(FULL, AE[P - 1], 1);             // neighbour's (E[P-1], E[C-1])
            if (lane == 0) {
                if (NW == 1 || warp == 0) {
                    // left of k = -W: MINUS_INF2; in the prologue the cell is not real yet (dead), and on d == W its left
                    // neighbour is the matrix-edge value E(W,0) (agatha_kernel.h:130)
                    if (!PRO) x = FLOORU16 << 16;
                    else x = (unsigned)((dd == W) ? (-(p.goe + p.ge * W) - p.goe + bias + D(W)) : (int)FLOORU16) << 16;
                } else x = sm->edgeE[((dd - 1) >> 1) & 1][warp - 1];
            }
            const unsigned ein = prmt(x, AE[P - 1], 0x5432);             // lo: neighbour's E[C-1], hi: own E[P-1]
            if (!wdead) {
                best2 = cells16<C, 0, TAILM>(A0, AE, AF, Qw, Rw, ein, k, vm2);
                AE[JP] = prmt(AE[JP], floor2, edge_sel);                // band-edge lane: nothing leaks into k = W+1
            }
            if (BLK) shift_ref_blk(); else shift_ref();
            if (TAILM) { if (has_phantom) phantom_patch16(dd + 1, UN{}); }   // inputs of the next anti-diagonal
            // prologue: the matrix-edge cells of this anti-diagonal, BEFORE the hand-over slot is published (an injected E / F
            // at the first or last cell of a warp is read by the neighbouring warp)
            if (PRO) { if (dd < W && !wdead) inject(dd, u_tag); }
            if (NW > 1) { if (lane == 0) sm->edgeF[(dd >> 1) & 1][warp] = AF[0]; }
        } else {
            unsigned y = __shfl_down_sync(FULL, AF[0], 1);               // neighbour's (F[0], F[P])
            if (lane == 31) {
                if (NW == 1 || warp == NW - 1) y = FLOORU16;             // right of the last lane: dead
                else y = sm->edgeF[((dd - 1) >> 1) & 1][warp + 1];
            }
            const unsigned fin = prmt(AF[0], y, 0x5432);                 // lo: own F[P], hi: neighbour's F[0]
            if (!wdead) {
                best2 = cells16<C, 1, TAILM>(A1, AE, AF, Qw, Rw, fin, k, vm2);
                // k = +W reads MINUS_INF2 from outside the band; in the prologue that cell is dead until F(0,W) is injected
                if (!PRO) AF[JP] = prmt(AF[JP], floor2, edge_sel);
            }
            if (BLK) shift_query_blk(); else shift_query();
            if (TAILM) { if (has_phantom) phantom_patch16(dd + 1, UN{}); }
            if (PRO) { if (dd < W && !wdead) inject(dd, u_tag); }
            if (NW > 1) { if (lane == 31) sm->edgeE[(dd >> 1) & 1][warp] = AE[P - 1]; }
        }
        if (!scan) {                                                     // computed, not scanned (d >= L before the wrap-up)
            if (NW > 1) __syncthreads();
            return false;
        }
        if (TAILM) {
            if (empty) {                                                 // reads as an empty ring slot (agatha_kernel.h:152,296-299)
                if (NW > 1) __syncthreads();
                ev_empty = true;
                return true;
            }
        }
        return (U == 0) ? scan_fast(best2, A0, dd, 0, TAILM) : scan_fast(best2, A1, dd, 1, TAILM);
    };

    // ---- multi-warp groups: the pipelined steady state ---------------------------------------------------------------------
    // In lock step (step16) a group pays one __syncthreads per anti-diagonal with a serial tail behind it (shared-memory read of
    // the per-warp maxima, warp reduction, test, then the shuffle and shared-memory read of the neighbour's edge value) while
    // the ALU pipe idles. Here a warp only ever waits for what it is about to read, and it has published its own part a whole
    // step earlier:
    //   * the edge value a step hands to the neighbouring warp does not depend on the edge value coming in during the same
    //     step (out_edge16), so step d computes it FIRST, publishes it and arrives on the barrier (phase d); only then does it
    //     wait for phase d-1 -- the neighbours' edge values of d-1, published at the top of THEIR step d-1. A warp can run a
    //     full anti-diagonal ahead of the slowest one before it has to wait (edge slots are double-buffered for this).
    //   * per-warp maxima are published at the end of a step and tested TWO steps later: phase d-1 complete means every warp
    //     has finished step d-2 and published its maximum, and the H values of d-2 are still in the registers step d is about
    //     to overwrite (new maximum: snapshot as usual; possible Z-drop: saved to shared memory, the cold path looks at them
    //     after the step). Terminations are therefore noticed a few anti-diagonals late -- results do not depend on cells
    //     computed after the one that stops the alignment.
    // The pipeline is entered behind a __syncthreads (pipe_enter) and left through pipe_drain, which tests what is pending.
    int pipe_lo = 0;                                                     // first anti-diagonal computed by the current pipeline run
    bool pipe_ready = false;                                             // the phase the next wait is about is already known to be complete
    int ev_d = 0;                                                        // anti-diagonal saved in evsnap
    auto pipe_enter = [&](int d0) {
        if (lane == 0) mbar_arrive(&sm->mbar[(d0 - 1) & 1]);             // phase "d0-1": those edge values are published (lock step)
        pipe_lo = d0;
        pipe_ready = false;
    };
    // test of anti-diagonal dt given the per-warp maxima `v` (lane w < NW: warp w's; INT_MIN elsewhere); A: its H values.
    // True = it might fire Z-drop (cold path).
    auto pipe_test = [&](int dt, int v, const unsigned (&A)[P], int u) -> bool {
        const int h = __reduce_max_sync(FULL, v);
        thrS += (u == 1) ? dod : de;                                     // into the units of dt
        if ((unsigned)(h - thrS) <= (unsigned)Zeff) return false;
        if (h > thrS) {                                                  // a new maximum
            const int ow = 31 - __clz((int)__ballot_sync(FULL, v == h));
            if (warp == ow) snapshot(A, 0u, false);
            snap_d = dt; snap_u = u; snap_w = ow; snap_h = h;
            mx_h = h; mx_d = dt;
            thrS = h - Zeff;
            return false;
        }
        ev_h = h;
        return true;
    };
    auto load_maxima = [&](int dt) -> int { return (lane < NW) ? sm->scan_h[dt & 3][lane < NW ? lane : 0] : INT_MIN; };
    // one steady-state anti-diagonal inside a block; true = the test of dd-2 wants the cold path (its H values are in evsnap;
    // step dd has been computed all the same)
    auto pipe_step = [&](int dd, auto u_tag) -> bool {
        constexpr int U = decltype(u_tag)::value;
        const int ws = (dd >> 1) & 1, rs = ((dd - 1) >> 1) & 1;          // edge slot this step writes / reads
        bool slow = false;
        unsigned best2;
        if (U == 0) {
            unsigned x = __shfl_up_sync(FULL, AE[P - 1], 1);
            const unsigned out = out_edge16<C, 0>(A0, AE, AF, Qw, Rw, k);
            if (lane == 0) { sm->edgeF[ws][warp] = out; mbar_arrive(&sm->mbar[1]); }      // odd anti-diagonal
            if (!pipe_ready) mbar_wait(&sm->mbar[0], mphase & 1u);
            mphase ^= 1u;
            const int v = load_maxima(dd - 2);
            if (lane == 0) x = (warp == 0) ? (FLOORU16 << 16) : sm->edgeE[rs][warp > 0 ? warp - 1 : 0];
            if (dd - 2 >= pipe_lo) slow = pipe_test(dd - 2, v, A0, 0);
            if (slow) { snapshot_ev(A0); ev_d = dd - 2; }
            const unsigned ein = prmt(x, AE[P - 1], 0x5432);
            best2 = cells16<C, 0, false>(A0, AE, AF, Qw, Rw, ein, k, 0u);
            // has everybody arrived for THIS anti-diagonal already? (asked here, answered by the time the next step wants to know:
            // the barrier unit takes some fifty cycles to answer, which the step after the next wait would otherwise sit out)
            pipe_ready = mbar_test(&sm->mbar[1], (mphase >> 1) & 1u);
            AE[JP] = prmt(AE[JP], floor2, edge_sel);
            shift_ref_blk();
        } else {
            unsigned y = __shfl_down_sync(FULL, AF[0], 1);
            const unsigned out = out_edge16<C, 1>(A1, AE, AF, Qw, Rw, k);
            if (lane == 31) { sm->edgeE[ws][warp] = out; mbar_arrive(&sm->mbar[0]); }     // even anti-diagonal
            if (!pipe_ready) mbar_wait(&sm->mbar[1], (mphase >> 1) & 1u);
            mphase ^= 2u;
            const int v = load_maxima(dd - 2);
            if (lane == 31) y = (warp == NW - 1) ? FLOORU16 : sm->edgeF[rs][warp < NW - 1 ? warp + 1 : 0];
            if (dd - 2 >= pipe_lo) slow = pipe_test(dd - 2, v, A1, 1);
            if (slow) { snapshot_ev(A1); ev_d = dd - 2; }
            const unsigned fin = prmt(AF[0], y, 0x5432);
            best2 = cells16<C, 1, false>(A1, AE, AF, Qw, Rw, fin, k, 0u);
            pipe_ready = mbar_test(&sm->mbar[0], mphase & 1u);
            AF[JP] = prmt(AF[JP], floor2, edge_sel);
            shift_query_blk();
        }
        const unsigned rhi = __reduce_max_sync(FULL, best2);
        const unsigned rlo = __reduce_max_sync(FULL, (unsigned)imad((int)best2, p.k65536, 0));
        // published by the lane that arrives in the NEXT step, so that its arrival orders this write as well
        if (lane == (U == 0 ? 31 : 0)) sm->scan_h[dd & 3][warp] = (int)(max(rhi, rlo) >> 16);
        return slow;
    };
    // Leave the pipeline in front of step dn (steps up to dn-1 are computed): test what is still pending -- dn-3 if
    // pipe_step(dn-1) saved it for the cold path (`slow`), then dn-2 and dn-1. True = the alignment stops (d = the anti-diagonal
    // that fired).
    int d = 0;
    bool fired = false;
    auto pipe_drain = [&](int dn, bool slow) -> bool {
        {
            // phase dn-1, the one arrival nobody has waited for yet (every phase is waited for exactly once, whatever happens
            // below): everybody has finished step dn-2
            const int b = (dn - 1) & 1;
            if (!pipe_ready) mbar_wait(&sm->mbar[b], (mphase >> b) & 1u);
            mphase ^= 1u << b;
            pipe_ready = false;
        }
#pragma unroll 1
        for (int i = slow ? 0 : 1; i < 3; i++) {
            const int dt = dn - 3 + i;
            if (i == 2) __syncthreads();                                 // the maxima of dn-1 were published after the last arrival
            if (dt < pipe_lo) continue;
            const int u = (dt & 1) ? 0 : 1;                              // W is odd: even anti-diagonals are parity class 1
            unsigned B[P];
#pragma unroll
            for (int jj = 0; jj < P; jj++) B[jj] = (i == 0) ? sm->evsnap[cta_warp][jj][lane] : (u ? A1[jj] : A0[jj]);
            bool look = true;                                            // i == 0: pipe_test has run (thrS, ev_h are those of dn-3)
            if (i) look = pipe_test(dt, load_maxima(dt), B, u);
            if (look) { if (scan_slow(B, dt, u, 0)) { d = dt; fired = true; return true; } }
        }
        return false;
    };

    // ---- prologue: matrix-edge injection at compile-time positions inside aligned blocks of 8 anti-diagonals ----------------
    // (see extend_kernel.cuh inject_static for the geometry.) The two virtual cells move by one cell every second
    // anti-diagonal, in opposite directions; with W = 7 (mod 8) their position INSIDE a group of four cells is a compile-time
    // function of the step's place S in its block; only the group is a run-time value, turned into PRMT selectors per block.
    int blk_qt = 0, blk_ql = 0;
    bool blk_last = false;
    unsigned selT[NGH], selTn[NGH], selL[NGH], selLp[NGH];
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
        constexpr bool EVEN = (U == 1);
        constexpr int PT = EVEN ? S : ((S + 1) & 3), PL = (2 - S) & 3;
        // H of the virtual cells in the units of anti-diagonal dd, their E / F in the units of dd + 1 (base is 0 in the prologue)
        const int hv = -(p.goe + p.ge * (dd + 1)), gv = hv - p.goe;
        const unsigned hv2 = pk(hv + bias + D(dd)), gv2 = pk(gv + bias + D(dd + 1));
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
    auto inject16_sw = [&](int dd, auto u_tag, int S) {
        switch (S) {
        case 0: inject16(dd, u_tag, S0{}); break;
        case 1: inject16(dd, u_tag, S1{}); break;
        case 2: inject16(dd, u_tag, S2{}); break;
        default: inject16(dd, u_tag, S3{}); break;
        }
    };

    if (NW > 1) {
        // hand-over slots must describe THIS alignment's initial state before the first step reads them
        if (lane == 31) { sm->edgeE[0][warp] = AE[P - 1]; sm->edgeE[1][warp] = AE[P - 1]; }
        if (lane == 0) { sm->edgeF[0][warp] = AF[0]; sm->edgeF[1][warp] = AF[0]; }
        __syncthreads();
    }

    using NOBLK = std::integral_constant<int, 0>;
    using INBLK = std::integral_constant<int, 1>;
    // The prologue, d = 0 .. W. Inside it no value can leave the 16-bit range (host-side bound, extend_dispatch.h), so there
    // is no range check. The scan -- including the search for the position of a low maximum -- runs BEFORE the injection, so
    // an injected edge value can never be mistaken for the anti-diagonal's maximum.
    auto inject_pro = [&](int dd, auto u_tag) { inject16_sw(dd, u_tag, (dd >> 1) & 3); };
    for (;;) {
        int ev = 0;
#pragma unroll 1
        for (; d < W; d += 2) {
            if ((d & 7) == 0) {                                          // a new block of 8
                blk_qt = ((d + W + 1) >> 3) - NG * gl;
                blk_ql = ((W + 1 - d) >> 3) - 1 - NG * gl;
                blk_last = (d + 7 == W);
                block_selectors();
            }
            if (step16(d, true, U1{}, MPRO{}, NOBLK{}, inject_pro)) { ev = 1; break; }
            if (step16(d + 1, true, U0{}, MPRO{}, NOBLK{}, inject_pro)) { ev = 2; break; }
        }
        if (!ev) break;
        // cold: a closer look at the anti-diagonal that might fire, then finish the pair of steps
        if (ev == 1) {
            if (scan_slow(A1, d, 1, 2)) { fired = true; break; }
            if (step16(d + 1, true, U0{}, MPRO{}, NOBLK{}, inject_pro)) ev = 2;
        }
        if (ev == 2) {
            if (scan_slow(A0, d + 1, 0, 2)) { fired = true; d++; break; }
        }
        d += 2;
    }

    bool redo = false, band_exit = false;
    int d_check = -64;                                                   // anti-diagonal of the last range check
    // feeds of a steady-state block: 8 query bases (first in the top nibble) and 8 target bases (first in the bottom nibble)
    auto block_feeds = [&]() {
        window_pos(d);
        const int nq = qtop + 1, nb = rbot + C;
        const uint32_t q0 = load_qword(pr, nq >> 3), q1 = load_qword(pr, (nq >> 3) + 1);
        const uint32_t t0 = load_tword(pr, nb >> 3), t1 = load_tword(pr, (nb >> 3) + 1);
        qfeed = __funnelshift_l(q1, q0, 4 * (nq & 7));
        rfeed = __funnelshift_r(t0, t1, 4 * (nb & 7)) & 0x77777777u;
    };
    if constexpr (NW > 1) {
        // ---- steady state of a multi-warp group: pipelined blocks of 16 anti-diagonals ------------------------------------------
        // The range monitor does not interrupt the pipeline: every RANGE16_PERIOD anti-diagonals each warp pushes its dead positions back, takes
        // the minimum / maximum of its own live values and publishes them; two anti-diagonals later (everybody has published by
        // then, and the period + 2 anti-diagonals is what the margins of low_ok / high_ok cover) the group-wide values are looked at. Only
        // when they call for a re-basing or a hand-over does the group leave the pipeline for the full check (check_range).
        while (!fired && d + 16 <= d_fast_hi) {
            if (d - d_check >= RANGE16_PERIOD) { if (!check_range(d, S1{})) { redo = true; break; } d_check = d; }
            pipe_enter(d);                                               // (everybody is behind a barrier here)
            int ev = 0;
            bool rng_pending = false;
            for (;;) {
                if (AGATHA_ASYNC_RANGE && d - d_check >= RANGE16_PERIOD) {
                    int mn, mx;
                    range_local(mn, mx, S1{});
                    if (lane == 31) { sm->rng[0][warp] = mn; sm->rng[1][warp] = mx; }   // lane 31 arrives next (even anti-diagonal)
                    rng_pending = true;
                    d_check = d;
                }
                block_feeds();
                const int dblk = d + 16;
#pragma unroll 1
                for (; d < dblk; d += 2) {
                    if (pipe_step(d, U1{})) { ev = 1; break; }
                    if (pipe_step(d + 1, U0{})) { ev = 2; break; }
                    if (rng_pending) {                                   // the wait of step d+1 is behind us: every warp's values are there
                        rng_pending = false;
                        const int mn = __reduce_min_sync(FULL, lane < NW ? sm->rng[0][lane < NW ? lane : 0] : INT_MAX);
                        const int mx = __reduce_max_sync(FULL, lane < NW ? sm->rng[1][lane < NW ? lane : 0] : 0);
                        if (range_alarm(mn, mx, d_check)) { ev = 3; d += 2; break; }
                    }
                }
                if (ev || d + 16 > d_fast_hi) break;
                if (!AGATHA_ASYNC_RANGE && d - d_check >= RANGE16_PERIOD) break;
            }
            // cold: out of the pipeline. ev == 1: steps up to d are computed, ev == 2: up to d + 1, otherwise up to d - 1.
            if (pipe_drain(d + (ev < 3 ? ev : 0), ev == 1 || ev == 2)) break;
            if (ev == 1) {
                // back to an even anti-diagonal in lock step (the feeds of the interrupted block are still in place)
                if (step16(d + 1, true, U0{}, MSTEADY{}, INBLK{}, inject_pro)) { if (scan_slow(A0, d + 1, 0, 0)) { fired = true; d++; break; } }
            }
            if (ev == 1 || ev == 2) d += 2;
            if (ev == 3 || rng_pending) d_check = d - RANGE16_PERIOD;                // the full check, now
        }
        if (!fired && !redo) { window_pos(d); refeed(); }
    } else if (!fired) {
        // ---- steady state: blocks of 16 anti-diagonals, range check every RANGE16_PERIOD --------------------------------------
        while (d + 16 <= d_fast_hi && !fired && !redo) {
            if (d - d_check >= RANGE16_PERIOD) { if (!check_range(d, S0{})) { redo = true; break; } d_check = d; }
            block_feeds();
            const int dblk = d + 16;
#if AGATHA_INLINE_EVENTS
#pragma unroll 1
            for (; d < dblk; d += 2) {
                if (step16(d, true, U1{}, MSTEADY{}, INBLK{}, inject_pro)) { if (scan_slow(A1, d, 1, 0)) { fired = true; break; } }
                if (step16(d + 1, true, U0{}, MSTEADY{}, INBLK{}, inject_pro)) { if (scan_slow(A0, d + 1, 0, 0)) { fired = true; d++; break; } }
            }
#else
            while (d < dblk) {
                int ev = 0;
#pragma unroll 1
                for (; d < dblk; d += 2) {
                    if (step16(d, true, U1{}, MSTEADY{}, INBLK{}, inject_pro)) { ev = 1; break; }
                    if (step16(d + 1, true, U0{}, MSTEADY{}, INBLK{}, inject_pro)) { ev = 2; break; }
                }
                if (!ev) break;
                if (ev == 1) {
                    if (scan_slow(A1, d, 1, 0)) { fired = true; break; }
                    if (step16(d + 1, true, U0{}, MSTEADY{}, INBLK{}, inject_pro)) ev = 2;   // finish the pair (cold copy of the second step)
                }
                if (ev == 2) {
                    if (scan_slow(A0, d + 1, 0, 0)) { fired = true; d++; break; }
                }
                d += 2;
            }
#endif
        }
        window_pos(d);                                                   // (d is even here unless fired)
        refeed();
    }

    if (!fired && !redo) {
        // ---- tail: the rest of the steady range, then the far matrix edges; slice schedule of the reference ----------------
        // (agatha_kernel.h:180-330): band exit is tested at every slice start, diagonals d >= L are computed but not scanned,
        // and when the slices end exactly on total_anti_diags 8 more anti-diagonals are scanned (wrap-up, :334-356).
        const int span = 8 * p.sw;
        const int n_slices = (pr.total + p.sw - 1) / p.sw;
        const bool wrap = (pr.total % p.sw) == 0;
        const int wrap_lo = wrap ? 8 * pr.total : INT_MAX;
        const int d_end = wrap ? 8 * pr.total + 8 : min(pr.L, 8 * n_slices * p.sw);   // nothing observable beyond
        int next_slice = ((d + span - 1) / span) * span;
        // (Measured alternatives, both slower on every workload: the patch outside the loop -- the loop leaves for cold code at
        // every anti-diagonal that needs it, up to 7 in every 8 * slice_width -- and a loop of single steps with everything
        // special between two runs. Leaving a hot loop costs more instruction fetch than the patch code costs inside it; what
        // did help is the patch's cheaper poke, a warp-uniform register choice instead of a select chain per register.)
        if (has_phantom) phantom_patch16(d, U1{});                       // inputs of the first tail step
        while (d < d_end && !fired && !redo && !band_exit) {
            if (d - d_check >= RANGE16_PERIOD) { if (!check_range(d, S1{})) { redo = true; break; } d_check = d; }
            const int dchunk = min(d_end, d_check + RANGE16_PERIOD);
            int ev = 0;
#pragma unroll 1
            for (; d < dchunk; d += 2) {
                if (d >= next_slice && d < 8 * pr.total) {
                    // slice bounds, agatha_kernel.h:183-191 (truncating division as in the reference)
                    const int i = next_slice >> 3;
                    int ss = max(0, i - pr.pq + 1);
                    ss = max(ss, (i * 8 + 8 - W) / 2 / 8);
                    int se = min(pr.pt - 1, i + p.sw - 1);
                    se = min(se, ((i + p.sw - 1) * 8 + 7 + W) / 2 / 8);
                    if (ss > se) { ev = 3; break; }
                    next_slice += span;
                }
                if (step16(d, d < pr.L || d >= wrap_lo, U1{}, MTAIL{}, NOBLK{}, inject_pro)) { ev = 1; break; }
                if (d + 1 < d_end) { if (step16(d + 1, d + 1 < pr.L || d + 1 >= wrap_lo, U0{}, MTAIL{}, NOBLK{}, inject_pro)) { ev = 2; break; } }
            }
            if (ev == 3) { stop = AGATHA_STOP_BANDEXIT; d_stop = min(next_slice, pr.L); band_exit = true; break; }
            if (ev == 1) {
                if (scan_slow(A1, d, 1, 1)) { fired = true; break; }
                if (d + 1 < d_end) { if (step16(d + 1, d + 1 < pr.L || d + 1 >= wrap_lo, U0{}, MTAIL{}, NOBLK{}, inject_pro)) ev = 2; }
            }
            if (ev == 2) {
                if (scan_slow(A0, d + 1, 0, 1)) { fired = true; d++; break; }
            }
            if (ev) d += 2;
        }
    }
    if (redo) return false;
    if (fired) { if (d < pr.L) { stop = AGATHA_STOP_ZDROP; d_stop = d + 1; } }
    resolve();
    sync_state();
    out_score = st.max; out_qend = st.mq; out_tend = st.mt; out_stop = stop; out_dstop = d_stop;
    return true;
}

// Persistent kernel: every group of NW warps pulls alignments from a queue ordered longest-first by the host scheduler.
template <int C, int NW, int JWS>
__global__ void __launch_bounds__(Shape16<C, NW>::threads, Shape16<C, NW>::min_blocks) extend16_kernel(JobArrays ja, KernelParams p)
{
    const int lane = threadIdx.x & 31;
    const int cta_warp = (int)(threadIdx.x >> 5);
    const int warp = NW == 1 ? 0 : cta_warp;
    __shared__ Shared16<C, NW> smem;
    Shared16<C, NW>* sm = &smem;
    unsigned mphase = 0u;                             // NW > 1: bit b = parity of the phase this warp waits for next on barrier b
    if (NW > 1) {
        if (threadIdx.x == 0) { mbar_init(&sm->mbar[0], NW); mbar_init(&sm->mbar[1], NW); }
        __syncthreads();
    }
    for (;;) {
        unsigned job = 0;
        if (NW == 1) {
            if (lane == 0) job = atomicAdd(ja.counter, 1u);
            job = __shfl_sync(FULL, job, 0);
        } else {
            __syncthreads();                          // everybody is done with the previous job's shared state
            if (threadIdx.x == 0) sm->job = atomicAdd(ja.counter, 1u);
            __syncthreads();
            job = sm->job;
        }
        if (job >= (unsigned)ja.n) break;
        const unsigned idx = ja.order ? __ldg(ja.order + job) : job;

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
        bool done = true;
        if (pr.qlen > 0 && pr.tlen > 0) done = run_pair16<C, NW, JWS>(pr, p, lane, warp, cta_warp, sm, mphase, score, qend, tend, stop, dstop);
        if (lane == 0 && warp == 0) {
            if (done) {
                ja.score[idx] = score; ja.qend[idx] = qend; ja.tend[idx] = tend;   // agatha_kernel.h:359-363
                if (ja.stop) ja.stop[idx] = stop;
                if (ja.dstop) ja.dstop[idx] = dstop;
            } else {
                ja.qend[idx] = REDO_MARK;             // the general kernel, launched next on this stream, takes it from here
            }
        }
    }
}

}  // namespace agatha
