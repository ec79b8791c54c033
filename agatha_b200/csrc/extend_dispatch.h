// Host-side choice of the extension-kernel variant for a set of alignment parameters: kernel shape (cells per lane, warps
// per alignment), band-edge template constant, run-time kernel parameters. Pure host logic with no CUDA call, shared by
// engine.cu (which launches the chosen variant) and by the SIMT emulation under tests/emu (which runs it on the CPU).
#pragma once
#include <cstdlib>

#include "extend_kernel.cuh"
#include "extend16_kernel.cuh"
#include "engine_internal.h"

namespace agatha {

// Kernel shape for a band width: a group of NW warps covers global cell indices g in [0, 32*NW*C); the band needs g <= W.
struct Shape { int C, NW; };
inline Shape shape_for(int W)
{
    if (W < 32 * 2) return {2, 1};
    if (W < 32 * 4) return {4, 1};
    if (W < 32 * 8) return {8, 1};
    if (W < 32 * 16) return {16, 1};
    if (W < 32 * 24) return {24, 1};
    if (W < 32 * 32) return {32, 1};
    if (W < 2 * 32 * 32) return {32, 2};
    if (W < 4 * 32 * 32) return {32, 4};
    if (W < 8 * 32 * 32) return {32, 8};
    return {0, 0};
}

inline bool fast_table_ok(const agatha_params_t* p)
{
    return p->match >= -128 && p->match <= 127 && p->mismatch >= 1 && p->mismatch <= 128;
}

// AGATHA_S16 (A/B measurements): "0" disables every 16-bit packed path, "1" = general kernel only, packed
// steady state, "2" = general kernel only, packed prologue + steady state (the round-1 default), "7" = also its tail (needs
// a -DAGATHA_TAIL16=1 build); unset = the packed kernel first, general kernel for what it marks. Read once per process; the
// emulation's tests override it through s16_mode().
inline int& s16_mode()
{
    static int mode = [] {
        const char* env = getenv("AGATHA_S16");
        if (!env || !env[0] || env[1]) return -1;
        return (env[0] >= '0' && env[0] <= '9') ? env[0] - '0' : -1;
    }();
    return mode;
}

inline int make_kernel_params(const agatha_params_t* p, KernelParams* kp)
{
    if (!p) return set_error(AGATHA_EINVAL, "params is NULL");
    if (p->band_width < 0) return set_error(AGATHA_EINVAL, "band_width < 0");
    if (p->slice_width < 1) return set_error(AGATHA_EINVAL, "slice_width < 1");
    const int C = shape_for(p->band_width).C;
    if (!C) return set_error(AGATHA_EUNSUPPORTED, "band_width %d > %d not supported by this build", p->band_width, agatha_max_band_width());
    kp->match = p->match; kp->mismatch = p->mismatch;
    kp->goe = p->gap_open + p->gap_extend;          // gasal_align.cu:301
    kp->ge = p->gap_extend;
    kp->sw = p->slice_width; kp->Z = p->z_threshold; kp->W = p->band_width;
    kp->LW = p->band_width / C; kp->JW = p->band_width % C;
    // PRMT table over x = query code ^ target code: 0 match, 1..3 mismatch, 4..7 N vs base (-N_PENALTY = -1);
    // x >= 8 selects the sign of entry x&7 replicated: 0xff == -1 for every negative entry (extend_kernel.cuh)
    const unsigned m = (unsigned)p->match & 0xffu, x = (unsigned)(-p->mismatch) & 0xffu;
    kp->tab_lo = m | (x << 8) | (x << 16) | (x << 24);
    kp->tab_hi = 0xffffffffu;
    kp->one = 1; kp->k32 = 32; kp->m16 = 0xffff; kp->k65536 = 65536;
    kp->force_generic = fast_table_ok(p) ? 0 : 1;
    // 16-bit packed steady state (extend_kernel.cuh run_fast16): needs small scoring values so that the per-window drift
    // bounds of its range monitor hold
    const int mode = s16_mode();
    kp->s16 = (!kp->force_generic && p->match >= 0 && p->match <= 100 && p->mismatch <= 100 && p->gap_open >= 0 && p->gap_extend >= 0 &&
               p->gap_open + 2 * p->gap_extend <= 2000 && mode != 0) ? 1 : 0;
    // bit 1: the prologue (anti-diagonals 0..W) may run packed too, without a range monitor. On those anti-diagonals every
    // live value lies in [-(2*goe + ge*(W+1)) - mismatch*(W+2)/2 - goe, match*(W+2)/2] (a cell is at most (W+2)/2 diagonal
    // steps away from a matrix-edge value) and a dead cell creeps up by at most match*(W+2)/2 from the floor (-30000):
    // both must stay well apart and inside 16 bits.
    if (kp->s16) {
        const long long half = (p->band_width + 2) / 2;
        const long long depth = 3LL * kp->goe + (long long)kp->ge * (p->band_width + 1) + (long long)(p->mismatch + p->match) * half + 256;
        if (depth < 24000 && mode != 1) kp->s16 |= 2;
        // bit 2: the tail (far matrix edges) packed as well (it keeps the range monitor, so it needs no bound of its own).
        // Opt-in (AGATHA_S16=7) and only present in builds with -DAGATHA_TAIL16=1: bit-exact and 12-14 % faster on equal-length
        // pairs, but on mixed-length batches the extra loop costs more in instruction fetch than it saves in issue slots
        // (C1: 21.2 -> 27.0 ms, ncu: no_instruction 1.6 -> 2.9 warps per issue).
        if (mode == 7) kp->s16 |= 4;
    }
    // The packed kernel (extend16_kernel.cuh, drifting representation). Table of biased scores s + X, X = mismatch: match + X,
    // 0, X - 1 (N_PENALTY = 1, gasal_kernels.h:48-50) -- all must fit a non-negative byte. bias16 places the prologue inside
    // the 16-bit window: on its anti-diagonals a live true value is at least -lowtrue (two gaps from the origin plus the
    // mismatches of at most half a band), a dead cell creeps up from the floor by at most `creep`, a live value rises to at
    // most (match + X) * half in stored units.
    kp->p16_ok = 0; kp->bias16 = 0; kp->tabb_lo = kp->tabb_hi = 0;
    if ((kp->s16 & 1) && p->mismatch >= 1 && p->match >= 0 && p->match + p->mismatch <= 127 && (p->band_width & 7) == 7) {
        const long long X = p->mismatch, half = (p->band_width + 2) / 2;
        const long long creep = (p->match + X) * half;
        const long long lowtrue = 3LL * kp->goe + (long long)kp->ge * (p->band_width + 2) + X * half + 128;
        const long long bias = 4000 /* FLOORU16 */ + creep + lowtrue + kp->goe + 2LL * kp->ge + RANGE16_PAIRS * (p->match + X) + 320;
        const bool fits = bias + creep + 2048 < 65535 - RANGE16_PAIRS * (p->match + X);
        // (MINUS_INF2 is not read during the prologue -- the band edges are still outside the matrix -- and the range check
        // that follows it tests the lowest live value against the sentinel before the steady state uses it)
        if (fits) {
            kp->p16_ok = 1; kp->bias16 = (int)bias;
            const unsigned m = (unsigned)(p->match + X), n = (unsigned)(X - 1);
            kp->tabb_lo = m;                                    // x = 0 match, 1..3 mismatch (0)
            kp->tabb_hi = n | (n << 8) | (n << 16) | (n << 24);  // x = 4..7: a base against N
        }
    }
    return AGATHA_OK;
}

// Calls l.run<C, NW, WODD, JWS>() for the variant that handles kp and stores its return value in *rc; false when no
// compiled shape covers the band width. The index of the band-edge cell inside its lane, JW = W % C, is a template constant
// for every band width that is 7 (mod 8) -- the only residue for which the reference's band is exact -- and a
// run-time value otherwise. The static-JW variants also compile the prologue with compile-time injection positions
// (STATIC_PRO in extend_kernel.cuh), which hold only for W = 7 (mod 8): for C = 4 the test JW == C-1 alone would also accept
// W = 3 (mod 8) (W = 67, 75, ...), hence the explicit residue test.
template <int C, int NW, class L>
inline int dispatch_c(const KernelParams& kp, const L& l)
{
    const bool wodd = kp.W & 1;
    if (wodd && (kp.W & 7) == 7) {
        constexpr int J0 = C < 8 ? C - 1 : 7;
        if (kp.JW == J0) return l.template run<C, NW, true, J0>();
        if constexpr (C > 15) { if (kp.JW == 15) return l.template run<C, NW, true, 15>(); }
        if constexpr (C > 23) { if (kp.JW == 23) return l.template run<C, NW, true, 23>(); }
        if constexpr (C > 31) { if (kp.JW == 31) return l.template run<C, NW, true, 31>(); }
    }
    if (wodd) return l.template run<C, NW, true, -1>();
    return l.template run<C, NW, false, -1>();
}

template <class L>
inline bool dispatch_variant(const KernelParams& kp, const L& l, int* rc)
{
    const Shape sh = shape_for(kp.W);
    switch (sh.C * 100 + sh.NW) {
        case 201: *rc = dispatch_c<2, 1>(kp, l); return true;
        case 401: *rc = dispatch_c<4, 1>(kp, l); return true;
        case 801: *rc = dispatch_c<8, 1>(kp, l); return true;
        case 1601: *rc = dispatch_c<16, 1>(kp, l); return true;
        case 2401: *rc = dispatch_c<24, 1>(kp, l); return true;
        case 3201: *rc = dispatch_c<32, 1>(kp, l); return true;
        case 3202: *rc = dispatch_c<32, 2>(kp, l); return true;
        case 3204: *rc = dispatch_c<32, 4>(kp, l); return true;
        case 3208: *rc = dispatch_c<32, 8>(kp, l); return true;
    }
    return false;
}

// The packed kernel (extend16_kernel.cuh) covers the band widths 7 (mod 8) from 135 up whose prologue provably stays inside
// 16 bits (kp.s16 bit 1), with table scoring. It runs first; the general kernel then redoes the pairs it marked.
inline bool packed_kernel_ok(const KernelParams& kp)
{
    const Shape sh = shape_for(kp.W);
    return sh.C >= 8 && kp.p16_ok && !kp.force_generic && s16_mode() < 0;
}

template <int C, int NW, class L>
inline int dispatch16_c(const KernelParams& kp, const L& l)
{
    if (kp.JW == 7) return l.template run16<C, NW, 7>();
    if constexpr (C > 15) { if (kp.JW == 15) return l.template run16<C, NW, 15>(); }
    if constexpr (C > 23) { if (kp.JW == 23) return l.template run16<C, NW, 23>(); }
    if constexpr (C > 31) { if (kp.JW == 31) return l.template run16<C, NW, 31>(); }
    return set_error(AGATHA_EUNSUPPORTED, "no packed kernel for band_width %d", kp.W);
}

template <class L>
inline bool dispatch16_variant(const KernelParams& kp, const L& l, int* rc)
{
    if (!packed_kernel_ok(kp)) return false;
    const Shape sh = shape_for(kp.W);
    switch (sh.C * 100 + sh.NW) {
        case 801: *rc = dispatch16_c<8, 1>(kp, l); return true;
        case 1601: *rc = dispatch16_c<16, 1>(kp, l); return true;
        case 2401: *rc = dispatch16_c<24, 1>(kp, l); return true;
        case 3201: *rc = dispatch16_c<32, 1>(kp, l); return true;
        case 3202: *rc = dispatch16_c<32, 2>(kp, l); return true;
        case 3204: *rc = dispatch16_c<32, 4>(kp, l); return true;
        case 3208: *rc = dispatch16_c<32, 8>(kp, l); return true;
    }
    return false;
}

}  // namespace agatha
