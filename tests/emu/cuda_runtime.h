// TEST INFRASTRUCTURE -- host-side SIMT emulation of the few CUDA features the kernels of agatha_b200/csrc use.
//
// This header SHADOWS <cuda_runtime.h> when tests/emu/emu_main.cpp compiles the product's kernel headers
// (extend_kernel.cuh, pack_kernel.cuh) with g++: every CUDA thread of one CTA becomes a fiber (own stack, cooperative
// switch), warp collectives (__shfl*, __ballot, __reduce_*, __any) and __syncthreads are rendezvous points. The kernels
// only call collectives under warp-uniform control flow with the full mask, which is what this emulation supports.
// It exists so that the kernel LOGIC can be checked against the oracle on a machine without a GPU (pytest -m "not gpu");
// nothing under agatha_b200/ includes or links it, and it says nothing about performance.
#pragma once

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define AGATHA_HOST_EMU 1

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static          /* one CTA at a time, one OS thread */

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct emu_dim3 { unsigned x, y, z; };

namespace emu {

struct Warp {
    int arrived = 0;
    unsigned long long gen = 0;
    long long in[32];
    long long in2[32];
    long long out[32];
};

struct Block {
    int arrived = 0;
    unsigned long long gen = 0;
    int nthreads = 0;
};

struct Thread {
    emu_dim3 tid;
    Warp* warp;
    int lane;
};

extern Thread* cur;            // the running fiber's CUDA thread
extern Block blk;
extern emu_dim3 g_blockDim, g_gridDim, g_blockIdx;
void yield();                   // back to the scheduler; returns when it is this fiber's turn again

enum Op { SHFL_IDX, SHFL_UP, SHFL_DOWN, BALLOT, RMAX, RMIN, RMAXU, RMINU };

inline long long warp_collective(Op op, long long v, long long arg)
{
    Warp& w = *cur->warp;
    const int lane = cur->lane;
    w.in[lane] = v; w.in2[lane] = arg;
    const unsigned long long gen = w.gen;
    if (++w.arrived == 32) {
        long long red = 0;
        switch (op) {
            case BALLOT: for (int i = 0; i < 32; i++) if (w.in[i]) red |= 1ll << i; break;
            case RMAX: red = LLONG_MIN; for (int i = 0; i < 32; i++) red = std::max(red, w.in[i]); break;
            case RMIN: red = LLONG_MAX; for (int i = 0; i < 32; i++) red = std::min(red, w.in[i]); break;
            default: break;
        }
        for (int i = 0; i < 32; i++) {
            switch (op) {
                case SHFL_IDX: w.out[i] = w.in[(int)(w.in2[i] & 31)]; break;
                case SHFL_UP: { const int s = i - (int)w.in2[i]; w.out[i] = s >= 0 ? w.in[s] : w.in[i]; } break;
                case SHFL_DOWN: { const int s = i + (int)w.in2[i]; w.out[i] = s < 32 ? w.in[s] : w.in[i]; } break;
                default: w.out[i] = red; break;
            }
        }
        w.arrived = 0;
        w.gen++;
    } else {
        while (w.gen == gen) yield();
    }
    return w.out[lane];
}

inline void block_barrier()
{
    const unsigned long long gen = blk.gen;
    if (++blk.arrived == blk.nthreads) { blk.arrived = 0; blk.gen++; }
    else while (blk.gen == gen) yield();
}

// mbarrier (split arrive / wait): bits 0..15 pending arrivals, 16..31 expected count, bit 32 parity of the current phase
inline void mbar_init(uint64_t* b, unsigned count) { *b = (uint64_t)count | ((uint64_t)count << 16); }
inline void mbar_arrive(uint64_t* b)
{
    uint64_t v = *b;
    unsigned pending = (unsigned)(v & 0xffffu) - 1u;
    const unsigned count = (unsigned)((v >> 16) & 0xffffu);
    uint64_t phase = (v >> 32) & 1u;
    if (pending == 0u) { pending = count; phase ^= 1u; }
    *b = (uint64_t)pending | ((uint64_t)count << 16) | (phase << 32);
}
inline void mbar_wait(uint64_t* b, unsigned parity)
{
    unsigned long long spins = 0;
    while ((unsigned)((*(volatile uint64_t*)b >> 32) & 1u) == (parity & 1u)) {
        if (++spins > 50000000ull) { fprintf(stderr, "emu: mbar_wait never completes (deadlock in the kernel under test)\n"); abort(); }
        yield();
    }
}

inline bool mbar_test(uint64_t* b, unsigned parity) { return (unsigned)((*(volatile uint64_t*)b >> 32) & 1u) != (parity & 1u); }

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_collective(emu::BALLOT, 0, 0); }

inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::warp_collective(emu::SHFL_IDX, v, src); }
inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)emu::warp_collective(emu::SHFL_IDX, v, src); }
inline int __shfl_up_sync(unsigned, int v, unsigned d) { return (int)emu::warp_collective(emu::SHFL_UP, v, d); }
inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned d) { return (unsigned)emu::warp_collective(emu::SHFL_UP, v, d); }
inline int __shfl_down_sync(unsigned, int v, unsigned d) { return (int)emu::warp_collective(emu::SHFL_DOWN, v, d); }
inline unsigned __shfl_down_sync(unsigned, unsigned v, unsigned d) { return (unsigned)emu::warp_collective(emu::SHFL_DOWN, v, d); }
inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)emu::warp_collective(emu::BALLOT, pred != 0, 0); }
inline int __any_sync(unsigned, int pred) { return emu::warp_collective(emu::BALLOT, pred != 0, 0) != 0; }
inline int __all_sync(unsigned, int pred) { return (unsigned)emu::warp_collective(emu::BALLOT, pred != 0, 0) == 0xffffffffu; }
inline int __reduce_max_sync(unsigned, int v) { return (int)emu::warp_collective(emu::RMAX, v, 0); }
inline int __reduce_min_sync(unsigned, int v) { return (int)emu::warp_collective(emu::RMIN, v, 0); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return (unsigned)emu::warp_collective(emu::RMAX, (long long)v, 0); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return (unsigned)emu::warp_collective(emu::RMIN, (long long)v, 0); }
inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    unsigned r = 0;
    for (int b = 0; b < 32; b++) if (emu::warp_collective(emu::BALLOT, (v >> b) & 1u, 0)) r |= 1u << b;
    return r;
}

template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicMax(int* p, int v) { const int o = *p; *p = std::max(o, v); return o; }
inline int atomicMin(int* p, int v) { const int o = *p; *p = std::min(o, v); return o; }

inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned n = (sel >> (4 * i)) & 15u;
        unsigned byte = (unsigned)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;     // sign-replicate mode
        r |= byte << (8 * i);
    }
    return r;
}
inline int __dp4a(int a, int b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
}
inline int __viaddmax_s32(int a, int b, int c) { return std::max((int)((unsigned)a + (unsigned)b), c); }
inline int __viaddmin_s32(int a, int b, int c) { return std::min((int)((unsigned)a + (unsigned)b), c); }
inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
inline unsigned emu_u16x2(unsigned lo, unsigned hi) { return (lo & 0xffffu) | (hi << 16); }
inline unsigned __viaddmax_u16x2(unsigned a, unsigned b, unsigned c)
{
    const unsigned lo = std::max((a + b) & 0xffffu, c & 0xffffu), hi = std::max(((a >> 16) + (b >> 16)) & 0xffffu, c >> 16);
    return emu_u16x2(lo, hi);
}
inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c)
{
    const unsigned lo = std::min((a + b) & 0xffffu, c & 0xffffu), hi = std::min(((a >> 16) + (b >> 16)) & 0xffffu, c >> 16);
    return emu_u16x2(lo, hi);
}
inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu_u16x2(std::max(a & 0xffffu, std::max(b & 0xffffu, c & 0xffffu)), std::max(a >> 16, std::max(b >> 16, c >> 16)));
}
inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c)
{
    return emu_u16x2(std::min(a & 0xffffu, std::min(b & 0xffffu, c & 0xffffu)), std::min(a >> 16, std::min(b >> 16, c >> 16)));
}
inline unsigned __vimax_u16x2(unsigned a, unsigned b) { return __vimax3_u16x2(a, b, b); }
inline unsigned __vimin_u16x2(unsigned a, unsigned b) { return __vimin3_u16x2(a, b, b); }
using std::max;
using std::min;
