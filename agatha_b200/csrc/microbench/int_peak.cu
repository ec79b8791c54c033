// Integer-pipe throughput microbenchmark for B200 (sm_100a).
// Measures warp-instruction issue rates for the ops the extension kernel is built from, so the
// roofline denominator ("INT32 op peak") is a measured number, not a guess.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 1 << 17;

enum Op { OP_IADD, OP_IMAD, OP_LEA, OP_VIADDMAX, OP_VIMAX3, OP_VIADDMAX16, OP_VIMAX316, OP_DP4A, OP_PRMT, OP_LOP3, OP_SHF,
          OP_MIX_ALU_FMA, OP_MIX_ALU2_FMA1, OP_MIX_CELL, OP_MIX_CELL_PRMT, OP_SHFL, OP_REDUX, OP_IMADHI, OP_IMADSHL, OP_MIX_CELL_HI, OP_COUNT };

template <int OP>
__global__ void __launch_bounds__(1024) bench(int* out, int a0, int b0, int c0, long long* clk)
{
    int v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = a0 + threadIdx.x * (i + 1);
    int b = b0 + (threadIdx.x & 3), c = c0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == OP_IADD) { asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b)); }
            else if (OP == OP_IMAD) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c)); }
            else if (OP == OP_LEA) { v[i] = v[i] * 32 + b; asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_VIADDMAX) { v[i] = __viaddmax_s32(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_VIMAX3) { v[i] = __vimax3_s32(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_VIADDMAX16) { v[i] = __viaddmax_s16x2(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_VIMAX316) { v[i] = __vimax3_s16x2(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_DP4A) { v[i] = __dp4a(b, c, v[i]); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_PRMT) { v[i] = __byte_perm(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c)); }
            else if (OP == OP_SHF) { v[i] = __funnelshift_l(b, v[i], 4); asm volatile("" : "+r"(v[i])); }
            else if (OP == OP_MIX_ALU_FMA) {
                // one ALU-pipe op + one FMA-pipe op per slot (counted as 2 ops)
                if (i & 1) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c)); }
                else { v[i] = __viaddmax_s32(v[i], b, c); asm volatile("" : "+r"(v[i])); }
            }
            else if (OP == OP_MIX_ALU2_FMA1) {
                // two ALU-pipe ops + one FMA-pipe op per slot (counted as 3 ops)
                int t = v[i]; asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(t) : "r"(b), "r"(c));
                int e = __viaddmax_s32(v[i], b, c);
                v[i] = __vimax3_s32(e, t, b); asm volatile("" : "+r"(v[i]));
            }
            else if (OP == OP_MIX_CELL_PRMT) {
                // cell mix with PRMT sign-extending extract + IADD instead of dp4a (8 ops)
                int s = __byte_perm(b, 0, 0x8880 | (i & 3));
                int m = v[i] + s; asm volatile("" : "+r"(m));
                int t = m + c;  asm volatile("" : "+r"(t));
                int h = __vimax3_s32(m, v[(i + 1) % ILP], v[(i + 2) % ILP]);
                int e = __viaddmax_s32(v[(i + 1) % ILP], b, t);
                int f = __viaddmax_s32(v[(i + 2) % ILP], b, t);
                int k = h * 32 + i;
                v[i] = __vimax3_s32(e, f, k);
                asm volatile("" : "+r"(v[i]));
            }
            else if (OP == OP_MIX_CELL) {
                // the per-cell mix of the extension kernel: dp4a, imad(t), lea(key), vimax3, 2x viaddmax, 0.5 vimax3
                int m = __dp4a(b, c, v[i]);
                int t = m + c;  asm volatile("" : "+r"(t));
                int h = __vimax3_s32(m, v[(i + 1) % ILP], v[(i + 2) % ILP]);
                int e = __viaddmax_s32(v[(i + 1) % ILP], b, t);
                int f = __viaddmax_s32(v[(i + 2) % ILP], b, t);
                int k = h * 32 + i;
                v[i] = __vimax3_s32(e, f, k);
                asm volatile("" : "+r"(v[i]));
            }
            else if (OP == OP_IMADHI) { asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c)); }
            else if (OP == OP_IMADSHL) { asm volatile("mul.lo.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b)); }
            else if (OP == OP_MIX_CELL_HI) {
                // cell mix with mad.hi (FMA pipe) instead of dp4a: 1.75 FMA ops for the score add
                int sh = v[i]; asm volatile("mul.lo.s32 %0, %0, %1;" : "+r"(sh) : "r"(b));
                int m = v[(i + 3) % ILP]; asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(m) : "r"(sh), "r"(c));
                int t = m; asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(t) : "r"(b), "r"(c));
                int h = __vimax3_s32(m, v[(i + 1) % ILP], v[(i + 2) % ILP]);
                int e = __viaddmax_s32(v[(i + 1) % ILP], b, t);
                int f = __viaddmax_s32(v[(i + 2) % ILP], b, t);
                int k = h; asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(k) : "r"(c), "r"(b));
                v[i] = __vimax3_s32(e, f, k);
                asm volatile("" : "+r"(v[i]));
            }
            else if (OP == OP_SHFL) { v[i] = __shfl_up_sync(0xffffffffu, v[i], 1); }
            else if (OP == OP_REDUX) { v[i] = __reduce_max_sync(0xffffffffu, v[i]) + b; }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, double ops_per_slot, int nsm, double clk_mhz, bool last)
{
    const int threads = 1024, blocks_per_sm = 1;
    const int blocks = nsm * blocks_per_sm;
    int* out; long long* clk;
    CK(cudaMalloc(&out, sizeof(int) * blocks * threads));
    CK(cudaMalloc(&clk, sizeof(long long) * blocks));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<OP><<<blocks, threads>>>(out, 1, 3, 5, clk);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        bench<OP><<<blocks, threads>>>(out, 1, 3, 5, clk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long hclk[8]; CK(cudaMemcpy(hclk, clk, sizeof(hclk), cudaMemcpyDeviceToHost));
    double thread_ops = (double)ITERS * ILP * ops_per_slot;
    double total = thread_ops * threads * (double)blocks;
    double tops = total / (best * 1e-3) / 1e12;
    // per-SM per-clock lane-ops, from the in-kernel cycle counter of block 0
    double per_sm_clk = thread_ops * threads * blocks_per_sm / (double)hclk[0];
    printf("  \"%s\": {\"ms\": %.4f, \"tera_lane_ops_per_s\": %.3f, \"lane_ops_per_clk_per_sm\": %.2f, \"eff_mhz\": %.0f}%s\n",
           name, best, tops, per_sm_clk, (double)hclk[0] / (best * 1e-3) / 1e6, last ? "" : ",");
    cudaFree(out); cudaFree(clk);
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d,\n", p.name, nsm, khz / 1000);
    run<OP_IADD>("iadd", 1, nsm, khz / 1e3, false);
    run<OP_IMAD>("imad", 1, nsm, khz / 1e3, false);
    run<OP_LEA>("lea_x32", 1, nsm, khz / 1e3, false);
    run<OP_VIADDMAX>("viaddmax_s32", 1, nsm, khz / 1e3, false);
    run<OP_VIMAX3>("vimax3_s32", 1, nsm, khz / 1e3, false);
    run<OP_VIADDMAX16>("viaddmax_s16x2", 1, nsm, khz / 1e3, false);
    run<OP_VIMAX316>("vimax3_s16x2", 1, nsm, khz / 1e3, false);
    run<OP_DP4A>("dp4a", 1, nsm, khz / 1e3, false);
    run<OP_PRMT>("prmt", 1, nsm, khz / 1e3, false);
    run<OP_LOP3>("lop3", 1, nsm, khz / 1e3, false);
    run<OP_SHF>("shf", 1, nsm, khz / 1e3, false);
    run<OP_MIX_ALU_FMA>("mix_viaddmax_imad", 1, nsm, khz / 1e3, false);
    run<OP_MIX_ALU2_FMA1>("mix_alu2_fma1", 3, nsm, khz / 1e3, false);
    run<OP_MIX_CELL>("mix_cell7", 7, nsm, khz / 1e3, false);
    run<OP_MIX_CELL_PRMT>("mix_cell8_prmt", 8, nsm, khz / 1e3, false);
    run<OP_SHFL>("shfl", 1, nsm, khz / 1e3, false);
    run<OP_REDUX>("redux_max", 1, nsm, khz / 1e3, false);
    run<OP_IMADHI>("imad_hi", 1, nsm, khz / 1e3, false);
    run<OP_IMADSHL>("imul_lo", 1, nsm, khz / 1e3, false);
    run<OP_MIX_CELL_HI>("mix_cell8_madhi", 8, nsm, khz / 1e3, true);
    printf("}\n");
    return 0;
}
