// Development probe (not product code): throughput of the integer multiply-add pipe (IMAD.WIDE.U32), of the FP64 pipe (DFMA)
// and of both issued together, in lane-operations per clock per SM.  Motivation: DESIGN.md section 7 -- field products
// are bound by IMAD.WIDE; an FP64-assisted limb product only pays if DFMA is fast AND overlaps with IMAD.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu ; run: tools/pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0 = IMAD.WIDE only, 1 = DFMA only, 2 = both in every warp, 3 = even warps IMAD / odd warps DFMA
__global__ void __launch_bounds__(256) k_probe(uint64_t *out, double *outd, int iters, uint32_t seed) {
    uint64_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    uint32_t x = seed * 2654435761u + threadIdx.x, y = x ^ 0x9e3779b9u;
    double d0 = 1.0 + threadIdx.x, d1 = d0 * 1.1, d2 = d0 * 1.2, d3 = d0 * 1.3, d4 = d0 * 1.4, d5 = d0 * 1.5, d6 = d0 * 1.6, d7 = d0 * 1.7;
    const double m = 1.0000001, c = 1e-9;
    const bool int_warp = (MODE == 0) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool fp_warp = (MODE == 1) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int i = 0; i < iters; i++) {
        if (int_warp) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a0 += (uint64_t)x * y; a1 += (uint64_t)x * (y + 1); a2 += (uint64_t)(x + 1) * y; a3 += (uint64_t)(x + 2) * y;
                a4 += (uint64_t)x * (y + 3); a5 += (uint64_t)(x + 4) * y; a6 += (uint64_t)x * (y + 5); a7 += (uint64_t)(x + 6) * y;
                x += (uint32_t)a0;
            }
        }
        if (fp_warp) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                d0 = fma(d0, m, c); d1 = fma(d1, m, c); d2 = fma(d2, m, c); d3 = fma(d3, m, c);
                d4 = fma(d4, m, c); d5 = fma(d5, m, c); d6 = fma(d6, m, c); d7 = fma(d7, m, c);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    outd[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
}

template <int MODE> static void run(const char *name, int sms, double clk_ghz, uint64_t *o, double *od) {
    const int iters = 20000, blocks = sms * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE><<<blocks, threads>>>(o, od, 100, 1);
    cudaEventRecord(e0);
    k_probe<MODE><<<blocks, threads>>>(o, od, iters, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lanes = (double)blocks * threads * iters * 32.0;   // ops of ONE kind per participating lane: 32 per iteration
    const double frac_int = MODE == 3 ? 0.5 : (MODE == 1 ? 0 : 1), frac_fp = MODE == 3 ? 0.5 : (MODE == 0 ? 0 : 1);
    const double clocks = ms * 1e-3 * clk_ghz * 1e9 * sms;
    printf("%-34s %8.3f ms   IMAD.WIDE %6.1f /clk/SM   DFMA %6.1f /clk/SM\n", name, ms, lanes * frac_int / clocks, lanes * frac_fp / clocks);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("%s, %d SMs, %.3f GHz (nominal max; rates below assume it)\n", p.name, p.multiProcessorCount, ghz);
    uint64_t *o; double *od; cudaMalloc(&o, 8 << 20); cudaMalloc(&od, 8 << 20);
    run<0>("IMAD.WIDE only", p.multiProcessorCount, ghz, o, od);
    run<1>("DFMA only", p.multiProcessorCount, ghz, o, od);
    run<2>("both, interleaved in every warp", p.multiProcessorCount, ghz, o, od);
    run<3>("even warps IMAD.WIDE, odd warps DFMA", p.multiProcessorCount, ghz, o, od);
    return 0;
}
