// FP64 issue-rate probe (B200): how many DMUL/DADD/DFMA warp-instructions per clock and SM does the FP64 pipe sustain?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_rate tools/micro/fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: 13 DMUL + 13 DADD per step (k_lpc4's pattern), 1: 13 DFMA per step, 2: 26 DADD per step
__global__ void __launch_bounds__(128) probe(double* out, int iters, double seed, double nz)
{
    __shared__ double sm[128][32];
    for (int i = threadIdx.x; i < 128 * 32; i += 128) sm[i / 32][i % 32] = seed + i;
    __syncthreads();
    double acc[13], win[13];
#pragma unroll
    for (int i = 0; i < 13; i++) { acc[i] = -0.0; win[i] = seed + i + threadIdx.x; }
    double v = seed * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 7) {   // DADD with two fresh register operands (no .reuse possible)
#pragma unroll
                for (int i = 0; i < 13; i++) { acc[i] = __dadd_rn(acc[i], win[i]); win[i] = __dadd_rn(win[i], acc[(i + 5) % 13]); }
            } else if (MODE == 8) {   // k_lpc4's pair with the products formed a step ahead: DMUL(w, v.reuse) / DADD(acc, p) on independent data
                v = __longlong_as_double(0x3ff0000000000000ll + (long long)(u + 1));
                double p[13];
#pragma unroll
                for (int i = 0; i < 13; i++) p[i] = __dmul_rn(win[i], v);
#pragma unroll
                for (int i = 0; i < 13; i++) acc[i] = __dadd_rn(acc[i], p[i]);
#pragma unroll
                for (int i = 0; i < 12; i++) win[i] = win[i + 1];
                win[12] = v;
            } else if (MODE == 5) {   // DMUL only
                v = __longlong_as_double(0x3ff0000000000000ll + (long long)(u + 1));
#pragma unroll
                for (int i = 0; i < 13; i++) { acc[i] = __dmul_rn(acc[i], v); win[i] = __dmul_rn(win[i], v); }
            } else if (MODE == 6) {   // groups of 13 DMUL and 13 DADD on independent registers (no mul -> add dependency)
                v = __longlong_as_double(0x3ff0000000000000ll + (long long)(u + 1));
#pragma unroll
                for (int i = 0; i < 13; i++) win[i] = __dmul_rn(win[i], v);
#pragma unroll
                for (int i = 0; i < 13; i++) acc[i] = __dadd_rn(acc[i], v);
            } else if (MODE == 4) {   // the product as DFMA(a, b, -0.0): the same rounding as DMUL
                v = __longlong_as_double(0x3ff0000000000000ll + (long long)(it * 8 + u));
                win[(it * 8 + u) % 13 == 0 ? 0 : 1] = v;
#pragma unroll
                for (int i = 0; i < 13; i++) acc[i] = __dadd_rn(acc[i], __fma_rn(win[i], v, nz));
            } else if (MODE == 0 || MODE == 3) {
                // a fresh v per step (its bits come from the integer pipe, or from shared memory: MODE 3), so the products are not loop-invariant
                if (MODE == 3) v = sm[(it * 8 + u) & 127][threadIdx.x & 31];
                else v = __longlong_as_double(0x3ff0000000000000ll + (long long)(it * 8 + u));
                win[(it * 8 + u) % 13 == 0 ? 0 : 1] = v;
#pragma unroll
                for (int i = 0; i < 13; i++) acc[i] = __dadd_rn(acc[i], __dmul_rn(win[i], v));
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 13; i++) acc[i] = fma(win[i], v, acc[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 13; i++) { acc[i] = __dadd_rn(acc[i], v); win[i] = __dadd_rn(win[i], v); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) s += acc[i] + win[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int sms, double clk_ghz, int fp64_per_step)
{
    double* out;
    const int grid = sms * ctas_per_sm, iters = 4000;
    cudaMalloc(&out, (size_t)grid * 128 * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<grid, 128>>>(out, 100, 1.0, -0.0);
    cudaEventRecord(a);
    probe<MODE><<<grid, 128>>>(out, iters, 1.0, -0.0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double inst = (double)grid * 4 * iters * 8 * fp64_per_step;   // warp instructions
    printf("%-28s warps/SM %2d  %.3f ms  %.3f FP64 warp-inst/clk/SM (at %.3f GHz)\n", name, ctas_per_sm * 4, ms, inst / sms / (ms * 1e-3) / (clk_ghz * 1e9), clk_ghz);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    for (int c : {1, 2, 4, 8}) {
        run<0>("13 DMUL + 13 DADD", c, p.multiProcessorCount, ghz, 26);
        run<1>("13 DFMA", c, p.multiProcessorCount, ghz, 13);
        run<2>("26 DADD", c, p.multiProcessorCount, ghz, 26);
        run<3>("13 DMUL + 13 DADD + LDS.64", c, p.multiProcessorCount, ghz, 26);
        run<4>("13 DFMA(a,b,-0) + 13 DADD", c, p.multiProcessorCount, ghz, 26);
        run<5>("26 DMUL", c, p.multiProcessorCount, ghz, 26);
        run<7>("26 DADD two fresh operands", c, p.multiProcessorCount, ghz, 26);
        run<8>("13 DMUL(w,v) + 13 DADD(acc,p)", c, p.multiProcessorCount, ghz, 26);
        run<6>("13 DMUL, 13 DADD independent", c, p.multiProcessorCount, ghz, 26);
    }
    return 0;
}
