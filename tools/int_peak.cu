// Measures the INT32 instruction issue rate of this GPU (the roofline denominator for the
// X-drop extension and banded-DP kernels; SURVEY.md section 8d asks for the measured figure).
// Each kernel runs 8 independent dependency chains per thread of one instruction class, with
// enough resident warps to saturate issue.  Prints JSON.
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, int seed) {
    int a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x + c;
    int b = seed * 3 + 1, m = seed | 5;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) a[c] = a[c] + a[(c + 3) & 7] + b;                 // IADD3 (alu pipe)
            if (MODE == 1) a[c] = a[c] * m + b;                          // IMAD  (fma pipe)
            if (MODE == 2) a[c] = (c & 1) ? a[c] * m + b : a[c] + a[(c + 2) & 7] + b;  // 50/50 alu + fma
            if (MODE == 3) a[c] = __vimax3_s32_relu(a[c], b - it, c - a[c]);  // DPX max3.relu (2 VIMNMX)
            if (MODE == 4) a[c] = (a[c] ^ a[(c + 3) & 7]) & (a[(c + 5) & 7] | m);  // LOP3 x2
            if (MODE == 5) a[c] = max(a[c] + b, c);                      // VIADDMNMX / add+max
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c];
    if (s == 0x7fffffff) out[threadIdx.x] = s;
}
template <int MODE>
double run(const char *name, int sms, double ops_per_iter) {
    int *out; cudaMalloc(&out, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, r + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * ITERS * CHAINS * ops_per_iter;
    double gops = ops / (best * 1e-3) / 1e9;
    printf("  \"%s_gops\": %.1f,\n", name, gops);
    cudaFree(out);
    return gops;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f,\n", p.name, p.multiProcessorCount, clk / 1000.0);
    double iadd = run<0>("iadd3", p.multiProcessorCount, 1);
    double imad = run<1>("imad", p.multiProcessorCount, 1);
    double mix = run<2>("mixed_iadd_imad", p.multiProcessorCount, 1);
    double dpx = run<3>("vimax3_relu_x2", p.multiProcessorCount, 2);
    double lop = run<4>("lop3", p.multiProcessorCount, 2);
    double addmax = run<5>("add_max", p.multiProcessorCount, 2);
    double best = iadd;
    if (imad > best) best = imad;
    if (mix > best) best = mix;
    if (lop > best) best = lop;
    if (addmax > best) best = addmax;
    if (dpx > best) best = dpx;
    printf("  \"gops_measured\": %.1f,\n", best);
    printf("  \"per_sm_per_clk_at_max_clock\": %.1f,\n", best * 1e9 / (p.multiProcessorCount * (clk * 1e3)));
    printf("  \"how\": \"tools/int_peak.cu: 8 independent chains/thread, 8 CTAs x 256 threads per SM, best of 5, CUDA events\"\n}\n");
    (void)dpx;
}
