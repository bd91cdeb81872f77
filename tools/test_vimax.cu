// Probe: does `B == operand` after __vimax3_s32_relu compile correctly on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const int *in, int *out, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int I = in[3 * t], M = in[3 * t + 1], D = in[3 * t + 2];
    int B = __vimax3_s32_relu(I, M, D);
    int code = (B == M) ? 0 : ((B == I) ? 1 : ((B == D) ? 2 : 3));
    int code2 = (M >= I && M >= D && M >= 0) ? 0 : ((I >= D && I >= 0) ? 1 : ((D >= 0) ? 2 : 3));
    out[3 * t] = B; out[3 * t + 1] = code; out[3 * t + 2] = code2;
}
int main() {
    const int vals[] = {-11, -3, -1, 0, 1, 5};
    int h[3 * 216], n = 0;
    for (int a : vals) for (int b : vals) for (int c : vals) { h[3*n]=a; h[3*n+1]=b; h[3*n+2]=c; n++; }
    int *di, *dout; cudaMalloc(&di, sizeof h); cudaMalloc(&dout, sizeof h);
    cudaMemcpy(di, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1, 256>>>(di, dout, n);
    int o[3 * 216]; cudaMemcpy(o, dout, sizeof o, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad2 = 0;
    for (int t = 0; t < n; t++) {
        int I = h[3*t], M = h[3*t+1], D = h[3*t+2];
        int B = I > M ? I : M; B = B > D ? B : D; B = B > 0 ? B : 0;
        int code = (B == M) ? 0 : ((B == I) ? 1 : ((B == D) ? 2 : 3));
        if (o[3*t] != B || o[3*t+1] != code) { if (bad1 < 8) printf("eq-form wrong: I %d M %d D %d -> B %d code %d (want %d %d)\n", I, M, D, o[3*t], o[3*t+1], B, code); bad1++; }
        if (o[3*t+2] != code) { if (bad2 < 8) printf("order-form wrong: I %d M %d D %d -> %d want %d\n", I, M, D, o[3*t+2], code); bad2++; }
    }
    printf("eq-form bad %d, order-form bad %d of %d\n", bad1, bad2, n);
}
