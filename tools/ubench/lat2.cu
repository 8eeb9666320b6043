// Latency of the r2p2_math.h functions (single warp, dependent chain) on B200.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../r2p2_b200/csrc/r2p2_math.h"
#define N 2048
__device__ double d_tab[R2P2_SINCOSTAB_LEN] = {R2P2_SINCOSTAB_VALUES};
template <int OP> __global__ void lat(double* out, const double* in, long long* cyc) {
    __shared__ double tab[R2P2_SINCOSTAB_LEN];
    for (int i = threadIdx.x; i < R2P2_SINCOSTAB_LEN; i += 32) tab[i] = d_tab[i];
    __syncwarp();
    double a = in[0], b = in[1];
    int ok = 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (OP == 0) a = r2p2_tanh(a) + b;
        if (OP == 1) { double s, c; r2p2_sincos(a, tab, &s, &c, &ok); a = s + c + b; }
        if (OP == 2) a = r2p2_atan2(a, b) + b;
        if (OP == 3) a = r2p2_sin(a, tab, &ok) + b;
        if (OP == 4) a = tanh(a) + b;
        if (OP == 5) { double s, c; sincos(a, &s, &c); a = s + c + b; }
        if (OP == 6) a = atan2(a, b) + b;
        if (OP == 7) a = r2p2_norm2(a, b) + b;
        if (OP == 8) { double s, c; r2p2_sincos(a * 7.0, tab, &s, &c, &ok); a = s + c + b; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + ok;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *in, *out; long long* cyc;
    cudaMalloc(&in, 64); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    double h[2] = {0.7, 0.61}; cudaMemcpy(in, h, 16, cudaMemcpyHostToDevice);
    const char* names[] = {"r2p2_tanh", "r2p2_sincos(small)", "r2p2_atan2", "r2p2_sin", "cuda tanh", "cuda sincos", "cuda atan2", "norm2(sqrt)", "r2p2_sincos(reduced)"};
    long long c;
#define RUN(op) for (int r = 0; r < 2; ++r) { lat<op><<<1, 32>>>(out, in, cyc); cudaDeviceSynchronize(); } cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %.0f cycles/call (incl. 1 DADD)\n", names[op], (double)c / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
