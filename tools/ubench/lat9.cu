// Dependent-chain cost of warp collectives and friends, one warp.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void lat(double* out, const double* in, long long* cyc, const unsigned* tab) {
    double a = in[0], v = in[4];
    unsigned k = (unsigned)in[5] + threadIdx.x;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        if (OP == 0) { k = __reduce_min_sync(0xffffffffu, k) + lane; }
        if (OP == 1) { k = min(k, __shfl_xor_sync(0xffffffffu, k, 1)) + lane; }
        if (OP == 2) { a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31); }
        if (OP == 3) { k = __ballot_sync(0xffffffffu, (k & 1u) != 0u) + lane; }
        if (OP == 4) { k = tab[k & 255u] + lane; }                 // LDG chain
        if (OP == 5) { k = __ldg(tab + (k & 255u)) + lane; }       // LDG.CONSTANT chain
        if (OP == 6) { a = __dadd_rn(a, v); k += (unsigned)__double2loint(__dadd_rd(a, 4503599627370496.0)); k = tab[k & 255u]; a += (double)(k & 1u); }
        if (OP == 7) { k = (k * 2654435761u) >> 3; }
        if (OP == 8) { if (k & 1u) k = k * 3u + 1u; else k = k >> 1; k += lane; }   // divergent branch
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + k;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *in, *out; long long* cyc; unsigned* tab;
    cudaMalloc(&in, 64); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&tab, 1024);
    double h[6] = {100.25, 200.5, 300.75, 400.125, 0.7071, 1.0}; cudaMemcpy(in, h, 48, cudaMemcpyHostToDevice);
    unsigned ht[256]; for (int i = 0; i < 256; ++i) ht[i] = (unsigned)((i * 7 + 3) & 255); cudaMemcpy(tab, ht, 1024, cudaMemcpyHostToDevice);
    const char* names[] = {"REDUX.MIN + add", "SHFL.BFLY + min + add", "SHFL f64 (2 shfl)", "BALLOT + add", "LDG chain", "LDG.CONSTANT chain", "DADD+floor+LDG+I2D chain", "IMAD+SHF chain", "divergent if/else"};
    long long c;
#define RUN(op) for (int r = 0; r < 2; ++r) { lat<op><<<1, 32>>>(out, in, cyc, tab); cudaDeviceSynchronize(); } cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-30s %.1f cycles/iter\n", names[op], (double)c / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
