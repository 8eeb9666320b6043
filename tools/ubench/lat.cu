// Latency microbenchmarks (single warp, dependent chains) on B200 -- build notes for DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void lat(double* out, const double* in, long long* cyc, const unsigned char* tab) {
    double a = in[0], b = in[1];
    int ia = (int)in[2];
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) a = __dadd_rn(a, b);
        if (OP == 1) a = __fma_rn(a, b, b);
        if (OP == 2) a = __dmul_rn(a, b);
        if (OP == 3) { ia = __double2int_rz(a); a = __dadd_rn(b, __int_as_float(0) + (double)0) ; a = a + (double)(ia & 1); }   // F2I + I2F-ish chain (approx)
        if (OP == 4) { ia = tab[ia]; }                         // LDG u8 dependent chain (L1 hit)
        if (OP == 5) { a = __ddiv_rn(b, a); }
        if (OP == 6) { a = __dsqrt_rn(a) + b; }
        if (OP == 7) { ia = ia * 3 + 1; }                      // IMAD chain
        if (OP == 8) { a = rint(a) + b; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + ia + acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *in, *out; long long* cyc; unsigned char* tab;
    cudaMalloc(&in, 64); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&tab, 256);
    double h[3] = {1.000001, 0.999999, 5.0}; cudaMemcpy(in, h, 24, cudaMemcpyHostToDevice);
    unsigned char ht[256]; for (int i = 0; i < 256; ++i) ht[i] = (unsigned char)((i * 7 + 3) & 255); cudaMemcpy(tab, ht, 256, cudaMemcpyHostToDevice);
    const char* names[] = {"DADD", "DFMA", "DMUL", "F2I+DADD+I2D", "LDG.U8 chain", "DDIV", "DSQRT+DADD", "IMAD", "rint+DADD"};
    long long c;
#define RUN(op) for (int r = 0; r < 2; ++r) { lat<op><<<1, 32>>>(out, in, cyc, tab); cudaDeviceSynchronize(); } cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-14s %.1f cycles/op\n", names[op], (double)c / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
