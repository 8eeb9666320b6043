// FP64 issue/latency structure for ONE warp: n independent DADD chains interleaved; F2I.F64; DADD.RD trick; predicated.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void lat(double* out, const double* in, long long* cyc) {
    double a = in[0], b = in[1], c = in[2], d = in[3], v = in[4];
    int acc = 0;
    const bool pr = in[5] > 0.0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (OP == 0) { a = __dadd_rn(a, v); }
        if (OP == 1) { a = __dadd_rn(a, v); b = __dadd_rn(b, v); }
        if (OP == 2) { a = __dadd_rn(a, v); b = __dadd_rn(b, v); c = __dadd_rn(c, v); d = __dadd_rn(d, v); }
        if (OP == 3) { if (pr) { a = __dadd_rn(a, v); b = __dadd_rn(b, v); } }
        if (OP == 4) { a = __dadd_rn(a, v); b = __dadd_rn(b, v); acc += __double2int_rz(a) ^ __double2int_rz(b); }
        if (OP == 5) { a = __dadd_rn(a, v); b = __dadd_rn(b, v); acc += __double2loint(__dadd_rd(a, 4503599627370496.0)) ^ __double2loint(__dadd_rd(b, 4503599627370496.0)); }
        if (OP == 6) { acc += __double2int_rz(a + (double)i); }
        if (OP == 7) { a = __dadd_rn(a, v); acc += (int)(__double_as_longlong(a) >> 40); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + b + c + d + acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *in, *out; long long* cyc;
    cudaMalloc(&in, 64); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    double h[6] = {100.25, 200.5, 300.75, 400.125, 0.7071, 1.0}; cudaMemcpy(in, h, 48, cudaMemcpyHostToDevice);
    const char* names[] = {"1 DADD chain", "2 chains", "4 chains", "2 chains predicated", "2 chains + 2 F2I.F64", "2 chains + 2 DADD.RD floor", "I2D+DADD+F2I independent", "1 chain + int use"};
    long long c;
#define RUN(op) for (int r = 0; r < 2; ++r) { lat<op><<<1, 32>>>(out, in, cyc); cudaDeviceSynchronize(); } cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-30s %.1f cycles/iter\n", names[op], (double)c / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    return 0;
}
