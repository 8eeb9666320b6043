// Latency of the fixed-shape MLP forward (NetNeuro, LPR=32) and of a team march, single warp.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../r2p2_b200/csrc/r2p2_kernels.cuh"
using namespace r2p2;
#define N 1024
__global__ void mlp_lat(double* out, const double* gen, long long* cyc, int act) {
    FixedMlp<32, NetNeuro> mlp;
    const int sub = threadIdx.x & 31;
    mlp.load(gen, sub);
    double v = 0.3, w = 0.1;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        double h[1];
        h[0] = (sub < 5) ? 25.0 / (20.0 + v + sub) : 0.0;
        mlp.forward(0xffffffffu, sub, h, act, v, w);
    }
    long long t1 = clock64();
    out[threadIdx.x] = v + w;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *gen, *out; long long* cyc;
    cudaMalloc(&gen, 156 * 8); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    double h[156]; for (int i = 0; i < 156; ++i) h[i] = ((i * 37) % 100) / 50.0 - 1.0; cudaMemcpy(gen, h, sizeof h, cudaMemcpyHostToDevice);
    long long c;
    for (int act = 0; act < 2; ++act) {
        for (int r = 0; r < 2; ++r) { mlp_lat<<<1, 32>>>(out, gen, cyc, act); cudaDeviceSynchronize(); }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FixedMlp<32,NetNeuro>::forward act=%d: %.0f cycles/call (incl. 1 DDIV)\n", act, (double)c / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
