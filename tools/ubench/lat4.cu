// Latency of team_march (5 sonar rays x 6 lanes, limit 30) and of the 32-lane collision-ray march, single warp.
#include <cstdio>
#include <cuda_runtime.h>
#include "chunked_march.cuh"
using namespace r2p2;
#define N 512
__global__ void march_lat(double* out, StageView sv, long long* cyc, int mode) {
    const int lane = threadIdx.x & 31;
    Team tm; MarchPlan mp;
    double limit;
    if (mode == 0) { tm.tl = 6; tm.j = lane % 6; tm.mask = 0x3fu << (lane - tm.j); tm.sync = 0xffffffffu; limit = 30.0; mp = make_plan(limit, 6); }
    else { tm.tl = 32; tm.j = lane; tm.mask = 0xffffffffu; tm.sync = 0xffffffffu; limit = 7.3; mp = make_plan(limit, 32); }
    const double ang = 0.3 + 0.9 * (lane / 6);
    const double vx = cos(ang), vy = sin(ang);
    double ox = 230.0, oy = 160.0;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        {
            const Hit h = team_march_chunked(sv, tm, mp, ox, oy, vx, vy, limit);
            acc += h.nsamp;
            ox = 230.0 + (double)(h.nsamp & 1u);
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = ox + acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    const int W = 460, H = 321, wpr = (W + 31) / 32;
    uint32_t* wall; uint8_t* dist; double* out; long long* cyc;
    cudaMalloc(&wall, wpr * H * 4); cudaMemset(wall, 0, wpr * H * 4);
    cudaMalloc(&dist, W * H); cudaMemset(dist, 200, W * H);
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    StageView sv; sv.wall = wall; sv.lvl50 = wall; sv.dist = dist; sv.W = W; sv.H = H; sv.wpr = wpr;
    long long c;
    for (int mode = 0; mode < 2; ++mode) {
        for (int r = 0; r < 2; ++r) { march_lat<<<1, 32>>>(out, sv, cyc, mode); cudaDeviceSynchronize(); }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("team_march mode %d: %.0f cycles/call\n", mode, (double)c / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
