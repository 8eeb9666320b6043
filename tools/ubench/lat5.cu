// Latency of the slow-path pieces of a tick for a robot stuck at a wall (single warp, 32 lanes per robot):
// collision-ray team march with a hit, make_plan, crash ring.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../r2p2_b200/csrc/r2p2_kernels.cuh"
using namespace r2p2;
#define N 512
__global__ void lat(double* out, StageView sv, long long* cyc, int mode, const double* crxy) {
    __shared__ double s_crx[360], s_cry[360];
    for (int i = threadIdx.x; i < 360; i += 32) { s_crx[i] = crxy[i]; s_cry[i] = crxy[360 + i]; }
    __syncwarp();
    const int lane = threadIdx.x & 31;
    Team tm; tm.tl = 32; tm.j = lane; tm.mask = 0xffffffffu; tm.sync = 0xffffffffu;
    Group<32> g; g.sub = lane; g.mask = 0xffffffffu;
    double limit = 7.3;
    const double vx = cos(0.3), vy = sin(0.3);
    double ox = 230.0, oy = 160.0;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (mode == 0) {            // miss, plan hoisted
            const MarchPlan mp = make_plan(7.3, 32);
            const Hit h = team_march_t<32, 1>(sv, tm, mp, ox, oy, vx, vy, 7.3);
            acc += h.nsamp; ox = 230.0 + (double)(h.nsamp & 1u);
        } else if (mode == 1) {     // make_plan in the chain
            const MarchPlan mp = make_plan(limit, 32);
            const Hit h = team_march_t<32, 1>(sv, tm, mp, ox, oy, vx, vy, limit);
            acc += h.nsamp; ox = 230.0 + (double)(h.nsamp & 1u); limit = 7.3 + 1e-9 * (double)(h.nsamp & 3u);
        } else if (mode == 2) {     // hit at k = 3 (wall column at x = 303)
            const MarchPlan mp = make_plan(limit, 32);
            const Hit h = team_march_t<32, 1>(sv, tm, mp, ox + 70.0, oy, vx, vy, limit);
            acc += h.nsamp; ox = 230.0 + (double)(h.k & 1); limit = 7.3 + 1e-9 * (double)(h.nsamp & 3u);
        } else if (mode == 3) {     // crash ring, clear
            const int c = crash_ring<32>(sv, g, ox, oy, 6.0, s_crx, s_cry);
            acc += c; ox = 230.0 + (double)(c & 1);
        } else if (mode == 4) {     // crash ring touching the wall
            const int c = crash_ring<32>(sv, g, ox + 67.5, oy, 6.0, s_crx, s_cry);
            acc += c; ox = 230.0 + (double)(c & 2);
        } else if (mode == 5) {     // ray_is_clear + norm2
            unsigned ns = 0;
            const double cl = R2P2_ADD(r2p2_norm2(ox - 229.0, oy - 159.0), 5.0);
            const bool c = ray_is_clear(sv, ox, oy, cl, ns);
            acc += ns; ox = 230.0 + (double)(c ? (ns & 1u) : 1u);
        }
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = ox + acc + limit;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    const int W = 460, H = 321, wpr = (W + 31) / 32;
    uint32_t* wall; uint8_t* dist; double* out; long long* cyc; double* crxy;
    uint32_t* hw = (uint32_t*)calloc(wpr * H, 4);
    for (int y = 0; y < H; ++y) for (int x = 303; x < 320; ++x) hw[y * wpr + (x >> 5)] |= 1u << (x & 31);
    cudaMalloc(&wall, wpr * H * 4); cudaMemcpy(wall, hw, wpr * H * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dist, W * H); cudaMemset(dist, 3, W * H);
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    double hc[720]; for (int i = 0; i < 360; ++i) { hc[i] = cos(i * M_PI / 180) * 6.0; hc[360 + i] = sin(i * M_PI / 180) * 6.0; }
    cudaMalloc(&crxy, sizeof hc); cudaMemcpy(crxy, hc, sizeof hc, cudaMemcpyHostToDevice);
    StageView sv; sv.wall = wall; sv.lvl50 = wall; sv.dist = dist; sv.W = W; sv.H = H; sv.wpr = wpr;
    long long c;
    const char* names[] = {"coll march miss, plan hoisted", "coll march miss + make_plan", "coll march hit k=3 + make_plan", "crash ring clear", "crash ring wall", "norm2 + ray_is_clear (not clear)"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int r = 0; r < 2; ++r) { lat<<<1, 32>>>(out, sv, cyc, mode, crxy); cudaDeviceSynchronize(); }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-36s %.0f cycles/call\n", names[mode], (double)c / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
