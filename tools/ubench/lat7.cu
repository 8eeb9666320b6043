// Where does a strided <4,8> ray go?  Pieces added one at a time (single warp, dependent iterations).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../r2p2_b200/csrc/r2p2_kernels.cuh"
using namespace r2p2;
#define N 512
template <int STAGE>
__global__ void lat(double* out, StageView s, long long* cyc, double ox0, double oy0, int kB, int kA, double qstar) {
    const int lane = threadIdx.x & 31;
    const int j = lane & 3;
    const double ang = 0.3 + 0.9 * (lane >> 2);
    const double vx = cos(ang), vy = sin(ang);
    double ox = ox0, oy = oy0;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N; ++it) {
        double x = ox, y = oy, hx = -1.0, hy = -1.0;
        unsigned key = 0xffffffffu;
        const int nwalk = min(4, kB) - 1;
#pragma unroll 4
        for (int i = 0; i < nwalk; ++i) if (i < j) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
        for (int k0 = j; k0 < kB; k0 += 32) {
            double xs[8], ys[8]; unsigned w[8]; int ixs[8]; bool inter[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                xs[m] = x; ys[m] = y;
                if (STAGE >= 1) {
                    const int ix = dtrunc(x), iy = dtrunc(y);
                    inter[m] = ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
                    ixs[m] = ix;
                    if (STAGE >= 2) w[m] = s.wall[inter[m] ? (iy * s.wpr + (ix >> 5)) : 0]; else w[m] = (unsigned)iy;
                }
                if (k0 + (m + 1) * 4 < kB) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
                }
            }
            if (STAGE == 0) { hx = xs[7]; hy = ys[3]; }
            if (STAGE >= 1) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int k = k0 + m * 4;
                    bool in_range = true;
                    if (STAGE >= 3) { if (k >= kA) in_range = r2p2_norm2sq(R2P2_SUB(xs[m], ox), R2P2_SUB(ys[m], oy)) < qstar; }
                    const unsigned kind = !inter[m] ? 3u : (!in_range ? 2u : (((w[m] >> (ixs[m] & 31)) & 1u) ? 1u : 0u));
                    if ((k < kB) && kind != 0u && key == 0xffffffffu) { key = ((unsigned)k << 2) | kind; hx = xs[m]; hy = ys[m]; }
                }
            }
        }
        unsigned best = key;
        if (STAGE >= 4) {
#pragma unroll
            for (int o = 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            const int src = (best != 0xffffffffu) ? (lane - j + (int)((best >> 2) & 3u)) : lane;
            hx = __shfl_sync(0xffffffffu, hx, src);
            hy = __shfl_sync(0xffffffffu, hy, src);
        }
        acc += best;
        ox = ox0 + ((hx + hy > 1e9) ? 1.0 : 0.0);
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
    cudaMalloc(&dist, W * H); cudaMemset(dist, 3, W * H);
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    StageView sv; sv.wall = wall; sv.lvl50 = wall; sv.dist = dist; sv.W = W; sv.H = H; sv.wpr = wpr;
    long long c;
    const char* names[] = {"chain + record", "+ trunc/interior", "+ bit-plane loads", "+ knife-edge norm", "+ reduce + shuffles"};
#define RUN(st) for (int r = 0; r < 2; ++r) { lat<st><<<1, 32>>>(out, sv, cyc, 100.0, 160.0, 32, 31, 961.0); cudaDeviceSynchronize(); } cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %.0f cycles/ray\n", names[st], (double)c / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
