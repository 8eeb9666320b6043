// Sonar-fan march variants, single warp, 5 rays from one origin (limit 31: kA = 31, kB = 32), runtime plan.
#include <cstdio>
#include <cuda_runtime.h>
#include "chunked_march.cuh"
using namespace r2p2;
#define N 512

template <int TL, int NOWN>
__device__ __forceinline__ Hit team_march_x(const StageView& s, const Team& tm, const MarchPlan& mp, double ox, double oy,
                                            double vx, double vy, double limit) {
    const int kA = mp.kA, kB = mp.kB;
    if (kB == 0x7fffffff) return march_plain(s, ox, oy, vx, vy, limit);
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = 0u; r.fault = false;
    const int j = tm.j;
    constexpr unsigned kNone = 0xffffffffu;
    constexpr unsigned kSampleSlow = 3u;
    unsigned key = kNone;
    double x = ox, y = oy, hx = -1.0, hy = -1.0;
    const int nwalk = min(TL, kB) - 1;
#pragma unroll 4
    for (int i = 0; i < nwalk; ++i) {
        if (i < j) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
    }
    for (int k0 = j; k0 < kB; k0 += TL * NOWN) {
        double xs[NOWN], ys[NOWN];
        unsigned w[NOWN];
        int ixs[NOWN];
        bool inter[NOWN];
#pragma unroll
        for (int m = 0; m < NOWN; ++m) {
            xs[m] = x; ys[m] = y;
            const int ix = dtrunc(x), iy = dtrunc(y);
            inter[m] = ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
            w[m] = s.wall[inter[m] ? (iy * s.wpr + (ix >> 5)) : 0];
            ixs[m] = ix;
            if (k0 + (m + 1) * TL < kB) {
#pragma unroll
                for (int u = 0; u < TL; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
            }
        }
#pragma unroll
        for (int m = 0; m < NOWN; ++m) {
            const int k = k0 + m * TL;
            bool in_range = true;
            if (k >= kA) in_range = r2p2_norm2sq(R2P2_SUB(xs[m], ox), R2P2_SUB(ys[m], oy)) < mp.qstar;
            const unsigned kind = !inter[m] ? kSampleSlow : (!in_range ? (unsigned)kSampleExit : (((w[m] >> (ixs[m] & 31)) & 1u) ? (unsigned)kSampleHit : 0u));
            if ((k < kB) && kind != 0u && key == kNone) { key = ((unsigned)k << 2) | kind; hx = xs[m]; hy = ys[m]; }
        }
    }
    unsigned best = key;
    if (TL == 32) best = __reduce_min_sync(tm.sync, key);
    else {
#pragma unroll
        for (int o = TL >> 1; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(tm.sync, best, o));
    }
    if (TL > 1) {
        const int lane = (int)(threadIdx.x & 31u);
        const int src = (best != kNone) ? (lane - j + (int)((best >> 2) & (unsigned)(TL - 1))) : lane;
        hx = __shfl_sync(tm.sync, hx, src);
        hy = __shfl_sync(tm.sync, hy, src);
    }
    if (best == kNone) { r.nsamp = (unsigned)kB; return r; }
    const int k = (int)(best >> 2);
    const unsigned kind = best & 3u;
    if (kind == kSampleSlow) return march_plain(s, ox, oy, vx, vy, limit, k, hx, hy);
    if (kind == kSampleHit) { r.x = hx; r.y = hy; r.k = k; r.nsamp = (unsigned)(k + 1); }
    else r.nsamp = (unsigned)k;
    return r;
}

// floor(v) for 0 <= v < 2^32 without F2I: the low word of RD(v + 2^52); ok is false outside that range (and for NaN)
__device__ __forceinline__ bool fast_floor_x(double v, int& idx) {
    const double t = __dadd_rd(v, 4503599627370496.0);
    idx = __double2loint(t);
    return __double2hiint(t) == 0x43300000;
}
template <int TL, int NOWN>
__device__ __forceinline__ Hit team_march_f(const StageView& s, const Team& tm, const MarchPlan& mp, double ox, double oy,
                                            double vx, double vy, double limit) {
    const int kA = mp.kA, kB = mp.kB;
    if (kB == 0x7fffffff) return march_plain(s, ox, oy, vx, vy, limit);
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = 0u; r.fault = false;
    const int j = tm.j;
    constexpr unsigned kNone = 0xffffffffu;
    constexpr unsigned kSampleSlow = 3u;
    unsigned key = kNone;
    double x = ox, y = oy, hx = -1.0, hy = -1.0;
    // to this lane's first sample: unpredicated additions of v or 0.0 (a predicated DADD chain runs at half speed)
    const int nwalk = min(TL, kB) - 1;
#pragma unroll 4
    for (int i = 0; i < nwalk; ++i) {
        const bool on = i < j;
        x = R2P2_ADD(x, on ? vx : 0.0); y = R2P2_ADD(y, on ? vy : 0.0);
    }
    const bool has_window = kB > kA;                       // uniform
    double xw = 0.0, yw = 0.0;
    for (int k0 = j; k0 < kB; k0 += TL * NOWN) {
        double xs[NOWN], ys[NOWN];
        unsigned w[NOWN];
        int ixs[NOWN];
        bool inter[NOWN];
#pragma unroll
        for (int m = 0; m < NOWN; ++m) {
            xs[m] = x; ys[m] = y;
            int ix, iy;
            const bool okx = fast_floor_x(x, ix), oky = fast_floor_x(y, iy);
            inter[m] = okx & oky & ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
            w[m] = __ldg(s.wall + (inter[m] ? (iy * s.wpr + (ix >> 5)) : 0));
            ixs[m] = ix;
            if (m + 1 < NOWN) {
#pragma unroll
                for (int u = 0; u < TL; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
            }
        }
        if (has_window) {
#pragma unroll
            for (int m = 0; m < NOWN; ++m) if (k0 + m * TL == kA) { xw = xs[m]; yw = ys[m]; }
        }
#pragma unroll
        for (int m = 0; m < NOWN; ++m) {
            const int k = k0 + m * TL;
            const unsigned kind = !inter[m] ? kSampleSlow : ((((w[m] >> (ixs[m] & 31)) & 1u) ? (unsigned)kSampleHit : 0u));
            if ((k < kB) && kind != 0u && key == kNone) { key = ((unsigned)k << 2) | kind; hx = xs[m]; hy = ys[m]; }
        }
        if (k0 + TL * NOWN < kB) {                          // another pass follows
#pragma unroll
            for (int u = 0; u < TL; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
        }
    }
    // the knife-edge sample (at most one, k = kA = kB - 1): the reference's limit test comes before its pixel test
    if (has_window && ((kA & (TL - 1)) == j)) {
        const bool in_range = r2p2_norm2sq(R2P2_SUB(xw, ox), R2P2_SUB(yw, oy)) < mp.qstar;
        const bool earlier = (key != kNone) && ((int)(key >> 2) < kA || (key & 3u) == kSampleSlow);
        if (!in_range && !earlier) { key = ((unsigned)kA << 2) | (unsigned)kSampleExit; hx = xw; hy = yw; }
    }
    unsigned best = key;
    if (TL == 32) best = __reduce_min_sync(tm.sync, key);
    else {
#pragma unroll
        for (int o = TL >> 1; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(tm.sync, best, o));
    }
    if (TL > 1) {
        const int lane = (int)(threadIdx.x & 31u);
        const int src = (best != kNone) ? (lane - j + (int)((best >> 2) & (unsigned)(TL - 1))) : lane;
        hx = __shfl_sync(tm.sync, hx, src);
        hy = __shfl_sync(tm.sync, hy, src);
    }
    if (best == kNone) { r.nsamp = (unsigned)kB; return r; }
    const int k = (int)(best >> 2);
    const unsigned kind = best & 3u;
    if (kind == kSampleSlow) return march_plain(s, ox, oy, vx, vy, limit, k, hx, hy);
    if (kind == kSampleHit) { r.x = hx; r.y = hy; r.k = k; r.nsamp = (unsigned)(k + 1); }
    else r.nsamp = (unsigned)k;
    return r;
}

__global__ void lat(double* out, StageView sv, long long* cyc, int mode, double limit, double ox0, double oy0) {
    const int lane = threadIdx.x & 31;
    Team tm; MarchPlan mp;
    int ray;
    if (mode == 0) { tm.tl = 6; tm.j = lane % 6; tm.mask = 0x3fu << (lane - tm.j); tm.sync = 0xffffffffu; mp = make_plan(limit, 6); ray = lane / 6; }
    else if (mode == 5 || mode == 4) { tm.tl = 32; tm.j = lane; tm.mask = 0xffffffffu; tm.sync = 0xffffffffu; mp = make_plan(limit, 32); ray = 0; }
    else { tm.tl = 4; tm.j = lane & 3; tm.mask = 0xfu << (lane - tm.j); tm.sync = 0xffffffffu; mp = make_plan(limit, 4); ray = lane / 4; }
    const double ang = 0.3 + 0.9 * ray;
    const double vx = cos(ang), vy = sin(ang);
    double ox = ox0, oy = oy0;
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        Hit h;
        if (mode == 0) h = team_march_chunked(sv, tm, mp, ox, oy, vx, vy, limit);
        else if (mode == 1) h = team_march_x<4, 8>(sv, tm, mp, ox, oy, vx, vy, limit);
        else if (mode == 2) h = team_march_x<4, 8>(sv, tm, mp, ox, oy, vx, vy, limit);
        else if (mode == 3) h = team_march_f<4, 8>(sv, tm, mp, ox, oy, vx, vy, limit);
        else if (mode == 4) h = team_march_f<32, 1>(sv, tm, mp, ox, oy, vx, vy, limit);
        else h = team_march_x<32, 1>(sv, tm, mp, ox, oy, vx, vy, limit);
        acc += h.nsamp + (unsigned)h.k;
        ox = ox0 + (double)(h.nsamp & 1u);
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = ox + acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    const int W = 460, H = 321, wpr = (W + 31) / 32;
    uint32_t* wall; uint8_t* dist; double* out; long long* cyc;
    uint32_t* hw = (uint32_t*)calloc(wpr * H, 4);
    for (int y = 0; y < H; ++y) for (int x = 250; x < 270; ++x) hw[y * wpr + (x >> 5)] |= 1u << (x & 31);
    cudaMalloc(&wall, wpr * H * 4); cudaMemcpy(wall, hw, wpr * H * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dist, W * H); cudaMemset(dist, 3, W * H);
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    StageView sv; sv.wall = wall; sv.lvl50 = wall; sv.dist = dist; sv.W = W; sv.H = H; sv.wpr = wpr;
    long long c;
    const char* names[] = {"chunked tl=6", "strided runtime tl=4", "strided <4,8>", "fast <4,8>", "fast collision <32,1> 7.3", "collision <32,1> limit 7.3"};
    for (int wallcase = 0; wallcase < 2; ++wallcase)
        for (int mode = 0; mode < 6; ++mode) {
            
            const double limit = (mode == 5 || mode == 4) ? 7.3 : 31.0;
            const double ox = wallcase ? ((mode == 5 || mode == 4) ? 246.0 : 230.0) : 100.0;
            for (int r = 0; r < 2; ++r) { lat<<<1, 32>>>(out, sv, cyc, mode, limit, ox, 160.0); cudaDeviceSynchronize(); }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-30s %s %.0f cycles/call\n", names[mode], wallcase ? "walls " : "open  ", (double)c / N);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
