// The first team march of this round (contiguous chunk of samples per lane), kept for the micro-benchmarks that compare
// it with the strided layout the kernels use now (lat4.cu, lat6.cu).  Not part of the library.
#pragma once
#include "../../r2p2_b200/csrc/r2p2_kernels.cuh"
namespace r2p2 {
// Team march: `tl` lanes (tm.mask, this lane is number tm.j) cooperate on ONE ray.  Every lane carries the position
// forward from the origin with the reference's sequential additions (redundant work is free under SIMT), but lane j
// only tests the samples of its own contiguous chunk [j*chunk, (j+1)*chunk); the earliest event of the whole ray --
// hit, loop exit or fault, in the reference's order -- is then picked with one warp-level min-reduction.  A 31-sample
// sonar ray costs ~ (tl-1)/tl * 31 chained additions + 31/tl pixel tests per lane instead of 31 of each, with no
// data-dependent branch inside the loops.  tl == 1 degenerates to the plain lock-step loop.
__device__ __forceinline__ Hit team_march_chunked(const StageView& s, const Team& tm, const MarchPlan& mp, double ox, double oy,
                                          double vx, double vy, double limit) {
    const int kA = mp.kA, kB = mp.kB;
    if (kB == 0x7fffffff) return march_plain(s, ox, oy, vx, vy, limit);     // non-finite / huge limit: uniform in the team
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = 0u; r.fault = false;
    const int tl = tm.tl;
    // this lane's samples below the norm window: [k0, k1) (empty for the lane reserved for the window)
    const bool window_lane = (tl == 1) | ((mp.nfast < tl) & (tm.j == tl - 1));
    const int k0 = (tl == 1) ? 0 : ((tm.j < mp.nfast) ? min(kA, tm.j * mp.chunk) : kA);
    const int k1 = (tl == 1) ? kA : ((tm.j < mp.nfast) ? min(kA, k0 + mp.chunk) : kA);
    double x = ox, y = oy;
    // phase 1: carry the position to this lane's first sample.  Lanes that are already there add 0.0 (exact), so the
    // loop is branch-free and the additions form the only dependent chain.
    const int pre = (tl == 1) ? 0 : ((mp.nfast < tl) ? kA : min((tl - 1) * mp.chunk, kA));
    for (int i = 0; i < pre; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool on = (i + u) < k0;
            x = R2P2_ADD(x, on ? vx : 0.0);
            y = R2P2_ADD(y, on ? vy : 0.0);
        }
    }
    // phase 2: test this lane's samples, four at a time and without a branch: positions first (the only dependent
    // chain), then four independent truncation / address / bit-plane-load / test pipelines.  A sample that is not on
    // an interior pixel (map border, off the map, NaN) makes the whole team fall back to the plain loop below.
    constexpr int kNone = 0x7fffffff;
    int key = kNone;
    double hx = -1.0, hy = -1.0;
    bool slow = false;
    const int n2 = (tl == 1) ? kA : mp.chunk;
    for (int c0 = 0; c0 < n2; c0 += 4) {
        double xs[4], ys[4];
        unsigned word[4];
        int ixs[4];
        bool act[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {           // advance only through this lane's own samples: it must end on sample k1
            act[u] = (k0 + c0 + u) < k1;
            xs[u] = x; ys[u] = y;
            x = R2P2_ADD(x, act[u] ? vx : 0.0);
            y = R2P2_ADD(y, act[u] ? vy : 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ix = dtrunc(xs[u]), iy = dtrunc(ys[u]);
            const bool interior = ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
            slow |= act[u] & !interior;
            ixs[u] = ix;
            word[u] = s.wall[interior ? (iy * s.wpr + (ix >> 5)) : 0];       // always a valid address
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool hit = act[u] & (((word[u] >> (ixs[u] & 31)) & 1u) != 0u);
            if (hit && key == kNone) { key = ((k0 + c0 + u) << 2) | kSampleHit; hx = xs[u]; hy = ys[u]; }
        }
    }
    // the norm window [kA, kB) (at most one sample): the reference's own limit test, sqrt-free through qstar
    for (int k = kA; k < kB; ++k) {
        const int ix = dtrunc(x), iy = dtrunc(y);
        const bool interior = ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
        const unsigned w = s.wall[interior ? (iy * s.wpr + (ix >> 5)) : 0];
        const bool in_range = r2p2_norm2sq(R2P2_SUB(x, ox), R2P2_SUB(y, oy)) < mp.qstar;
        const bool live = window_lane & (key == kNone);
        slow |= live & !interior;
        const int st = !in_range ? kSampleExit : (((w >> (ix & 31)) & 1u) ? kSampleHit : kSampleFree);
        if (live && st != kSampleFree) { key = (k << 2) | st; hx = x; hy = y; }
        x = R2P2_ADD(x, vx);
        y = R2P2_ADD(y, vy);
    }
    // Earliest event of the ray = the event of the lowest lane of the team that has one (lanes own ascending sample
    // ranges).  All collectives are issued with tm.sync, ONE mask value shared by every lane that executes this code
    // together: a warp-level instruction whose lanes carry different member masks is executed once per distinct
    // mask (the compiler emits a MATCH.ANY loop), which serialises the teams.
    int best = key;
    if (tl > 1) {
        const unsigned slow_b = __ballot_sync(tm.sync, slow);
        const unsigned ev_b = __ballot_sync(tm.sync, key != kNone) & tm.mask;
        const int src = ev_b ? (__ffs(ev_b) - 1) : (int)(threadIdx.x & 31);
        best = __shfl_sync(tm.sync, key, src);
        hx = __shfl_sync(tm.sync, hx, src);
        hy = __shfl_sync(tm.sync, hy, src);
        slow = (slow_b & tm.mask) != 0u;
    }
    if (slow) return march_plain(s, ox, oy, vx, vy, limit);
    if (best == kNone) { r.nsamp = (unsigned)kB; return r; }
    const int k = best >> 2, st = best & 3;
    if (st == kSampleHit) {
        r.x = hx; r.y = hy; r.k = k; r.nsamp = (unsigned)(k + 1);
    } else {
        r.nsamp = (unsigned)k;          // the while-condition failed at sample k
    }
    return r;
}

}  // namespace r2p2
