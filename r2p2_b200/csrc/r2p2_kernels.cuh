// r2p2_kernels.cuh -- hand-written sm_100a kernels for the r2p2 robot step.
//
// One launch of episode_kernel advances a whole population by up to T ticks of the reference's
// Robot.update (r2p2/robot.py:389-402): sonar ray-march (r2p2/utils.py:420-432 through
// robot.py:287-318), controller forward (neurocontroller.py:67-95 / inspyred.py:60-100), unicycle
// integration with wall / border / crash-radius handling (robot.py:181-260) and the fitness
// bookkeeping (neurocontroller.py:69-85,129-131).  State lives in registers across ticks; HBM is
// touched at robot start (genome, pose) and robot end only.
//
// Mapping: LPR lanes cooperate on one robot (LPR = 1, 8 or 32; a "group" is an aligned slice of a
// warp synchronised with its own member mask).  Sonar rays, MLP neurons and the 360 crash-radius
// probes are spread over the group's lanes; the strictly sequential per-robot scalar work is done
// redundantly by every lane of the group (free under SIMT, keeps values uniform without broadcasts).
// Robots are handed out through a global work counter, so a group that finishes an episode early
// (collision) immediately starts the next robot.
//
// Exactness: every FP64 operation that can reach an integer decision is issued with explicit
// rounding (r2p2_math.h); this translation unit is compiled with -fmad=false.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/r2p2_b200.h"
#include "r2p2_math.h"

namespace r2p2 {

#ifdef R2P2_PHASE_TIMING      // profiling build only (tools/phase_timing.py): cycles per tick phase of robot 0
__device__ unsigned long long g_phase_cycles[16];
__device__ unsigned long long g_robot_cycles[65536 * 12];   // per robot, per phase (last launch)
#define PT_DECL long long pt_last = clock64(); unsigned long long pt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PT(n) { const long long pt_now = clock64(); pt_acc[n] += (unsigned long long)(pt_now - pt_last); pt_last = pt_now; }
#define PT_FLUSH(cond) if (cond) { for (int q = 0; q < 12; ++q) atomicAdd(&g_phase_cycles[q], pt_acc[q]); } \
    if (g.sub == 0 && rb < 65536u) { for (int q = 0; q < 12; ++q) g_robot_cycles[rb * 12 + q] = pt_acc[q]; }
#else
#define PT_DECL
#define PT(n)
#define PT_FLUSH(cond)
#endif

#ifndef R2P2_RING_T
#define R2P2_RING_T 8           // one-lane mapping: above this many rings per converged warp every lane scans its own (measured: 4, 8, 12, 16)
#endif
// Threads per CTA are chosen per launch (128, 256, 512, or one CTA per SM of 32 * ceil(warps / SMs) threads for a
// population that is resident at once; blockDim.x): the 128-register instances hold 16 warps per SM
// whatever the split, and one large CTA stages the read-only tables once per SM instead of four times -- the shared
// memory that frees goes to L1, where the genomes of the robots in flight live (measured: config E 130 -> 122 ms per 200
// ticks, scenario-track 62 -> 57 ms).  The register-weight instances (up to 242 registers) run 128 threads.
#ifndef R2P2_MAXBLOCK
#define R2P2_MAXBLOCK 512
#endif
constexpr int kMaxBlock = R2P2_MAXBLOCK;   // threads per SM of the shared-memory / run-time-shaped instances (registers per thread = 65536 / this)
constexpr int kRegBlock = 128;
constexpr int kCinfoRow = 6;          // doubles per tick and robot in the collision / controller trace
// sin/cos table in shared memory: rows (sn, ssn, cs, ccs) padded to 5 doubles.  Lanes with different angles read
// different rows; at glibc's 4 doubles (8 words) per row they fall on 4 bank groups -- 8-way conflicts on each of the 8
// loads of a sincos (ncu: 57 % of the one-lane kernel's shared-load wavefronts were conflicts, the LSU pipe 66 % busy);
// 10 words per row spread them over 16.
constexpr int kSinCosRow = 5;
constexpr int kSinCosPadLen = (R2P2_SINCOSTAB_LEN / 4) * kSinCosRow;      // 550 doubles = 4400 bytes (a multiple of 16)
constexpr double kRad = 3.14159265358979323846 / 180.0;   // CPython math.radians multiplier
constexpr double kDeg = 180.0 / 3.14159265358979323846;   // CPython math.degrees multiplier
constexpr double kNormSlack = 1e-6;            // |norm(p_k - o) - k| bound used to skip the norm test (see march)

struct KParams {
    // stage
    int32_t W, H, wpr;                 // wpr = 32-bit words per bit-plane row
    int32_t has50;
    const uint32_t* g_wall;            // [H][wpr] bit x of row y set <=> level(x, y) == 0
    const uint32_t* g_lvl50;           //                         <=> level(x, y) == 50
    const uint8_t* g_dist;             // [H][W] Chebyshev distance to the nearest blocked pixel (wall or outer ring), <= 255
    uint32_t plane_bytes;              // bytes of one plane, multiple of 16
    const uint32_t* g_hot;             // padded "event" plane for the ray march (see HotPlane), pre-offset to pixel (0, 0); may be NULL
    int32_t hwpr, hpad;                // its words per row (one row = 4 pixel rows) and padding in pixels
    int32_t win_ww, win_wb, win_half;  // warp-synchronous kernel: per-robot window of the hot plane in shared memory -- words per
                                       // band (multiple of 4), bands, and the half-size in pixels a tick needs around the robot
    int32_t win_slack;                 // extra pixels on every side when a window is fetched: it is kept until the robot has moved that far
    int32_t crash_clear;               // distance from which the crash-radius ring cannot touch a wall: ceil(radius) + 2
    // robot
    int32_t R;
    double ray_deg[R2P2_MAX_RAYS];
    double vr0, vr1, radius, max_speed;
    double L, db_lo, db_hi;            // vr1 + radius, vr0*0.125 + radius, vr0 + radius
    double rad_p1;                     // radius + 1
    // controller
    int32_t ctrl_kind, n_layers, activation, G, maxw;
    int32_t w_in_smem;                 // LPR > 1: copy each robot's genome into shared memory at robot start
    int32_t tanh_in_smem;              // the tanh table (18 KB) is staged in shared memory next to the sin/cos table
    const double* g_tanhtab;           // its global copy
    int32_t layers[R2P2_MAX_LAYERS];
    const double* genomes;             // [P][G] (LPR > 1) or tiles [ceil(P/32)][G][32] (LPR == 1, see transpose_genomes_kernel)
    // run
    int32_t P, T, evolve;
    double delta, delta_ctrl;
    // tables
    const double* g_sincostab;         // 440 doubles (glibc's layout; kernels without a shared-memory copy)
    const double* g_sincostab_pad;     // the same rows padded to kSinCosRow doubles: what the episode kernels stage into shared memory
    const double* g_crxy;              // 360 x offsets then 360 y offsets (cos/sin(radians(i)) * radius)
    // state (SoA, P each)
    double *x, *y, *theta, *speed, *omega, *last_x, *last_y, *dist, *odom_x, *odom_y, *origin_x, *origin_y, *time, *fitness;
    uint32_t *flags, *ncollide, *ticks;
    unsigned long long *nsamp_sonar, *nsamp_coll;
    // sensor noise (robot.py:96,139-140,308-310)
    int32_t noise_mode;                // 0 off, 1 host stream, 2 device generator
    int32_t noise_len;                 // draws per robot in noise_stream
    unsigned long long noise_seed;
    uint32_t robot_index0;             // global index of robot 0 (sharded populations): key of the device noise generator
    const double* noise_stream;        // [P][noise_len], consumed one per unclamped ray in ray order
    uint32_t* ndraws;                  // [P] draws consumed so far
    // trace (optional)
    double* tr_pose; double* tr_dst; int32_t* tr_hitk; int32_t* tr_hitpx; uint32_t* tr_nsamp; uint32_t* tr_ncoll;
    int goal_on;                // goal flagging (r2p2_set_goal): predicate on the pose after each tick, a flag only
    int goal_is_wall;           //   the goal pixel itself is a wall: the predicate always holds (pid_controller.py:210-211)
    double goal_x, goal_y, goal_r;
    uint32_t* goal_tick;        //   [P] ticks completed when the predicate first held, 0xffffffff = never
    int skip_sense;             // r2p2_move: the tick starts at the controller output (command in `genomes`), no sensor sweep
    double* tr_cinfo;           // [T][P][6]: collide() calls of the tick (code, wall spot x, y, border coordinate) and the controller's
                                //            answer before the speed clamp (what its own log records), see r2p2_read_trace_collisions
    // Sequential_PID_Controller (R2P2_CTRL_PID): per-robot controller state, SoA [P]; the goal stack [R2P2_PID_MAX_GOALS][P]
    double *pid_acc_ang, *pid_acc_dist, *pid_last_ang, *pid_last_dist, *pid_target;
    int32_t *pid_state, *pid_backing, *pid_ngoals;
    double *pid_gx, *pid_gy;
    double pid_max_acc, pid_accel;
    // scheduling
    unsigned int* work_counter;
    int32_t block;                     // threads per CTA of this launch (host side; the kernels read blockDim.x)
    uint32_t rb_begin, rb_end;         // robots [rb_begin, rb_end) are this launch's work (the whole population, or one chunk of a
                                       // pipelined evaluation whose later genomes are still being uploaded)
    // Time slicing (warp-synchronous kernel, populations that run in several waves): the T ticks are cut into n_slices
    // slices of slice_T ticks; a warp's work item is (slice, batch of 32/LPR robots), handed out slice-major, and an item
    // waits until the previous slice of its batch has been written back (slice_done[batch] counts them).  Without it
    // the last, partly filled wave takes as long as a full one (32,768 32-ray robots = 3.46 waves ran like 3.8).
    int32_t n_slices, slice_T;         // n_slices <= 1: off
    uint32_t* slice_done;              // [batches], zeroed before the launch
};

struct StageView {
    const uint32_t* wall;
    const uint32_t* lvl50;
    const uint8_t* dist;
    int W, H, wpr;
    const uint32_t* hot;               // see hot_word / hot_shift; NULL: no fast path
    int hwpr;
    int hx0, hy0, hx1, hy1;            // pixels (inclusive) at which the fast plane in use may be read; empty: no fast path
    uint32_t win_sb;                   // window variant: shared-memory byte address of the word of pixel (0, 0) (may lie outside the window)
    int win_ww;                        //                 words per band of the window
};
__device__ __forceinline__ void view_global_hot(StageView& sv, const uint32_t* hot, int hwpr, int hpad) {
    sv.hot = hot; sv.hwpr = hwpr; sv.win_sb = 0u; sv.win_ww = 0;
    if (hot) { sv.hx0 = -hpad; sv.hy0 = -hpad; sv.hx1 = sv.W + hpad - 1; sv.hy1 = sv.H + hpad - 1; }
    else { sv.hx0 = 1; sv.hy0 = 1; sv.hx1 = 0; sv.hy1 = 0; }
}
__device__ __forceinline__ unsigned lds_u32(uint32_t saddr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// The "hot" plane: what the marches and the crash ring read.  One bit per pixel of the map AND of a margin of `hpad`
// pixels around it; a bit is set when the pixel is a wall (level 0) or is not an interior pixel of the map (the
// outermost pixel ring and everything outside) -- i.e. whenever a sample on that pixel is not simply "free, carry on".
// A 32-bit word covers 8 x 4 pixels (x by y), words of one 4-row band are contiguous: the 61 x 61 pixel neighbourhood a
// 30-pixel sonar fan touches is 16 bands x ~160 bytes instead of 61 rows of the row-major wall plane.  Word and bit of
// pixel (ix, iy), both possibly negative inside the margin (arithmetic shifts, hpad is a multiple of 8):
//     word = (iy >> 2) * hwpr + (ix >> 3)       relative to the pre-offset pointer
//     bit  = (ix + 8 * iy) & 31                 a permutation of the word's 32 pixels: one IMAD, and the funnel shift
//                                               that extracts it wraps the shift count by itself
__host__ __device__ __forceinline__ int hot_word(int ix, int iy, int hwpr) { return (iy >> 2) * hwpr + (ix >> 3); }
__host__ __device__ __forceinline__ int hot_shift(int ix, int iy) { return ix + 8 * iy; }
// floor(v) for |v| < 2^31 (negative values too) from the low word of RD(v + 1.5 * 2^52); no validity check
__device__ __forceinline__ int floor_lo(double v) { return __double2loint(__dadd_rd(v, 6755399441055744.0)); }

// one-lane genome tiles (see transpose_genomes_kernel)
__host__ __device__ __forceinline__ size_t tiled_base(unsigned rb, int G) { return (size_t)(rb >> 5) * (size_t)G * 32u + (rb & 31u); }
constexpr int kTileStride = 32;       // doubles between consecutive genes of one robot

// ---------------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(mbar) : "memory");
}

// release / acquire on a 32-bit flag in global memory (time slicing: a batch's next slice waits for the previous one)
__device__ __forceinline__ unsigned ld_acquire_u32(const uint32_t* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS), cached at L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src_gmem) : "memory");
}

// ---------------------------------------------------------------------------------------------
// pixel probes with the reference's (Python) index semantics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int dtrunc(double v) { return __double2int_rz(v); }   // int(v); saturates, NaN -> 0

// floor(v) for 0 <= v < 2^32 without an FP64 -> int conversion (F2I.F64 issues at a quarter of the DADD rate): the low
// word of RD(v + 2^52).  ok is false outside that range and for NaN; inside it floor == trunc, i.e. Python's int().
__device__ __forceinline__ bool fast_floor(double v, int& idx) {
    const double t = __dadd_rd(v, 4503599627370496.0);
    idx = __double2loint(t);
    return __double2hiint(t) == 0x43300000;
}

__device__ __forceinline__ bool plane_bit(const uint32_t* __restrict__ plane, int wpr, int ix, int iy) {
    return (plane[iy * wpr + (ix >> 5)] >> (ix & 31)) & 1u;
}
// env.item(ix, iy) == target-level?  Negative indices wrap, anything else out of range is an IndexError.
__device__ __forceinline__ bool px_probe(const uint32_t* __restrict__ plane, const StageView& s, int ix, int iy, bool& fault) {
    if (ix < 0) ix += s.W;
    if (iy < 0) iy += s.H;
    if ((unsigned)ix >= (unsigned)s.W || (unsigned)iy >= (unsigned)s.H) { fault = true; return false; }
    return plane_bit(plane, s.wpr, ix, iy);
}

struct Hit {
    double x, y;
    int k;            // hit step, -1 on miss
    unsigned nsamp;   // pixels tested (loop body executions)
    bool fault;
};

// utils.search_edge_in_angle_with_limit (r2p2/utils.py:420-432) with (vx, vy) = (cos, sin) already taken.
//
// Two tests of the reference loop are decided without arithmetic whenever that is provably safe:
//  * `norm(p - o) < limit`: p_k is o plus k accumulated unit steps, so |norm - k| < kNormSlack for every
//    in-map coordinate (|coord| < 2^15, k < 3e4: <= k * (ulp(2^15)/2) * sqrt(2) + k * 3e-16 < 2e-7).  The
//    test is therefore true for k < limit - slack and false for k >= limit + slack; only the (at most one)
//    k in between -- the knife-edge sample of SURVEY.md H5 -- is evaluated exactly, with the reference's
//    own arithmetic sqrt(fma(dy,dy, dx*dx)) < limit.
//  * `round(p.x) in range(0, W)`: true whenever the truncated pixel index is in [1, W-2]; otherwise the
//    half-to-even test is evaluated exactly.
__device__ __forceinline__ void norm_window(double limit, int& kA, int& kB) {
    kA = 0; kB = 0x7fffffff;              // k < kA: norm test true; k >= kB: false; else evaluated exactly
    if (limit > -1.0 && limit < 30000.0) {
        kA = (int)ceil(limit - kNormSlack);
        kB = (int)ceil(limit + kNormSlack);
        if (kA < 0) kA = 0;
        if (kB < 0) kB = 0;
    }
}

// One iteration of the reference loop at (x, y), evaluated with the reference's own tests.
// returns 0: body ran, no wall; 1: body ran, wall (hit); 2: while-condition false (loop ends); 3: the reference raises
enum { kSampleFree = 0, kSampleHit = 1, kSampleExit = 2, kSampleFault = 3 };
__device__ __noinline__ int sample_exact(const uint32_t* __restrict__ wall, int W, int H, int wpr, double x, double y,
                                         double ox, double oy, double limit, int need_norm) {
    const int ix = dtrunc(x), iy = dtrunc(y);
    const bool interior = ((unsigned)(ix - 1) < (unsigned)(W - 2)) & ((unsigned)(iy - 1) < (unsigned)(H - 2));
    if (!interior) {
        if (!isfinite(x) || !isfinite(y)) return kSampleFault;               // round(nan/inf) raises
        const double rx = rint(x), ry = rint(y);
        if (!(rx >= 0.0 && rx < (double)W && ry >= 0.0 && ry < (double)H)) return kSampleExit;
    }
    if (need_norm) {
        if (!(r2p2_norm2(R2P2_SUB(x, ox), R2P2_SUB(y, oy)) < limit)) return kSampleExit;
    }
    int jx = ix, jy = iy;                                                    // env.item(): negative indices wrap
    if (jx < 0) jx += W;
    if (jy < 0) jy += H;
    if ((unsigned)jx >= (unsigned)W || (unsigned)jy >= (unsigned)H) return kSampleFault;
    return plane_bit(wall, wpr, jx, jy) ? kSampleHit : kSampleFree;
}

// The plain restatement: every sample goes through sample_exact (kept as the in-kernel cross-check of march).
// (k0, x0, y0): resume the loop at sample k0, whose position is (x0, y0); the k0 samples before it ran the loop body.
__device__ __forceinline__ Hit march_plain(const StageView& s, double ox, double oy, double vx, double vy, double limit,
                                           int k0, double x0, double y0) {
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = (unsigned)k0; r.fault = false;
    int kA, kB;
    norm_window(limit, kA, kB);
    double x = x0, y = y0;
    for (int k = k0; k < kB; ++k) {
        const int st = sample_exact(s.wall, s.W, s.H, s.wpr, x, y, ox, oy, limit, k >= kA);
        if (st == kSampleFault) { r.fault = true; return r; }
        if (st == kSampleExit) return r;
        r.nsamp++;
        if (st == kSampleHit) { r.x = x; r.y = y; r.k = k; return r; }
        x = R2P2_ADD(x, vx);
        y = R2P2_ADD(y, vy);
    }
    return r;
}

__device__ __forceinline__ Hit march_plain(const StageView& s, double ox, double oy, double vx, double vy, double limit) {
    return march_plain(s, ox, oy, vx, vy, limit, 0, ox, oy);
}

// Distance-field march.  s.dist[iy*W + ix] is the Chebyshev distance (saturated at 255) from pixel (ix, iy) to the
// nearest BLOCKED pixel, blocked = wall (level 0) or the outermost pixel ring of the map.  If the sample at step k
// falls on a pixel with distance d >= 1 then, because consecutive samples are one unit step apart (pixel indices of
// sample k+j differ from those of sample k by at most j+1), the samples k .. k+max(1, d-1)-1 all lie on unblocked
// pixels strictly inside the ring: for them the reference's bounds test is true (pixel index in [1, W-2]), the pixel
// test is false, and -- for k below the norm window -- the limit test is true.  Only the position has to be carried
// forward, with the reference's own sequential FP64 additions.  Everything else goes through sample_exact.
__device__ __forceinline__ Hit march(const StageView& s, double ox, double oy, double vx, double vy, double limit) {
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = 0u; r.fault = false;
    int kA, kB;
    norm_window(limit, kA, kB);
    double x = ox, y = oy;
    int k = 0;
    while (k < kB) {
        unsigned d = 0u;
        bool wall_here = false;
        if (k < kA) {
            int ix, iy;
            const bool okx = fast_floor(x, ix), oky = fast_floor(y, iy);
            const bool inmap = okx & oky & ((unsigned)ix < (unsigned)s.W) & ((unsigned)iy < (unsigned)s.H);
            if (inmap) d = __ldg(s.dist + (size_t)iy * s.W + ix);
            // distance 0 on an interior pixel is a wall (the only other blocked pixels are the outer ring): a hit
            wall_here = inmap & (d == 0u) & ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
        }
        if (wall_here) {
            r.nsamp++;
            r.x = x; r.y = y; r.k = k;
            return r;
        }
        if (d == 0u) {
            const int st = sample_exact(s.wall, s.W, s.H, s.wpr, x, y, ox, oy, limit, k >= kA);
            if (st == kSampleFault) { r.fault = true; return r; }
            if (st == kSampleExit) return r;
            r.nsamp++;
            if (st == kSampleHit) { r.x = x; r.y = y; r.k = k; return r; }
            x = R2P2_ADD(x, vx);
            y = R2P2_ADD(y, vy);
            ++k;
        } else {
            int n = max(1, (int)d - 1);
            n = min(n, kA - k);
            r.nsamp += (unsigned)n;
            k += n;
            for (; n >= 4; n -= 4) {
                x = R2P2_ADD(R2P2_ADD(R2P2_ADD(R2P2_ADD(x, vx), vx), vx), vx);
                y = R2P2_ADD(R2P2_ADD(R2P2_ADD(R2P2_ADD(y, vy), vy), vy), vy);
            }
            for (; n > 0; --n) {
                x = R2P2_ADD(x, vx);
                y = R2P2_ADD(y, vy);
            }
        }
    }
    return r;
}

// A team: `tl` lanes of a warp (a power of two, an aligned segment) that cooperate on ONE ray.
struct Team {
    unsigned mask;    // lanes of this team
    unsigned sync;    // member mask for the collectives: the same value in every lane that calls team_march together
    int j, tl;
};

// Norm window of a ray of a given limit (hoisted out of the tick loop for the sonar rays, whose limit is a per-run
// constant).  nfast / chunk describe the contiguous-chunk split of the first team march (tools/ubench/chunked_march.cuh,
// kept for the micro-benchmarks that compare the two layouts); the kernels use kA, kB and qstar.
struct MarchPlan {
    int kA, kB;       // norm window
    int nfast;        // lanes that test the samples below the window; with a knife-edge sample and tl > 1 the last
                      // lane of the team is reserved for it
    int chunk;        // samples below the window per such lane
    double qstar;     // sqrt(q) < limit  <=>  q < qstar  (sqrt is monotone): the knife-edge test without the sqrt
};
__device__ __forceinline__ MarchPlan make_plan(double limit, int tl) {
    MarchPlan m;
    norm_window(limit, m.kA, m.kB);
    m.nfast = tl; m.chunk = 0; m.qstar = 0.0;
    if (m.kB == 0x7fffffff) return m;
    if (m.kB > m.kA) {
        // smallest double whose square root reaches the limit
        double q = R2P2_MUL(limit, limit);
        for (int it = 0; it < 8 && q > 0.0 && R2P2_SQRT(r2p2_from_bits(r2p2_bits(q) - 1)) >= limit; ++it) q = r2p2_from_bits(r2p2_bits(q) - 1);
        for (int it = 0; it < 8 && R2P2_SQRT(q) < limit; ++it) q = r2p2_from_bits(r2p2_bits(q) + 1);
        m.qstar = q;
        if (tl > 1) m.nfast = tl - 1;
    }
    // two divisions by (tl) and (tl - 1) rather than one by nfast: tl is a compile-time constant at the hot call site
    m.chunk = (m.nfast == tl) ? (m.kA + tl - 1) / tl : (m.kA + tl - 2) / (tl - 1);
    return m;
}

// Strided team march (the one the kernels use).  TL lanes -- a power of two, an aligned segment of the warp --
// cooperate on ONE ray; lane j owns the samples k = j, j + TL, j + 2 TL, ...  A warp issues in order, so the code is
// laid out as the hardware should see it: one unpredicated chain of the reference's sequential additions (a predicated
// DADD chain runs at half speed, measured), the pixel word of the hot plane requested whenever the lane passes one of
// its own samples, and the tests after the chain, when the loads have landed; NOWN own samples (TL * NOWN samples of
// the ray) per pass, fully unrolled.  Per own sample: 2 DADD (position), 2 DADD.RD (floor), word / shift arithmetic, one
// load, two funnel shifts that push the sample's event bit into a per-pass mask.  A set bit means "wall, or not an
// interior pixel"; the lane's first one ends its walk and is classified (hit / leave it to the exact loop).  The
// earliest event of the ray -- wall hit, while-condition false at the knife-edge sample, or a sample off the interior
// of the map -- is a min-reduction over the team of (k << 2 | kind); in the last case the exact loop takes over from
// that sample.  All collectives are issued with tm.sync, one mask value shared by every lane that executes this code
// together.
//
// Nothing is range-checked per sample: with |v| <= 1 per axis the walk stays within k + 1 pixels of the origin pixel
// after k steps and (passes are not cut at kB) ends below kB + TL * NOWN, so one box test per ray against the pixels the
// plane in use covers -- the hot plane's margin, or (WIN) the robot's window of it in shared memory -- decides.  Rays
// that fail it (far off the map, a limit beyond the margin, no plane) go to the exact loop from sample 0.
template <int TL, int NOWN, bool WIN = false>
__device__ __forceinline__ Hit team_march_t(const StageView& s, const Team& tm, const MarchPlan& mp, double ox, double oy,
                                            double vx, double vy, double limit) {
    const int kA = mp.kA, kB = mp.kB;
    if (kB == 0x7fffffff) return march_plain(s, ox, oy, vx, vy, limit);     // non-finite / huge limit: uniform in the team
    Hit r;
    r.x = -1.0; r.y = -1.0; r.k = -1; r.nsamp = 0u; r.fault = false;
    const int j = tm.j;
    constexpr unsigned kNone = 0xffffffffu;
    constexpr unsigned kSampleSlow = 3u;
    unsigned key = kNone;
    double x = ox, y = oy, hx = -1.0, hy = -1.0;
    // every pixel the walk can read -- within kB + TL * NOWN of the origin pixel -- must be inside the plane in use
    bool fast = (r2p2_fabs(ox) < 1e9) & (r2p2_fabs(oy) < 1e9) & (r2p2_fabs(vx) <= 1.0) & (r2p2_fabs(vy) <= 1.0);
    {
        const int iox = floor_lo(ox), ioy = floor_lo(oy), reach = kB + TL * NOWN + 1;
        fast &= (iox - reach >= s.hx0) & (iox + reach <= s.hx1) & (ioy - reach >= s.hy0) & (ioy + reach <= s.hy1);
    }
    const int kEnd = fast ? kB : 0;
    if (!fast && kB > 0) { key = kSampleSlow; hx = ox; hy = oy; }     // (0 << 2 | slow): the exact loop runs the whole ray
    const int nwalk = min(TL, kEnd) - 1;                   // to this lane's first sample: additions of v or of 0.0
#pragma unroll 4
    for (int i = 0; i < nwalk; ++i) {
        const bool on = i < j;
        x = R2P2_ADD(x, on ? vx : 0.0);
        y = R2P2_ADD(y, on ? vy : 0.0);
    }
    const bool has_window = kB > kA;                       // uniform
    const bool win_owner = has_window & ((kA & (TL - 1)) == j);
    double xw = 0.0, yw = 0.0;
    const uint32_t* __restrict__ hot = s.hot;
    const int hwpr = WIN ? s.win_ww : s.hwpr;
    for (int k0 = j; k0 < kEnd; k0 += TL * NOWN) {
        // Position of an own sample when it is needed afterwards (the lane's first event, the knife-edge sample).  One
        // lane per ray: re-walked from the pass start with the same additions (at most NOWN - 1 of them; no registers
        // held across the pass).  Teams: kept per sample -- a re-walk of up to TL * (NOWN - 1) dependent additions would
        // sit on the critical path of the latency-bound mappings (config B: +7 %).  The warp-synchronous kernel keeps them
        // too: there the re-walk is a loop whose trip count differs from lane to lane (measured on config E, 200 ticks:
        // 4 own samples kept 123 ms, re-walked 128 ms, 8 own samples re-walked 130 ms).
        constexpr int NKEEP = (TL > 1 || WIN) ? NOWN : 1;
        double xs[NKEEP], ys[NKEEP];
        xs[0] = x; ys[0] = y;                               // this lane's first sample of the pass
        unsigned w[NOWN];
        int sh[NOWN];
#pragma unroll
        for (int m = 0; m < NOWN; ++m) {
            if constexpr (NKEEP > 1) { xs[m] = x; ys[m] = y; }
            const int ix = floor_lo(x), iy = floor_lo(y);
            if constexpr (WIN) w[m] = lds_u32(s.win_sb + 4u * (unsigned)hot_word(ix, iy, hwpr));
            else w[m] = __ldg(hot + hot_word(ix, iy, hwpr));
            sh[m] = hot_shift(ix, iy);
            if (m + 1 < NOWN) {
#pragma unroll
                for (int u = 0; u < TL; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
            }
        }
        unsigned evm = 0u;                                  // sample m's bit ends up at bit 32 - NOWN + m
#pragma unroll
        for (int m = 0; m < NOWN; ++m) evm = __funnelshift_r(evm, __funnelshift_r(w[m], 0u, (unsigned)sh[m]), 1u);
        evm >>= (32 - NOWN);
        const int rem = kEnd - k0;                          // samples of the ray from this lane's first of the pass on
        if (rem <= TL * NOWN) {                             // last pass
            if (rem < TL * NOWN) evm &= (1u << ((rem + TL - 1) / TL)) - 1u;   // own samples at or beyond kB do not count
            if (win_owner) {                                // the knife-edge sample, if any, is the last one of the ray
                xw = xs[0]; yw = ys[0];
                if constexpr (NKEEP > 1) {
                    const int mw = (kA - k0) / TL;
#pragma unroll
                    for (int m = 1; m < NOWN; ++m) if (mw == m) { xw = xs[m]; yw = ys[m]; }
                } else {
                    for (int u = kA - k0; u > 0; --u) { xw = R2P2_ADD(xw, vx); yw = R2P2_ADD(yw, vy); }
                }
            }
        }
        if (evm) {                                          // this lane's first event: its position, its kind
            const int me = __ffs((int)evm) - 1;
            hx = xs[0]; hy = ys[0];
            if constexpr (NKEEP > 1) {
#pragma unroll
                for (int m = 1; m < NOWN; ++m) if (me == m) { hx = xs[m]; hy = ys[m]; }
            } else {
                for (int u = me; u > 0; --u) { hx = R2P2_ADD(hx, vx); hy = R2P2_ADD(hy, vy); }
            }
            const int ix = floor_lo(hx), iy = floor_lo(hy);
            const bool inter = ((unsigned)(ix - 1) < (unsigned)(s.W - 2)) & ((unsigned)(iy - 1) < (unsigned)(s.H - 2));
            key = ((unsigned)(k0 + me * TL) << 2) | (inter ? (unsigned)kSampleHit : kSampleSlow);
            break;                                          // this lane's later samples cannot come before its first event
        }
        // On to the lane's first sample of the next pass.  After the last pass the position is not read again: a lane that
        // walks its ray alone skips the test (the compiler turned it into four selects per pass: config E -1 %, 8,192
        // robots -3 %); teams keep it -- their TL additions would sit at the end of the in-order chain of the
        // latency-bound mappings (config B +2 % without it).
        if (TL == 1 || rem > TL * NOWN) {
#pragma unroll
            for (int u = 0; u < TL; ++u) { x = R2P2_ADD(x, vx); y = R2P2_ADD(y, vy); }
        }
    }
    // the knife-edge sample (at most one, k = kA = kB - 1): the reference's limit test comes before its pixel test, and
    // after its bounds test
    if (win_owner && fast) {
        const bool in_range = r2p2_norm2sq(R2P2_SUB(xw, ox), R2P2_SUB(yw, oy)) < mp.qstar;
        const bool earlier = (key != kNone) && ((int)(key >> 2) < kA || (key & 3u) == kSampleSlow);
        if (!in_range && !earlier) { key = ((unsigned)kA << 2) | (unsigned)kSampleExit; hx = xw; hy = yw; }
    }
    unsigned best = key;
    if (TL == 32) {
        best = __reduce_min_sync(tm.sync, key);
    } else {
#pragma unroll
        for (int o = TL >> 1; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(tm.sync, best, o));
    }
    if (TL > 1) {
        const int lane = (int)(threadIdx.x & 31u);
        const int src = (best != kNone) ? (lane - j + (int)((best >> 2) & (unsigned)(TL - 1))) : lane;   // the owner of that sample
        hx = __shfl_sync(tm.sync, hx, src);
        hy = __shfl_sync(tm.sync, hy, src);
    }
    if (best == kNone) { r.nsamp = (unsigned)kB; return r; }
    const int k = (int)(best >> 2);
    const unsigned kind = best & 3u;
    if (kind == kSampleSlow) return march_plain(s, ox, oy, vx, vy, limit, k, hx, hy);   // team-uniform
    if (kind == kSampleHit) {
        r.x = hx; r.y = hy; r.k = k; r.nsamp = (unsigned)(k + 1);
    } else {
        r.nsamp = (unsigned)k;          // the while-condition failed at sample k
    }
    return r;
}

// Team sizes the kernels use: sonar rays 1, 2, 4 or 8 lanes (the largest that gives every ray of the fan a team within
// the group), the collision ray the whole group.  MAXTL = lanes per robot.
constexpr int kMaxSonarTeam = 8;
// own samples per pass (maxown: of a lane that walks a ray alone, teamown: of teams of 2 or 4 lanes)
__host__ __device__ constexpr int sonar_nown(int tl, int maxown, int teamown) { return tl >= 32 ? 1 : tl >= 16 ? 2 : tl >= 8 ? 4 : tl >= 2 ? teamown : maxown; }
template <int MAXTL, int MAXOWN = 8, bool WIN = false, int TEAMOWN = MAXOWN>
__device__ __forceinline__ Hit team_march(const StageView& s, const Team& tm, const MarchPlan& mp, double ox, double oy,
                                          double vx, double vy, double limit) {
    constexpr int O = MAXOWN;
    switch (tm.tl) {                                       // uniform
        case 32: if constexpr (MAXTL >= 32) return team_march_t<32, 1, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
        case 16: if constexpr (MAXTL >= 16) return team_march_t<16, 2, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
        case 8: if constexpr (MAXTL >= 8) return team_march_t<8, 4, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
        case 4: if constexpr (MAXTL >= 4) return team_march_t<4, TEAMOWN, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
        case 2: if constexpr (MAXTL >= 2) return team_march_t<2, TEAMOWN, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
        default: return team_march_t<1, O, WIN>(s, tm, mp, ox, oy, vx, vy, limit);
    }
}

// Robot.__crashing_radius probes (r2p2/robot.py:181-186) of one lane: i = sub, sub + LPR, ... < 360, four in flight.
// returns the lane's earliest event as 2 i (IndexError at probe i) or 2 i + 1 (wall at probe i), 0xffffffff for none;
// the reference stops at the first wall, so an IndexError counts only if it comes before the first wall probe: after
// a min-reduction over the lanes, an even key is a fault, an odd key a crash.
template <int LPR, bool WIN = false>
__device__ __forceinline__ unsigned ring_scan(const StageView& sv, int sub, double x, double y, double radius,
                                              const double* s_crx, const double* s_cry) {
    unsigned key = 0xffffffffu;
    constexpr int kPer = (360 + LPR - 1) / LPR;
    // The whole ring strictly inside the map (the usual case: |offset| <= |radius|): every probe index is in range and
    // non-negative, so neither Python's index wrap nor an IndexError can occur and only the wall bit is read.
    const double ar = r2p2_fabs(radius);
    const bool inside = (R2P2_SUB(x, ar) >= 1.0) & (R2P2_ADD(x, ar) < (double)(sv.W - 1)) & (R2P2_SUB(y, ar) >= 1.0) & (R2P2_ADD(y, ar) < (double)(sv.H - 1));
    if (inside) {                                          // uniform in the group
        // (every probed pixel is an interior pixel, where the hot plane's bit is the wall bit)
        // the robot's window in shared memory when the whole ring lies inside it, else the hot plane, else the wall plane
        bool in_win = false;
        if constexpr (WIN) {
            const int icx = floor_lo(x), icy = floor_lo(y), rr = (int)ar + 2;
            in_win = (icx - rr >= sv.hx0) & (icx + rr <= sv.hx1) & (icy - rr >= sv.hy0) & (icy + rr <= sv.hy1);
        }
        const bool use_hot = sv.hot != nullptr;
        const uint32_t* __restrict__ plane = use_hot ? sv.hot : sv.wall;
#pragma unroll 1
        for (int m0 = 0; m0 < kPer; m0 += 4) {
            unsigned w[4];
            int bx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = sub + (m0 + u) * LPR;
                const int ii = (i < 360) ? i : 0;
                const int ix = floor_lo(R2P2_ADD(x, s_crx[ii])), iy = floor_lo(R2P2_ADD(y, s_cry[ii]));
                if (WIN && in_win) w[u] = lds_u32(sv.win_sb + 4u * (unsigned)hot_word(ix, iy, sv.win_ww));
                else w[u] = plane[use_hot ? hot_word(ix, iy, sv.hwpr) : (iy * sv.wpr + (ix >> 5))];
                bx[u] = use_hot ? hot_shift(ix, iy) : ix;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = sub + (m0 + u) * LPR;
                const bool hit = (i < 360) & ((__funnelshift_r(w[u], 0u, (unsigned)bx[u]) & 1u) != 0u);
                key = min(key, hit ? (unsigned)(2 * i + 1) : 0xffffffffu);
            }
        }
        return key;
    }
#pragma unroll 1
    for (int m0 = 0; m0 < kPer; m0 += 4) {
        unsigned w[4], ev[4];
        int bx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = sub + (m0 + u) * LPR;
            const bool ok = i < 360;
            const int ii = ok ? i : 0;
            const double px = R2P2_ADD(x, s_crx[ii]), py = R2P2_ADD(y, s_cry[ii]);
            int ix, iy;
            if (!(fast_floor(px, ix) & fast_floor(py, iy))) {   // negative / huge / NaN: Python's int() and index wrap, exactly
                ix = dtrunc(px); iy = dtrunc(py);
                if (ix < 0) ix += sv.W;
                if (iy < 0) iy += sv.H;
            }
            const bool f = ((unsigned)ix >= (unsigned)sv.W) | ((unsigned)iy >= (unsigned)sv.H);
            w[u] = sv.wall[f ? 0 : (iy * sv.wpr + (ix >> 5))];
            bx[u] = ix & 31;
            ev[u] = !ok ? 0xffffffffu : (f ? (unsigned)(2 * i) : 0xfffffffeu);   // 0xfffffffe: decided by the bit
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = sub + (m0 + u) * LPR;
            unsigned e = ev[u];
            if (e == 0xfffffffeu) e = ((w[u] >> bx[u]) & 1u) ? (unsigned)(2 * i + 1) : 0xffffffffu;
            key = min(key, e);
        }
    }
    return key;
}

// The same scan with a run-time number of lanes (the lanes of a warp that happen to be converged, one-lane mapping).
__device__ __forceinline__ unsigned ring_scan_rt(const StageView& sv, int sub, int LPR, double x, double y, double radius,
                                                 const double* s_crx, const double* s_cry) {
    unsigned key = 0xffffffffu;
    const int kPer = (360 + LPR - 1) / LPR;
    // The whole ring strictly inside the map (the usual case: |offset| <= |radius|): every probe index is in range and
    // non-negative, so neither Python's index wrap nor an IndexError can occur and only the wall bit is read.
    const double ar = r2p2_fabs(radius);
    const bool inside = (R2P2_SUB(x, ar) >= 1.0) & (R2P2_ADD(x, ar) < (double)(sv.W - 1)) & (R2P2_SUB(y, ar) >= 1.0) & (R2P2_ADD(y, ar) < (double)(sv.H - 1));
    if (inside) {                                          // uniform in the group
        // (every probed pixel is an interior pixel, where the hot plane's bit is the wall bit)
        const bool use_hot = sv.hot != nullptr;
        const uint32_t* __restrict__ plane = use_hot ? sv.hot : sv.wall;
#pragma unroll 1
        for (int m0 = 0; m0 < kPer; m0 += 4) {
            unsigned w[4];
            int bx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = sub + (m0 + u) * LPR;
                const int ii = (i < 360) ? i : 0;
                const int ix = floor_lo(R2P2_ADD(x, s_crx[ii])), iy = floor_lo(R2P2_ADD(y, s_cry[ii]));
                w[u] = plane[use_hot ? hot_word(ix, iy, sv.hwpr) : (iy * sv.wpr + (ix >> 5))];
                bx[u] = use_hot ? hot_shift(ix, iy) : ix;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = sub + (m0 + u) * LPR;
                const bool hit = (i < 360) & ((__funnelshift_r(w[u], 0u, (unsigned)bx[u]) & 1u) != 0u);
                key = min(key, hit ? (unsigned)(2 * i + 1) : 0xffffffffu);
            }
        }
        return key;
    }
#pragma unroll 1
    for (int m0 = 0; m0 < kPer; m0 += 4) {
        unsigned w[4], ev[4];
        int bx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = sub + (m0 + u) * LPR;
            const bool ok = i < 360;
            const int ii = ok ? i : 0;
            const double px = R2P2_ADD(x, s_crx[ii]), py = R2P2_ADD(y, s_cry[ii]);
            int ix, iy;
            if (!(fast_floor(px, ix) & fast_floor(py, iy))) {   // negative / huge / NaN: Python's int() and index wrap, exactly
                ix = dtrunc(px); iy = dtrunc(py);
                if (ix < 0) ix += sv.W;
                if (iy < 0) iy += sv.H;
            }
            const bool f = ((unsigned)ix >= (unsigned)sv.W) | ((unsigned)iy >= (unsigned)sv.H);
            w[u] = sv.wall[f ? 0 : (iy * sv.wpr + (ix >> 5))];
            bx[u] = ix & 31;
            ev[u] = !ok ? 0xffffffffu : (f ? (unsigned)(2 * i) : 0xfffffffeu);   // 0xfffffffe: decided by the bit
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = sub + (m0 + u) * LPR;
            unsigned e = ev[u];
            if (e == 0xfffffffeu) e = ((w[u] >> bx[u]) & 1u) ? (unsigned)(2 * i + 1) : 0xffffffffu;
            key = min(key, e);
        }
    }
    return key;
}

// A ray of this limit from (x, y) that cannot reach a blocked pixel: every sample k < kB lies within k (+1 for the
// truncation) pixels of the origin pixel, so with a distance-field value >= kB + 1 there all of them fall on unblocked
// interior pixels and the reference loop runs exactly kB times and misses.  Only taken when the norm window is empty
// (no knife-edge sample, which is the case unless the limit is within 2e-7 of an integer).
__device__ __forceinline__ bool ray_is_clear(const StageView& s, double x, double y, double limit, unsigned& nsamp) {
    int kA, kB;
    norm_window(limit, kA, kB);
    if ((kA != kB) | (kB > 250)) return false;
    const int ix = dtrunc(x), iy = dtrunc(y);
    if (!(((unsigned)ix < (unsigned)s.W) & ((unsigned)iy < (unsigned)s.H))) return false;   // also NaN / inf
    if ((int)__ldg(s.dist + (size_t)iy * s.W + ix) < kB + 1) return false;
    nsamp = (unsigned)kB;
    return true;
}

// ---------------------------------------------------------------------------------------------
// group (LPR lanes of one warp working on one robot)
// ---------------------------------------------------------------------------------------------
template <int LPR>
struct Group {
    unsigned mask;
    int sub;
    __device__ __forceinline__ Group() {
        const int lane = threadIdx.x & 31;
        sub = lane & (LPR - 1);
        mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
    }
    __device__ __forceinline__ void sync() const { if (LPR > 1) __syncwarp(mask); }
    __device__ __forceinline__ bool any(bool p) const { return (LPR == 1) ? p : (__ballot_sync(mask, p) != 0u); }
    __device__ __forceinline__ int min_i(int v) const {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(mask, v, o, 32));
        return v;
    }
    __device__ __forceinline__ unsigned sum_u(unsigned v) const {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 32);
        return v;
    }
    __device__ __forceinline__ unsigned min_u(unsigned v) const {
        if constexpr (LPR == 32) {
            return __reduce_min_sync(0xffffffffu, v);
        } else {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(mask, v, o, 32));
            return v;
        }
    }
    __device__ __forceinline__ unsigned bcast_u(unsigned v) const { return (LPR == 1) ? v : __shfl_sync(mask, v, 0, LPR); }
};

// ---------------------------------------------------------------------------------------------
// sensor noise: Robot.__generate_noise = np.random.normal(0, 0.5) (r2p2/robot.py:139-140)
// ---------------------------------------------------------------------------------------------
// mode 1 replays a host-generated stream (bit-exact with the reference: np.random.seed(s); np.random.normal(0, 0.5, n)
// is the reference's own draw sequence); mode 2 is the production generator: Philox4x32-10 keyed by the
// seed, counter = (draw index, robot), Box-Muller, sigma 0.5 -- statistically equivalent, order independent.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__device__ __forceinline__ double noise_draw(const KParams& p, unsigned rb, unsigned idx) {
    if (p.noise_mode == 1) return (idx < (unsigned)p.noise_len) ? __ldg(p.noise_stream + (size_t)rb * p.noise_len + idx) : 0.0;
    uint32_t c[4] = {idx, rb + p.robot_index0, 0u, 0u};
    uint32_t k0 = (uint32_t)p.noise_seed, k1 = (uint32_t)(p.noise_seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    const double u1 = ((double)(((unsigned long long)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
    const double u2 = (double)(((unsigned long long)c[2] << 21) ^ (c[3] >> 11)) * (1.0 / 9007199254740992.0);           // [0, 1)
    return 0.5 * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// Robot.__crashing_radius (r2p2/robot.py:181-186), only reached when the centre pixel is within ceil(radius)+1 of a
// blocked pixel.  returns 0 clear, 1 crashing, 2 the reference raises (IndexError before the first wall probe).
template <int LPR>
__device__ __forceinline__ int crash_ring(const StageView& sv, const Group<LPR>& g, double x, double y, double radius,
                                       const double* s_crx, const double* s_cry) {
    unsigned key = ring_scan<LPR>(sv, g.sub, x, y, radius, s_crx, s_cry);
    if (LPR > 1) key = g.min_u(key);
    if (key == 0xffffffffu) return 0;
    return (key & 1u) ? 1 : 2;
}

// (Measured: making sincos / tanh out-of-line calls to shrink the code footprint costs more in call overhead and
// spills than it saves in instruction-cache misses: 7.3 -> 8.5 ms on config B.  They stay inlined.)
struct SinCos {
    double s, c;
    int ok;
};
#ifdef R2P2_EXP_SINCOS_NOINLINE
__device__ __noinline__ SinCos sincos_call(double a, const double* tab, int rs = kSinCosRow) {
#else
__device__ __forceinline__ SinCos sincos_call(double a, const double* tab, int rs = kSinCosRow) {
#endif
    SinCos r;
    r.ok = 1;
    r2p2_sincos_rs(a, tab, rs, &r.s, &r.c, &r.ok);
    return r;
}

// tanhtab: the tanh table the kernel uses (its shared-memory copy when there is room for one, else the global one)
__device__ __forceinline__ double activate(int kind, double v, const double* tanhtab) {
    // tanh first, behind a plain (uniform) branch: a switch compiles to an indexed jump, whose target the warp waits for
    // once per neuron (8.7 % of the stall samples of the one-lane kernel)
    if (kind == R2P2_ACT_TANH) return r2p2_tanh_tab(v, tanhtab);
    if (kind == R2P2_ACT_RELU) {
        if (v != v) return v;
        if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
        return v > 0.0 ? v : 0.0;
    }
    if (kind == R2P2_ACT_LOGISTIC) return r2p2_logistic(v);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Sequential_PID_Controller.control (r2p2/controllers/pid_controller.py:40-262), one robot, scalar: every lane of a group
// evaluates it redundantly on the same values.  The arithmetic keeps the reference's evaluation order; its quirks are the
// reference's (see r2p2_b200/controllers/pid_controller.py for the list).  atan2 is r2p2_atan2 (<= 3 ulp from libm), the
// avoid state's cos / sin of +-90 degrees (+ turns) are the restated glibc ones.
// ---------------------------------------------------------------------------------------------
struct PidState {
    double acc_ang, acc_dist, last_ang, last_dist, target;
    int state, backing, ngoals;
};
__device__ __forceinline__ void pid_load(const KParams& p, unsigned rb, PidState& ps) {
    ps.acc_ang = p.pid_acc_ang[rb]; ps.acc_dist = p.pid_acc_dist[rb]; ps.last_ang = p.pid_last_ang[rb]; ps.last_dist = p.pid_last_dist[rb];
    ps.target = p.pid_target[rb]; ps.state = p.pid_state[rb]; ps.backing = p.pid_backing[rb]; ps.ngoals = p.pid_ngoals[rb];
}
__device__ __forceinline__ void pid_store(const KParams& p, unsigned rb, const PidState& ps) {
    p.pid_acc_ang[rb] = ps.acc_ang; p.pid_acc_dist[rb] = ps.acc_dist; p.pid_last_ang[rb] = ps.last_ang; p.pid_last_dist[rb] = ps.last_dist;
    p.pid_target[rb] = ps.target; p.pid_state[rb] = ps.state; p.pid_backing[rb] = ps.backing; p.pid_ngoals[rb] = ps.ngoals;
}
__device__ __forceinline__ double pymod360(double a) {      // Python float % 360
    double m = fmod(a, 360.0);
    if (m != 0.0) { if (m < 0.0) m = R2P2_ADD(m, 360.0); } else m = 0.0;
    return m;
}
__device__ __forceinline__ double pid_brake(const KParams& p, double& speed) {   // Robot.brake (robot.py:354-363)
    const bool neg = speed < 0.0;
    const double third = R2P2_DIV(speed, 3.0);
    const double new_speed = R2P2_SUB(R2P2_SUB(speed, p.pid_accel), third);
    if ((neg && new_speed > 0.0) || (!neg && new_speed < 0.0)) { speed = 0.0; return R2P2_DIV(r2p2_neg(p.pid_accel), 5.0); }
    return R2P2_SUB(r2p2_neg(p.pid_accel), third);
}
__device__ __forceinline__ double pid_angle_variation(PidState& ps, double theta, double ap, double ai, double ad) {
    const double e = R2P2_SUB(ps.target, theta);
    ps.acc_ang = R2P2_ADD(ps.acc_ang, e);
    const double de = R2P2_SUB(e, ps.last_ang);
    ps.last_ang = e;
    return R2P2_ADD(R2P2_ADD(R2P2_MUL(ap, e), R2P2_MUL(ai, ps.acc_ang)), R2P2_MUL(ad, de));
}
// dst: the robot's sonar distances (element i at dst[i * stride]); k: its six gains (element i at k[i * kstride]).
// returns false when the reference would raise (IndexError at the goal pixel) or the goal stack is full.
__device__ __forceinline__ bool pid_control(const KParams& p, const StageView& sv, PidState& ps, unsigned rb, bool writer, double x, double y,
                                            double theta, double& speed, const double* dst, int stride, const double* k, size_t kstride,
                                            const double* s_tab, unsigned& flags, double& v_out, double& w_out) {
    const int R = p.R;
    if (ps.backing) { ps.backing -= 1; v_out = theta; w_out = r2p2_neg(p.max_speed); return true; }
    const double rng = p.vr1;
    const double d0 = dst[0], d1 = dst[stride], d2 = dst[2 * stride], dm1 = dst[(R - 1) * stride], dm2 = dst[(R - 2) * stride];
    {
        const double tol = R2P2_MUL(R2P2_DIV(p.vr0, rng), 13.0);
        const bool danger = (R2P2_DIV(dm1, rng) < tol && R2P2_DIV(dm2, rng) < tol) || (R2P2_DIV(d1, rng) < tol && R2P2_DIV(d2, rng) < tol) ||
                            R2P2_DIV(d0, rng) < tol;
        ps.state = danger ? ((ps.state == 0) ? 1 : 2) : 0;
    }
    if (ps.state != 1) {                                   // control_advance
        if (ps.ngoals == 0) { v_out = 0.0; w_out = 0.0; return true; }
        const size_t slot = (size_t)(ps.ngoals - 1) * (size_t)p.P + rb;
        const double gx = p.pid_gx[slot], gy = p.pid_gy[slot];
        ps.target = R2P2_MUL(r2p2_atan2(R2P2_SUB(gy, y), R2P2_SUB(gx, x)), kDeg);
        const double angle = pid_angle_variation(ps, theta, k[0], k[kstride], k[2 * kstride]);
        const double speed0 = speed;
        double var;
        if (r2p2_fabs(R2P2_SUB(ps.target, theta)) > 45.0) {
            ps.acc_dist = R2P2_ADD(ps.acc_dist, pid_brake(p, speed));
            var = pid_brake(p, speed);
        } else {
            const double e = r2p2_norm2(R2P2_SUB(gx, x), R2P2_SUB(gy, y));
            if (p.pid_accel < p.pid_max_acc) ps.acc_dist = R2P2_ADD(ps.acc_dist, e);
            const double de = R2P2_SUB(e, ps.last_dist);
            ps.last_dist = e;
            var = R2P2_ADD(R2P2_ADD(R2P2_MUL(k[3 * kstride], e), R2P2_MUL(k[4 * kstride], ps.acc_dist)), R2P2_MUL(k[5 * kstride], de));
        }
        double acc = R2P2_ADD(p.pid_accel, var);
        if (acc > p.pid_max_acc) acc = p.pid_max_acc;
        else if (r2p2_neg(acc) < p.pid_max_acc) acc = r2p2_neg(p.pid_max_acc);
        v_out = R2P2_ADD(speed0, acc);
        w_out = angle;
        // is_at_goal: the goal pixel is a wall, or the robot is within 1.85 radii
        if (!(isfinite(gx) && isfinite(gy))) return false;                      // int(nan / inf) raises
        bool f = false;
        const bool wall = px_probe(sv.wall, sv, dtrunc(gx), dtrunc(gy), f);
        if (f) return false;
        if (wall || r2p2_norm2(R2P2_SUB(gx, x), R2P2_SUB(gy, y)) <= p.goal_r) { ps.ngoals -= 1; if (ps.state == 2) ps.state = 0; }
        return true;
    }
    // control_avoid
    double risk = 1.0;
    for (int i = 0; i < R; ++i) risk = R2P2_MUL(risk, R2P2_DIV(dst[i * stride], rng));
    if (R2P2_ADD(d1, d2) < R2P2_ADD(dm1, dm2)) risk = R2P2_DIV(risk, R2P2_MUL(R2P2_DIV(rng, dm1), R2P2_DIV(rng, dm2)));
    else risk = R2P2_DIV(risk, R2P2_MUL(R2P2_DIV(rng, d1), R2P2_DIV(rng, d2)));
    if (R2P2_DIV(d0, rng) < 0.45) risk = R2P2_MUL(risk, R2P2_DIV(d0, rng));
    double safety = R2P2_MUL(R2P2_MUL(R2P2_ADD(rng, p.radius), risk), 2.0);
    if (safety <= p.vr0) safety = R2P2_MUL(p.vr0, 2.0);
    const double tm = pymod360(ps.target);
    const double a = (R2P2_ADD(d1, d2) > R2P2_ADD(dm1, dm2)) ? R2P2_ADD(tm, 90.0) : R2P2_SUB(tm, 90.0);
    const SinCos sc = sincos_call(R2P2_MUL(R2P2_SUB(ps.target, a), kRad), s_tab);
    if (!sc.ok) { flags |= R2P2_F_TRIGRANGE; return false; }
    double x_var = R2P2_MUL(safety, sc.c), y_var = R2P2_MUL(safety, sc.s);
    const double am = pymod360(a);
    if (am > 90.0 && am < 270.0) x_var = r2p2_neg(x_var);
    if (am > 180.0) y_var = r2p2_neg(y_var);
    if (ps.ngoals >= R2P2_PID_MAX_GOALS) return false;
    if (writer) {
        const size_t slot = (size_t)ps.ngoals * (size_t)p.P + rb;
        p.pid_gx[slot] = R2P2_ADD(x, x_var); p.pid_gy[slot] = R2P2_ADD(y, y_var);
    }
    ps.ngoals += 1;
    (void)pid_angle_variation(ps, theta, k[0], k[kstride], k[2 * kstride]);
    ps.target = a;
    bool must_retreat = false;
    const double lim = R2P2_MUL(R2P2_DIV(p.vr0, rng), 12.0);
    for (int i = 0; i < R; ++i) must_retreat |= R2P2_DIV(dst[i * stride], rng) < lim;
    v_out = must_retreat ? r2p2_neg(p.max_speed) : speed;
    w_out = 0.0;
    return true;
}

// ---------------------------------------------------------------------------------------------
// the episode kernel
// ---------------------------------------------------------------------------------------------
// dynamic shared memory layout (bytes):
//   [0,16)                         mbarrier
//   [16, 16+4400)                  sin/cos table, rows padded to 5 doubles
//   [.., +5760)                    crash-radius offsets (360 x, 360 y)
//   [.., +18304)                   tanh table (tanh_in_smem only)
//   [.., +plane_bytes)             wall plane      (SMEM_STAGE only)
//   [.., +plane_bytes)             level-50 plane  (SMEM_STAGE and has50 only)
//   [.., +2*maxw*NR*8)             MLP / sonar buffers A and B, element i of robot slot r at [i*NR + r]; NR = buf_stride(LPR, blockDim.x / LPR)
//   [.., +NR*GS*8)                 genomes of the robots in flight (w_in_smem only), row r at [r*GS], GS = smem_gstride(LPR, G)
//   [.., +NR*win_words*4)          per-robot windows of the hot plane (warp-synchronous kernel)
// Row stride (doubles) of the genome copies in shared memory.  The groups of a warp read the same weight columns of
// different robots in one instruction (LPR consecutive doubles each); with a stride that is a multiple of 16 doubles all
// of them would hit the same banks (config E: G = 656 -> 4-way conflicts on every weight load).  A stride congruent to
// LPR modulo 16 spreads the 32 / LPR groups evenly over the 32 banks.
__host__ __device__ inline int smem_gstride(int lpr, int G) { return G + ((lpr - G) % 16 + 16) % 16; }

// Stride (doubles) of the activation buffers, element i of robot slot r at [i * stride + r].  With several lanes per robot
// the lanes of a group touch DIFFERENT elements of ONE robot (ray leaders writing their distances, the 25/dst pass,
// neurons writing their outputs): at a stride that is a multiple of 16 doubles all of them hit one bank pair (config E,
// 64 robots per CTA: 16 wavefronts per such instruction, a fifth of the kernel's shared-memory wavefronts).  An odd stride
// puts consecutive elements two banks apart.  One lane per robot: lanes are consecutive slots, no padding needed.
__host__ __device__ inline int buf_stride(int lpr, int robots) { return lpr == 1 ? robots : (robots | 1); }

__host__ __device__ inline size_t episode_smem_bytes(int block, int lpr, bool smem_stage, uint32_t plane_bytes, int has50, int maxw, int w_in_smem, int G,
                                                     int win_words = 0, int tanh_in_smem = 0) {
    size_t n = 16 + kSinCosPadLen * 8 + 720 * 8;
    if (tanh_in_smem) n += (size_t)R2P2_TANH_NINT * R2P2_TANH_ROW * 8;
    if (smem_stage) n += (size_t)plane_bytes * (has50 ? 2 : 1);
    n += (size_t)2 * maxw * buf_stride(lpr, block / lpr) * 8;
    if (w_in_smem) n += (size_t)(block / lpr) * smem_gstride(lpr, G) * 8;
    if (win_words) n += (size_t)(block / lpr) * (size_t)win_words * 4;
    return n;
}

// ---------------------------------------------------------------------------------------------
// fixed-shape MLP: weights in registers, activations exchanged by warp shuffles
// ---------------------------------------------------------------------------------------------
// The reference's shipped networks (conf/controller-neuro.json 5-10-7-4-2, controller-neuro-simple.json 5-1-2) get a
// kernel instance whose layer sizes are compile-time constants: every loop unrolls, lane j keeps the weights of
// "its" neurons (j, j+LPR, ...) in registers for the whole episode and reads the previous layer with __shfl_sync.
// The arithmetic is the generic path's (ordered FMA chain per neuron), so results are bit-identical to it.
template <int N0, int N1, int N2, int N3, int N4>
struct NetShape {
    static constexpr bool kSmemWeights = false;
    static constexpr int kLayers = (N4 > 0) ? 5 : (N3 > 0) ? 4 : (N2 > 0) ? 3 : 2;       // layer sizes incl. input/output
    __host__ __device__ static constexpr int sz(int l) { return l == 0 ? N0 : l == 1 ? N1 : l == 2 ? N2 : l == 3 ? N3 : N4; }
    __host__ __device__ static constexpr int woff(int l) {       // offset of layer l's weights in the flat genome
        int o = 0;
        for (int q = 0; q < l; ++q) o += sz(q) * sz(q + 1);
        return o;
    }
    __host__ __device__ static constexpr int sum_in() {           // sum of the input widths of all weight layers
        int t = 0;
        for (int q = 0; q + 1 < kLayers; ++q) t += sz(q);
        return t;
    }
    __host__ __device__ static constexpr int maxw() {
        int m = 0;
        for (int q = 0; q < kLayers; ++q) m = sz(q) > m ? sz(q) : m;
        return m;
    }
};
struct GenericNet {
    static constexpr int kLayers = 0;
    static constexpr bool kSmemWeights = false;
};
// Compile-time shape, weights read from the shared-memory copy of the genome instead of registers: for the mappings
// with few lanes per robot, where register-resident weight columns would cost the occupancy the run lives on.
template <class SHAPE>
struct SmemNet : SHAPE {
    static constexpr bool kSmemWeights = true;
};
using NetNeuro = NetShape<5, 10, 7, 4, 2>;     // conf/controller-neuro.json on robot-neuro.json (5 rays)
using NetSimple = NetShape<5, 1, 2, 0, 0>;     // conf/controller-neuro-simple.json
using NetSweep = NetShape<32, 16, 8, 2, 0>;    // BASELINE.json configs[4]: the 32-ray synthetic sweep

template <int LPR, class NET>
struct FixedMlp {
    static constexpr int kMaxW = NET::maxw();
    static constexpr int kSlots = (kMaxW + LPR - 1) / LPR;           // neurons per lane (layer width above LPR)
    static constexpr int kWPerLane = (LPR == 1) ? 1 : kSlots * NET::sum_in();   // upper bound; slots a layer does not need are never touched
    double w[kWPerLane];

    // lane `sub` loads the weight columns of its neurons: layer l, slot s, input i -> W_l[i][sub + s*LPR]
    __device__ __forceinline__ void load(const double* __restrict__ gen, int sub) {
        int q = 0;
#pragma unroll
        for (int l = 0; l + 1 < NET::kLayers; ++l) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const int j = sub + s * LPR;
#pragma unroll
                for (int i = 0; i < NET::sz(l); ++i) {
                    w[q] = (s * LPR < NET::sz(l + 1) && j < NET::sz(l + 1)) ? gen[NET::woff(l) + i * NET::sz(l + 1) + j] : 0.0;
                    ++q;
                }
            }
        }
    }
    // h[s] on entry: input (s*LPR + sub) of layer 0 (unused slots ignored); returns outputs 0 and 1 in all lanes
    __device__ __forceinline__ void forward(unsigned mask, int sub, double (&h)[kSlots], int activation, const double* tanhtab, double& out0, double& out1) const {
        int q = 0;
#pragma unroll
        for (int l = 0; l + 1 < NET::kLayers; ++l) {
            double hn[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NET::sz(l); ++i) {
                    const double hi = (LPR == 1) ? h[i] : __shfl_sync(mask, h[i / LPR], i % LPR, LPR);
                    if (s * LPR < NET::sz(l + 1)) acc = R2P2_FMA(hi, w[q], acc);
                    ++q;
                }
                hn[s] = (s * LPR < NET::sz(l + 1)) ? ((l + 2 < NET::kLayers) ? activate(activation, acc, tanhtab) : acc) : 0.0;
            }
#pragma unroll
            for (int s = 0; s < kSlots; ++s) h[s] = hn[s];
        }
        out0 = (LPR == 1) ? h[0] : __shfl_sync(mask, h[0], 0, LPR);
        out1 = (LPR == 1) ? h[1] : __shfl_sync(mask, h[(1 / LPR)], 1 % LPR, LPR);
    }
    // Same chains with the weights in shared memory (gs = this robot's genome copy): W_l[i][j] at gs[woff(l) + i*sz(l+1) + j],
    // compile-time i, lane-dependent j.
    __device__ __forceinline__ void forward_smem(const double* gs, unsigned mask, int sub, double (&h)[kSlots], int activation, const double* tanhtab,
                                                 double& out0, double& out1) const {
#pragma unroll
        for (int l = 0; l + 1 < NET::kLayers; ++l) {
            double hn[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const int j = sub + s * LPR;
                const int jc = (j < NET::sz(l + 1)) ? j : 0;          // lanes without a neuron read a valid column
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NET::sz(l); ++i) {
                    const double hi = (LPR == 1) ? h[i] : __shfl_sync(mask, h[i / LPR], i % LPR, LPR);
                    if (s * LPR < NET::sz(l + 1)) acc = R2P2_FMA(hi, gs[NET::woff(l) + i * NET::sz(l + 1) + jc], acc);
                }
                hn[s] = (s * LPR < NET::sz(l + 1)) ? ((l + 2 < NET::kLayers) ? activate(activation, acc, tanhtab) : acc) : 0.0;
            }
#pragma unroll
            for (int s = 0; s < kSlots; ++s) h[s] = hn[s];
        }
        out0 = (LPR == 1) ? h[0] : __shfl_sync(mask, h[0], 0, LPR);
        out1 = (LPR == 1) ? h[1] : __shfl_sync(mask, h[(1 / LPR)], 1 % LPR, LPR);
    }
    // One lane per robot: the whole net in this thread.  Activations stay in registers, every loop is unrolled, and
    // the weights stream from the robot's tile of the genome array (element q at wt[q * 32]: one coalesced line per warp
    // and weight).  Same ordered FMA chains as everywhere else.
    __device__ __forceinline__ void forward_global(const double* __restrict__ wt, double (&h)[kSlots], int activation, const double* tanhtab,
                                                   double& out0, double& out1) const {
#pragma unroll
        for (int l = 0; l + 1 < NET::kLayers; ++l) {
            double hn[kSlots];
#pragma unroll
            for (int j = 0; j < NET::sz(l + 1); ++j) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NET::sz(l); ++i)
                    acc = R2P2_FMA(h[i], __ldg(wt + (NET::woff(l) + i * NET::sz(l + 1) + j) * kTileStride), acc);
                hn[j] = (l + 2 < NET::kLayers) ? activate(activation, acc, tanhtab) : acc;
            }
#pragma unroll
            for (int j = 0; j < NET::sz(l + 1); ++j) h[j] = hn[j];
        }
        out0 = h[0];
        out1 = h[1];
    }
};
// One lane per robot, compile-time shape: activations in shared memory (stride NR), weights from the transposed genome
// tiles (wt[q * 32]: weight q of the warp's 32 robots is one coalesced 256-byte read); one activation instance per layer.
template <class NET, int L>
__device__ __forceinline__ void lane_layer(const double* __restrict__ wt, const double* hin, double* hout, int NR, int activation, const double* tanhtab) {
    constexpr int nin = NET::sz(L), nout = NET::sz(L + 1);
    constexpr bool last = (L + 2 == NET::kLayers);
    // Input-major: row i of the layer (W[i][0 .. nout-1] = nout consecutive genes of the robot, compile-time offsets from
    // its tile pointer) feeds all nout neurons at once.  Every neuron keeps the reference's ordered chain
    // acc_j = fma(h_i, w_ij, acc_j), i ascending, but the nout chains are independent: the FMA latency and the L2 latency
    // of the next row's loads overlap instead of being paid once per neuron.
    double acc[nout];
#pragma unroll
    for (int j = 0; j < nout; ++j) acc[j] = 0.0;
#pragma unroll
    for (int i = 0; i < nin; ++i) {
        double wv[nout];
#pragma unroll
        for (int j = 0; j < nout; ++j) wv[j] = __ldcg(wt + (NET::woff(L) + i * nout + j) * 32);   // L2 only: the stream of weight rows must not evict the hot plane from L1
        const double hi = hin[i * NR];
#pragma unroll
        for (int j = 0; j < nout; ++j) acc[j] = R2P2_FMA(hi, wv[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < nout; ++j) hout[j * NR] = acc[j];
    if (!last) {
        // (measured: a branch-free tanh evaluated for two neurons at a time, so that their chains overlap, is 5 % slower --
        // the table rows are picked per lane and the shared-memory pipe, not the chain, is what the activations wait for)
#pragma unroll 1
        for (int j = 0; j < nout; ++j) hout[j * NR] = activate(activation, hout[j * NR], tanhtab);
    }
}
template <class NET>
__device__ __forceinline__ void lane_mlp(const double* __restrict__ wt, double* a, double* b, int NR, int activation, const double* tanhtab,
                                         double& out0, double& out1) {
    lane_layer<NET, 0>(wt, a, b, NR, activation, tanhtab);
    double* res = b;
    if constexpr (NET::kLayers > 2) { lane_layer<NET, 1>(wt, b, a, NR, activation, tanhtab); res = a; }
    if constexpr (NET::kLayers > 3) { lane_layer<NET, 2>(wt, a, b, NR, activation, tanhtab); res = b; }
    if constexpr (NET::kLayers > 4) { lane_layer<NET, 3>(wt, b, a, NR, activation, tanhtab); res = a; }
    out0 = res[0];
    out1 = res[NR];
}

// Compile-time shape, LPR lanes per robot, activations in shared memory (element i of this robot at buf[i * NR]): lane
// `sub` owns the neurons sub, sub + LPR, ... of every layer.  One pass over the inputs feeds all of the lane's neurons
// (one LDS per input, broadcast to the group), the weights come through `gs` -- the robot's genome in shared memory or
// in the [P][G] array -- and each neuron keeps the reference's ordered chain acc = fma(h_i, w_ij, acc), i ascending.
// (The first version exchanged activations with __shfl_sync: two SHFL per double and input were 11 % of the synthetic
// sweep's instructions.)
// (Measured on config E: keeping only the two small layers' weights (144 of 656) in shared memory, at the price of the tanh
// table's place there, is 3.5 % slower than reading everything through L2.)
// (Measured on config E, where the genomes of the robots in flight fit neither shared memory nor L1 and every batch of
// weight loads waits for L2 -- 9 % of the run in long-scoreboard stalls: prefetch.global.L1 of the next batch's lines, one
// per lane and batch, made the run 20 % slower.)
template <int LPR, class NET, int L>
__device__ __forceinline__ void group_layer(const double* __restrict__ gs, const double* hin, double* hout, int NR, int sub, int activation, const double* tanhtab) {
    constexpr int nin = NET::sz(L), nout = NET::sz(L + 1), slots = (nout + LPR - 1) / LPR;
    constexpr bool last = (L + 2 == NET::kLayers);
    const double* wp[slots];
    double acc[slots];
#pragma unroll
    for (int q = 0; q < slots; ++q) {
        const int jn = sub + q * LPR;
        wp[q] = gs + NET::woff(L) + (jn < nout ? jn : 0);          // lanes without a neuron read a valid column
        acc[q] = 0.0;
    }
#pragma unroll 8
    for (int i = 0; i < nin; ++i) {
        const double hi = hin[i * NR];
#pragma unroll
        for (int q = 0; q < slots; ++q) acc[q] = R2P2_FMA(hi, wp[q][i * nout], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < slots; ++q) {
        const int jn = sub + q * LPR;
        if (jn < nout) hout[jn * NR] = last ? acc[q] : activate(activation, acc[q], tanhtab);
    }
}
// in: a[i * NR] = input i (already vr1 / dst); a and b are overwritten; returns outputs 0 and 1 in every lane of the group.
// `sync` is the group-wide barrier (a __syncwarp in the warp-synchronous kernel).
template <int LPR, class NET>
__device__ __forceinline__ void group_mlp(const double* __restrict__ gs, double* a, double* b, int NR, int sub, int activation, const double* tanhtab, double& out0, double& out1) {
    group_layer<LPR, NET, 0>(gs, a, b, NR, sub, activation, tanhtab);
    __syncwarp();
    double* res = b;
    if constexpr (NET::kLayers > 2) { group_layer<LPR, NET, 1>(gs, b, a, NR, sub, activation, tanhtab); __syncwarp(); res = a; }
    if constexpr (NET::kLayers > 3) { group_layer<LPR, NET, 2>(gs, a, b, NR, sub, activation, tanhtab); __syncwarp(); res = b; }
    if constexpr (NET::kLayers > 4) { group_layer<LPR, NET, 3>(gs, b, a, NR, sub, activation, tanhtab); __syncwarp(); res = a; }
    out0 = res[0];
    out1 = res[NR];
}

template <int LPR>
struct FixedMlp<LPR, GenericNet> {
    static constexpr int kSlots = 1;
};

// EXTRA: goal flagging / r2p2_move's skip_sense compiled in.  The plain instance is what evaluations run; keeping the
// rarely used switches out of it is worth 2-3 % on config B (register allocation of the latency-bound chain).
template <int LPR, bool SMEM_STAGE, class NET, bool EXTRA = false>
__global__ void __launch_bounds__((NET::kLayers > 0 && !NET::kSmemWeights) ? kRegBlock : kMaxBlock, 1) episode_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NRC = (int)blockDim.x / LPR;                // robots in flight per CTA
    const int NR = buf_stride(LPR, NRC);                  // stride of the activation buffers
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_tab = reinterpret_cast<double*>(smem_raw + 16);
    double* s_crx = s_tab + kSinCosPadLen;
    double* s_cry = s_crx + 360;
    unsigned char* cur = reinterpret_cast<unsigned char*>(s_cry + 360);
    const double* s_tanh = p.g_tanhtab;
    if (p.tanh_in_smem) { s_tanh = reinterpret_cast<const double*>(cur); cur += R2P2_TANH_NINT * R2P2_TANH_ROW * 8; }
    uint32_t* s_wall = nullptr;
    uint32_t* s_l50 = nullptr;
    if (SMEM_STAGE) {
        s_wall = reinterpret_cast<uint32_t*>(cur); cur += p.plane_bytes;
        if (p.has50) { s_l50 = reinterpret_cast<uint32_t*>(cur); cur += p.plane_bytes; }
    }
    double* s_bufA = reinterpret_cast<double*>(cur);
    double* s_bufB = s_bufA + (size_t)p.maxw * NR;
    double* s_gen = s_bufB + (size_t)p.maxw * NR;         // [NR][G], only when p.w_in_smem

    // ---- stage the read-only tables (and the bit planes) into shared memory with the TMA unit ----
    const uint32_t mb = smem_u32(mbar);
    if (threadIdx.x == 0) {
        mbar_init(mb, 1);
        uint32_t total = kSinCosPadLen * 8 + 720 * 8;
        if (SMEM_STAGE) total += p.plane_bytes * (p.has50 ? 2u : 1u);
        if (p.tanh_in_smem) total += R2P2_TANH_NINT * R2P2_TANH_ROW * 8;
        mbar_expect_tx(mb, total);
        tma_load_1d(s_tab, p.g_sincostab_pad, kSinCosPadLen * 8, mb);
        tma_load_1d(s_crx, p.g_crxy, 720 * 8, mb);
        if (p.tanh_in_smem) tma_load_1d(const_cast<double*>(s_tanh), p.g_tanhtab, R2P2_TANH_NINT * R2P2_TANH_ROW * 8, mb);
        if (SMEM_STAGE) {
            for (uint32_t off = 0; off < p.plane_bytes; off += 32768u) {
                const uint32_t n = min(32768u, p.plane_bytes - off);
                tma_load_1d(reinterpret_cast<unsigned char*>(s_wall) + off, reinterpret_cast<const unsigned char*>(p.g_wall) + off, n, mb);
                if (p.has50)
                    tma_load_1d(reinterpret_cast<unsigned char*>(s_l50) + off, reinterpret_cast<const unsigned char*>(p.g_lvl50) + off, n, mb);
            }
        }
    }
    __syncthreads();
    while (!mbar_try_wait(mb, 0)) { }

    StageView sv;
    sv.wall = SMEM_STAGE ? s_wall : p.g_wall;
    sv.lvl50 = SMEM_STAGE ? s_l50 : p.g_lvl50;
    sv.dist = p.g_dist;
    sv.W = p.W; sv.H = p.H; sv.wpr = p.wpr;
    view_global_hot(sv, p.g_hot, p.hwpr, p.hpad);

    const Group<LPR> g;
    const int slot = threadIdx.x / LPR;                   // robot slot inside the CTA
    double* bufA = s_bufA + slot;
    double* bufB = s_bufB + slot;
    const int R = p.R;
    const double delta = p.delta;
    const bool tracing = p.tr_pose != nullptr;
    int tl = 1;                                           // lanes per sonar ray: the largest power of two with R teams in the group
    while (LPR > 1 && tl * 2 * R <= LPR && tl < kMaxSonarTeam) tl *= 2;
    const int rays_per_round = LPR / tl;
    const int team_ray = g.sub / tl, team_j = g.sub % tl; // this lane's ray slot and position inside its team
    Team sonar_team;
    sonar_team.tl = tl; sonar_team.j = team_j; sonar_team.sync = g.mask;
    sonar_team.mask = (tl >= 32) ? 0xffffffffu : (((1u << tl) - 1u) << ((threadIdx.x & 31) - team_j));
    const MarchPlan sonar_plan = make_plan(p.L, tl);

    for (;;) {
        // ---- fetch the next robot --------------------------------------------------------------
        unsigned rb = 0;
        if (g.sub == 0) rb = atomicAdd(p.work_counter, 1u) + p.rb_begin;
        rb = g.bcast_u(rb);
        if (rb >= p.rb_end) break;

        double x = p.x[rb], y = p.y[rb], theta = p.theta[rb], speed = p.speed[rb], omega = p.omega[rb];
        double last_x = p.last_x[rb], last_y = p.last_y[rb];
        double dist = p.dist[rb], odom_x = p.odom_x[rb], odom_y = p.odom_y[rb];
        const double origin_x = p.origin_x[rb], origin_y = p.origin_y[rb];
        double tmr = p.time[rb];
        double fitness = p.fitness[rb];
        unsigned flags = p.flags[rb], ncollide = p.ncollide[rb], ticks = p.ticks[rb];
        unsigned ndraws = p.noise_mode ? p.ndraws[rb] : 0u;
        unsigned my_ns_sonar = 0u, ns_coll = 0u;          // my_ns_sonar: this lane's rays only
        // the genome is read every tick for T ticks: keep it next to the SM
        const double* gen = (LPR == 1) ? p.genomes + tiled_base(rb, p.G) : p.genomes + (size_t)rb * (size_t)p.G;   // one lane: the robot's column of its tile
        FixedMlp<LPR, NET> mlp;
        if constexpr (NET::kLayers > 0 && !NET::kSmemWeights) {
            if constexpr (LPR > 1) mlp.load(gen, g.sub);
        } else if (LPR > 1 && p.w_in_smem) {
            double* dst_g = s_gen + (size_t)slot * smem_gstride(LPR, p.G);
            for (int i = g.sub; i < p.G; i += LPR) dst_g[i] = __ldg(gen + i);
            g.sync();
            gen = dst_g;
        }

        unsigned goal_tick = (EXTRA && p.goal_on) ? p.goal_tick[rb] : 0xffffffffu;
        const bool pid = EXTRA && p.ctrl_kind == R2P2_CTRL_PID;      // Sequential_PID_Controller on the device
        PidState ps;
        ps.acc_ang = ps.acc_dist = ps.last_ang = ps.last_dist = ps.target = 0.0; ps.state = ps.backing = ps.ngoals = 0;
        if (pid) pid_load(p, rb, ps);
        PT_DECL
        int t = 0;
        unsigned ncollide0 = ncollide;
        for (; t < p.T; ++t) {
            if (flags & (R2P2_F_FAULT | R2P2_F_DONE)) break;
            ncollide0 = ncollide;
            const size_t trow = (size_t)t * (size_t)p.P + rb;
            unsigned ccode = 0u;                          // collide() calls of this tick, for the trace
            double ccx = 0.0, ccy = 0.0, cbv = 0.0;

            PT(11)
            // ---- sense: Robot.check_sensors (robot.py:287-318) --------------------------------
            bool fault = !(EXTRA && p.skip_sense) && !(isfinite(x) && isfinite(y));   // round(nan) / round(inf) raise
            const double ox = rint(x), oy = rint(y);      // Python round(): half to even
            if (!fault && !(EXTRA && p.skip_sense)) {                // (uniform in the group)
                // Rays are spread over the group: `tl` lanes per ray (team march) when the group has lanes to spare.
                // Every lane of the group walks through the same code (lanes without a ray march a dummy one), so that
                // the warp-level collectives inside team_march see one member mask.
                for (int base = 0; base < R; base += rays_per_round) {
                    const int i = base + team_ray;
                    const bool have_ray = (team_ray < rays_per_round) & (i < R);
                    {
                        const double ang = R2P2_ADD(p.ray_deg[have_ray ? i : 0], theta);
                        PT(11)
                        const SinCos sc = sincos_call(R2P2_MUL(ang, kRad), s_tab);
                        const double vs = sc.s, vc = sc.c;
                        PT(8)
                        if (!sc.ok && have_ray) { flags |= R2P2_F_TRIGRANGE; fault = true; }
                        Hit h;
                        if (LPR == 1) {
                            // one thread walks the whole ray: every sample's bit-plane load is independent of the others
                            // (the distance-field march chains each load to the one before it and diverges between robots)
                            h = sc.ok ? team_march_t<1, 8>(sv, sonar_team, sonar_plan, ox, oy, vc, vs, p.L) : Hit{-1.0, -1.0, -1, 0u, true};
                        } else {
                            h = team_march<(LPR < kMaxSonarTeam ? LPR : kMaxSonarTeam)>(sv, sonar_team, sonar_plan, ox, oy, vc, vs, p.L);
                        }
                        PT(9)
                        if (h.fault && have_ray) fault = true;
                        const bool leader = have_ray && (LPR == 1 || team_j == 0);
                        double d = 0.0;
                        bool unclamped = false;
                        if (leader) {
                            my_ns_sonar += h.nsamp;
                            d = (h.k >= 0) ? r2p2_norm2(R2P2_SUB(h.x, ox), R2P2_SUB(h.y, oy)) : r2p2_from_bits(0x7ff0000000000000ULL);
                            if ((d < p.db_hi && d > p.db_lo) || d > p.L) d = p.L;
                            else unclamped = true;
                        }
                        if (p.noise_mode) {
                            // robot.py:308-310: one N(0, 0.5) draw per unclamped ray, in ray order (= lane order here)
                            unsigned nb = unclamped ? 1u : 0u, below = 0u;
                            if (LPR > 1) {
                                const unsigned gbase = (threadIdx.x & 31u) & ~(unsigned)(LPR - 1);
                                nb = (__ballot_sync(g.mask, unclamped) >> gbase) & ((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u));
                                below = nb & ((1u << g.sub) - 1u);
                            }
                            if (unclamped) d = R2P2_ADD(d, noise_draw(p, rb, ndraws + (unsigned)__popc(below)));
                            ndraws += (unsigned)__popc(nb);
                        }
                        if (leader) {
                            bufA[i * NR] = d;
                            if (tracing) {
                                p.tr_dst[trow * R + i] = d;
                                p.tr_hitk[trow * R + i] = h.k;
                                p.tr_hitpx[(trow * R + i) * 2] = (h.k >= 0) ? dtrunc(h.x) : -1;
                                p.tr_hitpx[(trow * R + i) * 2 + 1] = (h.k >= 0) ? dtrunc(h.y) : -1;
                                p.tr_nsamp[trow * R + i] = h.nsamp;
                            }
                        }
                    }
                }
            }
            if (LPR > 1) {
                const unsigned tr_bit = __ballot_sync(g.mask, (flags & R2P2_F_TRIGRANGE) != 0u);
                if (tr_bit) flags |= R2P2_F_TRIGRANGE;
            }
            if (g.any(fault)) { flags |= R2P2_F_FAULT; break; }
            g.sync();

            PT(0)
            // ---- control: Neuro_controller.control / Inspyred_controller.control -----------------
            dist = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(x, odom_x), R2P2_SUB(y, odom_y)));
            odom_x = x; odom_y = y;
            if (p.ctrl_kind != R2P2_CTRL_CONST && !pid) {
                if (p.evolve) tmr = R2P2_SUB(tmr, p.delta_ctrl);
                if (tmr <= 0.0 && p.evolve) {             // episode over: fitness is frozen here
                    fitness = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(odom_x, origin_x), R2P2_SUB(odom_y, origin_y)));
                    flags |= R2P2_F_DONE;
                    speed = 0.0; omega = 0.0;
                    break;
                }
            }
            PT(1)
            double v, w;
            if constexpr (NET::kSmemWeights) {            // one lane per robot, compile-time shape (capi dispatches it for LPR == 1 only)
#pragma unroll
                for (int i = 0; i < NET::sz(0); ++i) bufA[i * NR] = R2P2_DIV(p.vr1, bufA[i * NR]);   // 25/dst (inf when dst == 0)
                lane_mlp<NET>(gen, bufA, bufB, NR, p.activation, s_tanh, v, w);
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
            } else if constexpr (NET::kLayers > 0) {
                double h[FixedMlp<LPR, NET>::kSlots];
#pragma unroll
                for (int s = 0; s < FixedMlp<LPR, NET>::kSlots; ++s) {
                    const int i = g.sub + s * LPR;
                    h[s] = (i < NET::sz(0)) ? R2P2_DIV(p.vr1, bufA[i * NR]) : 0.0;                     // 25/dst (inf when dst == 0)
                }
                if constexpr (LPR > 1) mlp.forward(g.mask, g.sub, h, p.activation, s_tanh, v, w);
                else mlp.forward_global(gen, h, p.activation, s_tanh, v, w);
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
                g.sync();                                  // bufA is rewritten by the next tick's sense
            } else if (p.ctrl_kind == R2P2_CTRL_NEURO) {
                for (int i = g.sub; i < R; i += LPR) bufA[i * NR] = R2P2_DIV(p.vr1, bufA[i * NR]);   // 25/dst (inf when dst == 0)
                g.sync();
                double* hin = bufA;
                double* hout = bufB;
                int woff = 0;
                for (int l = 0; l + 1 < p.n_layers; ++l) {
                    const int nin = p.layers[l], nout = p.layers[l + 1];
                    const bool last = (l + 2 == p.n_layers);
                    for (int j = g.sub; j < nout; j += LPR) {
                        double acc = 0.0;
                        if (LPR == 1) {
                            const double* wp = gen + (size_t)(woff + j) * kTileStride;
                            for (int i = 0; i < nin; ++i) acc = R2P2_FMA(hin[i * NR], __ldg(wp + (size_t)i * nout * kTileStride), acc);
                        } else {
                            // batches of 8: all loads first (independent of the chain), then the ordered FMA chain;
                            // out-of-range slots read a clamped (valid) address and are predicated off
                            const double* wp = gen + woff + j;
                            for (int i0 = 0; i0 < nin; i0 += 8) {
                                double wv[8], hv[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    const int i = min(i0 + u, nin - 1);
                                    wv[u] = wp[i * nout];
                                    hv[u] = hin[i * NR];
                                }
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    if (i0 + u < nin) acc = R2P2_FMA(hv[u], wv[u], acc);
                            }
                        }
                        hout[j * NR] = last ? acc : activate(p.activation, acc, s_tanh);
                    }
                    g.sync();
                    double* tsw = hin; hin = hout; hout = tsw;
                    woff += nin * nout;
                }
                v = hin[0];
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
                w = hin[NR];
                g.sync();                                  // buffers are rewritten by the next tick's sense
            } else if (p.ctrl_kind == R2P2_CTRL_LINEAR) {
                const int n = p.G / 2 - 1;
                const int m = n < R ? n : R;
                double lin = 0.0, angv = 0.0;
                if (LPR == 1) {
                    for (int i = 0; i < m; ++i) lin = R2P2_ADD(lin, R2P2_MUL(__ldg(gen + (size_t)i * kTileStride), bufA[i * NR]));
                    for (int i = 0; i < m; ++i) angv = R2P2_ADD(angv, R2P2_MUL(__ldg(gen + (size_t)(n + i) * kTileStride), bufA[i * NR]));
                } else {
                    for (int i = 0; i < m; ++i) lin = R2P2_ADD(lin, R2P2_MUL(gen[i], bufA[i * NR]));
                    for (int i = 0; i < m; ++i) angv = R2P2_ADD(angv, R2P2_MUL(gen[n + i], bufA[i * NR]));
                }
                v = lin; w = angv;
                g.sync();
            } else if (pid) {
                // every lane of the group evaluates the controller on the same values; lane 0 writes the goal stack
                const bool ok = (LPR == 1) ? pid_control(p, sv, ps, rb, true, x, y, theta, speed, bufA, NR, gen, (size_t)kTileStride, s_tab, flags, v, w)
                                           : pid_control(p, sv, ps, rb, g.sub == 0, x, y, theta, speed, bufA, NR, gen, (size_t)1, s_tab, flags, v, w);
                g.sync();                                  // bufA is rewritten by the next tick's sense; a pushed goal is visible to the group
                if (!ok) { flags |= R2P2_F_FAULT; break; }
            } else {
                if (LPR == 1) { v = __ldg(gen); w = __ldg(gen + kTileStride); }
                else { v = gen[0]; w = gen[1]; }
            }
            PT(2)
            if (tracing && g.sub == 0) {                   // the controller's return value, as its log_step records it
                double* ci = p.tr_cinfo + trow * kCinfoRow;
                ci[4] = v; ci[5] = w;
            }
            // Robot.control clamp (robot.py:386-387)
            if (r2p2_fabs(v) > p.max_speed) v = R2P2_MUL(R2P2_DIV(v, r2p2_fabs(v)), p.max_speed);
            speed = v; omega = w;

            // ---- integrate: __update_angle, __update_position (robot.py:188-267) ------------------
            theta = R2P2_ADD(theta, R2P2_MUL(omega, delta));
            {
                const SinCos hsc = sincos_call(R2P2_MUL(theta, kRad), s_tab);
                const double hs = hsc.s, hc = hsc.c;
                if (!hsc.ok) { flags |= R2P2_F_TRIGRANGE | R2P2_F_FAULT; break; }
                const double cx = R2P2_ADD(x, R2P2_MUL(R2P2_MUL(hc, speed), delta));
                const double cy = R2P2_ADD(y, R2P2_MUL(R2P2_MUL(hs, speed), delta));
                const double dx = R2P2_SUB(cx, x), dy = R2P2_SUB(cy, y);
                PT(3)
                double angle = R2P2_MUL(r2p2_atan2(dy, dx), kDeg);
                // Python float % 360 for |angle| <= 180: fmod is the identity, then the sign fix-up
                if (angle != 0.0) { if (angle < 0.0) angle = R2P2_ADD(angle, 360.0); } else angle = 0.0;
                const SinCos csc = sincos_call(R2P2_MUL(angle, kRad), s_tab);
                const double cs_ = csc.s, cc_ = csc.c;
                if (!csc.ok) { flags |= R2P2_F_TRIGRANGE | R2P2_F_FAULT; break; }   // NaN displacement: the reference faults two statements later
                PT(4)
                Hit ms;
                const double climit = R2P2_ADD(r2p2_norm2(dx, dy), p.radius);
                ms.x = -1.0; ms.y = -1.0; ms.k = -1; ms.nsamp = 0u; ms.fault = false;
                if (ray_is_clear(sv, x, y, climit, ms.nsamp)) {
                    // open space around the robot (group-uniform: same x, y, climit in every lane): the ray misses
                } else if (LPR == 1) {
                    ms = march(sv, x, y, cc_, cs_, climit);
                } else {                                   // the whole group is one team on the collision ray
                    Team tm;
                    tm.tl = LPR; tm.j = g.sub; tm.mask = g.mask; tm.sync = g.mask;
                    ms = team_march_t<LPR, (LPR >= 4 ? 32 / LPR : 8)>(sv, tm, make_plan(climit, LPR), x, y, cc_, cs_, climit);
                }
                ns_coll += ms.nsamp;
                if (ms.fault) { flags |= R2P2_F_FAULT; break; }
                PT(5)
                x = cx; y = cy;
                if (ms.k >= 0) {
                    if (!(isfinite(cx) && isfinite(cy))) { flags |= R2P2_F_FAULT; break; }   // int(inf) raises
                    bool f = false;
                    const bool inwall = px_probe(sv.wall, sv, dtrunc(cx), dtrunc(cy), f);
                    if (f) { flags |= R2P2_F_FAULT; break; }
                    if (inwall) {
                        ncollide++; tmr = 0.0;
                        ccode |= 1u; ccx = x; ccy = y;
                        const double ex = R2P2_SUB(ms.x, x), ey = R2P2_SUB(ms.y, y);
                        if (ex > 0.0) x = R2P2_ADD(ms.x, p.rad_p1); else if (ex < 0.0) x = R2P2_SUB(ms.x, p.rad_p1);
                        if (ey > 0.0) y = R2P2_ADD(ms.y, p.rad_p1); else if (ey < 0.0) y = R2P2_SUB(ms.y, p.rad_p1);
                    }
                }
                const double dW = (double)p.W, dH = (double)p.H, rad = p.radius;
                // (border cases 1..8 of robot.py:222-247, recorded for the collision log: the spot mixes map sizes and this coordinate)
                if (R2P2_ADD(x, rad) >= dW) {
                    ncollide++; tmr = 0.0; cbv = y;
                    if (R2P2_ADD(y, rad) < dH && R2P2_SUB(y, rad) >= 0.0) { ccode |= 1u << 1; }
                    else if (R2P2_ADD(y, rad) >= dH) { ccode |= 2u << 1; y = R2P2_SUB(R2P2_SUB(dH, 1.0), rad); }
                    else { ccode |= 3u << 1; y = rad; }
                    x = R2P2_SUB(dW, 1.0);
                } else if (R2P2_SUB(x, rad) < 0.0) {
                    ncollide++; tmr = 0.0; cbv = y;
                    if (R2P2_ADD(y, rad) < dH && R2P2_SUB(y, rad) >= 0.0) { ccode |= 4u << 1; }
                    else if (R2P2_ADD(y, rad) >= dH) { ccode |= 5u << 1; y = R2P2_SUB(R2P2_SUB(dH, 1.0), rad); }
                    else { ccode |= 6u << 1; y = rad; }
                    x = 0.0;
                } else if (R2P2_ADD(y, rad) > dH) {
                    ncollide++; tmr = 0.0; cbv = x; ccode |= 7u << 1;
                    y = R2P2_SUB(R2P2_SUB(dH, 1.0), rad);
                } else if (R2P2_SUB(y, rad) < 0.0) {
                    ncollide++; tmr = 0.0; cbv = x; ccode |= 8u << 1;
                    y = rad;
                }
                PT(6)
                // env[int(x), int(y)] is 50 or __crashing_radius (robot.py:249, 181-186)
                if (!(isfinite(x) && isfinite(y))) { flags |= R2P2_F_FAULT; break; }
                bool f = false;
                bool crash = false;
                if (p.has50) crash = px_probe(sv.lvl50, sv, dtrunc(x), dtrunc(y), f);
                else { int ix = dtrunc(x), iy = dtrunc(y); if (ix < 0) ix += p.W; if (iy < 0) iy += p.H; f = ((unsigned)ix >= (unsigned)p.W) | ((unsigned)iy >= (unsigned)p.H); }
                if (f) { flags |= R2P2_F_FAULT; break; }
                // Far enough from every blocked pixel (wall or outer ring): none of the 360 probes -- all within
                // ceil(radius)+1 pixels of the centre pixel -- can be a wall or leave the map, so the ring is skipped.
                bool ring_clear = false;
                {
                    const int cxp = dtrunc(x), cyp = dtrunc(y);
                    if (((unsigned)cxp < (unsigned)p.W) & ((unsigned)cyp < (unsigned)p.H))
                        ring_clear = (int)__ldg(p.g_dist + (size_t)cyp * p.W + cxp) >= p.crash_clear;
                }
                if constexpr (LPR == 1) {
                    // One lane per robot: a lane alone would walk 360 probes while the rest of its warp waits.  Instead
                    // the lanes that are at this point together (whatever robots and ticks they are on) serve the robots
                    // among them that need the ring one after the other: 360 / n probes per lane and one REDUX each.
                    const unsigned act = __activemask();
                    const unsigned lane = threadIdx.x & 31u;
                    const bool need = !crash && !ring_clear;
                    unsigned todo = __ballot_sync(act, need);
                    const int cnt = __popc(act), idx = __popc(act & ((1u << lane) - 1u));
                    unsigned mykey = 0xffffffffu;
                    if (__popc(todo) > R2P2_RING_T) {     // many of them need it: every lane scans its own ring, all in parallel
                        if (need) mykey = ring_scan<1>(sv, 0, x, y, p.radius, s_crx, s_cry);
                        todo = 0u;
                    }
                    while (todo) {
                        const int owner = __ffs(todo) - 1;
                        todo &= todo - 1u;
                        const double rx = __shfl_sync(act, x, owner), ry = __shfl_sync(act, y, owner);
                        const unsigned key = __reduce_min_sync(act, ring_scan_rt(sv, idx, cnt, rx, ry, p.radius, s_crx, s_cry));
                        if ((int)lane == owner) mykey = key;
                    }
                    if (need && mykey != 0xffffffffu) {
                        if (!(mykey & 1u)) { flags |= R2P2_F_FAULT; break; }
                        crash = true;
                    }
                } else if (!crash && !ring_clear) {
                    // 360 probes spread over the group; the reference stops at the first wall, so an
                    // IndexError only counts if it comes before the first wall probe
                    const int cr = crash_ring<LPR>(sv, g, x, y, p.radius, s_crx, s_cry);
                    if (cr == 2) { flags |= R2P2_F_FAULT; break; }
                    crash = cr == 1;
                }
                if (crash) { x = last_x; y = last_y; ncollide++; tmr = 0.0; ccode |= 1u << 5; }
                else { last_x = x; last_y = y; }
            }
            PT(7)
            ticks++;
            if (ncollide != ncollide0) flags |= R2P2_F_COLLIDED;
            if (pid && ncollide != ncollide0) ps.backing = 10;      // on_collision (pid_controller.py:166-167)
            if ((EXTRA && p.goal_on) && !(flags & R2P2_F_GOAL)) {   // Sequential_PID_Controller.is_at_goal (pid_controller.py:207-213), flag only
                if (p.goal_is_wall || r2p2_norm2(R2P2_SUB(p.goal_x, x), R2P2_SUB(p.goal_y, y)) <= p.goal_r) { flags |= R2P2_F_GOAL; goal_tick = ticks; }
            }
            if (tracing && g.sub == 0) {
                double* q = p.tr_pose + trow * 5;
                q[0] = x; q[1] = y; q[2] = theta; q[3] = speed; q[4] = omega;
                p.tr_ncoll[trow] = ncollide - ncollide0;
                double* ci = p.tr_cinfo + trow * kCinfoRow;
                ci[0] = (double)ccode; ci[1] = ccx; ci[2] = ccy; ci[3] = cbv;
            }
        }
        PT_FLUSH(rb == 0 && g.sub == 0)
        if (ncollide != p.ncollide[rb]) flags |= R2P2_F_COLLIDED;
        if (pid && ncollide != ncollide0) ps.backing = 10;          // collide() calls of a tick the reference raised in
        if (tracing && g.sub == 0) {     // rows of ticks that did not complete: frozen pose (as the oracle records them)
            for (int t2 = t; t2 < p.T; ++t2) {
                const size_t trow = (size_t)t2 * (size_t)p.P + rb;
                double* q = p.tr_pose + trow * 5;
                q[0] = x; q[1] = y; q[2] = theta; q[3] = speed; q[4] = omega;
                p.tr_ncoll[trow] = (t2 == t) ? (ncollide - ncollide0) : 0u;
            }
        }
        if (!(flags & R2P2_F_DONE))      // controller.fitness() after the last update
            fitness = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(odom_x, origin_x), R2P2_SUB(odom_y, origin_y)));
        const unsigned ns_sonar = (LPR == 1) ? my_ns_sonar : g.sum_u(my_ns_sonar);
        if (p.noise_mode == 1 && ndraws > (unsigned)p.noise_len) flags |= R2P2_F_NOISE;   // the replayed stream ran out
        if (g.sub == 0) {
            p.x[rb] = x; p.y[rb] = y; p.theta[rb] = theta; p.speed[rb] = speed; p.omega[rb] = omega;
            p.last_x[rb] = last_x; p.last_y[rb] = last_y;
            p.dist[rb] = dist; p.odom_x[rb] = odom_x; p.odom_y[rb] = odom_y;
            p.time[rb] = tmr; p.fitness[rb] = fitness;
            p.flags[rb] = flags; p.ncollide[rb] = ncollide; p.ticks[rb] = ticks;
            p.nsamp_sonar[rb] += ns_sonar; p.nsamp_coll[rb] += ns_coll;
            if (p.noise_mode) p.ndraws[rb] = ndraws;
            if ((EXTRA && p.goal_on)) p.goal_tick[rb] = goal_tick;
            if (pid) pid_store(p, rb, ps);
        }
        g.sync();
    }
}

// ---------------------------------------------------------------------------------------------
// small service kernels
// ---------------------------------------------------------------------------------------------
// levels[x*H + y] (uint8) -> bit planes [H][wpr]; one thread per output word
__global__ void pack_stage_kernel(const uint8_t* __restrict__ levels, int W, int H, int wpr,
                                  uint32_t* __restrict__ wall, uint32_t* __restrict__ lvl50, int* __restrict__ has50) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    const int wx = blockIdx.y;
    if (iy >= H) return;
    uint32_t bw = 0u, b5 = 0u;
    for (int b = 0; b < 32; ++b) {
        const int ix = wx * 32 + b;
        if (ix < W) {
            const uint8_t v = levels[(size_t)ix * H + iy];
            bw |= (uint32_t)(v == 0) << b;
            b5 |= (uint32_t)(v == 50) << b;
        }
    }
    wall[(size_t)iy * wpr + wx] = bw;
    lvl50[(size_t)iy * wpr + wx] = b5;
    if (b5) atomicOr(has50, 1);
}

// Hot plane (see hot_word / hot_shift): one thread per 32-bit word = 8 x 4 pixels of the padded map.
__global__ void hot_plane_kernel(const uint8_t* __restrict__ levels, int W, int H, int hpad, int hwpr, int hrows, uint32_t* __restrict__ hot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, rr = blockIdx.y;
    if (c >= hwpr || rr >= hrows) return;
    uint32_t bits = 0u;
    for (int dy = 0; dy < 4; ++dy) {
        const int iy = 4 * (rr - hpad / 4) + dy;
        for (int dx = 0; dx < 8; ++dx) {
            const int ix = 8 * (c - hpad / 8) + dx;
            const bool interior = (ix >= 1) & (ix <= W - 2) & (iy >= 1) & (iy <= H - 2);
            const bool blocked = !interior || levels[(size_t)ix * H + iy] == 0;
            bits |= (uint32_t)blocked << (hot_shift(ix, iy) & 31);
        }
    }
    hot[(size_t)rr * hwpr + c] = bits;
}

// Distance field, pass 1: per map row, distance along x to the nearest blocked pixel (wall or outer ring), <= 255.
// One thread per row (one-time setup work; the walk is sequential in x by nature).
__global__ void dist_rows_kernel(const uint8_t* __restrict__ levels, int W, int H, uint8_t* __restrict__ hrow) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    if (iy >= H) return;
    const bool ring_row = (iy == 0) | (iy == H - 1);
    int run = 0;
    for (int ix = 0; ix < W; ++ix) {
        const bool blocked = ring_row | (ix == 0) | (ix == W - 1) | (levels[(size_t)ix * H + iy] == 0);
        run = blocked ? 0 : min(run + 1, 255);
        hrow[(size_t)iy * W + ix] = (uint8_t)run;
    }
    run = 0;
    for (int ix = W - 1; ix >= 0; --ix) {
        const int cur = hrow[(size_t)iy * W + ix];
        run = (cur == 0) ? 0 : min(run + 1, 255);
        hrow[(size_t)iy * W + ix] = (uint8_t)min(cur, run);
    }
}
// pass 2: Chebyshev distance D(x, y) = min over dy of max(|dy|, hrow(x, y + dy)); rows outside the map are blocked.
__global__ void dist_cols_kernel(const uint8_t* __restrict__ hrow, int W, int H, uint8_t* __restrict__ dist) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    if (ix >= W) return;
    int best = hrow[(size_t)iy * W + ix];
    for (int dy = 1; dy < best; ++dy) {
        const int up = (iy - dy >= 0) ? hrow[(size_t)(iy - dy) * W + ix] : 0;
        const int dn = (iy + dy < H) ? hrow[(size_t)(iy + dy) * W + ix] : 0;
        best = min(best, max(dy, min(up, dn)));
    }
    dist[(size_t)iy * W + ix] = (uint8_t)best;
}

__global__ void i32_to_u8_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, size_t n, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = in[i];
    if (v < 0 || v > 255) atomicOr(bad, 1);
    out[i] = (uint8_t)v;
}

// One-lane mapping: the genomes in tiles of 32 robots, [ceil(P/32)][G][32] -- gene q of robot rb at
// ((rb >> 5) * G + q) * 32 + (rb & 31).  A warp's 32 robots read gene q as one coalesced 256-byte line, and for one robot
// consecutive genes are a compile-time 256 bytes apart: the loads of the compile-time-shaped MLP take immediate offsets
// from one pointer per robot (with a [G][P] array every weight cost four address instructions next to its load and FMA).
// Rows lo .. hi-1 of `in` ([P][G]); lo is a multiple of 32 or 0.
__global__ void transpose_genomes_kernel(const double* __restrict__ in, double* __restrict__ out, int lo, int hi, int G) {
    __shared__ double tile[32][33];
    const int p0 = lo + blockIdx.x * 32, g0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int pp = p0 + r, gg = g0 + threadIdx.x;
        if (pp < hi && gg < G) tile[r][threadIdx.x] = in[(size_t)pp * G + gg];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gg = g0 + r, pp = p0 + threadIdx.x;
        if (pp < hi && gg < G) out[tiled_base((unsigned)pp, G) + (size_t)gg * kTileStride] = tile[threadIdx.x][r];
    }
}

struct ResetParams {
    uint32_t* goal_tick;
    int P, n;
    const double *x0, *y0, *theta0;
    double time_budget;
    double *x, *y, *theta, *speed, *omega, *last_x, *last_y, *dist, *odom_x, *odom_y, *origin_x, *origin_y, *time, *fitness;
    uint32_t *flags, *ncollide, *ticks;
    unsigned long long *nsamp_sonar, *nsamp_coll;
    uint32_t* ndraws;
};
__global__ void reset_kernel(const ResetParams r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.P) return;
    const int s = (r.n == 1) ? 0 : i;
    const double x = r.x0[s], y = r.y0[s];
    r.x[i] = x; r.y[i] = y; r.theta[i] = r.theta0[s]; r.speed[i] = 0.0; r.omega[i] = 0.0;
    r.last_x[i] = x; r.last_y[i] = y; r.dist[i] = 0.0;
    r.odom_x[i] = x; r.odom_y[i] = y; r.origin_x[i] = x; r.origin_y[i] = y;
    r.time[i] = r.time_budget; r.fitness[i] = 0.0;
    r.flags[i] = 0u; r.ncollide[i] = 0u; r.ticks[i] = 0u;
    r.nsamp_sonar[i] = 0ull; r.nsamp_coll[i] = 0ull;
    r.ndraws[i] = 0u;
    r.goal_tick[i] = 0xffffffffu;
}

// Fresh Sequential_PID_Controllers: zero PID terms, state 0, the configured goal list on the stack (slot n-1 = goal[0]).
struct PidResetParams {
    int P, cap, n_goals, per_robot;
    const double* goals;             // [n_goals][2] or [P][n_goals][2]
    double *acc_ang, *acc_dist, *last_ang, *last_dist, *target;
    int32_t *state, *backing, *ngoals;
    double *gx, *gy;                 // [R2P2_PID_MAX_GOALS][cap]
};
__global__ void pid_reset_kernel(const PidResetParams q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.P) return;
    q.acc_ang[i] = 0.0; q.acc_dist[i] = 0.0; q.last_ang[i] = 0.0; q.last_dist[i] = 0.0; q.target[i] = 0.0;
    q.state[i] = 0; q.backing[i] = 0; q.ngoals[i] = q.n_goals;
    const double* gl = q.goals + (q.per_robot ? (size_t)i * q.n_goals * 2 : 0);
    for (int s = 0; s < q.n_goals; ++s) {
        q.gx[(size_t)s * q.cap + i] = gl[(q.n_goals - 1 - s) * 2];
        q.gy[(size_t)s * q.cap + i] = gl[(q.n_goals - 1 - s) * 2 + 1];
    }
}

__global__ void sum_counters_kernel(const unsigned long long* __restrict__ a, const unsigned long long* __restrict__ b,
                                    const uint32_t* __restrict__ ticks, int P, unsigned long long* __restrict__ out3) {
    unsigned long long sa = 0, sb = 0, st = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) { sa += a[i]; sb += b[i]; st += ticks[i]; }
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out3, sa); atomicAdd(out3 + 1, sb); atomicAdd(out3 + 2, st); }
}

// FP64 issue-rate probe: 8 independent DFMA chains per thread
__global__ void fp64_peak_kernel(double* __restrict__ out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// raw rays: one thread per ray, planes read through L1/L2
__global__ void cast_rays_kernel(StageView sv, int plain, const double* __restrict__ sincostab, int n,
                                 const double* __restrict__ ox, const double* __restrict__ oy, const double* __restrict__ ang,
                                 const double* __restrict__ limit, int32_t* __restrict__ status, double* __restrict__ hx,
                                 double* __restrict__ hy, int32_t* __restrict__ kk, uint32_t* __restrict__ nsamp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ok = 1;
    double vs, vc;
    r2p2_sincos(ang[i], sincostab, &vs, &vc, &ok);
    if (!ok) { status[i] = -1; hx[i] = -1; hy[i] = -1; kk[i] = -1; nsamp[i] = 0; return; }
    const Hit h = plain ? march_plain(sv, ox[i], oy[i], vc, vs, limit[i]) : march(sv, ox[i], oy[i], vc, vs, limit[i]);
    status[i] = h.fault ? -1 : (h.k >= 0 ? 1 : 0);
    hx[i] = h.x; hy[i] = h.y; kk[i] = h.k; nsamp[i] = h.nsamp;
}

// Grid-cost rasterisation for the path planner: path_planning.generate_grid (r2p2/path_planning.py:71-110) + Node.value
// (r2p2/node.py:64).  Cell (j, i) of a div_x x div_y grid takes the grey level at its centre pixel, or 0 if any pixel
// of its int(chunk_h) x int(chunk_w) block is a wall; cost = int(9 - value / 255 * 8), the only place the reference
// uses "darker = more expensive".  One thread per cell, same double arithmetic and truncations as the Python.
__global__ void cost_grid_kernel(const uint8_t* __restrict__ levels, int W, int H, int div_x, int div_y,
                                 int32_t* __restrict__ values, int32_t* __restrict__ costs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= div_x * div_y) return;
    const int j = c % div_x, i = c / div_x;
    const double cw = __ddiv_rn((double)W, (double)div_x), ch = __ddiv_rn((double)H, (double)div_y);
    const double x0 = __dmul_rn((double)j, cw), y0 = __dmul_rn((double)i, ch);
    const int cx = __double2int_rz(__dadd_rn(x0, __ddiv_rn(cw, 2.0))), cy = __double2int_rz(__dadd_rn(y0, __ddiv_rn(ch, 2.0)));
    int value = levels[(size_t)cx * H + cy];
    const int nk = __double2int_rz(ch), nl = __double2int_rz(cw);
    for (int k = 0; k < nk && value != 0; ++k) {
        const int y = __double2int_rz(__dadd_rn(y0, (double)k));
        for (int l = 0; l < nl; ++l) {
            const int x = __double2int_rz(__dadd_rn(x0, (double)l));
            if (levels[(size_t)x * H + y] == 0) { value = 0; break; }
        }
    }
    values[c] = value;
    costs[c] = __double2int_rz(__dsub_rn(9.0, __dmul_rn(__ddiv_rn((double)value, 255.0), 8.0)));
}

// Robot.check_sensors (r2p2/robot.py:287-318) alone, for controllers that run on the host (r2p2_sense): one thread per
// robot walks its fan in ray order -- the noise draw index of a ray is the number of unclamped rays before it -- and
// writes the distances the controller gets.  Poses are not touched; a robot the reference would raise for gets the
// fault flag, as in the episode kernel.
struct SenseParams {
    KParams k;
    double* dst;          // [P][R]
    int32_t* hitpx;       // [P][R][2] hit pixel or -1 (the `col` list of check_sensors), may be NULL
};
__global__ void sense_kernel(const __grid_constant__ SenseParams sp) {
    const KParams& p = sp.k;
    const unsigned rb = blockIdx.x * blockDim.x + threadIdx.x;
    if (rb >= (unsigned)p.P) return;
    unsigned flags = p.flags[rb];
    if (flags & (R2P2_F_FAULT | R2P2_F_DONE)) return;
    StageView sv;
    sv.wall = p.g_wall; sv.lvl50 = p.g_lvl50; sv.dist = p.g_dist; sv.W = p.W; sv.H = p.H; sv.wpr = p.wpr;
    view_global_hot(sv, p.g_hot, p.hwpr, p.hpad);
    const double x = p.x[rb], y = p.y[rb], theta = p.theta[rb];
    bool fault = !(isfinite(x) && isfinite(y));
    const double ox = rint(x), oy = rint(y);
    unsigned ndraws = p.noise_mode ? p.ndraws[rb] : 0u;
    unsigned long long ns = 0;
    for (int i = 0; i < p.R && !fault; ++i) {
        const SinCos sc = sincos_call(R2P2_MUL(R2P2_ADD(p.ray_deg[i], theta), kRad), p.g_sincostab, 4);
        if (!sc.ok) { flags |= R2P2_F_TRIGRANGE; fault = true; break; }
        const Hit h = march(sv, ox, oy, sc.c, sc.s, p.L);
        if (h.fault) { fault = true; break; }
        ns += h.nsamp;
        double d = (h.k >= 0) ? r2p2_norm2(R2P2_SUB(h.x, ox), R2P2_SUB(h.y, oy)) : r2p2_from_bits(0x7ff0000000000000ULL);
        if ((d < p.db_hi && d > p.db_lo) || d > p.L) d = p.L;
        else if (p.noise_mode) d = R2P2_ADD(d, noise_draw(p, rb, ndraws++));
        sp.dst[(size_t)rb * p.R + i] = d;
        if (sp.hitpx) {
            sp.hitpx[((size_t)rb * p.R + i) * 2] = (h.k >= 0) ? dtrunc(h.x) : -1;
            sp.hitpx[((size_t)rb * p.R + i) * 2 + 1] = (h.k >= 0) ? dtrunc(h.y) : -1;
        }
    }
    if (fault) flags |= R2P2_F_FAULT;
    if (p.noise_mode == 1 && ndraws > (unsigned)p.noise_len) flags |= R2P2_F_NOISE;
    p.flags[rb] = flags;
    p.nsamp_sonar[rb] += ns;
    if (p.noise_mode) p.ndraws[rb] = ndraws;
}

}  // namespace r2p2
