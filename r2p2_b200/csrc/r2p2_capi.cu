// r2p2_capi.cu -- C ABI (include/r2p2_b200.h) over the sm_100a kernels in r2p2_kernels.cuh.
//
// Host side of the drop-in boundary: owns the device buffers (bit-packed stage, SoA robot state,
// genomes), picks the lane mapping and launches exactly one episode kernel per r2p2_run.  CUDA runtime
// only -- no torch types, no library kernels, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "r2p2_kernel_ws.cuh"
#include "r2p2_lprtab.h"
#include "r2p2_ga.cuh"

using namespace r2p2;

namespace {
std::string g_create_error;

struct StateArrays {
    double *x = nullptr, *y = nullptr, *theta = nullptr, *speed = nullptr, *omega = nullptr, *last_x = nullptr, *last_y = nullptr,
           *dist = nullptr, *odom_x = nullptr, *odom_y = nullptr, *origin_x = nullptr, *origin_y = nullptr, *time = nullptr,
           *fitness = nullptr;
    uint32_t *flags = nullptr, *ncollide = nullptr, *ticks = nullptr, *ndraws = nullptr, *goal_tick = nullptr;
    unsigned long long *nsamp_sonar = nullptr, *nsamp_coll = nullptr;
};
}  // namespace

struct r2p2_sim {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    bool timing_off = false;         // r2p2_evaluate brackets its chunk launches itself
    cudaStream_t copy_stream = nullptr, aux_stream = nullptr;   // r2p2_evaluate: genome upload in two chunks; the second chunk's kernel
    cudaEvent_t ev_chunk[2] = {nullptr, nullptr}, ev_idle = nullptr, ev_aux = nullptr;
    int slot = 0;                    // launches go to (stream, d_work) [slot 0] or (aux_stream, d_work + 1) [slot 1]
    std::string err;
    int sm_count = 0;
    size_t smem_optin = 0;
    // stage
    int W = 0, H = 0, wpr = 0, has50 = 0;
    uint32_t plane_bytes = 0;
    uint32_t *d_wall = nullptr, *d_l50 = nullptr;
    uint8_t* d_dist = nullptr;       // Chebyshev distance field [H][W]
    uint8_t* d_levels = nullptr;     // grey levels env[x][y] as uploaded (path planner's cost grid)
    uint32_t* d_hot = nullptr;       // padded event plane of the marches (r2p2_kernels.cuh: hot_word / hot_shift), built lazily
    int hot_pad = 0, hot_wpr = 0, hot_rows = 0;
    uint64_t stage_epoch = 0, hot_epoch = 0;
    // robot
    bool have_robot = false;
    int R = 0;
    double ray_deg[R2P2_MAX_RAYS];
    double vr0 = 0, vr1 = 0, radius = 0, max_speed = 0;
    double* d_crxy = nullptr;
    double* d_sincostab = nullptr;
    double* d_sincostab_pad = nullptr;   // rows padded to kSinCosRow doubles (shared-memory copy of the episode kernels)
    // controller
    bool have_ctrl = false;
    int kind = 0, n_layers = 0, activation = 0, maxw = 1;
    int layers[R2P2_MAX_LAYERS];
    // population
    int P = 0, G = 0;
    double* d_genomes = nullptr;     // [P][G]
    double* d_genomes_t = nullptr;   // [ceil(P/32)][G][32] (see transpose_genomes_kernel), built lazily for the 1-lane mapping
    size_t genomes_cap = 0;
    bool genomes_t_valid = false;
    bool have_genomes = false, have_reset = false;
    StateArrays st;
    int state_cap = 0;
    double *d_x0 = nullptr, *d_y0 = nullptr, *d_t0 = nullptr;
    // mapping (the choice is cached until the configuration changes: the per-tick API calls r2p2_run once per frame)
    int lpr_req = 0, smem_req = -1;
    uint64_t cfg_epoch = 1, lpr_epoch = 0;
    int lpr_cached = 0;
    // sensor noise
    int noise_mode = 0, noise_len = 0, noise_P = 0;
    unsigned long long noise_seed = 0;
    uint32_t robot_index0 = 0;       // global index of robot 0 (sharded populations)
    double* d_noise = nullptr;
    // trace
    bool trace = false;
    int trace_T = 0, trace_P = 0, trace_R = 0;
    double *tr_pose = nullptr, *tr_dst = nullptr, *tr_cinfo = nullptr;
    int32_t *tr_hitk = nullptr, *tr_hitpx = nullptr;
    uint32_t *tr_nsamp = nullptr, *tr_ncoll = nullptr;
    // start pose / time budget of the last r2p2_reset with n == 1 (re-applied by r2p2_ga_evaluate)
    bool have_pose1 = false;
    double rx0 = 0, ry0 = 0, rt0 = 0, rbudget = 0;
    // goal flagging
    int goal_on = 0, goal_is_wall = 0;
    double goal_x = 0, goal_y = 0;
    // Sequential_PID_Controller on the device (r2p2_set_pid): goals as given, per-robot controller state
    bool have_pid = false;
    int pid_n_goals = 0, pid_per_robot = 0, pid_goals_P = 0, pid_cap = 0;
    double pid_max_acc = 5.0, pid_accel = 0.0;
    double* d_pid_goals = nullptr;      // [n_goals][2] or [P][n_goals][2]
    double *pid_acc_ang = nullptr, *pid_acc_dist = nullptr, *pid_last_ang = nullptr, *pid_last_dist = nullptr, *pid_target = nullptr;
    int32_t *pid_state = nullptr, *pid_backing = nullptr, *pid_ngoals = nullptr;
    double *pid_gx = nullptr, *pid_gy = nullptr;      // goal stacks [R2P2_PID_MAX_GOALS][pid_cap]
    // host-controller bridge (r2p2_sense / r2p2_move)
    double* d_sense_dst = nullptr;      // [P][R]
    int32_t* d_sense_px = nullptr;      // [P][R][2]
    double* d_cmd = nullptr;            // [P][2]
    int bridge_cap = 0, bridge_R = 0;
    // on-device GA (r2p2_ga.cuh)
    int ga_P = 0, ga_G = 0, ga_cur = 0, ga_tournament = 2, ga_elites = 1;
    double ga_cx = 0.9, ga_mut = 0.05, ga_sigma = 0.1, ga_min = -1.0, ga_max = 1.0;
    unsigned long long ga_seed = 0;
    uint32_t ga_generation = 0;
    double* d_ga_pop[2] = {nullptr, nullptr};
    double* d_ga_fit = nullptr;
    int* d_ga_elite = nullptr;
    // scheduling / counters
    unsigned int* d_work = nullptr;
    uint32_t* d_slice[2] = {nullptr, nullptr};   // time slicing: slices written back per batch of robots, one buffer per launch slot
    size_t slice_cap[2] = {0, 0};
    unsigned long long* d_cnt3 = nullptr;
    uint64_t launches = 0;
};

namespace {

int fail(r2p2_handle h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(h, R2P2_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// The caller's current device is restored when an entry point returns (a torch / multi-GPU host process keeps its own).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(h)                                                                                           \
    DeviceGuard dev_guard_((h)->device);                                                                       \
    if (dev_guard_.err != cudaSuccess) return fail((h), R2P2_E_CUDA, "cudaSetDevice(%d): %s", (h)->device, cudaGetErrorString(dev_guard_.err))

// Scratch device memory of one call: freed on every return path.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T)); }
    operator T*() const { return p; }
};

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

int ensure_state(r2p2_handle h, int P) {
    if (P <= h->state_cap) return R2P2_OK;
    StateArrays& s = h->st;
    double** dbl[] = {&s.x, &s.y, &s.theta, &s.speed, &s.omega, &s.last_x, &s.last_y, &s.dist, &s.odom_x, &s.odom_y,
                      &s.origin_x, &s.origin_y, &s.time, &s.fitness, &h->d_x0, &h->d_y0, &h->d_t0};
    for (double** p : dbl) CK(dalloc(p, (size_t)P));
    CK(dalloc(&s.flags, (size_t)P)); CK(dalloc(&s.ncollide, (size_t)P)); CK(dalloc(&s.ticks, (size_t)P));
    CK(dalloc(&s.nsamp_sonar, (size_t)P)); CK(dalloc(&s.nsamp_coll, (size_t)P)); CK(dalloc(&s.ndraws, (size_t)P)); CK(dalloc(&s.goal_tick, (size_t)P));
    h->state_cap = P;
    h->have_reset = false;
    return R2P2_OK;
}

int ensure_genomes(r2p2_handle h, int P, int G) {
    const size_t need = (size_t)((P + 31) / 32 * 32) * (size_t)G;      // whole tiles of 32 robots for the one-lane copy
    if (need > h->genomes_cap) {
        CK(dalloc(&h->d_genomes, need));
        CK(dalloc(&h->d_genomes_t, need));
        h->genomes_cap = need;
    }
    if (P != h->P) h->have_reset = false;
    if (P != h->P || G != h->G) h->cfg_epoch++;
    h->P = P; h->G = G;
    h->genomes_t_valid = false;
    return ensure_state(h, P);
}

int ensure_pid(r2p2_handle h, int P) {
    if (P <= h->pid_cap) return R2P2_OK;
    double** dbl[] = {&h->pid_acc_ang, &h->pid_acc_dist, &h->pid_last_ang, &h->pid_last_dist, &h->pid_target};
    for (double** p : dbl) CK(dalloc(p, (size_t)P));
    CK(dalloc(&h->pid_state, (size_t)P)); CK(dalloc(&h->pid_backing, (size_t)P)); CK(dalloc(&h->pid_ngoals, (size_t)P));
    CK(dalloc(&h->pid_gx, (size_t)P * R2P2_PID_MAX_GOALS)); CK(dalloc(&h->pid_gy, (size_t)P * R2P2_PID_MAX_GOALS));
    h->pid_cap = P;
    return R2P2_OK;
}

void free_trace(r2p2_handle h) {
    cudaFree(h->tr_pose); cudaFree(h->tr_dst); cudaFree(h->tr_hitk); cudaFree(h->tr_hitpx); cudaFree(h->tr_nsamp); cudaFree(h->tr_ncoll); cudaFree(h->tr_cinfo);
    h->tr_pose = h->tr_dst = h->tr_cinfo = nullptr; h->tr_hitk = h->tr_hitpx = nullptr; h->tr_nsamp = h->tr_ncoll = nullptr;
    h->trace_T = h->trace_P = h->trace_R = 0;
}

int pack_stage(r2p2_handle h, const uint8_t* d_levels, int W, int H) {
    const int wpr = (W + 31) / 32;
    size_t bytes = (size_t)wpr * 4 * (size_t)H;
    bytes = (bytes + 15) & ~(size_t)15;
    CK(dalloc(&h->d_wall, bytes / 4));
    CK(dalloc(&h->d_l50, bytes / 4));
    CK(cudaMemsetAsync(h->d_wall, 0, bytes, h->stream));
    CK(cudaMemsetAsync(h->d_l50, 0, bytes, h->stream));
    DevBuf<int> d_has;
    CK(d_has.alloc(1));
    CK(cudaMemsetAsync(d_has, 0, sizeof(int), h->stream));
    dim3 grid((H + 127) / 128, wpr);
    pack_stage_kernel<<<grid, 128, 0, h->stream>>>(d_levels, W, H, wpr, h->d_wall, h->d_l50, d_has);
    h->launches++;
    CK(cudaGetLastError());
    CK(dalloc(&h->d_levels, (size_t)W * H));
    CK(cudaMemcpyAsync(h->d_levels, d_levels, (size_t)W * H, cudaMemcpyDeviceToDevice, h->stream));
    // distance field (Chebyshev distance to the nearest wall / outer-ring pixel, saturated at 255)
    DevBuf<uint8_t> d_hrow;
    CK(d_hrow.alloc((size_t)W * H));
    CK(dalloc(&h->d_dist, (size_t)W * H));
    dist_rows_kernel<<<(H + 63) / 64, 64, 0, h->stream>>>(d_levels, W, H, d_hrow);
    dist_cols_kernel<<<dim3((W + 127) / 128, H), 128, 0, h->stream>>>(d_hrow, W, H, h->d_dist);
    h->launches += 2;
    CK(cudaGetLastError());
    int has = 0;
    CK(cudaMemcpyAsync(&has, d_has, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->W = W; h->H = H; h->wpr = wpr; h->has50 = has; h->plane_bytes = (uint32_t)bytes;
    h->cfg_epoch++;
    h->stage_epoch++;
    h->goal_on = 0;                     // the goal is judged against the stage it was set on
    return R2P2_OK;
}

// The marches' event plane: the map plus a margin wide enough for every sonar ray of this robot (limit + the overshoot
// of the unrolled walk), rebuilt when the stage or the margin changes.  Collision rays are short (speed * delta +
// radius); one that would not fit the margin takes the exact loop.  No plane above a 1024-pixel margin.
int ensure_hot(r2p2_handle h, double delta) {
    const double L = h->vr1 + h->radius;
    int kB = 16;
    if (L > -1.0 && L < 30000.0) kB = std::max(kB, (int)std::ceil(L + 1e-6));
    const double climit = std::fabs(h->max_speed * delta) + std::fabs(h->radius);
    int half = kB + 34;                                    // see window_geometry: the widest team walks 32 samples per pass
    if (climit < 4096.0) half = std::max(half, (int)std::ceil(climit + 1e-6) + 22);
    int pad = ((half + 8 + 96 + 31) / 32) * 32;
    if (h->d_hot && h->hot_epoch == h->stage_epoch && h->hot_pad >= pad && h->hot_pad <= 2 * pad + 64) return R2P2_OK;   // (a wider margin from an earlier run is as good)
    if (pad > 2048) {
        if (h->d_hot) { CK(cudaStreamSynchronize(h->stream)); cudaFree(h->d_hot); h->d_hot = nullptr; }
        h->hot_pad = 0; h->hot_epoch = h->stage_epoch;
        return R2P2_OK;
    }
    CK(cudaStreamSynchronize(h->stream));
    const int hwpr = ((h->W + 2 * pad + 7) / 8 + 3) / 4 * 4, hrows = (h->H + 2 * pad + 3) / 4;
    CK(dalloc(&h->d_hot, (size_t)hwpr * hrows));
    hot_plane_kernel<<<dim3((hwpr + 127) / 128, hrows), 128, 0, h->stream>>>(h->d_levels, h->W, h->H, pad, hwpr, hrows, h->d_hot);
    h->launches++;
    CK(cudaGetLastError());
    h->hot_pad = pad; h->hot_wpr = hwpr; h->hot_rows = hrows; h->hot_epoch = h->stage_epoch;
    h->cfg_epoch++;
    return R2P2_OK;
}

int upload_crxy(r2p2_handle h) {
    // robot.py:183 `math.cos(math.radians(i)) * self.radius`: per-config constants.  Evaluated on the host with the
    // same glibc-exact sin/cos restatement the device uses, so the numbers do not depend on the host's libm build.
    std::vector<double> t(720);
    for (int i = 0; i < 360; ++i) {
        const double a = (double)i * kRad;
        t[i] = r2p2_cos(a, r2p2_h_sincostab, nullptr) * h->radius;
        t[360 + i] = r2p2_sin(a, r2p2_h_sincostab, nullptr) * h->radius;
    }
    if (!h->d_crxy) CK(dalloc(&h->d_crxy, (size_t)720));
    CK(cudaMemcpyAsync(h->d_crxy, t.data(), 720 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

// occ_out != NULL: only report how many CTAs of the kernel this configuration selects fit on an SM (no launch)
struct OccEntry { const void* fn; int device; size_t smem; int block; int occ; };
std::vector<OccEntry> g_occ_cache;      // per kernel instance: dynamic shared memory it was sized for, CTAs per SM
std::mutex g_occ_mutex;                 // handles of different host threads share the cache

template <int LPR, class KERN>
int launch_kernel(r2p2_handle h, KERN kern, const KParams& kp_in, size_t smem, int* occ_out, bool can_slice = false) {
    KParams kp = kp_in;
    int occ = -1;
    std::lock_guard<std::mutex> lock(g_occ_mutex);
    for (const OccEntry& e : g_occ_cache)
        if (e.fn == (const void*)kern && e.device == h->device && e.smem == smem && e.block == kp.block) { occ = e.occ; break; }
    if (occ < 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kp.block, smem));
        // one entry per (kernel, device): a new size replaces the old one, the attribute is the last one set
        g_occ_cache.erase(std::remove_if(g_occ_cache.begin(), g_occ_cache.end(), [&](const OccEntry& e) { return e.fn == (const void*)kern && e.device == h->device; }), g_occ_cache.end());
        g_occ_cache.push_back({(const void*)kern, h->device, smem, kp.block, occ});
    }
    if (occ_out) { *occ_out = occ; return R2P2_OK; }
    if (occ < 1) return fail(h, R2P2_E_LIMIT, "episode kernel does not fit on an SM (smem %zu bytes)", smem);
    const int robots_per_cta = kp.block / LPR;
    long long ctas = ((long long)(kp.rb_end - kp.rb_begin) + robots_per_cta - 1) / robots_per_cta;
    ctas = std::max<long long>(1, std::min<long long>(ctas, (long long)occ * h->sm_count));   // persistent: the work counter feeds the rest
    cudaStream_t st = h->slot ? h->aux_stream : h->stream;
    // Time slicing (warp-synchronous kernel, see KParams): only when the population runs in several waves, and then fine
    // enough that the launch ends within ~1/48 of its length of the ideal; slices of at least 25 ticks (state and window
    // are re-read per slice).  R2P2_NO_SLICES=1: off (A/B measurements).
    kp.n_slices = 0; kp.slice_T = 0; kp.slice_done = nullptr;
    static const bool no_slices = [] { const char* e = getenv("R2P2_NO_SLICES"); return e && e[0] == '1'; }();
    if (can_slice && !no_slices && !kp.tr_pose && kp.T >= 50) {
        const long long warps = ((long long)(kp.rb_end - kp.rb_begin) * LPR + 31) / 32, cap = (long long)occ * h->sm_count * (kp.block / 32);
        if (warps > cap) {
            const double waves = (double)warps / (double)cap;
            int n = std::min((int)std::ceil(48.0 / waves), kp.T / 25);
            if (n >= 2) {
                kp.slice_T = (kp.T + n - 1) / n;
                kp.n_slices = (kp.T + kp.slice_T - 1) / kp.slice_T;
                const size_t batches = ((size_t)(kp.rb_end - kp.rb_begin) + (32 / LPR) - 1) / (32 / LPR);
                if (batches > h->slice_cap[h->slot]) { CK(dalloc(&h->d_slice[h->slot], batches)); h->slice_cap[h->slot] = batches; }
                kp.slice_done = h->d_slice[h->slot];
                CK(cudaMemsetAsync(kp.slice_done, 0, batches * sizeof(uint32_t), st));
                ctas = std::min<long long>((long long)occ * h->sm_count, ctas);
            }
        }
    }
    CK(cudaMemsetAsync(kp.work_counter, 0, sizeof(unsigned int), st));
    if (!h->timing_off) CK(cudaEventRecord(h->ev0, st));
    kern<<<(unsigned)ctas, kp.block, smem, st>>>(kp);
    if (!h->timing_off) { CK(cudaEventRecord(h->ev1, st)); h->ev_valid = true; }
    h->launches++;
    CK(cudaGetLastError());
    return R2P2_OK;
}

// 2..16 lanes per robot run the warp-synchronous kernel (r2p2_kernel_ws.cuh): same results, one instruction stream per
// warp.  R2P2_WS=0 in the environment selects the per-group kernel for A/B measurements.  A sonar limit without a norm
// window sends whole groups through the plain loop, which the lock-step kernel cannot mix with groups that have
// stopped; r2p2_move's skip_sense lives in the per-group kernel only.
bool ws_allowed(const KParams& kp) {
    static const bool ws = [] { const char* e = getenv("R2P2_WS"); return !(e && e[0] == '0'); }();
    return ws && !kp.skip_sense && kp.L > -1.0 && kp.L < 30000.0 && kp.g_hot && kp.win_ww > 0;
}

// Window geometry of the warp-synchronous kernel for `lpr` lanes per robot (same numbers as the kernel: sonar team size,
// own samples per pass).  half = the farthest pixel, counted from the sonar origin, a fast march of the tick can read.
void window_geometry(KParams& kp, int lpr) {
    kp.win_ww = kp.win_wb = kp.win_half = kp.win_slack = 0;
    static const bool no_win = [] { const char* e = getenv("R2P2_NO_WIN"); return e && e[0] == '1'; }();   // A/B measurements
    if (lpr < 2 || lpr > 16 || !kp.g_hot || no_win || !(kp.L > -1.0 && kp.L < 30000.0)) return;
    int tl = 1;
    while (tl * 2 * kp.R <= lpr && tl < kMaxSonarTeam) tl *= 2;
    const int own_s = sonar_nown(tl, lpr <= 2 ? kWsSonarOwn<2> : kWsSonarOwn<8>, lpr <= 2 ? kWsTeamOwn<2> : kWsTeamOwn<8>), own_c = (lpr >= 16 ? 1 : lpr >= 8 ? 2 : 4);
    const int kB_s = (int)std::ceil(kp.L + 1e-6);
    const double climit = std::fabs(kp.max_speed * kp.delta) + std::fabs(kp.radius);
    if (!(climit < 4096.0)) return;
    const int kB_c = (int)std::ceil(climit + 1e-6) + 1;
    const int half = std::max(std::max(kB_s + tl * own_s + 1, kB_c + lpr * own_c + 3), (int)std::ceil(std::fabs(kp.radius)) + 4) + 1;
    static const int slack_env = [] { const char* e = getenv("R2P2_WIN_SLACK"); return e ? atoi(e) : 0; }();   // experiments
    const int slack = slack_env > 0 ? slack_env : 8;       // a fetched window is kept until the robot has moved this far
    const int ww = (((2 * (half + slack) + 32 + 7) / 8) + 3) / 4 * 4, wb = (2 * (half + slack) + 4 + 3) / 4;
    if (half + slack + 96 > kp.hpad) return;                                             // the window must stay inside the plane's margin
    if ((size_t)(kp.block / lpr) * ww * wb * 4 * 128 > (size_t)64 * 1024 * kp.block) return;                 // too fat: the per-group kernel reads the plane through L1
    kp.win_ww = ww; kp.win_wb = wb; kp.win_half = half; kp.win_slack = slack;
}

template <int LPR, bool SMEM, class NET>
int launch_episode(r2p2_handle h, const KParams& kp, size_t smem, int* occ_out) {
    // 2..16 lanes per robot: the warp-synchronous kernel (r2p2_kernel_ws.cuh), same results, one instruction stream per
    // warp.  R2P2_WS=0 in the environment selects the per-group kernel for A/B measurements.
    const bool extra = kp.goal_on || kp.skip_sense || kp.ctrl_kind == R2P2_CTRL_PID;     // rarely used switches live in their own instances
    if constexpr (LPR >= 2 && LPR <= 16 && !SMEM) {
        if (ws_allowed(kp))
            return extra ? launch_kernel<LPR>(h, episode_kernel_ws<LPR, NET, true>, kp, smem, occ_out)
                         : launch_kernel<LPR>(h, episode_kernel_ws<LPR, NET, false>, kp, smem, occ_out, true);
    }
    if constexpr (!NET::kSmemWeights || LPR == 1) {
        return extra ? launch_kernel<LPR>(h, episode_kernel<LPR, SMEM, NET, true>, kp, smem, occ_out)
                     : launch_kernel<LPR>(h, episode_kernel<LPR, SMEM, NET, false>, kp, smem, occ_out);
    } else {
        return fail(h, R2P2_E_ARG, "internal: shared-memory-weight instance outside the warp-synchronous kernel");
    }
}

template <class NET>
bool net_matches(const KParams& kp) {
    if (kp.ctrl_kind != R2P2_CTRL_NEURO || kp.n_layers != NET::kLayers) return false;
    for (int l = 0; l < NET::kLayers; ++l)
        if (kp.layers[l] != NET::sz(l)) return false;
    return true;
}

template <int LPR>
int launch_episode_s(r2p2_handle h, KParams& kp, bool smem_stage, int* occ_out) {
    // the reference's shipped network shapes have fixed-shape instances (weights in registers; chosen unless the shared-memory
    // stage copy is requested, which keeps the run-time-shaped path reachable for tests); LPR == 1 would need
    // the whole net in one thread's registers, so it stays on the generic path
    // LPR >= 16: weights in registers (latency-bound runs).  Below that the extra registers cost occupancy (measured:
    // LPR 8 with register weights 25 ms vs 17 ms generic on 8192 robots; LPR 1 fully unrolled = 255 registers + spills,
    // 148 ms vs 81 ms on 65536 robots), so those mappings stay on the run-time-shaped path.
    if constexpr (LPR >= 16) if (!smem_stage) {
        if (net_matches<NetNeuro>(kp)) {
            kp.w_in_smem = 0;
            return launch_episode<LPR, false, NetNeuro>(h, kp, episode_smem_bytes(kp.block, LPR, false, kp.plane_bytes, kp.has50, kp.maxw, 0, kp.G, kp.win_ww * kp.win_wb, kp.tanh_in_smem), occ_out);
        }
        if (net_matches<NetSimple>(kp)) {
            kp.w_in_smem = 0;
            return launch_episode<LPR, false, NetSimple>(h, kp, episode_smem_bytes(kp.block, LPR, false, kp.plane_bytes, kp.has50, kp.maxw, 0, kp.G, kp.win_ww * kp.win_wb, kp.tanh_in_smem), occ_out);
        }
    }
    // one lane per robot: compile-time shape over the transposed genome array
    if constexpr (LPR == 1) if (!smem_stage) {
        const size_t sm = episode_smem_bytes(kp.block, 1, false, kp.plane_bytes, kp.has50, kp.maxw, 0, kp.G, 0, kp.tanh_in_smem);
        if (net_matches<NetNeuro>(kp)) return launch_episode<1, false, SmemNet<NetNeuro>>(h, kp, sm, occ_out);
        if (net_matches<NetSimple>(kp)) return launch_episode<1, false, SmemNet<NetSimple>>(h, kp, sm, occ_out);
    }
    // 2..8 lanes: compile-time shape, weights read through one pointer per robot -- its genome's copy in shared memory
    // when the robots in flight fit there (w_in_smem), else the [P][G] array itself (L1/L2)
    if constexpr (LPR >= 2 && LPR <= 8) if (!smem_stage && ws_allowed(kp)) {
        const size_t sm = episode_smem_bytes(kp.block, LPR, false, kp.plane_bytes, kp.has50, kp.maxw, kp.w_in_smem, kp.G, kp.win_ww * kp.win_wb, kp.tanh_in_smem);
        if (net_matches<NetNeuro>(kp)) return launch_episode<LPR, false, SmemNet<NetNeuro>>(h, kp, sm, occ_out);
        if (net_matches<NetSimple>(kp)) return launch_episode<LPR, false, SmemNet<NetSimple>>(h, kp, sm, occ_out);
        if constexpr (LPR >= 4) if (net_matches<NetSweep>(kp) && !(kp.goal_on || kp.skip_sense))
            return launch_kernel<LPR>(h, episode_kernel_ws<LPR, SmemNet<NetSweep>, false>, kp, sm, occ_out, true);
    }
    const size_t smem = episode_smem_bytes(kp.block, LPR, smem_stage, kp.plane_bytes, kp.has50, kp.maxw, kp.w_in_smem, kp.G, kp.win_ww * kp.win_wb, kp.tanh_in_smem);
    return smem_stage ? launch_episode<LPR, true, GenericNet>(h, kp, smem, occ_out) : launch_episode<LPR, false, GenericNet>(h, kp, smem, occ_out);
}

// genomes of the robots in flight live in shared memory when the group mapping leaves room for them
int w_in_smem_for(int block, int lpr, int G) {
    // R2P2_WSMEM_KB (experiments): largest genome block a CTA may hold in shared memory; two CTAs per SM must still fit
    static const size_t lim = [] { const char* e = getenv("R2P2_WSMEM_KB"); return (size_t)(e ? atoi(e) : 48) * 1024; }();
    return (lpr > 1 && (size_t)(block / lpr) * G * 8 * 128 <= lim * block) ? 1 : 0;
}

// Shared-memory plan of one launch for a given CTA size: windows first, then the genomes of the robots in flight, then
// the tanh table -- each only while 16 warps per SM (512 / block CTAs of the 128-register instances) still fit.
void plan_smem(KParams& kp, int lpr, bool smem_stage) {
    const size_t per_cta = (size_t)225 * 1024 / std::max(1, kMaxBlock / kp.block);   // what one CTA may take with kMaxBlock / block CTAs per SM (one CTA for 448 or 512 threads)
    kp.w_in_smem = w_in_smem_for(kp.block, lpr, kp.G);
    window_geometry(kp, lpr);
    if (kp.w_in_smem && kp.win_ww && episode_smem_bytes(kp.block, lpr, false, kp.plane_bytes, kp.has50, kp.maxw, 1, kp.G, kp.win_ww * kp.win_wb) > per_cta)
        kp.w_in_smem = 0;
    // the tanh table (18 KB): rows are picked per lane, so a global read is up to 32 L1 lookups per instruction against a
    // few shared-memory wavefronts
    static const bool no_tanh_smem = [] { const char* e = getenv("R2P2_NO_TANH_SMEM"); return e && e[0] == '1'; }();
    kp.tanh_in_smem = (!no_tanh_smem && !smem_stage && kp.ctrl_kind == R2P2_CTRL_NEURO && kp.activation == R2P2_ACT_TANH &&
                       episode_smem_bytes(kp.block, lpr, false, kp.plane_bytes, kp.has50, kp.maxw, kp.w_in_smem, kp.G, kp.win_ww * kp.win_wb, 1) <= per_cta) ? 1 : 0;
}

// Threads per CTA.  The register-weight instances run 128.  The others hold 16 warps per SM whatever the split, so the
// choice is about (a) how the population spreads over the SMs -- the run lasts as long as its fullest SM, so the split
// with the fewest warps on the fullest SM wins -- and (b) at equal spread, the largest CTA: its tables are staged once
// per SM and the shared memory that saves is L1 for the genomes (R2P2_BLOCK in the environment forces a size).
// A population that is resident all at once is a static partition (every robot runs its T ticks on the SM it landed on):
// 65,536 one-lane robots are 2,048 warps = 13.8 per SM, and 128 CTAs of 512 threads would leave 20 SMs idle with 16 warps
// on each of the others.  One CTA per SM of 32 * ceil(warps / SMs) threads (here 448: 14 warps on 147 SMs) spreads them
// evenly and still stages the tables once per SM.
int choose_block(r2p2_handle h, const KParams& kp, int lpr, bool smem_stage) {
    const bool regw = lpr >= 16 && !smem_stage && (net_matches<NetNeuro>(kp) || net_matches<NetSimple>(kp));
    if (regw) return kRegBlock;
    static const int forced = [] { const char* e = getenv("R2P2_BLOCK"); const int v = e ? atoi(e) : 0; return (v >= 128 && v <= kMaxBlock && v % 32 == 0) ? v : 0; }();
    if (forced) return forced;
    int best = 128;
    long long best_warps = -1;
    auto smem_of = [&](int block) {
        KParams k2 = kp;
        k2.block = block;
        plan_smem(k2, lpr, smem_stage);
        return episode_smem_bytes(block, lpr, smem_stage, k2.plane_bytes, k2.has50, k2.maxw, k2.w_in_smem, k2.G, k2.win_ww * k2.win_wb, k2.tanh_in_smem);
    };
    for (int block : {kMaxBlock, 256, 128}) {
        const size_t smem = smem_of(block);
        if (smem > h->smem_optin || smem * (kMaxBlock / block) > (size_t)225 * 1024) continue;   // 16 warps per SM must fit
        const long long ctas = ((long long)(kp.rb_end - kp.rb_begin) * lpr + block - 1) / block;
        const long long per_sm = std::min<long long>((ctas + h->sm_count - 1) / h->sm_count, kMaxBlock / block);
        const long long warps = per_sm * (block / 32);
        if (best_warps < 0 || warps < best_warps) { best_warps = warps; best = block; }
    }
    static const bool no_bal = [] { const char* e = getenv("R2P2_NO_BALANCE"); return e && e[0] == '1'; }();   // A/B measurements
    const long long total_warps = ((long long)(kp.rb_end - kp.rb_begin) * lpr + 31) / 32;
    const long long bal = 32 * ((total_warps + h->sm_count - 1) / h->sm_count);        // one CTA per SM, everything resident
    if (!no_bal && bal >= 128 && bal < kMaxBlock && (best_warps < 0 || bal / 32 < best_warps) && smem_of((int)bal) <= h->smem_optin) best = (int)bal;
    return best;
}

int dispatch_episode(r2p2_handle h, KParams& kp, int lpr, bool smem_stage, int* occ_out) {
    kp.block = choose_block(h, kp, lpr, smem_stage);
    plan_smem(kp, lpr, smem_stage);
    switch (lpr) {
        case 1: return launch_episode_s<1>(h, kp, smem_stage, occ_out);
        case 2: return launch_episode_s<2>(h, kp, smem_stage, occ_out);
        case 4: return launch_episode_s<4>(h, kp, smem_stage, occ_out);
        case 8: return launch_episode_s<8>(h, kp, smem_stage, occ_out);
        case 16: return launch_episode_s<16>(h, kp, smem_stage, occ_out);
        default: return launch_episode_s<32>(h, kp, smem_stage, occ_out);
    }
}

// Lanes per robot.  From 768 robots on: the fastest mapping of the measured sweep (tools/sweep_lpr.py on a B200 ->
// profiles/r02_lpr_sweep.json -> r2p2_lprtab.h), row by ray count, column by the population's power of two.  The rule it
// replaces -- the widest group whose whole population is resident at once -- agrees with the sweep within 3 % except where
// it picked 2 lanes (32,768 five-ray robots: 1 lane is 1.3x faster; 16,384 / 32,768 robots with 32 rays: 8 lanes are
// 1.7x / 1.2x faster); it still serves small populations (always resident: a warp per robot) and mappings the table
// names but this configuration cannot launch.
int choose_lpr(r2p2_handle h, const KParams& kp) {
    if (h->lpr_req) return h->lpr_req;
    if (h->P >= 768) {
        const LprRow* row = &kLprTable[0];
        for (const LprRow& r : kLprTable) { row = &r; if (h->R <= r.max_rays) break; }
        int col = 0;
        while (col < 7 && (double)h->P > 1024.0 * std::pow(2.0, col + 0.5)) ++col;       // nearest power of two (log scale)
        const int lpr = row->best[col];
        KParams k2 = kp;
        int occ = 0;
        if (dispatch_episode(h, k2, lpr, false, &occ) == R2P2_OK && occ >= 1) return lpr;
    }
    for (int lpr = 32; lpr > 1; lpr >>= 1) {
        KParams k2 = kp;
        int occ = 0;
        if (dispatch_episode(h, k2, lpr, false, &occ) != R2P2_OK) continue;
        const long long ctas = ((long long)h->P * lpr + k2.block - 1) / k2.block;
        if (ctas <= (long long)occ * h->sm_count) return lpr;
    }
    return h->R >= 16 ? 8 : 1;
}

}  // namespace

extern "C" {

int r2p2_version(void) { return 100; }

const char* r2p2_last_error(r2p2_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int r2p2_create(int device, r2p2_handle* out) {
    if (!out) return fail(nullptr, R2P2_E_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, R2P2_E_NODEV, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, R2P2_E_ARG, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, R2P2_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, R2P2_E_NODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    r2p2_sim* s = new r2p2_sim();
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    s->smem_optin = prop.sharedMemPerBlockOptin;
    DeviceGuard guard(device);                             // the caller's current device comes back when this returns
    e = guard.err;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    s->stream = s->own_stream;
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_work, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_cnt3, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_sincostab, R2P2_SINCOSTAB_LEN * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_sincostab, r2p2_h_sincostab, R2P2_SINCOSTAB_LEN * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_sincostab_pad, kSinCosPadLen * sizeof(double));
    if (e == cudaSuccess) {
        std::vector<double> pad(kSinCosPadLen, 0.0);
        for (int r = 0; r < R2P2_SINCOSTAB_LEN / 4; ++r)
            for (int j = 0; j < 4; ++j) pad[(size_t)r * kSinCosRow + j] = r2p2_h_sincostab[r * 4 + j];
        e = cudaMemcpy(s->d_sincostab_pad, pad.data(), kSinCosPadLen * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        fail(nullptr, R2P2_E_CUDA, "r2p2_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        r2p2_destroy(s);                                   // frees whatever was created (stream, events, buffers)
        return R2P2_E_CUDA;
    }
    *out = s;
    return R2P2_OK;
}

int r2p2_destroy(r2p2_handle h) {
    if (!h) return R2P2_OK;
    DeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    StateArrays& s = h->st;
    void* ptrs[] = {s.x, s.y, s.theta, s.speed, s.omega, s.last_x, s.last_y, s.dist, s.odom_x, s.odom_y, s.origin_x, s.origin_y,
                    s.time, s.fitness, s.flags, s.ncollide, s.ticks, s.ndraws, s.goal_tick, s.nsamp_sonar, s.nsamp_coll, h->d_noise, h->d_x0, h->d_y0, h->d_t0,
                    h->d_wall, h->d_l50, h->d_dist, h->d_crxy, h->d_sincostab, h->d_sincostab_pad, h->d_genomes, h->d_genomes_t, h->d_work, h->d_cnt3,
                    h->d_ga_pop[0], h->d_ga_pop[1], h->d_ga_fit, h->d_ga_elite, h->d_sense_dst, h->d_sense_px, h->d_cmd, h->d_levels, h->d_hot,
                    h->d_slice[0], h->d_slice[1], h->d_pid_goals, h->pid_acc_ang, h->pid_acc_dist, h->pid_last_ang, h->pid_last_dist, h->pid_target, h->pid_state, h->pid_backing,
                    h->pid_ngoals, h->pid_gx, h->pid_gy};
    for (void* p : ptrs) cudaFree(p);
    free_trace(h);
    for (cudaEvent_t e : h->ev_chunk) if (e) cudaEventDestroy(e);
    if (h->ev_idle) cudaEventDestroy(h->ev_idle);
    if (h->ev_aux) cudaEventDestroy(h->ev_aux);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return R2P2_OK;
}

int r2p2_set_stream(r2p2_handle h, void* cuda_stream) {
    if (!h) return R2P2_E_ARG;
    ON_DEVICE(h);
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return R2P2_OK;
}

int r2p2_fp64_peak(r2p2_handle h, double* dfma_per_s) {
    if (!h || !dfma_per_s) return R2P2_E_ARG;
    ON_DEVICE(h);
    const int blocks = h->sm_count * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> d;
    CK(d.alloc((size_t)blocks * threads));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(h->ev0, h->stream));
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(d, iters);
        CK(cudaEventRecord(h->ev1, h->stream));
        h->launches++;
        CK(cudaEventSynchronize(h->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    h->ev_valid = false;
    *dfma_per_s = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
    return R2P2_OK;
}

int r2p2_set_stage(r2p2_handle h, const uint8_t* levels_xy, int32_t W, int32_t H) {
    if (!h) return R2P2_E_ARG;
    if (!levels_xy || W < 3 || H < 3) return fail(h, R2P2_E_ARG, "stage must be at least 3x3");
    if (W > 32768 || H > 32768) return fail(h, R2P2_E_LIMIT, "stage larger than 32768 px per side");
    ON_DEVICE(h);
    DevBuf<uint8_t> d;
    CK(d.alloc((size_t)W * H));
    CK(cudaMemcpyAsync(d, levels_xy, (size_t)W * H, cudaMemcpyHostToDevice, h->stream));
    return pack_stage(h, d, W, H);
}

int r2p2_set_stage_i32(r2p2_handle h, const int32_t* levels_xy, int32_t W, int32_t H) {
    if (!h) return R2P2_E_ARG;
    if (!levels_xy || W < 3 || H < 3) return fail(h, R2P2_E_ARG, "stage must be at least 3x3");
    if (W > 32768 || H > 32768) return fail(h, R2P2_E_LIMIT, "stage larger than 32768 px per side");
    ON_DEVICE(h);
    const size_t n = (size_t)W * H;
    DevBuf<int32_t> d32; DevBuf<uint8_t> d8; DevBuf<int> d_bad;
    CK(d32.alloc(n)); CK(d8.alloc(n)); CK(d_bad.alloc(1));
    CK(cudaMemsetAsync(d_bad, 0, 4, h->stream));
    CK(cudaMemcpyAsync(d32, levels_xy, n * 4, cudaMemcpyHostToDevice, h->stream));
    i32_to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d32, d8, n, d_bad);
    h->launches++;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return bad ? fail(h, R2P2_E_ARG, "stage levels must be grey values 0..255") : pack_stage(h, d8, W, H);
}

int r2p2_set_robot(r2p2_handle h, double vr0, double vr1, double radius, double max_speed, const double* ray_deg, int32_t n_rays) {
    if (!h) return R2P2_E_ARG;
    if (!ray_deg || n_rays < 1) return fail(h, R2P2_E_ARG, "need at least one sonar ray");
    if (n_rays > R2P2_MAX_RAYS) return fail(h, R2P2_E_LIMIT, "at most %d sonar rays", R2P2_MAX_RAYS);
    ON_DEVICE(h);
    h->vr0 = vr0; h->vr1 = vr1; h->radius = radius; h->max_speed = max_speed; h->R = n_rays;
    memcpy(h->ray_deg, ray_deg, sizeof(double) * (size_t)n_rays);
    h->have_robot = true;
    h->cfg_epoch++;
    return upload_crxy(h);
}

int r2p2_set_controller(r2p2_handle h, int32_t kind, const int32_t* layers, int32_t n_layers, int32_t activation) {
    if (!h) return R2P2_E_ARG;
    if (kind < R2P2_CTRL_CONST || kind > R2P2_CTRL_PID) return fail(h, R2P2_E_ARG, "unknown controller kind %d", kind);
    if (activation < 0 || activation > R2P2_ACT_LOGISTIC) return fail(h, R2P2_E_ARG, "unknown activation %d", activation);
    h->maxw = 1;
    if (kind == R2P2_CTRL_NEURO) {
        if (!layers || n_layers < 2 || n_layers > R2P2_MAX_LAYERS) return fail(h, R2P2_E_LIMIT, "MLP needs 2..%d layer sizes", R2P2_MAX_LAYERS);
        for (int i = 0; i < n_layers; ++i) {
            if (layers[i] < 1 || layers[i] > R2P2_MAX_WIDTH) return fail(h, R2P2_E_LIMIT, "layer width must be 1..%d", R2P2_MAX_WIDTH);
            h->layers[i] = layers[i];
            h->maxw = std::max(h->maxw, layers[i]);
        }
        if (layers[n_layers - 1] != 2) return fail(h, R2P2_E_ARG, "the MLP must end in 2 outputs (speed, angular velocity)");
        h->n_layers = n_layers;
    } else {
        h->n_layers = 0;
    }
    h->kind = kind; h->activation = activation; h->have_ctrl = true;
    h->cfg_epoch++;
    return R2P2_OK;
}

int r2p2_set_noise(r2p2_handle h, int32_t mode, uint64_t seed, const double* stream, int32_t P, int32_t len) {
    if (!h) return R2P2_E_ARG;
    if (mode < 0 || mode > 2) return fail(h, R2P2_E_ARG, "noise mode must be 0 (off), 1 (host stream) or 2 (device generator)");
    ON_DEVICE(h);
    if (mode == 1) {
        if (!stream || P < 1 || len < 1) return fail(h, R2P2_E_ARG, "noise mode 1 needs a stream [P][len]");
        CK(cudaStreamSynchronize(h->stream));
        CK(dalloc(&h->d_noise, (size_t)P * len));
        CK(cudaMemcpyAsync(h->d_noise, stream, sizeof(double) * (size_t)P * len, cudaMemcpyHostToDevice, h->stream));
        if (h->st.ndraws && P <= h->state_cap) CK(cudaMemsetAsync(h->st.ndraws, 0, sizeof(uint32_t) * P, h->stream));   // a new stream is read from its start
        CK(cudaStreamSynchronize(h->stream));
        h->noise_P = P; h->noise_len = len;
    }
    h->noise_mode = mode; h->noise_seed = seed;
    return R2P2_OK;
}

int r2p2_set_robot_index_offset(r2p2_handle h, uint32_t first) {
    if (!h) return R2P2_E_ARG;
    h->robot_index0 = first;
    return R2P2_OK;
}

int r2p2_population_size(r2p2_handle h, int32_t* P, int32_t* G) {
    if (!h) return R2P2_E_ARG;
    if (P) *P = h->P;
    if (G) *G = h->G;
    return R2P2_OK;
}

int r2p2_write_pose(r2p2_handle h, int32_t first, int32_t count, const double* x, const double* y, const double* theta,
                    const double* last_x, const double* last_y) {
    if (!h) return R2P2_E_ARG;
    if (h->P < 1 || !h->have_reset) return fail(h, R2P2_E_ARG, "r2p2_write_pose: population not initialised (r2p2_reset first)");
    if (first < 0 || count < 0 || first + count > h->P) return fail(h, R2P2_E_ARG, "r2p2_write_pose: robots [%d, %d) out of range (P = %d)", first, first + count, h->P);
    ON_DEVICE(h);
    StateArrays& s = h->st;
    const size_t nb = sizeof(double) * (size_t)count;
    if (x) CK(cudaMemcpyAsync(s.x + first, x, nb, cudaMemcpyHostToDevice, h->stream));
    if (y) CK(cudaMemcpyAsync(s.y + first, y, nb, cudaMemcpyHostToDevice, h->stream));
    if (theta) CK(cudaMemcpyAsync(s.theta + first, theta, nb, cudaMemcpyHostToDevice, h->stream));
    if (last_x) CK(cudaMemcpyAsync(s.last_x + first, last_x, nb, cudaMemcpyHostToDevice, h->stream));
    if (last_y) CK(cudaMemcpyAsync(s.last_y + first, last_y, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));              // the host arrays may be pageable and reused by the caller
    return R2P2_OK;
}

int r2p2_read_noise_draws(r2p2_handle h, uint32_t* ndraws) {
    if (!h || !ndraws) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(ndraws, h->st.ndraws, sizeof(uint32_t) * h->P, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_set_mapping(r2p2_handle h, int32_t lanes_per_robot, int32_t stage_in_smem) {
    if (!h) return R2P2_E_ARG;
    const int l = lanes_per_robot;
    if (!(l == 0 || l == 1 || l == 2 || l == 4 || l == 8 || l == 16 || l == 32)) return fail(h, R2P2_E_ARG, "lanes_per_robot must be 0,1,2,4,8,16,32");
    if (stage_in_smem < -1 || stage_in_smem > 1) return fail(h, R2P2_E_ARG, "stage_in_smem must be -1, 0 or 1");
    h->lpr_req = l; h->smem_req = stage_in_smem;
    h->cfg_epoch++;
    return R2P2_OK;
}

int r2p2_upload_genomes(r2p2_handle h, const double* genomes, int32_t P, int32_t G) {
    if (!h) return R2P2_E_ARG;
    if (!genomes || P < 1 || G < 1) return fail(h, R2P2_E_ARG, "genomes: need P >= 1, G >= 1");
    ON_DEVICE(h);
    int rc = ensure_genomes(h, P, G);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_genomes, genomes, (size_t)P * G * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->have_genomes = true;
    return R2P2_OK;
}

int r2p2_set_genomes_device(r2p2_handle h, const double* d_genomes, int32_t P, int32_t G) {
    if (!h) return R2P2_E_ARG;
    if (!d_genomes || P < 1 || G < 1) return fail(h, R2P2_E_ARG, "genomes: need P >= 1, G >= 1");
    ON_DEVICE(h);
    int rc = ensure_genomes(h, P, G);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_genomes, d_genomes, (size_t)P * G * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    h->have_genomes = true;
    return R2P2_OK;
}

int r2p2_reset(r2p2_handle h, const double* x0, const double* y0, const double* theta0, int32_t n, double time_budget) {
    if (!h) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "upload genomes (population size) before r2p2_reset");
    if (!x0 || !y0 || !theta0 || !(n == 1 || n == h->P)) return fail(h, R2P2_E_ARG, "reset: n must be 1 or P");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(h->d_x0, x0, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_y0, y0, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_t0, theta0, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    ResetParams r;
    StateArrays& s = h->st;
    r.P = h->P; r.n = n; r.x0 = h->d_x0; r.y0 = h->d_y0; r.theta0 = h->d_t0; r.time_budget = time_budget;
    r.x = s.x; r.y = s.y; r.theta = s.theta; r.speed = s.speed; r.omega = s.omega; r.last_x = s.last_x; r.last_y = s.last_y;
    r.dist = s.dist; r.odom_x = s.odom_x; r.odom_y = s.odom_y; r.origin_x = s.origin_x; r.origin_y = s.origin_y;
    r.time = s.time; r.fitness = s.fitness; r.flags = s.flags; r.ncollide = s.ncollide; r.ticks = s.ticks;
    r.nsamp_sonar = s.nsamp_sonar; r.nsamp_coll = s.nsamp_coll; r.ndraws = s.ndraws; r.goal_tick = s.goal_tick;
    reset_kernel<<<(h->P + 255) / 256, 256, 0, h->stream>>>(r);
    h->launches++;
    CK(cudaGetLastError());
    if (h->have_pid) {                                 // fresh controllers: the goal list as configured, PID terms zero
        if (h->pid_per_robot && h->pid_goals_P != h->P)
            return fail(h, R2P2_E_ARG, "r2p2_set_pid gave goals for %d robots, the population is %d", h->pid_goals_P, h->P);
        const int rc = ensure_pid(h, h->P);
        if (rc) return rc;
        PidResetParams q;
        q.P = h->P; q.cap = h->P; q.n_goals = h->pid_n_goals; q.per_robot = h->pid_per_robot; q.goals = h->d_pid_goals;
        q.acc_ang = h->pid_acc_ang; q.acc_dist = h->pid_acc_dist; q.last_ang = h->pid_last_ang; q.last_dist = h->pid_last_dist; q.target = h->pid_target;
        q.state = h->pid_state; q.backing = h->pid_backing; q.ngoals = h->pid_ngoals; q.gx = h->pid_gx; q.gy = h->pid_gy;
        pid_reset_kernel<<<(h->P + 255) / 256, 256, 0, h->stream>>>(q);
        h->launches++;
        CK(cudaGetLastError());
    }
    if (n > 1) CK(cudaStreamSynchronize(h->stream));   // host arrays may be pageable and reused by the caller
    h->have_reset = true;
    if (n == 1) { h->have_pose1 = true; h->rx0 = x0[0]; h->ry0 = y0[0]; h->rt0 = theta0[0]; h->rbudget = time_budget; }
    return R2P2_OK;
}

static int fill_params(r2p2_handle h, KParams& kp, int32_t T, double delta, double delta_ctrl, int32_t evolve) {
    memset(&kp, 0, sizeof kp);
    kp.W = h->W; kp.H = h->H; kp.wpr = h->wpr; kp.has50 = h->has50;
    kp.g_wall = h->d_wall; kp.g_lvl50 = h->d_l50; kp.plane_bytes = h->plane_bytes;
    kp.g_dist = h->d_dist;
    {
        const int rc = ensure_hot(h, delta);
        if (rc) return rc;
        static const bool no_hot = [] { const char* e = getenv("R2P2_NO_HOT"); return e && e[0] == '1'; }();   // A/B measurements
        if (h->d_hot && !no_hot) { kp.g_hot = h->d_hot + (size_t)(h->hot_pad / 4) * h->hot_wpr + h->hot_pad / 8; kp.hwpr = h->hot_wpr; kp.hpad = h->hot_pad; }
    }
    kp.crash_clear = (int)std::ceil(std::fabs(h->radius)) + 2;
    if (!(h->radius >= 0.0 && h->radius < 250.0)) kp.crash_clear = 1 << 20;   // never skip the ring
    kp.R = h->R;
    memcpy(kp.ray_deg, h->ray_deg, sizeof(double) * (size_t)h->R);
    kp.vr0 = h->vr0; kp.vr1 = h->vr1; kp.radius = h->radius; kp.max_speed = h->max_speed;
    kp.L = h->vr1 + h->radius;
    {
        const volatile double m = h->vr0 * 0.125;   // two roundings, as Python evaluates vr0*0.125 + radius
        kp.db_lo = m + h->radius;
    }
    kp.db_hi = h->vr0 + h->radius;
    kp.rad_p1 = h->radius + 1.0;
    kp.ctrl_kind = h->kind; kp.n_layers = h->n_layers; kp.activation = h->activation; kp.G = h->G;
    kp.maxw = std::max(h->maxw, h->R);
    memcpy(kp.layers, h->layers, sizeof(int) * R2P2_MAX_LAYERS);
    kp.P = h->P; kp.T = T; kp.evolve = evolve ? 1 : 0;
    kp.rb_begin = 0u; kp.rb_end = (uint32_t)h->P;
    kp.delta = delta; kp.delta_ctrl = delta_ctrl;
    kp.g_sincostab = h->d_sincostab; kp.g_sincostab_pad = h->d_sincostab_pad; kp.g_crxy = h->d_crxy;
    {
        void* tt = nullptr;
        CK(cudaGetSymbolAddress(&tt, r2p2_d_tanhtab));
        kp.g_tanhtab = reinterpret_cast<const double*>(tt);
    }
    StateArrays& s = h->st;
    kp.x = s.x; kp.y = s.y; kp.theta = s.theta; kp.speed = s.speed; kp.omega = s.omega; kp.last_x = s.last_x; kp.last_y = s.last_y;
    kp.dist = s.dist; kp.odom_x = s.odom_x; kp.odom_y = s.odom_y; kp.origin_x = s.origin_x; kp.origin_y = s.origin_y;
    kp.time = s.time; kp.fitness = s.fitness; kp.flags = s.flags; kp.ncollide = s.ncollide; kp.ticks = s.ticks;
    kp.nsamp_sonar = s.nsamp_sonar; kp.nsamp_coll = s.nsamp_coll;
    kp.ndraws = s.ndraws;
    kp.noise_mode = h->noise_mode; kp.noise_len = h->noise_len; kp.noise_seed = h->noise_seed; kp.noise_stream = h->d_noise;
    kp.robot_index0 = h->robot_index0;
    if (h->noise_mode == 1 && h->noise_P != h->P) return fail(h, R2P2_E_ARG, "noise stream was uploaded for %d robots, population is %d", h->noise_P, h->P);
    kp.work_counter = h->d_work + h->slot;
    kp.genomes = h->d_genomes;
    kp.goal_on = h->goal_on; kp.goal_is_wall = h->goal_is_wall; kp.goal_x = h->goal_x; kp.goal_y = h->goal_y;
    kp.goal_r = h->radius * 1.85;                  // pid_controller.py:213
    kp.goal_tick = s.goal_tick;
    kp.pid_acc_ang = h->pid_acc_ang; kp.pid_acc_dist = h->pid_acc_dist; kp.pid_last_ang = h->pid_last_ang; kp.pid_last_dist = h->pid_last_dist;
    kp.pid_target = h->pid_target; kp.pid_state = h->pid_state; kp.pid_backing = h->pid_backing; kp.pid_ngoals = h->pid_ngoals;
    kp.pid_gx = h->pid_gx; kp.pid_gy = h->pid_gy; kp.pid_max_acc = h->pid_max_acc; kp.pid_accel = h->pid_accel;
    return R2P2_OK;
}

static int run_check(r2p2_handle h, int32_t T) {
    if (!h) return R2P2_E_ARG;
    if (!h->d_wall) return fail(h, R2P2_E_ARG, "r2p2_run: no stage (call r2p2_set_stage)");
    if (!h->have_robot) return fail(h, R2P2_E_ARG, "r2p2_run: no robot (call r2p2_set_robot)");
    if (!h->have_ctrl) return fail(h, R2P2_E_ARG, "r2p2_run: no controller (call r2p2_set_controller)");
    if (!h->have_genomes) return fail(h, R2P2_E_ARG, "r2p2_run: no genomes (call r2p2_upload_genomes)");
    if (!h->have_reset) return fail(h, R2P2_E_ARG, "r2p2_run: population not initialised (call r2p2_reset)");
    if (T < 0) return fail(h, R2P2_E_ARG, "T must be >= 0");
    if (h->kind == R2P2_CTRL_NEURO) {
        if (h->layers[0] != h->R) return fail(h, R2P2_E_ARG, "MLP input width %d != number of sonar rays %d", h->layers[0], h->R);
        int g = 0;
        for (int l = 0; l + 1 < h->n_layers; ++l) g += h->layers[l] * h->layers[l + 1];
        if (g != h->G) return fail(h, R2P2_E_ARG, "genome length %d != MLP weight count %d", h->G, g);
    } else if (h->kind == R2P2_CTRL_LINEAR) {
        if (h->G < 4 || (h->G & 1)) return fail(h, R2P2_E_ARG, "linear chromosome length must be even and >= 4 (got %d)", h->G);
    } else if (h->kind == R2P2_CTRL_PID) {
        if (h->G != 6) return fail(h, R2P2_E_ARG, "PID genome must be the six gains [ap, ai, ad, lp, li, ld] (got %d)", h->G);
        if (h->R < 3) return fail(h, R2P2_E_ARG, "the PID controller reads sonar rays 0, 1, 2, -1 and -2: at least three rays");
        if (!h->have_pid) return fail(h, R2P2_E_ARG, "r2p2_run: no goal list (call r2p2_set_pid, then r2p2_reset)");
        if (h->pid_cap < h->P) return fail(h, R2P2_E_ARG, "r2p2_run: controller state not initialised (call r2p2_reset after r2p2_set_pid)");
    } else if (h->G != 2) {
        return fail(h, R2P2_E_ARG, "constant-command genome must be [speed, angular_velocity]");
    }
    if (h->trace && (T > h->trace_T)) return fail(h, R2P2_E_ARG, "trace enabled for at most %d ticks per run", h->trace_T);
    return R2P2_OK;
}

// One episode-kernel launch over robots [lo, hi) (the whole population for r2p2_run; a chunk of it for r2p2_evaluate,
// whose later genomes are still on their way).  cap_out != NULL: no launch, only the number of robots resident at once.
static int run_range(r2p2_handle h, int32_t T, double delta, double delta_ctrl, int32_t evolve, int lo, int hi, long long* cap_out) {
    KParams kp;
    int rc0 = fill_params(h, kp, T, delta, delta_ctrl, evolve);
    if (rc0) return rc0;
    if (h->lpr_epoch != h->cfg_epoch) { h->lpr_cached = choose_lpr(h, kp); h->lpr_epoch = h->cfg_epoch; }
    const int lpr = h->lpr_cached;
    const bool smem_stage = h->smem_req == 1;   // default off: measured on B200, L1-cached global reads of the planes are as fast as the shared copy
    if (cap_out) {
        int occ = 0;
        rc0 = dispatch_episode(h, kp, lpr, smem_stage, &occ);
        if (rc0) return rc0;
        *cap_out = (long long)std::max(occ, 1) * h->sm_count * (kp.block / lpr);
        return R2P2_OK;
    }
    const bool whole = lo == 0 && hi == h->P;
    if (lpr == 1) {
        if (!whole || !h->genomes_t_valid) {
            dim3 grid((hi - lo + 31) / 32, (h->G + 31) / 32), block(32, 8);
            transpose_genomes_kernel<<<grid, block, 0, h->slot ? h->aux_stream : h->stream>>>(h->d_genomes, h->d_genomes_t, lo, hi, h->G);
            h->launches++;
            CK(cudaGetLastError());
            if (whole) h->genomes_t_valid = true;
        }
        kp.genomes = h->d_genomes_t;
    } else {
        kp.genomes = h->d_genomes;
    }

    if (h->trace) {
        if (h->trace_P != h->P || h->trace_R != h->R) return fail(h, R2P2_E_ARG, "trace buffers were sized for another population; call r2p2_set_trace again");
        if (lo == 0) {
            const size_t rows = (size_t)T * h->P;
            CK(cudaMemsetAsync(h->tr_pose, 0, rows * 5 * sizeof(double), h->stream));
            CK(cudaMemsetAsync(h->tr_dst, 0, rows * h->R * sizeof(double), h->stream));
            CK(cudaMemsetAsync(h->tr_hitk, 0xff, rows * h->R * sizeof(int32_t), h->stream));
            CK(cudaMemsetAsync(h->tr_hitpx, 0xff, rows * h->R * 2 * sizeof(int32_t), h->stream));
            CK(cudaMemsetAsync(h->tr_nsamp, 0, rows * h->R * sizeof(uint32_t), h->stream));
            CK(cudaMemsetAsync(h->tr_ncoll, 0, rows * sizeof(uint32_t), h->stream));
            CK(cudaMemsetAsync(h->tr_cinfo, 0, rows * kCinfoRow * sizeof(double), h->stream));
        }
        kp.tr_pose = h->tr_pose; kp.tr_dst = h->tr_dst; kp.tr_hitk = h->tr_hitk; kp.tr_hitpx = h->tr_hitpx;
        kp.tr_nsamp = h->tr_nsamp; kp.tr_ncoll = h->tr_ncoll; kp.tr_cinfo = h->tr_cinfo;
    }

    if (smem_stage) {
        const size_t with_stage = episode_smem_bytes(128, lpr, true, kp.plane_bytes, kp.has50, kp.maxw, w_in_smem_for(128, lpr, kp.G), kp.G);
        if (with_stage > h->smem_optin) return fail(h, R2P2_E_LIMIT, "stage planes (%zu bytes with tables) exceed shared memory", with_stage);
    }
    kp.rb_begin = (uint32_t)lo; kp.rb_end = (uint32_t)hi;
    return dispatch_episode(h, kp, lpr, smem_stage, nullptr);
}

int r2p2_run(r2p2_handle h, int32_t T, double delta, double delta_ctrl, int32_t evolve) {
    const int rc = run_check(h, T);
    if (rc) return rc;
    ON_DEVICE(h);
    return run_range(h, T, delta, delta_ctrl, evolve, 0, h->P, nullptr);
}

/* evolution.evaluate_ann for a whole population (r2p2/evolution.py:29-52) as one call: upload, reset, run, read back.  When
 * the population runs in three or more waves (a wave = as many robots as are resident at once) the genomes go up in two
 * pieces on a copy stream: the first few waves' worth, then the rest while the episode kernel already runs the first
 * piece.  The second piece gets its own launch on a second stream, so its CTAs move in as the first launch's CTAs retire
 * (one stream would put a full drain between the two). */
int r2p2_evaluate(r2p2_handle h, const double* genomes, int32_t P, int32_t G, const double* x0, const double* y0, const double* theta0,
                  int32_t n, double time_budget, int32_t T, double delta, double delta_ctrl, int32_t evolve, double* fitness) {
    if (!h) return R2P2_E_ARG;
    if (!genomes || P < 1 || G < 1) return fail(h, R2P2_E_ARG, "genomes: need P >= 1, G >= 1");
    if (!fitness) return fail(h, R2P2_E_ARG, "fitness is NULL");
    ON_DEVICE(h);
    int rc = ensure_genomes(h, P, G);
    if (rc) return rc;
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : h->ev_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_idle, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming));
    }
    h->have_genomes = true;
    rc = r2p2_reset(h, x0, y0, theta0, n, time_budget);
    if (rc) return rc;
    rc = run_check(h, T);
    if (rc) return rc;
    long long cap = 0;
    rc = run_range(h, T, delta, delta_ctrl, evolve, 0, P, &cap);
    if (rc) return rc;
    const long long waves = ((long long)P + cap - 1) / cap;
    const bool split = waves >= 3 && (size_t)P * G * sizeof(double) >= ((size_t)8 << 20) && !h->trace;
    const int cut = split ? (int)std::min<long long>(P, std::max<long long>(1, waves / 8) * cap) : P;
    // earlier launches on the compute stream may still read the genome buffer; the reset must be done before either launch
    CK(cudaEventRecord(h->ev_idle, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_idle, 0));
    CK(cudaMemcpyAsync(h->d_genomes, genomes, (size_t)cut * G * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->ev_chunk[0], h->copy_stream));
    if (split) {
        CK(cudaMemcpyAsync(h->d_genomes + (size_t)cut * G, genomes + (size_t)cut * G, (size_t)(P - cut) * G * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->ev_chunk[1], h->copy_stream));
    }
    h->genomes_t_valid = false;
    h->timing_off = true;
    cudaError_t e = cudaStreamWaitEvent(h->stream, h->ev_chunk[0], 0);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev0, h->stream);
    if (e == cudaSuccess) rc = run_range(h, T, delta, delta_ctrl, evolve, 0, cut, nullptr);
    if (e == cudaSuccess && rc == R2P2_OK && split) {
        e = cudaStreamWaitEvent(h->aux_stream, h->ev_idle, 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->aux_stream, h->ev_chunk[1], 0);
        if (e == cudaSuccess) {
            h->slot = 1;
            rc = run_range(h, T, delta, delta_ctrl, evolve, cut, P, nullptr);
            h->slot = 0;
        }
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_aux, h->aux_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_aux, 0);
    }
    h->timing_off = false;
    if (e != cudaSuccess) rc = fail(h, R2P2_E_CUDA, "r2p2_evaluate: %s", cudaGetErrorString(e));
    if (rc) { cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->aux_stream); cudaStreamSynchronize(h->stream); return rc; }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->ev_valid = true;
    h->genomes_t_valid = (h->lpr_cached == 1);
    CK(cudaMemcpyAsync(fitness, h->st.fitness, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

/* ---- path planner's cost grid ---- */
int r2p2_cost_grid(r2p2_handle h, int32_t div_x, int32_t div_y, int32_t* values, int32_t* costs) {
    if (!h) return R2P2_E_ARG;
    if (!h->d_levels) return fail(h, R2P2_E_ARG, "r2p2_cost_grid: no stage (call r2p2_set_stage)");
    if (div_x < 1 || div_y < 1 || div_x > h->W || div_y > h->H) return fail(h, R2P2_E_ARG, "grid divisions must be in 1..W x 1..H");
    if (!values && !costs) return fail(h, R2P2_E_ARG, "values and costs are both NULL");
    ON_DEVICE(h);
    const size_t n = (size_t)div_x * div_y;
    DevBuf<int32_t> d_v, d_c;
    CK(d_v.alloc(n)); CK(d_c.alloc(n));
    cost_grid_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->d_levels, h->W, h->H, div_x, div_y, d_v, d_c);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && values) e = cudaMemcpyAsync(values, d_v, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && costs) e = cudaMemcpyAsync(costs, d_c, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(h, R2P2_E_CUDA, "r2p2_cost_grid: %s", cudaGetErrorString(e));
    return R2P2_OK;
}

/* ---- Sequential_PID_Controller on the device ---- */
int r2p2_set_pid(r2p2_handle h, const double* goals, int32_t n_goals, int32_t per_robot, double max_acceleration, double acceleration) {
    if (!h) return R2P2_E_ARG;
    if (!goals || n_goals < 0) return fail(h, R2P2_E_ARG, "r2p2_set_pid: goals is NULL");
    if (n_goals > R2P2_PID_MAX_GOALS) return fail(h, R2P2_E_LIMIT, "at most %d goals", R2P2_PID_MAX_GOALS);
    if (per_robot && h->P < 1) return fail(h, R2P2_E_ARG, "r2p2_set_pid: per-robot goals need the population size (upload genomes first)");
    ON_DEVICE(h);
    const size_t n = (size_t)(per_robot ? h->P : 1) * (size_t)n_goals * 2;
    CK(cudaStreamSynchronize(h->stream));
    CK(dalloc(&h->d_pid_goals, n));
    if (n) CK(cudaMemcpyAsync(h->d_pid_goals, goals, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->pid_n_goals = n_goals; h->pid_per_robot = per_robot ? 1 : 0; h->pid_goals_P = per_robot ? h->P : 0;
    h->pid_max_acc = max_acceleration; h->pid_accel = acceleration;
    h->have_pid = true;
    h->have_reset = false;                               // the goal lists are loaded by r2p2_reset
    return R2P2_OK;
}

int r2p2_read_pid(r2p2_handle h, int32_t* goals_left, int32_t* state, int32_t* backing, double* target_angle, double* goal_x, double* goal_y) {
    if (!h) return R2P2_E_ARG;
    if (!h->have_pid || h->P < 1 || h->pid_cap < h->P) return fail(h, R2P2_E_ARG, "r2p2_read_pid: no PID controller state (r2p2_set_pid, r2p2_reset)");
    ON_DEVICE(h);
    const size_t P = (size_t)h->P;
    std::vector<int32_t> ng(P);
    CK(cudaMemcpyAsync(ng.data(), h->pid_ngoals, P * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (state) CK(cudaMemcpyAsync(state, h->pid_state, P * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (backing) CK(cudaMemcpyAsync(backing, h->pid_backing, P * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (target_angle) CK(cudaMemcpyAsync(target_angle, h->pid_target, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    std::vector<double> gx, gy;
    if (goal_x || goal_y) {
        gx.resize(P * R2P2_PID_MAX_GOALS); gy.resize(P * R2P2_PID_MAX_GOALS);
        CK(cudaMemcpyAsync(gx.data(), h->pid_gx, gx.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(gy.data(), h->pid_gy, gy.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < P; ++i) {
        if (goals_left) goals_left[i] = ng[i];
        const bool any = ng[i] > 0 && ng[i] <= R2P2_PID_MAX_GOALS;
        if (goal_x) goal_x[i] = any ? gx[(size_t)(ng[i] - 1) * P + i] : std::nan("");
        if (goal_y) goal_y[i] = any ? gy[(size_t)(ng[i] - 1) * P + i] : std::nan("");
    }
    return R2P2_OK;
}

/* ---- goal flagging ---- */
int r2p2_set_goal(r2p2_handle h, int32_t enable, double gx, double gy) {
    if (!h) return R2P2_E_ARG;
    if (!enable) { h->goal_on = 0; return R2P2_OK; }
    if (!h->d_wall) return fail(h, R2P2_E_ARG, "r2p2_set_goal: no stage (call r2p2_set_stage)");
    if (!(std::isfinite(gx) && std::isfinite(gy))) return fail(h, R2P2_E_ARG, "goal must be finite");
    long long ix = (long long)gx, iy = (long long)gy;          // Python int(): truncation
    if (ix < 0) ix += h->W;
    if (iy < 0) iy += h->H;
    if (ix < 0 || ix >= h->W || iy < 0 || iy >= h->H) return fail(h, R2P2_E_ARG, "goal pixel off the map (the reference raises IndexError)");
    ON_DEVICE(h);
    uint32_t word = 0;
    CK(cudaMemcpyAsync(&word, h->d_wall + (size_t)iy * h->wpr + (size_t)(ix >> 5), sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->goal_on = 1; h->goal_x = gx; h->goal_y = gy; h->goal_is_wall = (int)((word >> (ix & 31)) & 1u);
    return R2P2_OK;
}

int r2p2_read_goal(r2p2_handle h, uint32_t* goal_tick) {
    if (!h || !goal_tick) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(goal_tick, h->st.goal_tick, (size_t)h->P * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

/* ---- controllers that run on the host: the tick in two halves --------------------------------- */
static int bridge_ready(r2p2_handle h, const char* who) {
    if (!h) return R2P2_E_ARG;
    if (!h->d_wall) return fail(h, R2P2_E_ARG, "%s: no stage (call r2p2_set_stage)", who);
    if (!h->have_robot) return fail(h, R2P2_E_ARG, "%s: no robot (call r2p2_set_robot)", who);
    if (h->P < 1 || !h->have_reset) return fail(h, R2P2_E_ARG, "%s: population not initialised (upload genomes to size it, then r2p2_reset)", who);
    if (h->P > h->bridge_cap || h->R != h->bridge_R) {
        ON_DEVICE(h);
        CK(cudaStreamSynchronize(h->stream));
        CK(dalloc(&h->d_sense_dst, (size_t)h->P * h->R));
        CK(dalloc(&h->d_sense_px, (size_t)h->P * h->R * 2));
        CK(dalloc(&h->d_cmd, (size_t)h->P * 2));
        h->bridge_cap = h->P; h->bridge_R = h->R;
    }
    return R2P2_OK;
}

int r2p2_sense(r2p2_handle h, double* dst, int32_t* hitpx) {
    int rc = bridge_ready(h, "r2p2_sense"); if (rc) return rc;
    if (!dst) return fail(h, R2P2_E_ARG, "dst is NULL");
    ON_DEVICE(h);
    SenseParams sp;
    rc = fill_params(h, sp.k, 1, 0.0, 0.0, 0);
    if (rc) return rc;
    sp.dst = h->d_sense_dst; sp.hitpx = h->d_sense_px;
    CK(cudaMemsetAsync(h->d_sense_dst, 0, (size_t)h->P * h->R * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->d_sense_px, 0xff, (size_t)h->P * h->R * 2 * sizeof(int32_t), h->stream));
    sense_kernel<<<(h->P + 127) / 128, 128, 0, h->stream>>>(sp);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, h->d_sense_dst, (size_t)h->P * h->R * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (hitpx) CK(cudaMemcpyAsync(hitpx, h->d_sense_px, (size_t)h->P * h->R * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_move(r2p2_handle h, const double* commands, double delta) {
    int rc = bridge_ready(h, "r2p2_move"); if (rc) return rc;
    if (!commands) return fail(h, R2P2_E_ARG, "commands is NULL");
    if (h->trace && h->trace_T < 1) return fail(h, R2P2_E_ARG, "trace enabled for 0 ticks");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(h->d_cmd, commands, (size_t)h->P * 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    KParams kp;
    rc = fill_params(h, kp, 1, delta, 0.0, 0);
    if (rc) return rc;
    kp.ctrl_kind = R2P2_CTRL_CONST; kp.n_layers = 0; kp.G = 2; kp.genomes = h->d_cmd; kp.skip_sense = 1;
    kp.maxw = std::max(1, h->R);
    if (h->trace) {
        if (h->trace_P != h->P || h->trace_R != h->R) return fail(h, R2P2_E_ARG, "trace buffers were sized for another population; call r2p2_set_trace again");
        const size_t rows = (size_t)h->P;
        CK(cudaMemsetAsync(h->tr_pose, 0, rows * 5 * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->tr_ncoll, 0, rows * sizeof(uint32_t), h->stream));
        CK(cudaMemsetAsync(h->tr_cinfo, 0, rows * kCinfoRow * sizeof(double), h->stream));
        kp.tr_pose = h->tr_pose; kp.tr_dst = h->tr_dst; kp.tr_hitk = h->tr_hitk; kp.tr_hitpx = h->tr_hitpx;
        kp.tr_nsamp = h->tr_nsamp; kp.tr_ncoll = h->tr_ncoll; kp.tr_cinfo = h->tr_cinfo;
    }
    // the whole warp / a pair of lanes per robot: the command is read as genome[0], genome[1] of a [P][2] array
    const int lpr = ((long long)h->P * 32 <= (long long)h->sm_count * 512) ? 32 : 2;
    rc = dispatch_episode(h, kp, lpr, false, nullptr);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));     // `commands` may be pageable and reused by the caller
    return R2P2_OK;
}

int r2p2_sync(r2p2_handle h) {
    if (!h) return R2P2_E_ARG;
    ON_DEVICE(h);
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_last_run_ms(r2p2_handle h, float* ms) {
    if (!h || !ms) return R2P2_E_ARG;
    if (!h->ev_valid) return fail(h, R2P2_E_ARG, "no run yet");
    ON_DEVICE(h);
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return R2P2_OK;
}

int r2p2_read_fitness(r2p2_handle h, double* fitness) {
    if (!h || !fitness) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(fitness, h->st.fitness, sizeof(double) * h->P, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_read_state(r2p2_handle h, double* x, double* y, double* theta, double* speed, double* omega, uint32_t* flags,
                    uint32_t* ncollide, uint32_t* ticks) {
    if (!h) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    ON_DEVICE(h);
    const StateArrays& s = h->st;
    const size_t nb = sizeof(double) * h->P, nu = sizeof(uint32_t) * h->P;
    if (x) CK(cudaMemcpyAsync(x, s.x, nb, cudaMemcpyDeviceToHost, h->stream));
    if (y) CK(cudaMemcpyAsync(y, s.y, nb, cudaMemcpyDeviceToHost, h->stream));
    if (theta) CK(cudaMemcpyAsync(theta, s.theta, nb, cudaMemcpyDeviceToHost, h->stream));
    if (speed) CK(cudaMemcpyAsync(speed, s.speed, nb, cudaMemcpyDeviceToHost, h->stream));
    if (omega) CK(cudaMemcpyAsync(omega, s.omega, nb, cudaMemcpyDeviceToHost, h->stream));
    if (flags) CK(cudaMemcpyAsync(flags, s.flags, nu, cudaMemcpyDeviceToHost, h->stream));
    if (ncollide) CK(cudaMemcpyAsync(ncollide, s.ncollide, nu, cudaMemcpyDeviceToHost, h->stream));
    if (ticks) CK(cudaMemcpyAsync(ticks, s.ticks, nu, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

/* ---- checkpoint / resume of the population's episode state ---- */
/* every per-robot array the episode kernel carries between launches, in blob order */
static int copy_state(r2p2_handle h, char* blob, bool to_host) {
    StateArrays& s = h->st;
    const size_t P = (size_t)h->P;
    struct Arr { void* d; size_t n; };
    const Arr arrs[] = {
        {s.x, 8 * P}, {s.y, 8 * P}, {s.theta, 8 * P}, {s.speed, 8 * P}, {s.omega, 8 * P}, {s.last_x, 8 * P}, {s.last_y, 8 * P},
        {s.dist, 8 * P}, {s.odom_x, 8 * P}, {s.odom_y, 8 * P}, {s.origin_x, 8 * P}, {s.origin_y, 8 * P}, {s.time, 8 * P}, {s.fitness, 8 * P},
        {s.nsamp_sonar, 8 * P}, {s.nsamp_coll, 8 * P},
        {s.flags, 4 * P}, {s.ncollide, 4 * P}, {s.ticks, 4 * P}, {s.ndraws, 4 * P}, {s.goal_tick, 4 * P}};
    size_t off = 0;
    for (const Arr& a : arrs) {
        if (to_host) CK(cudaMemcpyAsync(blob + off, a.d, a.n, cudaMemcpyDeviceToHost, h->stream));
        else CK(cudaMemcpyAsync(a.d, blob + off, a.n, cudaMemcpyHostToDevice, h->stream));
        off += a.n;
    }
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

/* blob = StateHeader + the arrays of copy_state */
struct StateHeader {
    uint32_t magic, version;
    int32_t P, R, G, kind;
    uint32_t robot_index0, reserved;
};
static const uint32_t kStateMagic = 0x53503252u;   /* "R2PS" */
static const uint32_t kStateVersion = 2u;

int r2p2_state_bytes(r2p2_handle h, uint64_t* bytes) {
    if (!h || !bytes) return R2P2_E_ARG;
    *bytes = sizeof(StateHeader) + (uint64_t)h->P * (16 * 8 + 5 * 4);
    return R2P2_OK;
}

int r2p2_save_state(r2p2_handle h, void* blob) {
    if (!h || !blob) return R2P2_E_ARG;
    if (h->P < 1 || !h->have_reset) return fail(h, R2P2_E_ARG, "r2p2_save_state: population not initialised");
    ON_DEVICE(h);
    StateHeader hd = {kStateMagic, kStateVersion, h->P, h->R, h->G, h->have_ctrl ? h->kind : -1, h->robot_index0, 0u};
    memcpy(blob, &hd, sizeof hd);
    return copy_state(h, (char*)blob + sizeof hd, true);
}

int r2p2_load_state(r2p2_handle h, const void* blob) {
    if (!h || !blob) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "r2p2_load_state: no population (upload genomes first)");
    StateHeader hd;
    memcpy(&hd, blob, sizeof hd);
    if (hd.magic != kStateMagic) return fail(h, R2P2_E_ARG, "r2p2_load_state: not a state blob (bad magic)");
    if (hd.version != kStateVersion) return fail(h, R2P2_E_ARG, "r2p2_load_state: blob format version %u, this library reads %u", hd.version, kStateVersion);
    if (hd.P != h->P || hd.R != h->R || hd.G != h->G || hd.kind != (h->have_ctrl ? h->kind : -1))
        return fail(h, R2P2_E_ARG, "r2p2_load_state: blob was saved for P=%d R=%d G=%d controller %d; this handle has P=%d R=%d G=%d controller %d",
                    hd.P, hd.R, hd.G, hd.kind, h->P, h->R, h->G, h->have_ctrl ? h->kind : -1);
    ON_DEVICE(h);
    const int rc = copy_state(h, (char*)blob + sizeof hd, false);
    if (rc == R2P2_OK) { h->have_reset = true; h->robot_index0 = hd.robot_index0; }
    return rc;
}

int r2p2_read_counters(r2p2_handle h, uint64_t* sonar_samples, uint64_t* collision_samples, uint64_t* robot_steps) {
    if (!h) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    ON_DEVICE(h);
    CK(cudaMemsetAsync(h->d_cnt3, 0, 3 * sizeof(unsigned long long), h->stream));
    const int blocks = std::min((h->P + 255) / 256, 4 * h->sm_count);
    sum_counters_kernel<<<blocks, 256, 0, h->stream>>>(h->st.nsamp_sonar, h->st.nsamp_coll, h->st.ticks, h->P, h->d_cnt3);
    h->launches++;
    CK(cudaGetLastError());
    unsigned long long c[3];
    CK(cudaMemcpyAsync(c, h->d_cnt3, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (sonar_samples) *sonar_samples = c[0];
    if (collision_samples) *collision_samples = c[1];
    if (robot_steps) *robot_steps = c[2];
    return R2P2_OK;
}

int r2p2_fitness_device(r2p2_handle h, double** d_fitness) {
    if (!h || !d_fitness) return R2P2_E_ARG;
    if (h->P < 1) return fail(h, R2P2_E_ARG, "no population");
    *d_fitness = h->st.fitness;
    return R2P2_OK;
}

int r2p2_launch_count(r2p2_handle h, uint64_t* n) {
    if (!h || !n) return R2P2_E_ARG;
    *n = h->launches;
    return R2P2_OK;
}

int r2p2_set_trace(r2p2_handle h, int32_t enable, int32_t T_max) {
    if (!h) return R2P2_E_ARG;
    ON_DEVICE(h);
    CK(cudaStreamSynchronize(h->stream));
    free_trace(h);
    h->trace = false;
    if (!enable) return R2P2_OK;
    if (h->P < 1 || !h->have_robot) return fail(h, R2P2_E_ARG, "set the robot and the population before enabling the trace");
    if (T_max < 1) return fail(h, R2P2_E_ARG, "T_max must be >= 1");
    const size_t rows = (size_t)T_max * h->P;
    CK(dalloc(&h->tr_pose, rows * 5)); CK(dalloc(&h->tr_dst, rows * h->R));
    CK(dalloc(&h->tr_hitk, rows * h->R)); CK(dalloc(&h->tr_hitpx, rows * h->R * 2));
    CK(dalloc(&h->tr_nsamp, rows * h->R)); CK(dalloc(&h->tr_ncoll, rows)); CK(dalloc(&h->tr_cinfo, rows * kCinfoRow));
    h->trace = true; h->trace_T = T_max; h->trace_P = h->P; h->trace_R = h->R;
    return R2P2_OK;
}

int r2p2_read_trace(r2p2_handle h, int32_t T, double* pose, double* dst, int32_t* hitk, int32_t* hitpx, uint32_t* nsamp, uint32_t* ncoll) {
    if (!h) return R2P2_E_ARG;
    if (!h->trace) return fail(h, R2P2_E_ARG, "trace not enabled");
    if (T < 0 || T > h->trace_T) return fail(h, R2P2_E_ARG, "T out of range");
    ON_DEVICE(h);
    const size_t rows = (size_t)T * h->trace_P, R = (size_t)h->trace_R;
    if (pose) CK(cudaMemcpyAsync(pose, h->tr_pose, rows * 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (dst) CK(cudaMemcpyAsync(dst, h->tr_dst, rows * R * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (hitk) CK(cudaMemcpyAsync(hitk, h->tr_hitk, rows * R * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (hitpx) CK(cudaMemcpyAsync(hitpx, h->tr_hitpx, rows * R * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (nsamp) CK(cudaMemcpyAsync(nsamp, h->tr_nsamp, rows * R * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (ncoll) CK(cudaMemcpyAsync(ncoll, h->tr_ncoll, rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_read_trace_collisions(r2p2_handle h, int32_t T, double* cinfo) {
    if (!h) return R2P2_E_ARG;
    if (!h->trace) return fail(h, R2P2_E_ARG, "trace not enabled");
    if (T < 0 || T > h->trace_T) return fail(h, R2P2_E_ARG, "T out of range");
    if (!cinfo) return fail(h, R2P2_E_ARG, "cinfo is NULL");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(cinfo, h->tr_cinfo, (size_t)T * h->trace_P * kCinfoRow * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

/* ---- on-device GA (r2p2_ga.cuh) --------------------------------------------------------------- */
static int ga_ready(r2p2_handle h) {
    if (!h) return R2P2_E_ARG;
    if (!h->d_ga_pop[0]) return fail(h, R2P2_E_ARG, "no GA population (call r2p2_ga_create)");
    return R2P2_OK;
}

int r2p2_ga_create(r2p2_handle h, int32_t P, int32_t G, uint64_t seed, double min_gen, double max_gen) {
    if (!h) return R2P2_E_ARG;
    if (P < 2 || G < 1) return fail(h, R2P2_E_ARG, "GA: need P >= 2, G >= 1");
    if (!(min_gen < max_gen)) return fail(h, R2P2_E_ARG, "GA: need min_gen < max_gen");
    ON_DEVICE(h);
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 2; ++i) CK(dalloc(&h->d_ga_pop[i], (size_t)P * G));
    CK(dalloc(&h->d_ga_fit, (size_t)P));
    CK(dalloc(&h->d_ga_elite, (size_t)kGaMaxElites));
    CK(cudaMemsetAsync(h->d_ga_fit, 0, (size_t)P * sizeof(double), h->stream));
    h->ga_P = P; h->ga_G = G; h->ga_seed = seed; h->ga_min = min_gen; h->ga_max = max_gen; h->ga_cur = 0; h->ga_generation = 0;
    const size_t n = (size_t)P * G;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8);
    ga_init_kernel<<<blocks, 256, 0, h->stream>>>(h->d_ga_pop[0], P, G, min_gen, max_gen, seed);
    h->launches++;
    CK(cudaGetLastError());
    return R2P2_OK;
}

int r2p2_ga_set_operators(r2p2_handle h, int32_t tournament, double crossover_rate, double mutation_rate, double mutation_sigma, int32_t elites) {
    if (!h) return R2P2_E_ARG;
    if (tournament < 1 || tournament > 4) return fail(h, R2P2_E_ARG, "GA: tournament size must be 1..4");
    if (elites < 0 || elites > kGaMaxElites) return fail(h, R2P2_E_ARG, "GA: elites must be 0..%d", kGaMaxElites);
    if (!(crossover_rate >= 0.0 && crossover_rate <= 1.0 && mutation_rate >= 0.0 && mutation_rate <= 1.0 && mutation_sigma >= 0.0))
        return fail(h, R2P2_E_ARG, "GA: rates must be in [0, 1], sigma >= 0");
    h->ga_tournament = tournament; h->ga_cx = crossover_rate; h->ga_mut = mutation_rate; h->ga_sigma = mutation_sigma; h->ga_elites = elites;
    return R2P2_OK;
}

int r2p2_ga_population_device(r2p2_handle h, double** d_pop) {
    int rc = ga_ready(h); if (rc) return rc;
    if (!d_pop) return fail(h, R2P2_E_ARG, "d_pop is NULL");
    *d_pop = h->d_ga_pop[h->ga_cur];
    return R2P2_OK;
}

int r2p2_ga_fitness_device(r2p2_handle h, double** d_fit) {
    int rc = ga_ready(h); if (rc) return rc;
    if (!d_fit) return fail(h, R2P2_E_ARG, "d_fit is NULL");
    *d_fit = h->d_ga_fit;
    return R2P2_OK;
}

int r2p2_ga_write_fitness(r2p2_handle h, const double* fitness, int32_t first, int32_t count) {
    int rc = ga_ready(h); if (rc) return rc;
    if (!fitness || first < 0 || count < 0 || first + count > h->ga_P) return fail(h, R2P2_E_ARG, "GA: fitness range out of bounds");
    ON_DEVICE(h);
    CK(cudaMemcpyAsync(h->d_ga_fit + first, fitness, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_ga_read(r2p2_handle h, double* population, double* fitness) {
    int rc = ga_ready(h); if (rc) return rc;
    ON_DEVICE(h);
    if (population) CK(cudaMemcpyAsync(population, h->d_ga_pop[h->ga_cur], (size_t)h->ga_P * h->ga_G * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (fitness) CK(cudaMemcpyAsync(fitness, h->d_ga_fit, (size_t)h->ga_P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_ga_evaluate(r2p2_handle h, int32_t first, int32_t count, int32_t T, double delta, double delta_ctrl, int32_t evolve) {
    int rc = ga_ready(h); if (rc) return rc;
    if (first < 0 || count < 1 || first + count > h->ga_P) return fail(h, R2P2_E_ARG, "GA: individual range out of bounds");
    if (!h->have_pose1) return fail(h, R2P2_E_ARG, "GA: no start pose (call r2p2_reset with n == 1 once)");
    rc = r2p2_set_genomes_device(h, h->d_ga_pop[h->ga_cur] + (size_t)first * h->ga_G, count, h->ga_G);
    if (rc) return rc;
    h->robot_index0 = (uint32_t)first;                     // the device noise generator is keyed by the global individual index
    rc = r2p2_reset(h, &h->rx0, &h->ry0, &h->rt0, 1, h->rbudget);
    if (rc) return rc;
    rc = r2p2_run(h, T, delta, delta_ctrl, evolve);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_ga_fit + first, h->st.fitness, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return R2P2_OK;
}

int r2p2_ga_breed(r2p2_handle h) {
    int rc = ga_ready(h); if (rc) return rc;
    ON_DEVICE(h);
    GaParams g;
    g.P = h->ga_P; g.G = h->ga_G; g.tournament = h->ga_tournament; g.elites = std::min(h->ga_elites, h->ga_P);
    g.crossover_rate = h->ga_cx; g.mutation_rate = h->ga_mut; g.mutation_sigma = h->ga_sigma; g.min_gen = h->ga_min; g.max_gen = h->ga_max;
    g.seed = h->ga_seed; g.generation = h->ga_generation;
    g.pop = h->d_ga_pop[h->ga_cur]; g.fit = h->d_ga_fit; g.next = h->d_ga_pop[h->ga_cur ^ 1]; g.elite_idx = h->d_ga_elite;
    ga_elites_kernel<<<1, 1024, 0, h->stream>>>(h->d_ga_fit, h->ga_P, std::max(g.elites, 1), h->d_ga_elite);
    const long long threads = (long long)h->ga_P * 32;
    ga_breed_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, h->stream>>>(g);
    h->launches += 2;
    CK(cudaGetLastError());
    h->ga_cur ^= 1;
    h->ga_generation++;
    return R2P2_OK;
}

/* Best individual of the generation that was bred from last (index, fitness, genome [G]); any pointer may be NULL. */
int r2p2_ga_best(r2p2_handle h, int32_t* index, double* fitness, double* genome) {
    int rc = ga_ready(h); if (rc) return rc;
    if (h->ga_generation == 0) return fail(h, R2P2_E_ARG, "GA: no generation bred yet");
    ON_DEVICE(h);
    int idx = -1;
    CK(cudaMemcpyAsync(&idx, h->d_ga_elite, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (idx < 0 || idx >= h->ga_P) return fail(h, R2P2_E_CUDA, "GA: elite index out of range");
    if (index) *index = idx;
    if (fitness) CK(cudaMemcpyAsync(fitness, h->d_ga_fit + idx, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (genome) CK(cudaMemcpyAsync(genome, h->d_ga_pop[h->ga_cur ^ 1] + (size_t)idx * h->ga_G, (size_t)h->ga_G * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

struct GaHeader {
    uint32_t magic, version;
    int32_t P, G, tournament, elites, cur_unused, reserved;
    uint64_t seed;
    uint32_t generation, pad;
    double min_gen, max_gen, cx, mut, sigma;
};
static const uint32_t kGaMagic = 0x47503252u;      /* "R2PG" */

int r2p2_ga_state_bytes(r2p2_handle h, uint64_t* bytes) {
    int rc = ga_ready(h); if (rc) return rc;
    if (!bytes) return fail(h, R2P2_E_ARG, "bytes is NULL");
    *bytes = sizeof(GaHeader) + ((uint64_t)h->ga_P * h->ga_G + (uint64_t)h->ga_P) * sizeof(double);
    return R2P2_OK;
}

int r2p2_ga_save(r2p2_handle h, void* blob) {
    int rc = ga_ready(h); if (rc) return rc;
    if (!blob) return fail(h, R2P2_E_ARG, "blob is NULL");
    ON_DEVICE(h);
    GaHeader hd;
    memset(&hd, 0, sizeof hd);
    hd.magic = kGaMagic; hd.version = 1u; hd.P = h->ga_P; hd.G = h->ga_G; hd.tournament = h->ga_tournament; hd.elites = h->ga_elites;
    hd.seed = h->ga_seed; hd.generation = h->ga_generation;
    hd.min_gen = h->ga_min; hd.max_gen = h->ga_max; hd.cx = h->ga_cx; hd.mut = h->ga_mut; hd.sigma = h->ga_sigma;
    memcpy(blob, &hd, sizeof hd);
    char* q = (char*)blob + sizeof hd;
    const size_t npop = (size_t)h->ga_P * h->ga_G * sizeof(double);
    CK(cudaMemcpyAsync(q, h->d_ga_pop[h->ga_cur], npop, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(q + npop, h->d_ga_fit, (size_t)h->ga_P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

int r2p2_ga_load(r2p2_handle h, const void* blob, uint64_t bytes) {
    if (!h) return R2P2_E_ARG;
    if (!blob || bytes < sizeof(GaHeader)) return fail(h, R2P2_E_ARG, "r2p2_ga_load: blob too short");
    GaHeader hd;
    memcpy(&hd, blob, sizeof hd);
    if (hd.magic != kGaMagic || hd.version != 1u) return fail(h, R2P2_E_ARG, "r2p2_ga_load: not a GA blob of format version 1");
    if (hd.P < 2 || hd.G < 1 || bytes != sizeof(GaHeader) + ((uint64_t)hd.P * hd.G + (uint64_t)hd.P) * sizeof(double))
        return fail(h, R2P2_E_ARG, "r2p2_ga_load: blob size does not match its header (P=%d, G=%d)", hd.P, hd.G);
    int rc = r2p2_ga_create(h, hd.P, hd.G, hd.seed, hd.min_gen, hd.max_gen);
    if (rc) return rc;
    rc = r2p2_ga_set_operators(h, hd.tournament, hd.cx, hd.mut, hd.sigma, hd.elites);
    if (rc) return rc;
    ON_DEVICE(h);
    const char* q = (const char*)blob + sizeof hd;
    const size_t npop = (size_t)hd.P * hd.G * sizeof(double);
    CK(cudaMemcpyAsync(h->d_ga_pop[0], q, npop, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_ga_fit, q + npop, (size_t)hd.P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->ga_cur = 0; h->ga_generation = hd.generation;
    return R2P2_OK;
}

int r2p2_ga_generation(r2p2_handle h, uint32_t* generation) {
    if (!h || !generation) return R2P2_E_ARG;
    *generation = h->ga_generation;
    return R2P2_OK;
}

static int cast_rays_impl(r2p2_handle h, int plain, int32_t n, const double* ox, const double* oy, const double* angle_rad,
                          const double* limit, int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp);

int r2p2_cast_rays(r2p2_handle h, int32_t n, const double* ox, const double* oy, const double* angle_rad, const double* limit,
                   int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp) {
    return cast_rays_impl(h, 0, n, ox, oy, angle_rad, limit, status, hx, hy, k, nsamp);
}

int r2p2_cast_rays_plain(r2p2_handle h, int32_t n, const double* ox, const double* oy, const double* angle_rad, const double* limit,
                         int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp) {
    return cast_rays_impl(h, 1, n, ox, oy, angle_rad, limit, status, hx, hy, k, nsamp);
}

static int cast_rays_impl(r2p2_handle h, int plain, int32_t n, const double* ox, const double* oy, const double* angle_rad,
                          const double* limit, int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp) {
    if (!h) return R2P2_E_ARG;
    if (!h->d_wall) return fail(h, R2P2_E_ARG, "no stage");
    if (n < 1 || !ox || !oy || !angle_rad || !limit || !status || !hx || !hy || !k || !nsamp) return fail(h, R2P2_E_ARG, "cast_rays: bad arguments");
    ON_DEVICE(h);
    DevBuf<double> d_in, d_out; DevBuf<int32_t> d_i;
    CK(d_in.alloc((size_t)4 * n)); CK(d_out.alloc((size_t)2 * n)); CK(d_i.alloc((size_t)3 * n));
    const size_t nb = sizeof(double) * n;
    CK(cudaMemcpyAsync(d_in, ox, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in.p + n, oy, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in.p + 2 * n, angle_rad, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_in.p + 3 * n, limit, nb, cudaMemcpyHostToDevice, h->stream));
    StageView sv; sv.wall = h->d_wall; sv.lvl50 = h->d_l50; sv.dist = h->d_dist; sv.W = h->W; sv.H = h->H; sv.wpr = h->wpr;
    sv.hot = nullptr; sv.hwpr = 0; sv.hx0 = sv.hy0 = 1; sv.hx1 = sv.hy1 = 0; sv.win_sb = 0u; sv.win_ww = 0;       // r2p2_cast_rays: the distance-field march and the plain loop
    cast_rays_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(sv, plain, h->d_sincostab, n, d_in.p, d_in.p + n, d_in.p + 2 * n, d_in.p + 3 * n,
                                                             d_i.p, d_out.p, d_out.p + n, d_i.p + n, reinterpret_cast<uint32_t*>(d_i.p + 2 * n));
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(status, d_i, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(k, d_i.p + n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(nsamp, d_i.p + 2 * n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hx, d_out, nb, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hy, d_out.p + n, nb, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return R2P2_OK;
}

}  // extern "C"

#ifdef R2P2_PHASE_TIMING
// profiling build only (tools/phase_timing.py); not part of the ABI
extern "C" int r2p2_debug_phase_cycles(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16 && cudaMemcpyFromSymbol(out16, r2p2::g_phase_cycles, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(r2p2::g_phase_cycles, z, sizeof z) != cudaSuccess) return -1;
    }
    return 0;
}
extern "C" int r2p2_debug_robot_cycles(unsigned long long* out, int n_robots) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, r2p2::g_robot_cycles, (size_t)n_robots * 12 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
#endif
