// r2p2_kernel_ws.cuh -- warp-synchronous episode kernel for 2..16 lanes per robot.
//
// Same tick, same arithmetic and same results as episode_kernel (r2p2_kernels.cuh); what changes is the control
// structure.  With LPR < 32 a warp carries 32/LPR robots.  In episode_kernel each group runs its own control flow and
// synchronises with its own member mask; a warp-level instruction whose lanes carry different member masks is executed
// once per mask, and groups that are at different places in the code serialise, so a warp with four robots issues
// close to four robots' worth of instructions.  Here the WARP is the unit: it fetches 32/LPR robots at once and walks
// all of them through every tick in lock step.  Every collective uses the full-warp mask (segmented with
// width = LPR / bit slices of the ballot), a robot that has stopped (fault, episode over) is carried along with
// zero-length ray plans and all of its state updates predicated off, and the only divergence left is inside the math
// routines (different branches of glibc's sin/cos for different headings).  One instruction stream now serves all the
// robots of the warp.
#pragma once

#include "r2p2_kernels.cuh"

namespace r2p2 {

#ifdef R2P2_EXP_FETCH_ONLY      // experiment: windows are fetched but the marches read the global plane
constexpr bool kWsWin = false;
#else
constexpr bool kWsWin = true;
#endif

template <int LPR>
struct WGroup {
    int sub;
    unsigned gbase, gbits;
    __device__ __forceinline__ WGroup() {
        const unsigned lane = threadIdx.x & 31u;
        sub = (int)(lane & (unsigned)(LPR - 1));
        gbase = lane & ~(unsigned)(LPR - 1);
        gbits = (LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u);
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(0xffffffffu, p) >> gbase) & gbits; }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
    __device__ __forceinline__ int min_i(int v) const {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o, 32));
        return v;
    }
    __device__ __forceinline__ unsigned sum_u(unsigned v) const {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
        return v;
    }
    __device__ __forceinline__ unsigned min_u(unsigned v) const {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o, 32));
        return v;
    }
};

// own samples per pass of the sonar teams (below 8 lanes per ray) and of the collision team (the whole group, 16 samples
// per pass: collision rays are speed * delta + radius long); the host sizes the windows with the same numbers
#ifdef R2P2_EXP_SONAR_OWN8
template <int LPR> constexpr int kWsSonarOwn = 8;
#else
template <int LPR> constexpr int kWsSonarOwn = (LPR <= 2 ? 8 : 4);
#endif
template <int LPR> constexpr int kWsTeamOwn = kWsSonarOwn<LPR>;
template <int LPR> constexpr int kWsCollOwn = (LPR >= 16 ? 1 : LPR >= 8 ? 2 : 4);

template <int LPR, class NET, bool EXTRA = false>
__global__ void __launch_bounds__((NET::kLayers > 0 && !NET::kSmemWeights) ? kRegBlock : kMaxBlock, 1) episode_kernel_ws(const __grid_constant__ KParams p) {
    static_assert(LPR >= 2 && LPR <= 16, "warp-synchronous kernel: 2..16 lanes per robot");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NRC = (int)blockDim.x / LPR;                // robots in flight per CTA
    const int NR = buf_stride(LPR, NRC);                  // stride of the activation buffers
    constexpr int RPW = 32 / LPR;                         // robots per warp
    constexpr unsigned FULL = 0xffffffffu;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_tab = reinterpret_cast<double*>(smem_raw + 16);
    double* s_crx = s_tab + kSinCosPadLen;
    double* s_cry = s_crx + 360;
    const double* s_tanh = p.g_tanhtab;
    double* s_after = s_cry + 360;
    if (p.tanh_in_smem) { s_tanh = s_after; s_after += R2P2_TANH_NINT * R2P2_TANH_ROW; }
    double* s_bufA = s_after;
    double* s_bufB = s_bufA + (size_t)p.maxw * NR;
    double* s_gen = s_bufB + (size_t)p.maxw * NR;         // [NR][G], only when p.w_in_smem
    // per-robot windows of the hot plane (see below)
    const int win_words = p.win_ww * p.win_wb;
    uint32_t* s_win = reinterpret_cast<uint32_t*>(s_gen + (p.w_in_smem ? (size_t)NRC * smem_gstride(LPR, p.G) : 0));

    const uint32_t mb = smem_u32(mbar);
    if (threadIdx.x == 0) {
        mbar_init(mb, 1);
        mbar_expect_tx(mb, kSinCosPadLen * 8 + 720 * 8 + (p.tanh_in_smem ? R2P2_TANH_NINT * R2P2_TANH_ROW * 8 : 0));
        tma_load_1d(s_tab, p.g_sincostab_pad, kSinCosPadLen * 8, mb);
        tma_load_1d(s_crx, p.g_crxy, 720 * 8, mb);
        if (p.tanh_in_smem) tma_load_1d(const_cast<double*>(s_tanh), p.g_tanhtab, R2P2_TANH_NINT * R2P2_TANH_ROW * 8, mb);
    }
    __syncthreads();
    while (!mbar_try_wait(mb, 0)) { }

    StageView sv;
    sv.wall = p.g_wall; sv.lvl50 = p.g_lvl50; sv.dist = p.g_dist;
    sv.W = p.W; sv.H = p.H; sv.wpr = p.wpr;
    view_global_hot(sv, p.g_hot, p.hwpr, p.hpad);

    const WGroup<LPR> g;
    const unsigned lane = threadIdx.x & 31u;
    const int slot = threadIdx.x / LPR;
    double* bufA = s_bufA + slot;
    double* bufB = s_bufB + slot;
    const int R = p.R;
    const double delta = p.delta;
    const bool tracing = p.tr_pose != nullptr;
    int tl = 1;                                           // lanes per sonar ray: the largest power of two with R teams in the group
    while (tl * 2 * R <= LPR && tl < kMaxSonarTeam) tl *= 2;
    const int rays_per_round = LPR / tl;
    const int team_ray = g.sub / tl, team_j = g.sub % tl;
    Team sonar_team;
    sonar_team.tl = tl; sonar_team.j = team_j; sonar_team.sync = FULL;
    sonar_team.mask = (((1u << tl) - 1u) << (lane - (unsigned)team_j));
    const MarchPlan sonar_plan = make_plan(p.L, tl);
    Team coll_team;
    coll_team.tl = LPR; coll_team.j = g.sub; coll_team.mask = g.gbits << g.gbase; coll_team.sync = FULL;
    // The robot's window: win_ww x win_wb words of the hot plane around its sonar origin, copied once per tick by the
    // lanes of its group in 16-byte asynchronous copies (LDGSTS, L2 -> shared memory, not through L1) -- every sample of
    // the sonar fan, the collision ray and the crash ring of the tick is then a shared-memory read with a fixed latency,
    // whatever L1 holds.  (Measured: one cp.async.bulk per 64-byte band row is 2x slower than no window at all -- the TMA
    // unit's per-copy cost dominates at this size.)
    const uint32_t win_saddr = smem_u32(s_win + (size_t)slot * win_words);
    const int win_cpb = p.win_ww >> 2;                    // 16-byte chunks per band

    const unsigned n_batches = (p.rb_end - p.rb_begin + (unsigned)RPW - 1u) / (unsigned)RPW;
    const unsigned n_slices = p.n_slices > 1 ? (unsigned)p.n_slices : 1u;
    for (;;) {
        // ---- the warp fetches its next work item: 32/LPR robots, all T ticks or one slice of them -----------
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_batches * n_slices) break;
        const unsigned slice = item / n_batches, batch = item - slice * n_batches;
        const unsigned base = batch * (unsigned)RPW + p.rb_begin;
        if (slice > 0u) {
            // Items are handed out in order, so the previous slice of this batch was taken earlier by a warp that is
            // resident and running (an item only ever waits for a smaller one): the wait ends.  A bound on the polls turns
            // a broken invariant into a launch error instead of a hang.
            if (lane == 0) {
                unsigned polls = 0u;
                while (ld_acquire_u32(p.slice_done + batch) < slice) {
                    __nanosleep(256);
                    if (++polls > (1u << 26)) __trap();
                }
            }
            __syncwarp();
        }
        const unsigned rb_raw = base + lane / LPR;
        const bool valid = rb_raw < p.rb_end;
        const unsigned rb = valid ? rb_raw : p.rb_end - 1u;              // lanes without a robot read a valid one, commit nothing

        // (L2 loads: with time slicing the previous slice of this robot may have been written by another SM, and L1 is not
        // coherent between SMs)
        double x = __ldcg(p.x + rb), y = __ldcg(p.y + rb), theta = __ldcg(p.theta + rb), speed = __ldcg(p.speed + rb), omega = __ldcg(p.omega + rb);
        double last_x = __ldcg(p.last_x + rb), last_y = __ldcg(p.last_y + rb);
        double dist = __ldcg(p.dist + rb), odom_x = __ldcg(p.odom_x + rb), odom_y = __ldcg(p.odom_y + rb);
        const double origin_x = __ldcg(p.origin_x + rb), origin_y = __ldcg(p.origin_y + rb);
        double tmr = __ldcg(p.time + rb);
        double fitness = __ldcg(p.fitness + rb);
        unsigned flags = __ldcg(p.flags + rb), ncollide = __ldcg(p.ncollide + rb), ticks = __ldcg(p.ticks + rb);
        const unsigned ncollide_in = ncollide;
        unsigned ndraws = p.noise_mode ? __ldcg(p.ndraws + rb) : 0u;
        unsigned my_ns_sonar = 0u, ns_coll = 0u;
        const double* gen = p.genomes + (size_t)rb * (size_t)p.G;
        FixedMlp<LPR, NET> mlp;
        if constexpr (NET::kLayers > 0 && !NET::kSmemWeights) {
            mlp.load(gen, g.sub);
        } else if (p.w_in_smem) {
            double* dst_g = s_gen + (size_t)slot * smem_gstride(LPR, p.G);
            for (int i = g.sub; i < p.G; i += LPR) dst_g[i] = __ldg(gen + i);
            __syncwarp();
            gen = dst_g;
        }

        unsigned goal_tick = (EXTRA && p.goal_on) ? __ldcg(p.goal_tick + rb) : 0xffffffffu;
        const bool pid = EXTRA && p.ctrl_kind == R2P2_CTRL_PID;      // Sequential_PID_Controller on the device
        PidState ps;
        ps.acc_ang = ps.acc_dist = ps.last_ang = ps.last_dist = ps.target = 0.0; ps.state = ps.backing = ps.ngoals = 0;
        if (pid) pid_load(p, rb, ps);
        bool alive = valid && !(flags & (R2P2_F_FAULT | R2P2_F_DONE));
        StageView svw = sv;                               // the robot's view: its window instead of the global plane
        svw.win_ww = p.win_ww;
        svw.hx0 = 1; svw.hy0 = 1; svw.hx1 = 0; svw.hy1 = 0;   // (nothing fetched yet)
        int tdone = 0;                                    // ticks completed in this launch
        unsigned ncollide0 = ncollide;
        const int t_begin = (n_slices > 1u) ? (int)slice * p.slice_T : 0;
        const int t_end = (n_slices > 1u) ? min(p.T, t_begin + p.slice_T) : p.T;
        for (int t = t_begin; t < t_end; ++t) {
            if (!__any_sync(FULL, alive)) break;
            bool live = alive;
            if (live) ncollide0 = ncollide;
            const size_t trow = (size_t)t * (size_t)p.P + rb;
            unsigned ccode = 0u;                          // collide() calls of this tick, for the trace
            double ccx = 0.0, ccy = 0.0, cbv = 0.0;

            // ---- sense: Robot.check_sensors (robot.py:287-318) ----------------------------------------
            bool fault = live && !(isfinite(x) && isfinite(y));           // round(nan) / round(inf) raise
            const bool sensing = live && !fault;
            const double ox = rint(x), oy = rint(y);
            MarchPlan plan = sonar_plan;                                  // a stopped robot marches zero-length rays
            if (!sensing) { plan.kA = 0; plan.kB = 0; plan.chunk = 0; }
            {
                const bool near_map = sensing & (ox > -32.0) & (ox < (double)(p.W + 32)) & (oy > -32.0) & (oy < (double)(p.H + 32));   // the window stays inside the plane's margin
                const int iox = near_map ? __double2int_rn(ox) : 0, ioy = near_map ? __double2int_rn(oy) : 0;
                // the window fetched on an earlier tick still covers this tick's box: nothing to do (robots move a few pixels per tick)
                const bool covered = (iox - p.win_half >= svw.hx0) & (iox + p.win_half <= svw.hx1) & (ioy - p.win_half >= svw.hy0) & (ioy + p.win_half <= svw.hy1);
                if (sensing && !near_map) { svw.hx0 = 1; svw.hy0 = 1; svw.hx1 = 0; svw.hy1 = 0; }
                __syncwarp();                                             // last tick's reads of the window are done
#ifdef R2P2_EXP_STALE
                if (near_map && t == 0) {
#else
                if (near_map && !covered) {
#endif
                    const int wx0 = ((iox - p.win_half - p.win_slack) >> 5) << 5, wy0 = ((ioy - p.win_half - p.win_slack) >> 2) << 2;
                    const uint32_t* src = p.g_hot + hot_word(wx0, wy0, p.hwpr);
                    int b = 0, q = g.sub;
                    while (q >= win_cpb) { q -= win_cpb; ++b; }
                    while (b < p.win_wb) {
                        cp_async16(win_saddr + (uint32_t)(b * p.win_ww + q * 4) * 4u, src + (size_t)b * p.hwpr + q * 4);
                        q += LPR;
                        while (q >= win_cpb) { q -= win_cpb; ++b; }
                    }
                    svw.hx0 = wx0; svw.hy0 = wy0; svw.hx1 = wx0 + p.win_ww * 8 - 1; svw.hy1 = wy0 + p.win_wb * 4 - 1;
                    svw.win_sb = win_saddr - 4u * (unsigned)hot_word(wx0, wy0, p.win_ww);
                }
                asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
                __syncwarp();                                             // the group's copies are visible to all its lanes
            }
            for (int base_r = 0; base_r < R; base_r += rays_per_round) {
                const int i = base_r + team_ray;
                const bool have_ray = sensing & (team_ray < rays_per_round) & (i < R);
                const double ang = have_ray ? R2P2_ADD(p.ray_deg[i], theta) : 0.0;
                const SinCos sc = sincos_call(R2P2_MUL(ang, kRad), s_tab);
                if (!sc.ok && have_ray) { flags |= R2P2_F_TRIGRANGE; fault = true; }
                const Hit h = team_march<(LPR < kMaxSonarTeam ? LPR : kMaxSonarTeam), kWsSonarOwn<LPR>, kWsWin, kWsTeamOwn<LPR>>(kWsWin ? svw : sv, sonar_team, plan, ox, oy, sc.c, sc.s, p.L);
                if (h.fault && have_ray) fault = true;
                const bool leader = have_ray && (team_j == 0);
                double d = 0.0;
                bool unclamped = false;
                if (leader) {
                    my_ns_sonar += h.nsamp;
                    d = (h.k >= 0) ? r2p2_norm2(R2P2_SUB(h.x, ox), R2P2_SUB(h.y, oy)) : r2p2_from_bits(0x7ff0000000000000ULL);
                    if ((d < p.db_hi && d > p.db_lo) || d > p.L) d = p.L;
                    else unclamped = true;
                }
                if (p.noise_mode) {                       // robot.py:308-310: one draw per unclamped ray, ray order = lane order
                    const unsigned nb = g.ballot(unclamped);
                    const unsigned below = nb & ((1u << g.sub) - 1u);
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
            if (g.any((flags & R2P2_F_TRIGRANGE) != 0u) && live) flags |= R2P2_F_TRIGRANGE;
            if (g.any(fault)) { if (live) flags |= R2P2_F_FAULT; live = false; }
            __syncwarp();

            // ---- control: Neuro_controller.control / Inspyred_controller.control -----------------------
            if (live) {
                dist = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(x, odom_x), R2P2_SUB(y, odom_y)));
                odom_x = x; odom_y = y;
                if (p.ctrl_kind != R2P2_CTRL_CONST && !pid) {
                    if (p.evolve) tmr = R2P2_SUB(tmr, p.delta_ctrl);
                    if (tmr <= 0.0 && p.evolve) {         // episode over: fitness is frozen here
                        fitness = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(odom_x, origin_x), R2P2_SUB(odom_y, origin_y)));
                        flags |= R2P2_F_DONE;
                        speed = 0.0; omega = 0.0;
                        live = false;
                    }
                }
            }
            double v, w;
#ifdef R2P2_EXP_OLDMLP
            if constexpr (false) {
#else
            if constexpr (NET::kLayers > 0 && NET::kSmemWeights) {
#endif
                for (int i = g.sub; i < NET::sz(0); i += LPR) bufA[i * NR] = R2P2_DIV(p.vr1, bufA[i * NR]);   // 25/dst (inf when dst == 0)
                __syncwarp();
                group_mlp<LPR, NET>(gen, bufA, bufB, NR, g.sub, p.activation, s_tanh, v, w);
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
                __syncwarp();
            } else if constexpr (NET::kLayers > 0) {
                double h[FixedMlp<LPR, NET>::kSlots];
#pragma unroll
                for (int s = 0; s < FixedMlp<LPR, NET>::kSlots; ++s) {
                    const int i = g.sub + s * LPR;
                    h[s] = (i < NET::sz(0)) ? R2P2_DIV(p.vr1, bufA[i * NR]) : 0.0;
                }
                if constexpr (NET::kSmemWeights) mlp.forward_smem(gen, FULL, g.sub, h, p.activation, s_tanh, v, w);
                else mlp.forward(FULL, g.sub, h, p.activation, s_tanh, v, w);
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
                __syncwarp();
            } else if (p.ctrl_kind == R2P2_CTRL_NEURO) {
                for (int i = g.sub; i < R; i += LPR) bufA[i * NR] = R2P2_DIV(p.vr1, bufA[i * NR]);   // 25/dst (inf when dst == 0)
                __syncwarp();
                double* hin = bufA;
                double* hout = bufB;
                int woff = 0;
                for (int l = 0; l + 1 < p.n_layers; ++l) {
                    const int nin = p.layers[l], nout = p.layers[l + 1];
                    const bool last = (l + 2 == p.n_layers);
                    for (int j = g.sub; j < nout; j += LPR) {
                        double acc = 0.0;
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
                        hout[j * NR] = last ? acc : activate(p.activation, acc, s_tanh);
                    }
                    __syncwarp();
                    double* tsw = hin; hin = hout; hout = tsw;
                    woff += nin * nout;
                }
                v = hin[0];
                if (v > 180.0) v = R2P2_SUB(v, 360.0);
                w = hin[NR];
                __syncwarp();
            } else if (p.ctrl_kind == R2P2_CTRL_LINEAR) {
                const int n = p.G / 2 - 1;
                const int m = n < R ? n : R;
                double lin = 0.0, angv = 0.0;
                for (int i = 0; i < m; ++i) lin = R2P2_ADD(lin, R2P2_MUL(gen[i], bufA[i * NR]));
                for (int i = 0; i < m; ++i) angv = R2P2_ADD(angv, R2P2_MUL(gen[n + i], bufA[i * NR]));
                v = lin; w = angv;
                __syncwarp();
            } else if (pid) {
                // scalar per robot, no collectives inside: robots that have stopped skip it; lane 0 of a group writes the goal stack
                bool ok = true;
                v = 0.0; w = 0.0;
                if (live) ok = pid_control(p, sv, ps, rb, g.sub == 0, x, y, theta, speed, bufA, NR, gen, (size_t)1, s_tab, flags, v, w);
                __syncwarp();                              // bufA is rewritten by the next tick's sense; a pushed goal is visible to the group
                if (!ok) { flags |= R2P2_F_FAULT; live = false; }
            } else {
                v = gen[0]; w = gen[1];
            }
            if (tracing && live && g.sub == 0) {           // the controller's return value, as its log_step records it
                double* ci = p.tr_cinfo + trow * kCinfoRow;
                ci[4] = v; ci[5] = w;
            }
            if (live) {
                if (r2p2_fabs(v) > p.max_speed) v = R2P2_MUL(R2P2_DIV(v, r2p2_fabs(v)), p.max_speed);   // robot.py:386-387
                speed = v; omega = w;
                theta = R2P2_ADD(theta, R2P2_MUL(omega, delta));                                            // robot.py:257-267
            }

            // ---- integrate: __update_position (robot.py:188-255) -----------------------------------------
            const SinCos hsc = sincos_call(R2P2_MUL(live ? theta : 0.0, kRad), s_tab);
            if (live && !hsc.ok) { flags |= R2P2_F_TRIGRANGE | R2P2_F_FAULT; live = false; }
            const double cx = R2P2_ADD(x, R2P2_MUL(R2P2_MUL(hsc.c, speed), delta));
            const double cy = R2P2_ADD(y, R2P2_MUL(R2P2_MUL(hsc.s, speed), delta));
            const double dx = R2P2_SUB(cx, x), dy = R2P2_SUB(cy, y);
            double angle = R2P2_MUL(r2p2_atan2(live ? dy : 0.0, live ? dx : 1.0), kDeg);
            if (angle != 0.0) { if (angle < 0.0) angle = R2P2_ADD(angle, 360.0); } else angle = 0.0;   // Python float % 360
            const SinCos csc = sincos_call(R2P2_MUL(angle, kRad), s_tab);
            if (live && !csc.ok) { flags |= R2P2_F_TRIGRANGE | R2P2_F_FAULT; live = false; }   // NaN displacement: the reference faults two statements later
            const double climit = R2P2_ADD(r2p2_norm2(dx, dy), p.radius);
            Hit ms;
            ms.x = -1.0; ms.y = -1.0; ms.k = -1; ms.nsamp = 0u; ms.fault = false;
            const bool must_march = live && !ray_is_clear(sv, x, y, climit, ms.nsamp);
            if (__any_sync(FULL, must_march)) {           // robots in open space idle through it with a zero-length plan
                // a limit without a norm window (NaN, huge) takes the plain loop; it has no collectives, so it runs after
                // the team march, which every lane of the warp has to go through together
                const bool windowed = climit > -1.0 && climit < 30000.0;
                const bool teamed = must_march && windowed;
                MarchPlan cplan = make_plan(teamed ? climit : 0.0, LPR);
                if (!teamed) { cplan.kA = 0; cplan.kB = 0; cplan.chunk = 0; }
                const Hit mh = team_march_t<LPR, kWsCollOwn<LPR>, kWsWin>(kWsWin ? svw : sv, coll_team, cplan, x, y, csc.c, csc.s, climit);
                if (teamed) ms = mh;
                else if (must_march) ms = march_plain(sv, x, y, csc.c, csc.s, climit);
            }
            if (live) {
                ns_coll += ms.nsamp;
                if (ms.fault) { flags |= R2P2_F_FAULT; live = false; }
            }
            if (live) {
                x = cx; y = cy;
                if (ms.k >= 0) {
                    if (!(isfinite(cx) && isfinite(cy))) { flags |= R2P2_F_FAULT; live = false; }   // int(inf) raises
                    else {
                        bool f = false;
                        const bool inwall = px_probe(sv.wall, sv, dtrunc(cx), dtrunc(cy), f);
                        if (f) { flags |= R2P2_F_FAULT; live = false; }
                        else if (inwall) {
                            ncollide++; tmr = 0.0;
                            ccode |= 1u; ccx = x; ccy = y;
                            const double ex = R2P2_SUB(ms.x, x), ey = R2P2_SUB(ms.y, y);
                            if (ex > 0.0) x = R2P2_ADD(ms.x, p.rad_p1); else if (ex < 0.0) x = R2P2_SUB(ms.x, p.rad_p1);
                            if (ey > 0.0) y = R2P2_ADD(ms.y, p.rad_p1); else if (ey < 0.0) y = R2P2_SUB(ms.y, p.rad_p1);
                        }
                    }
                }
            }
            if (live) {
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
                // env[int(x), int(y)] is 50 or __crashing_radius (robot.py:249, 181-186)
                if (!(isfinite(x) && isfinite(y))) { flags |= R2P2_F_FAULT; live = false; }
            }
            bool crash = false, need_ring = false;
            if (live) {
                bool f = false;
                if (p.has50) crash = px_probe(sv.lvl50, sv, dtrunc(x), dtrunc(y), f);
                else { int ix = dtrunc(x), iy = dtrunc(y); if (ix < 0) ix += p.W; if (iy < 0) iy += p.H; f = ((unsigned)ix >= (unsigned)p.W) | ((unsigned)iy >= (unsigned)p.H); }
                if (f) { flags |= R2P2_F_FAULT; live = false; }
                else if (!crash) {
                    // far enough from every blocked pixel: none of the 360 probes can be a wall or leave the map
                    bool ring_clear = false;
                    const int cxp = dtrunc(x), cyp = dtrunc(y);
                    if (((unsigned)cxp < (unsigned)p.W) & ((unsigned)cyp < (unsigned)p.H))
                        ring_clear = (int)__ldg(p.g_dist + (size_t)cyp * p.W + cxp) >= p.crash_clear;
                    need_ring = !ring_clear;
                }
            }
            // 360 probes; the reference stops at the first wall, so an IndexError only counts if it comes before the first
            // wall probe (ring_scan's key).
            {
                // The whole warp serves one robot at a time: 12 probes per lane and one REDUX, instead of 360 / LPR probes
                // per lane of the group while the other groups idle (8,192 robots: 12.6 -> 11.9 ms).
                unsigned todo = __ballot_sync(FULL, need_ring && g.sub == 0);
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const double rx = __shfl_sync(FULL, x, leader), ry = __shfl_sync(FULL, y, leader);
                    StageView svr = sv;                      // that robot's window
                    svr.win_ww = p.win_ww;
                    svr.hx0 = __shfl_sync(FULL, svw.hx0, leader); svr.hy0 = __shfl_sync(FULL, svw.hy0, leader);
                    svr.hx1 = __shfl_sync(FULL, svw.hx1, leader); svr.hy1 = __shfl_sync(FULL, svw.hy1, leader);
                    svr.win_sb = __shfl_sync(FULL, svw.win_sb, leader);
                    const unsigned key = __reduce_min_sync(FULL, ring_scan<32, kWsWin>(svr, (int)lane, rx, ry, p.radius, s_crx, s_cry));
                    if ((int)(lane & ~(unsigned)(LPR - 1)) == leader && key != 0xffffffffu) {   // the lanes of that robot's group
                        if (key & 1u) crash = true;
                        else { flags |= R2P2_F_FAULT; live = false; }
                    }
                }
            }
            if (live && crash) { x = last_x; y = last_y; ncollide++; tmr = 0.0; ccode |= 1u << 5; }
            if (pid && ncollide != ncollide0) ps.backing = 10;      // on_collision (pid_controller.py:166-167), also in a tick the reference raised in
            if (live) {
                if (!crash) { last_x = x; last_y = y; }
                ticks++; tdone++;
                if (ncollide != ncollide0) flags |= R2P2_F_COLLIDED;
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
            alive = live;
        }
        if (valid) {
            if (ncollide != ncollide_in) flags |= R2P2_F_COLLIDED;
            if (tracing && g.sub == 0) {     // rows of ticks that did not complete: frozen pose (as the oracle records them)
                for (int t2 = tdone; t2 < p.T; ++t2) {
                    const size_t trow = (size_t)t2 * (size_t)p.P + rb;
                    double* q = p.tr_pose + trow * 5;
                    q[0] = x; q[1] = y; q[2] = theta; q[3] = speed; q[4] = omega;
                    p.tr_ncoll[trow] = (t2 == tdone) ? (ncollide - ncollide0) : 0u;
                }
            }
            if (!(flags & R2P2_F_DONE))      // controller.fitness() after the last update
                fitness = R2P2_ADD(dist, r2p2_norm2(R2P2_SUB(odom_x, origin_x), R2P2_SUB(odom_y, origin_y)));
        }
        const unsigned ns_sonar = g.sum_u(my_ns_sonar);
        if (p.noise_mode == 1 && ndraws > (unsigned)p.noise_len) flags |= R2P2_F_NOISE;   // the replayed stream ran out
        if (valid && g.sub == 0) {
            p.x[rb] = x; p.y[rb] = y; p.theta[rb] = theta; p.speed[rb] = speed; p.omega[rb] = omega;
            p.last_x[rb] = last_x; p.last_y[rb] = last_y;
            p.dist[rb] = dist; p.odom_x[rb] = odom_x; p.odom_y[rb] = odom_y;
            p.time[rb] = tmr; p.fitness[rb] = fitness;
            p.flags[rb] = flags; p.ncollide[rb] = ncollide; p.ticks[rb] = ticks;
            p.nsamp_sonar[rb] = __ldcg(p.nsamp_sonar + rb) + ns_sonar; p.nsamp_coll[rb] = __ldcg(p.nsamp_coll + rb) + ns_coll;
            if (p.noise_mode) p.ndraws[rb] = ndraws;
            if ((EXTRA && p.goal_on)) p.goal_tick[rb] = goal_tick;
            if (pid) pid_store(p, rb, ps);
        }
        __syncwarp();
        if (n_slices > 1u) {                               // this batch's state is in L2: its next slice may start
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(p.slice_done + batch, slice + 1u);
        }
    }
}

}  // namespace r2p2
