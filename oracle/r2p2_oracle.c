/* r2p2_oracle.c -- CPU restatement of the r2p2 per-tick robot step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: plain C, one robot at a time, written to follow the reference
 * statement by statement.  It is never linked into, imported by or called from the product
 * (r2p2_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it.
 *
 * Reference lines followed (paths under /root/reference):
 *   r2p2/utils.py:420-432    search_edge_in_angle_with_limit      -> orc_march
 *   r2p2/utils.py:456-476,497-514  ray fans (int form / list form) -> orc_sense (angles expanded by caller)
 *   r2p2/robot.py:287-318    Robot.check_sensors                  -> orc_sense
 *   r2p2/robot.py:373-387    Robot.control (+ speed clamp)        -> orc_tick
 *   r2p2/controllers/neurocontroller.py:55-60,67-95,129-131,139-160 -> orc_control (kind NEURO)
 *   r2p2/controllers/inspyred.py:60-100,122-124,132-144            -> orc_control (kind LINEAR)
 *   r2p2/controllers/naive_controller.py:35-55 (constant command)  -> orc_control (kind CONST)
 *   r2p2/robot.py:257-267    __update_angle / turn                 -> orc_tick
 *   r2p2/robot.py:188-255    __update_position                     -> orc_move
 *   r2p2/robot.py:181-186    __crashing_radius                     -> orc_crashing
 *   r2p2/robot.py:170-179    collide -> controller.on_collision (time = 0)
 *
 * Third-party arithmetic the reference reaches through its dependencies (none vendored):
 *   glibc 2.39 libm sin/cos/atan2 via CPython `math`      -> called directly here (libm mode)
 *   numpy/OpenBLAS `np.linalg.norm` of a 2-vector          -> sqrt(fma(dy,dy, dx*dx))  
 *   scikit-learn MLP forward (h @ W, activation)           -> sequential-FMA dot + libm tanh; numpy's
 *       BLAS summation order and SIMD tanh differ from this at ulp level, only the 1e-5 tolerance applies.
 *
 * Pinning: the reference ships no tests; this oracle is pinned against golden
 * vectors recorded from the unmodified reference in the build container (tests/golden/, made by
 * tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.
 *
 * Two build flavours (oracle/Makefile):
 *   libr2p2_oracle.so       "libm mode": libm sin/cos/atan2/tanh -- the faithful restatement.
 *   libr2p2_oracle_port.so  -DORC_PORTMATH: trig/tanh come from r2p2_b200/csrc/r2p2_math.h, the
 *       header the CUDA kernels use, so GPU results can be compared bit for bit at any size.  The
 *       two flavours agree exactly on sin/cos (tests/test_math_port.py) and to <= 3 ulp on atan2/tanh.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_PORTMATH
#include "../r2p2_b200/csrc/r2p2_math.h"
#define ORC_SIN(x, okp) r2p2_sin((x), r2p2_h_sincostab, (okp))
#define ORC_COS(x, okp) r2p2_cos((x), r2p2_h_sincostab, (okp))
#define ORC_ATAN2(y, x) r2p2_atan2((y), (x))
#define ORC_TANH(x) r2p2_tanh((x))
#define ORC_LOGISTIC(x) r2p2_logistic((x))
#else
#define ORC_SIN(x, okp) sin((x))
#define ORC_COS(x, okp) cos((x))
#define ORC_ATAN2(y, x) atan2((y), (x))
#define ORC_TANH(x) tanh((x))
#define ORC_LOGISTIC(x) (1.0 / (1.0 + exp(-(x))))
#endif

#define ORC_MAX_RAYS 64
#define ORC_MAX_LAYERS 8
#define ORC_MAX_WIDTH 64

enum { ORC_CTRL_CONST = 0, ORC_CTRL_NEURO = 1, ORC_CTRL_LINEAR = 2, ORC_CTRL_PID = 3 };
enum { ORC_ACT_IDENTITY = 0, ORC_ACT_TANH = 1, ORC_ACT_RELU = 2, ORC_ACT_LOGISTIC = 3 };

/* per-robot flag bits (same meaning as include/r2p2_b200.h) */
#define ORC_F_COLLIDED 1u   /* collide() was called at least once            */
#define ORC_F_FAULT 2u      /* the reference would have raised (IndexError / ValueError / OverflowError) */
#define ORC_F_DONE 4u       /* evolve: the episode ended (timer or collision) */
#define ORC_F_GOAL 16u      /* the goal predicate held after some tick */
#define ORC_F_NOISE 32u     /* the replayed noise stream ran out: later draws were 0.0 */
#define ORC_F_TRIGRANGE 8u  /* |angle| beyond the supported sin/cos range (port mode only) */

typedef struct {
    int32_t W, H;              /* env.shape == (W, H); env[x*H + y] */
    const int32_t* env;
    int32_t R;                 /* number of sonar rays */
    double ray_deg[ORC_MAX_RAYS];
    double vr0, vr1, radius, max_speed;
    int32_t ctrl_kind;
    int32_t n_layers;          /* NEURO: sizes incl. input (R) and output (2) */
    int32_t layers[ORC_MAX_LAYERS];
    int32_t activation;
    int32_t G;                 /* genome doubles per robot */
    int32_t evolve;            /* episode semantics of neurocontroller.py:72-85 */
    double delta;              /* Robot.update(env, delta)                       */
    double delta_ctrl;         /* u.delta seen by the controller timer            */
    double time_budget;        /* controller self.time at episode start           */
    const double* noise;       /* has_noise (robot.py:96,308-310): per-robot stream [P][noise_len] of N(0, 0.5) draws, consumed
                                  one per unclamped ray in ray order exactly as np.random.normal is called; NULL = noise off */
    int32_t noise_len;
    int32_t goal_on;           /* goal flagging: Sequential_PID_Controller.is_at_goal (controllers/pid_controller.py:207-213) as a */
    double goal_x, goal_y;     /* per-run predicate on the pose after each tick; a flag only, the dynamics never see it          */
    /* Sequential_PID_Controller (controllers/pid_controller.py:8-262): genome = [ap, ai, ad, lp, li, ld] per robot */
    double pid_max_acc;        /* self.max_acceleration (5)  */
    double pid_accel;          /* robot.acceleration (0: nothing in the tick changes it, robot.py:398 is commented out) */
    void* pid_state;           /* orc_pid[P]: PID terms, obstacle-avoid state, goal stack */
} orc_cfg;

#define ORC_PID_GCAP 64        /* goal stack capacity (configured goals + temporary avoid goals); overflow = fault */
typedef struct {
    double acc_ang, acc_dist, last_ang, last_dist, target_angle;
    int32_t state, backing, ngoals, pad;
    double gx[ORC_PID_GCAP], gy[ORC_PID_GCAP];   /* a stack: slot ngoals-1 is goal[0] of the reference's list */
} orc_pid;

typedef struct {
    double x, y, theta, speed, omega;
    double last_x, last_y;
    double dist, odom_x, odom_y, origin_x, origin_y;
    double time;
    double fitness;
    uint32_t flags, ncollide, ticks;
    uint64_t nsamp_sonar, nsamp_coll;
    uint64_t npixel;           /* every env.item() the reference would have executed */
    uint32_t ndraws;           /* noise draws consumed so far */
    uint32_t goal_tick;        /* ticks completed when the goal predicate first held (0xffffffff: never) */
} orc_robot;

static const double ORC_RAD = 3.14159265358979323846 / 180.0;  /* CPython degToRad */
static const double ORC_DEG = 180.0 / 3.14159265358979323846;  /* CPython radToDeg */

static inline double orc_norm2(double dx, double dy) { return sqrt(fma(dy, dy, dx * dx)); }

/* Python round(float) -> int: half to even; raises on nan/inf. */
static inline int orc_pyround(double v, int64_t* out) {
    if (!isfinite(v)) return 0;
    *out = (int64_t)rint(v);
    return 1;
}

/* env.item(ix, iy) with Python index rules. returns 0 on IndexError */
static inline int orc_px(const orc_cfg* c, int64_t ix, int64_t iy, int32_t* v) {
    if (ix < 0) ix += c->W;
    if (iy < 0) iy += c->H;
    if (ix < 0 || ix >= c->W || iy < 0 || iy >= c->H) return 0;
    *v = c->env[ix * (int64_t)c->H + iy];
    return 1;
}

/* int(float): truncation; raises on nan/inf */
static inline int orc_pyint(double v, int64_t* out) {
    if (!isfinite(v)) return 0;
    *out = (int64_t)v;   /* |v| is far below 2^63 for anything in-map; huge values fault at orc_px anyway */
    if (fabs(v) >= 9.0e18) *out = (v > 0) ? INT64_MAX / 2 : INT64_MIN / 2;
    return 1;
}

typedef struct { int hit; int fault; double x, y; int32_t k; uint32_t nsamp; } orc_hit;

/* utils.py:420-432.  (ox, oy) may be integers (sonar) or floats (collision ray). */
static orc_hit orc_march(const orc_cfg* c, double ox, double oy, double angle_rad, double limit, uint32_t* flags) {
    orc_hit r; memset(&r, 0, sizeof r); r.k = -1;
    int ok = 1;
    const double vx = ORC_COS(angle_rad, &ok), vy = ORC_SIN(angle_rad, &ok);
    if (!ok) { *flags |= ORC_F_TRIGRANGE | ORC_F_FAULT; r.fault = 1; return r; }
    double x = ox, y = oy;
    int32_t k = 0;
    for (;;) {
        int64_t rx, ry;
        if (!orc_pyround(x, &rx) || !orc_pyround(y, &ry)) { r.fault = 1; *flags |= ORC_F_FAULT; return r; }
        if (!(rx >= 0 && rx < c->W && ry >= 0 && ry < c->H)) break;
        if (!(orc_norm2(x - ox, y - oy) < limit)) break;
        int64_t ix, iy; int32_t v;
        orc_pyint(x, &ix); orc_pyint(y, &iy);
        r.nsamp++;
        if (!orc_px(c, ix, iy, &v)) { r.fault = 1; *flags |= ORC_F_FAULT; return r; }
        if (v == 0) { r.hit = 1; r.x = x; r.y = y; r.k = k; return r; }
        x = x + vx; y = y + vy; k++;
    }
    return r;
}

typedef struct {             /* optional per-tick trace rows for one robot (all may be NULL) */
    double* dst;             /* [R]  post-processed distances handed to the controller */
    int32_t* hitk;           /* [R]  hit step or -1 */
    int32_t* hitpx;          /* [R][2] hit pixel (int(x), int(y)) or -1,-1 */
    uint32_t* nsamp;         /* [R]  samples tested on the ray */
    double* cinfo;       /* [4]: collide() calls of the tick: code, wall spot x, y, border coordinate (include/r2p2_b200.h) */
} orc_trace_row;

/* robot.py:287-318 */
static int orc_sense(const orc_cfg* c, orc_robot* rb, double* dst, const orc_trace_row* tr, const double* noise_stream) {
    int64_t ox, oy;
    if (!orc_pyround(rb->x, &ox) || !orc_pyround(rb->y, &oy)) { rb->flags |= ORC_F_FAULT; return 0; }
    const double L = c->vr1 + c->radius;
    for (int i = 0; i < c->R; ++i) {
        const double ang = c->ray_deg[i] + rb->theta;
        orc_hit h = orc_march(c, (double)ox, (double)oy, ang * ORC_RAD, L, &rb->flags);
        rb->nsamp_sonar += h.nsamp; rb->npixel += h.nsamp;
        if (tr && tr->nsamp) tr->nsamp[i] = h.nsamp;
        if (h.fault) return 0;
        double d = h.hit ? orc_norm2(h.x - (double)ox, h.y - (double)oy) : INFINITY;
        if (tr && tr->hitk) tr->hitk[i] = h.k;
        if (tr && tr->hitpx) {
            tr->hitpx[2 * i] = h.hit ? (int32_t)(int64_t)h.x : -1;
            tr->hitpx[2 * i + 1] = h.hit ? (int32_t)(int64_t)h.y : -1;
        }
        if ((d < c->vr0 + c->radius && d > c->vr0 * 0.125 + c->radius) || d > c->vr1 + c->radius) d = c->vr1 + c->radius;
        else if (noise_stream) {
            if (rb->ndraws < (uint32_t)c->noise_len) d += noise_stream[rb->ndraws]; else { d += 0.0; rb->flags |= ORC_F_NOISE; }
            rb->ndraws++;
        }
        dst[i] = d;
        if (tr && tr->dst) tr->dst[i] = d;
    }
    return 1;
}

static inline double orc_act(int kind, double v) {
    switch (kind) {
        case ORC_ACT_TANH: return ORC_TANH(v);
        case ORC_ACT_RELU:   /* sklearn: np.clip(x, 0, finfo.max): NaN stays NaN, +inf becomes DBL_MAX */
            if (v != v) return v;
            if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
            return v > 0.0 ? v : 0.0;
        case ORC_ACT_LOGISTIC: return ORC_LOGISTIC(v);
        default: return v;
    }
}

static inline double orc_fitness(const orc_robot* rb) {   /* neurocontroller.py:129-131 */
    return rb->dist + orc_norm2(rb->odom_x - rb->origin_x, rb->odom_y - rb->origin_y);
}

/* controller.control(dst).  returns 0 when the episode ended on this call (evolve). */
static int orc_control_pid(const orc_cfg* c, orc_robot* rb, orc_pid* ps, const double* k, const double* dst, double* v_out, double* w_out);
static int orc_control(const orc_cfg* c, orc_robot* rb, const double* genome, const double* dst, double* v_out, double* w_out, orc_pid* ps) {
    rb->dist += orc_norm2(rb->x - rb->odom_x, rb->y - rb->odom_y);
    rb->odom_x = rb->x; rb->odom_y = rb->y;
    if (c->ctrl_kind == ORC_CTRL_PID) return orc_control_pid(c, rb, ps, genome, dst, v_out, w_out);
    if (c->ctrl_kind != ORC_CTRL_CONST) {
        if (c->evolve) rb->time -= c->delta_ctrl;
        if (rb->time <= 0 && c->evolve) {
            rb->fitness = orc_fitness(rb);
            rb->flags |= ORC_F_DONE;
            *v_out = 0; *w_out = 0;
            return 0;
        }
    }
    if (c->ctrl_kind == ORC_CTRL_NEURO) {
        double h[ORC_MAX_WIDTH], o[ORC_MAX_WIDTH];
        for (int i = 0; i < c->R; ++i) h[i] = c->vr1 / dst[i];
        const double* w = genome;
        for (int l = 0; l + 1 < c->n_layers; ++l) {
            const int nin = c->layers[l], nout = c->layers[l + 1];
            for (int j = 0; j < nout; ++j) {
                double acc = 0.0;
                for (int i = 0; i < nin; ++i) acc = fma(h[i], w[i * nout + j], acc);
                o[j] = (l + 2 < c->n_layers) ? orc_act(c->activation, acc) : acc;
            }
            memcpy(h, o, sizeof(double) * (size_t)nout);
            w += nin * nout;
        }
        double ang = h[0];
        if (ang > 180) ang = ang - 360;
        *v_out = ang;     /* Robot.control unpacks (speed, angular_velocity) = (out0', out1): SURVEY H8 */
        *w_out = h[1];
    } else if (c->ctrl_kind == ORC_CTRL_LINEAR) {
        /* chromosome [l_1..l_n, a_1..a_n, vmaxlin, vmaxang], n = G/2 - 1; zip truncates to min(n, R) */
        const int n = c->G / 2 - 1;
        const int m = n < c->R ? n : c->R;
        double lin = 0, ang = 0;
        for (int i = 0; i < m; ++i) lin = lin + genome[i] * dst[i];
        for (int i = 0; i < m; ++i) ang = ang + genome[n + i] * dst[i];
        *v_out = lin; *w_out = ang;
    } else {
        *v_out = genome[0]; *w_out = genome[1];
    }
    return 1;
}

static inline void orc_collide(orc_robot* rb) { rb->ncollide++; rb->flags |= ORC_F_COLLIDED; rb->time = 0; }

/* ---- Sequential_PID_Controller.control (controllers/pid_controller.py:40-262) ----------------------------------- */
static inline double orc_pymod360(double a) {          /* Python float % 360 */
    double m = fmod(a, 360.0);
    if (m != 0) { if (m < 0) m += 360.0; } else m = 0.0;
    return m;
}
/* Robot.brake (robot.py:354-363) */
static inline double orc_brake(const orc_cfg* c, orc_robot* rb) {
    const int neg = rb->speed < 0;
    const double new_speed = rb->speed - c->pid_accel - rb->speed / 3;
    if ((neg && new_speed > 0) || (!neg && new_speed < 0)) { rb->speed = 0; return -c->pid_accel / 5; }
    return -c->pid_accel - rb->speed / 3;
}
static inline double orc_pid_angle_variation(orc_robot* rb, orc_pid* ps, const double* k) {   /* :235-243 */
    const double e = ps->target_angle - rb->theta;
    ps->acc_ang += e;
    const double de = e - ps->last_ang;
    ps->last_ang = e;
    return k[0] * e + k[1] * ps->acc_ang + k[2] * de;
}
/* returns 0 on a fault (IndexError at the goal pixel, goal stack overflow) */
static int orc_control_pid(const orc_cfg* c, orc_robot* rb, orc_pid* ps, const double* k, const double* dst, double* v_out, double* w_out) {
    const int R = c->R;
    if (ps->backing) { ps->backing -= 1; *v_out = rb->theta; *w_out = -c->max_speed; return 1; }      /* :50-53 */
    {   /* manage_state :215-229 */
        const double rng = c->vr1, tol = c->vr0 / rng * 13;
        if ((dst[R - 1] / rng < tol && dst[R - 2] / rng < tol) || (dst[1] / rng < tol && dst[2] / rng < tol) || dst[0] / rng < tol)
            ps->state = (ps->state == 0) ? 1 : 2;
        else ps->state = 0;
    }
    if (ps->state == 0 || ps->state == 2) {             /* control_advance :60-72 */
        if (ps->ngoals == 0) { *v_out = 0; *w_out = 0; return 1; }
        const double gx = ps->gx[ps->ngoals - 1], gy = ps->gy[ps->ngoals - 1];
        ps->target_angle = ORC_ATAN2(gy - rb->y, gx - rb->x) * ORC_DEG;                               /* :245-250 */
        const double angle = orc_pid_angle_variation(rb, ps, k);
        const double speed0 = rb->speed;
        double var;                                      /* calculate_acceleration_variation :219-233 */
        if (fabs(ps->target_angle - rb->theta) > 45) {
            ps->acc_dist += orc_brake(c, rb);
            var = orc_brake(c, rb);
        } else {
            const double e = orc_norm2(gx - rb->x, gy - rb->y);
            if (c->pid_accel < c->pid_max_acc) ps->acc_dist += e;
            const double de = e - ps->last_dist;
            ps->last_dist = e;
            var = k[3] * e + k[4] * ps->acc_dist + k[5] * de;
        }
        double acc = c->pid_accel + var;                 /* choose_acceleration :143-155 */
        if (acc > c->pid_max_acc) acc = c->pid_max_acc;
        else if (-acc < c->pid_max_acc) acc = -c->pid_max_acc;
        const double speed = speed0 + acc;
        {   /* is_at_goal :207-213 */
            int64_t ix, iy; int32_t v = 1;
            rb->npixel++;
            if (!orc_pyint(gx, &ix) || !orc_pyint(gy, &iy) || !orc_px(c, ix, iy, &v)) { rb->flags |= ORC_F_FAULT; return 0; }
            int at = (v == 0);
            if (!at) at = orc_norm2(gx - rb->x, gy - rb->y) <= c->radius * 1.85;
            if (at) { ps->ngoals--; if (ps->state == 2) ps->state = 0; }
        }
        *v_out = speed; *w_out = angle;
        return 1;
    }
    /* control_avoid :73-124 */
    const double far = c->vr1;
    double risk = 1;
    for (int i = 0; i < R; ++i) risk *= (dst[i] / far);
    if (dst[1] + dst[2] < dst[R - 1] + dst[R - 2]) risk /= (far / dst[R - 1]) * (far / dst[R - 2]);
    else risk /= (far / dst[1]) * (far / dst[2]);
    if (dst[0] / far < 0.45) risk *= dst[0] / far;
    double safety = (far + c->radius) * risk * 2;
    if (safety <= c->vr0) safety = c->vr0 * 2;
    double a;
    if ((dst[1] + dst[2]) > (dst[R - 1] + dst[R - 2])) a = orc_pymod360(ps->target_angle) + 90;
    else a = orc_pymod360(ps->target_angle) - 90;
    int ok = 1;
    const double arg = (ps->target_angle - a) * ORC_RAD;
    double x_var = safety * ORC_COS(arg, &ok), y_var = safety * ORC_SIN(arg, &ok);
    if (!ok) { rb->flags |= ORC_F_TRIGRANGE | ORC_F_FAULT; return 0; }
    const double am = orc_pymod360(a);
    if (am > 90 && am < 270) x_var *= -1;
    if (am > 180) y_var *= -1;
    if (ps->ngoals >= ORC_PID_GCAP) { rb->flags |= ORC_F_FAULT; return 0; }
    ps->gx[ps->ngoals] = rb->x + x_var; ps->gy[ps->ngoals] = rb->y + y_var; ps->ngoals++;             /* goal.insert(0, ...) */
    (void)orc_pid_angle_variation(rb, ps, k);
    ps->target_angle = a;
    int must_retreat = 0;
    for (int i = 0; i < R; ++i) if (dst[i] / far < c->vr0 / far * 12) { must_retreat = 1; break; }
    *v_out = must_retreat ? -c->max_speed : rb->speed;
    *w_out = 0;
    return 1;
}

/* robot.py:181-186 ; returns 1 crashing, 0 clear, -1 fault */
static int orc_crashing(const orc_cfg* c, orc_robot* rb, const double* crx, const double* cry) {
    for (int i = 0; i < 360; ++i) {
        int64_t ix, iy; int32_t v;
        rb->npixel++;
        if (!orc_pyint(rb->x + crx[i], &ix) || !orc_pyint(rb->y + cry[i], &iy)) return -1;
        if (!orc_px(c, ix, iy, &v)) return -1;
        if (v == 0) return 1;
    }
    return 0;
}

/* robot.py:188-255 */
static int orc_move(const orc_cfg* c, orc_robot* rb, const double* crx, const double* cry, double* ci) {
    double ci_local[4];
    if (!ci) ci = ci_local;
    ci[0] = ci[1] = ci[2] = ci[3] = 0.0;
    unsigned code = 0;
    const double W = (double)c->W, H = (double)c->H, rad = c->radius, delta = c->delta;
    int ok = 1;
    const double hr = rb->theta * ORC_RAD;
    const double cx = rb->x + ORC_COS(hr, &ok) * rb->speed * delta;
    const double cy = rb->y + ORC_SIN(hr, &ok) * rb->speed * delta;
    if (!ok) { rb->flags |= ORC_F_TRIGRANGE | ORC_F_FAULT; return 0; }
    const double dx = cx - rb->x, dy = cy - rb->y;
    double angle = ORC_ATAN2(dy, dx) * ORC_DEG;
    {   /* Python float %: fmod, then sign fix-up */
        double m = fmod(angle, 360.0);
        if (m != 0) { if (m < 0) m += 360.0; } else m = 0.0;
        angle = m;
    }
    orc_hit ms = orc_march(c, rb->x, rb->y, angle * ORC_RAD, orc_norm2(dx, dy) + rad, &rb->flags);
    rb->nsamp_coll += ms.nsamp; rb->npixel += ms.nsamp;
    if (ms.fault) return 0;
    rb->x = cx; rb->y = cy;
    if (ms.hit) {
        int64_t ix, iy; int32_t v;
        rb->npixel++;
        if (!orc_pyint(cx, &ix) || !orc_pyint(cy, &iy) || !orc_px(c, ix, iy, &v)) { rb->flags |= ORC_F_FAULT; return 0; }
        if (v == 0) {
            orc_collide(rb);
            code |= 1u; ci[1] = rb->x; ci[2] = rb->y;
            if (ms.x - rb->x > 0) rb->x = ms.x + (rad + 1);
            else if (ms.x - rb->x < 0) rb->x = ms.x - (rad + 1);
            if (ms.y - rb->y > 0) rb->y = ms.y + (rad + 1);
            else if (ms.y - rb->y < 0) rb->y = ms.y - (rad + 1);
        }
    }
    if (rb->x + rad >= W) {
        ci[3] = rb->y;
        if (rb->y + rad < H && rb->y - rad >= 0) { orc_collide(rb); code |= 1u << 1; }
        else if (rb->y + rad >= H) { orc_collide(rb); code |= 2u << 1; rb->y = H - 1 - rad; }
        else { orc_collide(rb); code |= 3u << 1; rb->y = rad; }
        rb->x = W - 1;
    } else if (rb->x - rad < 0) {
        ci[3] = rb->y;
        if (rb->y + rad < H && rb->y - rad >= 0) { orc_collide(rb); code |= 4u << 1; }
        else if (rb->y + rad >= H) { orc_collide(rb); code |= 5u << 1; rb->y = H - 1 - rad; }
        else { orc_collide(rb); code |= 6u << 1; rb->y = rad; }
        rb->x = 0;
    } else if (rb->y + rad > H) {
        ci[3] = rb->x;
        orc_collide(rb); code |= 7u << 1; rb->y = H - 1 - rad;
    } else if (rb->y - rad < 0) {
        ci[3] = rb->x;
        orc_collide(rb); code |= 8u << 1; rb->y = rad;
    }
    ci[0] = (double)code;
    {
        int64_t ix, iy; int32_t v;
        rb->npixel++;
        if (!orc_pyint(rb->x, &ix) || !orc_pyint(rb->y, &iy) || !orc_px(c, ix, iy, &v)) { rb->flags |= ORC_F_FAULT; return 0; }
        int crash = (v == 50);
        if (!crash) {
            crash = orc_crashing(c, rb, crx, cry);
            if (crash < 0) { rb->flags |= ORC_F_FAULT; return 0; }
        }
        if (crash) { rb->x = rb->last_x; rb->y = rb->last_y; orc_collide(rb); code |= 1u << 5; ci[0] = (double)code; }
        else { rb->last_x = rb->x; rb->last_y = rb->y; }
    }
    return 1;
}

/* Robot.update (robot.py:389-402).  returns 1 if the robot keeps running. */
static int orc_tick(const orc_cfg* c, orc_robot* rb, const double* genome, const double* crx, const double* cry,
                    const orc_trace_row* tr, const double* noise_stream, orc_pid* ps) {
    double dst[ORC_MAX_RAYS];
    if (!orc_sense(c, rb, dst, tr, noise_stream)) return 0;
    double v, w;
    const uint32_t ncoll_before = rb->ncollide;
    if (!orc_control(c, rb, genome, dst, &v, &w, ps)) {
        if (c->ctrl_kind != ORC_CTRL_PID) { rb->speed = 0; rb->omega = 0; }
        return 0;
    }
    if (tr && tr->cinfo) { tr->cinfo[4] = v; tr->cinfo[5] = w; }        /* what the controller's log_step records (neurocontroller.py:94-98) */
    if (fabs(v) > c->max_speed) v = v / fabs(v) * c->max_speed;
    rb->speed = v; rb->omega = w;
    rb->theta = rb->theta + rb->omega * c->delta;
    const int moved = orc_move(c, rb, crx, cry, tr ? tr->cinfo : NULL);
    if (ps && rb->ncollide != ncoll_before) ps->backing = 10;          /* on_collision (pid_controller.py:166-167) */
    if (!moved) return 0;
    rb->ticks++;
    if (c->goal_on && !(rb->flags & ORC_F_GOAL)) {
        /* pid_controller.py:210-213: `npdata.item((int(gx), int(gy))) is 0` or norm(goal - pos) <= radius * 1.85 */
        int64_t gx, gy; int32_t v = 1;
        int at = 0;
        if (orc_pyint(c->goal_x, &gx) && orc_pyint(c->goal_y, &gy) && orc_px(c, gx, gy, &v) && v == 0) at = 1;
        if (!at) at = orc_norm2(c->goal_x - rb->x, c->goal_y - rb->y) <= c->radius * 1.85;
        if (at) { rb->flags |= ORC_F_GOAL; rb->goal_tick = rb->ticks; }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* exported API (ctypes)                                                                       */
/* ------------------------------------------------------------------------------------------- */

/* Raw ray (KAT1): returns 1 hit / 0 miss / -1 fault. */
int orc_ray(const int32_t* env, int32_t W, int32_t H, double ox, double oy, double angle_rad, double limit,
            double* hx, double* hy, int32_t* k, uint32_t* nsamp) {
    orc_cfg c; memset(&c, 0, sizeof c); c.W = W; c.H = H; c.env = env;
    uint32_t fl = 0;
    orc_hit h = orc_march(&c, ox, oy, angle_rad, limit, &fl);
    *hx = h.hit ? h.x : -1; *hy = h.hit ? h.y : -1; *k = h.k; *nsamp = h.nsamp;
    return h.fault ? -1 : h.hit;
}

void orc_init_robot(orc_robot* rb, double x, double y, double theta, double time_budget) {
    memset(rb, 0, sizeof *rb);
    rb->x = x; rb->y = y; rb->theta = theta;
    rb->last_x = x; rb->last_y = y;
    rb->odom_x = rb->origin_x = x; rb->odom_y = rb->origin_y = y;
    rb->time = time_budget;
    rb->goal_tick = 0xffffffffu;
}

typedef struct {
    const orc_cfg* c; orc_robot* robots; const double* genomes; int32_t p0, p1, P, T;
    double* tr_pose; double* tr_dst; int32_t* tr_hitk; int32_t* tr_hitpx; uint32_t* tr_nsamp; uint32_t* tr_ncoll; double* tr_cinfo;
} orc_job;

static void* orc_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    const orc_cfg* c = j->c;
    double crx[360], cry[360];
    for (int i = 0; i < 360; ++i) {   /* robot.py:183: math.cos(math.radians(i)) * self.radius */
        int ok = 1; (void)ok;
        crx[i] = ORC_COS(i * ORC_RAD, &ok) * c->radius;
        cry[i] = ORC_SIN(i * ORC_RAD, &ok) * c->radius;
    }
    const int R = c->R;
    for (int32_t p = j->p0; p < j->p1; ++p) {
        orc_robot* rb = &j->robots[p];
        const double* genome = j->genomes + (size_t)p * (size_t)c->G;
        for (int32_t t = 0; t < j->T; ++t) {
            const size_t row = (size_t)t * (size_t)j->P + (size_t)p;
            orc_trace_row tr;
            tr.dst = j->tr_dst ? j->tr_dst + row * R : NULL;
            tr.hitk = j->tr_hitk ? j->tr_hitk + row * R : NULL;
            tr.hitpx = j->tr_hitpx ? j->tr_hitpx + row * R * 2 : NULL;
            tr.nsamp = j->tr_nsamp ? j->tr_nsamp + row * R : NULL;
            tr.cinfo = j->tr_cinfo ? j->tr_cinfo + row * 6 : NULL;      /* code, wx, wy, b (orc_move); controller outputs (orc_tick) */
            const uint32_t ncoll0 = rb->ncollide;
            int alive = 0;
            if (!(rb->flags & (ORC_F_FAULT | ORC_F_DONE))) {
                const double* nz = c->noise ? c->noise + (size_t)p * (size_t)c->noise_len : NULL;
                orc_pid* ps = (c->ctrl_kind == ORC_CTRL_PID) ? ((orc_pid*)c->pid_state) + p : NULL;
                alive = orc_tick(c, rb, genome, crx, cry, &tr, nz, ps);
            }
            if (j->tr_pose) {
                double* q = j->tr_pose + row * 5;
                q[0] = rb->x; q[1] = rb->y; q[2] = rb->theta; q[3] = rb->speed; q[4] = rb->omega;
            }
            if (j->tr_ncoll) j->tr_ncoll[row] = rb->ncollide - ncoll0;
            if (!alive && !j->tr_pose && !j->tr_ncoll) break;
        }
        if (!(rb->flags & ORC_F_DONE)) rb->fitness = orc_fitness(rb);   /* fitness() after T updates */
    }
    return NULL;
}

/* Run robots[0..P) for T ticks.  Trace pointers may be NULL; layouts [T][P][...].  n_threads >= 1. */
int orc_run(const orc_cfg* c, orc_robot* robots, const double* genomes, int32_t P, int32_t T, int32_t n_threads,
            double* tr_pose, double* tr_dst, int32_t* tr_hitk, int32_t* tr_hitpx, uint32_t* tr_nsamp, uint32_t* tr_ncoll, double* tr_cinfo) {
    if (c->R > ORC_MAX_RAYS || c->n_layers > ORC_MAX_LAYERS) return -1;
    if (c->ctrl_kind == ORC_CTRL_PID && (c->R < 3 || c->G != 6 || !c->pid_state)) return -1;
    for (int l = 0; l < c->n_layers; ++l) if (c->layers[l] > ORC_MAX_WIDTH) return -1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > P) n_threads = P > 0 ? P : 1;
    orc_job* jobs = (orc_job*)calloc((size_t)n_threads, sizeof(orc_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        orc_job* j = &jobs[i];
        j->c = c; j->robots = robots; j->genomes = genomes; j->P = P; j->T = T;
        j->p0 = (int32_t)((int64_t)P * i / n_threads); j->p1 = (int32_t)((int64_t)P * (i + 1) / n_threads);
        j->tr_pose = tr_pose; j->tr_dst = tr_dst; j->tr_hitk = tr_hitk; j->tr_hitpx = tr_hitpx;
        j->tr_nsamp = tr_nsamp; j->tr_ncoll = tr_ncoll; j->tr_cinfo = tr_cinfo;
        if (n_threads == 1) orc_worker(j); else pthread_create(&th[i], NULL, orc_worker, j);
    }
    if (n_threads > 1) for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(jobs); free(th);
    return 0;
}

int orc_sizeof_cfg(void) { return (int)sizeof(orc_cfg); }
int orc_sizeof_pid(void) { return (int)sizeof(orc_pid); }
int orc_sizeof_robot(void) { return (int)sizeof(orc_robot); }
int orc_is_portmath(void) {
#ifdef ORC_PORTMATH
    return 1;
#else
    return 0;
#endif
}
