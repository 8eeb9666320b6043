"""BatchSim: Python face of the C ABI -- one population of robots on one GPU.

Host code only (numpy in, numpy out).  All arithmetic happens in the CUDA library.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import R2P2Error, check

CTRL = {"const": 0, "neuro": 1, "linear": 2, "pid": 3}
ACT = {"identity": 0, "tanh": 1, "relu": 2, "logistic": 3}
F_COLLIDED, F_FAULT, F_DONE, F_TRIGRANGE = 1, 2, 4, 8


def expand_sonars(sonars):
    """Ray angles as the reference builds them: a list is taken as given
    (utils.search_in_all_directions_with_angles_and_offset_with_limit, r2p2/utils.py:497-514); an int n
    means ``range(0, 360, floor(365/n))`` (r2p2/utils.py:465-466) -- 16 'sonars' are 17 rays."""
    if isinstance(sonars, (list, tuple, np.ndarray)):
        return [float(a) for a in sonars]
    return [float(i) for i in range(0, 360, math.floor(365 / int(sonars)))]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class BatchSim:
    """A population of identical robots on one stage, evaluated by one kernel launch per ``run``."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.r2p2_create(int(device), C.byref(h))
        if rc != 0:
            msg = self._L.r2p2_last_error(None)
            raise R2P2Error("r2p2_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self._h = h
        self.device = int(device)
        self.R = 0
        self._trace_T = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.r2p2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        check(self._h, rc)

    # The population size lives in the library (r2p2_ga_evaluate resizes it to the evaluated shard): every host buffer
    # below is sized from what the library reports, never from a value cached here.
    def _size(self):
        P, G = C.c_int32(), C.c_int32()
        self._ck(self._L.r2p2_population_size(self._h, C.byref(P), C.byref(G)))
        return P.value, G.value

    @property
    def P(self):
        return self._size()[0]

    @property
    def G(self):
        return self._size()[1]

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (integer cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self._L.r2p2_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def fp64_peak(self):
        """Measured DFMA thread-instructions per second on this device."""
        v = C.c_double()
        self._ck(self._L.r2p2_fp64_peak(self._h, C.byref(v)))
        return v.value

    # ---- configuration -----------------------------------------------------------------------
    def set_stage(self, env):
        """env[x, y] grey levels as the reference holds them after its transpose (shape (W, H))."""
        env = np.asarray(env)
        if env.ndim != 2:
            raise ValueError("stage must be 2-D (W, H)")
        if env.dtype == np.uint8:
            e = np.ascontiguousarray(env)
            self._ck(self._L.r2p2_set_stage(self._h, e.ctypes.data, e.shape[0], e.shape[1]))
        else:
            e = np.ascontiguousarray(env, dtype=np.int32)
            self._ck(self._L.r2p2_set_stage_i32(self._h, e.ctypes.data, e.shape[0], e.shape[1]))
        self.W, self.H = e.shape

    def set_robot(self, *, sonar_range, radius, max_speed, sonars):
        rays = _f64(expand_sonars(sonars))
        self.R = len(rays)
        self._ck(self._L.r2p2_set_robot(self._h, float(sonar_range[0]), float(sonar_range[1]), float(radius),
                                        float(max_speed), rays.ctypes.data, self.R))

    def set_controller(self, kind, hidden_layer=None, activation="tanh"):
        if kind == "neuro":
            layers = np.asarray([self.R] + list(hidden_layer) + [2], dtype=np.int32)
            self._ck(self._L.r2p2_set_controller(self._h, CTRL[kind], layers.ctypes.data, len(layers), ACT[activation]))
        else:
            self._ck(self._L.r2p2_set_controller(self._h, CTRL[kind], None, 0, 0))

    def set_pid(self, goals, max_acceleration=5.0, acceleration=0.0):
        """Goal list(s) of the device Sequential_PID_Controller (controller kind "pid", genome = [ap, ai, ad, lp, li, ld]):
        [n, 2] shared by every robot, or [P, n, 2].  Loaded by the next ``reset``; the reference's register_robot starts
        the robot at goals[0] -- pass that as the start pose."""
        g = _f64(goals)
        if g.ndim not in (2, 3) or g.shape[-1] != 2:
            raise ValueError("goals must be [n, 2] or [P, n, 2]")
        if g.ndim == 3 and g.shape[0] != self.P:
            raise ValueError("per-robot goals: %d lists for a population of %d" % (g.shape[0], self.P))
        self._ck(self._L.r2p2_set_pid(self._h, g.ctypes.data, g.shape[-2], int(g.ndim == 3), float(max_acceleration), float(acceleration)))

    def pid_state(self):
        """Controller state per robot: goals left, state (0 advancing, 1 avoiding, 2 advancing after avoiding), ticks of
        backing off left, target angle, current goal (NaN when the list is empty)."""
        P = self.P
        out = {k: np.empty(P, dtype=np.int32) for k in ("goals_left", "state", "backing")}
        out.update({k: np.empty(P) for k in ("target_angle", "goal_x", "goal_y")})
        self._ck(self._L.r2p2_read_pid(self._h, *[out[k].ctypes.data for k in ("goals_left", "state", "backing", "target_angle", "goal_x", "goal_y")]))
        return out

    def set_noise(self, mode="off", seed=0, stream=None):
        """Sensor noise (Robot.has_noise): "off", "replay" (stream[P, len] of N(0, 0.5) draws, bit-exact with the
        reference's np.random.normal sequence) or "device" (Philox generator keyed by seed)."""
        m = {"off": 0, "replay": 1, "device": 2}[mode]
        if m == 1:
            s = _f64(stream)
            if s.ndim != 2:
                raise ValueError("noise stream must be [P, len]")
            self._ck(self._L.r2p2_set_noise(self._h, 1, int(seed), s.ctypes.data, s.shape[0], s.shape[1]))
        else:
            self._ck(self._L.r2p2_set_noise(self._h, m, int(seed), None, 0, 0))

    def noise_draws(self):
        n = np.empty(self.P, dtype=np.uint32)
        self._ck(self._L.r2p2_read_noise_draws(self._h, n.ctypes.data))
        return n

    def set_mapping(self, lanes_per_robot=0, stage_in_smem=-1):
        self._ck(self._L.r2p2_set_mapping(self._h, int(lanes_per_robot), int(stage_in_smem)))

    # ---- population ----------------------------------------------------------------------------
    def upload_genomes(self, genomes):
        g = _f64(genomes)
        if g.ndim != 2:
            raise ValueError("genomes must be [P, G]")
        self._keep = [g]
        self._ck(self._L.r2p2_upload_genomes(self._h, g.ctypes.data, g.shape[0], g.shape[1]))

    def upload_genomes_ptr(self, host_ptr, P, G):
        """Host pointer variant (e.g. a pinned torch tensor's data_ptr()); the buffer must stay alive."""
        self._ck(self._L.r2p2_upload_genomes(self._h, C.c_void_p(int(host_ptr)), int(P), int(G)))

    def set_genomes_device(self, dev_ptr, P, G):
        self._ck(self._L.r2p2_set_genomes_device(self._h, C.c_void_p(int(dev_ptr)), int(P), int(G)))

    def reset(self, x0, y0, theta0=0.0, time_budget=20000.0):
        x0, y0, t0 = np.atleast_1d(_f64(x0)), np.atleast_1d(_f64(y0)), np.atleast_1d(_f64(theta0))
        n = max(len(x0), len(y0), len(t0))
        if n > 1:
            x0, y0, t0 = (_f64(np.broadcast_to(a, (n,))) for a in (x0, y0, t0))
        self._ck(self._L.r2p2_reset(self._h, x0.ctypes.data, y0.ctypes.data, t0.ctypes.data, n, float(time_budget)))

    def write_pose(self, first, x=None, y=None, theta=None, last_x=None, last_y=None):
        """Overwrite pose attributes of robots [first, first + n) as host code does in the reference (set_position,
        set_last_position, ``robot.orientation = ...``); odometry, fitness and flags are left alone."""
        arrs = [None if a is None else np.atleast_1d(_f64(a)) for a in (x, y, theta, last_x, last_y)]
        n = max(len(a) for a in arrs if a is not None)
        assert all(a is None or len(a) == n for a in arrs)
        self._ck(self._L.r2p2_write_pose(self._h, int(first), n, *[None if a is None else a.ctypes.data for a in arrs]))

    def set_robot_index_offset(self, first):
        """Index of robot 0 in the whole sharded population (keys the device noise generator)."""
        self._ck(self._L.r2p2_set_robot_index_offset(self._h, int(first)))

    # ---- execution -----------------------------------------------------------------------------
    def run(self, T, delta=0.1, delta_ctrl=None, evolve=False):
        self._ck(self._L.r2p2_run(self._h, int(T), float(delta), float(delta if delta_ctrl is None else delta_ctrl), int(bool(evolve))))

    def evaluate_ptr(self, genomes_host_ptr, P, G, x0, y0, theta0, time_budget, T, delta, delta_ctrl, evolve, fitness_host_ptr):
        """upload + reset + run + read fitness as one call with the genome upload pipelined against the kernel
        (r2p2_evaluate); raw host pointers, page-locked for the overlap."""
        x0, y0, t0 = np.atleast_1d(_f64(x0)), np.atleast_1d(_f64(y0)), np.atleast_1d(_f64(theta0))
        n = max(len(x0), len(y0), len(t0))
        if n > 1:
            x0, y0, t0 = (_f64(np.broadcast_to(a, (n,))) for a in (x0, y0, t0))
        self._ck(self._L.r2p2_evaluate(self._h, C.c_void_p(int(genomes_host_ptr)), int(P), int(G), x0.ctypes.data, y0.ctypes.data, t0.ctypes.data,
                                       n, float(time_budget), int(T), float(delta), float(delta_ctrl), int(bool(evolve)),
                                       C.c_void_p(int(fitness_host_ptr))))

    def sync(self):
        self._ck(self._L.r2p2_sync(self._h))

    def last_run_ms(self):
        ms = C.c_float()
        self._ck(self._L.r2p2_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    # ---- results ---------------------------------------------------------------------------------
    def fitness(self, out=None):
        P = self.P
        f = np.empty(P) if out is None else out
        if f.size < P or f.dtype != np.float64 or not f.flags["C_CONTIGUOUS"]:
            raise ValueError("fitness buffer must be a contiguous float64 array of at least %d elements" % P)
        self._ck(self._L.r2p2_read_fitness(self._h, f.ctypes.data))
        return f

    def read_fitness_ptr(self, host_ptr):
        self._ck(self._L.r2p2_read_fitness(self._h, C.c_void_p(int(host_ptr))))

    def state(self):
        P = self.P
        d = {k: np.empty(P) for k in ("x", "y", "theta", "speed", "omega")}
        u = {k: np.empty(P, dtype=np.uint32) for k in ("flags", "ncollide", "ticks")}
        self._ck(self._L.r2p2_read_state(self._h, *[d[k].ctypes.data for k in ("x", "y", "theta", "speed", "omega")],
                                         *[u[k].ctypes.data for k in ("flags", "ncollide", "ticks")]))
        d.update(u)
        return d

    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._L.r2p2_read_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"sonar_samples": a.value, "collision_samples": b.value, "robot_steps": c.value}

    def fitness_device_ptr(self):
        p = C.c_void_p()
        self._ck(self._L.r2p2_fitness_device(self._h, C.byref(p)))
        return p.value

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self._L.r2p2_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- trace -----------------------------------------------------------------------------------
    def set_trace(self, enable, T_max=1):
        self._ck(self._L.r2p2_set_trace(self._h, int(bool(enable)), int(T_max)))
        self._trace_T = int(T_max) if enable else 0
        self._trace_P = self.P if enable else 0                # the library copies rows of THIS population size

    def trace(self, T):
        P, R = self._trace_P, self.R
        out = {"pose": np.empty((T, P, 5)), "dst": np.empty((T, P, R)), "hitk": np.empty((T, P, R), dtype=np.int32),
               "hitpx": np.empty((T, P, R, 2), dtype=np.int32), "nsamp": np.empty((T, P, R), dtype=np.uint32),
               "ncoll": np.empty((T, P), dtype=np.uint32)}
        self._ck(self._L.r2p2_read_trace(self._h, int(T), *[out[k].ctypes.data for k in ("pose", "dst", "hitk", "hitpx", "nsamp", "ncoll")]))
        out["cinfo"] = np.empty((T, P, 6))        # collide() calls per tick (code, wall spot, border coordinate) + the controller's answer: r2p2_b200.logs
        self._ck(self._L.r2p2_read_trace_collisions(self._h, int(T), out["cinfo"].ctypes.data))
        return out

    # ---- checkpoint / resume ---------------------------------------------------------------------------
    def save_state(self):
        """Opaque snapshot (bytes) of everything the running episodes carry between launches."""
        import ctypes as C
        n = C.c_uint64()
        self._ck(self._L.r2p2_state_bytes(self._h, C.byref(n)))
        blob = np.empty(n.value, dtype=np.uint8)
        self._ck(self._L.r2p2_save_state(self._h, blob.ctypes.data))
        return blob.tobytes()

    def load_state(self, blob):
        b = np.frombuffer(blob, dtype=np.uint8)
        import ctypes as C
        n = C.c_uint64()
        self._ck(self._L.r2p2_state_bytes(self._h, C.byref(n)))
        if b.size != n.value:
            raise ValueError("snapshot of %d bytes does not fit a population of %d robots (%d bytes)" % (b.size, self.P, n.value))
        if b[:4].tobytes() != b"R2PS":
            raise ValueError("not a state snapshot (bad magic)")
        self._ck(self._L.r2p2_load_state(self._h, b.ctypes.data))

    # ---- path planner's cost grid --------------------------------------------------------------------
    def cost_grid(self, divider):
        """path_planning.generate_grid + Node.value for `divider` (int or [div_x, div_y]): (values, costs), both
        int32 [div_y, div_x] (grid[i][j] of the reference)."""
        dx, dy = (divider, divider) if np.isscalar(divider) else divider
        v = np.empty((int(dy), int(dx)), dtype=np.int32); c = np.empty_like(v)
        self._ck(self._L.r2p2_cost_grid(self._h, int(dx), int(dy), v.ctypes.data, c.ctypes.data))
        return v, c

    # ---- goal flagging ---------------------------------------------------------------------------------
    def set_goal(self, goal=None):
        """Per-run goal point (x, y) for the is_at_goal predicate (pid_controller.py:207-213); None switches it off."""
        if goal is None:
            self._ck(self._L.r2p2_set_goal(self._h, 0, 0.0, 0.0))
        else:
            self._ck(self._L.r2p2_set_goal(self._h, 1, float(goal[0]), float(goal[1])))

    def goal_ticks(self):
        """[P] ticks completed when each robot first satisfied the goal predicate (0xffffffff: never)."""
        g = np.empty(self.P, dtype=np.uint32)
        self._ck(self._L.r2p2_read_goal(self._h, g.ctypes.data))
        return g

    # ---- controllers that run on the host: the tick in two halves ----------------------------------
    def sense(self, hit_pixels=False):
        """Robot.check_sensors for every robot at its current pose: dst [P, R] (and the hit pixels [P, R, 2] or -1)."""
        dst = np.empty((self.P, self.R))
        px = np.empty((self.P, self.R, 2), dtype=np.int32) if hit_pixels else None
        self._ck(self._L.r2p2_sense(self._h, dst.ctypes.data, px.ctypes.data if hit_pixels else None))
        return (dst, px) if hit_pixels else dst

    def move(self, commands, delta=0.1):
        """The rest of the tick with the controllers' answers: commands [P, 2] = (speed, angular_velocity)."""
        c = _f64(commands).reshape(self.P, 2)
        self._ck(self._L.r2p2_move(self._h, c.ctypes.data, float(delta)))

    # ---- raw rays ----------------------------------------------------------------------------------
    def cast_rays(self, ox, oy, angle_rad, limit, plain=False):
        ox, oy, ang, lim = (_f64(a) for a in (ox, oy, angle_rad, limit))
        n = len(ox)
        st = np.empty(n, dtype=np.int32); k = np.empty(n, dtype=np.int32); ns = np.empty(n, dtype=np.uint32)
        hx = np.empty(n); hy = np.empty(n)
        fn = self._L.r2p2_cast_rays_plain if plain else self._L.r2p2_cast_rays
        self._ck(fn(self._h, n, ox.ctypes.data, oy.ctypes.data, ang.ctypes.data, lim.ctypes.data,
                st.ctypes.data, hx.ctypes.data, hy.ctypes.data, k.ctypes.data, ns.ctypes.data))
        return {"status": st, "hx": hx, "hy": hy, "k": k, "nsamp": ns}
