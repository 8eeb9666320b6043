"""ctypes binding of the CPU parity oracle (oracle/r2p2_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package (r2p2_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

MAX_RAYS, MAX_LAYERS = 64, 8
CTRL = {"const": 0, "neuro": 1, "linear": 2, "pid": 3}
PID_GCAP = 64
ACT = {"identity": 0, "tanh": 1, "relu": 2, "logistic": 3}
F_COLLIDED, F_FAULT, F_DONE, F_TRIGRANGE = 1, 2, 4, 8


class _Cfg(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("env", C.c_void_p), ("R", C.c_int32),
                ("ray_deg", C.c_double * MAX_RAYS),
                ("vr0", C.c_double), ("vr1", C.c_double), ("radius", C.c_double), ("max_speed", C.c_double),
                ("ctrl_kind", C.c_int32), ("n_layers", C.c_int32), ("layers", C.c_int32 * MAX_LAYERS),
                ("activation", C.c_int32), ("G", C.c_int32), ("evolve", C.c_int32),
                ("delta", C.c_double), ("delta_ctrl", C.c_double), ("time_budget", C.c_double),
                ("noise", C.c_void_p), ("noise_len", C.c_int32), ("goal_on", C.c_int32), ("goal_x", C.c_double), ("goal_y", C.c_double),
                ("pid_max_acc", C.c_double), ("pid_accel", C.c_double), ("pid_state", C.c_void_p)]


class _Robot(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("x", "y", "theta", "speed", "omega", "last_x", "last_y", "dist",
                                          "odom_x", "odom_y", "origin_x", "origin_y", "time", "fitness")] + \
               [("flags", C.c_uint32), ("ncollide", C.c_uint32), ("ticks", C.c_uint32),
                ("nsamp_sonar", C.c_uint64), ("nsamp_coll", C.c_uint64), ("npixel", C.c_uint64),
                ("ndraws", C.c_uint32), ("goal_tick", C.c_uint32)]


ROBOT_DTYPE = np.dtype([(n, "<f8") for n in ("x", "y", "theta", "speed", "omega", "last_x", "last_y", "dist",
                                             "odom_x", "odom_y", "origin_x", "origin_y", "time", "fitness")] +
                       [("flags", "<u4"), ("ncollide", "<u4"), ("ticks", "<u4"), ("_pad", "<u4"),
                        ("nsamp_sonar", "<u8"), ("nsamp_coll", "<u8"), ("npixel", "<u8"), ("ndraws", "<u4"), ("goal_tick", "<u4")])


PID_DTYPE = np.dtype([(n, "<f8") for n in ("acc_ang", "acc_dist", "last_ang", "last_dist", "target_angle")] +
                     [("state", "<i4"), ("backing", "<i4"), ("ngoals", "<i4"), ("pad", "<i4"), ("gx", "<f8", (PID_GCAP,)), ("gy", "<f8", (PID_GCAP,))])


def pid_initial_state(P, goals):
    """Controller state of P fresh Sequential_PID_Controllers with the goal list `goals` ([n][2], shared, or [P][n][2])."""
    g = np.asarray(goals, dtype=np.float64)
    if g.ndim == 2:
        g = np.broadcast_to(g, (P,) + g.shape)
    n = g.shape[1]
    assert n <= PID_GCAP
    ps = np.zeros(P, dtype=PID_DTYPE)
    ps["ngoals"] = n
    ps["gx"][:, :n] = g[:, ::-1, 0]                   # a stack: the last slot in use is goal[0]
    ps["gy"][:, :n] = g[:, ::-1, 1]
    return ps


def build(force=False):
    """Compile both oracle flavours into oracle/_build/ (gcc; a few seconds)."""
    libs = [os.path.join(_BUILD, n) for n in ("libr2p2_oracle.so", "libr2p2_oracle_port.so", "libr2p2_mathcheck.so")]
    if force or not all(os.path.exists(p) for p in libs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return libs


def expand_sonars(sonars):
    """Ray angle list as the reference builds it: list form as given (utils.py:504), int form
    ``range(0, 360, floor(365/n))`` (utils.py:465-466) -- so 16 'sonars' give 17 rays."""
    if isinstance(sonars, (list, tuple)):
        return [float(a) for a in sonars]
    import math
    return [float(i) for i in range(0, 360, math.floor(365 / int(sonars)))]


class Oracle:
    def __init__(self, port=False):
        libs = build()
        self.port = bool(port)
        self.lib = C.CDLL(libs[1] if port else libs[0])
        L = self.lib
        assert L.orc_sizeof_cfg() == C.sizeof(_Cfg), (L.orc_sizeof_cfg(), C.sizeof(_Cfg))
        assert L.orc_sizeof_robot() == C.sizeof(_Robot) == ROBOT_DTYPE.itemsize
        assert L.orc_is_portmath() == int(self.port)
        assert L.orc_sizeof_pid() == PID_DTYPE.itemsize, (L.orc_sizeof_pid(), PID_DTYPE.itemsize)
        L.orc_ray.restype = C.c_int
        L.orc_ray.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                              C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7

    @staticmethod
    def _env(env):
        env = np.ascontiguousarray(env, dtype=np.int32)
        assert env.ndim == 2
        return env

    def ray(self, env, ox, oy, angle_rad, limit):
        """search_edge_in_angle_with_limit -> (status, hx, hy, k, nsamp); status 1 hit, 0 miss, -1 fault."""
        env = self._env(env)
        hx, hy, k, ns = C.c_double(), C.c_double(), C.c_int32(), C.c_uint32()
        st = self.lib.orc_ray(env.ctypes.data, env.shape[0], env.shape[1], float(ox), float(oy), float(angle_rad),
                              float(limit), C.byref(hx), C.byref(hy), C.byref(k), C.byref(ns))
        return st, hx.value, hy.value, k.value, ns.value

    def run(self, env, *, rays, vr, radius, max_speed, ctrl, genomes, x0, y0, theta0, T, delta=0.1,
            layers=None, activation="tanh", evolve=False, delta_ctrl=None, time_budget=20000.0,
            noise=None, trace=False, threads=1, goal=None, state=None, pid_goals=None, pid_state=None, pid_max_acc=5.0, pid_accel=0.0):
        """Run P robots for T ticks.  Returns dict(state=structured array [P], plus traces [T,P,...]).
        state: continue from this structured array (a copy is advanced) instead of starting at x0 / y0 / theta0."""
        env = self._env(env)
        genomes = np.ascontiguousarray(genomes, dtype=np.float64)
        P, G = genomes.shape
        rays = list(rays)
        cfg = _Cfg()
        cfg.W, cfg.H, cfg.env = env.shape[0], env.shape[1], env.ctypes.data
        cfg.R = len(rays)
        for i, a in enumerate(rays):
            cfg.ray_deg[i] = a
        cfg.vr0, cfg.vr1, cfg.radius, cfg.max_speed = float(vr[0]), float(vr[1]), float(radius), float(max_speed)
        cfg.ctrl_kind = CTRL[ctrl]
        layers = list(layers or [])
        cfg.n_layers = len(layers)
        for i, n in enumerate(layers):
            cfg.layers[i] = n
        cfg.activation = ACT[activation]
        cfg.G, cfg.evolve = G, int(bool(evolve))
        cfg.delta = float(delta)
        cfg.delta_ctrl = float(delta if delta_ctrl is None else delta_ctrl)
        cfg.time_budget = float(time_budget)
        if goal is not None:
            cfg.goal_on, cfg.goal_x, cfg.goal_y = 1, float(goal[0]), float(goal[1])
        ps = None
        if ctrl == "pid":                                   # genomes = [P][6] PID gains (ap, ai, ad, lp, li, ld)
            ps = np.array(pid_state, dtype=PID_DTYPE, copy=True) if pid_state is not None else pid_initial_state(P, pid_goals)
            cfg.pid_max_acc, cfg.pid_accel, cfg.pid_state = float(pid_max_acc), float(pid_accel), ps.ctypes.data
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            assert noise.ndim == 2 and noise.shape[0] == P      # per-robot stream of N(0, 0.5) draws, consumed sequentially
            cfg.noise = noise.ctypes.data
            cfg.noise_len = noise.shape[1]
        if state is not None:
            st = np.array(state, dtype=ROBOT_DTYPE, copy=True)
            assert st.shape == (P,)
            x0 = y0 = theta0 = 0.0
        else:
            st = None
        st_given = st is not None
        if not st_given:
            st = np.zeros(P, dtype=ROBOT_DTYPE)
        x0 = np.broadcast_to(np.asarray(x0, dtype=np.float64), (P,))
        y0 = np.broadcast_to(np.asarray(y0, dtype=np.float64), (P,))
        th0 = np.broadcast_to(np.asarray(theta0, dtype=np.float64), (P,))
        if not st_given:
            for f, v in (("x", x0), ("last_x", x0), ("odom_x", x0), ("origin_x", x0),
                         ("y", y0), ("last_y", y0), ("odom_y", y0), ("origin_y", y0), ("theta", th0)):
                st[f] = v
            st["time"] = cfg.time_budget
            st["goal_tick"] = 0xFFFFFFFF
        out = {}
        ptrs = [None] * 7
        if trace:
            R = len(rays)
            out["pose"] = np.zeros((T, P, 5)); out["dst"] = np.zeros((T, P, R))
            out["hitk"] = np.full((T, P, R), -1, dtype=np.int32); out["hitpx"] = np.full((T, P, R, 2), -1, dtype=np.int32)
            out["nsamp"] = np.zeros((T, P, R), dtype=np.uint32); out["ncoll"] = np.zeros((T, P), dtype=np.uint32)
            out["cinfo"] = np.zeros((T, P, 6))
            ptrs = [out[k].ctypes.data for k in ("pose", "dst", "hitk", "hitpx", "nsamp", "ncoll", "cinfo")]
        rc = self.lib.orc_run(C.byref(cfg), st.ctypes.data, genomes.ctypes.data, P, int(T), int(threads), *ptrs)
        if rc != 0:
            raise ValueError("oracle: unsupported configuration")
        out["state"] = st
        out["fitness"] = st["fitness"].copy()
        if ps is not None:
            out["pid"] = ps
        return out


def _resume(self, env, state, **kw):
    """Advance an existing state array by T more ticks (see ``run``)."""
    kw.setdefault("x0", 0.0); kw.setdefault("y0", 0.0); kw.setdefault("theta0", 0.0)
    return self.run(env, state=state, **kw)


Oracle.resume = _resume


def cost_grid(env, divider):
    """path_planning.generate_grid (r2p2/path_planning.py:71-110) + Node.value (r2p2/node.py:64), restated: returns
    (values, costs) int32 [div_y, div_x].  The reference writes each cell's value back into its copy of the map; the
    write-back never reaches a pixel a later cell reads (cells are visited in increasing x, y and read at or beyond their
    own int(j*cw), int(i*ch)), so the original map decides every cell."""
    env = np.asarray(env)
    W, H = env.shape
    dx, dy = (divider, divider) if np.isscalar(divider) else divider
    cw, ch = W / dx, H / dy
    values = np.zeros((dy, dx), dtype=np.int32)
    for i in range(dy):
        for j in range(dx):
            value = int(env[int(j * cw + cw / 2), int(i * ch + ch / 2)])
            x0, y0 = int(j * cw), int(i * ch)
            if (env[x0:x0 + int(cw), y0:y0 + int(ch)] == 0).any():     # int(j*cw + l) == int(j*cw) + l for integer l
                value = 0
            values[i, j] = value
    costs = (9 - values / 255 * 8).astype(np.int32)                    # int(): truncation, values are >= 1
    return values, costs
