#!/usr/bin/env python3
"""Record golden vectors from the UNMODIFIED r2p2 reference.  Build-container only.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference ships no tests or fixtures, so parity is pinned by running
the reference itself (tests/golden/ref_harness.py, recipe of SURVEY.md section 8c) and committing
what it returns.  The GPU box has no /root/reference; tests there read only these files.

Three fixture files (float64 values are stored as-is, i.e. bit-exact):
  rays.npz      raw `utils.search_edge_in_angle_with_limit` calls (reference utils.py:420-432):
                integer and float origins, several limits, three stages; hit point, distance and
                the number of pixels the loop tested (counted through an ndarray subclass).
  ticks.npz     single `Robot.update` calls from random states with a scripted command (a
                Controller subclass returning a fixed (speed, angular_velocity)): everything the
                tick produces except the MLP -- sonar distances, raw ray hits, new pose, collide
                calls, exceptions.
  episodes.npz  free-running episodes with the reference's own Neuro_controller (sklearn MLP),
                Inspyred_controller (linear) and a constant command (KAT2/KAT3/KAT4 of SURVEY 8c).
ulp-level digits are specific to this image (glibc 2.39, numpy 2.3.5 / OpenBLAS 0.3.30, AVX-512
Xeon); the environment is recorded in each file's `meta`.
"""
import contextlib
import io
import json
import math
import os
import platform
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STAGES = ["track_2.png", "stage.png", "track.png"]


class CountingEnv(np.ndarray):
    """env whose .item() calls are counted (= pixels tested by the ray loops)."""
    count = 0

    def item(self, *a):
        CountingEnv.count += 1
        return super().item(*a)


def counting(env):
    return np.ascontiguousarray(env).view(CountingEnv)


def meta():
    import sklearn
    return json.dumps({"python": platform.python_version(), "numpy": np.__version__, "sklearn": sklearn.__version__,
                       "libc": platform.libc_ver()[1], "machine": platform.machine(),
                       "cpu_flags": [f for f in open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0].split()
                                     if f in ("fma", "avx2", "avx512f")]})


def free_points(env, rng, n, clearance):
    """Random float positions whose whole crash-radius disc is free (valid last_pos)."""
    W, H = env.shape
    wall = (np.asarray(env) == 0)
    pts = []
    while len(pts) < n:
        x = rng.uniform(clearance + 1, W - clearance - 2)
        y = rng.uniform(clearance + 1, H - clearance - 2)
        x0, x1, y0, y1 = int(x - clearance - 1), int(x + clearance + 2), int(y - clearance - 1), int(y + clearance + 2)
        if not wall[x0:x1, y0:y1].any():
            pts.append((x, y))
    return pts


def near_wall_points(env, rng, n, radius):
    """Free positions 5..9 px from some wall pixel (so that moves collide)."""
    W, H = env.shape
    wall = (np.asarray(env) == 0)
    pts = []
    while len(pts) < n:
        x = rng.uniform(radius + 2, W - radius - 3)
        y = rng.uniform(radius + 2, H - radius - 3)
        r_in, r_out = int(radius + 1), int(radius + 5)
        xi, yi = int(x), int(y)
        if wall[xi - r_in:xi + r_in + 1, yi - r_in:yi + r_in + 1].any():
            continue
        if wall[max(0, xi - r_out):xi + r_out + 1, max(0, yi - r_out):yi + r_out + 1].any():
            pts.append((x, y))
    return pts


def gen_rays(E):
    u = E.u
    rng = np.random.RandomState(20261019)
    rec = {k: [] for k in ("stage", "ox", "oy", "angle", "limit", "hit", "hx", "hy", "dst", "nsamp")}
    for si, stage in enumerate(STAGES):
        env = counting(E.load_env(stage))
        W, H = env.shape
        cases = []
        # KAT1 of SURVEY 8c first
        if stage == "track_2.png":
            cases += [((230, 250), math.radians(a), 30) for a in (0, 45, 70, 290, 315)]
            cases += [((193, 249), math.radians(a), 55) for a in range(0, 360, 45)]
        if stage == "stage.png":
            cases += [((260, 200), math.radians(a), 30) for a in (0, 45, 70, 290, 315)]
        for _ in range(1400):   # integer origins, integer limits (the sonar case, knife edge H5)
            o = (int(rng.randint(1, W - 1)), int(rng.randint(1, H - 1)))
            cases.append((o, math.radians(rng.uniform(-720, 720)), int(rng.choice([30, 45, 55]))))
        for _ in range(300):    # exact multiples of 45 degrees plus heading offsets that are integers
            o = (int(rng.randint(1, W - 1)), int(rng.randint(1, H - 1)))
            cases.append((o, math.radians(float(rng.randint(-8, 16) * 45)), 30))
        for _ in range(800):    # float origins, float limits (the collision ray)
            o = (float(rng.uniform(1, W - 2)), float(rng.uniform(1, H - 2)))
            cases.append((o, math.radians(rng.uniform(0, 360)), float(rng.uniform(5.0, 12.0))))
        for _ in range(200):    # rays that leave the map / start near the border
            o = (float(rng.uniform(-3, W + 3)), float(rng.uniform(-3, H + 3)))
            cases.append((o, math.radians(rng.uniform(0, 360)), float(rng.uniform(3.0, 60.0))))
        for (o, ang, lim) in cases:
            CountingEnv.count = 0
            try:
                res = u.search_edge_in_angle_with_limit(o, ang, env, lim)
                hit = int(res != (-1, -1))
            except IndexError:
                res, hit = (-1, -1), -1
            ns = CountingEnv.count
            d = float(np.linalg.norm((res[0] - o[0], res[1] - o[1]))) if hit == 1 else math.inf
            for k, v in zip(rec, (si, o[0], o[1], ang, lim, hit, res[0], res[1], d, ns)):
                rec[k].append(v)
    np.savez_compressed(os.path.join(OUT, "rays.npz"), meta=meta(), stages=np.array(STAGES),
                        **{k: np.asarray(v, dtype=(np.int32 if k in ("stage", "hit", "nsamp") else np.float64))
                           for k, v in rec.items()})
    print("rays:", len(rec["hit"]), "hits", sum(1 for h in rec["hit"] if h == 1), "faults", sum(1 for h in rec["hit"] if h < 0))


def make_scripted(E):
    class Scripted(E.ctrl.Controller):
        """Harness controller: returns the command it was given; counts collide() calls."""
        def __init__(self):
            super().__init__("SCRIPTED", None)
            self.cmd = (0.0, 0.0)
            self.ncollide = 0

        def control(self, dst):
            return self.cmd

        def on_collision(self, pos):
            self.ncollide += 1
    return Scripted


def gen_ticks(E):
    u = E.u
    Scripted = make_scripted(E)
    rng = np.random.RandomState(77)
    raw = []
    orig = u.search_edge_in_angle_with_limit

    def spy(origin, angle, image, limit):
        res = orig(origin, angle, image, limit)
        raw.append(res)
        return res
    u.search_edge_in_angle_with_limit = spy
    keys = ("stage", "robot", "delta", "x", "y", "theta", "v", "w", "nx", "ny", "ntheta", "nspeed", "ncollide", "exc",
            "nlast_x", "nlast_y", "nsamp")
    rec = {k: [] for k in keys}
    dsts, rays_hit = [], []
    robots = ["robot-neuro.json", "robot_omni.json", "robot-track.json"]
    try:
        for si, stage in enumerate(STAGES):
            env = counting(E.load_env(stage))
            for ri, rj in enumerate(robots):
                ctrl = Scripted()
                with contextlib.redirect_stdout(io.StringIO()):
                    r = E.make_robot(rj, ctrl)
                n = 700 if ri == 0 else 350
                pts = free_points(env, rng, n // 2, r.radius + 1) + near_wall_points(env, rng, n - n // 2, r.radius)
                for i, (x, y) in enumerate(pts):
                    theta = float(rng.uniform(-400, 400)) if i % 5 else float(rng.randint(-4, 9) * 45)
                    v = float(rng.uniform(-60, 60)) if i % 7 else 0.0
                    w = float(rng.uniform(-90, 90))
                    delta = 0.1 if i % 4 else float(rng.choice([0.033, 0.5, 1.0]))
                    r.x, r.y, r.orientation, r.last_pos = x, y, theta, (x, y)
                    ctrl.cmd, ctrl.ncollide = (v, w), 0
                    raw.clear()
                    CountingEnv.count = 0
                    exc = 0
                    try:
                        r.update(env, delta, False)
                    except IndexError:
                        exc = 1
                        import threading
                        r.lock = threading.Lock()
                    R = len(ctrl.dst)
                    d = np.full(8, np.nan)
                    d[:R] = ctrl.dst
                    dsts.append(d)
                    h = np.full((9, 2), np.nan)
                    for j, p in enumerate(raw[:9]):
                        h[j] = p
                    rays_hit.append(h)
                    for k, val in zip(keys, (si, ri, delta, x, y, theta, v, w, r.x, r.y, r.orientation, r.speed,
                                             ctrl.ncollide, exc, r.last_pos[0], r.last_pos[1], CountingEnv.count)):
                        rec[k].append(val)
    finally:
        u.search_edge_in_angle_with_limit = orig
    np.savez_compressed(os.path.join(OUT, "ticks.npz"), meta=meta(), stages=np.array(STAGES), robots=np.array(robots),
                        dst=np.asarray(dsts), raw_hits=np.asarray(rays_hit),
                        **{k: np.asarray(v, dtype=(np.int32 if k in ("stage", "robot", "ncollide", "exc", "nsamp") else np.float64))
                           for k, v in rec.items()})
    print("ticks:", len(rec["x"]), "with collide", sum(1 for c in rec["ncollide"] if c), "exceptions", sum(rec["exc"]))


def gen_episodes(E):
    Scripted = make_scripted(E)
    out = {}
    specs = []

    def run(name, stage, robot_json, ctrl, T, genome, kind, layers=(), delta=0.1):
        env = counting(E.load_env(stage))
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
        ncol = [0]
        orig = ctrl.on_collision

        def oc(pos):
            ncol[0] += 1
            orig(pos)
        ctrl.on_collision = oc
        R = len(r.sensors) if isinstance(r.sensors, list) else len(range(0, 360, math.floor(365 / r.sensors)))
        pose = np.zeros((T, 5)); dst = np.zeros((T, R)); ncoll = np.zeros(T, dtype=np.int32); fit = np.zeros(T); nsamp = np.zeros(T, dtype=np.int32)
        for t in range(T):
            CountingEnv.count = 0
            r.update(env, delta, False)
            pose[t] = (r.x, r.y, r.orientation, r.speed, r.angular_velocity)
            dst[t] = ctrl.dst
            ncoll[t] = ncol[0]
            nsamp[t] = CountingEnv.count
            fit[t] = ctrl.fitness() if hasattr(ctrl, "fitness") else 0.0
        out[name + "/pose"] = pose; out[name + "/dst"] = dst; out[name + "/ncoll"] = ncoll
        out[name + "/fitness"] = fit; out[name + "/genome"] = np.asarray(genome, dtype=np.float64); out[name + "/nsamp"] = nsamp
        specs.append({"name": name, "stage": stage, "robot": robot_json, "T": T, "kind": kind, "layers": list(layers), "delta": delta})
        print("episode", name, "T", T, "collides", ncol[0], "fitness", repr(float(fit[-1])))

    # KAT2: neuro 5-10-7-4-2 on track_2
    w = np.random.RandomState(1).uniform(-1, 1, 156)
    run("kat2_neuro", "track_2.png", "robot-neuro.json", E.neuro_controller(w, [10, 7, 4]), 1000, w, "neuro", (5, 10, 7, 4, 2))
    for seed in (2, 3, 4):
        w = np.random.RandomState(seed).uniform(-1, 1, 156)
        run("neuro_stage_s%d" % seed, "stage.png", "robot-neuro.json", E.neuro_controller(w, [10, 7, 4]), 400, w, "neuro", (5, 10, 7, 4, 2))
    w = np.random.RandomState(5).uniform(-1, 1, 7)
    run("neuro_simple_track", "track.png", "robot-track.json", E.neuro_controller(w, [1]), 400, w, "neuro", (5, 1, 2))
    # KAT3: linear controller, 8 int-form sonars
    c = np.random.RandomState(7).uniform(-1, 1, 18) * 0.05
    run("kat3_linear", "track_2.png", "robot_omni.json", E.linear_controller(c), 300, c, "linear")
    for seed in (8, 9):
        c = np.random.RandomState(seed).uniform(-1, 1, 18) * 0.03
        run("linear_s%d" % seed, "track_2.png", "robot_omni.json", E.linear_controller(c), 300, c, "linear")
    c = np.random.RandomState(10).uniform(-1, 1, 12) * 0.08     # BASELINE C2: robot-neuro ring + linear, G = 12
    run("linear_neuro_ring", "track_2.png", "robot-neuro.json", E.linear_controller(c), 300, c, "linear")
    # KAT4: constant command into a wall
    sc = Scripted(); sc.cmd = (30.0, 2.0)
    run("kat4_const", "track_2.png", "robot-neuro.json", sc, 100, [30.0, 2.0], "const")
    sc = Scripted(); sc.cmd = (-12.5, -7.0)
    run("const_stage", "stage.png", "robot.json", sc, 300, [-12.5, -7.0], "const")
    np.savez_compressed(os.path.join(OUT, "episodes.npz"), meta=meta(), specs=json.dumps(specs), **out)


def gen_noisy_episodes(E):
    """has_noise = True (the reference's default, robot.py:96,308-310): np.random.normal(0, 0.5) is drawn once per
    unclamped ray, in ray order, from the global legacy MT19937 stream.  Seeded runs; the same stream is reproduced
    for the oracle / device with np.random.seed(s); np.random.normal(0, 0.5, size=n)."""
    Scripted = make_scripted(E)
    out, specs = {}, []

    def run(name, stage, robot_json, ctrl, T, genome, kind, seed, layers=()):
        env = counting(E.load_env(stage))
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
        r.has_noise = True
        ncol = [0]
        orig = ctrl.on_collision

        def oc(pos):
            ncol[0] += 1
            orig(pos)
        ctrl.on_collision = oc
        R = len(r.sensors) if isinstance(r.sensors, list) else len(range(0, 360, math.floor(365 / r.sensors)))
        pose = np.zeros((T, 5)); dst = np.zeros((T, R)); ncoll = np.zeros(T, dtype=np.int32)
        np.random.seed(seed)
        for t in range(T):
            r.update(env, 0.1, False)
            pose[t] = (r.x, r.y, r.orientation, r.speed, r.angular_velocity)
            dst[t] = ctrl.dst
            ncoll[t] = ncol[0]
        nxt = np.random.normal(0, 0.5)                     # the draw after the last one consumed: fixes the draw count
        np.random.seed(seed)
        stream = np.random.normal(0, 0.5, size=T * R + 1)
        ndraws = int(np.argmax(stream == nxt))
        assert stream[ndraws] == nxt
        out[name + "/pose"] = pose; out[name + "/dst"] = dst; out[name + "/ncoll"] = ncoll
        out[name + "/genome"] = np.asarray(genome, dtype=np.float64); out[name + "/ndraws"] = np.int64(ndraws)
        out[name + "/fitness"] = np.float64(ctrl.fitness() if hasattr(ctrl, "fitness") else 0.0)
        specs.append({"name": name, "stage": stage, "robot": robot_json, "T": T, "kind": kind, "layers": list(layers), "seed": seed})
        print("noisy episode", name, "T", T, "collides", ncol[0], "draws", ndraws)

    c = np.random.RandomState(7).uniform(-1, 1, 18) * 0.05
    run("noise_linear_omni", "track_2.png", "robot_omni.json", E.linear_controller(c), 250, c, "linear", 42)
    c = np.random.RandomState(10).uniform(-1, 1, 12) * 0.08
    run("noise_linear_ring", "track_2.png", "robot-neuro.json", E.linear_controller(c), 250, c, "linear", 43)
    sc = Scripted(); sc.cmd = (12.0, 5.0)
    run("noise_const", "stage.png", "robot.json", sc, 200, [12.0, 5.0], "const", 44)
    w = np.random.RandomState(1).uniform(-1, 1, 156)
    run("noise_neuro", "track_2.png", "robot-neuro.json", E.neuro_controller(w, [10, 7, 4]), 300, w, "neuro", 45, (5, 10, 7, 4, 2))
    np.savez_compressed(os.path.join(OUT, "episodes_noise.npz"), meta=meta(), specs=json.dumps(specs), **out)


def gen_act_episodes(E):
    """Neuro episodes with the other activations sklearn's MLP accepts (neurocontroller.py:26,117) and other layer shapes
    -> tests/golden/episodes_act.npz."""
    out, specs = {}, []
    cases = [("relu_10_7_4", "track_2.png", "robot-neuro.json", [10, 7, 4], "relu", 21, 1.0, 250),
             ("logistic_10_7_4", "track_2.png", "robot-neuro.json", [10, 7, 4], "logistic", 22, 1.0, 250),
             ("identity_6", "stage.png", "robot-neuro.json", [6], "identity", 23, 0.2, 250),
             ("tanh_3_3", "track.png", "robot-track.json", [3, 3], "tanh", 24, 2.0, 250),
             ("relu_12", "stage.png", "robot.json", [12], "relu", 25, 0.7, 250),
             ("logistic_2_9", "track_2.png", "robot-neuro.json", [2, 9], "logistic", 26, 3.0, 250)]
    for name, stage, robot_json, hidden, act, seed, scale, T in cases:
        sizes = [5] + hidden + [2]
        G = sum(a * b for a, b in zip(sizes, sizes[1:]))
        w = np.random.RandomState(seed).uniform(-1, 1, G) * scale
        ctrl = E.neuro_controller(w, hidden, activation=act)
        env = E.load_env(stage)
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
        ncol = [0]
        orig = ctrl.on_collision

        def oc(pos, orig=orig):
            ncol[0] += 1
            orig(pos)
        ctrl.on_collision = oc
        pose = np.zeros((T, 5)); dst = np.zeros((T, 5)); ncoll = np.zeros(T, dtype=np.int32)
        for t in range(T):
            r.update(env, 0.1, False)
            pose[t] = (r.x, r.y, r.orientation, r.speed, r.angular_velocity)
            dst[t] = ctrl.dst
            ncoll[t] = ncol[0]
        out[name + "/pose"] = pose; out[name + "/dst"] = dst; out[name + "/ncoll"] = ncoll
        out[name + "/genome"] = w; out[name + "/fitness"] = np.float64(ctrl.fitness())
        specs.append({"name": name, "stage": stage, "robot": robot_json, "T": T, "layers": sizes, "activation": act})
        print("activation episode", name, "collides", ncol[0], "fitness", repr(float(ctrl.fitness())), "final", pose[-1, :3])
    np.savez_compressed(os.path.join(OUT, "episodes_act.npz"), meta=meta(), specs=json.dumps(specs), **out)


def gen_logs(E):
    """Log files the reference writes (robot.py:158-168, 337-352) for two short episodes -> tests/golden/logs.json."""
    Scripted = make_scripted(E)
    out = {}

    def run(name, stage, robot_json, ctrl, T, genome, kind, layers=(), delta=0.1):
        env = E.load_env(stage)
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
        cfg = json.load(open(os.path.join("..", "conf", robot_json)))
        done = 0
        try:
            for t in range(T):
                r.update(env, delta, True)
                done += 1
        except (IndexError, ValueError) as ex:      # the reference leaves the map: the tick is cut short
            print("  ", name, "raised", type(ex).__name__, "in tick", done)
        T = done
        r.log.flush(); r.sensors_output.flush()
        stem = os.path.join("..", "logs", "robot_" + str(r.identifier))
        out[name] = {"log": open(stem + ".log").read(), "sensors": open(stem + "_sensors.log").read(),
                     "identifier": r.identifier, "robot_json": cfg, "stage": stage, "T": T, "kind": kind, "layers": list(layers),
                     "delta": delta, "genome": [float(g) for g in genome], "acceleration": r.acceleration,
                     "final": [r.x, r.y, r.orientation], "shape": [int(env.shape[0]), int(env.shape[1])]}
        print("logs", name, len(out[name]["log"]), "bytes")

    c = np.random.RandomState(7).uniform(-1, 1, 18) * 0.05
    run("linear_omni", "track_2.png", "robot_omni.json", E.linear_controller(c), 40, c, "linear")
    sc = Scripted(); sc.cmd = (30.0, 2.0)
    run("const_wall", "track_2.png", "robot-neuro.json", sc, 70, [30.0, 2.0], "const")
    sc = Scripted(); sc.cmd = (-12.5, -7.0)
    run("const_stage", "stage.png", "robot.json", sc, 30, [-12.5, -7.0], "const")
    # long ticks: the robot jumps into walls (collide at the proposed position, robot.py:212)
    sc = Scripted(); sc.cmd = (50.0, 7.0)
    run("fast_walls", "stage.png", "robot.json", sc, 60, [50.0, 7.0], "const", delta=1.0)
    # an empty 400 x 300 stage written into the scratch tree: the robot reaches the left / top map border (robot.py:232-247)
    from PIL import Image
    Image.new("L", (400, 300), 255).save(os.path.join("..", "res", "blank.png"))
    sc = Scripted(); sc.cmd = (-45.0, 3.0)
    run("blank_border", "blank.png", "robot.json", sc, 40, [-45.0, 3.0], "const", delta=1.0)
    with open(os.path.join(OUT, "logs.json"), "w") as fp:
        json.dump(out, fp, indent=1)


def gen_ctrl_logs(E):
    """The controllers' own logs (log_step: neurocontroller.py:97-98 -> ../logs/neuro.log, inspyred.py:102-103 ->
    ../logs/inspyred.log) for two episodes -> tests/golden/ctrl_logs.json.  Lines carry time.time(); everything else is kept."""
    out = {}

    def run(name, stage, robot_json, ctrl, T, genome, kind, logfile, layers=(), delta=0.1):
        env = E.load_env(stage)
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
        cfg = json.load(open(os.path.join("..", "conf", robot_json)))
        for t in range(T):
            r.update(env, delta, True)
        ctrl.log.flush()
        out[name] = {"text": open(os.path.join("..", "logs", logfile)).read(), "robot_json": cfg, "stage": stage, "T": T, "kind": kind,
                     "layers": list(layers), "delta": delta, "genome": [float(g) for g in genome], "identifier": r.identifier,
                     "final": [float(r.x), float(r.y), float(r.orientation)]}
        print("ctrl log", name, len(out[name]["text"]), "bytes,", out[name]["text"].count("Odom:"), "lines")

    c = np.random.RandomState(7).uniform(-1, 1, 18) * 0.05
    run("linear_omni", "track_2.png", "robot_omni.json", E.linear_controller(c), 60, c, "linear", "inspyred.log")
    g = np.random.RandomState(21).uniform(-1, 1, 156)
    run("neuro_track", "track_2.png", "robot-neuro.json", E.neuro_controller(g, [10, 7, 4]), 120, g, "neuro", "neuro.log", layers=[5, 10, 7, 4, 2])
    with open(os.path.join(OUT, "ctrl_logs.json"), "w") as fp:
        json.dump(out, fp, indent=1)


def gen_goal(E):
    """is_at_goal of the reference's Sequential_PID_Controller (controllers/pid_controller.py:207-213) evaluated on the pose
    after every tick of two reference episodes, for several goal points -> tests/golden/goal.json."""
    Scripted = make_scripted(E)
    pid = E._import_controller("pid_controller")
    out = {}

    def run(name, stage, robot_json, ctrl, T, genome, kind, goals):
        env = E.load_env(stage)
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot(robot_json, ctrl)
            probe = pid.Sequential_PID_Controller({})      # only its is_at_goal is used, on the episode's robot
        probe.robot = r
        E.u.npdata = env
        hits = {str(i): [] for i in range(len(goals))}
        for t in range(T):
            r.update(env, 0.1, False)
            for i, gpt in enumerate(goals):
                probe.goal = [gpt]
                hits[str(i)].append(bool(probe.is_at_goal()))
        out[name] = {"stage": stage, "robot": robot_json, "kind": kind, "T": T, "genome": [float(g) for g in genome], "goals": [list(g) for g in goals],
                     "at_goal": hits}
        print("goal", name, {k: (v.index(True) if True in v else None) for k, v in hits.items()})

    c = np.random.RandomState(7).uniform(-1, 1, 18) * 0.05
    run("kat3_linear", "track_2.png", "robot_omni.json", E.linear_controller(c), 300, c, "linear",
        [(200, 248), (215.5, 244.25), (230, 236), (100, 100), (193, 249), (10, 10), (229.0, 239.9)])
    sc = Scripted(); sc.cmd = (30.0, 2.0)
    run("kat4_const", "track_2.png", "robot-neuro.json", sc, 100, [30.0, 2.0], "const", [(300, 255), (396, 266), (405.5, 268), (0, 0)])
    with open(os.path.join(OUT, "goal.json"), "w") as fp:
        json.dump(out, fp)


def gen_grids(E):
    """path_planning.generate_grid + node.Node values of the reference for the shipped stages -> tests/golden/grids.npz."""
    import importlib
    pp = importlib.import_module("path_planning")
    node = importlib.import_module("node")
    out = {}
    specs = []
    rs = np.random.RandomState(77)                   # a stage with intermediate grey levels (the shipped ones are 0 / 255 only)
    grey = np.full((300, 200), 255, dtype=np.int64)
    for _ in range(60):
        x, y = rs.randint(0, 290), rs.randint(0, 190)
        grey[x:x + rs.randint(2, 40), y:y + rs.randint(2, 40)] = int(rs.choice([0, 30, 50, 100, 128, 200, 254]))
    out["grey/env"] = grey.astype(np.uint8)
    for stage in STAGES + ["grey"]:
        env = grey if stage == "grey" else E.load_env(stage)
        for div in (10, 25, 64, [20, 13], [7, 40]):
            with contextlib.redirect_stdout(io.StringIO()):
                grid, cw, ch = pp.generate_grid(env, div)
            vals = np.asarray(grid, dtype=np.int32)
            costs = np.asarray([[node.Node(v, (0, 0), (0, 0)).value for v in row] for row in grid], dtype=np.int32)
            key = "%s/%s" % (stage, "x".join(map(str, div)) if isinstance(div, list) else str(div))
            out[key + "/values"] = vals; out[key + "/costs"] = costs
            specs.append({"key": key, "stage": stage, "divider": div})
            print("grid", key, vals.shape, "walls", int((vals == 0).sum()), "cost levels", sorted(set(costs.ravel().tolist())))
    np.savez_compressed(os.path.join(OUT, "grids.npz"), specs=json.dumps(specs), **out)


def gen_host_controllers(E):
    """Episodes driven by reference controllers that have no device form -- naive_controller.Naive_Controller (constant
    command) and pid_controller.Sequential_PID_Controller (goal list, PID terms, obstacle-avoid state machine, backing
    off after a collision; controller-PID.json and controller-PID2.json) -- run by the unmodified reference on stage.png
    with robot.json (scenario-naive-test / scenario-PID).  Per tick: the distances handed to control(), the command it
    returned, the pose after the tick, collide() calls, and the controller's own state -> tests/golden/host_controllers.npz."""
    out, specs = {}, []

    def run(name, ctrl, T, conf_name=None):
        env = counting(E.load_env("stage.png"))
        E.u.npdata = env                                   # utils.display_image publishes the transposed stage (utils.py:302-303)
        with contextlib.redirect_stdout(io.StringIO()):
            r = E.make_robot("robot.json", ctrl)
        ncol = [0]
        orig_oc, orig_ctl = ctrl.on_collision, ctrl.control
        cmds = []

        def oc(pos):
            ncol[0] += 1
            orig_oc(pos)

        def ctl(dst):
            c = orig_ctl(dst)
            cmds.append((float(c[0]), float(c[1])))
            return c
        ctrl.on_collision, ctrl.control = oc, ctl
        R = len(r.sensors)
        pose = np.zeros((T, 5)); dst = np.zeros((T, R)); ncoll = np.zeros(T, dtype=np.int32); aux = np.zeros((T, 4))
        start = (float(r.x), float(r.y), float(r.orientation))
        for t in range(T):
            r.update(env, 0.1, False)
            pose[t] = (r.x, r.y, r.orientation, r.speed, r.angular_velocity)
            dst[t] = [10000.0 if d == math.inf else d for d in ctrl.dst]
            ncoll[t] = ncol[0]
            aux[t] = (len(getattr(ctrl, "goal", [])), getattr(ctrl, "state", 0), float(getattr(ctrl, "backing", 0)), getattr(ctrl, "target_angle", 0.0))
        out[name + "/pose"] = pose; out[name + "/dst"] = dst; out[name + "/ncoll"] = ncoll; out[name + "/aux"] = aux
        out[name + "/cmd"] = np.asarray(cmds, dtype=np.float64); out[name + "/start"] = np.asarray(start)
        specs.append({"name": name, "T": T, "stage": "stage.png", "robot": "robot.json", "conf": conf_name})
        print("host controller", name, "T", T, "collides", ncol[0], "goals left", aux[-1, 0], "final pose", pose[-1, :3])

    naive = E._import_controller("naive_controller")
    run("naive", naive.Naive_Controller({}), 400)
    pidm = E._import_controller("pid_controller")
    import shutil
    for conf_name in ("controller-PID.json", "controller-PID2.json"):
        # the class reads a hard-coded ../conf/controller-PID.json (pid_controller.py:18)
        if conf_name != "controller-PID.json":
            shutil.copyfile(os.path.join(ref_harness.REF, "conf", conf_name), os.path.join(E.tmp, "conf", "controller-PID.json"))
        run("pid_" + conf_name.replace("controller-", "").replace(".json", ""), pidm.Sequential_PID_Controller({}), 1500, conf_name)
    shutil.copyfile(os.path.join(ref_harness.REF, "conf", "controller-PID.json"), os.path.join(E.tmp, "conf", "controller-PID.json"))
    np.savez_compressed(os.path.join(OUT, "host_controllers.npz"), meta=meta(), specs=json.dumps(specs), **out)


def gen_conf():
    """The reference's conf/*.json, byte for byte -> tests/golden/conf/ (robot / controller / scenario numbers are read
    from these copies by tests/common.py and bench.py instead of being typed in)."""
    import shutil
    dst = os.path.join(OUT, "conf")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(ref_harness.REF, "conf")
    for f in sorted(os.listdir(src)):
        if f.endswith(".json"):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    print("conf:", len(os.listdir(dst)), "files")


def main():
    if "--conf-only" in sys.argv:
        gen_conf()
        return
    E = ref_harness.RefEnv()
    try:
        if "--act-only" in sys.argv:
            gen_act_episodes(E)
            return
        if "--grids-only" in sys.argv:
            gen_grids(E)
            return
        if "--goal-only" in sys.argv:
            gen_goal(E)
            return
        if "--host-controllers-only" in sys.argv:
            gen_host_controllers(E)
            return
        if "--logs-only" in sys.argv:
            gen_logs(E)
            return
        if "--ctrl-logs-only" in sys.argv:
            gen_ctrl_logs(E)
            return
        if "--noise-only" in sys.argv:
            gen_noisy_episodes(E)
            return
        gen_conf()
        gen_rays(E)
        gen_ticks(E)
        gen_episodes(E)
        gen_noisy_episodes(E)
        gen_act_episodes(E)
        gen_logs(E)
        gen_ctrl_logs(E)
        gen_goal(E)
        gen_grids(E)
        gen_host_controllers(E)
        # stage bitmaps as the reference loads them (env[x, y], int32) -- fixtures for the GPU box
        for stage in STAGES:
            env = np.ascontiguousarray(E.load_env(stage)).astype(np.uint8)
            np.savez_compressed(os.path.join(OUT, "stage_" + stage.replace(".png", "") + ".npz"), env=env)
    finally:
        E.close()


if __name__ == "__main__":
    main()
