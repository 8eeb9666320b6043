#!/usr/bin/env python3
"""Record ~1e6 steps of the UNMODIFIED reference's neuro path at population scale.  Build-container only.

    python tests/golden/make_population_golden.py [--workers 8]      # rewrites tests/golden/population.npz

SURVEY.md hard-part 1 asks for the rate at which the device (ordered-FMA MLP, table tanh, restated sin/cos) and the
reference (sklearn -> OpenBLAS matmul, numpy SIMD tanh, glibc) part ways on integer outcomes to be MEASURED, not
assumed.  This script runs the reference's own Robot.update with its own Neuro_controller (neurocontroller.py:67-95)
for several hundred free-running episodes of the BASELINE.json shapes and stores, per episode, what a replay needs to be
compared tick by tick: the rounded sonar distances (= hit steps) of every tick, the cumulative collide() count, pose
check-points and the final fitness.  tests/test_population_golden.py replays them through the oracle (CPU) and the
CUDA path (GPU) and reports the divergence rates.

  shape B   scenario-neuro:   track_2.png, conf/robot-neuro.json, MLP 5-10-7-4-2          220 episodes x 2000 ticks
  shape C3  north star:       stage.png,   conf/robot-neuro.json, MLP 5-10-7-4-2          160 episodes x 2000 ticks
  shape D   scenario-track:   track.png,   conf/robot-track.json, MLP 5-10-7-4-2, 4 headings   120 episodes x 2000 ticks
  shape E   synthetic sweep:  1024 x 1024 crop of bench.py's synthetic stage, 32 list-form rays, MLP 32-16-8-2,
                              start poses in free space                                    40 episodes x 300 ticks
Genomes are np.random.RandomState(seed).uniform(-1, 1, G) * scale (mirrors evolution.py:65); seeds are stored.
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

CHECK_EVERY = 250
SYNTH_N = 1024            # the reference loads stages from PNG files: a 1024^2 crop of the 4096^2 synthetic stage keeps it small


def synth_stage():
    import bench
    env = bench.synthetic_stage()[:SYNTH_N, :SYNTH_N].copy()
    env[:8, :] = 0; env[-8:, :] = 0; env[:, :8] = 0; env[:, -8:] = 0
    return env


def episode_list():
    eps = []
    for i in range(220):
        eps.append(dict(shape="B", stage="track_2.png", robot="robot-neuro.json", hidden=[10, 7, 4], seed=10_000 + i, scale=1.0, T=2000))
    for i in range(160):
        eps.append(dict(shape="C3", stage="stage.png", robot="robot-neuro.json", hidden=[10, 7, 4], seed=20_000 + i, scale=1.0, T=2000))
    for i in range(120):
        eps.append(dict(shape="D", stage="track.png", robot="robot-track.json", hidden=[10, 7, 4], seed=30_000 + i, scale=1.0, T=2000,
                        theta0=float((0, 90, 180, 270)[i % 4])))
    env = synth_stage()
    rs = np.random.RandomState(4040)
    free = np.argwhere(env[16:-16, 16:-16] != 0) + 16
    c = 0
    while c < 40:
        x, y = free[rs.randint(len(free))]
        if (env[x - 8:x + 9, y - 8:y + 9] == 0).any():
            continue
        eps.append(dict(shape="E", stage="synthetic1024.png", robot="robot-synth32.json", hidden=[16, 8], seed=40_000 + c, scale=1.0, T=300,
                        x0=float(x) + 0.25, y0=float(y) + 0.5, theta0=float(rs.uniform(0, 360))))
        c += 1
    for i, e in enumerate(eps):
        e["index"] = i
    return eps


_E = None


def _env():
    global _E
    if _E is None:
        import ref_harness
        _E = ref_harness.RefEnv()
        from PIL import Image
        Image.fromarray(np.ascontiguousarray(synth_stage().T), mode="L").save(os.path.join(_E.tmp, "res", "synthetic1024.png"))
    return _E


def run_episode(ep):
    E = _env()
    import bench
    base = dict(bench.ROBOTS["robot-synth32"], id=0) if ep["robot"] == "robot-synth32.json" else json.load(open(os.path.join(HERE, "conf", ep["robot"])))
    rob = dict(base)
    rob["id"] = os.getpid()                                  # per-process log file names
    for k_src, k_dst in (("x0", "x"), ("y0", "y")):
        if k_src in ep:
            rob[k_dst] = ep[k_src]
    rname = "robot-pop-%d.json" % os.getpid()
    E.write_conf(rname, rob)
    R = len(rob["sonars"])
    sizes = [R] + ep["hidden"] + [2]
    G = sum(a * b for a, b in zip(sizes, sizes[1:]))
    w = np.random.RandomState(ep["seed"]).uniform(-1, 1, G) * ep["scale"]
    ctrl = E.neuro_controller(w, ep["hidden"])
    env = E.load_env(ep["stage"])
    with contextlib.redirect_stdout(io.StringIO()):
        r = E.make_robot(rname, ctrl)
    if "theta0" in ep:
        r.orientation = ep["theta0"]
    ncol = [0]
    orig = ctrl.on_collision

    def oc(pos):
        ncol[0] += 1
        orig(pos)
    ctrl.on_collision = oc
    T = ep["T"]
    dstq = np.zeros((T, R), dtype=np.uint8)
    ncoll = np.zeros(T, dtype=np.int32)
    checks = []
    for t in range(T):
        r.update(env, 0.1, False)
        dstq[t] = np.rint(ctrl.dst)
        ncoll[t] = ncol[0]
        if (t + 1) % CHECK_EVERY == 0 or t == T - 1:
            checks.append((t, r.x, r.y, r.orientation, ctrl.fitness()))
    for f in (r.file, r.sensors_file, r.benchmark_file) if hasattr(r, "file") else ():
        with contextlib.suppress(Exception):
            f.close()
    return ep["index"], dstq, ncoll, np.asarray(checks, dtype=np.float64)


def main():
    workers = int(sys.argv[sys.argv.index("--workers") + 1]) if "--workers" in sys.argv else (os.cpu_count() or 1)
    eps = episode_list()
    if "--limit" in sys.argv:
        eps = eps[::max(1, len(eps) // int(sys.argv[sys.argv.index("--limit") + 1]))]
    out = {}
    with mp.Pool(workers) as pool:
        for n, (idx, dstq, ncoll, checks) in enumerate(pool.imap_unordered(run_episode, eps, chunksize=2)):
            out["ep%04d/dstq" % idx] = dstq
            out["ep%04d/ncoll" % idx] = ncoll
            out["ep%04d/checks" % idx] = checks
            if n % 20 == 0:
                print("episode", n, "/", len(eps), flush=True)
    import platform
    import sklearn
    meta = {"python": platform.python_version(), "numpy": np.__version__, "sklearn": sklearn.__version__, "libc": platform.libc_ver()[1],
            "check_every": CHECK_EVERY, "synth_n": SYNTH_N, "steps": int(sum(e["T"] for e in eps))}
    np.savez_compressed(os.path.join(HERE, "population.npz"), specs=json.dumps(eps), meta=json.dumps(meta), **out)
    print("wrote population.npz:", len(eps), "episodes,", meta["steps"], "reference steps")


if __name__ == "__main__":
    main()
