"""Shared test data: stages, robot configs (numbers of the reference's conf/*.json), golden loaders."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
STAGES = ["track_2.png", "stage.png", "track.png"]

CONF = os.path.join(GOLDEN, "conf")          # byte copies of the reference's conf/*.json (make_golden.py --conf-only)


def conf(name):
    """One of the reference's conf/*.json files, as shipped."""
    with open(os.path.join(CONF, name)) as fp:
        return json.load(fp)


# the robot jsons shipped by the reference (conf/robot-neuro.json, robot_omni.json, robot-track.json, robot.json)
ROBOTS = {name: conf(name) for name in ("robot-neuro.json", "robot_omni.json", "robot-track.json", "robot.json")}


def load_stage(name):
    """env[x, y] int32 exactly as the reference builds it (fixture recorded by tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "stage_%s.npz" % name.replace(".png", "")))["env"].astype(np.int32)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def episode_specs():
    e = golden("episodes")
    return e, json.loads(str(e["specs"]))


def oracle_robot_kwargs(rj):
    from oracle.oracle import expand_sonars
    r = ROBOTS[rj]
    return dict(rays=expand_sonars(r["sonars"]), vr=r["sonar_range"], radius=r["radius"], max_speed=r["max_speed"])


def make_sim(stage, rj, ctrl, hidden=None, activation="tanh", lpr=0, smem=-1, device=0):
    import r2p2_b200
    r = ROBOTS[rj]
    sim = r2p2_b200.BatchSim(device)
    sim.set_stage(load_stage(stage) if isinstance(stage, str) else stage)
    sim.set_robot(sonar_range=r["sonar_range"], radius=r["radius"], max_speed=r["max_speed"], sonars=r["sonars"])
    sim.set_controller(ctrl, hidden_layer=hidden, activation=activation)
    sim.set_mapping(lpr, smem)
    return sim


# ---- population-scale reference episodes (tests/golden/make_population_golden.py) ----------------------------------
def synthetic_crop(n=1024):
    """The n x n crop of bench.py's synthetic stage the population goldens' shape E ran on (black border restored)."""
    import bench
    env = bench.synthetic_stage()[:n, :n].copy()
    env[:8, :] = 0; env[-8:, :] = 0; env[:, :8] = 0; env[:, -8:] = 0
    return env


def population_groups():
    """Episodes of tests/golden/population.npz grouped by shape: (shape, stage env, robot dict, hidden, T, episodes)."""
    import bench
    g = np.load(os.path.join(GOLDEN, "population.npz"))
    eps = json.loads(str(g["specs"]))
    groups = {}
    for e in eps:
        groups.setdefault(e["shape"], []).append(e)
    out = []
    for shape, lst in groups.items():
        e0 = lst[0]
        env = synthetic_crop(json.loads(str(g["meta"]))["synth_n"]) if e0["stage"].startswith("synthetic") else load_stage(e0["stage"])
        rob = dict(bench.ROBOTS["robot-synth32"]) if e0["robot"] == "robot-synth32.json" else conf(e0["robot"])
        out.append((shape, env, rob, e0["hidden"], e0["T"], lst))
    return g, out


def population_inputs(rob, hidden, lst):
    """Genomes and start poses of a group, as the generator built them."""
    from oracle.oracle import expand_sonars
    R = len(expand_sonars(rob["sonars"]))
    sizes = [R] + list(hidden) + [2]
    G = sum(a * b for a, b in zip(sizes, sizes[1:]))
    genomes = np.stack([np.random.RandomState(e["seed"]).uniform(-1, 1, G) * e["scale"] for e in lst])
    x0 = np.array([e.get("x0", float(rob["x"])) for e in lst], dtype=np.float64)
    y0 = np.array([e.get("y0", float(rob["y"])) for e in lst], dtype=np.float64)
    th0 = np.array([e.get("theta0", float(rob["orientation"])) for e in lst], dtype=np.float64)
    return genomes, x0, y0, th0, sizes


def population_divergence(g, lst, dst, ncoll_per_tick, pose, fitness_at):
    """Compare a replay (dst [T, P, R], collide() calls per tick [T, P], pose [T, P, 5], fitness_at(t) -> [P] or None) with
    the reference's record.  Returns a dict of rates; `first` is the first tick at which the rounded sonar distances or the
    cumulative collide() count differ (T = never)."""
    T, P, _ = dst.shape
    cum = np.cumsum(ncoll_per_tick, axis=0)
    first = np.full(P, T, dtype=np.int64)
    pose_rel = np.zeros(P)
    for p, e in enumerate(lst):
        key = "ep%04d/" % e["index"]
        bad = (np.rint(dst[:, p]).astype(np.int64) != g[key + "dstq"].astype(np.int64)).any(axis=1) | (cum[:, p] != g[key + "ncoll"])
        if bad.any():
            first[p] = int(np.argmax(bad))
        for t, x, y, th, fit in g[key + "checks"]:
            t = int(t)
            if t < first[p]:
                ref = np.array([x, y, th])
                pose_rel[p] = max(pose_rel[p], float(np.max(np.abs(pose[t, p, :3] - ref) / np.maximum(1.0, np.abs(ref)))))
    return {"episodes": P, "ticks_per_episode": T, "episodes_identical": int((first == T).sum()),
            "steps_before_divergence": int(first.sum()), "steps": int(P * T),
            "first_divergence_ticks": sorted(int(f) for f in first if f < T),
            "pose_max_rel_diff_before_divergence": float(pose_rel.max())}
