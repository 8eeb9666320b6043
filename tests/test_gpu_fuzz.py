"""Randomised parity: synthetic stages (walls touching the map border, level-50 patches, open borders), random robots
(fractional radius, odd sonar fans, long ticks), random controllers and every lane mapping, against the port-flavour
oracle -- bit for bit.  The shipped stages keep robots 13+ pixels away from the map border; these cases drive the
paths they never reach: Python index wrap and IndexError in the crash ring, rays leaving the map, border collisions,
robots starting inside walls, knife-edge limits."""
import os

import numpy as np
import pytest

import r2p2_b200

pytestmark = pytest.mark.gpu


def random_stage(rs):
    W, H = int(rs.randint(40, 260)), int(rs.randint(40, 200))
    env = np.full((W, H), 255, dtype=np.int32)
    for _ in range(rs.randint(0, 14)):
        w, h = rs.randint(1, max(2, W // 3)), rs.randint(1, max(2, H // 3))
        x, y = rs.randint(-w // 2, W), rs.randint(-h // 2, H)
        env[max(0, x):x + w, max(0, y):y + h] = 0
    if rs.rand() < 0.5:                                   # closed map: a wall ring of random thickness
        t = rs.randint(1, 6)
        env[:t, :] = 0; env[-t:, :] = 0; env[:, :t] = 0; env[:, -t:] = 0
    if rs.rand() < 0.4:                                   # "crash" patches (level 50, robot.py:249)
        for _ in range(rs.randint(1, 4)):
            x, y = rs.randint(0, W - 4), rs.randint(0, H - 4)
            env[x:x + rs.randint(1, 12), y:y + rs.randint(1, 12)] = 50
    if rs.rand() < 0.3:                                   # other grey levels are free space for the robot
        env[rs.randint(0, W), :] = 128
    return env


def random_case(seed):
    rs = np.random.RandomState(seed)
    env = random_stage(rs)
    W, H = env.shape
    R = int(rs.choice([1, 2, 3, 5, 8, 17, 32]))
    rays = sorted(rs.uniform(0, 360, R).round(rs.randint(0, 3)).tolist()) if rs.rand() < 0.6 else [float(i * (360 // R)) for i in range(R)]
    vr1 = float(rs.choice([5, 12.5, 25, 30, 50, 55.5]))
    radius = float(rs.choice([0.5, 1, 2.75, 5, 6, 8.25]))
    vr0 = float(rs.choice([0.1, 0.5, 1, 2]))
    kind = str(rs.choice(["neuro", "neuro", "linear", "const"]))
    hidden = None
    if kind == "neuro":
        hidden = [int(h) for h in rs.randint(1, 12, rs.randint(1, 4))]
        if R == 5 and rs.rand() < 0.5:
            hidden = [10, 7, 4] if rs.rand() < 0.5 else [1]   # the fixed-shape instances
        sizes = [R] + hidden + [2]
        G = sum(a * b for a, b in zip(sizes, sizes[1:]))
        scale = float(rs.choice([0.5, 1.0, 3.0]))
    elif kind == "linear":
        G = 2 * (int(rs.randint(1, R + 3)) + 1)
        scale = float(rs.choice([0.02, 0.1, 0.5]))
    else:
        G, scale = 2, 1.0
    P = int(rs.choice([33, 100, 257]))
    genomes = rs.uniform(-1, 1, (P, G)) * scale
    if kind == "const":
        genomes = np.stack([rs.uniform(-70, 70, P), rs.uniform(-40, 40, P)], axis=1)
    x0 = rs.uniform(-3, W + 3, P); y0 = rs.uniform(-3, H + 3, P); th0 = rs.uniform(-400, 400, P)
    if rs.rand() < 0.5:                                   # integer starts: rays start exactly on pixel centres
        x0, y0 = np.floor(x0), np.floor(y0)
    return dict(env=env, rays=rays, vr=[vr0, vr1], radius=radius, max_speed=float(rs.choice([5, 20, 50, 80])), kind=kind, hidden=hidden,
                activation=str(rs.choice(["tanh", "tanh", "relu", "logistic", "identity"])), genomes=genomes, x0=x0, y0=y0, th0=th0,
                T=int(rs.choice([40, 90])), delta=float(rs.choice([0.1, 0.1, 0.5, 1.0, 2.0])), evolve=bool(rs.rand() < 0.4),
                lprs=[int(v) for v in rs.choice([1, 2, 4, 8, 16, 32], 3, replace=False)],
                noise=(rs.normal(0, 0.5, (P, R * int(rs.choice([40, 90])))) if rs.rand() < 0.3 else None),
                goal=((float(rs.uniform(0, W - 1)), float(rs.uniform(0, H - 1))) if rs.rand() < 0.5 else None))


@pytest.mark.parametrize("seed", range(int(os.environ.get("R2P2_FUZZ_CASES", "160"))))     # soak runs: R2P2_FUZZ_CASES=4000
def test_fuzz_vs_port_oracle(seed, oracle_port):
    c = random_case(1000 + seed)
    layers = ([len(c["rays"])] + c["hidden"] + [2]) if c["hidden"] else []
    ref = oracle_port.run(c["env"], rays=c["rays"], vr=c["vr"], radius=c["radius"], max_speed=c["max_speed"], ctrl=c["kind"], genomes=c["genomes"],
                          x0=c["x0"], y0=c["y0"], theta0=c["th0"], T=c["T"], delta=c["delta"], layers=layers, activation=c["activation"],
                          evolve=c["evolve"], delta_ctrl=40.0, time_budget=2000.0, threads=4, noise=c["noise"], goal=c["goal"])
    st = ref["state"]
    for lpr in c["lprs"] + [0]:
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(c["env"])
        sim.set_robot(sonar_range=c["vr"], radius=c["radius"], max_speed=c["max_speed"], sonars=c["rays"])
        sim.set_controller(c["kind"], hidden_layer=c["hidden"], activation=c["activation"])
        sim.set_mapping(lpr, -1)
        sim.upload_genomes(c["genomes"])
        if c["noise"] is not None:
            sim.set_noise("replay", stream=c["noise"])     # has_noise: the reference's own draws, replayed (robot.py:308-310)
        sim.set_goal(c["goal"])
        sim.reset(c["x0"], c["y0"], c["th0"], time_budget=2000.0)
        sim.run(c["T"], delta=c["delta"], delta_ctrl=40.0, evolve=c["evolve"])
        s, f, cnt = sim.state(), sim.fitness(), sim.counters()
        tag = (seed, lpr, c["kind"], c["hidden"], len(c["rays"]))
        assert np.array_equal(s["flags"] & 23, st["flags"] & 23), tag      # collided, fault, done, goal
        assert np.array_equal(sim.goal_ticks(), st["goal_tick"]), tag
        ok = (st["flags"] & 2) == 0                       # robots the reference keeps alive: everything bit for bit
        for k in ("x", "y", "theta", "speed", "omega", "ncollide", "ticks"):
            assert np.array_equal(s[k][ok], st[k][ok]), (k,) + tag
        assert np.array_equal(f[ok], ref["fitness"][ok]), tag
        bad = ~ok                                         # faulted robots: frozen where the oracle froze them
        assert np.array_equal(s["ticks"][bad], st["ticks"][bad]) and np.array_equal(s["ncollide"][bad], st["ncollide"][bad]), tag
        assert cnt["sonar_samples"] == int(st["nsamp_sonar"].sum()) and cnt["collision_samples"] == int(st["nsamp_coll"].sum()), tag
        sim.close()
