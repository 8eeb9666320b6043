"""Parity of the CUDA path (through the C ABI) on a real B200.

Three anchors:
  * golden vectors recorded from the unmodified reference (rays, single ticks, episodes) -- bit-exact wherever the
    reference's own arithmetic is reproducible (everything except the sklearn/numpy MLP forward);
  * the port-flavour oracle on seeded populations -- bit-exact on every output (it runs the same r2p2_math.h);
  * size-independent properties at BASELINE.json's full sizes (mapping invariance, determinism, counters).
"""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

MAPPINGS = [(1, 0), (1, 1), (8, 0), (8, 1), (32, 0), (32, 1), (4, -1), (16, -1)]


def test_rays_golden():
    """r2p2_cast_rays == utils.search_edge_in_angle_with_limit (reference r2p2/utils.py:420-432) on 8118 rays."""
    import r2p2_b200
    r = common.golden("rays")
    for si, st in enumerate(list(r["stages"])):
        m = r["stage"] == si
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(common.load_stage(st))
        out = sim.cast_rays(r["ox"][m], r["oy"][m], r["angle"][m], r["limit"][m])
        assert np.array_equal(out["status"], r["hit"][m])
        assert np.array_equal(out["hx"], r["hx"][m]) and np.array_equal(out["hy"], r["hy"][m])
        assert np.array_equal(out["nsamp"], r["nsamp"][m].astype(np.uint32))
        plain = sim.cast_rays(r["ox"][m], r["oy"][m], r["angle"][m], r["limit"][m], plain=True)
        for k in out:
            assert np.array_equal(out[k], plain[k]), k
        sim.close()


def test_distance_field_march_equals_plain_march():
    """The distance-field march (skips pixel look-ups where no wall can be) against the sample-by-sample loop on
    400k random rays per stage, long limits included: status, hit point bits, hit step and sample count."""
    import r2p2_b200
    rs = np.random.RandomState(42)
    for st in common.STAGES:
        env = common.load_stage(st)
        W, H = env.shape
        n = 400_000
        ox = np.where(rs.rand(n) < 0.5, rs.randint(0, W, n).astype(float), rs.uniform(-2, W + 2, n))
        oy = np.where(rs.rand(n) < 0.5, rs.randint(0, H, n).astype(float), rs.uniform(-2, H + 2, n))
        ang = np.radians(rs.uniform(-720, 720, n))
        lim = np.where(rs.rand(n) < 0.5, rs.choice([30.0, 45.0, 55.0, 200.0, 1000.0], n), rs.uniform(0.5, 80.0, n))
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(env)
        a = sim.cast_rays(ox, oy, ang, lim)
        b = sim.cast_rays(ox, oy, ang, lim, plain=True)
        for k in a:
            assert np.array_equal(a[k], b[k]), (st, k)
        assert (a["status"] == 1).sum() > n // 4
        sim.close()


@pytest.mark.parametrize("lpr,smem", MAPPINGS)
def test_single_ticks_golden(lpr, smem):
    """One Robot.update from 4200 reference-recorded states, scripted command: pose, speed, sonar distances and
    collide() count bit-identical to the reference for every lane mapping."""
    t = common.golden("ticks")
    sn, rn = list(t["stages"]), list(t["robots"])
    total = 0
    for si, st in enumerate(sn):
        for ri, rj in enumerate(rn):
            for delta in np.unique(t["delta"]):
                m = (t["stage"] == si) & (t["robot"] == ri) & (t["delta"] == delta)
                if not m.any():
                    continue
                sim = common.make_sim(st, rj, "const", lpr=lpr, smem=smem)
                sim.upload_genomes(np.stack([t["v"][m], t["w"][m]], axis=1))
                sim.reset(t["x"][m], t["y"][m], t["theta"][m])
                sim.set_trace(True, 1)
                sim.run(1, delta=float(delta))
                s, tr = sim.state(), sim.trace(1)
                for a, b in (("x", "nx"), ("y", "ny"), ("theta", "ntheta"), ("speed", "nspeed")):
                    assert np.array_equal(s[a], t[b][m]), (st, rj, a)
                assert np.array_equal(s["ncollide"], t["ncollide"][m].astype(np.uint32))
                assert np.array_equal(tr["dst"][0], t["dst"][m][:, :sim.R])
                assert not (s["flags"] & 2).any()
                total += int(m.sum())
                sim.close()
    assert total == len(t["x"])


@pytest.mark.parametrize("lpr", [1, 8, 32])
def test_episodes_golden(lpr):
    """Free-running episodes against the reference: linear / constant-command runs bit-identical on every tick; MLP
    runs (numpy BLAS order + SIMD tanh in the reference) identical in hit steps and collisions, poses and fitness
    within 1e-9 relative (north-star tolerance 1e-5)."""
    e, specs = common.episode_specs()
    for sp in specs:
        n, rob = sp["name"], common.ROBOTS[sp["robot"]]
        sim = common.make_sim(sp["stage"], sp["robot"], sp["kind"], hidden=sp["layers"][1:-1] if sp["layers"] else None, lpr=lpr)
        sim.upload_genomes(e[n + "/genome"][None, :])
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.set_trace(True, sp["T"])
        sim.run(sp["T"], delta=sp["delta"])
        tr = sim.trace(sp["T"])
        pose, gp = tr["pose"][:, 0], e[n + "/pose"]
        assert np.array_equal(np.cumsum(tr["ncoll"][:, 0]), e[n + "/ncoll"]), n
        if sp["kind"] == "neuro":
            assert np.array_equal(np.rint(tr["dst"][:, 0]), np.rint(e[n + "/dst"])), n
            assert np.max(np.abs(pose[:, :3] - gp[:, :3]) / np.maximum(1, np.abs(gp[:, :3]))) < 1e-9, n
            gf = e[n + "/fitness"][-1]
            assert abs(sim.fitness()[0] - gf) <= 1e-9 * abs(gf), n
        else:
            assert np.array_equal(pose, gp), n
            assert np.array_equal(tr["dst"][:, 0], e[n + "/dst"]), n
            if sp["kind"] == "linear":
                assert sim.fitness()[0] == e[n + "/fitness"][-1], n
        sim.close()


CASES = [("track_2.png", "robot-neuro.json", "neuro", [10, 7, 4], 156, 1.0),
         ("stage.png", "robot-neuro.json", "neuro", [10, 7, 4], 156, 1.0),
         ("track.png", "robot-track.json", "neuro", [1], 7, 3.0),
         ("track_2.png", "robot_omni.json", "linear", None, 18, 0.05),
         ("track_2.png", "robot-neuro.json", "linear", None, 12, 0.08)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%s" % (c[0], c[1], c[2]))
@pytest.mark.parametrize("evolve", [False, True])
def test_population_vs_port_oracle(case, evolve, oracle_port):
    """Seeded populations, 300 ticks: every pose, fitness, flag, tick count, collide count and the sonar / collision
    sample counters bit-identical to the port-flavour oracle, for 1-, 8- and 32-lane mappings."""
    stage, rj, kind, hidden, G, scale = case
    P, T = 384, 300
    genomes = np.random.RandomState(1234).uniform(-1, 1, (P, G)) * scale
    rob = common.ROBOTS[rj]
    kw = common.oracle_robot_kwargs(rj)
    layers = [len(kw["rays"])] + hidden + [2] if hidden else []
    ref = oracle_port.run(common.load_stage(stage), ctrl=kind, genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0, T=T, layers=layers,
                          evolve=evolve, delta_ctrl=33.0, time_budget=20000.0, threads=8, **kw)
    st = ref["state"]
    for lpr, smem in ((1, -1), (8, 1), (8, 0), (32, -1)):
        sim = common.make_sim(stage, rj, kind, hidden=hidden, lpr=lpr, smem=smem)
        sim.upload_genomes(genomes)
        sim.reset(rob["x"], rob["y"], 0.0, time_budget=20000.0)
        sim.run(T, delta=0.1, delta_ctrl=33.0, evolve=evolve)
        s, f, c = sim.state(), sim.fitness(), sim.counters()
        for k in ("x", "y", "theta", "speed", "omega", "ncollide", "flags", "ticks"):
            assert np.array_equal(s[k], st[k]), (k, lpr, smem)
        assert np.array_equal(f, ref["fitness"])
        assert c["sonar_samples"] == int(st["nsamp_sonar"].sum()) and c["collision_samples"] == int(st["nsamp_coll"].sum())
        assert c["robot_steps"] == int(st["ticks"].sum())
        sim.close()


def test_trace_matches_port_oracle(oracle_port):
    """Per-tick trace (hit step, hit pixel, distance, samples per ray, pose) bit-identical to the oracle."""
    stage, rj = "track_2.png", "robot-neuro.json"
    P, T = 64, 120
    genomes = np.random.RandomState(5).uniform(-1, 1, (P, 156))
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    ref = oracle_port.run(common.load_stage(stage), ctrl="neuro", genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0, T=T,
                          layers=[5, 10, 7, 4, 2], trace=True, **kw)
    for lpr in (1, 8, 32):
        sim = common.make_sim(stage, rj, "neuro", hidden=[10, 7, 4], lpr=lpr)
        sim.upload_genomes(genomes)
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.set_trace(True, T)
        sim.run(T)
        tr = sim.trace(T)
        for k in ("pose", "dst", "hitk", "hitpx", "nsamp", "ncoll"):
            assert np.array_equal(tr[k], ref[k]), (k, lpr)
        sim.close()


def test_per_tick_api_equals_one_launch():
    """Robot.update drop-in use: T launches of one tick == one launch of T ticks (state round-trips HBM)."""
    stage, rj = "track_2.png", "robot-neuro.json"
    genomes = np.random.RandomState(11).uniform(-1, 1, (200, 156))
    rob = common.ROBOTS[rj]
    a = common.make_sim(stage, rj, "neuro", hidden=[10, 7, 4])
    b = common.make_sim(stage, rj, "neuro", hidden=[10, 7, 4])
    for s in (a, b):
        s.upload_genomes(genomes)
        s.reset(rob["x"], rob["y"], 0.0)
    a.run(50)
    for _ in range(50):
        b.run(1)
    sa, sb = a.state(), b.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a.fitness(), b.fitness())
    assert a.counters() == b.counters()
    a.close(); b.close()


def test_start_poses_per_robot_and_faults(oracle_port):
    """Per-robot start poses (BASELINE config D/E style) incl. robots that start inside a wall or off the map: the
    reference raises there (IndexError / ValueError); the device sets the fault flag and freezes the robot."""
    stage, rj = "track.png", "robot-track.json"
    env = common.load_stage(stage)
    rs = np.random.RandomState(3)
    P = 512
    x0 = rs.uniform(-20, env.shape[0] + 20, P); y0 = rs.uniform(-20, env.shape[1] + 20, P); th0 = rs.uniform(0, 360, P)
    x0[:4] = [np.nan, np.inf, 100.0, 1e300]
    genomes = rs.uniform(-3, 3, (P, 7))
    kw = common.oracle_robot_kwargs(rj)
    ref = oracle_port.run(env, ctrl="neuro", genomes=genomes, x0=x0, y0=y0, theta0=th0, T=80, layers=[5, 1, 2], threads=4, **kw)
    for lpr in (1, 8, 32):
        sim = common.make_sim(stage, rj, "neuro", hidden=[1], lpr=lpr)
        sim.upload_genomes(genomes)
        sim.reset(x0, y0, th0)
        sim.run(80)
        s = sim.state()
        assert np.array_equal(s["flags"] & 7, ref["state"]["flags"] & 7)
        ok = (ref["state"]["flags"] & 2) == 0          # compare poses of the robots the reference keeps alive
        for k in ("x", "y", "theta", "ncollide", "ticks"):
            assert np.array_equal(s[k][ok], ref["state"][k][ok]), k
        assert np.array_equal(sim.fitness()[ok], ref["fitness"][ok])
        assert int(((ref["state"]["flags"] & 2) != 0).sum()) > 20
        sim.close()


def test_full_size_properties():
    """BASELINE config B at full size (1024 x 1000): mapping invariance (1/8/32 lanes, smem / L1), run-to-run
    determinism, sample counters consistent with the trace-free counters of a second run."""
    stage, rj = "track_2.png", "robot-neuro.json"
    genomes = np.random.RandomState(1234).uniform(-1, 1, (1024, 156))
    rob = common.ROBOTS[rj]
    res = []
    for lpr, smem in ((32, 1), (32, 0), (8, 1), (1, 0), (32, 1)):
        sim = common.make_sim(stage, rj, "neuro", hidden=[10, 7, 4], lpr=lpr, smem=smem)
        sim.upload_genomes(genomes)
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.run(1000)
        res.append((sim.fitness(), sim.state(), sim.counters()))
        sim.close()
    f0, s0, c0 = res[0]
    assert c0["robot_steps"] == 1024 * 1000 and c0["sonar_samples"] > 100 * 1024 * 1000
    for f, s, c in res[1:]:
        assert np.array_equal(f, f0) and c == c0
        for k in s0:
            assert np.array_equal(s[k], s0[k]), k


def test_full_size_vs_oracle_sample(oracle_port):
    """Config B: a 96-robot sample of the full 1000-tick run bit-identical to the port oracle (fitness + counters)."""
    stage, rj = "track_2.png", "robot-neuro.json"
    genomes = np.random.RandomState(1234).uniform(-1, 1, (1024, 156))
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    sim = common.make_sim(stage, rj, "neuro", hidden=[10, 7, 4])
    sim.upload_genomes(genomes)
    sim.reset(rob["x"], rob["y"], 0.0)
    sim.run(1000)
    f = sim.fitness()
    sel = np.arange(0, 1024, 11)[:96]
    ref = oracle_port.run(common.load_stage(stage), ctrl="neuro", genomes=genomes[sel], x0=rob["x"], y0=rob["y"], theta0=0.0, T=1000,
                          layers=[5, 10, 7, 4, 2], threads=8, **kw)
    assert np.array_equal(f[sel], ref["fitness"])
    sim.close()


def test_activations_and_widths(oracle_port):
    """relu / logistic / identity activations and a wider net (32-16-8-2 on 32 list-form rays, BASELINE config E shape)."""
    import r2p2_b200
    env = common.load_stage("stage.png")
    rays = [i * 11.25 for i in range(32)]
    P, T = 128, 60
    rs = np.random.RandomState(9)
    free = np.argwhere(env[8:-8, 8:-8] != 0) + 8
    pick = free[rs.randint(0, len(free), P)]
    x0, y0, th0 = pick[:, 0].astype(float) + 0.25, pick[:, 1].astype(float) + 0.5, rs.uniform(0, 360, P)
    for act in ("tanh", "relu", "logistic", "identity"):
        genomes = rs.uniform(-1, 1, (P, 32 * 16 + 16 * 8 + 8 * 2)) * (0.3 if act in ("relu", "identity") else 1.0)
        ref = oracle_port.run(env, rays=rays, vr=(0.5, 25), radius=5, max_speed=50, ctrl="neuro", genomes=genomes, x0=x0, y0=y0, theta0=th0,
                              T=T, layers=[32, 16, 8, 2], activation=act, threads=4)
        for lpr in (1, 8, 32):
            sim = r2p2_b200.BatchSim(0)
            sim.set_stage(env)
            sim.set_robot(sonar_range=(0.5, 25), radius=5, max_speed=50, sonars=rays)
            sim.set_controller("neuro", hidden_layer=[16, 8], activation=act)
            sim.set_mapping(lpr, -1)
            sim.upload_genomes(genomes)
            sim.reset(x0, y0, th0)
            sim.run(T)
            s = sim.state()
            assert np.array_equal(s["flags"], ref["state"]["flags"]), (act, lpr)
            ok = (s["flags"] & 2) == 0
            assert np.array_equal(s["x"][ok], ref["state"]["x"][ok]) and np.array_equal(s["theta"][ok], ref["state"]["theta"][ok]), (act, lpr)
            assert np.array_equal(sim.fitness()[ok], ref["fitness"][ok]), (act, lpr)
            sim.close()


def test_error_conventions():
    """Errors come back as codes + message (never an abort): bad call order, bad sizes."""
    import r2p2_b200
    sim = r2p2_b200.BatchSim(0)
    with pytest.raises(r2p2_b200.R2P2Error):
        sim.run(1)                                           # nothing configured
    sim.set_stage(common.load_stage("track.png"))
    sim.set_robot(sonar_range=(0.5, 25), radius=5, max_speed=50, sonars=[0, 45])
    sim.set_controller("neuro", hidden_layer=[3])
    sim.upload_genomes(np.zeros((4, 5)))                     # wrong genome length for 2-3-2
    sim.reset(100.0, 100.0)
    with pytest.raises(r2p2_b200.R2P2Error) as ei:
        sim.run(1)
    assert "genome length" in str(ei.value)
    with pytest.raises(r2p2_b200.R2P2Error):
        sim.set_mapping(3, 0)
    sim.close()


@pytest.mark.parametrize("lpr", [1, 8, 32])
def test_noise_replay_golden(lpr):
    """Robot.has_noise = True (reference default): with the reference's own np.random.normal draws replayed on the device
    (one per unclamped ray, ray order), poses / noisy distances / collide counts / number of draws are bit-identical to the
    seeded reference runs (linear and constant-command controllers; MLP runs to 1e-9)."""
    import json
    e = common.golden("episodes_noise")
    for sp in json.loads(str(e["specs"])):
        n, rob = sp["name"], common.ROBOTS[sp["robot"]]
        sim = common.make_sim(sp["stage"], sp["robot"], sp["kind"], hidden=sp["layers"][1:-1] if sp["layers"] else None, lpr=lpr)
        np.random.seed(sp["seed"])
        stream = np.random.normal(0, 0.5, size=sp["T"] * sim.R + 8)
        sim.upload_genomes(e[n + "/genome"][None, :])
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.set_noise("replay", stream=stream[None, :])
        sim.set_trace(True, sp["T"])
        sim.run(sp["T"])
        tr = sim.trace(sp["T"])
        assert int(sim.noise_draws()[0]) == int(e[n + "/ndraws"]), n
        assert np.array_equal(np.cumsum(tr["ncoll"][:, 0]), e[n + "/ncoll"]), n
        if sp["kind"] == "neuro":
            assert np.max(np.abs(tr["pose"][:, 0, :3] - e[n + "/pose"][:, :3])) < 1e-9 * 500, n
        else:
            assert np.array_equal(tr["pose"][:, 0], e[n + "/pose"]), n
            assert np.array_equal(tr["dst"][:, 0], e[n + "/dst"]), n
        sim.close()


def test_noise_replay_population_vs_oracle(oracle_port):
    """Per-robot streams on a population: bit-identical to the port oracle for every mapping."""
    stage, rj = "track_2.png", "robot_omni.json"
    P, T = 200, 150
    rs = np.random.RandomState(8)
    genomes = rs.uniform(-1, 1, (P, 18)) * 0.05
    streams = rs.normal(0, 0.5, (P, T * 8))
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    ref = oracle_port.run(common.load_stage(stage), ctrl="linear", genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0, T=T, noise=streams, threads=4, **kw)
    for lpr in (1, 4, 8, 32):
        sim = common.make_sim(stage, rj, "linear", lpr=lpr)
        sim.upload_genomes(genomes)
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.set_noise("replay", stream=streams)
        sim.run(T)
        s = sim.state()
        assert np.array_equal(s["x"], ref["state"]["x"]) and np.array_equal(s["theta"], ref["state"]["theta"]), lpr
        assert np.array_equal(sim.fitness(), ref["fitness"]) and np.array_equal(sim.noise_draws(), ref["state"]["ndraws"]), lpr
        sim.close()


def test_noise_device_generator_statistics():
    """Mode "device": Philox + Box-Muller.  Draws are N(0, 0.5): sample mean / std / kurtosis over ~1e5 draws, reproducible
    for a seed, different across seeds and robots."""
    stage, rj = "stage.png", "robot_omni.json"
    P, T = 4096, 8
    genomes = np.zeros((P, 2))                                  # constant command (0, 0): the robots stand still
    res = {}
    for seed, mode in ((0, "off"), (1, "device"), (1, "device"), (2, "device")):
        sim = common.make_sim(stage, rj, "const", lpr=8)
        sim.upload_genomes(genomes)
        sim.reset(260.0, 200.0, np.linspace(0, 359, P))
        sim.set_noise(mode, seed=seed)
        sim.set_trace(True, T)
        sim.run(T)
        res.setdefault((seed, mode), []).append(sim.trace(T)["dst"].copy())
        sim.close()
    clean = res[(0, "off")][0]
    a, a2 = res[(1, "device")]
    b = res[(2, "device")][0]
    assert np.array_equal(a, a2) and not np.array_equal(a, b)
    unclamped = clean < 55.0                                     # clamped rays (dst == L) get no noise
    assert np.array_equal(a[~unclamped], clean[~unclamped])
    d = (a - clean)[unclamped]
    assert d.size > 30_000
    assert abs(d.mean()) < 0.01 and abs(d.std() - 0.5) < 0.01
    assert abs(((d / d.std()) ** 4).mean() - 3.0) < 0.1
    assert abs(np.corrcoef(d[:-1], d[1:])[0, 1]) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("lpr", [1, 4, 32])
def test_activation_episodes_golden(lpr):
    """relu / logistic / identity networks and other layer shapes against episodes recorded from the reference."""
    import json
    e = common.golden("episodes_act")
    for sp in json.loads(str(e["specs"])):
        n = sp["name"]
        rob = common.ROBOTS[sp["robot"]]
        sim = common.make_sim(sp["stage"], sp["robot"], "neuro", hidden=sp["layers"][1:-1], activation=sp["activation"], lpr=lpr)
        sim.upload_genomes(e[n + "/genome"][None, :])
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.set_trace(True, sp["T"])
        sim.run(sp["T"])
        tr = sim.trace(sp["T"])
        gp = e[n + "/pose"]
        assert np.array_equal(np.cumsum(tr["ncoll"][:, 0]), e[n + "/ncoll"]), n
        assert np.array_equal(np.rint(tr["dst"][:, 0]), np.rint(e[n + "/dst"])), n
        assert np.max(np.abs(tr["pose"][:, 0, :3] - gp[:, :3]) / np.maximum(1, np.abs(gp[:, :3]))) < 1e-9, n
        gf = float(e[n + "/fitness"])
        assert abs(sim.fitness()[0] - gf) <= 1e-9 * max(1.0, abs(gf)), n
        sim.close()


@pytest.mark.gpu
def test_checkpoint_resume_is_bit_exact():
    """save_state after 120 ticks, finish the episode, load the snapshot into a fresh handle, finish again: same bits
    (poses, fitness, flags, counters, noise draws), with evolve semantics and replayed noise on."""
    rj = "robot-neuro.json"
    rob = common.ROBOTS[rj]
    P = 300
    genomes = np.random.RandomState(77).uniform(-1, 1, (P, 156))
    noise = np.random.RandomState(78).normal(0, 0.5, (P, 5 * 400))

    def fresh():
        sim = common.make_sim("track_2.png", rj, "neuro", hidden=[10, 7, 4])
        sim.upload_genomes(genomes)
        sim.set_noise("replay", stream=noise)
        sim.set_goal((240.0, 252.0))
        return sim
    a = fresh()
    a.reset(rob["x"], rob["y"], 0.0, time_budget=30000.0)
    a.run(120, delta_ctrl=100.0, evolve=True)
    snap = a.save_state()
    a.run(280, delta_ctrl=100.0, evolve=True)
    b = fresh()
    b.load_state(snap)
    b.run(280, delta_ctrl=100.0, evolve=True)
    sa, sb = a.state(), b.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a.fitness(), b.fitness()) and a.counters() == b.counters()
    assert np.array_equal(a.noise_draws(), b.noise_draws()) and np.array_equal(a.goal_ticks(), b.goal_ticks())
    with pytest.raises(ValueError):
        b.load_state(snap[:-8])
    a.close(); b.close()


@pytest.mark.gpu
def test_sharding_invariance():
    """SURVEY 8e result check: the fitness vector does not depend on how the population is cut into shards (the automatic
    lane mapping changes with the shard size: 4 096 robots -> 8 lanes, 1 024 -> 32, 512 -> 32), so N = 1, 2, 4, 8 GPUs give
    bit-identical generations."""
    from r2p2_b200.evaluator import PopulationEvaluator
    from r2p2_b200.distributed import shard_bounds
    rob = common.ROBOTS["robot-neuro.json"]
    P, T = 4096, 200
    genomes = np.random.RandomState(5).uniform(-1, 1, (P, 156))
    ev = PopulationEvaluator(common.load_stage("stage.png"), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                             sonars=rob["sonars"], controller="neuro", hidden_layer=[10, 7, 4], x=rob["x"], y=rob["y"], ticks=T,
                             delta_ctrl=100.0, time_budget=15000.0, evolve=True)
    full = ev.evaluate(genomes).copy()
    for world in (2, 4, 8):
        parts = [ev.evaluate(genomes[lo:hi]).copy() for lo, hi in (shard_bounds(P, world, r) for r in range(world))]
        assert np.array_equal(np.concatenate(parts), full), world
