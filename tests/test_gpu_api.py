"""C-ABI behaviour around the hot path: population size bookkeeping, pose writes, checkpoints with headers, EA
checkpoints, noise-stream exhaustion, the global robot index of sharded evaluations, the caller's current device."""
import numpy as np
import pytest

import common
from r2p2_b200 import ea

pytestmark = pytest.mark.gpu


def _sim(lpr=0):
    return common.make_sim("track_2.png", "robot-neuro.json", "neuro", hidden=[10, 7, 4], lpr=lpr)


def test_population_size_follows_the_library():
    """r2p2_ga_evaluate resizes the population to the evaluated shard; every read sizes its buffer from the library
    (round-1 advice: a cached P == 1 made sim.fitness() overrun an 8-byte buffer)."""
    rob = common.ROBOTS["robot-neuro.json"]
    sim = _sim()
    sim.upload_genomes(np.zeros((1, 156)))                    # the 1-row dummy population of examples/evolve_neuro.py
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
    assert (sim.P, sim.G) == (1, 156)
    ga = ea.DeviceGA(sim, 512, 156, seed=11)
    ga.evaluate(40, first=128, count=200)                     # a shard, as ShardedDeviceGA does
    assert (sim.P, sim.G) == (200, 156)
    f, st = sim.fitness(), sim.state()
    assert f.shape == (200,) and st["x"].shape == (200,) and sim.goal_ticks().shape == (200,) and sim.noise_draws().shape == (200,)
    _, gfit = ga.read(population=False)
    assert np.array_equal(gfit[128:328], f)
    with pytest.raises(ValueError):
        sim.fitness(out=np.empty(8))
    sim.close()


def test_write_pose_is_a_plain_assignment(oracle_port):
    """Robot.set_position / `robot.orientation = ...` on a running robot: pose changes, odometry and fitness carry on
    (the jump is counted as travelled distance at the next control(), neurocontroller.py:69-71)."""
    rj = "robot-neuro.json"
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    P = 64
    genomes = np.random.RandomState(3).uniform(-1, 1, (P, 156))
    sim = _sim()
    sim.upload_genomes(genomes)
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
    sim.run(30)
    before = sim.state()
    nx, ny, nth = before["x"][8:24] + 3.0, before["y"][8:24] - 2.0, before["theta"][8:24] + 45.0
    sim.write_pose(8, x=nx, y=ny, theta=nth)
    after = sim.state()
    assert np.array_equal(after["x"][8:24], nx) and np.array_equal(after["theta"][8:24], nth)
    assert np.array_equal(after["x"][:8], before["x"][:8]) and np.array_equal(after["ticks"], before["ticks"])
    sim.run(30)
    # the oracle with the same assignment between the two runs
    ref = oracle_port.run(common.load_stage("track_2.png"), ctrl="neuro", genomes=genomes, x0=float(rob["x"]), y0=float(rob["y"]), theta0=0.0,
                          T=30, layers=[5, 10, 7, 4, 2], **kw)
    st = ref["state"].copy()
    st["x"][8:24], st["y"][8:24], st["theta"][8:24] = nx, ny, nth
    ref2 = oracle_port.resume(common.load_stage("track_2.png"), st, ctrl="neuro", genomes=genomes, T=30, layers=[5, 10, 7, 4, 2], **kw)
    s = sim.state()
    for k in ("x", "y", "theta", "ncollide", "flags", "ticks"):
        assert np.array_equal(s[k], ref2["state"][k]), k
    assert np.array_equal(sim.fitness(), ref2["fitness"])
    with pytest.raises(Exception):
        sim.write_pose(60, x=np.zeros(8))
    sim.close()


def test_checkpoint_header_is_checked():
    rob = common.ROBOTS["robot-neuro.json"]
    a = _sim()
    a.upload_genomes(np.random.RandomState(1).uniform(-1, 1, (32, 156)))
    a.reset(float(rob["x"]), float(rob["y"]), 0.0)
    a.run(10)
    snap = a.save_state()
    assert snap[:4] == b"R2PS"
    b = common.make_sim("track_2.png", "robot_omni.json", "linear")     # same byte size per robot, other rays / controller
    b.upload_genomes(np.zeros((32, 18)))
    b.reset(100.0, 100.0, 0.0)
    import r2p2_b200
    with pytest.raises(r2p2_b200.R2P2Error) as ei:
        b.load_state(snap)
    assert "saved for P=32 R=5 G=156" in str(ei.value)
    bad = bytearray(snap); bad[0] ^= 0xFF
    with pytest.raises(ValueError):
        a.load_state(bytes(bad))
    a.load_state(snap)                                                    # and the matching handle takes it
    a.close(); b.close()


def test_ga_checkpoint_resumes_reproducibly():
    """Two generations, save, two more -- against a fresh handle that loads the blob and runs the same two: identical
    populations, fitness vectors and best records."""
    rob = common.ROBOTS["robot-neuro.json"]
    P, G, T = 192, 156, 60

    def fresh():
        sim = _sim()
        sim.upload_genomes(np.zeros((1, G)))
        sim.reset(float(rob["x"]), float(rob["y"]), 0.0, time_budget=20000.0)
        return sim
    s1 = fresh()
    ga = ea.DeviceGA(s1, P, G, seed=21, tournament=3, mutation_rate=0.1, elites=2)
    for _ in range(2):
        ga.step(T)
    blob = ga.save()
    recs = [ga.step(T) for _ in range(2)]
    pop1, fit1 = ga.read()
    s2 = fresh()
    gb = ea.DeviceGA(s2, 4, 3, seed=0)                                    # some other GA: load replaces it
    gb.load(blob)
    gb.P, gb.G = P, G
    assert gb.generation == 2
    recs2 = [gb.step(T) for _ in range(2)]
    pop2, fit2 = gb.read()
    assert np.array_equal(pop1, pop2) and np.array_equal(fit1, fit2)
    assert [(r["best"], r["best_index"]) for r in recs] == [(r["best"], r["best_index"]) for r in recs2]
    import r2p2_b200
    with pytest.raises(r2p2_b200.R2P2Error):
        gb.load(blob[:-8])
    s1.close(); s2.close()


@pytest.mark.parametrize("lpr", [1, 8, 32])
def test_noise_stream_exhaustion_is_flagged(lpr, oracle_port):
    """A replayed stream shorter than the run: the draws past its end are 0.0 as before, and the robot carries
    R2P2_F_NOISE (32) so that the divergence from the reference is visible."""
    rj = "robot_omni.json"
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    P, T = 96, 60
    rs = np.random.RandomState(2)
    genomes = rs.uniform(-1, 1, (P, 18)) * 0.05
    streams = rs.normal(0, 0.5, (P, 40))                                 # 8 rays x 60 ticks would need up to 480
    ref = oracle_port.run(common.load_stage("track_2.png"), ctrl="linear", genomes=genomes, x0=float(rob["x"]), y0=float(rob["y"]), theta0=0.0,
                          T=T, noise=streams, **kw)
    sim = common.make_sim("track_2.png", rj, "linear", lpr=lpr)
    sim.upload_genomes(genomes)
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
    sim.set_noise("replay", stream=streams)
    sim.run(T)
    s = sim.state()
    assert np.array_equal(s["flags"], ref["state"]["flags"]) and (s["flags"] & 32).any()
    assert np.array_equal((s["flags"] & 32) != 0, sim.noise_draws() > 40)
    assert np.array_equal(s["x"], ref["state"]["x"]) and np.array_equal(sim.fitness(), ref["fitness"])
    sim.close()


def test_device_noise_is_keyed_by_the_global_robot_index():
    """Mode "device": an individual draws the same noise whichever shard (GPU) evaluates it."""
    rj = "robot_omni.json"
    rob = common.ROBOTS[rj]
    P, T = 256, 40
    genomes = np.random.RandomState(4).uniform(-1, 1, (P, 18)) * 0.05

    def run(lo, hi):
        sim = common.make_sim("track_2.png", rj, "linear")
        sim.upload_genomes(genomes[lo:hi])
        sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
        sim.set_noise("device", seed=9)
        sim.set_robot_index_offset(lo)
        sim.run(T)
        out = (sim.fitness().copy(), sim.state()["x"].copy())
        sim.close()
        return out
    full = run(0, P)
    parts = [run(lo, lo + 64) for lo in range(0, P, 64)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full[0])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full[1])
    wrong = run(64, 128)
    sim = common.make_sim("track_2.png", rj, "linear")                  # without the offset the shard draws robot 0..63's noise
    sim.upload_genomes(genomes[64:128]); sim.reset(float(rob["x"]), float(rob["y"]), 0.0); sim.set_noise("device", seed=9); sim.run(T)
    assert not np.array_equal(sim.fitness(), wrong[0])
    sim.close()


def test_current_device_is_left_alone():
    """Entry points switch to the handle's device and back (a torch process keeps its own current device)."""
    import ctypes as C
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = C.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        pytest.skip("libcudart not loadable by name")
    n = C.c_int()
    assert rt.cudaGetDeviceCount(C.byref(n)) == 0
    if n.value < 2:
        pytest.skip("needs two GPUs to see a device switch")
    assert rt.cudaSetDevice(0) == 0
    sim = common.make_sim("track_2.png", "robot-neuro.json", "neuro", hidden=[10, 7, 4], device=1)
    sim.upload_genomes(np.zeros((4, 156)))
    sim.reset(230.0, 250.0, 0.0)
    sim.run(3)
    sim.fitness()
    cur = C.c_int(-1)
    assert rt.cudaGetDevice(C.byref(cur)) == 0 and cur.value == 0
    sim.close()
    assert rt.cudaGetDevice(C.byref(cur)) == 0 and cur.value == 0


def test_evaluate_call_equals_the_four_calls_and_pipelines_the_upload():
    """r2p2_evaluate (upload + reset + run + read; populations of three or more waves go up in two pieces, the second one
    while the first already runs, each with its own launch) against the separate calls: identical fitness, state and
    counters -- the 8-lane window kernel (config E's shape, 5 waves) and the one-lane kernel with its piece-wise
    transposition (250 000 linear robots, 4 waves), from pinned and pageable memory."""
    import torch
    import bench
    import r2p2_b200
    from r2p2_b200.evaluator import PopulationEvaluator
    cases = []
    w = bench.WORKLOADS["E"]
    rob = bench.ROBOTS[w["robot"]]
    P = 40000
    cases.append((bench.synthetic_stage(), rob, "neuro", w["hidden"], bench.slice_poses(bench.start_poses(w, P), 0, P), bench.make_genomes(w, 11, 0, P), 30, 2))
    w1 = bench.WORKLOADS["C1"]
    rob1 = bench.ROBOTS[w1["robot"]]
    P1 = 250000
    g1 = np.random.RandomState(3).uniform(-1, 1, (P1, 18)) * 0.05
    cases.append((bench.load_stage(w1["stage"]), rob1, "linear", None, (float(rob1["x"]), float(rob1["y"]), np.random.RandomState(4).uniform(0, 360, P1)), g1, 40, 2))
    for stage, rb, ctrl, hidden, (px, py, pt), genomes, T, min_launches in cases:
        P, G = genomes.shape
        ev = PopulationEvaluator(stage, sonar_range=rb["sonar_range"], radius=rb["radius"], max_speed=rb["max_speed"], sonars=rb["sonars"],
                                 controller=ctrl, hidden_layer=hidden, x=px, y=py, orientation=pt, ticks=T)
        sim = ev.sim
        sim.upload_genomes(genomes)
        sim.reset(px, py, pt)
        sim.run(T)
        ref = (sim.fitness().copy(), sim.state(), sim.counters())
        # pinned host buffers (the overlap case)
        gp = torch.from_numpy(genomes).pin_memory()
        fp = torch.empty(P, dtype=torch.float64).pin_memory()
        n0 = sim.launch_count()
        ev.evaluate_ptr(gp.data_ptr(), P, G, fp.data_ptr())
        launches = sim.launch_count() - n0
        assert launches >= 1 + min_launches, launches            # reset + one episode launch per chunk (+ transpositions)
        assert np.array_equal(fp.numpy(), ref[0])
        st, c = sim.state(), sim.counters()
        for k in st:
            assert np.array_equal(st[k], ref[1][k]), k
        assert c == ref[2]
        # pageable numpy memory, and a second evaluation of other genomes on the same handle
        g2 = genomes[::-1].copy()
        f2 = ev.evaluate(g2)
        assert f2.shape == (P,)                               # (start poses are per robot: only the next evaluation is comparable)
        f3 = ev.evaluate(genomes)
        assert np.array_equal(f3, ref[0])
        assert sim.last_run_ms() > 0
        ev.close()
