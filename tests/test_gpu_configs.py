"""BASELINE.json configs D and E at their own shapes, and config A's literal combination, against the oracle.

  D  scenario-track: res/track.png, conf/robot-track.json (5 rays, range [0.5, 40] -> limit 45), MLP 5-10-7-4-2, four start
     headings (0 / 90 / 180 / 270)
  E  synthetic sweep: the 4096 x 4096 stage bench.py generates (wpr = 128 words, 16 MB distance field, 2.3 MB hot plane),
     32 list-form rays, MLP 32-16-8-2, start poses in free space, the lane mapping the library picks by itself
  A  scenario-neuro-simple as shipped: track_2.png + robot-neuro.json + 5-1-2 (controller-neuro-simple.json)

State, flags, fitness and the sonar / collision sample counters are compared bit for bit with the port-flavour oracle
(same r2p2_math.h as the kernels) and fitness with the libm-flavour oracle (the reference's own libm calls; differs from
the device only through tanh / atan2 at the ulp level) to 1e-9 relative -- the north-star tolerance is 1e-5.
"""
import numpy as np
import pytest

import bench
import common

pytestmark = pytest.mark.gpu


def _check(sim, ref, libm, what):
    s, f, c = sim.state(), sim.fitness(), sim.counters()
    st = ref["state"]
    for k in ("x", "y", "theta", "speed", "omega", "ncollide", "flags", "ticks"):
        assert np.array_equal(s[k], st[k]), (what, k)
    assert np.array_equal(f, ref["fitness"]), what
    assert c["sonar_samples"] == int(st["nsamp_sonar"].sum()), what
    assert c["collision_samples"] == int(st["nsamp_coll"].sum()), what
    assert c["robot_steps"] == int(st["ticks"].sum()), what
    ok = (st["flags"] & 2) == 0
    rel = np.abs(f[ok] - libm["fitness"][ok]) / np.maximum(1.0, np.abs(libm["fitness"][ok]))
    # The two flavours differ by <= 3 ulp in tanh / atan2, hence at the 1e-16 level in the headings: knife-edge samples
    #  may be counted differently without any effect on the distances.  A robot whose collision history differs
    # (a knife-edge decided a HIT the other way) is allowed but must be rare; measured: none in these populations.
    same_history = libm["state"]["ncollide"][ok] == st["ncollide"][ok]
    assert same_history.mean() > 0.99, what
    assert rel[same_history].max() < 1e-9, (what, rel[same_history].max())


@pytest.mark.parametrize("lpr", [0, 1, 8, 32])
def test_config_D_track(lpr, oracle_port, oracle_libm):
    """scenario-track shape: 4 start headings x genomes, 300 ticks, every mapping."""
    rj = "robot-track.json"
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    P, T = 512, 300
    genomes = np.random.RandomState(31).uniform(-1, 1, (P, 156))
    th0 = np.tile([0.0, 90.0, 180.0, 270.0], P // 4)
    env = common.load_stage("track.png")
    args = dict(ctrl="neuro", genomes=genomes, x0=float(rob["x"]), y0=float(rob["y"]), theta0=th0, T=T, layers=[5, 10, 7, 4, 2], threads=8, **kw)
    ref, libm = oracle_port.run(env, **args), oracle_libm.run(env, **args)
    sim = common.make_sim("track.png", rj, "neuro", hidden=[10, 7, 4], lpr=lpr)
    sim.upload_genomes(genomes)
    sim.reset(float(rob["x"]), float(rob["y"]), th0)
    sim.run(T)
    _check(sim, ref, libm, ("D", lpr))
    assert int(ref["state"]["ncollide"].astype(bool).sum()) > P // 8       # the population does run into walls
    sim.close()


@pytest.fixture(scope="module")
def sweep_stage():
    return bench.synthetic_stage()


@pytest.mark.parametrize("lpr", [0, 2, 4, 8, 16])
def test_config_E_sweep(lpr, sweep_stage, oracle_port, oracle_libm):
    """Synthetic sweep shape on the full 4096 x 4096 stage: 320 robots x 240 ticks, the automatic mapping (lpr 0: eight
    lanes per robot, windows of the hot plane in shared memory) and the other warp-synchronous mappings."""
    import r2p2_b200
    w = bench.WORKLOADS["E"]
    rob = bench.ROBOTS[w["robot"]]
    P, T = 320, 240
    x0, y0, th0 = bench.slice_poses(bench.start_poses(w, 4096), 0, P)
    genomes = bench.make_genomes(w, 99, 0, P)
    # a few robots start right at the map's edge region so that rays, rings and windows reach into the margin
    x0[:8] = [9.5, 12.0, 4085.5, 4086.0, 200.0, 300.0, 9.0, 4086.9]
    y0[:8] = [9.5, 4086.0, 12.0, 4085.0, 9.2, 4086.5, 2000.0, 2000.0]
    args = dict(rays=rob["sonars"], vr=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], ctrl="neuro", genomes=genomes,
                x0=x0, y0=y0, theta0=th0, T=T, layers=[32, 16, 8, 2], threads=8)
    env = sweep_stage.astype(np.int32)
    ref, libm = oracle_port.run(env, **args), oracle_libm.run(env, **args)
    sim = r2p2_b200.BatchSim(0)
    sim.set_stage(sweep_stage)
    sim.set_robot(sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], sonars=rob["sonars"])
    sim.set_controller("neuro", hidden_layer=[16, 8])
    sim.set_mapping(lpr, -1)
    sim.upload_genomes(genomes)
    sim.reset(x0, y0, th0)
    sim.run(T)
    _check(sim, ref, libm, ("E", lpr))
    assert int(ref["state"]["ncollide"].astype(bool).sum()) > P // 8
    sim.close()


def test_config_E_large_population_mapping_invariance(sweep_stage):
    """More robots than fit at once (the regime of the full run): 40 000 robots x 40 ticks, the automatic mapping against
    one lane per robot (a different kernel, the global hot plane instead of windows): bit-identical state and counters."""
    import r2p2_b200
    w = bench.WORKLOADS["E"]
    rob = bench.ROBOTS[w["robot"]]
    P, T = 40000, 40
    x0, y0, th0 = bench.slice_poses(bench.start_poses(w, P), 0, P)
    genomes = bench.make_genomes(w, 7, 0, P)
    res = []
    for lpr in (0, 1):
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(sweep_stage)
        sim.set_robot(sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], sonars=rob["sonars"])
        sim.set_controller("neuro", hidden_layer=[16, 8])
        sim.set_mapping(lpr, -1)
        sim.upload_genomes(genomes)
        sim.reset(x0, y0, th0)
        sim.run(T)
        res.append((sim.state(), sim.fitness(), sim.counters()))
        sim.close()
    (s0, f0, c0), (s1, f1, c1) = res
    for k in s0:
        assert np.array_equal(s0[k], s1[k]), k
    assert np.array_equal(f0, f1) and c0 == c1
    assert c0["robot_steps"] == P * T


@pytest.mark.parametrize("evolve,noise", [(False, "off"), (True, "device")])
def test_time_slicing_of_multi_wave_populations(evolve, noise, sweep_stage):
    """A population that runs in several waves on the warp-synchronous kernel is time-sliced (a warp's work item is one
    slice of the ticks of one batch of robots; the next slice of a batch waits for the previous one's write-back):
    24 000 32-ray robots x 130 ticks = 2.5 waves -> five slices of 26 ticks.  One lane per robot (another kernel, never
    sliced) must give the same bits -- state, flags, fitness, sample counters, noise draws -- with robots that stop early
    (evolve: a collision ends the episode) and with the device noise generator."""
    import r2p2_b200
    w = bench.WORKLOADS["E"]
    rob = bench.ROBOTS[w["robot"]]
    P, T = 24000, 130
    x0, y0, th0 = bench.slice_poses(bench.start_poses(w, P), 0, P)
    genomes = bench.make_genomes(w, 11, 0, P)
    res = []
    for lpr in (8, 1):
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(sweep_stage)
        sim.set_robot(sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], sonars=rob["sonars"])
        sim.set_controller("neuro", hidden_layer=[16, 8])
        sim.set_noise(noise, seed=77)
        sim.set_mapping(lpr, -1)
        sim.upload_genomes(genomes)
        sim.reset(x0, y0, th0, time_budget=9000.0)
        sim.run(T, delta=0.1, delta_ctrl=100.0, evolve=evolve)        # evolve: the timer runs out after 90 control calls
        res.append((sim.state(), sim.fitness(), sim.counters(), sim.noise_draws()))
        sim.close()
    (s0, f0, c0, n0), (s1, f1, c1, n1) = res
    for k in s0:
        assert np.array_equal(s0[k], s1[k]), k
    assert np.array_equal(f0, f1) and c0 == c1 and np.array_equal(n0, n1)
    if evolve:
        assert 0 < c0["robot_steps"] < P * T and int((s0["ticks"] < 90).sum()) > P // 20     # collisions ended episodes early
    else:
        assert c0["robot_steps"] == P * T


def test_config_A_as_shipped(oracle_port, oracle_libm):
    """scenario-neuro-simple.json as shipped: track_2.png + robot-neuro.json + controller-neuro-simple.json (hidden [1], tanh,
    7 weights), single robots, 600 and 1000 ticks."""
    scen = common.conf("scenario-neuro-simple.json")
    ctrl = common.conf(scen["controller"].split("/")[-1])
    rj = scen["robot"].split("/")[-1]
    stage = scen["stage"].split("/")[-1]
    assert (stage, rj, ctrl["hidden_layer"], ctrl["activation"], len(ctrl["weights"])) == ("track_2.png", "robot-neuro.json", [1], "tanh", 7)
    rob, kw = common.ROBOTS[rj], common.oracle_robot_kwargs(rj)
    genomes = np.random.RandomState(65).uniform(-1, 1, (16, 7)) * 3.0
    env = common.load_stage(stage)
    for T in (600, 1000):
        args = dict(ctrl="neuro", genomes=genomes, x0=float(rob["x"]), y0=float(rob["y"]), theta0=float(rob["orientation"]), T=T,
                    layers=[5] + ctrl["hidden_layer"] + [2], activation=ctrl["activation"], threads=4, **kw)
        ref, libm = oracle_port.run(env, **args), oracle_libm.run(env, **args)
        for lpr in (0, 1, 32):
            sim = common.make_sim(stage, rj, "neuro", hidden=ctrl["hidden_layer"], activation=ctrl["activation"], lpr=lpr)
            sim.upload_genomes(genomes)
            sim.reset(float(rob["x"]), float(rob["y"]), float(rob["orientation"]))
            sim.run(T)
            _check(sim, ref, libm, ("A", T, lpr))
            sim.close()
