"""The oracle (oracle/r2p2_oracle.c) against golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle; the reference ships no tests of its own."""
import numpy as np
import pytest

import common


@pytest.mark.parametrize("flavour", ["libm", "port"])
def test_rays_bit_exact(flavour, oracle_libm, oracle_port):
    """utils.search_edge_in_angle_with_limit (reference r2p2/utils.py:420-432): hit flag, hit point bits and the
    number of pixels tested, for 8118 rays (integer/float origins, integer/float limits, rays leaving the map)."""
    O = oracle_libm if flavour == "libm" else oracle_port
    r = common.golden("rays")
    names = list(r["stages"])
    envs = {n: common.load_stage(n) for n in names}
    bad = 0
    for i in range(len(r["hit"])):
        st, hx, hy, k, ns = O.ray(envs[names[r["stage"][i]]], r["ox"][i], r["oy"][i], r["angle"][i], r["limit"][i])
        bad += not (st == r["hit"][i] and hx == r["hx"][i] and hy == r["hy"][i] and ns == r["nsamp"][i])
    assert bad == 0


def test_kat1_known_answers(oracle_libm):
    """SURVEY.md 8c KAT1, typed in by hand from the survey (independent of the fixture files)."""
    import math
    env = common.load_stage("track_2.png")
    exp = {70: ((238, 272), 24.000000000000416), 290: ((237, 230), 20.999999999999943), 315: ((249, 230), 27.00000000000022)}
    for a in (0, 45, 70, 290, 315):
        st, hx, hy, k, ns = oracle_libm.ray(env, 230, 250, math.radians(a), 30)
        if a in exp:
            assert st == 1 and (int(hx), int(hy)) == exp[a][0]
            assert float(np.linalg.norm((hx - 230, hy - 250))) == exp[a][1]
        else:
            assert st == 0


@pytest.mark.parametrize("flavour", ["libm", "port"])
def test_single_ticks_bit_exact(flavour, oracle_libm, oracle_port):
    """One Robot.update (reference r2p2/robot.py:389-402) from 4200 random states with a scripted command: new
    pose, speed, sonar distances, collide() count, last_pos and the exact count of env.item() calls."""
    O = oracle_libm if flavour == "libm" else oracle_port
    t = common.golden("ticks")
    sn, rn = list(t["stages"]), list(t["robots"])
    envs = {n: common.load_stage(n) for n in sn}
    bad = 0
    for i in range(len(t["x"])):
        kw = common.oracle_robot_kwargs(rn[t["robot"][i]])
        out = O.run(envs[sn[t["stage"][i]]], ctrl="const", genomes=np.array([[t["v"][i], t["w"][i]]]), x0=t["x"][i], y0=t["y"][i],
                    theta0=t["theta"][i], T=1, delta=t["delta"][i], trace=True, **kw)
        s = out["state"][0]
        R = len(kw["rays"])
        ok = (s["x"] == t["nx"][i] and s["y"] == t["ny"][i] and s["theta"] == t["ntheta"][i] and s["speed"] == t["nspeed"][i]
              and s["ncollide"] == t["ncollide"][i] and s["last_x"] == t["nlast_x"][i] and s["last_y"] == t["nlast_y"][i]
              and s["npixel"] == t["nsamp"][i] and np.array_equal(out["dst"][0, 0], t["dst"][i][:R]))
        bad += not ok
    assert bad == 0
    assert int((t["ncollide"] > 0).sum()) > 200     # the fixture does exercise collisions


@pytest.mark.parametrize("flavour", ["libm", "port"])
def test_episodes(flavour, oracle_libm, oracle_port):
    """Free-running episodes.  Linear (inspyred.py:60-100) and constant-command runs must be bit-identical on every
    tick (pose, distances, collisions, fitness).  The MLP goes through numpy's BLAS summation order and SIMD tanh in
    the reference, which the oracle restates only to ulp level: there hit steps and collisions must be identical and
    poses / fitness within 1e-9 relative (north-star tolerance: 1e-5)."""
    O = oracle_libm if flavour == "libm" else oracle_port
    e, specs = common.episode_specs()
    for sp in specs:
        n = sp["name"]
        rob = common.ROBOTS[sp["robot"]]
        out = O.run(common.load_stage(sp["stage"]), ctrl=sp["kind"], genomes=e[n + "/genome"][None, :], x0=rob["x"], y0=rob["y"],
                    theta0=0.0, T=sp["T"], delta=sp["delta"], layers=sp["layers"], trace=True, **common.oracle_robot_kwargs(sp["robot"]))
        pose, gp = out["pose"][:, 0], e[n + "/pose"]
        assert np.array_equal(np.cumsum(out["ncoll"][:, 0]), e[n + "/ncoll"]), n
        if sp["kind"] == "neuro":
            assert np.array_equal(np.rint(out["dst"][:, 0]), np.rint(e[n + "/dst"])), n      # same hit step on every ray
            assert np.max(np.abs(pose[:, :3] - gp[:, :3]) / np.maximum(1, np.abs(gp[:, :3]))) < 1e-9, n
            gf = e[n + "/fitness"][-1]
            assert abs(out["fitness"][0] - gf) <= 1e-9 * abs(gf), n
        else:
            assert np.array_equal(pose, gp), n
            assert np.array_equal(out["dst"][:, 0], e[n + "/dst"]), n
            if sp["kind"] == "linear":
                assert out["fitness"][0] == e[n + "/fitness"][-1], n


def test_kat_values_from_survey(oracle_libm):
    """KAT2/KAT3 numbers as printed in SURVEY.md 8c (typed by hand)."""
    rob = common.ROBOTS["robot_omni.json"]
    out = oracle_libm.run(common.load_stage("track_2.png"), ctrl="linear", genomes=(np.random.RandomState(7).uniform(-1, 1, 18) * 0.05)[None, :],
                          x0=rob["x"], y0=rob["y"], theta0=0.0, T=300, trace=True, **common.oracle_robot_kwargs("robot_omni.json"))
    assert out["pose"][0, 0, 0] == 193.08076961494612 and out["pose"][0, 0, 1] == 248.99986179682114
    assert out["pose"][299, 0, 0] == 231.961294130248 and out["pose"][299, 0, 2] == -16.249508036780878
    assert out["fitness"][0] == 82.48484273315702
    assert int(np.argmax(np.cumsum(out["ncoll"][:, 0]) > 0)) == 224 and int(out["ncoll"].sum()) == 76
    rob = common.ROBOTS["robot-neuro.json"]
    w = np.random.RandomState(1).uniform(-1, 1, 156)
    out = oracle_libm.run(common.load_stage("track_2.png"), ctrl="neuro", genomes=w[None, :], x0=rob["x"], y0=rob["y"], theta0=0.0, T=1000,
                          layers=[5, 10, 7, 4, 2], trace=True, **common.oracle_robot_kwargs("robot-neuro.json"))
    assert abs(out["fitness"][0] - 61.26121583287353) < 1e-9
    assert int(np.argmax(np.cumsum(out["ncoll"][:, 0]) > 0)) == 761 and int(out["ncoll"].sum()) == 239


def test_evolve_semantics(oracle_libm):
    """neurocontroller.py:72-85: with evolve the episode ends at the first control() after a collision (or when the
    timer runs out) and the fitness is the one the reference would have written to fitness.txt there."""
    rob = common.ROBOTS["robot-neuro.json"]
    w = np.random.RandomState(1).uniform(-1, 1, 156)
    kw = dict(ctrl="neuro", genomes=w[None, :], x0=rob["x"], y0=rob["y"], theta0=0.0, layers=[5, 10, 7, 4, 2], **common.oracle_robot_kwargs("robot-neuro.json"))
    env = common.load_stage("track_2.png")
    free = oracle_libm.run(env, T=1000, **kw)
    ev = oracle_libm.run(env, T=1000, evolve=True, delta_ctrl=33.0, time_budget=1e9, **kw)
    assert ev["state"]["ticks"][0] == 762 and ev["state"]["flags"][0] & 4          # first collide at tick 761 (KAT2)
    assert ev["fitness"][0] == free["fitness"][0]                                    # pose frozen afterwards (SURVEY KAT2)
    tm = oracle_libm.run(env, T=1000, evolve=True, delta_ctrl=100.0, time_budget=20000.0, **kw)   # timer: 200 control() calls
    assert tm["state"]["ticks"][0] == 199 and tm["state"]["flags"][0] & 4


@pytest.mark.parametrize("flavour", ["libm", "port"])
def test_noisy_episodes(flavour, oracle_libm, oracle_port):
    """has_noise = True (reference default, robot.py:96,308-310): seeded reference runs; the oracle consumes the same
    np.random.normal(0, 0.5) stream one draw per unclamped ray in ray order.  Linear / constant-command runs are
    bit-identical (pose, noisy distances, collide counts, number of draws); the MLP run to 1e-9."""
    import json
    O = oracle_libm if flavour == "libm" else oracle_port
    e = common.golden("episodes_noise")
    for sp in json.loads(str(e["specs"])):
        n, rob, kw = sp["name"], common.ROBOTS[sp["robot"]], common.oracle_robot_kwargs(sp["robot"])
        np.random.seed(sp["seed"])
        stream = np.random.normal(0, 0.5, size=sp["T"] * len(kw["rays"]) + 8)
        out = O.run(common.load_stage(sp["stage"]), ctrl=sp["kind"], genomes=e[n + "/genome"][None, :], x0=rob["x"], y0=rob["y"], theta0=0.0,
                    T=sp["T"], layers=sp["layers"], noise=stream[None, :], trace=True, **kw)
        assert int(out["state"]["ndraws"][0]) == int(e[n + "/ndraws"]), n
        assert np.array_equal(np.cumsum(out["ncoll"][:, 0]), e[n + "/ncoll"]), n
        if sp["kind"] == "neuro":
            assert np.max(np.abs(out["pose"][:, 0, :3] - e[n + "/pose"][:, :3])) < 1e-9 * 500, n
        else:
            assert np.array_equal(out["pose"][:, 0], e[n + "/pose"]), n
            assert np.array_equal(out["dst"][:, 0], e[n + "/dst"]), n


@pytest.mark.parametrize("flavour", ["libm", "port"])
def test_activation_episodes(flavour, oracle_libm, oracle_port):
    """The other activations sklearn's MLP accepts (relu, logistic, identity) and other layer shapes, recorded from the
    reference (tests/golden/episodes_act.npz): same hit step on every ray of every tick, same collisions, poses and
    fitness to 1e-9 (BLAS summation order / SIMD exp of the reference are only restated to the ulp)."""
    import json
    O = oracle_libm if flavour == "libm" else oracle_port
    e = common.golden("episodes_act")
    for sp in json.loads(str(e["specs"])):
        n = sp["name"]
        rob = common.ROBOTS[sp["robot"]]
        out = O.run(common.load_stage(sp["stage"]), ctrl="neuro", genomes=e[n + "/genome"][None, :], x0=rob["x"], y0=rob["y"], theta0=0.0,
                    T=sp["T"], layers=sp["layers"], activation=sp["activation"], trace=True, **common.oracle_robot_kwargs(sp["robot"]))
        gp = e[n + "/pose"]
        assert np.array_equal(np.cumsum(out["ncoll"][:, 0]), e[n + "/ncoll"]), n
        assert np.array_equal(np.rint(out["dst"][:, 0]), np.rint(e[n + "/dst"])), n
        assert np.max(np.abs(out["pose"][:, 0, :3] - gp[:, :3]) / np.maximum(1, np.abs(gp[:, :3]))) < 1e-9, n
        gf = float(e[n + "/fitness"])
        assert abs(out["fitness"][0] - gf) <= 1e-9 * max(1.0, abs(gf)), n
