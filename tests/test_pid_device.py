"""Sequential_PID_Controller as a device controller (R2P2_CTRL_PID; r2p2/controllers/pid_controller.py:8-262).

The goal list, the PID terms on heading and distance, the obstacle-avoid state machine and the ten ticks of backing off
after a collision run inside the step kernel, one controller per robot, genome = the six gains.

CPU: the oracle's restatement (oracle/r2p2_oracle.c: orc_control_pid) is pinned to tests/golden/host_controllers.npz --
1 500-tick episodes the UNMODIFIED reference ran with controller-PID.json and controller-PID2.json on stage.png (41 / 118
collide() calls, every goal visited, avoid states, backing off): the libm flavour reproduces every pose bit, every sonar
distance, the goal count, the state and the backing counter of every tick.
GPU: the kernels against the port-flavour oracle, bit for bit (pose, flags, counters, controller state), for every lane
mapping, with per-robot gains and goal lists; and against the reference's own episodes (distances and collisions on every
tick identical, poses to 1e-9 -- the device atan2 is within 3 ulp of libm's).
"""
import json
import os

import numpy as np
import pytest

import common

GAINS = ("ap", "ai", "ad", "lp", "li", "ld")


def _episodes():
    g = np.load(os.path.join(common.GOLDEN, "host_controllers.npz"))
    return g, [sp for sp in json.loads(str(g["specs"])) if sp["name"].startswith("pid")]


def _gains(c):
    return np.array([[c[k] for k in GAINS]], dtype=np.float64)


def test_oracle_pid_reproduces_reference_episodes(oracle_libm):
    from oracle.oracle import pid_initial_state
    g, specs = _episodes()
    assert len(specs) == 2
    for sp in specs:
        n, c = sp["name"], common.conf(sp["conf"])
        env = common.load_stage(sp["stage"])
        kw = common.oracle_robot_kwargs(sp["robot"])
        x0, y0, th0 = g[n + "/start"]
        assert (x0, y0) == tuple(c["goal"][0])                # register_robot moves the robot to goal[0]
        out = oracle_libm.run(env, ctrl="pid", genomes=_gains(c), x0=[x0], y0=[y0], theta0=[th0], T=sp["T"], trace=True, pid_goals=c["goal"], **kw)
        assert np.array_equal(out["pose"][:, 0], g[n + "/pose"]), n
        assert np.array_equal(out["dst"][:, 0], g[n + "/dst"]), n
        assert np.array_equal(np.cumsum(out["ncoll"][:, 0]), g[n + "/ncoll"]), n
        assert int(out["pid"]["ngoals"][0]) == int(g[n + "/aux"][-1][0]) == 0
        # tick by tick: goals left, state, backing counter (aux columns 0..2 as the reference's controller held them)
        st, ps = None, pid_initial_state(1, c["goal"])
        for t in range(400):
            o = oracle_libm.run(env, ctrl="pid", genomes=_gains(c), x0=[x0], y0=[y0], theta0=[th0], T=1, state=st, pid_state=ps, **kw)
            st, ps = o["state"], o["pid"]
            aux = g[n + "/aux"][t]
            assert (int(ps["ngoals"][0]), int(ps["state"][0]), int(ps["backing"][0])) == (int(aux[0]), int(aux[1]), int(aux[2])), (n, t)
            assert float(ps["target_angle"][0]) == float(aux[3]), (n, t)
        assert np.array_equal(np.array([st[k][0] for k in ("x", "y", "theta", "speed", "omega")]), g[n + "/pose"][399])


def test_oracle_pid_quirks(oracle_libm):
    """Empty goal list -> (0, 0); a goal pixel off the map -> the reference raises (fault flag); goal-stack overflow faults."""
    env = common.load_stage("stage.png")
    kw = common.oracle_robot_kwargs("robot.json")
    k = _gains(common.conf("controller-PID.json"))
    o = oracle_libm.run(env, ctrl="pid", genomes=k, x0=[600.0], y0=[300.0], theta0=[0.0], T=5, pid_goals=np.zeros((0, 2)), **kw)
    assert (o["state"]["x"][0], o["state"]["y"][0], o["state"]["flags"][0] & 2) == (600.0, 300.0, 0)
    o = oracle_libm.run(env, ctrl="pid", genomes=k, x0=[600.0], y0=[300.0], theta0=[0.0], T=5, pid_goals=[[5000.0, 10.0]], **kw)
    assert o["state"]["flags"][0] & 2 and o["state"]["ticks"][0] == 0


# ---- device ---------------------------------------------------------------------------------------------------------
def _population(P, n_goals, seed):
    """Per-robot gains around the shipped ones and goal lists drawn from free pixels of stage.png (start = goal[0])."""
    env = common.load_stage("stage.png")
    rs = np.random.RandomState(seed)
    base = _gains(common.conf("controller-PID.json"))[0]
    gains = base * rs.uniform(0.5, 1.5, (P, 6))
    free = np.argwhere(env[20:-20, 20:-20] > 60) + 20
    goals = free[rs.randint(0, len(free), (P, n_goals))].astype(np.float64) + rs.uniform(0, 1, (P, n_goals, 2)).round(2)
    goals[: P // 4] = np.floor(goals[: P // 4])               # integer goals like the shipped lists
    return env, gains, goals


def _device(env, gains, goals, T, lpr, ticks_per_launch=None, theta0=0.0):
    sim = common.make_sim(env, "robot.json", "pid", lpr=lpr)
    sim.upload_genomes(gains)
    sim.set_pid(goals)
    per_robot = np.ndim(goals) == 3
    x0 = goals[:, 0, 0] if per_robot else float(goals[0][0])
    y0 = goals[:, 0, 1] if per_robot else float(goals[0][1])
    sim.reset(x0, y0, theta0)
    for _ in range(T // (ticks_per_launch or T)):
        sim.run(ticks_per_launch or T)
    return sim


def _same(sim, ref, what):
    s, c, ps = sim.state(), sim.counters(), sim.pid_state()
    st, rp = ref["state"], ref["pid"]
    for k in ("x", "y", "theta", "speed", "omega", "ncollide", "flags", "ticks"):
        assert np.array_equal(s[k], st[k]), (what, k, int(np.argmax(s[k] != st[k])))
    assert c["sonar_samples"] == int(st["nsamp_sonar"].sum()) and c["collision_samples"] == int(st["nsamp_coll"].sum()), what
    assert np.array_equal(ps["goals_left"], rp["ngoals"]) and np.array_equal(ps["state"], rp["state"]) and np.array_equal(ps["backing"], rp["backing"]), what
    assert np.array_equal(ps["target_angle"], rp["target_angle"]), what
    idx = np.arange(len(rp))
    top = np.maximum(rp["ngoals"] - 1, 0)
    some = rp["ngoals"] > 0
    assert np.array_equal(ps["goal_x"][some], rp["gx"][idx, top][some]) and np.array_equal(ps["goal_y"][some], rp["gy"][idx, top][some]), what
    assert np.isnan(ps["goal_x"][~some]).all()
    assert np.array_equal(sim.fitness(), ref["fitness"]), what


@pytest.mark.gpu
@pytest.mark.parametrize("lpr", [0, 1, 2, 8, 16, 32])
def test_device_pid_equals_oracle(lpr, oracle_port):
    P, T = 384, 500
    env, gains, goals = _population(P, 6, 2024)
    th0 = np.random.RandomState(5).uniform(-180, 180, P)
    kw = common.oracle_robot_kwargs("robot.json")
    ref = oracle_port.run(env, ctrl="pid", genomes=gains, x0=goals[:, 0, 0], y0=goals[:, 0, 1], theta0=th0, T=T, pid_goals=goals, threads=8, **kw)
    sim = _device(env, gains, goals, T, lpr, theta0=th0)
    _same(sim, ref, ("pid", lpr))
    st = ref["state"]
    # the population exercises the whole controller: goals reached, obstacle avoidance (pushed goals), collisions -> backing off
    assert (ref["pid"]["ngoals"] < 6).sum() > P // 4 and (ref["pid"]["state"] == 1).sum() > 0 and (st["ncollide"] > 0).sum() > P // 16
    sim.close()


@pytest.mark.gpu
def test_device_pid_per_tick_api_and_shared_goals(oracle_port):
    """One launch per tick (the Robot.update drop-in) equals one long launch; a goal list shared by every robot."""
    P, T = 64, 300
    c = common.conf("controller-PID2.json")
    env = common.load_stage("stage.png")
    gains = np.repeat(_gains(c), P, axis=0) * np.random.RandomState(3).uniform(0.8, 1.2, (P, 6))
    goals = np.array(c["goal"], dtype=np.float64)
    kw = common.oracle_robot_kwargs("robot.json")
    ref = oracle_port.run(env, ctrl="pid", genomes=gains, x0=goals[0, 0], y0=goals[0, 1], theta0=0.0, T=T, pid_goals=goals, threads=8, **kw)
    for lpr, tpl in ((0, None), (32, 1), (8, 7), (1, 50)):
        sim = _device(env, gains, goals, T - T % (tpl or T), lpr, tpl)
        if T % (tpl or T):
            sim.run(T % tpl)
        _same(sim, ref, ("pid per tick", lpr, tpl))
        sim.close()


@pytest.mark.gpu
def test_device_pid_reproduces_reference_episodes():
    """The reference's own 1 500-tick PID episodes: sonar distances and collide() calls identical on every tick, all goals
    visited, poses within 1e-9 relative (atan2 on the device is within 3 ulp of libm's)."""
    g, specs = _episodes()
    for sp in specs:
        n, c = sp["name"], common.conf(sp["conf"])
        for lpr in (32, 8, 1):
            sim = common.make_sim(sp["stage"], sp["robot"], "pid", lpr=lpr)
            sim.upload_genomes(_gains(c))
            sim.set_pid(c["goal"])
            sim.set_trace(True, sp["T"])
            x0, y0, th0 = g[n + "/start"]
            sim.reset(x0, y0, th0)
            sim.run(sp["T"])
            tr = sim.trace(sp["T"])
            assert np.array_equal(tr["dst"][:, 0], g[n + "/dst"]), (n, lpr)
            assert np.array_equal(np.cumsum(tr["ncoll"][:, 0]), g[n + "/ncoll"]), (n, lpr)
            ref = g[n + "/pose"]
            rel = np.abs(tr["pose"][:, 0] - ref) / np.maximum(1.0, np.abs(ref))
            assert rel.max() < 1e-9, (n, lpr, rel.max())
            assert sim.pid_state()["goals_left"][0] == 0
            sim.close()


@pytest.mark.gpu
def test_device_pid_argument_checks():
    import r2p2_b200
    sim = common.make_sim("stage.png", "robot.json", "pid")
    sim.upload_genomes(np.ones((4, 6)))
    sim.reset(100.0, 100.0, 0.0)
    with pytest.raises(r2p2_b200.R2P2Error, match="r2p2_set_pid"):
        sim.run(1)
    sim.set_pid([[100.0, 100.0], [120.0, 100.0]])
    with pytest.raises(r2p2_b200.R2P2Error, match="r2p2_reset"):
        sim.run(1)                                            # goals are loaded by reset
    sim.reset(100.0, 100.0, 0.0)
    sim.run(3)
    assert (sim.state()["ticks"] == 3).all()
    sim.upload_genomes(np.ones((4, 5)))
    sim.reset(100.0, 100.0, 0.0)
    with pytest.raises(r2p2_b200.R2P2Error, match="six gains"):
        sim.run(1)
    with pytest.raises(ValueError):
        sim.set_pid(np.zeros((3, 2, 2)))                      # per-robot lists for 3 robots, population of 4
    with pytest.raises(r2p2_b200.R2P2Error, match="at most"):
        sim.set_pid(np.zeros((65, 2)))
    sim.close()
