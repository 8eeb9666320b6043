"""The r2p2-facing Python surface on a real B200: Robot.update, scenario loading, evolution callback."""
import os

import numpy as np
import pytest

import common
from test_host_logic import write_conf

pytestmark = pytest.mark.gpu


def test_robot_update_reproduces_reference_episode():
    """`for t: robot.update(env, 0.1)` with the linear controller == the reference's KAT3 run, tick by tick, bit for bit
    (pose, sonar distances handed to the controller, collide() calls, fitness)."""
    from r2p2_b200.controllers.inspyred import Inspyred_controller
    from r2p2_b200.robot import Robot
    e, specs = common.episode_specs()
    sp = next(s for s in specs if s["name"] == "kat3_linear")
    rob = common.ROBOTS[sp["robot"]]
    env = common.load_stage(sp["stage"])
    ctrl = Inspyred_controller({"controllers": [{"class": "inspyred.Inspyred_controller", "weights": list(e["kat3_linear/genome"]), "time": 20, "evolve": False}]})
    r = Robot(identifier=0, x=rob["x"], y=rob["y"], orientation=rob["orientation"], vision_range=tuple(rob["sonar_range"]), sensors=rob["sonars"],
              radius=rob["radius"], max_speed=rob["max_speed"], controller=ctrl)
    assert r.has_noise is True                     # the reference's default (robot.py:96); golden runs switch it off
    r.has_noise = False
    gp, gd, gc = e["kat3_linear/pose"], e["kat3_linear/dst"], e["kat3_linear/ncoll"]
    for t in range(120):
        r.update(env, 0.1, True)
        assert (r.x, r.y, r.orientation, r.speed, r.angular_velocity) == tuple(gp[t]), t
        assert ctrl.dst == list(gd[t]), t
        assert r.collisions == gc[t]
    assert ctrl.fitness() == e["kat3_linear/fitness"][119]


def test_robot_facade_neuro_and_check_sensors():
    from r2p2_b200.controllers.neurocontroller import Neuro_controller
    from r2p2_b200.robot import Robot
    e, specs = common.episode_specs()
    rob = common.ROBOTS["robot-neuro.json"]
    env = common.load_stage("track_2.png")
    ctrl = Neuro_controller({"controllers": [{"class": "neurocontroller.Neuro_controller", "weights": list(e["kat2_neuro/genome"]),
                                              "hidden_layer": [10, 7, 4], "activation": "tanh", "time": 20, "evolve": False}]})
    r = Robot(identifier=0, x=rob["x"], y=rob["y"], orientation=0, vision_range=tuple(rob["sonar_range"]), sensors=rob["sonars"],
              radius=rob["radius"], max_speed=rob["max_speed"], controller=ctrl)
    r.has_noise = False
    col, dst, ang = r.check_sensors(env)
    assert dst == list(e["kat2_neuro/dst"][0]) and ang == [0.0, 45.0, 70.0, 290.0, 315.0]
    assert (r.x, r.y) == (rob["x"], rob["y"])
    gp = e["kat2_neuro/pose"]
    for t in range(100):
        r.update(env, 0.1)
        assert abs(r.x - gp[t, 0]) < 1e-9 and abs(r.y - gp[t, 1]) < 1e-9 and abs(r.orientation - gp[t, 2]) < 1e-9
        assert np.array_equal(np.rint(ctrl.dst), np.rint(e["kat2_neuro/dst"][t]))


def test_robot_facade_with_noise_replays_numpy_stream():
    """has_noise = True: the facade feeds numpy's global legacy stream to the device, so a seeded run equals the
    reference's seeded run draw for draw (golden: noise_linear_omni, np.random.seed(42))."""
    import json
    from r2p2_b200.controllers.inspyred import Inspyred_controller
    from r2p2_b200.robot import Robot
    e = common.golden("episodes_noise")
    sp = next(s for s in json.loads(str(e["specs"])) if s["name"] == "noise_linear_omni")
    rob = common.ROBOTS[sp["robot"]]
    env = common.load_stage(sp["stage"])
    ctrl = Inspyred_controller({"controllers": [{"class": "inspyred.Inspyred_controller", "weights": list(e["noise_linear_omni/genome"]), "time": 20, "evolve": False}]})
    r = Robot(identifier=0, x=rob["x"], y=rob["y"], orientation=0, vision_range=tuple(rob["sonar_range"]), sensors=rob["sonars"],
              radius=rob["radius"], max_speed=rob["max_speed"], controller=ctrl)
    np.random.seed(sp["seed"])
    for t in range(80):
        r.update(env, 0.1)
        assert (r.x, r.y, r.orientation) == tuple(e["noise_linear_omni/pose"][t, :3]), t
        assert ctrl.dst == list(e["noise_linear_omni/dst"][t]), t


def test_scenario_files_and_evolution_callback(tmp_path, oracle_port):
    """Unchanged scenario / robot / controller JSON files -> Simulation and the batched evolution callback; the callback's
    fitness list equals the port oracle's for the same genomes (evolve semantics, 30 Hz time units of SURVEY H17)."""
    from r2p2_b200 import evolution
    from r2p2_b200.simulation import load_simulation
    conf = write_conf(str(tmp_path))
    sim = load_simulation(os.path.join(conf, "scenario-neuro-simple.json"))
    assert len(sim.robots) == 1 and sim.env.shape == (460, 321)
    sim.robots[0].controller.set_network_params(list(np.random.RandomState(5).uniform(-1, 1, 7)))
    sim.run(20, delta=0.1)
    assert sim.robots[0].x != 230

    ev = evolution.configure(os.path.join(conf, "scenario-neuro.json"), ticks=300, history_path=str(tmp_path / "history.log"),
                             fitness_path=str(tmp_path / "fitness.txt"))
    genomes = np.random.RandomState(3).uniform(-1, 1, (64, 156))
    fit = evolution.evaluator(genomes.tolist(), {})
    rob, kw = common.ROBOTS["robot-neuro.json"], common.oracle_robot_kwargs("robot-neuro.json")
    ref = oracle_port.run(common.load_stage("track_2.png"), ctrl="neuro", genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0, T=300,
                          delta=1.0 / 30.0, delta_ctrl=1000.0 / 30.0, time_budget=20000.0, evolve=True, layers=[5, 10, 7, 4, 2], **kw)
    assert fit == ref["fitness"].tolist()
    # the reference's side files: history records for every candidate at least as good as the best so far, the mailbox file
    recs = open(tmp_path / "history.log").read().split("\n\n")[1:]
    best, want = 0, []
    for g, f in zip(genomes.tolist(), fit):
        if f >= best:
            want.append("[" + str(f) + "]: " + str(g).replace(" ", ""))
            best = f
    assert recs == want and len(recs) >= 1
    assert float(open(tmp_path / "fitness.txt").read()) == fit[-1]
    assert evolution.evaluate_ann(genomes[7].tolist(), {}) == fit[7]
    assert evolution.evaluator([[0.0] * 3], {}) == [0]            # evolution.py:49-52: a failing candidate scores 0
    ev.close()


from r2p2_b200.controller import Controller as _Controller  # noqa: E402


class _HostLinear(_Controller):
    """A controller the library knows nothing about (no device form): the arithmetic of the reference's
    Inspyred_controller.control (controllers/inspyred.py:60-100) in plain Python."""

    def __init__(self, chromosome):
        super().__init__("HOST", None)
        n = len(chromosome) // 2 - 1
        self.lin, self.angc = list(chromosome[:n]), list(chromosome[n:2 * n])
        self.calls = 0

    def control(self, dst):
        self.calls += 1
        v = 0
        w = 0
        for coef, d in zip(self.lin, dst):
            v += coef * d
        for coef, d in zip(self.angc, dst):
            w += coef * d
        return v, w


@pytest.mark.gpu
def test_host_controller_bridge_reproduces_reference_episode():
    """Robot.update with a controller that has no device form: r2p2_sense -> controller.control on the host ->
    r2p2_move.  Same poses as the unmodified reference's run of the same linear controller (golden kat3_linear), bit
    for bit on every tick, and the same robot logs."""
    import json, os, tempfile
    from r2p2_b200.robot import Robot
    e, specs = common.episode_specs()
    sp = next(s for s in specs if s["name"] == "kat3_linear")
    rj = common.ROBOTS[sp["robot"]]
    ctrl = _HostLinear([float(c) for c in e["kat3_linear/genome"]])
    tmp = tempfile.mkdtemp()
    r = Robot(identifier=0, x=rj["x"], y=rj["y"], orientation=rj["orientation"], max_speed=rj["max_speed"], sensors=rj["sonars"],
              vision_range=rj["sonar_range"], radius=rj["radius"], controller=ctrl, log_dir=tmp)
    r.has_noise = False
    env = common.load_stage(sp["stage"])
    T = 120
    for t in range(T):
        r.update(env, sp["delta"], True)
        gp = e["kat3_linear/pose"][t]
        assert (r.x, r.y, r.orientation, r.speed, r.angular_velocity) == tuple(gp), t
        assert [float(d) for d in ctrl.dst] == list(e["kat3_linear/dst"][t]), t
    assert ctrl.calls == T
    r._logs.close()
    with open(os.path.join(common.GOLDEN, "logs.json")) as fp:
        g = json.load(fp)["linear_omni"]                # same robot, same chromosome, 40 ticks with write_stats
    assert open(os.path.join(tmp, "robot_0_sensors.log")).read().startswith(g["sensors"])
    assert open(os.path.join(tmp, "robot_0.log")).read().startswith(g["log"])


@pytest.mark.gpu
def test_sense_move_equal_device_controller(oracle_port):
    """Population level: r2p2_sense + host linear algebra + r2p2_move per tick == the device's own linear controller."""
    rj = "robot_omni.json"
    rob = common.ROBOTS[rj]
    P, T = 200, 60
    genomes = np.random.RandomState(11).uniform(-1, 1, (P, 18)) * 0.06
    ref = common.make_sim("track_2.png", rj, "linear")
    ref.upload_genomes(genomes); ref.reset(rob["x"], rob["y"], 0.0); ref.run(T)
    sim = common.make_sim("track_2.png", rj, "const")
    sim.upload_genomes(np.zeros((P, 2))); sim.reset(rob["x"], rob["y"], 0.0)
    n = 18 // 2 - 1
    for _ in range(T):
        dst = sim.sense()
        v = np.zeros(P); w = np.zeros(P)
        for i in range(n):                              # the reference's accumulation order (inspyred.py:88-93)
            v = v + genomes[:, i] * dst[:, i]
            w = w + genomes[:, n + i] * dst[:, i]
        sim.move(np.stack([v, w], axis=1), 0.1)
    a, b = sim.state(), ref.state()
    for k in ("x", "y", "theta", "speed", "omega", "ncollide", "flags", "ticks"):
        assert np.array_equal(a[k], b[k]), k
    sim.close(); ref.close()
