"""Controllers without a device form -- the reference's naive and sequential-goal PID controllers.

tests/golden/host_controllers.npz holds episodes the UNMODIFIED reference ran with its own naive_controller.Naive_Controller
and pid_controller.Sequential_PID_Controller (controller-PID.json and controller-PID2.json) on stage.png: per tick the
distances handed to control(), the command returned, the pose after the tick, collide() calls.

CPU: the controller classes -- this package's, and the reference's own files loaded unmodified through the plugin loader
when an r2p2 checkout is available -- are fed the recorded distances and robot states and must return the recorded
commands bit for bit.  GPU: the closed loop (r2p2_sense -> control() on the host -> r2p2_move) reproduces the recorded
episode, and a list of robots advanced by Simulation in one launch per tick equals the robots advanced one by one.
"""
import json
import os

import numpy as np
import pytest

import common

REF_CONTROLLERS = os.path.join(os.environ.get("R2P2_REFERENCE", "/root/reference"), "r2p2", "controllers")


def _episodes():
    g = np.load(os.path.join(common.GOLDEN, "host_controllers.npz"))
    return g, json.loads(str(g["specs"]))


def _scenario(conf_name):
    ctrl = common.conf(conf_name or "controller-naive.json")
    return {"controllers": [ctrl], "stage": "../res/stage.png"}


def _make_controller(sp, plugin_dir=None):
    from r2p2_b200.controllers.controllers import load_controller, plugin_modules
    if sp["name"] == "naive":
        return load_controller("naive_controller.Naive_Controller", plugin_dir)(_scenario(None))
    cls = load_controller("pid_controller.Sequential_PID_Controller", plugin_dir)
    if plugin_dir:
        # the reference class reads a hard-coded ../conf/controller-PID.json (pid_controller.py:18): give it that file
        import tempfile
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "conf")); os.makedirs(os.path.join(tmp, "r2p2"))
        with open(os.path.join(tmp, "conf", "controller-PID.json"), "w") as fp:
            json.dump(common.conf(sp["conf"]), fp)
        cwd = os.getcwd()
        os.chdir(os.path.join(tmp, "r2p2"))
        try:
            return cls({})
        finally:
            os.chdir(cwd)
    return cls(_scenario(sp["conf"]))


def _robot(ctrl):
    from r2p2_b200.robot import Robot
    rob = common.ROBOTS["robot.json"]
    r = Robot(identifier=0, x=rob["x"], y=rob["y"], orientation=rob["orientation"], vision_range=tuple(rob["sonar_range"]), sensors=rob["sonars"],
              radius=rob["radius"], max_speed=rob["max_speed"], controller=ctrl)
    r.has_noise = False
    return r, rob


def _open_loop(sp, g, plugin_dir=None):
    """Feed the recorded distances / robot states; the commands must be the recorded ones."""
    from r2p2_b200.controllers.controllers import plugin_modules
    from r2p2_b200.logs import python_dst
    n = sp["name"]
    env = common.load_stage(sp["stage"])
    ctrl = _make_controller(sp, plugin_dir)
    r, rob = _robot(ctrl)
    if plugin_dir:
        with plugin_modules(plugin_dir, env=env):
            pass                                              # publishes utils.npdata for is_at_goal
    else:
        ctrl.env = env
    assert (float(r.x), float(r.y), float(r.orientation)) == tuple(g[n + "/start"])
    clamp = rob["sonar_range"][1] + rob["radius"]
    rays = rob["sonars"]
    ncoll_prev = 0
    for t in range(sp["T"]):
        dst = python_dst(g[n + "/dst"][t], clamp)
        ctrl.set_dst(dst)
        ctrl.set_ang([a + r.orientation for a in rays])
        cmd = ctrl.control(dst)
        assert (float(cmd[0]), float(cmd[1])) == tuple(g[n + "/cmd"][t]), (n, t)
        # Robot.control's clamp, then the recorded outcome of the tick
        x, y, th, speed, omega = g[n + "/pose"][t]
        r.x, r.y, r.orientation, r.speed, r.angular_velocity = float(x), float(y), float(th), float(speed), float(omega)
        for _ in range(int(g[n + "/ncoll"][t]) - ncoll_prev):
            ctrl.on_collision((r.x, r.y))
        ncoll_prev = int(g[n + "/ncoll"][t])
        aux = g[n + "/aux"][t]
        if hasattr(ctrl, "goal"):
            assert (len(ctrl.goal), ctrl.state, float(ctrl.backing)) == (int(aux[0]), int(aux[1]), float(aux[2])), (n, t)


def test_controllers_reproduce_recorded_commands():
    g, specs = _episodes()
    for sp in specs:
        _open_loop(sp, g)


@pytest.mark.skipif(not os.path.isdir(REF_CONTROLLERS), reason="no r2p2 checkout (the reference tree does not travel to the GPU box)")
def test_unmodified_reference_plugins_load_and_reproduce_recorded_commands():
    """controllers.load_controller imports naive_controller.py / pid_controller.py from an unmodified r2p2/controllers
    directory (their `from controller import Controller`, `import utils as u`, `import path_planning as pp` resolve to the
    shims / the checkout's own files) and the classes behave as recorded."""
    g, specs = _episodes()
    for sp in specs:
        _open_loop(sp, g, plugin_dir=REF_CONTROLLERS)
    from r2p2_b200.controller import Controller
    from r2p2_b200.controllers.controllers import load_controller
    cls = load_controller("pid_controller.Sequential_PID_Controller", REF_CONTROLLERS)
    assert issubclass(cls, Controller) and cls.__module__ == "pid_controller"


# ---- closed loop on the device ------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("use_reference_files", [False, True])
def test_closed_loop_reproduces_reference_episodes(use_reference_files):
    """Robot.update with a host controller: r2p2_sense -> control() -> r2p2_move.  Distances, commands, poses and
    collide() calls equal the reference's own run on every tick, for the naive controller and both PID parameter sets
    (1 500 ticks each, 41 / 118 collisions, all fourteen goals visited)."""
    if use_reference_files and not os.path.isdir(REF_CONTROLLERS):
        pytest.skip("no r2p2 checkout on this box")
    from r2p2_b200.controllers.controllers import publish_stage
    g, specs = _episodes()
    for sp in specs:
        n = sp["name"]
        env = common.load_stage(sp["stage"])
        ctrl = _make_controller(sp, REF_CONTROLLERS if use_reference_files else None)
        r, rob = _robot(ctrl)
        publish_stage(env, 0.1)
        cmds = []
        orig = ctrl.control

        def spy(dst, orig=orig):
            c = orig(dst)
            cmds.append((float(c[0]), float(c[1])))
            return c
        ctrl.control = spy
        for t in range(sp["T"]):
            r.update(env, 0.1)
            assert [float(d) for d in ctrl.dst] == list(g[n + "/dst"][t]), (n, t)
            if r._group.host_mode:                             # (this package's naive controller has a device form: no host call)
                assert cmds[t] == tuple(g[n + "/cmd"][t]), (n, t)
            assert (r.x, r.y, r.orientation, r.speed, r.angular_velocity) == tuple(g[n + "/pose"][t]), (n, t)
            assert r.collisions == int(g[n + "/ncoll"][t]), (n, t)
        if hasattr(ctrl, "goal"):
            assert len(ctrl.goal) == int(g[n + "/aux"][-1][0])


@pytest.mark.gpu
def test_simulation_batches_robot_lists():
    """utils.update's `for r in robots: r.update(...)` as one launch per group per tick: a list of robots (two groups:
    neuro controllers and PID controllers, different start poses, sensor noise on) advanced by Simulation equals the same
    robots advanced one by one -- poses, controller inputs, collide() callbacks -- and the launch count per tick is per
    group, not per robot."""
    from r2p2_b200.controllers.neurocontroller import Neuro_controller
    from r2p2_b200.controllers.pid_controller import Sequential_PID_Controller
    from r2p2_b200.group import RobotGroup, group_key
    from r2p2_b200.robot import Robot
    env = common.load_stage("stage.png")
    rob = common.ROBOTS["robot.json"]

    def make():
        robots = []
        rs = np.random.RandomState(12)
        for i in range(6):
            w = rs.uniform(-1, 1, 156)
            c = Neuro_controller({"controllers": [{"class": "neurocontroller.Neuro_controller", "weights": list(w), "hidden_layer": [10, 7, 4],
                                                   "activation": "tanh", "time": 20, "evolve": False}]})
            robots.append(Robot(identifier=i, x=rob["x"] + 7 * i, y=rob["y"] + 3 * i, orientation=30 * i, vision_range=tuple(rob["sonar_range"]),
                                sensors=rob["sonars"], radius=rob["radius"], max_speed=rob["max_speed"], controller=c))
        for i in range(3):
            pid = dict(common.conf("controller-PID.json"))
            pid["goal"] = pid["goal"][i:]                      # different start goals
            c = Sequential_PID_Controller({"controllers": [pid]})
            c.env = env
            robots.append(Robot(identifier=10 + i, x=0, y=0, orientation=0, vision_range=tuple(rob["sonar_range"]), sensors=rob["sonars"],
                                radius=rob["radius"], max_speed=rob["max_speed"], controller=c))
        return robots

    T = 150
    one_by_one = make()
    np.random.seed(5)
    for t in range(T):
        for r in one_by_one:                                   # the reference's loop
            r.update(env, 0.1)
    batched = make()
    buckets = {}
    for r in batched:
        buckets.setdefault(group_key(r), []).append(r)
    groups = [RobotGroup(rs, env) for rs in buckets.values()]
    assert sorted(len(g.robots) for g in groups) == [3, 6]
    np.random.seed(5)
    l0 = [g.sim.launch_count() for g in groups]
    for t in range(T):
        for g in groups:
            g.update(0.1)
    launches = [g.sim.launch_count() - a for g, a in zip(groups, l0)]
    assert all(n <= 3 * T + 8 for n in launches), launches      # sense + move (+ the noise pre-pass), whatever the group size
    moved = 0
    for a, b in zip(one_by_one, batched):
        # numpy's global noise stream is consumed robot after robot; the groups here are contiguous in the robot list, so
        # the batched pass (group by group) draws in the reference's order and the runs are identical
        assert (a.x, a.y, a.orientation, a.speed, a.angular_velocity) == (b.x, b.y, b.orientation, b.speed, b.angular_velocity), a.identifier
        assert (a.collisions, a.flags, a.last_pos) == (b.collisions, b.flags, b.last_pos), a.identifier
        assert a.controller.dst == b.controller.dst, a.identifier
        moved += int((a.x, a.y) != (rob["x"], rob["y"]))
    assert moved == len(batched)
    for g in groups:
        g.close()
