#!/usr/bin/env python3
"""The Robot facade measured: `for r in robots: r.update(...)` (r2p2/utils.py:226-227) through RobotGroup -- one launch and
one read-back per group and tick, Python objects refreshed for every robot -- against the same robots advanced one by one
(one launch per robot and tick, what round 1's Simulation.update did).  Neuro controllers on stage.png, noise off."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from r2p2_b200.controllers.neurocontroller import Neuro_controller
from r2p2_b200.group import RobotGroup
from r2p2_b200.robot import Robot

ap = argparse.ArgumentParser()
ap.add_argument("--ticks", type=int, default=100)
a = ap.parse_args()
env = common.load_stage("stage.png")
rob = common.ROBOTS["robot-neuro.json"]


def make(n):
    rs = np.random.RandomState(1)
    out = []
    for i in range(n):
        c = Neuro_controller({"controllers": [{"class": "neurocontroller.Neuro_controller", "weights": list(rs.uniform(-1, 1, 156)),
                                               "hidden_layer": [10, 7, 4], "activation": "tanh", "time": 20, "evolve": False}]})
        r = Robot(identifier=i, x=rob["x"], y=rob["y"], orientation=(37 * i) % 360, vision_range=tuple(rob["sonar_range"]),
                  sensors=rob["sonars"], radius=rob["radius"], max_speed=rob["max_speed"], controller=c)
        r.has_noise = False
        out.append(r)
    return out


for n in (1, 64, 1024, 8192):
    robots = make(n)
    g = RobotGroup(robots, env)
    g.update(0.1)
    t0 = time.perf_counter()
    for _ in range(a.ticks):
        g.update(0.1)
    dt = time.perf_counter() - t0
    g.close()
    line = "%5d robots: group %.3f ms per tick (%.3g robot-steps/s)" % (n, 1e3 * dt / a.ticks, n * a.ticks / dt)
    if n <= 64:
        solo = make(n)
        for r in solo:
            r.update(env, 0.1)
        t0 = time.perf_counter()
        for _ in range(a.ticks):
            for r in solo:
                r.update(env, 0.1)
        dt1 = time.perf_counter() - t0
        line += "; one by one %.3f ms per tick" % (1e3 * dt1 / a.ticks)
    print(line, flush=True)
