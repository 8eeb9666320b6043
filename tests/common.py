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
