"""Reference-format log files (r2p2/robot.py:158-168, 337-352) rebuilt from a trace: byte-identical to what the
unmodified reference wrote for the same episodes (tests/golden/logs.json, recorded by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import common
from r2p2_b200 import logs


def _golden():
    with open(os.path.join(common.GOLDEN, "logs.json")) as fp:
        return json.load(fp)


CASES = ["linear_omni", "const_wall", "const_stage", "fast_walls", "blank_border"]


def _stage(g):
    if g["stage"] == "blank.png":       # written into the reference's scratch tree by make_golden.py: all white
        return np.full(tuple(g["shape"]), 255, dtype=np.int32)
    return common.load_stage(g["stage"])


def _texts(g, pose, dst, cinfo):
    rj = g["robot_json"]
    return logs.episode_logs(pose, dst, cinfo=cinfo, shape=g["shape"], position0=(rj["x"], rj["y"]), identifier=g["identifier"], sensors=rj["sonars"], orientation0=rj["orientation"],
                             vision_range=rj["sonar_range"], radius=rj["radius"], max_speed=rj["max_speed"],
                             acceleration=g["acceleration"], write_stats=True)


@pytest.mark.parametrize("name", CASES)
def test_logs_from_oracle_trace(name, oracle_libm):
    from oracle.oracle import expand_sonars
    g = _golden()[name]
    rj = g["robot_json"]
    out = oracle_libm.run(_stage(g), ctrl=g["kind"], genomes=np.asarray(g["genome"])[None, :], x0=rj["x"], y0=rj["y"],
                          theta0=float(rj["orientation"]), T=g["T"], delta=g["delta"], layers=g["layers"], trace=True,
                          rays=expand_sonars(rj["sonars"]), vr=rj["sonar_range"], radius=rj["radius"], max_speed=rj["max_speed"])
    t = _texts(g, out["pose"][:, 0], out["dst"][:, 0], out["cinfo"][:, 0])
    assert t["sensors"] == g["sensors"]
    assert t["log"] == g["log"]


def test_sensor_block_types():
    """Clamped readings print the JSON-typed range (int), tick 0 prints the JSON-typed orientation."""
    text, line = logs.format_sensor_data(3, [0, 45], 0, logs.python_dst([55.0, 33.25], 55))
    assert text == "[ROBOT ID 3 SENSOR DATA:\nSENSOR: 0º - DISTANCE: 55\nSENSOR: 45º - DISTANCE: 33.25\n]"
    assert line == "55;33.25\n"


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_logs_from_device_trace(name):
    import r2p2_b200
    g = _golden()[name]
    rj = g["robot_json"]
    sim = r2p2_b200.BatchSim(0)
    sim.set_stage(_stage(g))
    sim.set_robot(sonar_range=rj["sonar_range"], radius=rj["radius"], max_speed=rj["max_speed"], sonars=rj["sonars"])
    sim.set_controller(g["kind"])
    sim.upload_genomes(np.asarray(g["genome"])[None, :])
    sim.reset(float(rj["x"]), float(rj["y"]), float(rj["orientation"]))
    sim.set_trace(True, g["T"])
    sim.run(g["T"], delta=g["delta"])
    tr = sim.trace(g["T"])
    t = _texts(g, tr["pose"][:, 0], tr["dst"][:, 0], tr["cinfo"][:, 0])
    assert t["sensors"] == g["sensors"]
    assert t["log"] == g["log"]


@pytest.mark.gpu
def test_robot_facade_writes_reference_logs(tmp_path):
    """``Robot(..., log_dir=...)`` + ``update(env, delta, True)``: the files the reference's Robot writes, same bytes."""
    from r2p2_b200.robot import Robot
    from r2p2_b200.controllers.inspyred import Inspyred_controller
    g = _golden()["linear_omni"]
    rj = g["robot_json"]
    ctrl = Inspyred_controller({"weights": g["genome"], "time": 20, "evolve": False})
    r = Robot(identifier=rj["id"], x=rj["x"], y=rj["y"], orientation=rj["orientation"], max_speed=rj["max_speed"], sensors=rj["sonars"],
              vision_range=rj["sonar_range"], radius=rj["radius"], controller=ctrl, log_dir=str(tmp_path))
    r.has_noise = False
    env = _stage(g)
    for _ in range(g["T"]):
        r.update(env, g["delta"], True)
    r._logs.close()
    assert (tmp_path / "robot_0_sensors.log").read_text() == g["sensors"]
    assert (tmp_path / "robot_0.log").read_text() == g["log"]
