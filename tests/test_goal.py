"""Goal flagging: the reference's Sequential_PID_Controller.is_at_goal (controllers/pid_controller.py:207-213) evaluated
by the reference itself on every tick of two golden episodes (tests/golden/goal.json) vs the oracle and the device."""
import json
import os

import numpy as np
import pytest

import common


def _golden():
    with open(os.path.join(common.GOLDEN, "goal.json")) as fp:
        return json.load(fp)


def _first_tick(hits):
    return (hits.index(True) + 1) if True in hits else 0xFFFFFFFF      # ticks completed when the predicate first held


@pytest.mark.parametrize("name", ["kat3_linear", "kat4_const"])
def test_oracle_goal_tick_equals_reference(name, oracle_libm):
    g = _golden()[name]
    rob = common.ROBOTS[g["robot"]]
    for i, goal in enumerate(g["goals"]):
        out = oracle_libm.run(common.load_stage(g["stage"]), ctrl=g["kind"], genomes=np.asarray(g["genome"])[None, :], x0=rob["x"], y0=rob["y"],
                              theta0=0.0, T=g["T"], goal=goal, **common.oracle_robot_kwargs(g["robot"]))
        want = _first_tick(g["at_goal"][str(i)])
        assert int(out["state"]["goal_tick"][0]) == want, (name, goal)
        assert bool(out["state"]["flags"][0] & 16) == (want != 0xFFFFFFFF)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat3_linear", "kat4_const"])
def test_device_goal_tick_equals_reference(name):
    g = _golden()[name]
    rob = common.ROBOTS[g["robot"]]
    for lpr in (1, 8, 32):
        sim = common.make_sim(g["stage"], g["robot"], g["kind"], lpr=lpr)
        sim.upload_genomes(np.asarray(g["genome"])[None, :])
        for i, goal in enumerate(g["goals"]):
            sim.set_goal(goal)
            sim.reset(rob["x"], rob["y"], 0.0)
            sim.run(g["T"])
            want = _first_tick(g["at_goal"][str(i)])
            assert int(sim.goal_ticks()[0]) == want, (name, goal, lpr)
            assert bool(sim.state()["flags"][0] & 16) == (want != 0xFFFFFFFF)
        sim.set_goal(None)
        sim.reset(rob["x"], rob["y"], 0.0)
        sim.run(5)
        assert int(sim.goal_ticks()[0]) == 0xFFFFFFFF and not (sim.state()["flags"][0] & 16)
        with pytest.raises(Exception):
            sim.set_goal((1e9, 5))                       # off the map: the reference raises IndexError
        sim.close()
