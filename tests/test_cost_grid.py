"""Path planner's cost grid (path_planning.generate_grid + Node.value) against grids produced by the reference itself
(tests/golden/grids.npz: the three shipped stages and a synthetic one with intermediate grey levels)."""
import json

import numpy as np
import pytest

import common


def _cases():
    g = common.golden("grids")
    return g, json.loads(str(g["specs"]))


def _env(g, stage):
    return g["grey/env"].astype(np.int32) if stage == "grey" else common.load_stage(stage)


def test_restatement_equals_reference():
    from oracle.oracle import cost_grid
    g, specs = _cases()
    assert len(specs) == 20
    for sp in specs:
        v, c = cost_grid(_env(g, sp["stage"]), sp["divider"])
        assert np.array_equal(v, g[sp["key"] + "/values"]), sp["key"]
        assert np.array_equal(c, g[sp["key"] + "/costs"]), sp["key"]
    assert sorted(set(g["grey/64/costs"].ravel().tolist())) == [1, 2, 4, 5, 7, 8, 9]     # intermediate cost levels are exercised


@pytest.mark.gpu
def test_device_equals_reference():
    import r2p2_b200
    g, specs = _cases()
    for stage in sorted({sp["stage"] for sp in specs}):
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(_env(g, stage))
        for sp in (s for s in specs if s["stage"] == stage):
            v, c = sim.cost_grid(sp["divider"])
            assert np.array_equal(v, g[sp["key"] + "/values"]), sp["key"]
            assert np.array_equal(c, g[sp["key"] + "/costs"]), sp["key"]
        with pytest.raises(Exception):
            sim.cost_grid(0)
        sim.close()
