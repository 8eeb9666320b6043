"""Host-side logic that needs no GPU: config merging, sonar expansion, plugin loading, stage loading."""
import json
import os

import numpy as np
import pytest

import common
from r2p2_b200 import expand_sonars
from r2p2_b200.config_manager import ConfigManager
from r2p2_b200.controllers.controllers import get_controllers, load_controller


def write_conf(tmp, stage="track_2"):
    """A scenario tree with the shape of the reference's conf/ + res/ (numbers of the shipped JSON files)."""
    from PIL import Image
    os.makedirs(os.path.join(tmp, "conf")); os.makedirs(os.path.join(tmp, "res"))
    env = common.load_stage(stage + ".png").astype(np.uint8)                # env[x, y]
    Image.fromarray(np.ascontiguousarray(env.T), mode="L").save(os.path.join(tmp, "res", stage + ".png"))
    for name in ("robot-neuro.json", "controller-neuro.json", "controller-neuro-simple.json", "controller-inspyred.json"):
        json.dump(common.conf(name), open(os.path.join(tmp, "conf", name), "w"))      # the reference's files, as shipped
    for name, ctrl in (("scenario-neuro.json", "controller-neuro.json"), ("scenario-neuro-simple.json", "controller-neuro-simple.json"),
                       ("scenario-inspyred.json", "controller-inspyred.json")):
        json.dump({"stage": "../res/%s.png" % stage, "robot": "../conf/robot-neuro.json", "controller": "../conf/" + ctrl, "gui": False},
                  open(os.path.join(tmp, "conf", name), "w"))
    return os.path.join(tmp, "conf")


def test_expand_sonars_matches_reference_quirk():
    """utils.py:465-466: range(0, 360, floor(365/n)) -- 8 -> 8 rays, 16 -> 17, 32 -> 33."""
    assert expand_sonars([0, 45, 70, 290, 315]) == [0.0, 45.0, 70.0, 290.0, 315.0]
    assert len(expand_sonars(8)) == 8 and expand_sonars(8)[:3] == [0.0, 45.0, 90.0]
    assert len(expand_sonars(16)) == 17 and len(expand_sonars(32)) == 33


def test_config_manager_merge_rules(tmp_path):
    conf = write_conf(str(tmp_path))
    cm = ConfigManager(os.path.join(conf, "scenario-neuro.json"), params={"evolve": False, "extra": 1})
    cfg = cm.get_config()
    assert cfg["extra"] == 1 and cfg["gui"] is False
    assert cfg["controllers"][0]["class"] == "neurocontroller.Neuro_controller"
    assert cfg["controllers"][0]["evolve"] is False                     # scenario / CLI keys override controller keys
    assert ConfigManager(os.path.join(conf, "scenario-neuro.json")).get_config()["controllers"][0]["evolve"] is True
    cm2 = ConfigManager(os.path.join(conf, "scenario-neuro.json"), controller_config=os.path.join(conf, "controller-inspyred.json"))
    assert cm2.get_config()["controllers"][0]["class"] == "inspyred.Inspyred_controller"


def test_plugin_loading_by_class_string(tmp_path):
    assert load_controller("neurocontroller.Neuro_controller").__name__ == "Neuro_controller"
    assert load_controller("inspyred.Inspyred_controller").device_kind == "linear"
    conf = write_conf(str(tmp_path))
    c = get_controllers(ConfigManager(os.path.join(conf, "scenario-neuro.json")).get_config())[0]
    assert c.hidden_layer == [10, 7, 4] and c.activation == "tanh" and c.time == 20000 and c.evolve is True
    assert c.genome_length(5) == 156
    s = get_controllers(ConfigManager(os.path.join(conf, "scenario-neuro-simple.json")).get_config())[0]   # no "class" key
    assert s.hidden_layer == [1] and s.genome_length(5) == 7
    lin = get_controllers(ConfigManager(os.path.join(conf, "scenario-inspyred.json")).get_config())[0]
    lin.set_network_params(list(range(12)))
    assert lin.linCoefs == [0, 1, 2, 3, 4] and lin.angCoefs == [5, 6, 7, 8, 9] and (lin.vlinmax, lin.vangmax) == (10, 11)


def test_stage_loading_is_the_reference_transpose(tmp_path):
    """PNG -> load_image -> flipud(rot90()) equals the env the reference builds (fixture recorded from it)."""
    from r2p2_b200.simulation import load_image, stage_env
    conf = write_conf(str(tmp_path))
    env = stage_env(load_image(os.path.join(conf, "..", "res", "track_2.png")))
    ref = common.load_stage("track_2.png")
    assert env.shape == ref.shape == (460, 321) and np.array_equal(env, ref)


def test_controller_base_contract():
    from r2p2_b200.controller import Controller, upd_sensor_angles
    import math
    c = Controller("X")
    c.set_ang([1, 2]); c.set_dst([3.0, math.inf])
    assert c.control([3.0, math.inf]) == (0, 0)
    assert c.cur_detected_edges_distances == [3.0, 10000]                 # controller.py:121-125

    class K(Controller):
        @upd_sensor_angles
        def control(self, dst):
            return 1, 2
    k = K(); k.set_ang([0])
    assert k.control([5.0]) == (1, 2) and k.actual_sensor_angles == [0]


def test_ga_improves_and_is_deterministic():
    """The GA driver (r2p2_b200/ea.py) on the reference's stub evaluator evaluate_test = sum(candidate) (evolution.py:54-56)."""
    from r2p2_b200.ea import GAConfig, GeneticAlgorithm
    ev = lambda pop: pop.sum(axis=1)
    runs = []
    for _ in range(2):
        ga = GeneticAlgorithm(GAConfig(pop_size=64, genome_len=10, min_gen=0, max_gen=2, seed=3), ev)
        best, fit = ga.run(25)
        runs.append((best.copy(), fit, [r["best"] for r in ga.log]))
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]          # same seed, same run
    bests = runs[0][2]
    assert bests[-1] > bests[0] and all(b2 >= b1 for b1, b2 in zip(bests, bests[1:]))     # elitism: monotone
    assert runs[0][1] > 17.0 and np.all(runs[0][0] <= 2.0) and np.all(runs[0][0] >= 0.0)


def test_history_log_records(tmp_path):
    """evolution.evaluate_ann's history records (r2p2/evolution.py:31,45-47): '\\n\\n[fitness]: [genes without spaces]' for every
    candidate that is at least as good as the best so far, in evaluation order."""
    import io
    from r2p2_b200.evolution import record_history
    buf = io.StringIO()
    best = record_history([[0.1, 0.2], [0.3, -0.4], [0.5, 0.6]], [1.5, 1.0, 1.5], buf, 0)
    assert best == 1.5
    assert buf.getvalue() == "\n\n[1.5]: [0.1,0.2]\n\n[1.5]: [0.5,0.6]"
    best = record_history([np.array([1.0, 2.0])], [np.float64(2.25)], buf, best)       # numpy inputs print like the reference's lists
    assert best == 2.25 and buf.getvalue().endswith("\n\n[2.25]: [1.0,2.0]")


def test_lanes_per_robot_table_matches_the_committed_sweep():
    """r2p2_lprtab.h (what choose_lpr looks up) is generated from profiles/r02_lpr_sweep.json (tools/sweep_lpr.py on a B200):
    the committed header must be the generator's output for the committed measurements, and every entry the fastest
    mapping of its row."""
    import json, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    table = json.load(open(os.path.join(root, "profiles", "r02_lpr_sweep.json")))["table"]
    hdr = open(os.path.join(root, "r2p2_b200", "csrc", "r2p2_lprtab.h")).read()
    rows = re.findall(r"\{(\d+), \{([0-9, ]+)\}\}", hdr)
    assert len(rows) == 3
    for (rmax, best), shape in zip(rows, ("C3", "C1", "E")):
        best = [int(b) for b in best.split(",")]
        for col, lpr in enumerate(best):
            row = table["%s/%d" % (shape, 1024 << col)]
            times = {int(k): v for k, v in row.items() if k.isdigit()}
            assert lpr == row["best"] == min(times, key=times.get), (shape, 1024 << col)
