"""Fitness callbacks with the signatures of the reference's ``evolution.py`` (r2p2/evolution.py:12-56).

The reference evaluates ONE candidate per call through a file mailbox (``../res/weights.json`` ->
``../res/fitness.txt``, >= 0.5 s per candidate).  Here ``configure`` binds a scenario once and the whole
population handed over by the EA is one kernel launch.  ``evaluate_ann(candidate, args)`` keeps the
per-candidate signature; ``evaluator(candidates, args)`` is the batched form an inspyred-style EA calls
(inspyred.ec.evaluators.evaluator turns the former into the latter).
"""
import functools

import numpy as np

from .config_manager import ConfigManager
from .evaluator import PopulationEvaluator
from .simulation import load_image, stage_env

_default = None
_history = None          # open file of the run's history log (evolution.py:9), or None
_fitness_path = None     # the mailbox file the reference's controller answers through (neurocontroller.py:62-65), or None
cur_best = 0


def record_history(candidates, fits, history, best):
    """What evaluate_ann appends to ``../logs/history_<time>.log`` (r2p2/evolution.py:31,45-47), for a whole list of
    candidates in the order the reference would have evaluated them: a record for every candidate whose fitness is >=
    the best so far.  Returns the new best."""
    for c, f in zip(candidates, fits):
        if f >= best:
            string = str([float(v) for v in c]).replace(" ", "")          # the reference's candidates are lists of Python floats
            history.write("\n\n[" + str(float(f)) + "]: " + string)
            best = f
    history.flush()
    return best


def generate_const(random, args):
    return [0.1] * args.get("size", 10)


def generate_float(random, args):
    """evolution.generate_float (r2p2/evolution.py:18-22)."""
    lo, hi, size = args.get("min_gen", 0), args.get("max_gen", 2), args.get("size", 10)
    return [random.uniform(lo, hi) for _ in range(size)]


def make_evaluator(scenario, ticks=600, delta=1.0 / 30.0, delta_ctrl=1000.0 / 30.0, evolve=True, device=0, **kw):
    """Bind a scenario JSON (or a ConfigManager) to a PopulationEvaluator.

    Defaults reproduce the reference's live loop: 30 Hz (``clock.tick(30)``, utils.py:101-103), the robot
    integrating seconds and the controller timer counting milliseconds against ``time * 1000``."""
    cm = scenario if isinstance(scenario, ConfigManager) else ConfigManager(scenario)
    cfg = cm.get_config()
    import json
    robot = json.load(open(cm.resolve(cfg["robot"] if not isinstance(cfg["robot"], list) else cfg["robot"][0])))
    ctrl = cfg["controllers"][0]
    env = stage_env(load_image(cm.resolve(cfg["stage"])))
    neuro = "hidden_layer" in ctrl
    budget = float(ctrl.get("time", 20)) * 1000.0
    return PopulationEvaluator(env, sonar_range=robot["sonar_range"], radius=robot["radius"], max_speed=robot["max_speed"],
                               sonars=robot["sonars"], controller="neuro" if neuro else "linear",
                               hidden_layer=ctrl.get("hidden_layer"), activation=ctrl.get("activation", "tanh"),
                               x=robot["x"], y=robot["y"], orientation=robot["orientation"], ticks=ticks, delta=delta,
                               delta_ctrl=delta_ctrl, time_budget=budget, evolve=evolve, device=device, **kw)


def configure(scenario, history_path=None, fitness_path=None, **kw):
    """Set the evaluator used by ``evaluate_ann`` / ``evaluator`` when ``args`` carries none.

    history_path: write the reference's history log there (``'../logs/history_' + str(int(time.time())) + '.log'`` in the
    reference, evolution.py:9); fitness_path: also leave the last candidate's fitness in the reference's mailbox file
    (``'../res/fitness.txt'``) for tools that watch it."""
    global _default, _history, _fitness_path, cur_best
    _default = make_evaluator(scenario, **kw)
    if _history is not None:
        _history.close()
    _history = open(history_path, "a+") if history_path else None
    _fitness_path = fitness_path
    cur_best = 0
    return _default


def evaluator(candidates, args=None):
    """Batched form: list of genomes -> list of fitness (one launch)."""
    ev = (args or {}).get("r2p2_evaluator", _default)
    if ev is None:
        raise RuntimeError("call r2p2_b200.evolution.configure(scenario) first, or pass args['r2p2_evaluator']")
    global cur_best
    try:
        fits = ev.evaluate(np.asarray(candidates, dtype=np.float64)).tolist()
        if _history is not None and ev is _default:
            cur_best = record_history(candidates, fits, _history, cur_best)
        if _fitness_path and ev is _default and fits:
            with open(_fitness_path, "w") as fp:
                fp.write(str(fits[-1]))
        return fits
    except Exception as e:                                   # evolution.py:49-52: any failure scores 0
        print(__file__)
        print(e)
        return [0] * len(candidates)


def evaluate_ann(candidate, args=None):
    """Per-candidate signature of the reference (r2p2/evolution.py:29-52)."""
    return evaluator([candidate], args)[0]


def evaluate_test(candidate, args=None):
    return sum(candidate)


def batch_evaluator(evaluate):
    """What ``inspyred.ec.evaluators.evaluator`` does: f(candidate, args) -> f(candidates, args)."""
    @functools.wraps(evaluate)
    def wrapper(candidates, args):
        return [evaluate(c, args) for c in candidates]
    return wrapper
