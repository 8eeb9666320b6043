"""Drive the UNMODIFIED r2p2 reference (read-only at /root/reference) as the parity oracle.

Test infrastructure only.  Used by ``make_golden.py`` in the build container to record golden
vectors; nothing here runs on the GPU box (the reference tree does not travel).

Recipe follows SURVEY.md section 8c:
  * matplotlib / pygame are absent -> empty stub modules are inserted in ``sys.modules``;
  * the reference opens ``../logs/*.log`` and reads ``../conf/*.json`` relative to cwd ->
    a scratch copy of ``r2p2/ conf/ res/`` plus an empty ``logs/`` is made under a temp dir and
    cwd is moved into ``<tmp>/r2p2``;
  * every Robot gets its own ``threading.Lock`` (the default one is shared by all robots,
    reference r2p2/robot.py:54).
"""
import json
import os
import shutil
import sys
import tempfile
import threading
import types

REF = os.environ.get("R2P2_REFERENCE", "/root/reference")


def _stub_modules():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    mpl = mod("matplotlib")
    for sub in ("colors", "pyplot", "animation", "widgets"):
        m = mod("matplotlib." + sub)
        setattr(mpl, sub, m)
    sys.modules["matplotlib.colors"].get_named_colors_mapping = lambda: {"blue": "#0000FF"}
    sys.modules["matplotlib.pyplot"].imshow = lambda *a, **k: None
    sys.modules["matplotlib.widgets"].Button = object
    mod("pygame")


class RefEnv:
    """Scratch tree + imported reference modules (``utils``, ``robot``, ``controller``)."""

    def __init__(self):
        self.tmp = tempfile.mkdtemp(prefix="r2p2_ref_")
        for d in ("r2p2", "conf", "res"):
            shutil.copytree(os.path.join(REF, d), os.path.join(self.tmp, d))
        os.makedirs(os.path.join(self.tmp, "logs"), exist_ok=True)
        self.cwd0 = os.getcwd()
        os.chdir(os.path.join(self.tmp, "r2p2"))
        sys.path.insert(0, os.path.join(self.tmp, "r2p2"))
        _stub_modules()
        import warnings
        warnings.simplefilter("ignore")
        import utils as u            # reference r2p2/utils.py
        import robot as robot_mod    # reference r2p2/robot.py
        import controller as ctrl    # reference r2p2/controller.py
        self.u, self.robot_mod, self.ctrl = u, robot_mod, ctrl

    def close(self):
        os.chdir(self.cwd0)
        shutil.rmtree(self.tmp, ignore_errors=True)

    # -- stage exactly as the reference builds it (utils.py:273-280 + 302-303) --
    def load_env(self, stage_png):
        import numpy as np
        npdata = self.u.load_image(os.path.join("..", "res", stage_png))
        return np.flipud(np.rot90(npdata))

    def write_conf(self, name, obj):
        with open(os.path.join(self.tmp, "conf", name), "w") as fp:
            json.dump(obj, fp)

    def make_robot(self, robot_json, controller):
        r = self.u.create_robot(os.path.join("..", "conf", robot_json), controller)
        r.lock = threading.Lock()
        r.has_noise = False
        return r

    def _import_controller(self, module):
        """Same mechanism as the reference loader (controllers/controllers.py:20-26)."""
        cdir = os.path.join(self.tmp, "r2p2", "controllers")
        sys.path.insert(0, cdir)
        try:
            return __import__(module)
        finally:
            sys.path.remove(cdir)

    def neuro_controller(self, weights, hidden, activation="tanh"):
        self.write_conf("controller-neuro.json", {
            "class": "neurocontroller.Neuro_controller", "weights": list(map(float, weights)),
            "hidden_layer": list(hidden), "activation": activation, "time": 20, "evolve": False})
        import contextlib, io
        neurocontroller = self._import_controller("neurocontroller")
        with contextlib.redirect_stdout(io.StringIO()):
            c = neurocontroller.Neuro_controller({})
        return c

    def linear_controller(self, chromosome):
        self.write_conf("controller-inspyred.json", {
            "class": "inspyred.Inspyred_controller", "weights": list(map(float, chromosome)),
            "time": 20, "evolve": False})
        import contextlib, io
        insp_ctrl = self._import_controller("inspyred")   # reference controllers/inspyred.py, not the PyPI package
        with contextlib.redirect_stdout(io.StringIO()):
            c = insp_ctrl.Inspyred_controller({})
        return c
