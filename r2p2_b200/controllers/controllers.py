"""Controller loader with the reference's contract (r2p2/controllers/controllers.py:20-39):
``"class": "module.Class"`` names a controller class; it is constructed with the whole scenario dict.

Classes are looked up first in a directory of UNMODIFIED reference-style plugins when one is given (``plugin_dir=`` or
the environment variable ``R2P2_CONTROLLERS_DIR``, e.g. the ``r2p2/controllers`` directory of an r2p2 checkout), then
in this package (device-backed Neuro / Inspyred controllers, naive, PID).  Reference plugins do
``from controller import Controller`` and ``import utils as u`` with ``controllers/`` put on ``sys.path``
(controllers.py:3-26); ``plugin_modules`` installs the two names they need -- ``controller`` (this package's
``Controller`` / ``upd_sensor_angles``, same surface) and ``utils`` (``delta`` and ``npdata``, the two globals plugins
read: neurocontroller.py:73, inspyred.py:66, pid_controller.py:211) -- and puts the plugin directory and its parent
(for ``path_planning``) on ``sys.path`` for the duration of the import.
"""
import importlib
import os
import sys
import types


class plugin_modules:
    """Context manager: what an unmodified reference plugin imports, importable."""

    def __init__(self, plugin_dir, env=None, delta=None):
        self.dirs = [os.path.abspath(plugin_dir), os.path.dirname(os.path.abspath(plugin_dir))]
        self.env, self.delta = env, delta
        self.saved = {}

    def __enter__(self):
        from .. import controller as our_controller
        for name in ("controller", "utils"):
            self.saved[name] = sys.modules.get(name)
        sys.modules["controller"] = our_controller
        u = sys.modules.get("utils")
        if u is None or not getattr(u, "_r2p2_b200_shim", False):
            u = types.ModuleType("utils")
            u._r2p2_b200_shim = True
            u.delta = 0.1
            u.npdata = None
            u.pressed = []
            sys.modules["utils"] = u
        if self.env is not None:
            u.npdata = self.env
        if self.delta is not None:
            u.delta = self.delta
        for d in self.dirs:
            sys.path.insert(0, d)
        return u

    def __exit__(self, *exc):
        for d in self.dirs:
            try:
                sys.path.remove(d)
            except ValueError:
                pass
        return False


def publish_stage(env, delta=None):
    """What utils.display_image / calculate_delta publish for plugins (utils.py:101-103, 302-303): the transposed stage as
    ``utils.npdata`` and the tick length as ``utils.delta`` -- a no-op unless reference plugins were loaded."""
    u = sys.modules.get("utils")
    if u is not None and getattr(u, "_r2p2_b200_shim", False):
        u.npdata = env
        if delta is not None:
            u.delta = delta


def load_controller(name, plugin_dir=None):
    module, cls = name.split(".", 1)
    plugin_dir = plugin_dir or os.environ.get("R2P2_CONTROLLERS_DIR")
    mod = None
    if plugin_dir and os.path.exists(os.path.join(plugin_dir, module + ".py")):
        with plugin_modules(plugin_dir):
            mod = importlib.import_module(module)
    if mod is None:
        mod = importlib.import_module("r2p2_b200.controllers." + module)
    obj = mod
    for comp in cls.split("."):
        obj = getattr(obj, comp)
    return obj


def get_controllers(config, plugin_dir=None):
    if "class" in config:
        return load_controller(config["class"], plugin_dir)(config)
    out = []
    for ctrl in config.get("controllers", []):
        name = ctrl.get("class")
        if name is None and "hidden_layer" in ctrl:          # controller-neuro-simple.json has no "class" 
            name = "neurocontroller.Neuro_controller"
        if name is None and "goal" in ctrl and "ap" in ctrl:   # controller-PID2.json has no "class" either
            name = "pid_controller.Sequential_PID_Controller"
        if name is None:
            raise KeyError('controller configuration without a "class" attribute')
        out.append(load_controller(name, plugin_dir)(config))
    return out
