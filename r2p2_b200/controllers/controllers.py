"""Controller loader with the reference's contract (r2p2/controllers/controllers.py:20-39):
``"class": "module.Class"`` names a class in this package; it is constructed with the whole scenario dict."""
import importlib


def load_controller(name):
    module, cls = name.split(".", 1)
    mod = importlib.import_module("r2p2_b200.controllers." + module)
    obj = mod
    for comp in cls.split("."):
        obj = getattr(obj, comp)
    return obj


def get_controllers(config):
    if "class" in config:
        return load_controller(config["class"])(config)
    out = []
    for ctrl in config.get("controllers", []):
        name = ctrl.get("class")
        if name is None and "hidden_layer" in ctrl:          # controller-neuro-simple.json has no "class" 
            name = "neurocontroller.Neuro_controller"
        if name is None:
            raise KeyError('controller configuration without a "class" attribute')
        out.append(load_controller(name)(config))
    return out
