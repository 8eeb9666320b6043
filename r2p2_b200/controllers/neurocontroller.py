"""Neuro_controller -- the reference's MLP controller (r2p2/controllers/neurocontroller.py:12-178), device form.

Configuration keys are the reference's (``weights``, ``hidden_layer``, ``activation``, ``time``, ``evolve``).
The reference ignores the dict it is given and re-reads a hard-coded ``../conf/controller-neuro.json``
(neurocontroller.py:22); here the controller dict of the scenario is used (``config['controllers']``), which
also makes ``scenario-neuro-simple.json`` loadable (its controller JSON has no "class" key).
"""
from ..controller import Controller


def _controller_dict(config, wanted):
    if config is None:
        return {}
    for c in config.get("controllers", []):
        if c.get("class", "").endswith(wanted) or ("hidden_layer" in c and wanted == "Neuro_controller"):
            return c
    return config


class Neuro_controller(Controller):
    device_kind = "neuro"

    def __init__(self, config=None):
        super().__init__("NEURO", config)
        f = _controller_dict(config, "Neuro_controller")
        self.weights = list(f.get("weights", []))
        self.hidden_layer = list(f.get("hidden_layer", [10, 7, 4]))
        self.activation = f.get("activation", "tanh")
        self.time = f.get("time", 20) * 1000            # neurocontroller.py:27
        self.epoch_time = self.time
        self.evolve = bool(f.get("evolve", False))
        self.distance = 0.0
        self.odom = (-1, -1)
        self.origin = (-1, -1)
        self._fitness = 0.0
        self.n_sonar = 0

    def register_robot(self, r):
        super().register_robot(r)
        self.odom = (r.x, r.y)
        self.origin = (r.x, r.y)
        self.n_sonar = len(r.sensors) if isinstance(r.sensors, list) else r.sensors

    def genome_length(self, n_rays):
        sizes = [n_rays] + self.hidden_layer + [2]
        return sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))

    def set_network_params(self, weights):
        """Flat genome, layer after layer, each ``reshape((n_in, n_out))`` in C order (neurocontroller.py:139-160)."""
        self.weights = [float(w) for w in weights]
        if self.robot is not None:
            self.robot._genome_changed()

    def on_collision(self, pos):
        self.time = 0

    def fitness(self):
        """distance travelled + displacement from the origin (neurocontroller.py:129-131), as mirrored from the device."""
        return self._fitness

    def has_cur_detected_edge_list(self):
        return True

    def device_spec(self, n_rays):
        g = list(self.weights)
        need = self.genome_length(n_rays)
        if len(g) != need:                      # the shipped JSONs carry placeholder weights ([0.0, 0.0])
            g = (g + [0.0] * need)[:need]
        return dict(kind="neuro", hidden_layer=self.hidden_layer, activation=self.activation, genome=g)
