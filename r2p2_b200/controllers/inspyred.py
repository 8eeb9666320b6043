"""Inspyred_controller -- the reference's linear controller (r2p2/controllers/inspyred.py:10-156), device form.

Chromosome ``[l_1..l_n, a_1..a_n, vmaxlin, vmaxang]``: ``lin = sum(l_i * dst_i)``, ``ang = sum(a_i * dst_i)``
(inspyred.py:87-100, 132-144).  Note the reference's time units: the first episode uses ``time`` unscaled,
later ones ``time * 1000`` (inspyred.py:26-27); both are exposed.
"""
from ..controller import Controller
from .neurocontroller import _controller_dict


class Inspyred_controller(Controller):
    device_kind = "linear"

    def __init__(self, config=None):
        super().__init__("Inspyred", config)
        f = _controller_dict(config, "Inspyred_controller")
        self.weights = list(f.get("weights", []))
        self.time = f.get("time", 20)
        self.epoch_time = self.time * 1000
        self.evolve = bool(f.get("evolve", False))
        self.distance = 0.0
        self.odom = (-1, -1)
        self.origin = (-1, -1)
        self._fitness = 0.0
        self.set_network_params(self.weights)

    def register_robot(self, r):
        super().register_robot(r)
        self.odom = (r.x, r.y)
        self.origin = (r.x, r.y)

    def set_network_params(self, chromosome):
        chromosome = [float(c) for c in chromosome]
        self.weights = chromosome
        if len(chromosome) >= 4:
            self.vlinmax, self.vangmax = chromosome[-2], chromosome[-1]
            self.linCoefs = chromosome[:int(len(chromosome) / 2 - 1)]
            self.angCoefs = chromosome[int(len(chromosome) / 2) - 1:-2]
        if getattr(self, "robot", None) is not None:
            self.robot._genome_changed()

    def on_collision(self, pos):
        self.time = 0

    def fitness(self):
        return self._fitness

    def device_spec(self, n_rays):
        g = list(self.weights)
        if len(g) < 4 or len(g) % 2:
            g = [0.0] * (2 * n_rays + 2)
        return dict(kind="linear", hidden_layer=None, activation="tanh", genome=g)
