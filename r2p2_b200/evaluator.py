"""Batched fitness evaluation: the public call that replaces the reference's file mailbox.

Reference: ``evolution.evaluate_ann(candidate, args)`` (r2p2/evolution.py:29-52) writes ONE genome to
``../res/weights.json``, waits for the simulator process to run a whole episode at 30 Hz and polls
``../res/fitness.txt``.  Here a whole population is one kernel launch.

``PopulationEvaluator.evaluate(genomes)`` is the public API bench.py times end to end: host genomes in,
host fitness out, every copy included.
"""
import numpy as np

from .sim import BatchSim


class PopulationEvaluator:
    """Evaluate [P, G] genomes -> fitness [P] for one (stage, robot, controller) configuration.

    Episode semantics follow neurocontroller.py:67-85 / inspyred.py:60-79 when ``evolve`` is true (an
    individual stops at the first control() call after a collision or when its timer runs out); with
    ``evolve`` false every individual runs exactly ``ticks`` updates (fixed work).
    """

    def __init__(self, stage, *, sonar_range, radius, max_speed, sonars, controller, hidden_layer=None,
                 activation="tanh", x=0.0, y=0.0, orientation=0.0, ticks=1000, delta=0.1, delta_ctrl=None,
                 time_budget=20000.0, evolve=False, device=0, lanes_per_robot=0, stage_in_smem=-1):
        self.sim = BatchSim(device)
        self.sim.set_stage(stage)
        self.sim.set_robot(sonar_range=sonar_range, radius=radius, max_speed=max_speed, sonars=sonars)
        self.sim.set_controller(controller, hidden_layer=hidden_layer, activation=activation)
        self.sim.set_mapping(lanes_per_robot, stage_in_smem)
        self.x0, self.y0, self.th0 = x, y, orientation
        self.ticks, self.delta = int(ticks), float(delta)
        self.delta_ctrl = float(delta if delta_ctrl is None else delta_ctrl)
        self.time_budget, self.evolve = float(time_budget), bool(evolve)

    # -- whole evaluation from host memory (the e2e path) --
    def evaluate(self, genomes, out=None):
        g = np.ascontiguousarray(genomes, dtype=np.float64)
        if g.ndim != 2:
            raise ValueError("genomes must be [P, G]")
        f = np.empty(g.shape[0]) if out is None else out
        if f.size < g.shape[0] or f.dtype != np.float64 or not f.flags["C_CONTIGUOUS"]:
            raise ValueError("fitness buffer must be a contiguous float64 array of at least %d elements" % g.shape[0])
        self.evaluate_ptr(g.ctypes.data, g.shape[0], g.shape[1], f.ctypes.data)
        return f

    # -- pieces, for callers that keep genomes on the device or use pinned buffers --
    def evaluate_ptr(self, genomes_host_ptr, P, G, fitness_host_ptr):
        """Same as evaluate() with raw host pointers (e.g. pinned torch tensors): no numpy temporaries.  One library call
        (r2p2_evaluate): the genome upload is pipelined against the episode kernel chunk by chunk."""
        self.sim.evaluate_ptr(genomes_host_ptr, P, G, self.x0, self.y0, self.th0, self.time_budget, self.ticks, self.delta,
                              self.delta_ctrl, self.evolve, fitness_host_ptr)

    def launch_resident(self):
        """Reset + run with the genomes already on the device (asynchronous)."""
        self.sim.reset(self.x0, self.y0, self.th0, self.time_budget)
        self.sim.run(self.ticks, self.delta, self.delta_ctrl, self.evolve)

    def close(self):
        self.sim.close()
