"""A generational GA around the batched evaluator -- the loop the reference leaves to the user.

The reference's own GA driver is not available (r2p2/ga.py.gpg is encrypted) and ``evolution.__main__`` says
"TODO: Implement your EA here" (r2p2/evolution.py:58-68); what it fixes is the contract: candidates are flat
float lists from ``generate_float`` (uniform in [min_gen, max_gen], evolution.py:18-22), the evaluator maps
candidates to fitness, higher is better, and every new best genome goes to ``history_*.log``
(evolution.py:45-47).  This module follows inspyred's public ``ec.GA`` shape on the host: tournament
selection, n-point / blend crossover, Gaussian mutation, generational replacement with elitism.  All ranks of
a multi-GPU run hold the same EA state (same seed) and evaluate only their shard of each generation
(r2p2_b200/distributed.py); the fitness all-gather is the only exchange.
"""
import json
import time

import numpy as np


class GAConfig:
    def __init__(self, pop_size=1024, genome_len=156, min_gen=-1.0, max_gen=1.0, tournament=2, crossover_rate=0.9,
                 mutation_rate=0.05, mutation_sigma=0.1, elites=1, seed=0):
        self.__dict__.update(locals())
        del self.__dict__["self"]


class GeneticAlgorithm:
    def __init__(self, config, evaluate, history_path=None):
        """evaluate(genomes[P, G]) -> fitness[P] (e.g. PopulationEvaluator.evaluate or ShardedEvaluator.evaluate)."""
        self.c = config
        self.evaluate = evaluate
        self.rng = np.random.RandomState(config.seed)
        self.population = self.rng.uniform(config.min_gen, config.max_gen, (config.pop_size, config.genome_len))   # generate_float
        self.fitness = None
        self.generation = 0
        self.best_fitness = -np.inf
        self.best_genome = None
        self.history = open(history_path, "a+") if history_path else None
        self.log = []

    def _record_best(self):
        i = int(np.argmax(self.fitness))
        if self.fitness[i] >= self.best_fitness:                       # evolution.py:45-47 (>=: ties are logged too)
            self.best_fitness, self.best_genome = float(self.fitness[i]), self.population[i].copy()
            if self.history:
                self.history.write("\n\n[" + str(self.best_fitness) + "]: " + json.dumps(self.best_genome.tolist()).replace(" ", ""))
                self.history.flush()

    def step(self):
        """Evaluate the current generation, then breed the next one."""
        c, rng = self.c, self.rng
        t0 = time.perf_counter()
        self.fitness = np.asarray(self.evaluate(self.population), dtype=np.float64)
        eval_s = time.perf_counter() - t0
        self._record_best()
        P, G = self.population.shape
        order = np.argsort(-self.fitness, kind="stable")
        nxt = np.empty_like(self.population)
        ne = min(c.elites, P)
        nxt[:ne] = self.population[order[:ne]]
        # tournament selection of 2 * (P - ne) parents
        n_children = P - ne
        cand = rng.randint(0, P, (2 * n_children, c.tournament))
        winners = cand[np.arange(2 * n_children), np.argmax(self.fitness[cand], axis=1)]
        pa, pb = self.population[winners[:n_children]], self.population[winners[n_children:]]
        # uniform crossover
        do_x = rng.rand(n_children, 1) < c.crossover_rate
        mask = rng.rand(n_children, G) < 0.5
        child = np.where(do_x & mask, pb, pa)
        # Gaussian mutation, clipped to the generator's range
        mut = rng.rand(n_children, G) < c.mutation_rate
        child = np.where(mut, child + rng.normal(0.0, c.mutation_sigma, (n_children, G)), child)
        nxt[ne:] = np.clip(child, c.min_gen, c.max_gen)
        self.log.append({"generation": self.generation, "best": float(self.fitness.max()), "mean": float(self.fitness.mean()),
                         "eval_s": eval_s})
        self.population = nxt
        self.generation += 1
        return self.log[-1]

    def run(self, generations, verbose=False):
        for _ in range(generations):
            rec = self.step()
            if verbose:
                print("gen %4d  best %10.4f  mean %10.4f  eval %.1f ms" % (rec["generation"], rec["best"], rec["mean"], 1e3 * rec["eval_s"]))
        return self.best_genome, self.best_fitness


# ---------------------------------------------------------------------------------------------------------------
# The generation step on the device (r2p2_b200/csrc/r2p2_ga.cuh).  Its numpy restatement, the checker of
# tests/test_device_ga.py, lives with the rest of the test infrastructure, outside this package.
# ---------------------------------------------------------------------------------------------------------------
class DeviceGA:
    """Population, fitness and breeding resident on the GPU: a generation is `evaluate` + `breed`, all launches.

    sim is a configured ``BatchSim`` (stage, robot, controller set, and ``reset`` called once with the start pose
    and time budget).  In a multi-GPU run every rank owns a DeviceGA with the same seed, evaluates its shard
    (``evaluate(first, count)``), all-gathers into ``fitness_device_ptr`` and breeds the identical next generation."""

    def __init__(self, sim, pop_size, genome_len, seed=0, min_gen=-1.0, max_gen=1.0, tournament=2, crossover_rate=0.9,
                 mutation_rate=0.05, mutation_sigma=0.1, elites=1, history_path=None):
        import ctypes as C
        self._C = C
        self.sim, self.P, self.G = sim, int(pop_size), int(genome_len)
        self._L, self._h, self._ck = sim._L, sim._h, sim._ck
        self._ck(self._L.r2p2_ga_create(self._h, self.P, self.G, int(seed), float(min_gen), float(max_gen)))
        self._ck(self._L.r2p2_ga_set_operators(self._h, int(tournament), float(crossover_rate), float(mutation_rate),
                                               float(mutation_sigma), int(elites)))
        self.best_fitness, self.best_genome = -np.inf, None
        self.history = open(history_path, "a+") if history_path else None

    @property
    def generation(self):
        g = self._C.c_uint32()
        self._ck(self._L.r2p2_ga_generation(self._h, self._C.byref(g)))
        return g.value

    def population_device_ptr(self):
        p = self._C.c_void_p()
        self._ck(self._L.r2p2_ga_population_device(self._h, self._C.byref(p)))
        return p.value

    def fitness_device_ptr(self):
        p = self._C.c_void_p()
        self._ck(self._L.r2p2_ga_fitness_device(self._h, self._C.byref(p)))
        return p.value

    def evaluate(self, ticks, delta=0.1, delta_ctrl=None, evolve=False, first=0, count=None):
        count = self.P - first if count is None else count
        self._ck(self._L.r2p2_ga_evaluate(self._h, int(first), int(count), int(ticks), float(delta),
                                          float(delta if delta_ctrl is None else delta_ctrl), int(bool(evolve))))

    def save(self):
        """Checkpoint of the EA (bytes): seed, range, operators, generation counter, current generation and fitness."""
        n = self._C.c_uint64()
        self._ck(self._L.r2p2_ga_state_bytes(self._h, self._C.byref(n)))
        blob = np.empty(n.value, dtype=np.uint8)
        self._ck(self._L.r2p2_ga_save(self._h, blob.ctypes.data))
        return blob.tobytes()

    def load(self, blob):
        """Resume from ``save()``: the next ``breed`` draws exactly what the saved run would have drawn."""
        b = np.frombuffer(blob, dtype=np.uint8)
        self._ck(self._L.r2p2_ga_load(self._h, b.ctypes.data, int(b.size)))

    def write_fitness(self, fitness, first=0):
        f = np.ascontiguousarray(fitness, dtype=np.float64)
        self._ck(self._L.r2p2_ga_write_fitness(self._h, f.ctypes.data, int(first), int(f.size)))

    def breed(self):
        self._ck(self._L.r2p2_ga_breed(self._h))

    def read(self, population=True, fitness=True):
        pop = np.empty((self.P, self.G)) if population else None
        fit = np.empty(self.P) if fitness else None
        self._ck(self._L.r2p2_ga_read(self._h, pop.ctypes.data if population else None, fit.ctypes.data if fitness else None))
        return pop, fit

    def best(self, genome=True):
        """(index, fitness, genome) of the best individual of the generation last bred from."""
        i, f = self._C.c_int32(), self._C.c_double()
        g = np.empty(self.G) if genome else None
        self._ck(self._L.r2p2_ga_best(self._h, self._C.byref(i), self._C.byref(f), g.ctypes.data if genome else None))
        return i.value, f.value, g

    def step(self, ticks, delta=0.1, delta_ctrl=None, evolve=False):
        """One generation on one GPU: evaluate everyone, breed, and keep the best-genome history the reference's
        evaluator keeps (evolution.py:45-47)."""
        self.evaluate(ticks, delta, delta_ctrl, evolve)
        self.breed()
        i, f, g = self.best(genome=self.history is not None or True)
        if f >= self.best_fitness:
            self.best_fitness, self.best_genome = f, g
            if self.history:
                self.history.write("\n\n[" + str(f) + "]: " + json.dumps(g.tolist()).replace(" ", ""))
                self.history.flush()
        return {"generation": self.generation - 1, "best": f, "best_index": i}
