"""Generation step on the device (r2p2_b200/csrc/r2p2_ga.cuh) against its host restatement (r2p2_b200/ea.py)."""
import numpy as np
import pytest

import common
from oracle import ga_host
from r2p2_b200 import ea


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert [int(a) for a in ga_host.philox4x32(0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ff = 0xFFFFFFFF
    assert [int(a) for a in ga_host.philox4x32(ff, ff, ff, ff, (ff << 32) | ff)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(a) for a in ga_host.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, (0x299f31d0 << 32) | 0xa4093822)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_breed_host_properties():
    P, G, seed = 64, 20, 7
    pop = ga_host.init_population_host(P, G, seed)
    assert pop.min() >= -1.0 and pop.max() < 1.0 and len(np.unique(pop)) == P * G
    fit = np.random.RandomState(0).uniform(0, 100, P)
    nxt, info = ga_host.breed_host(pop, fit, 0, seed, elites=2)
    order = np.argsort(-fit, kind="stable")
    assert np.array_equal(nxt[0], pop[order[0]]) and np.array_equal(nxt[1], pop[order[1]])
    assert nxt.min() >= -1.0 and nxt.max() <= 1.0
    # unmutated genes come from one of the two parents
    pa, pb = pop[info["parents"][0]], pop[info["parents"][1]]
    child = nxt[2:]
    keep = ~info["mutated"]
    assert np.all((child[keep] == pa[keep]) | (child[keep] == pb[keep]))
    assert 0.02 < info["mutated"].mean() < 0.09
    # tournament pressure: parents are fitter than average
    assert fit[info["parents"][0]].mean() > fit.mean()
    # deterministic, and a different generation draws differently
    nxt2, _ = ga_host.breed_host(pop, fit, 0, seed, elites=2)
    nxt3, _ = ga_host.breed_host(pop, fit, 1, seed, elites=2)
    assert np.array_equal(nxt, nxt2) and not np.array_equal(nxt, nxt3)


def _sim():
    sim = common.make_sim("track_2.png", "robot-neuro.json", "neuro", hidden=[10, 7, 4])
    return sim


@pytest.mark.gpu
def test_device_init_and_breed_match_host():
    P, G, seed = 1000, 156, 1234
    sim = _sim()
    ga = ea.DeviceGA(sim, P, G, seed=seed, tournament=3, crossover_rate=0.8, mutation_rate=0.07, mutation_sigma=0.2, elites=3)
    pop, _ = ga.read()
    assert np.array_equal(pop, ga_host.init_population_host(P, G, seed))
    fit = np.random.RandomState(5).uniform(0, 100, P)
    fit[17] = fit[400] = fit.max() + 1.0            # a tie at the top: the lower index is the first elite
    ga.write_fitness(fit)
    ga.breed()
    nxt, _ = ga.read()
    ref, info = ga_host.breed_host(pop, fit, 0, seed, tournament=3, crossover_rate=0.8, mutation_rate=0.07, mutation_sigma=0.2, elites=3)
    assert list(info["elites"][:2]) == [17, 400]
    keep = np.ones_like(ref, dtype=bool)
    keep[3:] = ~info["mutated"]
    assert np.array_equal(nxt[keep], ref[keep])                       # elites, selection, crossover: bit for bit
    assert np.allclose(nxt[~keep], ref[~keep], rtol=0, atol=1e-13)    # Gaussian mutation: to the ulps of log / cospi
    assert ga.generation == 1
    i, f, g = ga.best()
    assert i == 17 and f == fit[17] and np.array_equal(g, pop[17])
    # second generation uses new draws
    ga.write_fitness(fit)
    ga.breed()
    nxt2, _ = ga.read()
    ref2, info2 = ga_host.breed_host(nxt, fit, 1, seed, tournament=3, crossover_rate=0.8, mutation_rate=0.07, mutation_sigma=0.2, elites=3)
    keep2 = np.ones_like(ref2, dtype=bool)
    keep2[3:] = ~info2["mutated"]
    assert np.array_equal(nxt2[keep2], ref2[keep2])


@pytest.mark.gpu
def test_device_generation_loop_matches_evaluator():
    """evaluate + breed on the device: fitness equals what PopulationEvaluator gives for the same genomes, the best
    fitness never decreases with elitism, and no genome leaves the device in between."""
    from r2p2_b200.evaluator import PopulationEvaluator
    rob = common.ROBOTS["robot-neuro.json"]
    P, G, T = 256, 156, 150
    sim = _sim()
    sim.upload_genomes(np.zeros((P, G)))
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0, time_budget=20000.0)
    ga = ea.DeviceGA(sim, P, G, seed=3, elites=2)
    ev = PopulationEvaluator(common.load_stage("track_2.png"), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                             sonars=rob["sonars"], controller="neuro", hidden_layer=[10, 7, 4], x=rob["x"], y=rob["y"], ticks=T)
    best = []
    for gen in range(4):
        pop, _ = ga.read(fitness=False)
        rec = ga.step(T)
        _, fit = ga.read(population=False)
        assert np.array_equal(fit, ev.evaluate(pop)), gen
        assert rec["best"] == fit.max() and rec["best_index"] == int(np.argmax(fit))
        best.append(rec["best"])
    assert all(b2 >= b1 for b1, b2 in zip(best, best[1:]))
