"""~1e6 steps of the unmodified reference's neuro path (tests/golden/population.npz) replayed through the oracle (CPU) and
the CUDA path (GPU): the MEASURED rate at which integer outcomes part ways (SURVEY.md hard-part 1).

The reference's MLP runs through sklearn -> OpenBLAS (summation order depends on the shape) and numpy's SIMD tanh; the
oracle and the device use an ordered FMA chain and tanh restatements.  Heading differences at the 1e-16 level can flip a
knife-edge sonar sample, after which the two trajectories are different trajectories.  The tests assert that
this is rare and that, up to the first such flip, poses agree to 1e-9 (north-star tolerance 1e-5); the measured numbers
are printed and, for the device, committed under profiles/ by tools/divergence_report.py.
"""
import numpy as np
import pytest

import common


def _check(stats, shape):
    print("population golden", shape, stats)
    assert stats["episodes_identical"] >= 0.95 * stats["episodes"], (shape, stats)
    assert stats["steps_before_divergence"] >= 0.97 * stats["steps"], (shape, stats)
    assert stats["pose_max_rel_diff_before_divergence"] < 1e-9, (shape, stats)


def test_oracle_vs_reference_population(oracle_libm):
    g, groups = common.population_groups()
    total = 0
    for shape, env, rob, hidden, T, lst in groups:
        from oracle.oracle import expand_sonars
        genomes, x0, y0, th0, sizes = common.population_inputs(rob, hidden, lst)
        out = oracle_libm.run(env, rays=expand_sonars(rob["sonars"]), vr=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                              ctrl="neuro", genomes=genomes, x0=x0, y0=y0, theta0=th0, T=T, layers=sizes, trace=True, threads=8)
        stats = common.population_divergence(g, lst, out["dst"], out["ncoll"], out["pose"], None)
        _check(stats, shape)
        total += stats["steps"]
    assert total >= 900_000


@pytest.mark.gpu
@pytest.mark.parametrize("lpr", [0, 1])
def test_device_vs_reference_population(lpr):
    import r2p2_b200
    g, groups = common.population_groups()
    for shape, env, rob, hidden, T, lst in groups:
        genomes, x0, y0, th0, sizes = common.population_inputs(rob, hidden, lst)
        sim = r2p2_b200.BatchSim(0)
        sim.set_stage(env)
        sim.set_robot(sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], sonars=rob["sonars"])
        sim.set_controller("neuro", hidden_layer=hidden)
        sim.set_mapping(lpr, -1)
        sim.upload_genomes(genomes)
        sim.reset(x0, y0, th0)
        sim.set_trace(True, T)
        sim.run(T)
        tr = sim.trace(T)
        stats = common.population_divergence(g, lst, tr["dst"], tr["ncoll"], tr["pose"], None)
        _check(stats, shape)
        sim.close()
