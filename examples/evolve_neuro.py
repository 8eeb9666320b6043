#!/usr/bin/env python3
"""Neuro-evolution of the reference's scenario-neuro setup on one or more B200s.

    python examples/evolve_neuro.py --generations 30 --pop 1024
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/evolve_neuro.py --pop 65536

Each generation is one kernel launch per GPU (1000 ticks, episode semantics of neurocontroller.py:72-85) plus, on
several GPUs, one NCCL all-gather of the fitness shards.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (stage fixtures + robot numbers)
from r2p2_b200.ea import GAConfig, GeneticAlgorithm  # noqa: E402
from r2p2_b200.evaluator import PopulationEvaluator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pop", type=int, default=1024); ap.add_argument("--generations", type=int, default=20)
ap.add_argument("--ticks", type=int, default=1000); ap.add_argument("--stage", default="track_2"); ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--device-ga", action="store_true", help="keep population, fitness and breeding on the GPU (r2p2_ga_*)")
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
rob = bench.ROBOTS["robot-neuro.json"]
ev = PopulationEvaluator(bench.load_stage(a.stage), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"], sonars=rob["sonars"],
                         controller="neuro", hidden_layer=[10, 7, 4], x=rob["x"], y=rob["y"], ticks=a.ticks, delta=0.1, delta_ctrl=100.0,
                         time_budget=a.ticks * 100.0, evolve=True, device=local)
if a.device_ga:
    # whole generations on the device: evaluate + breed are launches, the host reads the best fitness only
    import time
    from r2p2_b200.ea import DeviceGA
    sim = ev.sim
    sim.upload_genomes(np.zeros((a.pop if world == 1 else 1, 156)))
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0, time_budget=a.ticks * 100.0)
    ga = DeviceGA(sim, a.pop, 156, seed=a.seed)
    stepper = ga
    if world > 1:
        import torch
        import torch.distributed as dist
        from r2p2_b200.distributed import ShardedDeviceGA
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        stepper = ShardedDeviceGA(ga, "cuda:%d" % local)
    best = -np.inf
    for g in range(a.generations):
        t0 = time.perf_counter()
        rec = stepper.step(a.ticks, delta=0.1, delta_ctrl=100.0, evolve=True)
        best = max(best, rec["best"])
        if rank == 0:
            print("gen %4d  best %10.4f  %.1f ms" % (rec["generation"], rec["best"], 1e3 * (time.perf_counter() - t0)))
    if rank == 0:
        print("best fitness %.6f after %d generations of %d individuals (device GA)" % (best, a.generations, a.pop))
    sys.exit(0)
evaluate = ev.evaluate
if world > 1:
    import torch
    import torch.distributed as dist
    from r2p2_b200.distributed import ShardedEvaluator
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    evaluate = ShardedEvaluator(ev.evaluate, device="cuda:%d" % local).evaluate
ga = GeneticAlgorithm(GAConfig(pop_size=a.pop, genome_len=156, seed=a.seed), evaluate)
best, fit = ga.run(a.generations, verbose=(rank == 0))
if rank == 0:
    print("best fitness %.6f after %d generations of %d individuals" % (fit, a.generations, a.pop))
