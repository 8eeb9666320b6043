#!/usr/bin/env python3
"""Small driver for ncu: config B (or another bench workload), a few evaluations, nothing else."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from r2p2_b200.evaluator import PopulationEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="B"); ap.add_argument("--lpr", type=int, default=0); ap.add_argument("--smem", type=int, default=-1)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--pop", type=int, default=0); ap.add_argument("--ticks", type=int, default=0)
a = ap.parse_args()
w = dict(bench.WORKLOADS[a.workload])
if a.ticks: w["T"] = a.ticks
if a.pop: w["P"] = a.pop
rob = bench.ROBOTS[w["robot"]]
px, py, pt = bench.start_poses(w, w["P"])
ev = PopulationEvaluator(bench.load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                         sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=px, y=py, orientation=pt, ticks=w["T"],
                         lanes_per_robot=a.lpr, stage_in_smem=a.smem)
g = bench.make_genomes(w, 1234, 0, w["P"])
for i in range(a.iters):
    f = ev.evaluate(g)
    print("iter", i, "kernel ms %.3f" % ev.sim.last_run_ms(), "fitness sum", repr(float(f.sum())))
print(ev.sim.counters())
