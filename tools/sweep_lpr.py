#!/usr/bin/env python3
"""Measured choice of lanes per robot: every mapping x population x workload shape, kernel time of a T-tick launch.

    python tools/sweep_lpr.py [--ticks 300] [--out profiles/r02_lpr_sweep.json]

Writes the table and, per (shape, population), the fastest mapping -- the numbers behind `choose_lpr`'s thresholds in
r2p2_b200/csrc/r2p2_capi.cu (the automatic choice is printed next to the measured best, so a stale rule shows up as a
mismatch).  Shapes: the bench workloads (C3 = 5 rays + 5-10-7-4-2 on stage.png, E = 32 rays + 32-16-8-2 on the 4096^2
synthetic stage, C1 = 8 rays + linear controller on track_2.png)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from r2p2_b200.evaluator import PopulationEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--ticks", type=int, default=300)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_lpr_sweep.json"))
ap.add_argument("--shapes", default="C3,E,C1")
ap.add_argument("--pops", default="1024,2048,4096,8192,16384,32768,65536,131072")
a = ap.parse_args()
table = {}
for shape in a.shapes.split(","):
    base = dict(bench.WORKLOADS[shape])
    base["T"] = a.ticks
    rob = bench.ROBOTS[base["robot"]]
    for P in [int(x) for x in a.pops.split(",")]:
        w = dict(base, P=P)
        poses = bench.start_poses(w, P)
        g = bench.make_genomes(w, 1234, 0, P)
        row = {}
        for lpr in (0, 1, 2, 4, 8, 16, 32):
            if lpr and P * lpr > 262144 * 8:          # (hours of run time for nothing)
                continue
            ev = PopulationEvaluator(bench.load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                                     sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=poses[0], y=poses[1], orientation=poses[2],
                                     ticks=a.ticks, lanes_per_robot=lpr)
            ms = []
            for _ in range(3):
                f = ev.evaluate(g)
                ms.append(ev.sim.last_run_ms())
            row["auto" if lpr == 0 else str(lpr)] = min(ms[1:])
            ev.close()
        best = min((k for k in row if k != "auto"), key=lambda k: row[k])
        row["best"] = int(best)
        row["auto_vs_best"] = row["auto"] / row[best]
        table["%s/%d" % (shape, P)] = row
        print("%-3s P=%-7d " % (shape, P) + "  ".join("%s:%.2f" % (k, row[k]) for k in ("1", "2", "4", "8", "16", "32") if k in row) +
              "  | best %s  auto %.2f ms (x%.2f of best)" % (best, row["auto"], row["auto_vs_best"]), flush=True)
json.dump({"ticks": a.ticks, "unit": "ms per launch (kernel time, best of 2 after one warm-up)", "table": table}, open(a.out, "w"), indent=1)
