#!/usr/bin/env python3
"""Per-tick API (`Robot.update` for a list of robots = one r2p2_run(h, 1, ...) per frame): state round-trips HBM on
every launch (2 x 11 x 8 state + 8 G genome + 8 R dst bytes per robot-step).  Reports robot-steps/s and the
HBM-roofline fraction next to the persistent-episode number of the same population."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from r2p2_b200.evaluator import PopulationEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="B64k"); ap.add_argument("--pop", type=int, default=0); ap.add_argument("--frames", type=int, default=300)
a = ap.parse_args()
w = dict(bench.WORKLOADS[a.workload])
if a.pop: w["P"] = a.pop
rob = bench.ROBOTS[w["robot"]]
px, py, pt = bench.start_poses(w, w["P"])
ev = PopulationEvaluator(bench.load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                         sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=px, y=py, orientation=pt, ticks=a.frames)
sim = ev.sim
g = bench.make_genomes(w, 1234, 0, w["P"])
P, G, R = w["P"], g.shape[1], sim.R
# persistent episode
f_ref = ev.evaluate(g).copy()
t_persist = sim.last_run_ms()
# the same frames, one launch each
sim.upload_genomes(g)
sim.reset(px, py, pt)
sim.run(1); sim.sync()
sim.reset(px, py, pt)
sim.sync()
t0 = time.perf_counter()
for _ in range(a.frames):
    sim.run(1)
sim.sync()
dt = time.perf_counter() - t0
f_tick = sim.fitness()
steps = int(sim.counters()["robot_steps"])
bytes_per_step = 2 * 11 * 8 + 8 * G + 8 * R
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6543.7))
out = {"workload": a.workload, "P": P, "frames": a.frames, "per_tick_robot_steps_per_s": steps / dt, "us_per_frame": 1e6 * dt / a.frames,
       "persistent_robot_steps_per_s": P * a.frames / (t_persist * 1e-3), "persistent_ms": t_persist,
       "hbm_bytes_per_robot_step": bytes_per_step, "hbm_achieved_gbs": steps * bytes_per_step / dt / 1e9, "hbm_peak_gbs": hbm,
       "hbm_frac": steps * bytes_per_step / dt / 1e9 / hbm, "fitness_equal_to_one_launch": bool(np.array_equal(f_ref, f_tick))}
print(json.dumps(out))
