#!/usr/bin/env python3
"""Cycles per tick phase of robot 0 (clock64 instrumentation, separate -DR2P2_PHASE_TIMING build)."""
import argparse, ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "ubench", "libr2p2_timing.so")
if "--build" in sys.argv:
    from r2p2_b200 import _lib
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + _lib.NVCC_FLAGS + ["-DR2P2_PHASE_TIMING", "-o", LIB, os.path.join(_lib.CSRC, "r2p2_capi.cu")])
    sys.exit(0)
from r2p2_b200 import _lib
_lib.LIB_PATH = LIB
import bench
from r2p2_b200.evaluator import PopulationEvaluator
ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="B"); ap.add_argument("--lpr", type=int, default=0); ap.add_argument("--ticks", type=int, default=1000)
a = ap.parse_args()
w = dict(bench.WORKLOADS[a.workload]); w["T"] = a.ticks
rob = bench.ROBOTS[w["robot"]]
ev = PopulationEvaluator(bench.load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                         sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=rob["x"], y=rob["y"], ticks=w["T"], lanes_per_robot=a.lpr)
g = bench.make_genomes(w, 1234, 0, w["P"])
L = _lib.lib()
out = (C.c_ulonglong * 16)()
ev.evaluate(g); L.r2p2_debug_phase_cycles(out, 1)
ev.evaluate(g); L.r2p2_debug_phase_cycles(out, 1)
names = ["sense (total incl. below)", "control pre-MLP", "MLP", "clamp+theta+heading sincos+cx,cy", "atan2+angle+sincos", "collision march",
         "push-out+border", "50-test+crash ring", "  sense: sincos", "  sense: march", "-", "loop top / sense prologue"]
tot = sum(out[i] for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11))
print("kernel ms %.3f; robot 0: %d ticks" % (ev.sim.last_run_ms(), ev.sim.state()["ticks"][0]))
for i, n in enumerate(names):
    if n != "-": print("%-36s %9.0f cycles/tick  %5.1f%%" % (n, out[i] / a.ticks, 100.0 * out[i] / tot))
print("total %.0f cycles/tick" % (tot / a.ticks))

import numpy as np
n = min(w["P"], 65536)
rc = np.zeros((n, 12), dtype=np.uint64)
L.r2p2_debug_robot_cycles(rc.ctypes.data_as(C.c_void_p), n)
rc = rc.astype(np.float64) / a.ticks
tot_r = rc.sum(1)
st = ev.sim.state()
order = np.argsort(tot_r)
print("per-robot cycles/tick: min %.0f  median %.0f  p90 %.0f  max %.0f" % (tot_r.min(), np.median(tot_r), np.percentile(tot_r, 90), tot_r.max()))
print("phase means (all robots):", " ".join("%d:%.0f" % (i, rc[:, i].mean()) for i in range(12)))
for r in list(order[-5:]) + list(order[:2]):
    print("robot %5d total %.0f ticks %d coll %d flags %x |" % (r, tot_r[r], st["ticks"][r], st["ncollide"][r], st["flags"][r]), " ".join("%d:%.0f" % (i, rc[r, i]) for i in range(12)))
