#!/usr/bin/env python3
"""First on-GPU shake-out: golden rays/ticks/episodes, oracle[port] bit-parity, timing by mapping."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import common
from oracle.oracle import Oracle
import r2p2_b200

def main():
    O = Oracle(port=True)
    # 1. rays
    r = common.golden("rays"); names = list(r["stages"])
    for si, st in enumerate(names):
        m = r["stage"] == si
        sim = r2p2_b200.BatchSim(0); sim.set_stage(common.load_stage(st))
        out = sim.cast_rays(r["ox"][m], r["oy"][m], r["angle"][m], r["limit"][m])
        bad = (out["status"] != r["hit"][m]) | (out["hx"] != r["hx"][m]) | (out["hy"] != r["hy"][m]) | (out["nsamp"] != r["nsamp"][m])
        print("rays", st, "mismatch", int(bad.sum()), "of", int(m.sum()))
        sim.close()
    # 2. single ticks vs golden (const controller)
    t = common.golden("ticks"); sn = list(t["stages"]); rn = list(t["robots"])
    for lpr in (1, 8, 32):
        for smem in (0, 1):
            tot = bad = 0
            for si, st in enumerate(sn):
                for ri, rj in enumerate(rn):
                    for delta in np.unique(t["delta"]):
                        m = (t["stage"] == si) & (t["robot"] == ri) & (t["delta"] == delta)
                        if not m.any(): continue
                        sim = common.make_sim(st, rj, "const", lpr=lpr, smem=smem)
                        sim.upload_genomes(np.stack([t["v"][m], t["w"][m]], axis=1))
                        sim.reset(t["x"][m], t["y"][m], t["theta"][m])
                        sim.set_trace(True, 1)
                        sim.run(1, delta=float(delta))
                        s = sim.state(); tr = sim.trace(1)
                        R = sim.R
                        b = (s["x"] != t["nx"][m]) | (s["y"] != t["ny"][m]) | (s["theta"] != t["ntheta"][m]) | (s["speed"] != t["nspeed"][m]) | (s["ncollide"] != t["ncollide"][m])
                        b |= np.any(tr["dst"][0] != t["dst"][m][:, :R], axis=1)
                        tot += int(m.sum()); bad += int(b.sum())
                        sim.close()
            print("ticks lpr", lpr, "smem", smem, "mismatch", bad, "of", tot)
    # 3. episodes vs golden
    e, specs = common.episode_specs()
    for sp in specs:
        n = sp["name"]; rob = common.ROBOTS[sp["robot"]]
        for lpr in (1, 8, 32):
            sim = common.make_sim(sp["stage"], sp["robot"], sp["kind"], hidden=sp["layers"][1:-1] if sp["layers"] else None, lpr=lpr)
            sim.upload_genomes(e[n + "/genome"][None, :]); sim.reset(rob["x"], rob["y"], 0.0); sim.set_trace(True, sp["T"])
            sim.run(sp["T"], delta=sp["delta"])
            tr = sim.trace(sp["T"]); pose = tr["pose"][:, 0]; gp = e[n + "/pose"]
            eq = np.all(pose == gp, axis=1)
            rel = np.max(np.abs(pose[:, :3] - gp[:, :3]) / np.maximum(1, np.abs(gp[:, :3])))
            keq = np.all(np.rint(tr["dst"][:, 0]) == np.rint(e[n + "/dst"]), axis=1)
            nc = np.cumsum(tr["ncoll"][:, 0])
            f = sim.fitness()[0]; gf = e[n + "/fitness"][-1]
            print("episode", n, "lpr", lpr, "pose-biteq", int(eq.sum()), "/", sp["T"], "maxrel %.1e" % rel, "k-eq", int(keq.sum()), "ncoll-eq", bool(np.array_equal(nc, e[n + "/ncoll"])), "fit", f, gf)
            sim.close()
    # 4. population vs oracle[port], bit-exact
    for (stage, rj, kind, hidden, G) in (("track_2.png", "robot-neuro.json", "neuro", [10, 7, 4], 156), ("stage.png", "robot-neuro.json", "neuro", [10, 7, 4], 156),
                                         ("track_2.png", "robot_omni.json", "linear", None, 18)):
        P, T = 256, 300
        rs = np.random.RandomState(1234)
        genomes = rs.uniform(-1, 1, (P, G)) * (1.0 if kind == "neuro" else 0.05)
        rob = common.ROBOTS[rj]
        layers = [5] + hidden + [2] if hidden else []
        for evolve in (False, True):
            ref = O.run(common.load_stage(stage), ctrl=kind, genomes=genomes, x0=rob["x"], y0=rob["y"], theta0=0.0, T=T, layers=layers,
                        evolve=evolve, delta_ctrl=33.0, time_budget=20000.0, threads=8, **common.oracle_robot_kwargs(rj))
            for lpr in (1, 8, 32):
                for smem in (0, 1):
                    sim = common.make_sim(stage, rj, kind, hidden=hidden, lpr=lpr, smem=smem)
                    sim.upload_genomes(genomes); sim.reset(rob["x"], rob["y"], 0.0, time_budget=20000.0)
                    sim.run(T, delta=0.1, delta_ctrl=33.0, evolve=evolve)
                    s = sim.state(); f = sim.fitness(); c = sim.counters()
                    st = ref["state"]
                    bad = (s["x"] != st["x"]) | (s["y"] != st["y"]) | (s["theta"] != st["theta"]) | (f != ref["fitness"]) | (s["ncollide"] != st["ncollide"]) | (s["flags"] != st["flags"]) | (s["ticks"] != st["ticks"])
                    print("pop", stage, kind, "evolve", evolve, "lpr", lpr, "smem", smem, "mismatch", int(bad.sum()), "of", P,
                          "samples", c["sonar_samples"], int(st["nsamp_sonar"].sum()), c["collision_samples"], int(st["nsamp_coll"].sum()), "ms %.3f" % sim.last_run_ms())
                    sim.close()
    # 5. timing config B
    rs = np.random.RandomState(1234)
    genomes = rs.uniform(-1, 1, (1024, 156))
    for lpr in (1, 2, 4, 8, 16, 32):
        for smem in (0, 1):
            sim = common.make_sim("track_2.png", "robot-neuro.json", "neuro", hidden=[10, 7, 4], lpr=lpr, smem=smem)
            sim.upload_genomes(genomes)
            ms = []
            for it in range(4):
                sim.reset(230, 250, 0.0); sim.run(1000, delta=0.1, evolve=False); ms.append(sim.last_run_ms())
            c = sim.counters()
            print("configB lpr", lpr, "smem", smem, "ms", ["%.3f" % m for m in ms], "steps/s %.3e" % (1024 * 1000 / (min(ms) * 1e-3)), c)
            sim.close()

if __name__ == "__main__":
    main()
