#!/usr/bin/env python3
"""bench.py -- robot-steps/s of the r2p2 per-tick robot step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload E]

A "step" is one evaluation of a whole population: P robots x T ticks of Robot.update in one kernel launch per GPU.

Headline workload (``value``, ``e2e``, ``roofline``, ``cpu_baseline``): BASELINE.json configs[4], the synthetic sweep --
4096 x 4096 stage, 32-ray sonar, MLP 32-16-8-2, 262,144 robots x 1,000 ticks -- the largest configuration that fits one
GPU.  At N > 1 the SAME 262,144 robots are sharded over the N GPUs ("strong" scaling, as configs[4] asks: "262,144
robots at 1/2/4/8 GPUs"); the per-generation fitness all-gather (NCCL) is inside the timed step, there is no data-path
collective.

Two more measurements ride in the same JSON line, each with the full protocol (warm-up, device events on the launch
stream, L2 flushed between iterations, max over ranks):
  north_star   the 65,536-individual neuro population on res/stage.png (robot-neuro.json, MLP 5-10-7-4-2) sharded over
               the N GPUs -- 8,192 per GPU at N = 8, the configuration BASELINE.json's target (>= 1e10 sonar-ray
               samples/s on 8 x B200) is quoted on
  config_B     BASELINE.json configs[1] (scenario-neuro: 1,024 individuals x 1,000 ticks on track_2.png), rank 0's GPU
  per_tick_api the Robot.update drop-in (one launch per frame, state round-trips HBM) against the HBM roofline, N = 1

Keys of the headline:
  value     robot-steps/s with the genomes already resident in HBM (device events on the launch stream)
  e2e       the same through the public API (PopulationEvaluator.evaluate_ptr) from pinned HOST buffers:
            H2D of the genomes and D2H of the fitness inside the timed region
  roofline  the episode kernel against the FP64-issue ceiling (the path is bound by the FP64
            pipe + on-chip probes, not by HBM; the HBM figure is reported beside it as roofline_hbm)
  cpu_baseline  the C restatement of the reference (oracle/, libm flavour) on the host cores, bounded sample

--impl reference times that CPU restatement alone on the same config (the reference itself is pure Python at ~650
robot-steps/s/core and cannot travel to the GPU box; the oracle is pinned bit-exactly against it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _conf(name):
    """One of the reference's conf/*.json, byte copies under tests/golden/conf (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, "conf", name)) as fp:
        return json.load(fp)


ROBOTS = {name: _conf(name) for name in ("robot-neuro.json", "robot_omni.json", "robot-track.json")}
# 32 list-form rays i * 11.25 deg, vr [0.5, 25], radius 5 (the synthetic robot has no shipped json)
ROBOTS["robot-synth32"] = dict(x=100.0, y=100.0, orientation=0.0, sonar_range=[0.5, 25], radius=5, sonars=[i * 11.25 for i in range(32)], max_speed=50)
_NEURO_HIDDEN = _conf("controller-neuro.json")["hidden_layer"]          # [10, 7, 4]

WORKLOADS = {
    # BASELINE.json configs[1]: scenario-neuro.json
    "B": dict(name="scenario-neuro: P=1024 x T=1000, track_2.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
              stage="track_2", robot="robot-neuro.json", ctrl="neuro", hidden=_NEURO_HIDDEN, P=1024, T=1000, scale=1.0),
    # north-star headline: 65,536-individual neuro population on res/stage.png
    "C3": dict(name="north-star headline: 65,536 neuro individuals x T=1000 on stage.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
               stage="stage", robot="robot-neuro.json", ctrl="neuro", hidden=_NEURO_HIDDEN, P=65536, T=1000, scale=1.0),
    # shipped scenario-inspyred.json (C1): linear controller, 8 int-form sonars
    "C1": dict(name="scenario-inspyred as shipped: P=65536 x T=1000, track_2.png, robot_omni.json, linear G=18",
               stage="track_2", robot="robot_omni.json", ctrl="linear", hidden=None, P=65536, T=1000, scale=0.05),
    "B64k": dict(name="P=65536 x T=1000 on one GPU, track_2.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
                 stage="track_2", robot="robot-neuro.json", ctrl="neuro", hidden=_NEURO_HIDDEN, P=65536, T=1000, scale=1.0),
    # BASELINE.json configs[3]: track robot on res/track.png, 16,384 robots x 4 episodes (4 start headings)
    "D": dict(name="scenario-track: 16384 robots x 4 start headings (0/90/180/270), track.png, robot-track.json, MLP 5-10-7-4-2 tanh, T=1000",
              stage="track", robot="robot-track.json", ctrl="neuro", hidden=_NEURO_HIDDEN, P=65536, T=1000, scale=1.0, start="headings4"),
    # BASELINE.json configs[4]: synthetic sweep: 4096^2 stage, 32 list-form rays, MLP 32-16-8-2
    "E": dict(name="synthetic sweep: 4096x4096 stage (4096 random rectangles), 32 rays, MLP 32-16-8-2 tanh, P=262144 x T=1000",
              stage="synthetic4096", robot="robot-synth32", ctrl="neuro", hidden=[16, 8], P=262144, T=1000, scale=1.0, start="free_space"),
}
GENOME_BLOCK = 8192          # genomes are generated in blocks of this many individuals, seed = base + block index
_stage_cache = {}


def synthetic_stage(n=4096, n_rect=4096, seed=2026):
    """white n x n stage, black border 8 px, n_rect random axis-aligned black rectangles of side U(8,128)."""
    rs = np.random.RandomState(seed)
    env = np.full((n, n), 255, dtype=np.uint8)
    env[:8, :] = 0; env[-8:, :] = 0; env[:, :8] = 0; env[:, -8:] = 0
    for _ in range(n_rect):
        w_, h_ = rs.randint(8, 129, 2)
        x_, y_ = rs.randint(0, n - w_), rs.randint(0, n - h_)
        env[x_:x_ + w_, y_:y_ + h_] = 0
    return env


def start_poses(w, P, seed=4321):
    """(x0, y0, theta0) for the whole population of P robots (scalars when everybody starts at the robot json pose)."""
    rob = ROBOTS[w["robot"]]
    kind = w.get("start")
    if kind == "headings4":
        return (np.full(P, float(rob["x"])), np.full(P, float(rob["y"])), np.tile(np.array([0.0, 90.0, 180.0, 270.0]), P // 4 + 1)[:P])
    if kind == "free_space":
        env = load_stage(w["stage"])
        rs = np.random.RandomState(seed)
        wall = env == 0
        # clearance >= 8 px: erode free space with an integral image of the wall mask
        c = 8
        ii = np.pad(wall.astype(np.int32), ((1, 0), (1, 0))).cumsum(0).cumsum(1)
        n = env.shape[0]
        xs, ys = [], []
        while len(xs) < P:
            cx, cy = rs.randint(c + 1, n - c - 1, 2 * P), rs.randint(c + 1, n - c - 1, 2 * P)
            cnt = ii[cx + c + 1, cy + c + 1] - ii[cx - c, cy + c + 1] - ii[cx + c + 1, cy - c] + ii[cx - c, cy - c]
            ok = cnt == 0
            xs.extend((cx[ok] + rs.uniform(0, 1, ok.sum())).tolist()); ys.extend((cy[ok] + rs.uniform(0, 1, ok.sum())).tolist())
        return np.array(xs[:P]), np.array(ys[:P]), rs.uniform(0, 360, P)
    return float(rob["x"]), float(rob["y"]), float(rob["orientation"])


def slice_poses(poses, lo, hi):
    return tuple(np.ascontiguousarray(a[lo:hi]) if np.ndim(a) else a for a in poses)


def load_stage(name):
    if name not in _stage_cache:
        _stage_cache[name] = synthetic_stage() if name == "synthetic4096" else np.load(os.path.join(GOLDEN, "stage_%s.npz" % name))["env"]
    return _stage_cache[name]


def genome_len(w):
    rob = ROBOTS[w["robot"]]
    from r2p2_b200 import expand_sonars
    R = len(expand_sonars(rob["sonars"]))
    if w["ctrl"] == "neuro":
        sizes = [R] + list(w["hidden"]) + [2]
        return sum(a * b for a, b in zip(sizes[:-1], sizes[1:])), R
    return 2 * R + 2, R


def make_genomes(w, seed, lo, hi):
    """Individuals [lo, hi) of the generation: uniform(-1, 1) (mirrors evolution.py:65), generated block by block so that
    every sharding of the population sees the same individuals."""
    G, _ = genome_len(w)
    out = np.empty((hi - lo, G))
    b = lo // GENOME_BLOCK
    while b * GENOME_BLOCK < hi:
        blk = np.random.RandomState(seed + b).uniform(-1, 1, (GENOME_BLOCK, G))
        a, z = max(lo, b * GENOME_BLOCK), min(hi, (b + 1) * GENOME_BLOCK)
        out[a - lo:z - lo] = blk[a - b * GENOME_BLOCK:z - b * GENOME_BLOCK]
        b += 1
    return out * w["scale"]


def shard(P, world, rank):
    base, extra = divmod(P, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def config_of(w):
    """The `config` object of the JSON line: identical for both arms (--impl ours / reference)."""
    G, R = genome_len(w)
    return {"workload": w["name"], "population": w["P"], "ticks": w["T"], "genome_len": G, "rays": R, "evolve": False,
            "delta": 0.1, "noise": "off", "l2": "flushed between timed iterations (256 MiB write)"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(w, genomes, threads, reps=1, poses=None):
    """The oracle (C restatement of the reference, libm flavour) on the host cores. Returns (robot_steps/s, out, seconds)."""
    from oracle.oracle import Oracle, expand_sonars
    rob = ROBOTS[w["robot"]]
    x0, y0, t0_ = poses if poses is not None else (rob["x"], rob["y"], rob["orientation"])
    O = Oracle(port=False)
    env = load_stage(w["stage"]).astype(np.int32)
    _, R = genome_len(w)
    layers = ([R] + list(w["hidden"]) + [2]) if w["ctrl"] == "neuro" else []
    t0 = time.perf_counter()
    steps = 0
    for _ in range(reps):
        out = O.run(env, rays=expand_sonars(rob["sonars"]), vr=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                    ctrl=w["ctrl"], genomes=genomes, x0=x0, y0=y0, theta0=t0_, T=w["T"], layers=layers,
                    evolve=False, threads=threads)
        steps += int(out["state"]["ticks"].sum())
    dt = time.perf_counter() - t0
    return steps / dt, out, dt


CPU_SAMPLE = 4096        # individuals of the workload the CPU arm evaluates per step (the first ones of the population)


def run_reference(args, w):
    """--impl reference: the CPU implementation of the path on this box's host cores, same config object."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_ind = min(w["P"], CPU_SAMPLE)
    genomes = make_genomes(w, 1234, 0, n_ind)
    poses = slice_poses(start_poses(w, w["P"]), 0, n_ind)
    for _ in range(args.warmup):
        cpu_port_run(w, genomes[:64], cores, poses=slice_poses(poses, 0, 64))
    t = []
    for _ in range(args.steps):
        _, _, dt = cpu_port_run(w, genomes, cores, poses=poses)
        t.append(dt)
    total_steps = n_ind * w["T"] * args.steps
    value = total_steps / sum(t)
    sample = "each step = the first %d of the %d individuals x %d ticks (C restatement of the reference, %d pthreads); warm-up steps use 64 individuals" % (
        n_ind, w["P"], w["T"], cores)
    line = {"impl": "reference", "metric": "robot-steps/sec", "value": value, "unit": "robot-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w),
            "cpu_baseline": {"value": value, "unit": "robot-steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "the reference itself is pure Python (~650 robot-steps/s/core measured in the build container); the port is pinned bit-exactly against it; "
                                     "throughput of a bounded sample, the same at every N (host cores do not scale with --gpus)"},
            "e2e": {"value": value, "unit": "robot-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure(ctx, w, steps, warmup, lpr=0, smem=-1, sample_clocks=False, want_cpu=False, local_only=False):
    """One workload, full protocol.  The population w['P'] is sharded over the ranks (local_only: rank 0 alone runs all
    of it, the others wait at the barriers).  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from r2p2_b200.evaluator import PopulationEvaluator
    from r2p2_b200.distributed import gather_fitness

    world, rank, dev, stream = ctx.world, ctx.rank, ctx.dev, ctx.stream
    nshards = 1 if local_only else world
    active = (rank == 0) or not local_only
    rob = ROBOTS[w["robot"]]
    P_total, T = w["P"], w["T"]
    G, R = genome_len(w)
    lo, hi = shard(P_total, nshards, rank if not local_only else 0)
    P = hi - lo
    res = None
    ev = None
    if active:
        poses = slice_poses(start_poses(w, P_total), lo, hi)
        ev = PopulationEvaluator(load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                                 sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=poses[0], y=poses[1],
                                 orientation=poses[2], ticks=T, delta=0.1, evolve=False, device=ctx.local,
                                 lanes_per_robot=lpr, stage_in_smem=smem)
        sim = ev.sim
        sim.set_stream(stream.cuda_stream)
        genomes_np = make_genomes(w, 1234, lo, hi)             # this rank's block of the generation
        g_pin = torch.from_numpy(genomes_np).pin_memory()
        g_dev = g_pin.to(dev, non_blocking=True)
        f_pin = torch.empty(P, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()

        def exchange():
            if nshards > 1:
                gather_fitness(_as_tensor(sim.fitness_device_ptr(), P, dev), P_total)   # every rank gets all fitness values

        def one_step_resident():
            ev.launch_resident()
            exchange()

        sim.set_genomes_device(g_dev.data_ptr(), P, G)
        for _ in range(max(warmup, 3)):
            one_step_resident()
        torch.cuda.synchronize()
        launches0 = sim.launch_count()

    sampler = ClockSampler(ctx.local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms, e2e_s, kern_ms = 0.0, 0.0, []
    if active:
        evs = []
        for _ in range(steps):
            ctx.flush.zero_()                               # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one_step_resident()
            e1.record(stream)
            evs.append((e0, e1))
            torch.cuda.synchronize()
            kern_ms.append(sim.last_run_ms())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if active:
        launches = sim.launch_count() - launches0
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        counters = sim.counters()                            # totals of the last evaluation (reset zeroes them)
        fitness_dev = sim.fitness().copy()
        # ---- e2e: host buffers in, host fitness out, through the public API ---------------------------
        for _ in range(2):
            ev.evaluate_ptr(g_pin.data_ptr(), P, G, f_pin.data_ptr())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if active:
        t0 = time.perf_counter()
        for _ in range(steps):
            ev.evaluate_ptr(g_pin.data_ptr(), P, G, f_pin.data_ptr())
            if nshards > 1:
                exchange()
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_fit_equal = bool(np.array_equal(f_pin.numpy(), fitness_dev))
    clocks = sampler.stop() if sampler else None

    tt = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tt[0]), float(tt[1])
    steps_per_eval = P_total * T                             # evolve=False: every robot runs exactly T ticks
    if rank == 0:
        value = steps_per_eval * steps / (total_ms * 1e-3)
        e2e_value = steps_per_eval * steps / (e2e_ms * 1e-3)
        # ---- roofline of the episode kernel (per launch, this rank's shard) -----------------------------
        ns = counters["sonar_samples"] / counters["robot_steps"]
        nc = counters["collision_samples"] / counters["robot_steps"]
        macs = (G if w["ctrl"] == "neuro" else 2 * R)
        a_fp64 = 7.0 * (ns + nc) + 720.0 + macs + 60.0 * (R + 2)
        k_ms = float(np.mean(kern_ms))
        achieved = counters["robot_steps"] * a_fp64 / (k_ms * 1e-3)
        hbm_bytes = P * (8 * G + 68)                                  # persistent mode: genome + pose in, results out
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(ctx.workload_key.get(w["name"], ""))
        except (OSError, ValueError):
            pass
        roofline = {"bound": "fp64-issue", "achieved": achieved / 1e9, "peak": ctx.fp64_peak / 1e9, "unit": "G FP64 inst/s",
                    "frac": achieved / ctx.fp64_peak, "traffic": traffic,
                    "peak_source": "measured in this run: register-resident DFMA loop on all SMs (r2p2_fp64_peak)",
                    "kernel": "episode_kernel", "kernel_ms": k_ms, "a_fp64_per_robot_step": a_fp64,
                    "sonar_samples_per_step": ns, "collision_samples_per_step": nc,
                    "robots_per_launch": P, "algorithmic_hbm_bytes_per_launch": hbm_bytes}
        roofline_hbm = {"bound": "hbm", "achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / ctx.hbm_peak, "peak_source": ctx.hbm_source,
                        "note": "state stays in registers for T ticks; HBM sees 8G+68 bytes per robot-episode"}
        res = {"value": value, "ms_per_step": total_ms / steps, "steps": steps, "warmup": max(warmup, 3),
               "population": P_total, "population_per_gpu": P, "gpus": nshards,
               "sonar_rays_per_s": value * R, "sonar_samples_per_s": value * ns,
               "e2e": {"value": e2e_value, "unit": "robot-steps/s", "h2d_bytes_per_step": P * G * 8 + 24, "d2h_bytes_per_step": P * 8,
                       "ms_per_step": e2e_ms / steps, "api": "PopulationEvaluator.evaluate_ptr (pinned host genomes -> host fitness)",
                       "fitness_equal_to_resident_run": e2e_fit_equal},
               "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm, "clocks": clocks}
        # ---- CPU baseline beside it (bounded sample, rank 0, N = 1 only) --------------------------------
        if want_cpu:
            cores = os.cpu_count() or 1
            n_ind = min(P, CPU_SAMPLE)
            reps = 1
            cposes = slice_poses(poses, 0, n_ind)
            v, out, dt = cpu_port_run(w, genomes_np[:n_ind], cores, poses=cposes)
            while dt * reps < 8.0 and reps < 16:
                reps *= 2
            if reps > 1:
                v, out, dt = cpu_port_run(w, genomes_np[:n_ind], cores, reps, poses=cposes)
            cpu_fit = out["fitness"]
            rel = np.abs(cpu_fit - fitness_dev[:n_ind]) / np.maximum(1e-300, np.abs(cpu_fit))
            res["cpu_baseline"] = {"value": v, "unit": "robot-steps/s", "cores": cores, "kind": "port",
                                   "sample": "the first %d individuals x %d ticks x %d repeats of the same genomes (%.1f s)" % (n_ind, T, reps, dt),
                                   "fitness_max_rel_diff_vs_gpu": float(rel.max()), "fitness_bit_equal": int((cpu_fit == fitness_dev[:n_ind]).sum())}
    if ev is not None:
        ev.close()
    return res


def per_tick_api(ctx, frames=200, P=262144):
    """Robot.update drop-in: one r2p2_run(h, 1, ...) per frame for P robots of config B's shape; the SoA state and the
    genomes cross HBM on every launch (2*11*8 + 8G + 8R bytes per robot-step)."""
    from r2p2_b200.evaluator import PopulationEvaluator
    w = dict(WORKLOADS["B"], P=P)
    rob = ROBOTS[w["robot"]]
    G, R = genome_len(w)
    ev = PopulationEvaluator(load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                             sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=float(rob["x"]), y=float(rob["y"]),
                             orientation=0.0, ticks=frames, device=ctx.local)
    sim = ev.sim
    sim.upload_genomes(make_genomes(w, 1234, 0, P))
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
    for _ in range(3):
        sim.run(1)
    sim.sync()
    sim.reset(float(rob["x"]), float(rob["y"]), 0.0)
    sim.sync()
    t0 = time.perf_counter()
    for _ in range(frames):
        sim.run(1)
    sim.sync()
    dt = time.perf_counter() - t0
    steps = int(sim.counters()["robot_steps"])
    ev.close()
    bps = 2 * 11 * 8 + 8 * G + 8 * R
    return {"workload": "Robot.update drop-in: %d robots of config B's shape, %d frames, one launch per frame" % (P, frames),
            "robot_steps_per_s": steps / dt, "us_per_frame": 1e6 * dt / frames, "hbm_bytes_per_robot_step": bps,
            "roofline": {"bound": "hbm", "achieved": steps * bps / dt / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
                         "frac": steps * bps / dt / 1e9 / ctx.hbm_peak, "peak_source": ctx.hbm_source}}


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import r2p2_b200

    ctx = Ctx()
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
    ctx.dev = torch.device("cuda", ctx.local)
    ctx.stream = torch.cuda.Stream(device=ctx.dev)      # everything (our kernels, NCCL, timing events) is ordered on this stream
    torch.cuda.set_stream(ctx.stream)
    ctx.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=ctx.dev)   # > 126 MB L2
    ctx.workload_key = {v["name"]: k for k, v in WORKLOADS.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    ctx.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ctx.hbm_source = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    probe = r2p2_b200.BatchSim(ctx.local)
    ctx.fp64_peak = probe.fp64_peak()
    probe.close()

    main_res = measure(ctx, w, args.steps, args.warmup, lpr=args.lpr, smem=args.smem, sample_clocks=True, want_cpu=(ctx.world == 1 and not args.no_cpu))
    extras = {}
    if not args.only:
        k2 = max(3, min(args.steps, 10))
        if args.workload != "C3":
            extras["north_star"] = measure(ctx, WORKLOADS["C3"], k2, args.warmup)
        if args.workload != "B":
            extras["config_B"] = measure(ctx, WORKLOADS["B"], k2, args.warmup, local_only=True)
        if ctx.world == 1:
            extras["per_tick_api"] = per_tick_api(ctx) if ctx.rank == 0 else None

    if ctx.rank == 0:
        r = main_res
        cfg = config_of(w)
        line = {"metric": "robot-steps/sec", "value": r["value"], "unit": "robot-steps/s", "n_gpus": ctx.world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "population_per_gpu": r["population_per_gpu"], "lanes_per_robot": args.lpr,
                "sonar_rays_per_s": r["sonar_rays_per_s"], "sonar_samples_per_s": r["sonar_samples_per_s"],
                "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"], "roofline": r["roofline"],
                "roofline_hbm": r["roofline_hbm"], "cpu_baseline": r.get("cpu_baseline")}
        for k, v in extras.items():
            if v is None:
                continue
            if k in ("north_star", "config_B"):
                v = dict(v)
                v.pop("clocks", None)
                v["config"] = config_of(WORKLOADS["C3" if k == "north_star" else "B"])
                v["unit"] = "robot-steps/s"
                if k == "north_star":
                    v["target"] = ">= 1e10 sonar-ray samples/s on 8 x B200 (BASELINE.json north_star)"
                    v["scaling"] = "strong (65,536 individuals sharded over the N GPUs)"
            line[k] = v
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _as_tensor(dev_ptr, n, dev):
    """Zero-copy float64 torch view of a device pointer owned by the library (for the NCCL all-gather)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(dev_ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="E", choices=sorted(WORKLOADS))
    ap.add_argument("--pop", type=int, default=0, help="override the workload's population (experiments)")
    ap.add_argument("--ticks", type=int, default=0, help="override the workload's ticks (experiments)")
    ap.add_argument("--lpr", type=int, default=0, help="lanes per robot (0 = auto)")
    ap.add_argument("--smem", type=int, default=-1, help="stage bit planes in shared memory: 1/0/-1 auto")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only", action="store_true", help="headline workload only (no north_star / config_B / per_tick_api sections)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.pop:
        w["P"] = args.pop
        w["name"] += " [population overridden: %d]" % args.pop
    if args.ticks:
        w["T"] = args.ticks
        w["name"] += " [ticks overridden: %d]" % args.ticks
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
