#!/usr/bin/env python3
"""bench.py -- robot-steps/s of the r2p2 per-tick robot step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload B]

A "step" is one evaluation of the whole population: P robots x T ticks of Robot.update in one kernel
launch.  At N = 1 the workload is BASELINE.json configs[1] (scenario-neuro: 1,024 individuals x 1,000
ticks, track_2.png, robot-neuro.json, MLP 5-10-7-4-2); at N > 1 every rank evaluates its own 1,024
individuals (weak scaling, population sharded with no data-path collective) and the per-generation
fitness all-gather (NCCL) is inside the timed step.

One JSON line is printed by rank 0:
  value     robot-steps/s with the genomes already resident in HBM (device events on the launch stream)
  e2e       the same through the public API (PopulationEvaluator.evaluate_ptr) from pinned HOST buffers:
            H2D of the genomes and D2H of the fitness inside the timed region
  roofline  the episode kernel against the FP64-issue ceiling (the path is bound by the FP64
            pipe + on-chip probes, not by HBM; the HBM figure is reported beside it as roofline_hbm)
  cpu_baseline  the C restatement of the reference (oracle/, libm flavour) on the host cores, same workload

--impl reference times that CPU restatement alone (the reference itself is pure Python at ~650
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

ROBOTS = {   # numbers of the reference's conf/robot-neuro.json, conf/robot_omni.json
    "robot-neuro.json": dict(x=230.0, y=250.0, orientation=0.0, sonar_range=[0.5, 25], radius=5, sonars=[0, 45, 70, 290, 315], max_speed=50),
    "robot_omni.json": dict(x=193.0, y=249.0, orientation=0.0, sonar_range=[1, 50], radius=5, sonars=8, max_speed=50),
}
WORKLOADS = {
    # BASELINE.json configs[1]: scenario-neuro.json
    "B": dict(name="scenario-neuro: P=1024 x T=1000, track_2.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
              stage="track_2", robot="robot-neuro.json", ctrl="neuro", hidden=[10, 7, 4], P=1024, T=1000, scale=1.0),
    # north-star headline shape (C3): neuro population on stage.png, 8192 individuals per GPU
    "C3": dict(name="scenario-inspyred headline: P=8192/GPU x T=1000, stage.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
               stage="stage", robot="robot-neuro.json", ctrl="neuro", hidden=[10, 7, 4], P=8192, T=1000, scale=1.0),
    # shipped scenario-inspyred.json (C1): linear controller, 8 int-form sonars
    "C1": dict(name="scenario-inspyred as shipped: P=8192/GPU x T=1000, track_2.png, robot_omni.json, linear G=18",
               stage="track_2", robot="robot_omni.json", ctrl="linear", hidden=None, P=8192, T=1000, scale=0.05),
    "B64k": dict(name="P=65536 x T=1000 on one GPU, track_2.png, robot-neuro.json, MLP 5-10-7-4-2 tanh",
                 stage="track_2", robot="robot-neuro.json", ctrl="neuro", hidden=[10, 7, 4], P=65536, T=1000, scale=1.0),
    # BASELINE.json configs[3]: track robot on res/track.png, 16,384 robots x 4 episodes (4 start headings)
    "D": dict(name="scenario-track: 16384 robots x 4 start headings (0/90/180/270), track.png, robot-track.json, MLP 5-10-7-4-2 tanh, T=1000",
              stage="track", robot="robot-track.json", ctrl="neuro", hidden=[10, 7, 4], P=65536, T=1000, scale=1.0, start="headings4"),
    # BASELINE.json configs[4]: synthetic sweep: 4096^2 stage, 32 list-form rays, MLP 32-16-8-2
    "E": dict(name="synthetic sweep: 4096x4096 stage (4096 random rectangles), 32 rays, MLP 32-16-8-2 tanh, P=262144 x T=1000",
              stage="synthetic4096", robot="robot-synth32", ctrl="neuro", hidden=[16, 8], P=262144, T=1000, scale=1.0, start="free_space"),
}
ROBOTS["robot-track.json"] = dict(x=260.0, y=250.0, orientation=0.0, sonar_range=[0.5, 40], radius=5, sonars=[0, 45, 70, 290, 315], max_speed=50)
ROBOTS["robot-synth32"] = dict(x=100.0, y=100.0, orientation=0.0, sonar_range=[0.5, 25], radius=5, sonars=[i * 11.25 for i in range(32)], max_speed=50)
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


def start_poses(w, P, seed):
    """(x0, y0, theta0) per robot for the workloads that do not start everybody at the robot json pose."""
    rob = ROBOTS[w["robot"]]
    kind = w.get("start")
    if kind == "headings4":
        return (np.full(P, rob["x"]), np.full(P, rob["y"]), np.tile(np.array([0.0, 90.0, 180.0, 270.0]), P // 4 + 1)[:P])
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
    return rob["x"], rob["y"], rob["orientation"]


def load_stage(name):
    if name not in _stage_cache:
        _stage_cache[name] = synthetic_stage() if name == "synthetic4096" else np.load(os.path.join(GOLDEN, "stage_%s.npz" % name))["env"]
    return _stage_cache[name]


def genome_len(w):
    rob = ROBOTS[w["robot"]]
    from r2p2_b200 import expand_sonars
    R = len(expand_sonars(rob["sonars"]))
    if w["ctrl"] == "neuro":
        sizes = [R] + w["hidden"] + [2]
        return sum(a * b for a, b in zip(sizes[:-1], sizes[1:])), R
    return 2 * R + 2, R


def make_genomes(w, seed, P):
    G, _ = genome_len(w)
    return np.random.RandomState(seed).uniform(-1, 1, (P, G)) * w["scale"]   # mirrors evolution.py:65


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
    """The oracle (C restatement of the reference, libm flavour) on the host cores. Returns (robot_steps/s, dict)."""
    from oracle.oracle import Oracle, expand_sonars
    rob = ROBOTS[w["robot"]]
    x0, y0, t0_ = poses if poses is not None else (rob["x"], rob["y"], rob["orientation"])
    O = Oracle(port=False)
    env = load_stage(w["stage"]).astype(np.int32)
    _, R = genome_len(w)
    layers = ([R] + w["hidden"] + [2]) if w["ctrl"] == "neuro" else []
    t0 = time.perf_counter()
    steps = 0
    for _ in range(reps):
        out = O.run(env, rays=expand_sonars(rob["sonars"]), vr=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                    ctrl=w["ctrl"], genomes=genomes, x0=x0, y0=y0, theta0=t0_, T=w["T"], layers=layers,
                    evolve=False, threads=threads)
        steps += int(out["state"]["ticks"].sum())
    dt = time.perf_counter() - t0
    return steps / dt, out, dt


def run_reference(args, w):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    P = min(w["P"], 1024)
    genomes = make_genomes(w, 1234, P)
    px, py, pt = start_poses(w, w["P"], 4321)
    poses = tuple(np.asarray(a)[:P] if np.ndim(a) else a for a in (px, py, pt))
    for _ in range(args.warmup):
        cpu_port_run(w, genomes[:64], cores, poses=tuple(a[:64] if np.ndim(a) else a for a in poses))
    t = []
    steps_per_s = []
    for _ in range(args.steps):
        v, _, dt = cpu_port_run(w, genomes, cores, poses=poses)
        steps_per_s.append(v); t.append(dt)
    total_steps = P * w["T"] * args.steps
    value = total_steps / sum(t)
    sample = "%d individuals x %d ticks per step (C restatement of the reference, %d pthreads)" % (P, w["T"], cores)
    line = {"impl": "reference", "metric": "robot-steps/sec", "value": value, "unit": "robot-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "population_per_step": P, "ticks": w["T"]},
            "cpu_baseline": {"value": value, "unit": "robot-steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "the reference itself is pure Python (~650 robot-steps/s/core measured in the build container); the port is pinned bit-exactly against it"},
            "e2e": {"value": value, "unit": "robot-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import r2p2_b200
    from r2p2_b200.evaluator import PopulationEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)        # everything (our kernels, NCCL, timing events) is ordered on this stream
    torch.cuda.set_stream(stream)

    rob = ROBOTS[w["robot"]]
    P, T = w["P"], w["T"]
    G, R = genome_len(w)
    px, py, pt = start_poses(w, P, 4321 + rank)
    ev = PopulationEvaluator(load_stage(w["stage"]), sonar_range=rob["sonar_range"], radius=rob["radius"], max_speed=rob["max_speed"],
                             sonars=rob["sonars"], controller=w["ctrl"], hidden_layer=w["hidden"], x=px, y=py,
                             orientation=pt, ticks=T, delta=0.1, evolve=False, device=local,
                             lanes_per_robot=args.lpr, stage_in_smem=args.smem)
    sim = ev.sim
    sim.set_stream(stream.cuda_stream)

    genomes_np = make_genomes(w, 1234 + rank, P)           # one generation, sharded by rank
    g_dev = torch.from_numpy(genomes_np).to(dev)
    g_pin = torch.from_numpy(genomes_np).pin_memory()
    f_pin = torch.empty(P, dtype=torch.float64).pin_memory()
    from r2p2_b200.distributed import gather_fitness
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def one_step_resident():
        ev.launch_resident()
        if world > 1:
            local_fit = _as_tensor(sim.fitness_device_ptr(), P, dev)
            gather_fitness(local_fit, P * world)   # per-generation exchange: every rank gets all P*world fitness values

    sim.set_genomes_device(g_dev.data_ptr(), P, G)
    for _ in range(max(args.warmup, 3)):
        one_step_resident()
    torch.cuda.synchronize()
    launches0 = sim.launch_count()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    kern_ms = []
    for _ in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations
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
    launches = sim.launch_count() - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    counters = sim.counters()                            # totals of the last evaluation (reset zeroes them)
    fitness_dev = sim.fitness().copy()

    # ---- e2e: host buffers in, host fitness out, through the public API -------------------------------
    for _ in range(2):
        ev.evaluate_ptr(g_pin.data_ptr(), P, G, f_pin.data_ptr())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.evaluate_ptr(g_pin.data_ptr(), P, G, f_pin.data_ptr())
        if world > 1:
            local_fit = _as_tensor(sim.fitness_device_ptr(), P, dev)
            gather_fitness(local_fit, P * world)   # per-generation exchange: every rank gets all P*world fitness values
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_fit_equal = bool(np.array_equal(f_pin.numpy(), fitness_dev))
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tt[0]), float(tt[1])
    steps_per_eval = P * T * world                       # evolve=False: every robot runs exactly T ticks
    value = steps_per_eval * args.steps / (total_ms * 1e-3)
    e2e_value = steps_per_eval * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        # ---- roofline of the episode kernel (per launch, this rank) ----------------------------------
        ns = counters["sonar_samples"] / counters["robot_steps"]
        nc = counters["collision_samples"] / counters["robot_steps"]
        macs = (G if w["ctrl"] == "neuro" else 2 * R)
        a_fp64 = 7.0 * (ns + nc) + 720.0 + macs + 60.0 * (R + 2)
        k_ms = float(np.mean(kern_ms))
        fp64_peak = sim.fp64_peak()
        achieved = counters["robot_steps"] * a_fp64 / (k_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = P * (8 * G + 68)                                  # persistent mode: genome + pose in, results out
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(args.workload)
        except (OSError, ValueError):
            pass
        roofline = {"bound": "fp64-issue", "achieved": achieved / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64 inst/s",
                    "frac": achieved / fp64_peak, "traffic": traffic,
                    "peak_source": "measured in this run: register-resident DFMA loop on all SMs (r2p2_fp64_peak)",
                    "kernel": "episode_kernel", "kernel_ms": k_ms, "a_fp64_per_robot_step": a_fp64,
                    "sonar_samples_per_step": ns, "collision_samples_per_step": nc,
                    "sonar_samples_per_s": counters["sonar_samples"] * world / (k_ms * 1e-3)}
        roofline_hbm = {"bound": "hbm", "achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                        "note": "state stays in registers for T ticks; HBM sees 8G+68 bytes per robot-episode"}
        # ---- CPU baseline beside it (bounded sample, rank 0, N = 1 only) --------------------------------
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n_ind = min(P, 1024)
            reps = 1
            cposes = tuple(np.asarray(a)[:n_ind] if np.ndim(a) else a for a in (px, py, pt))
            v, out, dt = cpu_port_run(w, genomes_np[:n_ind], cores, poses=cposes)
            while dt * reps < 8.0 and reps < 16:
                reps *= 2
            if reps > 1:
                v, out, dt = cpu_port_run(w, genomes_np[:n_ind], cores, reps, poses=cposes)
            cpu_fit = out["fitness"]
            rel = np.abs(cpu_fit - fitness_dev[:n_ind]) / np.maximum(1e-300, np.abs(cpu_fit))
            cpu = {"value": v, "unit": "robot-steps/s", "cores": cores, "kind": "port",
                   "sample": "%d individuals x %d ticks x %d repeats of the same genomes (%.1f s)" % (n_ind, T, reps, dt),
                   "fitness_max_rel_diff_vs_gpu": float(rel.max()), "fitness_bit_equal": int((cpu_fit == fitness_dev[:n_ind]).sum())}
        line = {"metric": "robot-steps/sec", "value": value, "unit": "robot-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "population_per_gpu": P, "ticks": T, "genome_len": G, "rays": R,
                           "l2": "flushed between timed iterations (256 MiB write)", "evolve": False,
                           "lanes_per_robot": args.lpr, "stage_in_smem": args.smem},
                "sonar_rays_per_s": value * R, "sonar_samples_per_s": value * ns,
                "e2e": {"value": e2e_value, "unit": "robot-steps/s", "h2d_bytes_per_step": P * G * 8 + 24, "d2h_bytes_per_step": P * 8,
                        "ms_per_step": e2e_ms / args.steps, "api": "PopulationEvaluator.evaluate_ptr (pinned host genomes -> host fitness)",
                        "fitness_equal_to_resident_run": e2e_fit_equal},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--lpr", type=int, default=0, help="lanes per robot (0 = auto)")
    ap.add_argument("--smem", type=int, default=-1, help="stage bit planes in shared memory: 1/0/-1 auto")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        if args.steps > 8:
            args.steps = 8
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
