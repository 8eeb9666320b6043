"""Population sharding over the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Robots never interact, so the evaluation itself needs no collective: rank r evaluates the contiguous block
``shard_bounds(P, world, r)`` of the generation's genomes on its own GPU.  Per generation there is one
all-gather of the fitness shards (8 bytes per individual) and, for the EA's elitism, one broadcast of the best
genome from the rank that owns it.  Backend ``nccl`` on GPUs (NVLink/NVSwitch);
``gloo`` works for the host-logic tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(P, world, rank):
    """[lo, hi) of rank's block; blocks differ by at most one individual and cover 0..P in rank order."""
    base, extra = divmod(P, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_fitness(local_fit, P, group=None):
    """All ranks get the full fitness vector [P] (rank-ordered concatenation of the shards).

    local_fit: 1-D float64 tensor on the device the backend works with (CUDA for nccl, CPU for gloo)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(P, world, r)[1] - shard_bounds(P, world, r)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty(P, dtype=local_fit.dtype, device=local_fit.device)
        dist.all_gather_into_tensor(out, local_fit.contiguous(), group=group)
        return out
    # uneven shards: pad to the largest shard (collectives want equal sizes), then cut the padding out
    m = max(sizes)
    padded = torch.zeros(m, dtype=local_fit.dtype, device=local_fit.device)
    padded[:local_fit.numel()] = local_fit
    out = torch.empty(world * m, dtype=local_fit.dtype, device=local_fit.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)])


def broadcast_best(genomes_local, fitness_all, P, group=None):
    """Best individual of the generation on every rank: (index, fitness, genome[G]).  The owner rank broadcasts."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    best = int(torch.argmax(fitness_all).item())
    owner = next(r for r in range(world) if shard_bounds(P, world, r)[0] <= best < shard_bounds(P, world, r)[1])
    G = genomes_local.shape[1]
    buf = torch.empty(G, dtype=genomes_local.dtype, device=genomes_local.device)
    if rank == owner:
        buf.copy_(genomes_local[best - shard_bounds(P, world, rank)[0]])
    dist.broadcast(buf, src=owner if group is None else dist.get_global_rank(group, owner), group=group)
    return best, float(fitness_all[best].item()), buf


class ShardedEvaluator:
    """evaluate(genomes[P, G]) -> fitness[P] on every rank, each rank running only its shard.

    ``local_eval(genomes_shard) -> fitness_shard`` is the per-GPU evaluation (PopulationEvaluator.evaluate on a
    GPU box); ``device`` is where the gathered vector lives."""

    def __init__(self, local_eval, device="cpu", group=None):
        self.local_eval, self.device, self.group = local_eval, torch.device(device), group

    def evaluate(self, genomes):
        genomes = np.ascontiguousarray(genomes, dtype=np.float64)
        P = genomes.shape[0]
        lo, hi = shard_bounds(P, dist.get_world_size(self.group), dist.get_rank(self.group))
        local = np.asarray(self.local_eval(genomes[lo:hi]), dtype=np.float64)
        t = torch.from_numpy(local).to(self.device)
        return gather_fitness(t, P, self.group).cpu().numpy()


def device_view(dev_ptr, n, device):
    """Zero-copy float64 torch view of n doubles at a device pointer owned by the library."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(dev_ptr), False), "version": 3}
    return torch.as_tensor(h, device=device)


class ShardedDeviceGA:
    """The on-device EA (ea.DeviceGA) over the GPUs of one box.

    Every rank holds the whole population and the whole fitness vector in its own HBM and the same seed.  Per
    generation: rank r evaluates its block of individuals (one episode-kernel launch), the fitness blocks are
    all-gathered over NCCL straight into the GA's device fitness vector, and every rank breeds the identical next
    generation (Philox draws depend only on seed / generation / individual / gene).  No genome ever crosses a link."""

    def __init__(self, ga, device, group=None):
        self.ga, self.device, self.group = ga, torch.device(device), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lo, self.hi = shard_bounds(ga.P, self.world, self.rank)

    def step(self, ticks, delta=0.1, delta_ctrl=None, evolve=False):
        ga = self.ga
        ga.evaluate(ticks, delta, delta_ctrl, evolve, first=self.lo, count=self.hi - self.lo)
        full = device_view(ga.fitness_device_ptr(), ga.P, self.device)
        ga.sim.sync()                                       # the library's stream -> torch's stream
        gathered = gather_fitness(full[self.lo:self.hi].clone(), ga.P, self.group)
        full.copy_(gathered)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        ga.breed()
        i, f, _ = ga.best(genome=False)
        return {"generation": ga.generation - 1, "best": f, "best_index": i}
