"""Host restatement of the device GA (r2p2_b200/csrc/r2p2_ga.cuh) -- TEST INFRASTRUCTURE ONLY.

numpy versions of Philox4x32-10, the generation-0 draw and the breeding step, used by tests/test_device_ga.py to check
the CUDA kernels bit for bit (selection, crossover, mutation sites) and to the ulps of log / cospi (mutated values).
Nothing under r2p2_b200/ imports this module.
"""
import numpy as np

GA_INIT, GA_TOURNAMENT, GA_CROSSFLAG, GA_GENE, GA_NORMAL = 1, 2, 3, 4, 5


def philox4x32(c0, c1, c2, c3, seed):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays of counters, key = the 64-bit seed -- the generator the device
    code uses.  returns four uint32 arrays."""
    c = [np.asarray(a, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for a in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [a.astype(np.uint32) for a in c]


def _u53(a, b):
    return ((a.astype(np.uint64) << np.uint64(21)) ^ (b.astype(np.uint64) >> np.uint64(11))).astype(np.float64) * (1.0 / 9007199254740992.0)


def init_population_host(P, G, seed, min_gen=-1.0, max_gen=1.0):
    """generate_float (evolution.py:18-22) with the device's draws: bit-identical to r2p2_ga_create."""
    g, i = np.meshgrid(np.arange(G), np.arange(P))
    c = philox4x32(g, i, 0, GA_INIT, seed)
    return min_gen + _u53(c[0], c[1]) * (max_gen - min_gen)


def breed_host(pop, fit, generation, seed, *, tournament=2, crossover_rate=0.9, mutation_rate=0.05, mutation_sigma=0.1,
               elites=1, min_gen=-1.0, max_gen=1.0):
    """Host restatement of r2p2_ga_breed: elites, tournament winners, crossover masks and mutation sites are
    bit-identical to the device; mutated gene values agree to the last ulps of log / cospi."""
    pop = np.asarray(pop, dtype=np.float64)
    fit = np.asarray(fit, dtype=np.float64)
    P, G = pop.shape
    ne = min(elites, P)
    nxt = np.empty_like(pop)
    order = np.argsort(-fit, kind="stable")
    nxt[:ne] = pop[order[:ne]]
    nc = P - ne
    child = np.arange(nc)
    parents = []
    for s in (0, 1):
        c = philox4x32(s, child, generation, GA_TOURNAMENT, seed)
        cand = np.stack([((c[q].astype(np.uint64) * np.uint64(P)) >> np.uint64(32)).astype(np.int64) for q in range(tournament)], axis=1)
        parents.append(cand[child, np.argmax(fit[cand], axis=1)])      # ties: the first drawn
    cf = philox4x32(0, child, generation, GA_CROSSFLAG, seed)
    do_x = _u53(cf[0], cf[1]) < crossover_rate
    g, ch = np.meshgrid(np.arange(G), child)
    c = philox4x32(g, ch, generation, GA_GENE, seed)
    from_b = do_x[:, None] & (_u53(c[0], c[1]) < 0.5)
    v = np.where(from_b, pop[parents[1]], pop[parents[0]])
    mut = _u53(c[2], c[3]) < mutation_rate
    n = philox4x32(g, ch, generation, GA_NORMAL, seed)
    u1 = (((n[0].astype(np.uint64) << np.uint64(21)) ^ (n[1].astype(np.uint64) >> np.uint64(11))).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * _u53(n[2], n[3]))
    v = np.where(mut, v + mutation_sigma * z, v)
    nxt[ne:] = np.minimum(np.maximum(v, min_gen), max_gen)
    return nxt, {"elites": order[:ne], "parents": parents, "from_b": from_b, "mutated": mut}
