// r2p2_ga.cuh -- the generation step of the neuro-evolution loop on the device.
//
// The reference leaves the EA to the user ("TODO: Implement your EA here", r2p2/evolution.py:58-68; its own driver
// ga.py.gpg is encrypted) and fixes only the contract: candidates are flat float lists drawn uniformly from
// [min_gen, max_gen] (generate_float, evolution.py:18-22), the evaluator maps candidates to fitness, higher is better.
// The operators here are the ones of inspyred's public ec.GA, which the reference imports: tournament selection,
// uniform crossover, Gaussian mutation clipped to the generator's range, generational replacement with elitism.
// With the population, the fitness vector and the breeding all in HBM a generation is a handful of launches and no
// host round trip.
//
// Every random number is a pure function of (seed, generation, individual, gene, purpose) through Philox4x32-10, so
// the result does not depend on the launch geometry, every GPU of a multi-GPU run breeds the same next generation
// from the all-gathered fitness without exchanging genomes, and a numpy restatement kept with the test infrastructure
// reproduces it: selection and crossover bit for bit, mutated genes to the last ulps of log / cospi.
#pragma once

#include <stdint.h>

#include "r2p2_kernels.cuh"

namespace r2p2 {

enum { kGaInit = 1, kGaTournament = 2, kGaCrossFlag = 3, kGaGene = 4, kGaNormal = 5 };
constexpr int kGaMaxElites = 64;

struct GaParams {
    int P, G, tournament, elites;
    double crossover_rate, mutation_rate, mutation_sigma, min_gen, max_gen;
    unsigned long long seed;
    uint32_t generation;            // of the parents
    const double* pop;              // [P][G] parents
    const double* fit;              // [P]
    double* next;                   // [P][G]
    int* elite_idx;                 // [kGaMaxElites] best individuals of the parent generation, best first
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], unsigned long long seed) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {     // [0, 1)
    return (double)(((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11)) * (1.0 / 9007199254740992.0);
}

// generate_float (evolution.py:18-22): gene = min_gen + u * (max_gen - min_gen)
__global__ void ga_init_kernel(double* __restrict__ pop, int P, int G, double min_gen, double max_gen, unsigned long long seed) {
    const size_t n = (size_t)P * (size_t)G;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)(i % (size_t)G), (uint32_t)(i / (size_t)G), 0u, (uint32_t)kGaInit};
        philox4x32_10(c, seed);
        pop[i] = __dadd_rn(min_gen, __dmul_rn(u53(c[0], c[1]), __dsub_rn(max_gen, min_gen)));
    }
}

// The `count` best individuals, best first, ties to the lower index (np.argsort(-fitness, kind="stable")).  One CTA;
// count passes of a block-wide arg-max that skips the winners of the earlier passes.
__global__ void ga_elites_kernel(const double* __restrict__ fit, int P, int count, int* __restrict__ elite_idx) {
    __shared__ double s_f[32];
    __shared__ int s_i[32];
    __shared__ int s_won[kGaMaxElites];
    for (int e = 0; e < count; ++e) {
        double bf = 0.0;
        int bi = -1;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            bool taken = false;
            for (int q = 0; q < e; ++q) taken |= (s_won[q] == i);
            const double f = fit[i];
            if (!taken && (bi < 0 || f > bf)) { bf = f; bi = i; }          // ascending i: the first maximum wins
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double of = __shfl_xor_sync(0xffffffffu, bf, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || of > bf || (of == bf && oi < bi))) { bf = of; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_f[threadIdx.x >> 5] = bf; s_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int nw = blockDim.x >> 5;
            bf = (threadIdx.x < nw) ? s_f[threadIdx.x] : 0.0;
            bi = (threadIdx.x < nw) ? s_i[threadIdx.x] : -1;
            for (int o = 16; o > 0; o >>= 1) {
                const double of = __shfl_xor_sync(0xffffffffu, bf, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || of > bf || (of == bf && oi < bi))) { bf = of; bi = oi; }
            }
            if (threadIdx.x == 0) { s_won[e] = bi; elite_idx[e] = bi; }
        }
        __syncthreads();
    }
}

// One warp per individual of the next generation: the first `elites` are copies of the best parents, the rest are
// children of two tournament winners.
__global__ void __launch_bounds__(128) ga_breed_kernel(const GaParams p) {
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= p.P) return;
    double* dst = p.next + (size_t)warp * p.G;
    if (warp < p.elites) {
        const double* src = p.pop + (size_t)p.elite_idx[warp] * p.G;
        for (int g = lane; g < p.G; g += 32) dst[g] = src[g];
        return;
    }
    const uint32_t child = (uint32_t)(warp - p.elites);
    // two parents: tournaments of `tournament` (<= 4) uniform candidates, the fittest wins, ties to the first drawn
    int parent[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        uint32_t c[4] = {(uint32_t)s, child, p.generation, (uint32_t)kGaTournament};
        philox4x32_10(c, p.seed);
        int best = (int)__umulhi(c[0], (uint32_t)p.P);
        double bf = p.fit[best];
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            if (q < p.tournament) {
                const int cand = (int)__umulhi(c[q], (uint32_t)p.P);
                const double f = p.fit[cand];
                if (f > bf) { bf = f; best = cand; }
            }
        }
        parent[s] = best;
    }
    uint32_t cf[4] = {0u, child, p.generation, (uint32_t)kGaCrossFlag};
    philox4x32_10(cf, p.seed);
    const bool do_x = u53(cf[0], cf[1]) < p.crossover_rate;
    const double* pa = p.pop + (size_t)parent[0] * p.G;
    const double* pb = p.pop + (size_t)parent[1] * p.G;
    for (int g = lane; g < p.G; g += 32) {
        uint32_t c[4] = {(uint32_t)g, child, p.generation, (uint32_t)kGaGene};
        philox4x32_10(c, p.seed);
        const bool from_b = do_x && (u53(c[0], c[1]) < 0.5);
        double v = from_b ? pb[g] : pa[g];
        if (u53(c[2], c[3]) < p.mutation_rate) {
            uint32_t n[4] = {(uint32_t)g, child, p.generation, (uint32_t)kGaNormal};
            philox4x32_10(n, p.seed);
            const double u1 = ((double)(((unsigned long long)n[0] << 21) ^ (unsigned long long)(n[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
            const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u53(n[2], n[3]));
            v = __dadd_rn(v, __dmul_rn(p.mutation_sigma, z));
        }
        dst[g] = fmin(fmax(v, p.min_gen), p.max_gen);
    }
}

}  // namespace r2p2
