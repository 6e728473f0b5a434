// Device helpers shared by the evaluation kernels (gcp_eval.cu, gcp_fused.cu).
#pragma once
#include "gcp_common.cuh"

#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t hash_block(uint32_t b) { return b * 2654435761u; }
__device__ __forceinline__ uint32_t part_block(uint32_t b) { return (b * 0x85ebca6bu) >> 12; }

__device__ __forceinline__ int popc128(uint4 a)
{
    return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w);
}

__device__ __forceinline__ uint4 and128(uint4 a, uint4 b)
{
    return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w);
}
__device__ __forceinline__ void or128(uint4& a, uint4 b)
{
    a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
}

__device__ __forceinline__ double np_sum_order(const double* a, int n)
{
    // numpy pairwise_sum for n < 128 (verified against np.sum, tests/test_oracle_golden)
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}


// Row (agent) that holds position i of the compacted gene sequence: the smallest a with
// s_base[a + 1] > i (s_base = exclusive prefix sums of the row lengths, s_base[A] = n > i; rows of
// length 0 share their base with the next row and are skipped).  Bisection: A <= 32 -> 5 probes.
__device__ __forceinline__ int owner_row(const int* s_base, int A, int i)
{
    int lo = 0, hi = A - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_base[mid + 1] > i) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// -coverage (mappy.py:358-364, from the integer pixel counts c[0..3]) and travel cost
// (geneticalgorithm.py:196: np.sum over the per-agent sums, times flight_time_scale).
__device__ __forceinline__ void objective_values(int A, const gcp_params& prm, const gcp_map_consts& mc,
                                                 const uint32_t* c, const double* dist, const double* turn,
                                                 double& neg_cov, double& travel)
{
    const long long w_mask = 255ll * c[0] + c[2];
    const long long w_all = 255ll * c[1] + c[3];
    const double wall = __ddiv_rn(__ddiv_rn((double)(mc.gsum_mask + w_mask), 255.0),
                                  (double)mc.mask_size);
    const double sum_draw = __dadd_rn(mc.num_occluded, __ddiv_rn((double)w_all, 255.0));
    const double internal = __ddiv_rn(__dsub_rn(sum_draw, mc.num_occluded),
                                      __dsub_rn((double)mc.map_size, mc.num_occluded));
    const double coverage = __dadd_rn(__dmul_rn(prm.coverage_blend, wall),
                                      __dmul_rn(prm.one_minus_blend, internal));
    neg_cov = -coverage;
    const double td = np_sum_order(dist, A);
    const double tc = np_sum_order(turn, A);
    travel = __dmul_rn(__dadd_rn(td, tc), prm.flight_time_scale);
}

// The constraint half of geneticalgorithm.py:183-201 by ONE WARP (all 32 lanes must call): lane a owns
// row a of the loop-closure matrix.  Returns feasibility, ncomp = connected components of the graph
// lc[a][b] + lc[b][a] >= 1 (== A - #(eig(L) > 1e-6)).  Every lane gets the same result.
__device__ __forceinline__ bool constraints_warp(const int* s_lc, int A, const gcp_params& prm, int& ncomp)
{
    const int lane = threadIdx.x & 31;
    uint32_t m = 0;
    int rowsum = 0;
    bool sb = false;
    if (lane < A) {
        for (int b = 0; b < A; ++b) {
            const int v = s_lc[lane * A + b];
            if (b == lane) { sb = v < prm.min_solo_lcs; continue; }
            rowsum += v;                                             // entries <= 32,767 groups: no overflow
            if (v + s_lc[b * A + lane] >= 1) m |= 1u << b;
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) rowsum += __shfl_xor_sync(FULL, rowsum, o);
    const bool solo_bad = __any_sync(FULL, sb);
    // breadth-first search over the adjacency masks, identical (hence convergent) in every lane;
    // the mask of agent x is fetched from lane x
    ncomp = 0;
    uint32_t seen = 0;
    for (int a = 0; a < A; ++a) {
        if ((seen >> a) & 1u) continue;
        ++ncomp;
        uint32_t fr = 1u << a;
        while (fr) {
            seen |= fr;
            uint32_t nf = 0;
            for (uint32_t f = fr; f; f &= f - 1) nf |= __shfl_sync(FULL, m, __ffs(f) - 1);
            fr = nf & ~seen;
        }
    }
    return !(solo_bad || rowsum < prm.min_comb_lcs || ncomp > 1);
}

// Constraints + objectives of one organism (geneticalgorithm.py:177-201) from its loop-closure
// matrix (shared memory, A x A), integer coverage counters cov[0..3] and per-agent cost sums.
// Called by ALL 32 lanes of one warp: the A x A scan runs one row per lane (a single thread would
// keep the whole CTA - and, at one CTA per SM, the whole SM - waiting for A*A dependent iterations),
// lane 1 works out the objective values meanwhile, lane 0 writes the results.
__device__ __forceinline__ void finish_objectives(const int* s_lc, int A, const gcp_params& prm,
                                                  const gcp_map_consts& mc, const uint32_t* c,
                                                  const double* dist, const double* turn,
                                                  uint32_t org, double* __restrict__ obj,
                                                  double* __restrict__ neg_cov_out,
                                                  uint8_t* __restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    double neg_cov = 0.0, travel = 0.0;
    if (lane == 1) objective_values(A, prm, mc, c, dist, turn, neg_cov, travel);
    int ncomp;
    const bool feasible = constraints_warp(s_lc, A, prm, ncomp);
    neg_cov = __shfl_sync(FULL, neg_cov, 1);
    travel = __shfl_sync(FULL, travel, 1);
    if (lane != 0) return;
    if (neg_cov_out) neg_cov_out[org] = neg_cov;
    if (obj) {
        obj[(size_t)org * 2 + 0] = feasible ? neg_cov : 0.0;
        obj[(size_t)org * 2 + 1] = travel;
    }
    if (flags) {
        flags[(size_t)org * 4 + 0] = feasible ? 1 : 0;
        flags[(size_t)org * 4 + 2] = (uint8_t)ncomp;
    }
}
