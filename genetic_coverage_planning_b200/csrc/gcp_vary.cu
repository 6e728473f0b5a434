// Selection + variation kernels (Philox counter-based RNG), sm_100a.
//
// Reference semantics restated (citations into /root/reference/src):
//   gori_tools.py:207-229        lowVarSample      -> k_low_var_select
//   geneticalgorithm.py:211-249  Organism.crossover / matchWaypt (:311-333) / padMinDNA (:203-209)
//                                                   -> k_crossover
//   geneticalgorithm.py:251-268  mutation, :270-309 muterpolate, :346-357 pruneUTurns,
//   :335-344 addTelomere, constructor order :156-175 -> k_mutate
//   pathmaker.py:175-213         makeMeAPath / assignWeights -> walk()
//   geneticalgorithm.py:37-53    firstGeneration's random paths -> k_random_walks
//   geneticalgorithm.py:78-84    arena truncation   -> k_gather_survivors
// The reference draws from numpy's global MT19937 stream; here every (generation, organism,
// agent) owns a Philox4x32-10 stream, so results are reproducible for a seed and independent
// of the launch shape, but not equal draw-for-draw to numpy: parity for these kernels is by
// invariants and distributions (tests/test_gpu_variation.py), as SURVEY.md section 8a states.
#include "gcp_common.cuh"
#include <cstring>

#define FULL 0xffffffffu
constexpr int MAX_MUT_POINTS = 64;

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint4 ctr;
    uint2 key;
    uint4 out;
    int have;
    __device__ Philox(uint64_t seed, uint32_t gen, uint32_t stream, uint32_t id)
    {
        key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        ctr = make_uint4(0u, id, stream, gen);
        have = 0;
        out = make_uint4(0, 0, 0, 0);
    }
    __device__ void round_(uint4& c, uint2& k)
    {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    __device__ uint32_t next()
    {
        if (have == 0) {
            uint4 c = ctr;
            uint2 k = key;
            #pragma unroll
            for (int r = 0; r < 10; ++r) round_(c, k);
            out = c;
            ctr.x += 1u;
            have = 4;
        }
        const uint32_t v = (have == 4) ? out.x : (have == 3) ? out.y : (have == 2) ? out.z : out.w;
        --have;
        return v;
    }
    __device__ float uniform() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }   // [0,1)
    // 16-bit coin: true with probability thr16 / 65536; two coins per 32-bit draw
    uint32_t coin_bits = 0; int coin_have = 0;
    __device__ bool coin(uint32_t thr16)
    {
        if (coin_have == 0) { coin_bits = next(); coin_have = 2; }
        const uint32_t v = coin_bits & 0xFFFFu;
        coin_bits >>= 16; --coin_have;
        return v < thr16;
    }
    __device__ double uniform64()
    {
        const uint64_t a = next(), b = next();
        return (double)(((a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ int randint(int n) { return n <= 1 ? 0 : min(n - 1, (int)(uniform() * (float)n)); }
};

enum { STREAM_ORG = 1, STREAM_XOVER = 2, STREAM_MUT = 3, STREAM_INIT = 4, STREAM_MUTERP = 5 };
__device__ __forceinline__ uint32_t stream_id(uint32_t stream, uint32_t salt) { return stream | (salt << 8); }

struct VaryParams {
    int A, L;
    int min_dna_len, time_thresh, num_muterp, srch_dist, path_memory;
    float xover_prob, mut_prob, muterp_prob, sub_prob, inv_sigma;
    uint32_t org_off;    // global id of local organism 0 (population shards): Philox keys only
};

// All row-level routines below are WARP-COOPERATIVE: one warp owns one agent row held in
// shared memory, every lane carries the same Philox state and takes the same branches
// (control flow is warp-uniform), and the lanes split the data-parallel parts (adjacency
// slots of a walk step, match counting, copies, compaction).

__device__ __forceinline__ int w_valid_len(const int32_t* r, int L, int lane)
{
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t m = __ballot_sync(FULL, i < L && r[i] == -1);
        if (m) return i0 + __ffs(m) - 1;
    }
    return L;
}

// pathmaker.py:188-213.  r[pos] is the current node; appends up to n_steps nodes (stops at
// L).  The no-return memory is the last `path_memory` nodes of the running path, never
// reaching below row index mem_floor: that is just r[max(mem_floor, pos-mem+1) .. pos].
// Lane e handles the e-th out-edge of the current node.  Returns the last index written.
__device__ int w_walk(const GcpTables& t, const VaryParams& vp, int32_t* r, int pos, int n_steps,
                      int mem_floor, Philox& rng, int lane)
{
    const float PI = 3.14159265358979f;
    const float heading = rng.uniform() * 2.0f * PI - PI;            // :190, never updated
    const float k2 = -0.5f * vp.inv_sigma * vp.inv_sigma;
    const int W = t.adj8 ? 8 : 32;                                   // lanes that can hold an edge
    int cur = r[pos];
    for (int s = 0; s < n_steps && pos + 1 < vp.L; ++s) {
        uint32_t c = 0xFFFFFFFFu;
        float th = 0.f;
        if (t.adj8) {                                                // one 64-B record per node
            const uint2 a = __ldg(t.adj8 + (size_t)cur * 8 + (lane & 7));
            if (lane < 8) { c = a.x; th = __uint_as_float(a.y); }
        } else {
            const uint32_t lo = __ldg(t.row_ptr + cur), hi = __ldg(t.row_ptr + cur + 1);
            if (lo + lane < hi) { c = __ldg(t.col + lo + lane); th = (float)__ldg(t.edge_theta + lo + lane); }
        }
        const bool valid = c != 0xFFFFFFFFu;
        int nxt = cur;
        if (__any_sync(FULL, valid)) {
            float ang = th - heading;                                // assignWeights :179-181
            ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
            const float w = valid ? __expf(k2 * ang * ang) : 0.f;
            bool seen = false;
            for (int q = max(mem_floor, pos - vp.path_memory + 1); q <= pos; ++q)
                seen |= ((uint32_t)r[q] == c);
            const float wf = (valid && !seen) ? w : 0.f;
            float ia = w, ifr = wf;                                  // inclusive scans in edge order
            for (int o = 1; o < W; o <<= 1) {
                const float va = __shfl_up_sync(FULL, ia, o), vf = __shfl_up_sync(FULL, ifr, o);
                if (lane >= o) { ia += va; ifr += vf; }
            }
            const float tot_all = __shfl_sync(FULL, ia, W - 1), tot_fresh = __shfl_sync(FULL, ifr, W - 1);
            const bool use_fresh = tot_fresh > 0.f;                  // :199-207
            const float u = rng.uniform() * (use_fresh ? tot_fresh : tot_all);
            const bool elig = valid && (!use_fresh || !seen);
            const uint32_t hit = __ballot_sync(FULL, elig && u < (use_fresh ? ifr : ia));
            const uint32_t em = __ballot_sync(FULL, elig);
            const int src = hit ? (__ffs(hit) - 1) : (31 - __clz(em));
            nxt = (int)__shfl_sync(FULL, c, src);
        }
        ++pos;
        GCP_DASSERT(pos < vp.L && nxt >= 0 && (uint32_t)nxt < t.N);
        if (lane == 0) r[pos] = nxt;
        __syncwarp();
        cur = nxt;
    }
    return pos;
}

// Thread-level walk (one thread = one row in GLOBAL memory), same rule as w_walk.  Used
// where many rows walk at once (mutation tails, initial population): 32 independent chains
// per warp hide the dependent adjacency loads, and nothing is replicated across lanes.
constexpr int PATH_MEM_MAX = 16;

// MEMT: compile-time capacity of the no-return memory (10 = the driver's path_memory, else 16)
template <typename G, int MEMT = PATH_MEM_MAX>
__device__ int t_walk(const GcpTables& t, const VaryParams& vp, G* r, int pos, int n_steps,
                      int mem_floor, Philox& rng)
{
    const float PI = 3.14159265358979f;
    const float heading = rng.uniform() * 2.0f * PI - PI;            // :190, never updated
    const float k2 = -0.5f * vp.inv_sigma * vp.inv_sigma;
    const int mem = min(vp.path_memory, MEMT);
    int hist[MEMT];
    #pragma unroll
    for (int k = 0; k < MEMT; ++k) {
        const int idx = pos - k;
        hist[k] = (k < mem && idx >= mem_floor) ? (int)r[idx] : -2;
    }
    int cur = (int)r[pos];
    for (int s = 0; s < n_steps && pos + 1 < vp.L; ++s) {
        uint32_t c[8];
        float w[8];
        int deg = 0;
        uint32_t lo = 0;
        if (t.adj8) {
            const uint4* rec = reinterpret_cast<const uint4*>(t.adj8 + (size_t)cur * 8);
            #pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint4 q = __ldg(rec + h);
                c[2 * h] = q.x; w[2 * h] = __uint_as_float(q.y);
                c[2 * h + 1] = q.z; w[2 * h + 1] = __uint_as_float(q.w);
            }
            deg = 8;
        } else {
            lo = __ldg(t.row_ptr + cur);
            deg = (int)(__ldg(t.row_ptr + cur + 1) - lo);
        }
        int nxt = cur;
        if (deg > 0) {
            float tot_all = 0.f, tot_fresh = 0.f;
            uint32_t seen_mask = 0;
            if (t.adj8) {
                #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool valid = c[e] != 0xFFFFFFFFu;
                    float ang = w[e] - heading;                      // assignWeights :179-181
                    ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
                    w[e] = valid ? __expf(k2 * ang * ang) : 0.f;
                    bool seen = !valid;
                    #pragma unroll
                    for (int k = 0; k < MEMT; ++k) seen |= (hist[k] == (int)c[e]);
                    seen_mask |= (seen ? 1u : 0u) << e;
                    tot_all += w[e];
                    if (!seen) tot_fresh += w[e];
                }
                const bool use_fresh = tot_fresh > 0.f;              // :199-207
                const float u = rng.uniform() * (use_fresh ? tot_fresh : tot_all);
                float acc = 0.f;
                int last_ok = -1;
                nxt = -1;
                #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool valid = c[e] != 0xFFFFFFFFu;
                    const bool skip = !valid || (use_fresh && ((seen_mask >> e) & 1u));
                    if (!skip && nxt < 0) {
                        acc += w[e];
                        last_ok = (int)c[e];
                        if (u < acc) nxt = (int)c[e];
                    }
                }
                if (nxt < 0) nxt = (last_ok >= 0) ? last_ok : cur;
            } else {
                for (int e = 0; e < deg; ++e) {
                    float ang = (float)__ldg(t.edge_theta + lo + e) - heading;
                    ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
                    const float ww = __expf(k2 * ang * ang);
                    const int cc = (int)__ldg(t.col + lo + e);
                    bool seen = false;
                    #pragma unroll
                    for (int k = 0; k < MEMT; ++k) seen |= (hist[k] == cc);
                    tot_all += ww;
                    if (!seen) tot_fresh += ww;
                }
                const bool use_fresh = tot_fresh > 0.f;
                const float u = rng.uniform() * (use_fresh ? tot_fresh : tot_all);
                float acc = 0.f;
                int last_ok = -1;
                nxt = -1;
                for (int e = 0; e < deg; ++e) {
                    const int cc = (int)__ldg(t.col + lo + e);
                    bool seen = false;
                    if (use_fresh) {
                        #pragma unroll
                        for (int k = 0; k < MEMT; ++k) seen |= (hist[k] == cc);
                    }
                    if (seen) continue;
                    float ang = (float)__ldg(t.edge_theta + lo + e) - heading;
                    ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
                    acc += __expf(k2 * ang * ang);
                    last_ok = cc;
                    if (u < acc) { nxt = cc; break; }
                }
                if (nxt < 0) nxt = (last_ok >= 0) ? last_ok : cur;
            }
        }
        #pragma unroll
        for (int k = MEMT - 1; k > 0; --k) hist[k] = (k < mem) ? hist[k - 1] : -2;
        hist[0] = (mem > 0) ? nxt : -2;
        GCP_DASSERT(pos + 1 < vp.L && nxt >= 0 && (uint32_t)nxt < t.N);
        r[++pos] = (G)nxt;
        cur = nxt;
    }
    return pos;
}

// pruneUTurns (:346-357) of one row held in shared memory, warp-cooperative.
__device__ int w_prune(int32_t* r, int len, int lane)
{
    // pruneUTurns :346-357: position i goes away if the waypoint at i recurs at i+1 or
    // i+2, or if i-1 and i+1 hold the same waypoint (flags from the ORIGINAL row)
    {
        int w = 0, carry_prev = -3;
        for (int i0 = 0; i0 < len; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < len) ? r[i] : -6;
            const int n1 = (i + 1 < len) ? r[i + 1] : -4;
            const int n2 = (i + 2 < len) ? r[i + 2] : -5;
            int prev = __shfl_up_sync(FULL, g, 1);
            if (lane == 0) prev = carry_prev;
            carry_prev = __shfl_sync(FULL, g, 31);
            const bool keep = (i < len) && !((g == n1) || (g == n2) || (prev == n1));
            const uint32_t m = __ballot_sync(FULL, keep);
            __syncwarp();
            if (keep) r[w + __popc(m & ((1u << lane) - 1u))] = g;
            w += __popc(m);
            __syncwarp();
        }
        for (int x = w + lane; x < len; x += 32) r[x] = -1;
        len = w;
        __syncwarp();
    }
    return len;
}

// ------------------------------------------------------------------ lowVarSample
// w_i = (1+gamma)^(-fit_i - max(-fit)); normalised; u_m = r + m/M; index = first i whose
// inclusive cumulative weight reaches u_m (gori_tools.py:207-229).  Four small grid-wide
// kernels over tiles of 1024: max, weights + tile sums, cumulative weights, bisection.
constexpr int LV_THREADS = 256;
constexpr int LV_ITEMS = 4;
constexpr int LV_TILE = LV_THREADS * LV_ITEMS;

__device__ __forceinline__ double block_reduce(double v, bool is_max, double* s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(FULL, v, o);
        v = is_max ? fmax(v, t) : v + t;
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double r = s_red[0];
    for (int w = 1; w < (int)blockDim.x / 32; ++w) r = is_max ? fmax(r, s_red[w]) : r + s_red[w];
    return r;
}

__global__ void __launch_bounds__(LV_THREADS)
k_lv_max(const double* __restrict__ fit, uint32_t M, double* __restrict__ tile_max)
{
    __shared__ double s_red[LV_THREADS / 32];
    double mx = -1.0e300;
    for (int k = 0; k < LV_ITEMS; ++k) {
        const uint32_t i = blockIdx.x * LV_TILE + k * LV_THREADS + threadIdx.x;
        if (i < M) mx = fmax(mx, -fit[i]);
    }
    mx = block_reduce(mx, true, s_red);
    if (threadIdx.x == 0) tile_max[blockIdx.x] = mx;
}

__global__ void __launch_bounds__(LV_THREADS)
k_lv_weights(const double* __restrict__ fit, uint32_t M, double gamma, uint32_t n_tiles,
             const double* __restrict__ tile_max, double* __restrict__ cum, double* __restrict__ tile_sum)
{
    __shared__ double s_red[LV_THREADS / 32];
    double mx = -1.0e300;
    for (uint32_t k = threadIdx.x; k < n_tiles; k += LV_THREADS) mx = fmax(mx, tile_max[k]);
    mx = block_reduce(mx, true, s_red);
    const double base = gamma + 1.0;
    const uint32_t b = blockIdx.x * LV_TILE + threadIdx.x * LV_ITEMS;
    double loc = 0.0;
    #pragma unroll
    for (int k = 0; k < LV_ITEMS; ++k)
        if (b + k < M) { const double w = pow(base, -fit[b + k] - mx); cum[b + k] = w; loc += w; }
    loc = block_reduce(loc, false, s_red);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = loc;
}

__global__ void __launch_bounds__(LV_THREADS)
k_lv_cumulate(uint32_t M, uint32_t n_tiles, const double* __restrict__ tile_sum, double* __restrict__ cum)
{
    __shared__ double s_red[LV_THREADS / 32];
    __shared__ double s_w[LV_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double before = 0.0, all = 0.0;
    for (uint32_t k = threadIdx.x; k < n_tiles; k += LV_THREADS) {
        const double v = tile_sum[k];
        all += v;
        if (k < blockIdx.x) before += v;
    }
    before = block_reduce(before, false, s_red);
    const double total = block_reduce(all, false, s_red);
    const uint32_t b = blockIdx.x * LV_TILE + threadIdx.x * LV_ITEMS;
    double w[LV_ITEMS], loc = 0.0;
    #pragma unroll
    for (int k = 0; k < LV_ITEMS; ++k) { w[k] = (b + k < M) ? cum[b + k] : 0.0; loc += w[k]; }
    double inc = loc;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    double run = before + inc - loc;
    for (int w2 = 0; w2 < warp; ++w2) run += s_w[w2];
    #pragma unroll
    for (int k = 0; k < LV_ITEMS; ++k) { run += w[k]; if (b + k < M) cum[b + k] = run / total; }
}

__global__ void k_lv_search(const double* __restrict__ cum, uint32_t M, double r0, int32_t* __restrict__ out)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double u = r0 + (double)m / (double)M;
    uint32_t lo = 0, hi = M - 1;                          // first i with cum[i] >= u
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cum[mid] >= u) hi = mid; else lo = mid + 1; }
    out[m] = (int32_t)lo;
}

// ------------------------------------------------------------------ gene rows in global memory
// Chromosomes live in global memory as int32 or, when every waypoint id fits (N <= 32767), int16
// genes: G.  A row goes through shared memory as int32.  Rows are moved with 16-byte accesses
// when their byte length allows it (L = 200: 400 B int16 / 800 B int32), which is what keeps the
// reads of a PEER GPU's shard over NVLink at full request size.
template <typename G> struct GeneVec;
template <> struct GeneVec<int32_t> { static constexpr int N = 4; };
template <> struct GeneVec<int16_t> { static constexpr int N = 8; };

template <typename G>
__device__ __forceinline__ void w_load_row(const G* __restrict__ g, int32_t* r, int L, int lane)
{
    constexpr int V = GeneVec<G>::N;
    if ((L % V) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        const int4* g4 = reinterpret_cast<const int4*>(g);
        for (int x = lane; x < L / V; x += 32) {
            const int4 v = g4[x];
            if constexpr (V == 4) {
                reinterpret_cast<int4*>(r)[x] = v;
            } else {
                int4 lo, hi;
                lo.x = (int)(short)(v.x & 0xFFFF); lo.y = v.x >> 16;
                lo.z = (int)(short)(v.y & 0xFFFF); lo.w = v.y >> 16;
                hi.x = (int)(short)(v.z & 0xFFFF); hi.y = v.z >> 16;
                hi.z = (int)(short)(v.w & 0xFFFF); hi.w = v.w >> 16;
                reinterpret_cast<int4*>(r)[2 * x] = lo;
                reinterpret_cast<int4*>(r)[2 * x + 1] = hi;
            }
        }
    } else {
        for (int x = lane; x < L; x += 32) r[x] = (int)g[x];
    }
}

// g[x] = x < n ? r[x] : -1 for x < L (addTelomere :335-344)
template <typename G>
__device__ __forceinline__ void w_store_row(G* __restrict__ g, const int32_t* r, int n, int L, int lane)
{
    constexpr int V = GeneVec<G>::N;
    if ((L % V) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        int4* g4 = reinterpret_cast<int4*>(g);
        for (int x = lane; x < L / V; x += 32) {
            int e[V];
            #pragma unroll
            for (int k = 0; k < V; ++k) { const int i = x * V + k; e[k] = (i < n) ? r[i] : -1; }
            int4 v;
            if constexpr (V == 4) { v = make_int4(e[0], e[1], e[2], e[3]); }
            else {
                v.x = (e[0] & 0xFFFF) | (e[1] << 16); v.y = (e[2] & 0xFFFF) | (e[3] << 16);
                v.z = (e[4] & 0xFFFF) | (e[5] << 16); v.w = (e[6] & 0xFFFF) | (e[7] << 16);
            }
            g4[x] = v;
        }
    } else {
        for (int x = lane; x < L; x += 32) g[x] = (G)((x < n) ? r[x] : -1);
    }
}

// Shards of one logical [P_global][A][L] gene array: organism g lives on shard g / per at
// local index g % per.  base[] holds every shard as mapped into THIS process (own shard: plain
// device memory, peers: cudaIpcOpenMemHandle -> NVLink peer loads).  world == 1: a plain array.
struct ShardTab {
    const void* base[GCP_MAX_SHARDS];
    uint32_t world, per;
};

template <typename G>
__device__ __forceinline__ const G* shard_org(const ShardTab& sh, uint32_t g, size_t org_genes)
{
    const uint32_t s = (sh.world == 1) ? 0u : g / sh.per;
    const uint32_t loc = g - s * sh.per;
    GCP_DASSERT(s < sh.world && loc < sh.per);
    return reinterpret_cast<const G*>(sh.base[s]) + (size_t)loc * org_genes;
}

// ------------------------------------------------------------------ crossover (+ fused mutation)
// One warp per (pair k, agent a): parents strong[k] ("self") and strong[k + P/2] ("mate")
// -> children 2k, 2k+1.  Rows live in shared memory: pm | pd | c1 | c2 | match-index heads.
// Pairs [pair_begin, pair_begin + n_pairs) of the P_global / 2 pairs are built by this launch
// (population shards); the children go to children[2 (k - pair_begin)], +1.
constexpr int VARY_WARPS = 4;
constexpr int XO_BUCKETS = 256;
__device__ __forceinline__ int xo_bucket(int gene) { return (int)(((uint32_t)gene * 2654435761u) >> 24); }

__device__ __forceinline__ size_t vary_row_stride(int L) { return (size_t)((L + 31) & ~31); }

template <typename G>
__global__ void __launch_bounds__(VARY_WARPS * 32)
k_crossover(ShardTab parents, const int32_t* __restrict__ strong, uint32_t P, uint32_t pair_begin,
            uint32_t n_pairs, GcpTables t, VaryParams vp, uint64_t seed, uint32_t gen,
            G* __restrict__ children)
{
    extern __shared__ int32_t s_rows[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lid = blockIdx.x * VARY_WARPS + warp;              // local (pair, agent)
    const uint32_t half = P / 2;
    if (lid >= n_pairs * vp.A) return;
    const uint32_t kl = lid / vp.A, a = lid % vp.A;
    const uint32_t k = pair_begin + kl;                               // global pair
    const uint32_t gid = k * vp.A + a;                                // global (pair, agent): Philox key
    const int L = vp.L;
    const size_t RS = vary_row_stride(L);
    int32_t* pm = s_rows + (size_t)warp * (4 * RS + XO_BUCKETS);
    int32_t* pd = pm + RS;
    int32_t* c1 = pd + RS;
    int32_t* c2 = c1 + RS;
    int32_t* s_head = c2 + RS;                                         // [XO_BUCKETS] chain heads of the match index
    const size_t OG = (size_t)vp.A * L;
    const G* gm = shard_org<G>(parents, (uint32_t)strong[k], OG) + (size_t)a * L;
    const G* gd = shard_org<G>(parents, (uint32_t)strong[k + half], OG) + (size_t)a * L;
    w_load_row<G>(gm, pm, L, lane);
    w_load_row<G>(gd, pd, L, lane);
    __syncwarp();
    const int lm = w_valid_len(pm, L, lane), ld = w_valid_len(pd, L, lane);
    Philox rng(seed, gen, STREAM_XOVER, gid + (vp.org_off / 2) * vp.A);   // global (pair, agent) id
    const bool want = rng.uniform() < vp.xover_prob;                   // :212,217
    int sel_i = -1, sel_j = -1;
    if (want) {
        // matchWaypt :311-333: i, j >= 1, equal waypoint that is a crossover candidate,
        // |i - j| < time_thresh; candidates enumerated row-major like np.where.
        // The mate's positions j >= 1 are indexed by waypoint in 256 hash buckets (heads in
        // s_head, chains in c2), so counting the matches of self[i] walks one short chain
        // instead of scanning the whole time window.  cnt[i] -> c1 (scratch until the children
        // are built).
        for (int x = lane; x < XO_BUCKETS; x += 32) s_head[x] = 0;
        __syncwarp();
        for (int j = 1 + lane; j < ld; j += 32)
            c2[j] = atomicExch(&s_head[xo_bucket(pd[j])], j);          // 0 terminates: j >= 1
        __syncwarp();
        int total = 0;
        for (int i0 = 1; i0 < lm; i0 += 32) {
            const int i = i0 + lane;
            int cnt = 0;
            if (i < lm) {
                const int w = pm[i];
                if (__ldg(t.xover + w))
                    for (int p = s_head[xo_bucket(w)]; p; p = c2[p])
                        cnt += (pd[p] == w && abs(i - p) < vp.time_thresh);
                c1[i] = cnt;
            }
            total += cnt;
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
        __syncwarp();
        if (total > 0) {
            int pick = rng.randint(total);                             // np.random.choice :221
            for (int i0 = 1; i0 < lm; i0 += 32) {
                const int i = i0 + lane;
                const int cnt = (i < lm) ? c1[i] : 0;
                int inc = cnt;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += v;
                }
                const uint32_t m = __ballot_sync(FULL, inc > pick);
                if (m) {
                    const int src = __ffs(m) - 1;
                    int jj = -1;
                    if (lane == src) {
                        const int rem = pick - (inc - cnt);                // rem-th match in ascending j
                        const int w = pm[i];
                        const int b = xo_bucket(w);
                        for (int p = s_head[b]; p && jj < 0; p = c2[p]) {
                            if (!(pd[p] == w && abs(i - p) < vp.time_thresh)) continue;
                            int below = 0;
                            for (int q = s_head[b]; q; q = c2[q])
                                below += (q < p && pd[q] == w && abs(i - q) < vp.time_thresh);
                            if (below == rem) jj = p;
                        }
                    }
                    sel_i = i0 + src;
                    sel_j = __shfl_sync(FULL, jj, src);
                    break;
                }
                pick -= __shfl_sync(FULL, inc, 31);
            }
        }
        __syncwarp();
    }
    GCP_DASSERT(sel_i < 0 || (sel_i >= 1 && sel_i < lm && sel_j >= 1 && sel_j < ld && pm[sel_i] == pd[sel_j]));
    int n1, n2;
    if (sel_i >= 0) {
        // dna1 = self[0:i] + mate[j:], dna2 = mate[0:j] + self[i:]  (:225-229), cut at L
        n1 = min(L, sel_i + ld - sel_j);
        n2 = min(L, sel_j + lm - sel_i);
        for (int x = lane; x < n1; x += 32) c1[x] = (x < sel_i) ? pm[x] : pd[sel_j + x - sel_i];
        for (int x = lane; x < n2; x += 32) c2[x] = (x < sel_j) ? pd[x] : pm[sel_i + x - sel_j];
        __syncwarp();
        // padMinDNA :203-209: extend short children with a fresh walk from their last node
        if (n1 > 0 && n1 < vp.min_dna_len) n1 = w_walk(t, vp, c1, n1 - 1, vp.min_dna_len - n1, n1 - 1, rng, lane) + 1;
        if (n2 > 0 && n2 < vp.min_dna_len) n2 = w_walk(t, vp, c2, n2 - 1, vp.min_dna_len - n2, n2 - 1, rng, lane) + 1;
    } else {
        n1 = lm; n2 = ld;
        for (int x = lane; x < n1; x += 32) c1[x] = pm[x];
        for (int x = lane; x < n2; x += 32) c2[x] = pd[x];
    }
    __syncwarp();
    w_store_row<G>(children + ((size_t)(2 * kl) * vp.A + a) * L, c1, n1, L, lane);       // addTelomere :335-344
    w_store_row<G>(children + ((size_t)(2 * kl + 1) * vp.A + a) * L, c2, n2, L, lane);
}

// ------------------------------------------------------------------ mutation (:251-268)
// Organisms flip the mutation coin as a whole (:159-161); every agent row of a mutating
// organism regrows its tail from a random cut.  k_mutation_list compacts the rows that walk
// into a global list (order is irrelevant: every row owns its Philox stream), then
// k_mutation_walks gives each listed row to one THREAD: 32 independent dependent-load
// chains per warp, nothing replicated across lanes, no thread walks twice.
constexpr int MW_THREADS = 64;

// first -1 of a row that is a valid prefix followed by -1 padding, by bisection
template <typename G>
__device__ __forceinline__ int row_len(const G* r, int L)
{
    int lo = 0, hi = L;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (r[mid] == (G)-1) hi = mid; else lo = mid + 1; }
    return lo;
}

// the cut and the tail length of a mutating row (:252-258), from the row's own Philox stream
template <typename G>
__device__ __forceinline__ bool mutation_cut(const G* r, const VaryParams& vp, Philox& rng, int* len_out,
                                             int* idx_out, int* tail_out)
{
    const int L = vp.L;
    const int len = row_len<G>(r, L);
    *len_out = len;
    if (len < 3) return false;
    const int idx = 1 + rng.randint(len - 2);                          // randint(1, len-1)
    int len_tail = len - idx;
    if (len + 1 < L) len_tail += rng.randint(L - len);                 // :256-258
    *idx_out = idx; *tail_out = len_tail;
    return true;
}

// The rows that walk are counting-sorted by tail length (longest first), so that the 32 walks of a
// warp end together: with rows in arbitrary order half of a warp's lanes sit idle behind its
// longest walk.  k_mutation_list emits {row, class} pairs and a histogram of the classes,
// k_mutation_bases scans it, k_mutation_place scatters the rows into the sorted list.
constexpr int MW_BINS = 256;
__device__ __forceinline__ int mw_class(int tail, int L) { return min(MW_BINS - 1, (tail * MW_BINS) / (L + 1)); }

// counter layout (u32): [0] number of listed rows, [1 .. 1+MW_BINS) histogram -> cursors,
// [1+MW_BINS .. 1+2*MW_BINS) class bases.  which = 0: mutation coin (:159-161), pairs = {row, class};
// which = 1: muterpolate coin (:163-165), list = rows, counter[0] = their number.
template <typename G>
__global__ void k_mutation_list(const G* __restrict__ dna, uint32_t P, VaryParams vp, uint64_t seed,
                                uint32_t gen, uint32_t salt, int which, uint32_t* __restrict__ counter,
                                uint32_t* __restrict__ list, uint2* __restrict__ pairs)
{
    const uint32_t org = blockIdx.x * blockDim.x + threadIdx.x;
    if (org >= P) return;
    Philox org_rng(seed, gen, stream_id(STREAM_ORG, salt), org + vp.org_off);
    const float u0 = org_rng.uniform(), u1 = org_rng.uniform();
    if (which == 1) {
        if (u1 < vp.muterp_prob) {
            const uint32_t base = atomicAdd(counter, (uint32_t)vp.A);
            for (int a = 0; a < vp.A; ++a) list[base + a] = org * vp.A + a;
        }
        return;
    }
    if (!(u0 < vp.mut_prob)) return;
    for (int a = 0; a < vp.A; ++a) {
        const uint32_t gid = org * vp.A + a;
        Philox rng(seed, gen, stream_id(STREAM_MUT, salt), gid + vp.org_off * vp.A);
        int len, idx, tail;
        if (!mutation_cut<G>(dna + (size_t)gid * vp.L, vp, rng, &len, &idx, &tail)) continue;
        const int c = mw_class(tail, vp.L);
        GCP_DASSERT(c >= 0 && c < MW_BINS && tail >= 1 && idx >= 1 && idx < len);
        pairs[atomicAdd(counter, 1u)] = make_uint2(gid, (uint32_t)c);
        atomicAdd(&counter[1 + c], 1u);
    }
}

// one CTA of MW_BINS threads: bases of the classes in DESCENDING tail length; cursors restart at 0
__global__ void __launch_bounds__(MW_BINS)
k_mutation_bases(uint32_t* __restrict__ counter)
{
    __shared__ uint32_t s_w[MW_BINS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = MW_BINS - 1 - tid;                                   // thread 0 = longest class
    const uint32_t v = counter[1 + c];
    uint32_t inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += x; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    counter[1 + MW_BINS + c] = off + inc - v;
    counter[1 + c] = 0;
}

__global__ void k_mutation_place(uint32_t* __restrict__ counter, const uint2* __restrict__ pairs,
                                 uint32_t* __restrict__ list)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= counter[0]) return;
    const uint2 pr = pairs[j];
    const uint32_t at = counter[1 + MW_BINS + pr.y] + atomicAdd(&counter[1 + pr.y], 1u);
    GCP_DASSERT(pr.y < (uint32_t)MW_BINS && at < counter[0]);
    list[at] = pr.x;
}

template <typename G, int MEMT>
__global__ void __launch_bounds__(MW_THREADS)
k_mutation_walks(G* __restrict__ dna, GcpTables t, VaryParams vp, uint64_t seed,
                 uint32_t gen, uint32_t salt, const uint32_t* __restrict__ counter,
                 const uint32_t* __restrict__ list)
{
    const uint32_t j = blockIdx.x * MW_THREADS + threadIdx.x;
    if (j >= counter[0]) return;
    const uint32_t gid = list[j];
    G* r = dna + (size_t)gid * vp.L;
    Philox rng(seed, gen, stream_id(STREAM_MUT, salt), gid + vp.org_off * vp.A);
    int len, idx, len_tail;
    if (!mutation_cut<G>(r, vp, rng, &len, &idx, &len_tail)) return;
    const int last = t_walk<G, MEMT>(t, vp, r, idx, len_tail, 0, rng);
    for (int x = last + 1; x < len; ++x) r[x] = (G)-1;                 // addTelomere
}

// ------------------------------------------------------------------ muterpolate (:270-309)
// One THREAD per row of the organisms whose muterpolate coin came up (compacted list), the row
// edited in place in global memory (it stays in L1).  num_muterpolations anchors, searched
// back from idx1 + srch_dist + 1 to idx1 + 2, each tried with probability sub_prob: a direct
// edge pt1 -> pt2 is a no-op (the reference discards its np.delete, :293), otherwise a random
// common out-neighbour replaces idx1 + 1 and idx1 + 2 .. idx2 - 2 are deleted.
__device__ __forceinline__ void t_nbrs(const GcpTables& t, int node, uint32_t* n)
{
    const uint4* rec = reinterpret_cast<const uint4*>(t.adj8 + (size_t)node * 8);
    #pragma unroll
    for (int h = 0; h < 4; ++h) { const uint4 q = __ldg(rec + h); n[2 * h] = q.x; n[2 * h + 1] = q.z; }
}

__device__ __forceinline__ bool t_is_edge_csr(const GcpTables& t, int a, int b)
{
    for (uint32_t e = __ldg(t.row_ptr + a); e < __ldg(t.row_ptr + a + 1); ++e)
        if ((int)__ldg(t.col + e) == b) return true;
    return false;
}

// `spread` (a power of two <= 32): only every spread-th thread takes a row.  The rows of a warp follow
// different control paths (coins, break / continue), which SIMT serialises: a warp's time is the SUM
// of its rows' instruction streams.  With few rows to process the machine has warps to spare, and one
// row per warp finishes 32 x sooner than 32 rows in one warp; the host picks `spread` so that the rows
// still fill the SMs.
// One row: `r` points at the row (shared-memory copy or global memory), gid = organism * A + agent.
// Returns whether the row was changed.
template <typename G>
__device__ __forceinline__ bool muterpolate_row(G* r, uint32_t gid, const GcpTables& t, const VaryParams& vp,
                                                uint64_t seed, uint32_t gen, uint32_t salt)
{
    const int L = vp.L;
    const int len = row_len<G>(r, L);
    const int range = len - (vp.srch_dist + 3);
    if (range <= 0) return false;
    Philox rng(seed, gen, stream_id(STREAM_MUTERP, salt), gid + vp.org_off * vp.A);
    int pts[MAX_MUT_POINTS];
    const int np_ = min(vp.num_muterp, MAX_MUT_POINTS);
    for (int q = 0; q < np_; ++q) {                                    // choice with replacement,
        const int v = rng.randint(range);                              // kept sorted descending
        int x = q;
        while (x > 0 && pts[x - 1] < v) { pts[x] = pts[x - 1]; --x; }
        pts[x] = v;
    }
    const uint32_t sub_thr = (uint32_t)fminf(65536.f, fmaxf(0.f, vp.sub_prob * 65536.f + 0.5f));
    // np.delete shrinks the padded row; anchors come in descending order, so every deletion
    // lies at or left of the previous one: a GAP BUFFER (logical index x lives at x below the gap,
    // at x + gap_len above it) makes each deletion O(distance to the previous one) instead of
    // shifting the whole tail.
    int gap_lo = L, gap_len = 0;                                       // logical length = L - gap_len
    bool changed = false;
    for (int q = 0; q < np_; ++q) {
        const int idx1 = pts[q];
        const int pt1 = (int)r[idx1];                                  // idx1 + 1 < every gap position
        uint32_t n1[8];
        if (t.adj8) t_nbrs(t, pt1, n1);
        for (int idx2 = idx1 + vp.srch_dist + 1; idx2 > idx1 + 1; --idx2) {
            if (!rng.coin(sub_thr)) continue;                          // do_it < P_muterpolate_each :285
            if (idx2 >= L - gap_len) break;                            // IndexError -> break :289
            int pt2 = (int)r[idx2 < gap_lo ? idx2 : idx2 + gap_len];
            if (pt2 < 0) pt2 += (int)t.N;                              // numpy index -1
            int step = -1;
            if (t.adj8) {
                bool direct = false;
                #pragma unroll
                for (int e = 0; e < 8; ++e) direct |= (n1[e] == (uint32_t)pt2);
                if (direct) break;                                     // np.delete discarded :293
                uint32_t n2[8];
                t_nbrs(t, pt2, n2);
                uint32_t cm = 0;                                       // common out-neighbours :298
                #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    bool hit = false;
                    #pragma unroll
                    for (int f = 0; f < 8; ++f) hit |= (n1[e] == n2[f]);
                    if (hit && n1[e] != 0xFFFFFFFFu) cm |= 1u << e;
                }
                if (cm == 0) continue;
                for (int pick = rng.randint(__popc(cm)); pick > 0; --pick) cm &= cm - 1;
                const int e = __ffs(cm) - 1;
                #pragma unroll
                for (int k = 0; k < 8; ++k) if (k == e) step = (int)n1[k];
            } else {
                if (t_is_edge_csr(t, pt1, pt2)) break;
                int nopt = 0;
                for (uint32_t e1 = __ldg(t.row_ptr + pt1); e1 < __ldg(t.row_ptr + pt1 + 1); ++e1)
                    nopt += t_is_edge_csr(t, pt2, (int)__ldg(t.col + e1));
                if (nopt == 0) continue;
                int pick = rng.randint(nopt);
                for (uint32_t e1 = __ldg(t.row_ptr + pt1); e1 < __ldg(t.row_ptr + pt1 + 1); ++e1) {
                    const int c = (int)__ldg(t.col + e1);
                    if (t_is_edge_csr(t, pt2, c)) { if (pick == 0) { step = c; break; } --pick; }
                }
            }
            GCP_DASSERT(idx1 + 1 < gap_lo && step >= 0 && (uint32_t)step < t.N);
            r[idx1 + 1] = (G)step;                                     // :302
            changed = true;
            const int ndel = idx2 - idx1 - 3;                          // delete idx1+2 .. idx2-2
            if (ndel > 0) {
                const int new_lo = idx1 + 2;                           // <= gap_lo always
                GCP_DASSERT(new_lo <= gap_lo && gap_lo - 1 + gap_len < L);
                // move the gap left: four independent reads, then their writes (a move to the right walked
                // from the right end: a block that is read before it is written can never be clobbered)
                int x = gap_lo - 1;
                for (; x - 3 >= new_lo; x -= 4) {
                    const G m0 = r[x], m1 = r[x - 1], m2 = r[x - 2], m3 = r[x - 3];
                    r[x + gap_len] = m0; r[x - 1 + gap_len] = m1; r[x - 2 + gap_len] = m2; r[x - 3 + gap_len] = m3;
                }
                for (; x >= new_lo; --x) r[x + gap_len] = r[x];
                gap_lo = new_lo;
                gap_len += ndel;
            }
            break;
        }
    }
    if (changed) {                                                     // close the gap, strip -1 (:307)
        // below the gap nothing moved unless a -1 hole sits there (w == x: nothing to write); above it the
        // survivors slide left - four reads at a time, their writes land at or below what has been read
        int w = 0, x = 0;
        for (; x < gap_lo; ++x) {
            const G g = r[x];
            if (g != (G)-1) { if (w != x) r[w] = g; ++w; }
        }
        x = gap_lo + gap_len;
        for (; x + 3 < L; x += 4) {
            const G g0 = r[x], g1 = r[x + 1], g2 = r[x + 2], g3 = r[x + 3];
            if (g0 != (G)-1) r[w++] = g0;
            if (g1 != (G)-1) r[w++] = g1;
            if (g2 != (G)-1) r[w++] = g2;
            if (g3 != (G)-1) r[w++] = g3;
        }
        for (; x < L; ++x) {
            const G g = r[x];
            if (g != (G)-1) r[w++] = g;
        }
        for (x = w; x < L; ++x) r[x] = (G)-1;
    }
    return changed;
}

// STAGE: the rows of a warp are copied to shared memory with coalesced loads first, edited there by
// their owner lanes, and the changed ones copied back.  The edit is a serial chain of single-element
// reads and writes per row; in global memory every one of them is an L1 / L2 round trip (long
// scoreboard: 52 % of the stall samples, profiles/ncu r02w), in shared memory 29 cycles.
template <typename G, bool STAGE>
__global__ void __launch_bounds__(MW_THREADS)
k_muterpolate_rows(G* __restrict__ dna, GcpTables t, VaryParams vp, uint64_t seed, uint32_t gen,
                   uint32_t salt, const uint32_t* __restrict__ counter, const uint32_t* __restrict__ list,
                   uint32_t spread, uint32_t stride)
{
    extern __shared__ __align__(16) unsigned char mp_smem[];
    const uint32_t g = blockIdx.x * MW_THREADS + threadIdx.x;
    const uint32_t j = g / spread;
    const bool mine = (g & (spread - 1u)) == 0 && j < *counter;
    const int L = vp.L;
    const uint32_t gid = mine ? list[j] : 0u;
    if (!STAGE) {
        if (mine) muterpolate_row<G>(dna + (size_t)gid * L, gid, t, vp, seed, gen, salt);
        return;
    }
    const int lane = threadIdx.x & 31;
    // slot of a lane's row: rows of the block back to back, `stride` bytes apart (odd number of words:
    // lanes that touch the same index of their rows hit different banks)
    G* slot = reinterpret_cast<G*>(mp_smem + (size_t)(threadIdx.x / spread) * stride);
    const uint32_t have = __ballot_sync(0xffffffffu, mine);
    for (uint32_t m = have; m; m &= m - 1) {                           // copy in, one row at a time, all lanes
        const int src = __ffs(m) - 1;
        const uint32_t sg = __shfl_sync(0xffffffffu, gid, src);
        G* dst = reinterpret_cast<G*>(mp_smem + (size_t)((threadIdx.x - lane + src) / spread) * stride);
        const G* from = dna + (size_t)sg * L;
        for (int x = lane; x < L; x += 32) dst[x] = from[x];
    }
    __syncwarp();
    bool changed = false;
    if (mine) changed = muterpolate_row<G>(slot, gid, t, vp, seed, gen, salt);
    __syncwarp();
    for (uint32_t m = __ballot_sync(0xffffffffu, changed); m; m &= m - 1) {     // copy the changed rows back
        const int src = __ffs(m) - 1;
        const uint32_t sg = __shfl_sync(0xffffffffu, gid, src);
        const G* from = reinterpret_cast<const G*>(mp_smem + (size_t)((threadIdx.x - lane + src) / spread) * stride);
        G* dst = dna + (size_t)sg * L;
        for (int x = lane; x < L; x += 32) dst[x] = from[x];
    }
}

// ------------------------------------------------------------------ pruneUTurns (:346-357)
// One warp per (organism, agent); the row goes through shared memory and back in place.
template <typename G>
__global__ void __launch_bounds__(VARY_WARPS * 32)
k_post_mutate(G* __restrict__ dna, uint32_t P, GcpTables t, VaryParams vp, uint64_t seed,
              uint32_t gen, uint32_t salt)
{
    extern __shared__ int32_t s_rows[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gid = blockIdx.x * VARY_WARPS + warp;
    if (gid >= P * vp.A) return;
    const int L = vp.L;
    const size_t RS = vary_row_stride(L);
    int32_t* r = s_rows + (size_t)warp * RS;
    G* g = dna + (size_t)gid * L;
    w_load_row<G>(g, r, L, lane);
    __syncwarp();
    int len = w_valid_len(r, L, lane);
    len = w_prune(r, len, lane);
    w_store_row<G>(g, r, len, L, lane);
}

// ------------------------------------------------------------------ initial random paths
__global__ void __launch_bounds__(128)
k_random_walks(int32_t* __restrict__ dna, uint32_t P, GcpTables t, VaryParams vp,
               const int32_t* __restrict__ start_idx, int path_len, uint64_t seed, uint32_t batch)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P * vp.A) return;
    int32_t* r = dna + (size_t)gid * vp.L;
    Philox rng(seed, batch, STREAM_INIT, gid);
    r[0] = start_idx[gid % vp.A];
    const int last = t_walk<int32_t, PATH_MEM_MAX>(t, vp, r, 0, path_len, 0, rng);
    for (int x = last + 1; x < vp.L; ++x) r[x] = -1;
}

// ------------------------------------------------------------------ arena truncation
// new_parent[k] = gladiator[order[k]] for k in [k_begin, k_begin + n_local); gladiators =
// parents (0..P) ++ children, each a sharded array (a peer's survivors come over NVLink).
// CTAs past n_local copy the objective / fitness records of ALL P survivors (replicated, 24 B each).
__global__ void k_gather_survivors(const int32_t* __restrict__ order, uint32_t P, uint32_t k_begin,
                                   uint32_t n_local, uint32_t org_bytes, ShardTab par, ShardTab kid,
                                   const double* __restrict__ glad_obj, const double* __restrict__ glad_fit,
                                   void* __restrict__ out_dna, double* __restrict__ out_obj,
                                   double* __restrict__ out_fit)
{
    if (blockIdx.x >= n_local) {
        const uint32_t k = (blockIdx.x - n_local) * blockDim.x + threadIdx.x;
        if (k < P) {
            const uint32_t src = (uint32_t)order[k];
            reinterpret_cast<double2*>(out_obj)[k] = reinterpret_cast<const double2*>(glad_obj)[src];
            out_fit[k] = glad_fit[src];
        }
        return;
    }
    const uint32_t k = k_begin + blockIdx.x;
    const uint32_t src = (uint32_t)order[k];
    GCP_DASSERT(src < 2u * P);
    const ShardTab& sh = (src < P) ? par : kid;
    const uint32_t g = (src < P) ? src : src - P;
    const uint32_t s = (sh.world == 1) ? 0u : g / sh.per;
    const char* sp = reinterpret_cast<const char*>(sh.base[s]) + (size_t)(g - s * sh.per) * org_bytes;
    char* d = reinterpret_cast<char*>(out_dna) + (size_t)blockIdx.x * org_bytes;
    if ((org_bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        const int4* s4 = reinterpret_cast<const int4*>(sp);
        int4* d4 = reinterpret_cast<int4*>(d);
        for (uint32_t x = threadIdx.x; x < org_bytes / 16; x += blockDim.x) d4[x] = s4[x];
    } else {
        for (uint32_t x = threadIdx.x; x < org_bytes; x += blockDim.x) d[x] = sp[x];
    }
}

// ------------------------------------------------------------------ rendezvous over peer memory
// One thread per rank: announce `epoch` in slot `rank` of rank w's flags, then wait for rank w's
// announcement in the own flags.  release / acquire at system scope order the stores of the
// kernels that ran before this one against the peer's loads after its barrier.
__global__ void k_peer_barrier(ShardTab flags, uint32_t rank, uint32_t epoch, uint32_t* __restrict__ timed_out)
{
    const uint32_t w = threadIdx.x;
    if (w >= flags.world) return;
    uint32_t* theirs = reinterpret_cast<uint32_t*>(const_cast<void*>(flags.base[w])) + rank;
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.base[rank]) + w;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(theirs), "r"(epoch) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { if (timed_out) *timed_out = 1u; break; }      // 10 s: a peer died
        __nanosleep(200);
    }
}

// ------------------------------------------------------------------ host launchers
static VaryParams make_vp(gcp_ctx* c)
{
    VaryParams vp;
    vp.A = c->p.num_agents; vp.L = c->p.max_dna_len;
    vp.min_dna_len = c->vary.min_dna_len; vp.time_thresh = c->vary.crossover_time_thresh;
    vp.num_muterp = c->vary.num_muterpolations; vp.srch_dist = c->vary.muterpolation_srch_dist;
    vp.path_memory = c->vary.path_memory;
    vp.xover_prob = (float)c->vary.crossover_prob; vp.mut_prob = (float)c->vary.mutation_prob;
    vp.muterp_prob = (float)c->vary.muterpolate_prob; vp.sub_prob = (float)c->vary.muterpolation_sub_prob;
    vp.inv_sigma = (float)(1.0 / c->vary.log_scale_weight);
    vp.org_off = c->vary_org_offset;
    return vp;
}

static int make_tab(gcp_ctx* c, const gcp_shards* s, ShardTab& tab)
{
    if (!s || s->world < 1 || s->world > GCP_MAX_SHARDS) return gcp_fail(c, -1, "shard table: world must be in [1,%d]", GCP_MAX_SHARDS);
    if (s->world > 1 && s->per_shard == 0) return gcp_fail(c, -1, "shard table: per_shard is 0");
    tab.world = s->world; tab.per = s->per_shard;
    for (uint32_t w = 0; w < GCP_MAX_SHARDS; ++w) tab.base[w] = (w < s->world) ? s->base[w] : nullptr;
    for (uint32_t w = 0; w < s->world; ++w)
        if (!tab.base[w]) return gcp_fail(c, -1, "shard table: base[%u] is NULL", w);
    return 0;
}

int launch_low_var_select(gcp_ctx* c, const double* fit, uint32_t M, double gamma, double r0,
                          double* scratch /*dev double [M + 2 * ceil(M/1024)]*/, int32_t* out,
                          cudaStream_t st)
{
    if (M == 0) return 0;
    StageTimer tm(c, 5, st);
    const uint32_t nt = (M + LV_TILE - 1) / LV_TILE;
    double* cum = scratch;
    double* tile_max = scratch + M;
    double* tile_sum = tile_max + nt;
    k_lv_max<<<nt, LV_THREADS, 0, st>>>(fit, M, tile_max);
    k_lv_weights<<<nt, LV_THREADS, 0, st>>>(fit, M, gamma, nt, tile_max, cum, tile_sum);
    k_lv_cumulate<<<nt, LV_THREADS, 0, st>>>(M, nt, tile_sum, cum);
    k_lv_search<<<(M + 255) / 256, 256, 0, st>>>(cum, M, r0, out);
    c->launches += 4;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

static int vary_smem(gcp_ctx* c, int rows, int extra, size_t* out)
{
    const size_t RS = (size_t)((c->p.max_dna_len + 31) & ~31);
    *out = (size_t)VARY_WARPS * (rows * RS + extra) * 4;
    if (*out > (size_t)c->max_smem_optin)
        return gcp_fail(c, -3, "variation kernel needs %zu B shared memory (max_dna_len too large)", *out);
    return 0;
}

static int check_gene_bytes(gcp_ctx* c, int gene_bytes)
{
    if (gene_bytes != 2 && gene_bytes != 4) return gcp_fail(c, -1, "gene_bytes must be 2 (int16) or 4 (int32)");
    if (gene_bytes == 2 && c->t.N > 32767) return gcp_fail(c, -1, "int16 genes need N <= 32767 waypoints (N=%u)", c->t.N);
    return 0;
}

template <typename G>
static int mutate_impl(gcp_ctx* c, G* dna, uint32_t P, uint32_t org_offset, uint64_t seed, uint32_t gen,
                       uint32_t salt, cudaStream_t st)
{
    size_t smem; int r = vary_smem(c, 1, 0, &smem); if (r) return r;
    GCP_CUDA(c, cudaFuncSetAttribute(k_post_mutate<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(c, 7, st);
    VaryParams vp = make_vp(c);
    vp.org_off += org_offset;
    const uint32_t n_rows = P * (uint32_t)vp.A;
    if (vp.mut_prob > 0.f || vp.muterp_prob > 0.f) {
        const size_t need = 4096 + (size_t)n_rows * 12;           // counters | sorted list | {row, class} pairs
        if (c->vary_scratch_bytes < need) {
            if (c->vary_scratch) cudaFree(c->vary_scratch);
            c->vary_scratch = nullptr; c->vary_scratch_bytes = 0;
            GCP_CUDA(c, cudaMalloc(&c->vary_scratch, need));
            c->vary_scratch_bytes = need;
        }
    }
    uint32_t* counter = reinterpret_cast<uint32_t*>(c->vary_scratch);      // [1 + 2 * MW_BINS] + muterpolate counter at [768]
    uint2* pairs = reinterpret_cast<uint2*>(counter + 1024);
    uint32_t* list = reinterpret_cast<uint32_t*>(pairs + n_rows);
    const uint32_t row_grid = (n_rows + MW_THREADS - 1) / MW_THREADS;
    if (vp.mut_prob > 0.f) {                       // mutation (:159-161): tail regrowth walks
        GCP_CUDA(c, cudaMemsetAsync(counter, 0, 4 * (1 + MW_BINS), st));
        k_mutation_list<G><<<(P + 255) / 256, 256, 0, st>>>(dna, P, vp, seed, gen, salt, 0, counter, list, pairs);
        k_mutation_bases<<<1, MW_BINS, 0, st>>>(counter);
        k_mutation_place<<<(n_rows + 255) / 256, 256, 0, st>>>(counter, pairs, list);
        if (vp.path_memory <= 10)
            k_mutation_walks<G, 10><<<row_grid, MW_THREADS, 0, st>>>(dna, c->t, vp, seed, gen, salt, counter, list);
        else
            k_mutation_walks<G, PATH_MEM_MAX><<<row_grid, MW_THREADS, 0, st>>>(dna, c->t, vp, seed, gen, salt, counter, list);
        c->launches += 4;
    }
    if (vp.muterp_prob > 0.f && vp.num_muterp > 0) {   // muterpolate (:163-165)
        GCP_CUDA(c, cudaMemsetAsync(counter + 768, 0, 4, st));
        k_mutation_list<G><<<(P + 255) / 256, 256, 0, st>>>(dna, P, vp, seed, gen, salt, 1, counter + 768, list, pairs);
        // expected rows = organisms x coin probability x agents; warps available = SMs x 48
        const double rows_est = (double)P * vp.muterp_prob * vp.A + 64.0;
        uint32_t spread = 32;
        while (spread > 1 && rows_est * spread > (double)c->num_sms * 48.0 * 32.0) spread >>= 1;
        const uint32_t mgrid = (uint32_t)(((uint64_t)n_rows * spread + MW_THREADS - 1) / MW_THREADS);
        // rows staged in shared memory while they fit (stride: row bytes rounded up to an odd number of words)
        const uint32_t stride = ((((uint32_t)vp.L * (uint32_t)sizeof(G) + 3u) / 4u) | 1u) * 4u;
        const size_t mp_smem = (size_t)(MW_THREADS / spread) * stride;
        if (mp_smem <= 96u * 1024u) {
            GCP_CUDA(c, cudaFuncSetAttribute(k_muterpolate_rows<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mp_smem));
            k_muterpolate_rows<G, true><<<mgrid, MW_THREADS, mp_smem, st>>>(dna, c->t, vp, seed, gen, salt, counter + 768, list, spread, stride);
        } else {
            k_muterpolate_rows<G, false><<<mgrid, MW_THREADS, 0, st>>>(dna, c->t, vp, seed, gen, salt, counter + 768, list, spread, stride);
        }
        c->launches += 2;
    }
    k_post_mutate<G><<<(n_rows + VARY_WARPS - 1) / VARY_WARPS, VARY_WARPS * 32, smem, st>>>(dna, P, c->t, vp,
                                                                                            seed, gen, salt);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_mutate_genes(gcp_ctx* c, void* dna, int gene_bytes, uint32_t P, uint32_t org_offset, uint64_t seed,
                        uint32_t gen, uint32_t salt, cudaStream_t st)
{
    if (P == 0) return 0;
    int r = check_gene_bytes(c, gene_bytes); if (r) return r;
    if (gene_bytes == 2) return mutate_impl<int16_t>(c, reinterpret_cast<int16_t*>(dna), P, org_offset, seed, gen, salt, st);
    return mutate_impl<int32_t>(c, reinterpret_cast<int32_t*>(dna), P, org_offset, seed, gen, salt, st);
}

int launch_mutate(gcp_ctx* c, int32_t* dna, uint32_t P, uint64_t seed, uint32_t gen, uint32_t salt,
                  cudaStream_t st)
{
    return launch_mutate_genes(c, dna, 4, P, 0, seed, gen, salt, st);
}

template <typename G>
static int crossover_impl(gcp_ctx* c, const ShardTab& tab, const int32_t* strong, uint32_t P, uint32_t pair_begin,
                          uint32_t n_pairs, uint64_t seed, uint32_t gen, G* children, cudaStream_t st)
{
    size_t smem; int r = vary_smem(c, 4, XO_BUCKETS, &smem); if (r) return r;
    GCP_CUDA(c, cudaFuncSetAttribute(k_crossover<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(c, 6, st);
    const uint32_t n = n_pairs * c->p.num_agents;
    k_crossover<G><<<(n + VARY_WARPS - 1) / VARY_WARPS, VARY_WARPS * 32, smem, st>>>(
        tab, strong, P, pair_begin, n_pairs, c->t, make_vp(c), seed, gen, children);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_crossover_shards(gcp_ctx* c, const gcp_shards* parents, int gene_bytes, const int32_t* strong, uint32_t P,
                            uint32_t pair_begin, uint32_t n_pairs, uint64_t seed, uint32_t gen, void* children,
                            cudaStream_t st)
{
    if (n_pairs == 0) return 0;
    if (P & 1) return gcp_fail(c, -1, "population size must be even (pather_driver.py:107)");
    if ((uint64_t)pair_begin + n_pairs > P / 2) return gcp_fail(c, -1, "pair range exceeds P / 2");
    int r = check_gene_bytes(c, gene_bytes); if (r) return r;
    ShardTab tab;
    if ((r = make_tab(c, parents, tab))) return r;
    if (gene_bytes == 2) return crossover_impl<int16_t>(c, tab, strong, P, pair_begin, n_pairs, seed, gen,
                                                        reinterpret_cast<int16_t*>(children), st);
    return crossover_impl<int32_t>(c, tab, strong, P, pair_begin, n_pairs, seed, gen,
                                   reinterpret_cast<int32_t*>(children), st);
}

int launch_crossover(gcp_ctx* c, const int32_t* parents, const int32_t* strong, uint32_t P,
                     uint64_t seed, uint32_t gen, int32_t* children, cudaStream_t st)
{
    if (P < 2) return 0;
    gcp_shards one;
    memset(&one, 0, sizeof(one));
    one.world = 1; one.per_shard = P; one.base[0] = parents;
    // `strong` holds P entries even when only the pairs of one population shard are built here
    // (vary_org_offset keys the Philox streams by global ids)
    return launch_crossover_shards(c, &one, 4, strong, P, 0, P / 2, seed, gen, children, st);
}

int launch_random_walks(gcp_ctx* c, int32_t* dna, uint32_t P, const int32_t* start_idx_dev,
                        int path_len, uint64_t seed, uint32_t batch, cudaStream_t st)
{
    if (P == 0) return 0;
    const uint32_t n = P * c->p.num_agents;
    k_random_walks<<<(n + 127) / 128, 128, 0, st>>>(dna, P, c->t, make_vp(c), start_idx_dev, path_len,
                                                   seed, batch);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_gather_shards(gcp_ctx* c, const int32_t* order, uint32_t P, uint32_t k_begin, uint32_t n_local,
                         const gcp_shards* parents, const gcp_shards* children, int gene_bytes,
                         const double* glad_obj, const double* glad_fit, void* out_dna, double* out_obj,
                         double* out_fit, cudaStream_t st)
{
    if (P == 0) return 0;
    if ((uint64_t)k_begin + n_local > P) return gcp_fail(c, -1, "survivor range exceeds the population");
    int r = check_gene_bytes(c, gene_bytes); if (r) return r;
    ShardTab pt, kt;
    if ((r = make_tab(c, parents, pt)) || (r = make_tab(c, children, kt))) return r;
    const uint32_t org_bytes = (uint32_t)(c->p.num_agents * c->p.max_dna_len * gene_bytes);
    const uint32_t rec_blocks = (out_obj && out_fit) ? (P + 127) / 128 : 0;
    k_gather_survivors<<<n_local + rec_blocks, 128, 0, st>>>(order, P, k_begin, n_local, org_bytes, pt, kt,
                                                             glad_obj, glad_fit, out_dna, out_obj, out_fit);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_peer_barrier(gcp_ctx* c, const gcp_shards* flags, uint32_t rank, uint32_t epoch, uint32_t* timed_out,
                        cudaStream_t st)
{
    ShardTab tab;
    int r = make_tab(c, flags, tab); if (r) return r;
    if (rank >= tab.world) return gcp_fail(c, -1, "peer barrier: rank %u of %u", rank, tab.world);
    if (tab.world == 1) return 0;
    k_peer_barrier<<<1, 32, 0, st>>>(tab, rank, epoch, timed_out);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_gather_survivors(gcp_ctx* c, const int32_t* order, uint32_t P, const int32_t* par_dna,
                            const int32_t* kid_dna, const double* glad_obj, const double* glad_fit,
                            int32_t* out_dna, double* out_obj, double* out_fit, cudaStream_t st)
{
    gcp_shards par, kid;
    memset(&par, 0, sizeof(par)); memset(&kid, 0, sizeof(kid));
    par.world = kid.world = 1; par.per_shard = kid.per_shard = P;
    par.base[0] = par_dna; kid.base[0] = kid_dna;
    return launch_gather_shards(c, order, P, 0, P, &par, &kid, 4, glad_obj, glad_fit, out_dna, out_obj, out_fit, st);
}
