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

#define FULL 0xffffffffu
constexpr int VARY_THREADS = 128;
constexpr int MAX_MUT_POINTS = 64;
constexpr int PATH_MEM_MAX = 16;

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
    __device__ double uniform64()
    {
        const uint64_t a = next(), b = next();
        return (double)(((a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ int randint(int n) { return n <= 1 ? 0 : min(n - 1, (int)(uniform() * (float)n)); }
};

enum { STREAM_ORG = 1, STREAM_XOVER = 2, STREAM_MUT = 3, STREAM_INIT = 4 };

struct VaryParams {
    int A, L;
    int min_dna_len, time_thresh, num_muterp, srch_dist, path_memory;
    float xover_prob, mut_prob, muterp_prob, sub_prob, inv_sigma;
};

__device__ __forceinline__ int valid_len(const int32_t* r, int L)
{
    int n = 0;
    while (n < L && r[n] != -1) ++n;
    return n;
}

__device__ __forceinline__ bool is_edge(const GcpTables& t, int a, int b)
{
    for (uint32_t e = t.row_ptr[a]; e < t.row_ptr[a + 1]; ++e)
        if ((int)t.col[e] == b) return true;
    return false;
}

// pathmaker.py:188-213.  r[pos] is the current node; appends up to n_steps nodes (stops at
// L).  The no-return memory is the last `path_memory` nodes of the running path, which
// never reaches below row index mem_floor.  Returns the index of the last node written.
__device__ int walk(const GcpTables& t, const VaryParams& vp, int32_t* r, int pos, int n_steps,
                    int mem_floor, Philox& rng)
{
    const float PI = 3.14159265358979f;
    const float heading = rng.uniform() * 2.0f * PI - PI;            // :190, never updated
    int hist[PATH_MEM_MAX];
    const int mem = min(vp.path_memory, PATH_MEM_MAX);
    #pragma unroll
    for (int k = 0; k < PATH_MEM_MAX; ++k) {
        const int idx = pos - k;
        hist[k] = (k < mem && idx >= mem_floor) ? r[idx] : -2;
    }
    int cur = r[pos];
    for (int s = 0; s < n_steps && pos + 1 < vp.L; ++s) {
        const uint32_t lo = t.row_ptr[cur], hi = t.row_ptr[cur + 1];
        int nxt = cur;
        if (hi > lo) {
            float tot_all = 0.f, tot_fresh = 0.f;
            for (uint32_t e = lo; e < hi; ++e) {
                float ang = (float)t.edge_theta[e] - heading;        // assignWeights :179-181
                ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
                const float w = __expf(-0.5f * ang * ang * vp.inv_sigma * vp.inv_sigma);
                const int c = (int)t.col[e];
                bool seen = false;
                #pragma unroll
                for (int k = 0; k < PATH_MEM_MAX; ++k) seen |= (hist[k] == c);
                tot_all += w;
                if (!seen) tot_fresh += w;
            }
            const bool use_fresh = tot_fresh > 0.f;                  // :199-207
            const float u = rng.uniform() * (use_fresh ? tot_fresh : tot_all);
            float acc = 0.f;
            nxt = -1;
            int last_ok = -1;
            for (uint32_t e = lo; e < hi; ++e) {
                const int c = (int)t.col[e];
                bool seen = false;
                if (use_fresh) {
                    #pragma unroll
                    for (int k = 0; k < PATH_MEM_MAX; ++k) seen |= (hist[k] == c);
                }
                if (seen) continue;
                float ang = (float)t.edge_theta[e] - heading;
                ang -= 2.0f * PI * floorf((ang + PI) * (0.5f / PI));
                acc += __expf(-0.5f * ang * ang * vp.inv_sigma * vp.inv_sigma);
                last_ok = c;
                if (u < acc) { nxt = c; break; }
            }
            if (nxt < 0) nxt = (last_ok >= 0) ? last_ok : cur;
        }
        #pragma unroll
        for (int k = PATH_MEM_MAX - 1; k > 0; --k) hist[k] = (k < mem) ? hist[k - 1] : -2;
        hist[0] = nxt;
        r[++pos] = nxt;
        cur = nxt;
    }
    return pos;
}

// ------------------------------------------------------------------ lowVarSample
// w_i = (1+gamma)^(-fit_i - max(-fit)); normalised; u_m = r + m/M; index = first i whose
// inclusive cumulative weight reaches u_m.  One CTA (M <= a few 10^5).
__global__ void __launch_bounds__(1024)
k_low_var_select(const double* __restrict__ fit, uint32_t M, double gamma, double r0,
                 double* __restrict__ cum /*scratch [M]*/, int32_t* __restrict__ out)
{
    __shared__ double s_red[32];
    __shared__ double s_val;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // max(-fit)
    double mx = -1.0e300;
    for (uint32_t i = tid; i < M; i += 1024) mx = fmax(mx, -fit[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) { double m = s_red[0]; for (int w = 1; w < 32; ++w) m = fmax(m, s_red[w]); s_val = m; }
    __syncthreads();
    mx = s_val;
    const double base = gamma + 1.0;
    // contiguous chunk per thread -> in-order cumulative sum
    const uint32_t per = (M + 1023) / 1024;
    const uint32_t b = tid * per, e = min(M, b + per);
    double loc = 0.0;
    for (uint32_t i = b; i < e; ++i) { const double w = pow(base, -fit[i] - mx); cum[i] = w; loc += w; }
    double inc = loc;
    for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
    __syncthreads();
    if (lane == 31) s_red[warp] = inc;
    __syncthreads();
    double woff = 0.0, total = 0.0;
    for (int w = 0; w < 32; ++w) { if (w < warp) woff += s_red[w]; total += s_red[w]; }
    double run = woff + inc - loc;
    for (uint32_t i = b; i < e; ++i) { run += cum[i]; cum[i] = run / total; }
    __syncthreads();
    for (uint32_t m = tid; m < M; m += 1024) {
        const double u = r0 + (double)m / (double)M;
        uint32_t lo = 0, hi = M - 1;                      // first i with cum[i] >= u
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cum[mid] >= u) hi = mid; else lo = mid + 1; }
        out[m] = (int32_t)lo;
    }
}

// ------------------------------------------------------------------ crossover
// One thread per (pair k, agent a): parents strong[k] ("self") and strong[k + P/2] ("mate").
__global__ void __launch_bounds__(VARY_THREADS)
k_crossover(const int32_t* __restrict__ parents, const int32_t* __restrict__ strong, uint32_t P,
            GcpTables t, VaryParams vp, uint64_t seed, uint32_t gen, int32_t* __restrict__ children)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t half = P / 2;
    if (gid >= half * vp.A) return;
    const uint32_t k = gid / vp.A, a = gid % vp.A;
    const int L = vp.L;
    const int32_t* pm = parents + ((size_t)strong[k] * vp.A + a) * L;
    const int32_t* pd = parents + ((size_t)strong[k + half] * vp.A + a) * L;
    int32_t* c1 = children + ((size_t)(2 * k) * vp.A + a) * L;
    int32_t* c2 = children + ((size_t)(2 * k + 1) * vp.A + a) * L;
    const int lm = valid_len(pm, L), ld = valid_len(pd, L);
    Philox rng(seed, gen, STREAM_XOVER, gid);
    const bool want = rng.uniform() < vp.xover_prob;                   // :212,217
    int sel_i = -1, sel_j = -1;
    if (want) {
        // matchWaypt :311-333: i, j >= 1, equal waypoint that is a crossover candidate,
        // |i - j| < time_thresh; candidates enumerated row-major like np.where
        int cnt = 0;
        for (int i = 1; i < lm; ++i) {
            const int w = pm[i];
            if (!t.xover[w]) continue;
            const int j0 = max(1, i - vp.time_thresh + 1), j1 = min(ld - 1, i + vp.time_thresh - 1);
            for (int j = j0; j <= j1; ++j) cnt += (pd[j] == w);
        }
        if (cnt > 0) {
            int pick = rng.randint(cnt);                               // np.random.choice :221
            for (int i = 1; i < lm && sel_i < 0; ++i) {
                const int w = pm[i];
                if (!t.xover[w]) continue;
                const int j0 = max(1, i - vp.time_thresh + 1), j1 = min(ld - 1, i + vp.time_thresh - 1);
                for (int j = j0; j <= j1; ++j)
                    if (pd[j] == w) { if (pick == 0) { sel_i = i; sel_j = j; break; } --pick; }
            }
        }
    }
    int n1, n2;
    if (sel_i >= 0) {
        // dna1 = self[0:i] + mate[j:], dna2 = mate[0:j] + self[i:]  (:225-229), cut at L
        n1 = 0;
        for (int x = 0; x < sel_i && n1 < L; ++x) c1[n1++] = pm[x];
        for (int x = sel_j; x < ld && n1 < L; ++x) c1[n1++] = pd[x];
        n2 = 0;
        for (int x = 0; x < sel_j && n2 < L; ++x) c2[n2++] = pd[x];
        for (int x = sel_i; x < lm && n2 < L; ++x) c2[n2++] = pm[x];
        // padMinDNA :203-209: extend short children with a fresh walk from their last node
        if (n1 > 0 && n1 < vp.min_dna_len) n1 = walk(t, vp, c1, n1 - 1, vp.min_dna_len - n1, n1 - 1, rng) + 1;
        if (n2 > 0 && n2 < vp.min_dna_len) n2 = walk(t, vp, c2, n2 - 1, vp.min_dna_len - n2, n2 - 1, rng) + 1;
    } else {
        for (n1 = 0; n1 < lm; ++n1) c1[n1] = pm[n1];
        for (n2 = 0; n2 < ld; ++n2) c2[n2] = pd[n2];
    }
    for (int x = n1; x < L; ++x) c1[x] = -1;                           // addTelomere :335-344
    for (int x = n2; x < L; ++x) c2[x] = -1;
}

// ------------------------------------------------------------------ mutation / muterpolate / prune
// One thread per (organism, agent); works in place on the organism's row.
__global__ void __launch_bounds__(VARY_THREADS)
k_mutate(int32_t* __restrict__ dna, uint32_t P, GcpTables t, VaryParams vp, uint64_t seed,
         uint32_t gen, uint32_t stream_salt)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P * vp.A) return;
    const uint32_t org = gid / vp.A;
    const int L = vp.L;
    int32_t* r = dna + (size_t)gid * L;
    int len = valid_len(r, L);
    // organism-level coin flips, identical for all agents of the organism (:159-165)
    Philox org_rng(seed, gen, STREAM_ORG + stream_salt, org);
    const bool do_mut = org_rng.uniform() < vp.mut_prob;
    const bool do_mp = org_rng.uniform() < vp.muterp_prob;
    Philox rng(seed, gen, STREAM_MUT + stream_salt, gid);

    if (do_mut && len >= 3) {                                          // mutation :251-268
        const int idx = 1 + rng.randint(len - 2);                      // randint(1, len-1)
        int len_tail = len - idx;
        if (len + 1 < L) len_tail += rng.randint(L - len);             // :256-258
        len = walk(t, vp, r, idx, len_tail, 0, rng) + 1;
        for (int x = len; x < L; ++x) r[x] = -1;
    }
    if (do_mp && len - (vp.srch_dist + 3) > 0) {                       // muterpolate :270-309
        int pts[MAX_MUT_POINTS];
        const int np_ = min(vp.num_muterp, MAX_MUT_POINTS);
        const int range = len - (vp.srch_dist + 3);
        for (int q = 0; q < np_; ++q) {                                // choice with replacement,
            const int v = rng.randint(range);                          // kept sorted descending
            int x = q;
            while (x > 0 && pts[x - 1] < v) { pts[x] = pts[x - 1]; --x; }
            pts[x] = v;
        }
        int arr_len = L;                                               // np.delete shrinks the padded row
        for (int q = 0; q < np_; ++q) {
            const int idx1 = pts[q];
            for (int idx2 = idx1 + vp.srch_dist + 1; idx2 > idx1 + 1; --idx2) {
                if (!(rng.uniform() < vp.sub_prob)) continue;
                if (idx2 >= arr_len) break;                            // IndexError -> break :289
                const int pt1 = r[idx1];
                int pt2 = r[idx2];
                if (pt2 < 0) pt2 += (int)t.N;                          // numpy index -1
                if (is_edge(t, pt1, pt2)) break;                       // np.delete discarded :293
                int nopt = 0;                                          // common out-neighbours :298
                for (uint32_t e1 = t.row_ptr[pt1]; e1 < t.row_ptr[pt1 + 1]; ++e1)
                    nopt += is_edge(t, pt2, (int)t.col[e1]);
                if (nopt == 0) continue;
                int pick = rng.randint(nopt), step = -1;
                for (uint32_t e1 = t.row_ptr[pt1]; e1 < t.row_ptr[pt1 + 1]; ++e1) {
                    const int c = (int)t.col[e1];
                    if (is_edge(t, pt2, c)) { if (pick == 0) { step = c; break; } --pick; }
                }
                r[idx1 + 1] = step;                                    // :302
                const int ndel = idx2 - idx1 - 3;                      // delete idx1+2 .. idx2-2
                if (ndel > 0) {
                    for (int x = idx1 + 2; x + ndel < arr_len; ++x) r[x] = r[x + ndel];
                    for (int x = arr_len - ndel; x < arr_len; ++x) r[x] = -1;
                    arr_len -= ndel;
                }
                break;
            }
        }
        int w = 0;                                                     // strip -1 :307
        for (int x = 0; x < L; ++x) { const int g = r[x]; if (g != -1) r[w++] = g; }
        len = w;
        for (int x = len; x < L; ++x) r[x] = -1;
    }
    // pruneUTurns :346-357: position i goes away if the waypoint at i recurs at i+1 or
    // i+2, or if i-1 and i+1 hold the same waypoint (one pass over the original row)
    {
        int w = 0, prev = -3;
        for (int i = 0; i < len; ++i) {
            const int g = r[i];
            const int n1 = (i + 1 < len) ? r[i + 1] : -4;
            const int n2 = (i + 2 < len) ? r[i + 2] : -5;
            const bool gone = (g == n1) || (g == n2) || (prev == n1);
            prev = g;
            if (!gone) r[w++] = g;
        }
        for (int x = w; x < len; ++x) r[x] = -1;
        len = w;
    }
}

// ------------------------------------------------------------------ initial random paths
__global__ void __launch_bounds__(VARY_THREADS)
k_random_walks(int32_t* __restrict__ dna, uint32_t P, GcpTables t, VaryParams vp,
               const int32_t* __restrict__ start_idx, int path_len, uint64_t seed, uint32_t batch)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P * vp.A) return;
    int32_t* r = dna + (size_t)gid * vp.L;
    Philox rng(seed, batch, STREAM_INIT, gid);
    r[0] = start_idx[gid % vp.A];
    const int last = walk(t, vp, r, 0, path_len, 0, rng);
    for (int x = last + 1; x < vp.L; ++x) r[x] = -1;
}

// ------------------------------------------------------------------ arena truncation
// new_parent[k] = gladiator[order[k]] for k < P; gladiators = parents (0..P) ++ children.
__global__ void k_gather_survivors(const int32_t* __restrict__ order, uint32_t P, uint32_t row_ints,
                                   const int32_t* __restrict__ par_dna, const int32_t* __restrict__ kid_dna,
                                   const double* __restrict__ glad_obj, const double* __restrict__ glad_fit,
                                   int32_t* __restrict__ out_dna, double* __restrict__ out_obj,
                                   double* __restrict__ out_fit)
{
    const uint32_t k = blockIdx.x;
    if (k >= P) return;
    const uint32_t src = (uint32_t)order[k];
    const int32_t* s = (src < P) ? par_dna + (size_t)src * row_ints : kid_dna + (size_t)(src - P) * row_ints;
    int32_t* d = out_dna + (size_t)k * row_ints;
    for (uint32_t x = threadIdx.x; x < row_ints; x += blockDim.x) d[x] = s[x];
    if (threadIdx.x == 0) {
        out_obj[2 * (size_t)k] = glad_obj[2 * (size_t)src];
        out_obj[2 * (size_t)k + 1] = glad_obj[2 * (size_t)src + 1];
        out_fit[k] = glad_fit[src];
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
    return vp;
}

int launch_low_var_select(gcp_ctx* c, const double* fit, uint32_t M, double gamma, double r0,
                          double* scratch, int32_t* out, cudaStream_t st)
{
    if (M == 0) return 0;
    StageTimer tm(c, 5, st);
    k_low_var_select<<<1, 1024, 0, st>>>(fit, M, gamma, r0, scratch, out);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_crossover(gcp_ctx* c, const int32_t* parents, const int32_t* strong, uint32_t P,
                     uint64_t seed, uint32_t gen, int32_t* children, cudaStream_t st)
{
    if (P < 2) return 0;
    if (P & 1) return gcp_fail(c, -1, "population size must be even (pather_driver.py:107)");
    StageTimer tm(c, 6, st);
    const uint32_t n = (P / 2) * c->p.num_agents;
    k_crossover<<<(n + VARY_THREADS - 1) / VARY_THREADS, VARY_THREADS, 0, st>>>(
        parents, strong, P, c->t, make_vp(c), seed, gen, children);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_mutate(gcp_ctx* c, int32_t* dna, uint32_t P, uint64_t seed, uint32_t gen, uint32_t salt,
                  cudaStream_t st)
{
    if (P == 0) return 0;
    StageTimer tm(c, 7, st);
    const uint32_t n = P * c->p.num_agents;
    k_mutate<<<(n + VARY_THREADS - 1) / VARY_THREADS, VARY_THREADS, 0, st>>>(dna, P, c->t, make_vp(c),
                                                                            seed, gen, salt);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_random_walks(gcp_ctx* c, int32_t* dna, uint32_t P, const int32_t* start_idx_dev,
                        int path_len, uint64_t seed, uint32_t batch, cudaStream_t st)
{
    if (P == 0) return 0;
    const uint32_t n = P * c->p.num_agents;
    k_random_walks<<<(n + VARY_THREADS - 1) / VARY_THREADS, VARY_THREADS, 0, st>>>(
        dna, P, c->t, make_vp(c), start_idx_dev, path_len, seed, batch);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}

int launch_gather_survivors(gcp_ctx* c, const int32_t* order, uint32_t P, const int32_t* par_dna,
                            const int32_t* kid_dna, const double* glad_obj, const double* glad_fit,
                            int32_t* out_dna, double* out_obj, double* out_fit, cudaStream_t st)
{
    if (P == 0) return 0;
    const uint32_t row = (uint32_t)(c->p.num_agents * c->p.max_dna_len);
    k_gather_survivors<<<P, 128, 0, st>>>(order, P, row, par_dna, kid_dna, glad_obj, glad_fit,
                                          out_dna, out_obj, out_fit);
    c->launches++;
    GCP_CUDA(c, cudaGetLastError());
    return 0;
}
