// Loop closures for genomes too long for the shared-memory hash of k_loop_closures
// (A * L beyond ~7,000 genes: BASELINE config C4 with 16 / 32 agents x 640 genes).
//
// Same semantics (mappy.py:372-437), different machinery: the directed
// edge keys of a CHUNK of organisms are sorted at once with the stable LSD radix sort of
// gcp_sort.cu, key = organism | from | to, value = position | owner | split flag.  Stability
// keeps positions ascending inside a group of equal keys, so the "consecutive ascending gap"
// test of the canonical order is a comparison of neighbours.  One thread per group start walks
// its group and adds to the organism's A x A matrix; a last kernel applies the constraints and
// assembles the objectives (geneticalgorithm.py:177-201).
#include "gcp_common.cuh"
#include "gcp_device.cuh"

size_t radix_hist_bytes(uint32_t n);
int radix_sort_pairs_bits(gcp_ctx* ctx, unsigned long long* keys, uint32_t* vals, uint32_t n,
                          unsigned long long* keys_tmp, uint32_t* vals_tmp, uint32_t* hist,
                          int n_bits, cudaStream_t st, unsigned long long** out_keys,
                          uint32_t** out_vals);

constexpr int LB_THREADS = 256;
constexpr int LB_NW = LB_THREADS / 32;
constexpr unsigned long long LB_PAD = ~0ull;

// One CTA per organism of the chunk: compaction of the valid genes (row-major), then
// key_i = (wps[i], wps[(i+1) mod n]) incl. boundary / wrap edges (mappy.py:390-391).
__global__ void __launch_bounds__(LB_THREADS)
k_lcb_keys(const int32_t* __restrict__ dna, uint32_t org0, int A, int L, int nb /*bits per node id*/,
           int32_t* __restrict__ seq /*[chunk][A*L]*/, unsigned long long* __restrict__ keys,
           uint32_t* __restrict__ vals)
{
    __shared__ int s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t o = blockIdx.x;
    const int AL = A * L;
    const int32_t* d = dna + (size_t)(org0 + o) * AL;
    int32_t* sq = seq + (size_t)o * AL;
    unsigned long long* ky = keys + (size_t)o * AL;
    uint32_t* vl = vals + (size_t)o * AL;
    for (int a = warp; a < A; a += LB_NW) {
        int cnt = 0;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            cnt += __popc(__ballot_sync(FULL, (i < L) && (d[a * L + i] != -1)));
        }
        if (lane == 0) s_len[a] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int b = 0;
        for (int a = 0; a < A; ++a) { s_base[a] = b; b += s_len[a]; }
        s_base[A] = b;
    }
    __syncthreads();
    const int n = s_base[A];
    for (int a = warp; a < A; a += LB_NW) {
        int run = s_base[a];
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? d[a * L + i] : -1;
            const bool v = g != -1;
            const uint32_t bal = __ballot_sync(FULL, v);
            if (v) {
                const int pos = run + __popc(bal & ((1u << lane) - 1u));
                sq[pos] = g;
                vl[pos] = (uint32_t)pos | ((uint32_t)a << 16) |
                          ((a >= 1 && pos == s_base[a]) ? (1u << 21) : 0u);   // path split (F10)
            }
            run += __popc(bal);
        }
    }
    __syncthreads();                      // the CTA's own global writes are visible to it
    const unsigned long long mask = (1ull << nb) - 1ull;
    for (int i = tid; i < AL; i += LB_THREADS) {
        if (i < n) {
            const unsigned long long f = (unsigned long long)(uint32_t)sq[i] & mask;
            const unsigned long long g = (unsigned long long)(uint32_t)sq[(i + 1 == n) ? 0 : i + 1] & mask;
            ky[i] = ((unsigned long long)o << (2 * nb)) | (f << nb) | g;
        } else {
            ky[i] = LB_PAD;               // sorts behind every real key
            vl[i] = 0;
        }
    }
}

// One thread per sorted item; group starts walk their group.
__global__ void __launch_bounds__(LB_THREADS)
k_lcb_groups(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t M,
             uint32_t org0, int A, int nb, int thresh, int32_t* __restrict__ lc /*[P][A][A]*/)
{
    const uint32_t j = blockIdx.x * LB_THREADS + threadIdx.x;
    if (j >= M) return;
    const unsigned long long key = keys[j];
    if (key == LB_PAD) return;
    if (j > 0 && keys[j - 1] == key) return;            // not the first member of its group
    uint32_t owners = 0, cnt = 0;
    bool split = false, gap = false;
    uint32_t prev_pos = 0;
    for (uint32_t k = j; k < M && keys[k] == key; ++k) {
        const uint32_t v = vals[k];
        const uint32_t pos = v & 0xFFFFu;
        owners |= 1u << ((v >> 16) & 31u);
        split |= ((v >> 21) & 1u) != 0;
        if (cnt > 0 && (int)(pos - prev_pos) > thresh) gap = true;   // ascending: stable sort
        prev_pos = pos;
        ++cnt;
    }
    if (cnt < 2 || split) return;
    int32_t* m = lc + (size_t)(org0 + (uint32_t)(key >> (2 * nb))) * A * A;
    if (__popc(owners) == 1) {
        if (gap) { const int a = __ffs(owners) - 1; atomicAdd(&m[a * A + a], 1); }
    } else {
        for (uint32_t mx = owners; mx; mx &= mx - 1) {
            const int x = __ffs(mx) - 1;
            for (uint32_t my = mx & (mx - 1); my; my &= my - 1)
                atomicAdd(&m[x * A + (__ffs(my) - 1)], 1);
        }
    }
}

// constraints + objectives from the finished matrices, one warp per organism
__global__ void __launch_bounds__(32)
k_lcb_finish(uint32_t P, int A, gcp_params prm, gcp_map_consts mc, const int32_t* __restrict__ lc,
             const double* __restrict__ dist, const double* __restrict__ turn,
             const uint32_t* __restrict__ cov, double* __restrict__ obj,
             double* __restrict__ neg_cov_out, uint8_t* __restrict__ flags)
{
    __shared__ int s_lc[GCP_MAX_AGENTS * GCP_MAX_AGENTS];
    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    for (int i = threadIdx.x; i < A * A; i += 32) s_lc[i] = lc[(size_t)org * A * A + i];
    __syncwarp();
    if (threadIdx.x == 0)
        finish_objectives(s_lc, A, prm, mc, cov + (size_t)org * 8, dist + (size_t)org * A,
                          turn + (size_t)org * A, org, obj, neg_cov_out, flags);
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static int bits_for(uint32_t x) { int b = 1; while (b < 32 && (1u << b) <= x) ++b; return b; }

int launch_loop_closures_big(gcp_ctx* ctx, const int32_t* dna, uint32_t P, const double* dist,
                             const double* turn, const uint32_t* cov, int32_t* lc_out, double* obj,
                             double* neg_cov, uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    const int A = ctx->p.num_agents, L = ctx->p.max_dna_len;
    const size_t AL = (size_t)A * L;
    const int nb = bits_for(ctx->t.N - 1);
    // chunk of organisms sorted at once: <= 16 Mi items, organism index must fit the key
    uint32_t chunk = (uint32_t)((16u << 20) / AL);
    if (chunk < 1) chunk = 1;
    if (chunk > P) chunk = P;
    while (bits_for(chunk - 1) + 2 * nb > 63 && chunk > 1) chunk >>= 1;
    const size_t M = (size_t)chunk * AL;
    const size_t need = 2 * al256(M * 8) + 2 * al256(M * 4) + al256(M * 4) +
                        al256(radix_hist_bytes((uint32_t)M)) + al256((size_t)P * A * A * 4) + 256;
    if (ctx->lcb_scratch_bytes < need) {
        if (ctx->lcb_scratch) cudaFree(ctx->lcb_scratch);
        ctx->lcb_scratch = nullptr; ctx->lcb_scratch_bytes = 0;
        GCP_CUDA(ctx, cudaMalloc(&ctx->lcb_scratch, need));
        ctx->lcb_scratch_bytes = need;
    }
    char* p = reinterpret_cast<char*>(ctx->lcb_scratch);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(p);     p += al256(M * 8);
    unsigned long long* keys_tmp = reinterpret_cast<unsigned long long*>(p); p += al256(M * 8);
    uint32_t* vals = reinterpret_cast<uint32_t*>(p);                         p += al256(M * 4);
    uint32_t* vals_tmp = reinterpret_cast<uint32_t*>(p);                     p += al256(M * 4);
    int32_t* seq = reinterpret_cast<int32_t*>(p);                            p += al256(M * 4);
    uint32_t* hist = reinterpret_cast<uint32_t*>(p);                         p += al256(radix_hist_bytes((uint32_t)M));
    int32_t* lc = lc_out ? lc_out : reinterpret_cast<int32_t*>(p);
    StageTimer tm(ctx, 2, st);
    GCP_CUDA(ctx, cudaMemsetAsync(lc, 0, (size_t)P * A * A * 4, st));
    for (uint32_t o0 = 0; o0 < P; o0 += chunk) {
        const uint32_t n_org = (P - o0 < chunk) ? (P - o0) : chunk;
        const uint32_t m = (uint32_t)((size_t)n_org * AL);
        k_lcb_keys<<<n_org, LB_THREADS, 0, st>>>(dna, o0, A, L, nb, seq, keys, vals);
        unsigned long long* sk = nullptr; uint32_t* sv = nullptr;
        // pad keys are all ones: sorting on the bits real keys use keeps them behind only if
        // the bit just above is included
        const int n_bits = bits_for(n_org - 1) + 2 * nb + 1;
        int r = radix_sort_pairs_bits(ctx, keys, vals, m, keys_tmp, vals_tmp, hist, n_bits, st, &sk, &sv);
        if (r) return r;
        k_lcb_groups<<<(m + LB_THREADS - 1) / LB_THREADS, LB_THREADS, 0, st>>>(
            sk, sv, m, o0, A, nb, ctx->p.solo_sep_thresh, lc);
        ctx->launches += 2;
    }
    k_lcb_finish<<<P, 32, 0, st>>>(P, A, ctx->p, ctx->mc, lc, dist, turn, cov, obj, neg_cov, flags);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
