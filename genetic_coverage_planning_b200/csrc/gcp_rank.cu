// Ranking kernels: maximin fitness (GeneticalGorithm.rackNstack, geneticalgorithm.py:86-123)
// and the arena order (got.sortBy, gori_tools.py:232-236, canonical stable order).
//
// fit_i = max_{j != i} min(sc0_i - sc0_j, sc1_i - sc1_j) in fp64: subtraction is the only
// rounding operation and min/max are selections, so the tiled evaluation below is
// bit-identical to the reference's double loop.  Bound by the FP64 pipe (4 ops / pair),
// not by memory: the j tile is broadcast from shared memory.
#include "gcp_common.cuh"
#include <math_constants.h>

constexpr int RK_THREADS = 256;
constexpr int RK_IPT = 4;                 // rows per thread
constexpr int RK_ROWS = RK_THREADS * RK_IPT;
constexpr int RK_JT = 512;                // j tile staged in shared memory

// sc = [o0, o1 + 1000 if o0 > constr]  (geneticalgorithm.py:101-107)
__global__ void k_scale_obj(const double* __restrict__ obj, uint32_t n, double constr,
                            double2* __restrict__ sc, double* __restrict__ sc_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double o0 = obj[2 * (size_t)i], o1 = obj[2 * (size_t)i + 1];
    const double s1 = (o0 > constr) ? __dadd_rn(o1, 1000.0) : o1;
    sc[i] = make_double2(o0, s1);
    if (sc_out) { sc_out[2 * (size_t)i] = o0; sc_out[2 * (size_t)i + 1] = s1; }
}

__global__ void __launch_bounds__(RK_THREADS)
k_maximin(const double2* __restrict__ sc, uint32_t n, uint32_t row_begin, uint32_t row_end,
          uint32_t j_per_cta, double* __restrict__ partial /*[gridDim.y][row_end-row_begin]*/)
{
    __shared__ double2 s_j[RK_JT];
    const uint32_t rows = row_end - row_begin;
    const uint32_t i0 = row_begin + blockIdx.x * RK_ROWS + threadIdx.x;
    double2 me[RK_IPT];
    double best[RK_IPT];
    #pragma unroll
    for (int r = 0; r < RK_IPT; ++r) {
        const uint32_t i = i0 + r * RK_THREADS;
        me[r] = (i < row_end) ? sc[i] : make_double2(0.0, 0.0);
        best[r] = -CUDART_INF;
    }
    const uint32_t jb = blockIdx.y * j_per_cta;
    const uint32_t je = min(n, jb + j_per_cta);
    for (uint32_t j0 = jb; j0 < je; j0 += RK_JT) {
        const uint32_t cnt = min((uint32_t)RK_JT, je - j0);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < cnt; k += RK_THREADS) s_j[k] = sc[j0 + k];
        __syncthreads();
        // does this tile contain any of my rows?  (uniform per thread, rare)
        bool diag = false;
        #pragma unroll
        for (int r = 0; r < RK_IPT; ++r) {
            const uint32_t i = i0 + r * RK_THREADS;
            diag |= (i >= j0 && i < j0 + cnt);
        }
        if (!diag) {
            #pragma unroll 4
            for (uint32_t k = 0; k < cnt; ++k) {
                const double2 o = s_j[k];
                #pragma unroll
                for (int r = 0; r < RK_IPT; ++r) {
                    const double m = fmin(__dsub_rn(me[r].x, o.x), __dsub_rn(me[r].y, o.y));
                    best[r] = fmax(best[r], m);
                }
            }
        } else {
            for (uint32_t k = 0; k < cnt; ++k) {
                const double2 o = s_j[k];
                #pragma unroll
                for (int r = 0; r < RK_IPT; ++r) {
                    const uint32_t i = i0 + r * RK_THREADS;
                    const double m = fmin(__dsub_rn(me[r].x, o.x), __dsub_rn(me[r].y, o.y));
                    if (j0 + k != i) best[r] = fmax(best[r], m);
                }
            }
        }
    }
    #pragma unroll
    for (int r = 0; r < RK_IPT; ++r) {
        const uint32_t i = i0 + r * RK_THREADS;
        if (i < row_end) partial[(size_t)blockIdx.y * rows + (i - row_begin)] = best[r];
    }
}

__global__ void k_maximin_reduce(const double* __restrict__ partial, uint32_t rows, uint32_t js,
                                 uint32_t row_begin, double* __restrict__ fit)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double b = partial[i];
    for (uint32_t s = 1; s < js; ++s) b = fmax(b, partial[(size_t)s * rows + i]);
    fit[row_begin + i] = b;
}

static void rank_grid(gcp_ctx* ctx, uint32_t rows, uint32_t n, uint32_t* itiles, uint32_t* js,
                      uint32_t* j_per_cta)
{
    *itiles = (rows + RK_ROWS - 1) / RK_ROWS;
    uint32_t want = (uint32_t)(4 * ctx->num_sms + *itiles - 1) / *itiles;   // ~4 CTAs per SM
    uint32_t maxjs = (n + RK_JT - 1) / RK_JT;
    if (want < 1) want = 1;
    if (want > maxjs) want = maxjs;
    if (want > 256) want = 256;
    uint32_t per = (n + want - 1) / want;
    per = ((per + RK_JT - 1) / RK_JT) * RK_JT;
    *j_per_cta = per;
    *js = (n + per - 1) / per;
}

size_t sort_scratch_bytes(uint32_t n);
int launch_maximin_sweep(gcp_ctx* ctx, const double2* sc, uint32_t n, uint32_t r0, uint32_t r1,
                         double* fit, void* scratch, cudaStream_t st);
int launch_arena_order_sort(gcp_ctx* ctx, const double* fit, uint32_t n, int32_t* order,
                            uint8_t* nondom, void* scratch, cudaStream_t st);

static inline size_t sc_bytes(uint32_t n) { return (((size_t)n * 16) + 255) & ~(size_t)255; }

// rank_algo: 0 auto (sort-based for n >= 1024), 1 dominance-matrix tiles (O(N^2)), 2 sort-based
static bool use_sort(gcp_ctx* ctx, uint32_t n)
{
    return ctx->rank_algo == 2 || (ctx->rank_algo == 0 && n >= 1024);
}

size_t rank_scratch_bytes(gcp_ctx* ctx, uint32_t n)
{
    // sc [n] double2 | partial [256][n] worst case is too big: bound by the grid choice
    uint32_t itiles, js, per;
    rank_grid(ctx, n, n, &itiles, &js, &per);
    size_t a = (size_t)n * 16;                 // sc
    size_t b = (size_t)js * n * 8;             // maximin partials (also reused as u64 keys)
    if (b < (size_t)n * 8) b = (size_t)n * 8;
    size_t c = (size_t)n * 4;                  // ranks
    size_t tiles = a + b + c + 1024;
    size_t sorted = sc_bytes(n) + sort_scratch_bytes(n);
    return tiles > sorted ? tiles : sorted;
}

int launch_maximin(gcp_ctx* ctx, const double* obj, uint32_t n, double constr, uint32_t r0,
                   uint32_t r1, double* obj_sc, double* fit, void* scratch, size_t sb,
                   cudaStream_t st)
{
    if (n == 0 || r1 <= r0) return 0;
    if (r1 > n) return gcp_fail(ctx, -1, "maximin: row range [%u,%u) outside n=%u", r0, r1, n);
    if (sb < rank_scratch_bytes(ctx, n)) return gcp_fail(ctx, -1, "maximin: scratch too small");
    StageTimer tm(ctx, 3, st);
    double2* sc = reinterpret_cast<double2*>(scratch);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + (size_t)n * 16);
    k_scale_obj<<<(n + 255) / 256, 256, 0, st>>>(obj, n, constr, sc, obj_sc);
    if (use_sort(ctx, n)) {
        ctx->launches++;
        return launch_maximin_sweep(ctx, sc, n, r0, r1, fit,
                                    reinterpret_cast<char*>(scratch) + sc_bytes(n), st);
    }
    uint32_t itiles, js, per;
    const uint32_t rows = r1 - r0;
    rank_grid(ctx, rows, n, &itiles, &js, &per);
    k_maximin<<<dim3(itiles, js), RK_THREADS, 0, st>>>(sc, n, r0, r1, per, partial);
    k_maximin_reduce<<<(rows + 255) / 256, 256, 0, st>>>(partial, rows, js, r0, fit);
    ctx->launches += 3;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------
// arena order: rank_i = #{ j : (fit_j, j) < (fit_i, i) } by counting (exact, stable,
// deterministic); order[rank_i] = i.  Keys are the order-preserving u64 image of fit.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_key(double f)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    if (b == 0x8000000000000000ull) b = 0;                 // -0.0 == +0.0 for argsort
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_fit_keys(const double* __restrict__ fit, uint32_t n,
                           unsigned long long* __restrict__ keys, uint32_t* __restrict__ rank,
                           uint8_t* __restrict__ nondom)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double f = fit[i];
    keys[i] = f64_key(f);
    rank[i] = 0;
    if (nondom) nondom[i] = f < 0.0 ? 1 : 0;
}

__global__ void __launch_bounds__(RK_THREADS)
k_rank_count(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t j_per_cta,
             uint32_t* __restrict__ rank)
{
    __shared__ unsigned long long s_j[RK_JT];
    const uint32_t i0 = blockIdx.x * RK_ROWS + threadIdx.x;
    unsigned long long me[RK_IPT];
    uint32_t cnt_lt[RK_IPT];
    #pragma unroll
    for (int r = 0; r < RK_IPT; ++r) {
        const uint32_t i = i0 + r * RK_THREADS;
        me[r] = (i < n) ? keys[i] : 0ull;
        cnt_lt[r] = 0;
    }
    const uint32_t jb = blockIdx.y * j_per_cta;
    const uint32_t je = min(n, jb + j_per_cta);
    for (uint32_t j0 = jb; j0 < je; j0 += RK_JT) {
        const uint32_t cnt = min((uint32_t)RK_JT, je - j0);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < cnt; k += RK_THREADS) s_j[k] = keys[j0 + k];
        __syncthreads();
        #pragma unroll 4
        for (uint32_t k = 0; k < cnt; ++k) {
            const unsigned long long o = s_j[k];
            const uint32_t j = j0 + k;
            #pragma unroll
            for (int r = 0; r < RK_IPT; ++r) {
                const uint32_t i = i0 + r * RK_THREADS;
                cnt_lt[r] += (o < me[r]) || (o == me[r] && j < i);
            }
        }
    }
    #pragma unroll
    for (int r = 0; r < RK_IPT; ++r) {
        const uint32_t i = i0 + r * RK_THREADS;
        if (i < n && cnt_lt[r]) atomicAdd(&rank[i], cnt_lt[r]);
    }
}

__global__ void k_rank_scatter(const uint32_t* __restrict__ rank, uint32_t n,
                               int32_t* __restrict__ order)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[rank[i]] = (int32_t)i;
}

int launch_arena_order(gcp_ctx* ctx, const double* fit, uint32_t n, int32_t* order,
                       uint8_t* nondom, void* scratch, size_t sb, cudaStream_t st)
{
    if (n == 0) return 0;
    if (sb < rank_scratch_bytes(ctx, n)) return gcp_fail(ctx, -1, "arena_order: scratch too small");
    StageTimer tm(ctx, 4, st);
    if (use_sort(ctx, n))
        return launch_arena_order_sort(ctx, fit, n, order, nondom,
                                       reinterpret_cast<char*>(scratch) + sc_bytes(n), st);
    unsigned long long* keys =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + (size_t)n * 16);
    uint32_t itiles, js, per;
    rank_grid(ctx, n, n, &itiles, &js, &per);
    uint32_t* rank = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(keys) +
                                                 ((size_t)js * n * 8 > (size_t)n * 8 ? (size_t)js * n * 8 : (size_t)n * 8));
    k_fit_keys<<<(n + 255) / 256, 256, 0, st>>>(fit, n, keys, rank, nondom);
    k_rank_count<<<dim3(itiles, js), RK_THREADS, 0, st>>>(keys, n, per, rank);
    k_rank_scatter<<<(n + 255) / 256, 256, 0, st>>>(rank, n, order);
    ctx->launches += 3;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
