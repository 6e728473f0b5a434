// Sort-based ranking kernels (sm_100a):
//   * stable LSD radix sort of (u64 key, u32 value) pairs           -> arena order, a-order
//   * exact O(N log N) maximin fitness for two objectives            -> k_maximin_sweep
//
// Reference semantics (citations into /root/reference/src):
//   geneticalgorithm.py:86-123  rackNstack: fit_i = max_{j!=i} min(sc0_i-sc0_j, sc1_i-sc1_j)
//   gori_tools.py:232-236       sortBy -> argsort(fit), canonical stable order
//
// Why the sweep is bit-exact: fp64 subtraction is monotone in its
// second operand, so with gladiators sorted by a = sc0 ascending, x_(k) = rn(a_i - a_(k))
// is non-increasing in k and Y(k) = rn(b_i - min_{j<=k, j!=i} b_(j)) is non-decreasing;
// max_j min(x_j, y_j) = max_{k != pos(i)} min(x_(k), Y(k)) and the maximum sits next to
// the crossing, found by bisection.  Only differences of actual elements are evaluated.
#include "gcp_common.cuh"
#include <math_constants.h>

#define FULL 0xffffffffu

constexpr int RS_THREADS = 256;
constexpr int RS_NW = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                          // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 2048 keys per CTA
constexpr int RS_BINS = 256;

__device__ __forceinline__ unsigned long long f64_sort_key(double f)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    if (b == 0x8000000000000000ull) b = 0;                 // -0.0 == +0.0
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// keys[i] = order-preserving image of src[i*stride]; vals[i] = i
__global__ void k_rs_make_keys(const double* __restrict__ src, uint32_t stride, uint32_t n,
                               unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                               uint8_t* __restrict__ neg_flag /*optional: src < 0*/)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double f = src[(size_t)i * stride];
    keys[i] = f64_sort_key(f);
    vals[i] = i;
    if (neg_flag) neg_flag[i] = f < 0.0 ? 1 : 0;
}

// Warp w of a CTA owns keys [tile0 + w*256, +256) in 8 rounds of 32 consecutive keys, so
// (CTA, warp, round, lane) order == input order: ranks assigned in that order are stable.
// With `done` (a zeroed u32 ticket counter) the LAST CTA to finish also turns the histogram into
// the exclusive offsets k_rs_scatter reads (what k_rs_scan_bins + k_rs_scan_totals do as two more
// launches): for small populations a pass is then two launches instead of four, and the ranking of a
// small population is launch-latency bound.
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
          uint32_t* __restrict__ hist /*[256][nblk]*/, uint32_t nblk, uint32_t* __restrict__ done)
{
    __shared__ uint32_t s_h[RS_BINS];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x;
    s_h[tid] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + tid;
        if (i < n) atomicAdd(&s_h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)tid * nblk + blockIdx.x] = s_h[tid];
    if (!done) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done, 1u) == nblk - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // thread = bin: exclusive scan of its row over the tiles, then of the row totals over the bins
    uint32_t* row = hist + (size_t)tid * nblk;
    uint32_t run = 0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += 16) {           // 16 independent loads in flight per thread
        uint32_t v[16];
        #pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (b0 + k < nblk) ? __ldcg(row + b0 + k) : 0u;
        #pragma unroll
        for (int k = 0; k < 16; ++k) { if (b0 + k < nblk) row[b0 + k] = run; run += v[k]; }
    }
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t inc = run;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += x; }
    __shared__ uint32_t s_w[RS_BINS / 32];
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    hist[(size_t)RS_BINS * nblk + tid] = off + inc - run;
    if (tid == 0) *done = 0u;                               // ready for the next pass
}

// Exclusive scan of hist in (bin, block) order in two steps: one CTA per bin scans its row
// of nblk counts in place and records the row total at hist[256 * nblk + bin]; a 256-thread CTA
// then turns the 256 totals into bin bases.
__global__ void __launch_bounds__(1024)
k_rs_scan_bins(uint32_t* __restrict__ hist, uint32_t nblk)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry, s_tot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* row = hist + (size_t)blockIdx.x * nblk;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < nblk; c0 += 1024 * 4) {
        uint32_t v[4], loc = 0;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t i = c0 + tid * 4 + k;
            v[k] = (i < nblk) ? row[i] : 0u;
            loc += v[k];
        }
        uint32_t inc = loc;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_w[lane];
            uint32_t winc = w;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, winc, o);
                if (lane >= o) winc += t;
            }
            s_w[lane] = winc - w;                       // exclusive warp offsets
            if (lane == 31) s_tot = winc;
        }
        __syncthreads();
        uint32_t run = s_carry + s_w[warp] + inc - loc;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t i = c0 + tid * 4 + k;
            if (i < nblk) row[i] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == 0) s_carry += s_tot;
        __syncthreads();
    }
    if (tid == 0) hist[(size_t)RS_BINS * nblk + blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(RS_BINS)
k_rs_scan_totals(uint32_t* __restrict__ bin_total)
{
    __shared__ uint32_t s_w[RS_BINS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t v = bin_total[tid];
    uint32_t inc = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    bin_total[tid] = off + inc - v;
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
             uint32_t n, int shift, const uint32_t* __restrict__ hist, uint32_t nblk,
             unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    __shared__ uint32_t s_cnt[RS_NW][RS_BINS];          // per-warp digit counts -> bases
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < RS_NW * RS_BINS; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    unsigned long long k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? keys[i] : ~0ull;
        v[r] = ok ? vals[i] : 0u;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        const uint32_t act = __ballot_sync(FULL, ok);
        if (ok) {
            const uint32_t peers = __match_any_sync(act, d);
            if ((peers & lt) == 0) s_cnt[warp][d] += __popc(peers);   // leader, warp-private row
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit `tid`: exclusive scan over the warps + global base of (digit, CTA)
        uint32_t run = hist[(size_t)RS_BINS * nblk + tid] + hist[(size_t)tid * nblk + blockIdx.x];
        #pragma unroll
        for (int w = 0; w < RS_NW; ++w) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        const uint32_t act = __ballot_sync(FULL, ok);
        uint32_t peers = 0;
        if (ok) {
            peers = __match_any_sync(act, d);
            const uint32_t pos = s_cnt[warp][d] + __popc(peers & lt);
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
        __syncwarp();
        if (ok && (peers & lt) == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// One launch per pass ("onesweep"): the digit histograms of ALL passes come from one upfront kernel
// (the multiset of keys does not change between passes), and k_os_scatter finds the number of keys
// with its digit in the tiles in front of it by DECOUPLED LOOK-BACK: every tile publishes
// {count | flag} per digit - first its own count (AGG), then the inclusive prefix (PRE) - and walks
// back over its predecessors until it meets a prefix.  Tiles are numbered by an atomic ticket, so a
// tile only ever waits for tiles that started before it.  A sort of 131,072 pairs is 1 + 8 launches
// instead of 32; the ranking of a population is bound by launch latency, not by bandwidth.
// ---------------------------------------------------------------------------------
constexpr uint32_t OS_AGG = 1u << 30, OS_PRE = 1u << 31, OS_VAL = (1u << 30) - 1u;
constexpr uint32_t RS_ONESWEEP_MAX_KEYS = 1u << 20;        // status words: passes * tiles * 256 * 4 B (4 MiB here)
constexpr int OS_MAX_PASSES = 8;

// ghist [OS_MAX_PASSES][256] zeroed by the caller -> exclusive digit bases per pass; `done` zeroed ticket
__global__ void __launch_bounds__(RS_THREADS)
k_os_hist(const unsigned long long* __restrict__ keys, uint32_t n, int passes,
          uint32_t* __restrict__ ghist, uint32_t* __restrict__ done, uint32_t nblk)
{
    __shared__ uint32_t s_h[OS_MAX_PASSES][RS_BINS];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x;
    for (int p = 0; p < passes; ++p) s_h[p][tid] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + tid;
        if (i < n) {
            const unsigned long long k = keys[i];
            for (int p = 0; p < passes; ++p) atomicAdd(&s_h[p][(uint32_t)(k >> (8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) { const uint32_t c = s_h[p][tid]; if (c) atomicAdd(&ghist[p * RS_BINS + tid], c); }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done, 1u) == nblk - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: counts -> exclusive bases, one pass after the other (thread = digit)
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_w[RS_BINS / 32];
    for (int p = 0; p < passes; ++p) {
        const uint32_t v = __ldcg(ghist + p * RS_BINS + tid);
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        uint32_t off = 0;
        for (int w = 0; w < warp; ++w) off += s_w[w];
        ghist[p * RS_BINS + tid] = off + inc - v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS)
k_os_scatter(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
             uint32_t n, int shift, const uint32_t* __restrict__ gbase /*[256] of this pass*/,
             uint32_t* status /*[nblk][256] of this pass, zeroed*/, uint32_t* ticket /*zeroed*/,
             unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    __shared__ uint32_t s_cnt[RS_NW][RS_BINS];          // per-warp digit counts -> bases
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_NW * RS_BINS; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t wbase = tile * RS_TILE + warp * (RS_ITEMS * 32);
    unsigned long long k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? keys[i] : ~0ull;
        v[r] = ok ? vals[i] : 0u;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        const uint32_t act = __ballot_sync(FULL, ok);
        if (ok) {
            const uint32_t peers = __match_any_sync(act, d);
            if ((peers & lt) == 0) s_cnt[warp][d] += __popc(peers);   // leader, warp-private row
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit `tid`: tile total, look-back over the tiles in front, then the per-warp bases
        uint32_t total = 0;
        #pragma unroll
        for (int w = 0; w < RS_NW; ++w) total += s_cnt[w][tid];
        volatile uint32_t* st = status;
        uint32_t excl = 0;
        if (tile == 0) {
            st[tid] = total | OS_PRE;
        } else {
            st[(size_t)tile * RS_BINS + tid] = total | OS_AGG;
            for (uint32_t t = tile; t-- > 0;) {
                uint32_t w, spins = 0;
                while (((w = st[(size_t)t * RS_BINS + tid]) & (OS_AGG | OS_PRE)) == 0)
                    if (++spins > (1u << 28)) __trap();                // a predecessor never published: fail loudly, never hang
                excl += w & OS_VAL;
                if (w & OS_PRE) break;
            }
            st[(size_t)tile * RS_BINS + tid] = (excl + total) | OS_PRE;
        }
        uint32_t run = gbase[tid] + excl;
        #pragma unroll
        for (int w = 0; w < RS_NW; ++w) {
            const uint32_t c = s_cnt[w][tid];
            s_cnt[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    #pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        const uint32_t act = __ballot_sync(FULL, ok);
        uint32_t peers = 0;
        if (ok) {
            peers = __match_any_sync(act, d);
            const uint32_t pos = s_cnt[warp][d] + __popc(peers & lt);
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
        __syncwarp();
        if (ok && (peers & lt) == 0) s_cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
}

static inline uint32_t rs_blocks(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// hist scratch: (256 * nblk + 256) u32
static inline size_t onesweep_words(uint32_t n)
{   // [digit bases: 8 x 256][tickets: 16][status: 8 passes x tiles x 256]
    return (size_t)OS_MAX_PASSES * RS_BINS + 16 + (size_t)OS_MAX_PASSES * rs_blocks(n) * RS_BINS;
}
size_t radix_hist_bytes(uint32_t n)
{
    const size_t staged = ((size_t)rs_blocks(n) * RS_BINS + RS_BINS) * 4 + 256;
    const size_t one = n <= RS_ONESWEEP_MAX_KEYS ? onesweep_words(n) * 4 + 256 : 0;
    return one > staged ? one : staged;
}
// the single last CTA scans 256 rows of nblk counts: a win up to 16 tiles (32,768 keys: 0.39 -> 0.30 ms
// for maximin + order at 16,384 gladiators), a loss from 64 tiles on, where the two scan kernels stay
constexpr uint32_t RS_FUSED_SCAN_MAX_TILES = 16;

// Stable LSD radix sort of (keys, vals) on the low n_bits of the keys, ping-pong through the
// tmp buffers; *out_keys / *out_vals say where the sorted data ends up.
int radix_sort_pairs_bits(gcp_ctx* ctx, unsigned long long* keys, uint32_t* vals, uint32_t n,
                          unsigned long long* keys_tmp, uint32_t* vals_tmp, uint32_t* hist,
                          int n_bits, cudaStream_t st, unsigned long long** out_keys,
                          uint32_t** out_vals)
{
    const uint32_t nblk = rs_blocks(n);
    unsigned long long* kin = keys; unsigned long long* kout = keys_tmp;
    uint32_t* vin = vals; uint32_t* vout = vals_tmp;
    const int passes = (n_bits + 7) / 8;
    if (ctx->sort_onesweep && n <= RS_ONESWEEP_MAX_KEYS && passes <= OS_MAX_PASSES && n > 0) {
        uint32_t* gbase = hist;
        uint32_t* tickets = hist + OS_MAX_PASSES * RS_BINS;               // [0] histogram kernel, [1 + pass] scatter
        uint32_t* status = tickets + 16;
        GCP_CUDA(ctx, cudaMemsetAsync(hist, 0, (OS_MAX_PASSES * RS_BINS + 16 + (size_t)passes * nblk * RS_BINS) * 4, st));
        k_os_hist<<<nblk, RS_THREADS, 0, st>>>(kin, n, passes, gbase, tickets, nblk);
        for (int pass = 0; pass < passes; ++pass) {
            k_os_scatter<<<nblk, RS_THREADS, 0, st>>>(kin, vin, n, pass * 8, gbase + pass * RS_BINS,
                                                      status + (size_t)pass * nblk * RS_BINS, tickets + 1 + pass,
                                                      kout, vout);
            unsigned long long* tk = kin; kin = kout; kout = tk;
            uint32_t* tv = vin; vin = vout; vout = tv;
        }
        ctx->launches += 2 + passes;
        GCP_CUDA(ctx, cudaGetLastError());
        if (out_keys) *out_keys = kin;
        if (out_vals) *out_vals = vin;
        return 0;
    }
    const bool fused_scan = nblk <= RS_FUSED_SCAN_MAX_TILES;
    uint32_t* done = hist + (size_t)RS_BINS * nblk + RS_BINS;          // ticket counter behind the table
    if (fused_scan) GCP_CUDA(ctx, cudaMemsetAsync(done, 0, 4, st));
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * 8;
        if (fused_scan) {
            k_rs_hist<<<nblk, RS_THREADS, 0, st>>>(kin, n, shift, hist, nblk, done);
            ctx->launches += 2;
        } else {
            k_rs_hist<<<nblk, RS_THREADS, 0, st>>>(kin, n, shift, hist, nblk, nullptr);
            k_rs_scan_bins<<<RS_BINS, 1024, 0, st>>>(hist, nblk);
            k_rs_scan_totals<<<1, RS_BINS, 0, st>>>(hist + (size_t)RS_BINS * nblk);
            ctx->launches += 4;
        }
        k_rs_scatter<<<nblk, RS_THREADS, 0, st>>>(kin, vin, n, shift, hist, nblk, kout, vout);
        unsigned long long* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    GCP_CUDA(ctx, cudaGetLastError());
    if (out_keys) *out_keys = kin;
    if (out_vals) *out_vals = vin;
    return 0;
}

// Sorts (keys, vals) in place (8 passes: the data ends up in the original buffers).
static int radix_sort_pairs(gcp_ctx* ctx, unsigned long long* keys, uint32_t* vals, uint32_t n,
                            unsigned long long* keys_tmp, uint32_t* vals_tmp, uint32_t* hist,
                            cudaStream_t st)
{
    return radix_sort_pairs_bits(ctx, keys, vals, n, keys_tmp, vals_tmp, hist, 64, st, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------
// prefix (min1, argmin1 position, min2) of b along the a-sorted order; one CTA.
//   rec[k] = {m1, m2}, arg[k] = sorted position of m1;  inv[perm[k]] = k; sa[k] = a[perm[k]]
// ---------------------------------------------------------------------------------
struct Min2 { double m1, m2; uint32_t i1; };

__device__ __forceinline__ Min2 min2_combine(const Min2& L, const Min2& R)
{
    Min2 o;
    if (L.m1 <= R.m1) { o.m1 = L.m1; o.i1 = L.i1; o.m2 = fmin(L.m2, R.m1); }
    else              { o.m1 = R.m1; o.i1 = R.i1; o.m2 = fmin(L.m1, R.m2); }
    return o;
}

__device__ __forceinline__ Min2 min2_shfl_up(const Min2& v, int o)
{
    Min2 r;
    r.m1 = __shfl_up_sync(FULL, v.m1, o);
    r.m2 = __shfl_up_sync(FULL, v.m2, o);
    r.i1 = __shfl_up_sync(FULL, v.i1, o);
    return r;
}

// sa[k] = a[perm[k]], sb[k] = b[perm[k]] (with -0.0 folded), inv[perm[k]] = k
__global__ void k_gather_sorted(const double2* __restrict__ sc, const uint32_t* __restrict__ perm,
                                uint32_t n, double* __restrict__ sa, double* __restrict__ sb,
                                uint32_t* __restrict__ inv)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = perm[k];
    const double2 s = sc[p];
    sa[k] = s.x + 0.0;
    sb[k] = s.y + 0.0;
    inv[p] = k;
}

// Block-wide inclusive scan of Min2 over 256 threads x 4 consecutive elements (1024 per CTA).
// mode 0: write only the CTA total to partial[blockIdx.x]; mode 1: apply the exclusive prefix
// of the preceding CTAs (prefix[blockIdx.x]) and write rec / arg.
constexpr int PM_THREADS = 256;
constexpr int PM_ITEMS = 4;
constexpr int PM_TILE = PM_THREADS * PM_ITEMS;

__device__ __forceinline__ Min2 min2_identity()
{
    Min2 m; m.m1 = CUDART_INF; m.m2 = CUDART_INF; m.i1 = 0xFFFFFFFFu;
    return m;
}
__device__ __forceinline__ void min2_push(Min2& m, double v, uint32_t k)
{
    if (v < m.m1) { m.m2 = m.m1; m.m1 = v; m.i1 = k; }
    else if (v < m.m2) m.m2 = v;
}

template <int MODE>
__global__ void __launch_bounds__(PM_THREADS)
k_prefix_min2(const double* __restrict__ sb, uint32_t n, Min2* __restrict__ partial,
              const Min2* __restrict__ prefix, double2* __restrict__ rec, uint32_t* __restrict__ arg)
{
    __shared__ Min2 s_w[PM_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = blockIdx.x * PM_TILE + tid * PM_ITEMS;
    double v[PM_ITEMS];
    Min2 loc = min2_identity();
    #pragma unroll
    for (int k = 0; k < PM_ITEMS; ++k) {
        v[k] = (b + k < n) ? sb[b + k] : CUDART_INF;
        if (b + k < n) min2_push(loc, v[k], b + k);
    }
    Min2 inc = loc;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Min2 t = min2_shfl_up(inc, o);
        if (lane >= o) inc = min2_combine(t, inc);
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (MODE == 0) {
        if (tid == 0) {
            Min2 tot = s_w[0];
            for (int w = 1; w < PM_THREADS / 32; ++w) tot = min2_combine(tot, s_w[w]);
            partial[blockIdx.x] = tot;
        }
        return;
    }
    Min2 ex = prefix[blockIdx.x];
    for (int w = 0; w < warp; ++w) ex = min2_combine(ex, s_w[w]);
    {
        const Min2 prev = min2_shfl_up(inc, 1);
        if (lane > 0) ex = min2_combine(ex, prev);
    }
    Min2 run = ex;
    #pragma unroll
    for (int k = 0; k < PM_ITEMS; ++k) {
        if (b + k < n) {
            min2_push(run, v[k], b + k);
            rec[b + k] = make_double2(run.m1, run.m2);
            arg[b + k] = run.i1;
        }
    }
}

// exclusive scan of the CTA totals, one CTA (sequential per thread chunk + warp/block scan)
__global__ void __launch_bounds__(1024)
k_prefix_min2_partials(const Min2* __restrict__ partial, uint32_t nb, Min2* __restrict__ prefix)
{
    __shared__ Min2 s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t b = min(nb, tid * per), e = min(nb, b + per);
    Min2 loc = min2_identity();
    for (uint32_t k = b; k < e; ++k) loc = min2_combine(loc, partial[k]);
    Min2 inc = loc;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Min2 t = min2_shfl_up(inc, o);
        if (lane >= o) inc = min2_combine(t, inc);
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    Min2 ex = min2_identity();
    for (int w = 0; w < warp; ++w) ex = min2_combine(ex, s_w[w]);
    {
        const Min2 prev = min2_shfl_up(inc, 1);
        if (lane > 0) ex = min2_combine(ex, prev);
    }
    for (uint32_t k = b; k < e; ++k) { prefix[k] = ex; ex = min2_combine(ex, partial[k]); }
}

__global__ void __launch_bounds__(256)
k_maximin_sweep(const double2* __restrict__ sc, uint32_t n, const double* __restrict__ sa,
                const double2* __restrict__ rec, const uint32_t* __restrict__ arg,
                const uint32_t* __restrict__ inv, uint32_t row_begin, uint32_t row_end,
                double* __restrict__ fit)
{
    const uint32_t i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_end) return;
    const double2 me = sc[i];
    const double a = me.x + 0.0, b = me.y + 0.0;
    const uint32_t p = inv[i];
    // largest k with x_(k) >= Y(k); -1 if none
    int lo = -1, hi = (int)n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;                // hi >= 0 here, so the sum is >= 0
        const double2 r = rec[mid];
        const double pm = (arg[mid] == p) ? r.y : r.x;
        const double x = __dsub_rn(a, sa[mid]);
        const double y = __dsub_rn(b, pm);                 // b - inf = -inf for an empty prefix
        if (x >= y) lo = mid; else hi = mid - 1;
    }
    double best = -CUDART_INF;
    const int kl = (lo == (int)p) ? lo - 1 : lo;
    const int kr = (lo + 1 == (int)p) ? lo + 2 : lo + 1;
    #pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int k = c ? kr : kl;
        if (k >= 0 && k < (int)n) {
            const double2 r = rec[k];
            const double pm = (arg[k] == p) ? r.y : r.x;
            const double x = __dsub_rn(a, sa[k]);
            const double y = __dsub_rn(b, pm);
            best = fmax(best, fmin(x, y));
        }
    }
    fit[i] = best;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_scratch_bytes(uint32_t n)
{
    // keys x2, vals x2, hist, sa, sb, rec, arg, inv
    return 2 * al256((size_t)n * 8) + 2 * al256((size_t)n * 4) +
           al256(radix_hist_bytes(n)) + 2 * al256((size_t)n * 8) +
           al256((size_t)n * 16) + 2 * al256((size_t)n * 4) +
           al256((size_t)2 * ((n + 1023) / 1024 + 1) * 24) + 256;
}

struct SortScratch {
    unsigned long long *keys, *keys_tmp;
    uint32_t *vals, *vals_tmp, *hist, *arg, *inv;
    double *sa, *sb;
    double2* rec;
    Min2* part;
};

static SortScratch carve(void* base, uint32_t n)
{
    SortScratch s;
    char* p = reinterpret_cast<char*>(base);
    s.keys = reinterpret_cast<unsigned long long*>(p);      p += al256((size_t)n * 8);
    s.keys_tmp = reinterpret_cast<unsigned long long*>(p);  p += al256((size_t)n * 8);
    s.vals = reinterpret_cast<uint32_t*>(p);                p += al256((size_t)n * 4);
    s.vals_tmp = reinterpret_cast<uint32_t*>(p);            p += al256((size_t)n * 4);
    s.hist = reinterpret_cast<uint32_t*>(p);                p += al256(radix_hist_bytes(n));
    s.sa = reinterpret_cast<double*>(p);                    p += al256((size_t)n * 8);
    s.sb = reinterpret_cast<double*>(p);                    p += al256((size_t)n * 8);
    s.rec = reinterpret_cast<double2*>(p);                  p += al256((size_t)n * 16);
    s.arg = reinterpret_cast<uint32_t*>(p);                 p += al256((size_t)n * 4);
    s.inv = reinterpret_cast<uint32_t*>(p);                 p += al256((size_t)n * 4);
    s.part = reinterpret_cast<Min2*>(p);
    return s;
}

// sc: dev double2 [n] (scaled objectives); fit rows [r0, r1) are written.
int launch_maximin_sweep(gcp_ctx* ctx, const double2* sc, uint32_t n, uint32_t r0, uint32_t r1,
                         double* fit, void* scratch, cudaStream_t st)
{
    SortScratch s = carve(scratch, n);
    k_rs_make_keys<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double*>(sc), 2, n,
                                                    s.keys, s.vals, nullptr);
    ctx->launches++;
    int r = radix_sort_pairs(ctx, s.keys, s.vals, n, s.keys_tmp, s.vals_tmp, s.hist, st);
    if (r) return r;
    k_gather_sorted<<<(n + 255) / 256, 256, 0, st>>>(sc, s.vals, n, s.sa, s.sb, s.inv);
    {
        const uint32_t nb = (n + PM_TILE - 1) / PM_TILE;
        k_prefix_min2<0><<<nb, PM_THREADS, 0, st>>>(s.sb, n, s.part, nullptr, nullptr, nullptr);
        k_prefix_min2_partials<<<1, 1024, 0, st>>>(s.part, nb, s.part + nb);
        k_prefix_min2<1><<<nb, PM_THREADS, 0, st>>>(s.sb, n, nullptr, s.part + nb, s.rec, s.arg);
        ctx->launches += 2;
    }
    const uint32_t rows = r1 - r0;
    k_maximin_sweep<<<(rows + 255) / 256, 256, 0, st>>>(sc, n, s.sa, s.rec, s.arg, s.inv, r0, r1, fit);
    ctx->launches += 3;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

__global__ void k_copy_order(const uint32_t* __restrict__ vals, uint32_t n, int32_t* __restrict__ order)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[i] = (int32_t)vals[i];
}

int launch_arena_order_sort(gcp_ctx* ctx, const double* fit, uint32_t n, int32_t* order,
                            uint8_t* nondom, void* scratch, cudaStream_t st)
{
    SortScratch s = carve(scratch, n);
    k_rs_make_keys<<<(n + 255) / 256, 256, 0, st>>>(fit, 1, n, s.keys, s.vals, nondom);
    ctx->launches++;
    int r = radix_sort_pairs(ctx, s.keys, s.vals, n, s.keys_tmp, s.vals_tmp, s.hist, st);
    if (r) return r;
    k_copy_order<<<(n + 255) / 256, 256, 0, st>>>(s.vals, n, order);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
