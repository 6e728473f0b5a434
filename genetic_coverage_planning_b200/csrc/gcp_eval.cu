// Evaluation kernels: path costs (gather), coverage (bitset union + popcount),
// loop closures + constraints + objective assembly.  sm_100a, no tensor cores: this
// path is integer/bit work bound by L2/HBM gathers and shared-memory atomics.
//
// Reference semantics restated (citations into /root/reference/src):
//   mappy.py:330-370   getCoverage      -> k_path_costs + k_coverage
//   mappy.py:372-437   getLoopClosures  -> k_loop_closures
//   geneticalgorithm.py:177-201 calcObj -> tail of k_loop_closures
#include "gcp_common.cuh"

#define FULL 0xffffffffu

// ============================================================================
// K1  path costs: one CTA per organism.
//   step rule (mappy.py:343-345): for idx, wpt in row: stop at wpt == -1 or idx == L-1
//   next gene -1 addresses column N-1 (numpy negative index)
//   non-edge -> zeros: dist 0, turn 0, single-pixel footprint = null edge E+wpt (F4)
//   per-agent sums are sequential fp64 adds in step order (mappy.py:351-352) so the
//   result is bit-identical to the reference's `+=` loop.
// ============================================================================
constexpr int K1_THREADS = 256;
constexpr int K1_STAGE_ELEMS = 2048;   // double2 entries (32 KB)

__global__ void __launch_bounds__(K1_THREADS)
k_path_costs(const int32_t* __restrict__ dna, uint32_t P, GcpTables t, int A, int L,
             uint32_t* __restrict__ edge_ids, double* __restrict__ dist,
             double* __restrict__ turn, uint8_t* __restrict__ flags)
{
    __shared__ int s_first_neg[GCP_MAX_AGENTS];
    __shared__ int s_bad;       // bit0: non-traversable step, bit1: gene out of range
    __shared__ double2 s_stage[K1_STAGE_ELEMS];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x;
    const int AL = A * L;
    const int32_t* d = dna + (size_t)org * AL;

    if (tid < A) s_first_neg[tid] = L;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int idx = tid; idx < AL; idx += K1_THREADS) {
        int g = d[idx];
        if (g < 0) {
            atomicMin(&s_first_neg[idx / L], idx % L);
            if (g != -1) atomicOr(&s_bad, 2);
        } else if ((uint32_t)g >= t.N) {
            atomicOr(&s_bad, 2);
        }
    }
    __syncthreads();

    const int CL = min(L, K1_STAGE_ELEMS / A);     // columns staged per pass
    double acc_d = 0.0, acc_t = 0.0;               // lane a of warp 0 owns agent a
    for (int c0 = 0; c0 < L; c0 += CL) {
        const int cl = min(CL, L - c0);
        for (int k = tid; k < A * cl; k += K1_THREADS) {
            const int a = k / cl, ii = k - a * cl, i = c0 + ii;
            const int nsteps = min(s_first_neg[a], L - 1);
            uint32_t e = GCP_INVALID_EDGE;
            double2 c = make_double2(0.0, 0.0);
            if (i < nsteps) {
                const uint32_t w = (uint32_t)d[a * L + i];
                int nx = d[a * L + i + 1];
                const bool sentinel = nx < 0;
                if (nx < 0) nx += (int)t.N;                  // numpy wrap: -1 -> N-1
                if (w < t.N && nx >= 0 && (uint32_t)nx < t.N) {
                    const uint32_t lo = t.row_ptr[w], hi = t.row_ptr[w + 1];
                    e = t.E + w;                             // null edge unless found
                    bool found = false;
                    for (uint32_t q = lo; q < hi; ++q)
                        if (t.col[q] == (uint32_t)nx) { e = q; found = true; break; }
                    if (!found && !sentinel) atomicOr(&s_bad, 1);
                    c = t.edge_cost[e];
                }
            }
            edge_ids[(size_t)org * AL + a * L + i] = e;
            s_stage[ii * A + a] = c;
        }
        __syncthreads();
        if (tid < A) {
            for (int ii = 0; ii < cl; ++ii) {
                const double2 c = s_stage[ii * A + tid];
                acc_d = __dadd_rn(acc_d, c.x);
                acc_t = __dadd_rn(acc_t, c.y);
            }
        }
        __syncthreads();
    }
    if (tid < A) {
        dist[(size_t)org * A + tid] = acc_d;
        turn[(size_t)org * A + tid] = acc_t;
    }
    if (tid == 0 && flags) {
        flags[(size_t)org * 4 + 1] = (s_bad & 1) ? 0 : 1;   // traversable
        flags[(size_t)org * 4 + 3] = (s_bad & 2) ? 1 : 0;   // error: gene out of range
    }
}

int launch_path_costs(gcp_ctx* ctx, const int32_t* dna, uint32_t P, uint32_t* edge_ids,
                      double* dist, double* turn, uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    StageTimer tm(ctx, 0, st);
    k_path_costs<<<P, K1_THREADS, 0, st>>>(dna, P, ctx->t, ctx->p.num_agents,
                                           ctx->p.max_dna_len, edge_ids, dist, turn, flags);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

// ============================================================================
// K2  coverage: union of the visited edges' footprint sub-tiles, popcount against the
//     map planes.  One CTA per organism.
//
//   phase 1 (thread per step): each step's edge contributes 1..4 sub-tiles, each aligned
//     to one 64x64 map block.  Sub-tiles are grouped by block with an open-addressing
//     hash table in shared memory (key = block id) and per-block linked lists (one
//     atomicExch per sub-tile).
//   phase 2 (warp per block, "owner computes"): the warp holds the block's 64 rows in
//     registers (lane l = rows 2l,2l+1 = one uint4), walks the list, ORs 128-bit
//     coalesced loads of the stored quarters, then popcounts against the mask planes.
//     No shared-memory bitmap, no zeroing pass, no bitmap atomics.
//   If the table or pair store overflows (very long genomes), the organism is redone in
//   R = 2,4,8.. spatial partitions of the blocks; results are exact in every case.
// ============================================================================
__device__ __forceinline__ uint32_t hash_block(uint32_t b) { return b * 2654435761u; }
__device__ __forceinline__ uint32_t part_block(uint32_t b) { return (b * 0x85ebca6bu) >> 12; }

__device__ __forceinline__ uint4 ld_quarter(const uint4* __restrict__ tile, uint32_t p, int lane)
{
    const uint32_t qm = p & 15u;
    const int q = lane >> 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((qm >> q) & 1u) {
        const size_t idx = ((size_t)(p >> 4) + __popc(qm & ((1u << q) - 1u))) * 8 + (lane & 7);
        v = __ldg(tile + idx);
    }
    return v;
}

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

constexpr uint32_t K2_NIL = 0xFFFFu;

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_coverage(const uint32_t* __restrict__ edge_ids, uint32_t P, GcpTables t, int AL,
           int T /*pow2*/, int MAXP /*<=65535*/, uint32_t* __restrict__ cov,
           uint8_t* __restrict__ flags)
{
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    uint4*    s_scr  = reinterpret_cast<uint4*>(smem);                  // [NW][32] gray scratch
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_scr + NW * 32);    // [T] block id + 1
    uint32_t* s_head = s_keys + T;                                       // [T] first pair | NIL
    uint32_t* s_pay  = s_head + T;                                       // [MAXP]
    uint16_t* s_next = reinterpret_cast<uint16_t*>(s_pay + MAXP);        // [MAXP]
    __shared__ uint32_t s_npairs, s_overflow;
    __shared__ uint32_t s_tot[6];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* eids = edge_ids + (size_t)org * AL;
    const int logT = 31 - __clz(T);

    uint32_t R = 1;
    bool failed = false;
    uint32_t n_mask0, n_free0, wg_mask, wg_all, n_gmask, n_all;

    for (;;) {                                   // retry loop over partition count R
        n_mask0 = n_free0 = wg_mask = wg_all = n_gmask = n_all = 0;
        bool overflow = false;
        for (uint32_t r = 0; r < R && !overflow; ++r) {
            // ---- clear table ----
            for (int i = tid; i < T; i += THREADS) { s_keys[i] = 0; s_head[i] = K2_NIL; }
            if (tid == 0) { s_npairs = 0; s_overflow = 0; }
            __syncthreads();
            // ---- phase 1: group sub-tiles by block ----
            for (int s0 = warp * 32; s0 < AL; s0 += THREADS) {
                const int s = s0 + lane;
                uint32_t lo = 0, hi = 0;
                if (s < AL) {
                    const uint32_t e = eids[s];
                    if (e != GCP_INVALID_EDGE) { lo = t.sub_ptr[e]; hi = t.sub_ptr[e + 1]; }
                }
                uint32_t mine = hi - lo;
                if (R > 1) {
                    mine = 0;
                    for (uint32_t k = lo; k < hi; ++k)
                        mine += ((part_block(t.sub_block[k]) & (R - 1)) == r);
                }
                // warp-aggregated allocation of pair slots
                uint32_t incl = mine;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t total = __shfl_sync(FULL, incl, 31);
                if (total == 0) continue;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_npairs, total);
                base = __shfl_sync(FULL, base, 0);
                if (base + total > (uint32_t)MAXP) {
                    if (lane == 0) s_overflow = 1;
                    continue;
                }
                uint32_t idx = base + incl - mine;
                for (uint32_t k = lo; k < hi; ++k) {
                    const uint32_t blk = t.sub_block[k];
                    if (R > 1 && (part_block(blk) & (R - 1)) != r) continue;
                    const uint32_t key = blk + 1;
                    uint32_t h = hash_block(blk) >> (32 - logT);
                    int probes = 0;
                    for (;;) {
                        const uint32_t cur = ((volatile uint32_t*)s_keys)[h];
                        if (cur == key) break;
                        if (cur == 0) {
                            const uint32_t old = atomicCAS(&s_keys[h], 0u, key);
                            if (old == 0 || old == key) break;
                        }
                        h = (h + 1) & (T - 1);
                        if (++probes >= T) { s_overflow = 1; break; }
                    }
                    if (probes >= T) break;
                    s_pay[idx] = t.sub_payload[k];
                    s_next[idx] = (uint16_t)atomicExch(&s_head[h], idx);
                    ++idx;
                }
            }
            __syncthreads();
            overflow = (s_overflow != 0);      // table or pair store full: more partitions
            if (overflow) { __syncthreads(); break; }
            // ---- phase 2: owner warps ----
            for (int g0 = warp * 32; g0 < T; g0 += THREADS) {
                const uint32_t hd = s_head[g0 + lane];
                uint32_t m = __ballot_sync(FULL, hd != K2_NIL);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    uint32_t i = __shfl_sync(FULL, hd, b);
                    const uint32_t blk = s_keys[g0 + b] - 1;
                    uint4 acc = make_uint4(0, 0, 0, 0);
                    while (i != K2_NIL) {
                        uint32_t p0, p1 = 0, p2 = 0, p3 = 0;
                        p0 = s_pay[i]; i = s_next[i];
                        if (i != K2_NIL) { p1 = s_pay[i]; i = s_next[i]; }
                        if (i != K2_NIL) { p2 = s_pay[i]; i = s_next[i]; }
                        if (i != K2_NIL) { p3 = s_pay[i]; i = s_next[i]; }
                        const uint4 v0 = ld_quarter(t.tile, p0, lane);
                        const uint4 v1 = ld_quarter(t.tile, p1, lane);
                        const uint4 v2 = ld_quarter(t.tile, p2, lane);
                        const uint4 v3 = ld_quarter(t.tile, p3, lane);
                        or128(acc, v0); or128(acc, v1); or128(acc, v2); or128(acc, v3);
                    }
                    const uint4 mk = __ldg(t.blk_mask0 + (size_t)blk * 32 + lane);
                    const uint4 fr = __ldg(t.blk_free0 + (size_t)blk * 32 + lane);
                    n_all   += popc128(acc);
                    n_mask0 += popc128(and128(acc, mk));
                    n_free0 += popc128(and128(acc, fr));
                    const uint32_t ga = t.gray_ptr[blk], gb = t.gray_ptr[blk + 1];
                    if (gb > ga) {                     // gray pixels (0 < g < 255), SURVEY F6
                        s_scr[warp * 32 + lane] = acc;
                        __syncwarp();
                        const uint64_t* rows = reinterpret_cast<const uint64_t*>(s_scr + warp * 32);
                        for (uint32_t g = ga + lane; g < gb; g += 32) {
                            const uint32_t ent = t.gray_ent[g];
                            const uint32_t pos = ent & 4095u;
                            if ((rows[pos >> 6] >> (pos & 63u)) & 1ull) {
                                const uint32_t w = (ent >> 12) & 255u;
                                wg_all += w;
                                if ((ent >> 20) & 1u) { wg_mask += w; n_gmask += 1; }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
        }
        if (!overflow) break;
        R <<= 1;
        if (R > 4096) { failed = true; break; }
    }

    // ---- reduce the six counters over the CTA ----
    if (tid < 6) s_tot[tid] = 0;
    __syncthreads();
    uint32_t vals[6] = {n_mask0, n_free0, wg_mask, wg_all, n_mask0 + n_gmask, n_all};
    #pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t v = vals[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(&s_tot[k], v);
    }
    __syncthreads();
    if (tid < 8) cov[(size_t)org * 8 + tid] = (tid < 6 && !failed) ? s_tot[tid] : 0u;
    if (tid == 0 && failed && flags) flags[(size_t)org * 4 + 3] |= 2;
}

static size_t cov_smem_bytes(int threads, int T, int MAXP)
{
    return (size_t)(threads / 32) * 32 * 16 + (size_t)T * 8 + (size_t)MAXP * 4 +
           (size_t)((MAXP * 2 + 15) / 16) * 16;
}

int launch_coverage(gcp_ctx* ctx, const uint32_t* edge_ids, uint32_t P, uint32_t* cov,
                    uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    const int AL = ctx->p.num_agents * ctx->p.max_dna_len;
    // pair store sized for ~3 sub-tiles per step, table for a quarter of that
    int maxp = ctx->cov_maxpairs;
    if (maxp <= 0) {
        maxp = 1024;
        while (maxp < 3 * AL && maxp < 16384) maxp <<= 1;
    }
    if (maxp > 65534) maxp = 65534;
    int T = ctx->cov_table;
    if (T <= 0) { T = 256; while (T < maxp / 4) T <<= 1; }
    constexpr int THREADS = 256;
    const size_t smem = cov_smem_bytes(THREADS, T, maxp);
    if (smem > (size_t)ctx->max_smem_optin)
        return gcp_fail(ctx, -3, "coverage kernel needs %zu B shared memory", smem);
    static bool attr_set = false;
    if (!attr_set || smem > 48 * 1024) {
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_coverage<THREADS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        attr_set = true;
    }
    StageTimer tm(ctx, 1, st);
    k_coverage<THREADS><<<P, THREADS, smem, st>>>(edge_ids, P, ctx->t, AL, T, maxp, cov, flags);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

// ============================================================================
// K3  loop closures + constraints + objectives.  One CTA per organism.
//   wps   = dna[dna != -1]                       (row-major compaction)  mappy.py:390
//   key_i = (wps[i], wps[(i+1) mod n])            incl. boundary / wrap edges  :391
//   groups of equal keys, size > 1, none of whose positions is a path split (F10)
//   one owner  -> lc[a][a] += 1 iff some consecutive (ascending) gap > thresh  :413
//   >= 2 owners -> lc[x][y] += 1 for every owner pair x < y                    :417-424
//   The (key, position) pairs are sorted as packed u64 (from:24 | to:24 | pos:16) with a
//   shared-memory bitonic network, so groups are contiguous and positions ascend
//   (= the stable-argsort canonical order).
// ============================================================================
constexpr int K3_THREADS = 256;

__device__ double np_sum_order(const double* a, int n)
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

__global__ void __launch_bounds__(K3_THREADS)
k_loop_closures(const int32_t* __restrict__ dna, uint32_t P, int A, int L, int SNMAX,
                gcp_params prm, gcp_map_consts mc,
                const double* __restrict__ dist, const double* __restrict__ turn,
                const uint32_t* __restrict__ cov,
                int32_t* __restrict__ lc_out, double* __restrict__ obj,
                double* __restrict__ neg_cov_out, uint8_t* __restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem);   // [SNMAX]
    int32_t* s_seq = reinterpret_cast<int32_t*>(s_key + SNMAX);                 // [A*L]
    uint8_t* s_own = reinterpret_cast<uint8_t*>(s_seq + A * L);                 // [A*L] owner | split<<7
    __shared__ int s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    __shared__ int s_lc[GCP_MAX_AGENTS * GCP_MAX_AGENTS];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K3_THREADS / 32;
    const int32_t* d = dna + (size_t)org * A * L;

    for (int i = tid; i < A * A; i += K3_THREADS) s_lc[i] = 0;
    // ---- valid-gene counts per row ----
    for (int a = warp; a < A; a += NW) {
        int cnt = 0;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const bool v = (i < L) && (d[a * L + i] != -1);
            cnt += __popc(__ballot_sync(FULL, v));
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
    // ---- compaction: seq[pos] = gene, own[pos] = agent | first-of-row(a>=1) << 7 ----
    for (int a = warp; a < A; a += NW) {
        int run = s_base[a];
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? d[a * L + i] : -1;
            const bool v = g != -1;
            const uint32_t bal = __ballot_sync(FULL, v);
            if (v) {
                const int pos = run + __popc(bal & ((1u << lane) - 1u));
                s_seq[pos] = g;
                s_own[pos] = (uint8_t)(a | ((a >= 1 && pos == s_base[a]) ? 0x80 : 0));
            }
            run += __popc(bal);
        }
    }
    __syncthreads();
    int SN = 2;
    while (SN < n) SN <<= 1;
    if (SN > SNMAX) SN = SNMAX;          // cannot happen: SNMAX >= A*L
    for (int i = tid; i < SN; i += K3_THREADS) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const unsigned long long f = (uint32_t)s_seq[i] & 0xFFFFFFu;
            const unsigned long long g = (uint32_t)s_seq[(i + 1 == n) ? 0 : i + 1] & 0xFFFFFFu;
            k = (f << 40) | (g << 16) | (unsigned long long)i;
        }
        s_key[i] = k;
    }
    __syncthreads();
    // ---- bitonic sort, ascending ----
    for (int k = 2; k <= SN; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int tt = tid; tt < (SN >> 1); tt += K3_THREADS) {
                const int i = ((tt & ~(j - 1)) << 1) | (tt & (j - 1));
                const int l = i | j;
                const unsigned long long x = s_key[i], y = s_key[l];
                const bool asc = (i & k) == 0;
                if ((x > y) == asc) { s_key[i] = y; s_key[l] = x; }
            }
            __syncthreads();
        }
    }
    // ---- group analysis: the first element of each group walks it ----
    const int thresh = prm.solo_sep_thresh;
    for (int i = tid; i < n; i += K3_THREADS) {
        const unsigned long long ki = s_key[i] >> 16;
        if (i > 0 && (s_key[i - 1] >> 16) == ki) continue;
        if (i + 1 >= n || (s_key[i + 1] >> 16) != ki) continue;          // size 1
        uint32_t owners = 0;
        bool split = false, gap = false;
        int prev = -1;
        for (int j = i; j < n && (s_key[j] >> 16) == ki; ++j) {
            const int pos = (int)(s_key[j] & 0xFFFFu);
            const uint8_t o = s_own[pos];
            owners |= 1u << (o & 31);
            split |= (o & 0x80) != 0;
            if (prev >= 0 && pos - prev > thresh) gap = true;
            prev = pos;
        }
        if (split) continue;                                             // F10
        if (__popc(owners) == 1) {
            if (gap) { const int a = __ffs(owners) - 1; atomicAdd(&s_lc[a * A + a], 1); }
        } else {
            for (uint32_t mx = owners; mx; mx &= mx - 1) {
                const int x = __ffs(mx) - 1;
                for (uint32_t my = mx & (mx - 1); my; my &= my - 1)
                    atomicAdd(&s_lc[x * A + (__ffs(my) - 1)], 1);
            }
        }
    }
    __syncthreads();
    if (lc_out)
        for (int i = tid; i < A * A; i += K3_THREADS) lc_out[(size_t)org * A * A + i] = s_lc[i];

    // ---- constraints + objectives (geneticalgorithm.py:177-201), one thread ----
    if (tid == 0) {
        bool solo_bad = false;
        long long combo = 0;
        uint32_t adj[GCP_MAX_AGENTS];
        for (int a = 0; a < A; ++a) {
            if (s_lc[a * A + a] < prm.min_solo_lcs) solo_bad = true;
            uint32_t m = 0;
            for (int b = 0; b < A; ++b) {
                if (b == a) continue;
                combo += s_lc[a * A + b];
                if (s_lc[a * A + b] + s_lc[b * A + a] >= 1) m |= 1u << b;
            }
            adj[a] = m;
        }
        int ncomp = 0;
        uint32_t seen = 0;
        for (int a = 0; a < A; ++a) {
            if ((seen >> a) & 1u) continue;
            ++ncomp;
            uint32_t fr = 1u << a;
            while (fr) {
                seen |= fr;
                uint32_t nf = 0;
                for (uint32_t m = fr; m; m &= m - 1) nf |= adj[__ffs(m) - 1];
                fr = nf & ~seen;
            }
        }
        const bool feasible = !(solo_bad || combo < (long long)prm.min_comb_lcs || ncomp > 1);

        // coverage (mappy.py:358-364) from the integer pixel counts
        const uint32_t* c = cov + (size_t)org * 8;
        const long long w_mask = 255ll * c[0] + c[2];
        const long long w_all = 255ll * c[1] + c[3];
        const double wall = __ddiv_rn(__ddiv_rn((double)(mc.gsum_mask + w_mask), 255.0),
                                      (double)mc.mask_size);
        const double sum_draw = __dadd_rn(mc.num_occluded, __ddiv_rn((double)w_all, 255.0));
        const double internal = __ddiv_rn(__dsub_rn(sum_draw, mc.num_occluded),
                                          __dsub_rn((double)mc.map_size, mc.num_occluded));
        const double coverage = __dadd_rn(__dmul_rn(prm.coverage_blend, wall),
                                          __dmul_rn(prm.one_minus_blend, internal));
        const double neg_cov = -coverage;
        const double td = np_sum_order(dist + (size_t)org * A, A);
        const double tc = np_sum_order(turn + (size_t)org * A, A);
        const double travel = __dmul_rn(__dadd_rn(td, tc), prm.flight_time_scale);
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
}

int launch_loop_closures(gcp_ctx* ctx, const int32_t* dna, uint32_t P, const double* dist,
                         const double* turn, const uint32_t* cov, int32_t* lc, double* obj,
                         double* neg_cov, uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    const int A = ctx->p.num_agents, L = ctx->p.max_dna_len;
    int SNMAX = 2;
    while (SNMAX < A * L) SNMAX <<= 1;
    const size_t smem = (size_t)SNMAX * 8 + (size_t)A * L * 4 + (size_t)((A * L + 15) / 16) * 16;
    if (smem > (size_t)ctx->max_smem_optin)
        return gcp_fail(ctx, -3, "loop-closure kernel needs %zu B shared memory (A*L too large)", smem);
    GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(ctx, 2, st);
    k_loop_closures<<<P, K3_THREADS, smem, st>>>(dna, P, A, L, SNMAX, ctx->p, ctx->mc, dist, turn,
                                                 cov, lc, obj, neg_cov, flags);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
