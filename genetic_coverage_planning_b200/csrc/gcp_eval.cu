// Evaluation kernels: path costs (gather), coverage (bitset union + popcount),
// loop closures + constraints + objective assembly.  sm_100a, no tensor cores: this
// path is integer/bit work bound by L2/HBM gathers and shared-memory atomics.
//
// Reference semantics restated (citations into /root/reference/src):
//   mappy.py:330-370   getCoverage      -> k_path_costs + k_coverage
//   mappy.py:372-437   getLoopClosures  -> k_loop_closures
//   geneticalgorithm.py:177-201 calcObj -> tail of k_loop_closures
#include "gcp_common.cuh"
#include "gcp_device.cuh"

// ============================================================================
// K1  path costs: one CTA per organism.
//   step rule (mappy.py:343-345): for idx, wpt in row: stop at wpt == -1 or idx == L-1
//   next gene -1 addresses column N-1 (numpy negative index)
//   non-edge -> zeros: dist 0, turn 0, single-pixel footprint = null edge E+wpt (F4)
//   per-agent sums are sequential fp64 adds in step order (mappy.py:351-352) so the
//   result is bit-identical to the reference's `+=` loop.
// ============================================================================
// double2 entries per staging buffer, two buffers in dynamic shared memory: 2 x 32 KB for the 1,024-thread
// kernel of long genomes, 2 x 16 KB for the 256-thread one
template <int K1_THREADS> struct K1Stage { static constexpr int ELEMS = K1_THREADS >= 1024 ? 2048 : 1024; };

template <int K1_THREADS>
__global__ void __launch_bounds__(K1_THREADS)
k_path_costs(const int32_t* __restrict__ dna, uint32_t P, GcpTables t, int A, int L,
             uint32_t* __restrict__ edge_ids, uint32_t* __restrict__ step_info,
             const uint32_t* __restrict__ info_table /*m_sub_info or m_q_info*/,
             double* __restrict__ dist, double* __restrict__ turn, uint8_t* __restrict__ flags)
{
    __shared__ int s_first_neg[GCP_MAX_AGENTS];
    __shared__ int s_bad;       // bit0: non-traversable step, bit1: gene out of range
    extern __shared__ __align__(16) unsigned char k1_smem[];
    constexpr int K1_STAGE_ELEMS = K1Stage<K1_THREADS>::ELEMS;
    double2* s_stage_all = reinterpret_cast<double2*>(k1_smem);        // [2][K1_STAGE_ELEMS]

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
    // Two staging buffers: the A lanes that add up pass c (sequentially, in step order) do so while the other
    // warps are already gathering pass c + 1 into the other buffer - one barrier per pass; buffer c & 1 is
    // written again in pass c + 2, behind the barrier of pass c + 1, which the summing lanes only reach when
    // their sums of pass c are done.
    int pass = 0;
    for (int c0 = 0; c0 < L; c0 += CL, ++pass) {
        const int cl = min(CL, L - c0);
        double2* s_stage = s_stage_all + (pass & 1) * K1_STAGE_ELEMS;
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
                    const uint32_t lo = __ldg(t.row_ptr + w);
                    e = t.E + w;                             // null edge unless found
                    bool found = false;
                    if (t.adjc16) {                          // 8 x u16 out-neighbours in one 16-B load
                        const uint4 q = __ldg(t.adjc16 + w);
                        const uint32_t pat = (uint32_t)nx * 0x00010001u;
                        const uint32_t m0 = __vcmpeq2(q.x, pat), m1 = __vcmpeq2(q.y, pat);
                        const uint32_t m2 = __vcmpeq2(q.z, pat), m3 = __vcmpeq2(q.w, pat);
                        const uint32_t hit = ((m0 & 1u) | ((m0 >> 15) & 2u)) | (((m1 & 1u) | ((m1 >> 15) & 2u)) << 2) |
                                             (((m2 & 1u) | ((m2 >> 15) & 2u)) << 4) | (((m3 & 1u) | ((m3 >> 15) & 2u)) << 6);
                        if (hit) { e = lo + (uint32_t)(__ffs(hit) - 1); found = true; }
                    } else {
                        const uint32_t hi = __ldg(t.row_ptr + w + 1);
                        for (uint32_t q = lo; q < hi; ++q)
                            if (__ldg(t.col + q) == (uint32_t)nx) { e = q; found = true; break; }
                    }
                    if (!found && !sentinel) atomicOr(&s_bad, 1);
                    c = __ldg(t.edge_cost + e);
                }
            }
            if (edge_ids) edge_ids[(size_t)org * AL + a * L + i] = e;
            if (step_info)
                step_info[(size_t)org * AL + a * L + i] =
                    (e != GCP_INVALID_EDGE) ? __ldg(info_table + e) : 0u;
            s_stage[ii * A + a] = c;
        }
        __syncthreads();
        if (tid < A) {
            int ii = 0;
            for (; ii + 4 <= cl; ii += 4) {                   // four loads in flight, adds in step order
                const double2 c0_ = s_stage[ii * A + tid], c1_ = s_stage[(ii + 1) * A + tid];
                const double2 c2_ = s_stage[(ii + 2) * A + tid], c3_ = s_stage[(ii + 3) * A + tid];
                acc_d = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc_d, c0_.x), c1_.x), c2_.x), c3_.x);
                acc_t = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc_t, c0_.y), c1_.y), c2_.y), c3_.y);
            }
            for (; ii < cl; ++ii) {
                const double2 c = s_stage[ii * A + tid];
                acc_d = __dadd_rn(acc_d, c.x);
                acc_t = __dadd_rn(acc_t, c.y);
            }
        }
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
                      uint32_t* step_info, double* dist, double* turn, uint8_t* flags,
                      cudaStream_t st, int quarter_info)
{
    if (P == 0) return 0;
    if (step_info && !ctx->have_masked)
        return gcp_fail(ctx, -1, "step_info requested but no masked footprint store uploaded");
    StageTimer tm(ctx, 0, st);
    const uint32_t* info_table = quarter_info ? ctx->t.m_q_info : ctx->t.m_sub_info;
    if (ctx->p.num_agents * ctx->p.max_dna_len > 4096) {    // long genomes: more lookups in flight per organism
        const size_t smem = (size_t)2 * K1Stage<1024>::ELEMS * sizeof(double2);
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_path_costs<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_path_costs<1024><<<P, 1024, smem, st>>>(dna, P, ctx->t, ctx->p.num_agents, ctx->p.max_dna_len,
                                                  edge_ids, step_info, info_table, dist, turn, flags);
    } else {
        const size_t smem = (size_t)2 * K1Stage<256>::ELEMS * sizeof(double2);
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_path_costs<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_path_costs<256><<<P, 256, smem, st>>>(dna, P, ctx->t, ctx->p.num_agents, ctx->p.max_dna_len,
                                                edge_ids, step_info, info_table, dist, turn, flags);
    }
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

// ============================================================================
// K2  coverage: union of the visited edges' footprint sub-tiles, popcount against the
//     map planes.  One CTA per organism.
//
//   phase 1 (thread per step): each step's edge contributes 1..4 sub-tiles, each aligned
//     to one 64x64 map block.  Sub-tiles are bucketed by block: an open-addressing hash
//     table in shared memory maps block id -> slot, one shared atomicAdd per sub-tile
//     returns its rank inside the slot; a CTA-wide exclusive scan of the slot counts then
//     turns (slot, rank) into a position, i.e. a counting sort with one atomic per item.
//   phase 2 (warp per block, "owner computes"): the warp holds the block's 64 rows in
//     registers (lane l = rows 2l,2l+1 = one uint4), reads its bucket 32 payloads at a
//     time, broadcasts them by shuffle and ORs 128-bit coalesced loads of the stored
//     quarters (4 loads in flight per lane), then popcounts against the mask planes.
//     No shared-memory bitmap, no zeroing pass, no bitmap atomics.
//   If the table or pair store overflows (very long genomes), the organism is redone in
//   R = 2,4,8.. hash partitions of the blocks; results are exact in every case.
// ============================================================================

__device__ __forceinline__ uint4 ld_quarter(const uint4* __restrict__ tile, uint32_t p,
                                            uint32_t q, uint32_t below, uint32_t l7)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((p >> q) & 1u) {
        const uint32_t idx = (((p >> 4) + __popc(p & below)) << 3) | l7;
        v = __ldg(tile + idx);
    }
    return v;
}



// NEED_FREE: also count against the (g==0) plane (internal coverage, blend < 1 or detail)
// NEED_ALL : also count every painted pixel (detail output n_cov_all)
// HAS_GRAY : the map has pixels with 0 < g < 255
template <int THREADS, bool NEED_FREE, bool NEED_ALL, bool HAS_GRAY>
__global__ void __launch_bounds__(THREADS)
k_coverage(const uint32_t* __restrict__ edge_ids, uint32_t P, GcpTables t, int AL,
           int T /*pow2, multiple of THREADS*/, int MAXP /*<=65535*/, uint32_t R0,
           uint32_t* __restrict__ cov, uint8_t* __restrict__ flags)
{
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    uint4*    s_scr  = reinterpret_cast<uint4*>(smem);                  // [NW][32] gray scratch
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_scr + (HAS_GRAY ? NW * 32 : 0)); // [T] block+1
    uint32_t* s_cnt  = s_keys + T;                                       // [T] count -> base
    uint32_t* s_pay  = s_cnt + T;                                        // [MAXP] payload
    uint32_t* s_sr   = s_pay + MAXP;                                     // [MAXP] slot<<16 | rank
    uint16_t* s_sort = reinterpret_cast<uint16_t*>(s_sr + MAXP);         // [MAXP] pair index
    __shared__ uint32_t s_npairs, s_overflow, s_group;
    __shared__ uint32_t s_wsum[NW];
    __shared__ uint32_t s_tot[6];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* eids = edge_ids + (size_t)org * AL;
    const int logT = 31 - __clz(T);
    const uint32_t q = 0u + (lane >> 3), below = (1u << q) - 1u, l7 = lane & 7;

    uint32_t R = R0;
    bool failed = false;
    uint32_t n_mask0, n_free0, wg_mask, wg_all, n_gmask, n_all;

    for (;;) {                                   // retry loop over partition count R
        n_mask0 = n_free0 = wg_mask = wg_all = n_gmask = n_all = 0;
        bool overflow = false;
        for (uint32_t r = 0; r < R && !overflow; ++r) {
            // ---- clear table ----
            for (int i = tid; i < T; i += THREADS) { s_keys[i] = 0; s_cnt[i] = 0; }
            if (tid == 0) { s_npairs = 0; s_overflow = 0; s_group = 0; }
            __syncthreads();
            // ---- phase 1: slot + rank for every sub-tile ----
            for (int s0 = warp * 32; s0 < AL; s0 += THREADS) {
                const int s = s0 + lane;
                uint32_t lo = 0, hi = 0;
                if (s < AL) {
                    const uint32_t e = eids[s];
                    if (e != GCP_INVALID_EDGE) {
                        const uint32_t info = __ldg(t.sub_info + e);
                        lo = info & 0x0FFFFFFFu;
                        hi = lo + (info >> 28);
                    }
                }
                uint32_t mine = hi - lo;
                if (R > 1) {
                    mine = 0;
                    for (uint32_t k = lo; k < hi; ++k)
                        mine += ((part_block(__ldg(t.sub_desc + k).x) & (R - 1)) == r);
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
                    const uint2 dsc = __ldg(t.sub_desc + k);
                    const uint32_t blk = dsc.x;
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
                    const uint32_t rank = atomicAdd(&s_cnt[h], 1u);
                    GCP_DASSERT(h < (uint32_t)T && rank < 0x10000u);
                    s_pay[idx] = dsc.y;
                    s_sr[idx] = (h << 16) | rank;
                    ++idx;
                }
            }
            __syncthreads();
            overflow = (s_overflow != 0);      // table or pair store full: more partitions
            if (overflow) { __syncthreads(); break; }
            const uint32_t npairs = s_npairs;
            // ---- exclusive scan of the slot counts (T entries, T/THREADS per thread) ----
            {
                const int per = T / THREADS;
                uint32_t loc = 0;
                for (int k = 0; k < per; ++k) loc += s_cnt[k * THREADS + tid];
                uint32_t inc = loc;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                uint32_t woff = 0;
                for (int w2 = 0; w2 < warp; ++w2) woff += s_wsum[w2];
                uint32_t run = woff + inc - loc;
                for (int k = 0; k < per; ++k) {
                    const uint32_t c = s_cnt[k * THREADS + tid];
                    // keep the count in the high half: cnt[slot] = count<<16 | base
                    s_cnt[k * THREADS + tid] = (c << 16) | run;
                    run += c;
                }
            }
            __syncthreads();
            // ---- place: sorted position = base[slot] + rank ----
            for (uint32_t i = tid; i < npairs; i += THREADS) {
                const uint32_t sr = s_sr[i];
                s_sort[(s_cnt[sr >> 16] & 0xFFFFu) + (sr & 0xFFFFu)] = (uint16_t)i;
            }
            __syncthreads();
            // ---- phase 2: owner warps, groups of 32 slots handed out dynamically ----
            for (;;) {
                uint32_t g = 0;
                if (lane == 0) g = atomicAdd(&s_group, 1u);
                g = __shfl_sync(FULL, g, 0);
                if (g * 32 >= (uint32_t)T) break;
                const uint32_t cb = s_cnt[g * 32 + lane];
                const uint32_t ky = s_keys[g * 32 + lane];
                uint32_t m = __ballot_sync(FULL, (cb >> 16) != 0);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t cbb = __shfl_sync(FULL, cb, b);
                    const uint32_t blk = __shfl_sync(FULL, ky, b) - 1;
                    const uint32_t cnt = cbb >> 16, base = cbb & 0xFFFFu;
                    const uint4 mk = __ldg(t.blk_mask0 + (size_t)blk * 32 + lane);
                    uint4 fr = make_uint4(0, 0, 0, 0);
                    if (NEED_FREE) fr = __ldg(t.blk_free0 + (size_t)blk * 32 + lane);
                    uint4 acc = make_uint4(0, 0, 0, 0);
                    for (uint32_t c0 = 0; c0 < cnt; c0 += 32) {
                        uint32_t my = 0;
                        if (c0 + lane < cnt) my = s_pay[s_sort[base + c0 + lane]];
                        const uint32_t nn = min(32u, cnt - c0);
                        for (uint32_t k = 0; k < nn; k += 4) {
                            const uint32_t p0 = __shfl_sync(FULL, my, k);
                            const uint32_t p1 = __shfl_sync(FULL, my, k + 1);
                            const uint32_t p2 = __shfl_sync(FULL, my, k + 2);
                            const uint32_t p3 = __shfl_sync(FULL, my, k + 3);
                            const uint4 v0 = ld_quarter(t.tile, p0, q, below, l7);
                            const uint4 v1 = ld_quarter(t.tile, p1, q, below, l7);
                            const uint4 v2 = ld_quarter(t.tile, p2, q, below, l7);
                            const uint4 v3 = ld_quarter(t.tile, p3, q, below, l7);
                            or128(acc, v0); or128(acc, v1); or128(acc, v2); or128(acc, v3);
                        }
                    }
                    n_mask0 += popc128(and128(acc, mk));
                    if (NEED_FREE) n_free0 += popc128(and128(acc, fr));
                    if (NEED_ALL) n_all += popc128(acc);
                    if (HAS_GRAY) {
                        const uint32_t ga = t.gray_ptr[blk], gb = t.gray_ptr[blk + 1];
                        if (gb > ga) {                 // gray pixels (0 < g < 255), SURVEY F6
                            s_scr[warp * 32 + lane] = acc;
                            __syncwarp();
                            const uint64_t* rows =
                                reinterpret_cast<const uint64_t*>(s_scr + warp * 32);
                            for (uint32_t gi = ga + lane; gi < gb; gi += 32) {
                                const uint32_t ent = t.gray_ent[gi];
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

constexpr int K2_THREADS = 256;

static size_t cov_smem_bytes(bool gray, int T, int MAXP)
{
    return (gray ? (size_t)(K2_THREADS / 32) * 32 * 16 : 0) + (size_t)T * 8 + (size_t)MAXP * 8 +
           (size_t)((MAXP * 2 + 15) / 16) * 16;
}

template <bool F, bool A, bool G>
static int launch_cov_variant(gcp_ctx* ctx, const uint32_t* edge_ids, uint32_t P, int AL, int T,
                              int maxp, size_t smem, uint32_t* cov, uint8_t* flags, cudaStream_t st)
{
    GCP_CUDA(ctx, cudaFuncSetAttribute(k_coverage<K2_THREADS, F, A, G>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t R0 = 1;                      // expected ~2.2 sub-tiles per step
    while ((size_t)R0 * maxp < (size_t)AL * 5 / 2 && R0 < 4096) R0 <<= 1;
    k_coverage<K2_THREADS, F, A, G><<<P, K2_THREADS, smem, st>>>(edge_ids, P, ctx->t, AL, T, maxp,
                                                                R0, cov, flags);
    return 0;
}

int launch_coverage(gcp_ctx* ctx, const uint32_t* edge_ids, uint32_t P, uint32_t* cov,
                    uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    const int AL = ctx->p.num_agents * ctx->p.max_dna_len;
    // pair store sized for 3 sub-tiles per step (average is ~2), table for a third of that
    int maxp = ctx->cov_maxpairs > 0 ? ctx->cov_maxpairs : 3 * AL;
    if (maxp < 256) maxp = 256;
    if (maxp > 4096) maxp = 4096;          // longer genomes run in R0 block partitions
    int T = ctx->cov_table;
    if (T <= 0) { T = K2_THREADS; while (T < maxp / 3) T <<= 1; }
    if (T > 32768) T = 32768;
    const bool gray = ctx->n_gray > 0;
    const bool need_free = ctx->cov_detail || ctx->p.coverage_blend != 1.0;
    const bool need_all = ctx->cov_detail != 0;
    const size_t smem = cov_smem_bytes(gray, T, maxp);
    if (smem > (size_t)ctx->max_smem_optin)
        return gcp_fail(ctx, -3, "coverage kernel needs %zu B shared memory", smem);
    StageTimer tm(ctx, 1, st);
    int r;
    if (gray)                     r = launch_cov_variant<true, true, true>(ctx, edge_ids, P, AL, T, maxp, smem, cov, flags, st);
    else if (need_all)            r = launch_cov_variant<true, true, false>(ctx, edge_ids, P, AL, T, maxp, smem, cov, flags, st);
    else if (need_free)           r = launch_cov_variant<true, false, false>(ctx, edge_ids, P, AL, T, maxp, smem, cov, flags, st);
    else                          r = launch_cov_variant<false, false, false>(ctx, edge_ids, P, AL, T, maxp, smem, cov, flags, st);
    if (r) return r;
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

// ============================================================================
// K2m  wall coverage from the PRE-MASKED footprint store (objectives-only fast path).
//   Every stored sub-tile is footprint AND buffer_mask, so the wall-coverage pixel count is
//   just popcount(union): no plane loads, fewer and emptier sub-tiles (the mask is a thin
//   band along walls).  Same counting-sort bucketing as k_coverage, but phase 2 walks the
//   flat sorted payload array: each warp owns a contiguous range snapped to bucket
//   boundaries, keeps 4 independent 128-bit loads in flight per lane across bucket
//   boundaries, and closes a bucket (popcount, clear) when it meets the "last" flag.
//   sorted entry = payload << 1 | last   (payload = offset128 << 4 | qmask, offset < 2^27)
// ============================================================================
constexpr int K2M_THREADS = 256;
constexpr int K2M_PPT = 12;               // pairs staged per thread (MAXP <= 12 * 256)

__global__ void __launch_bounds__(K2M_THREADS, 5)
k_coverage_masked(const uint32_t* __restrict__ step_info, uint32_t P,
                  const uint2* __restrict__ sub_desc, const uint4* __restrict__ tile, int AL,
                  int T /*pow2, multiple of THREADS*/, int MAXP /*<= PPT*THREADS*/, uint32_t R0,
                  uint32_t* __restrict__ cov, uint8_t* __restrict__ flags)
{
    constexpr int THREADS = K2M_THREADS;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem);   // [T] block id + 1
    uint32_t* s_cnt  = s_keys + T;                           // [T] count -> count<<16 | base
    uint32_t* s_pay  = s_cnt + T;                            // [MAXP] payload, then sorted entries
    uint32_t* s_sr   = s_pay + MAXP;                         // [MAXP] slot<<16 | rank
    __shared__ uint32_t s_npairs, s_overflow;
    __shared__ uint32_t s_wsum[NW];
    __shared__ uint32_t s_total;

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* info = step_info + (size_t)org * AL;
    const int logT = 31 - __clz(T);
    const uint32_t q1 = 1u + (lane >> 3), below1 = ((1u << q1) - 1u) & ~1u, l7 = lane & 7;

    uint32_t R = R0;
    bool failed = false;
    uint32_t n_cov;

    for (;;) {                                   // retry loop over partition count R
        n_cov = 0;
        bool overflow = false;
        for (uint32_t r = 0; r < R && !overflow; ++r) {
            for (int i = tid; i < T; i += THREADS) { s_keys[i] = 0; s_cnt[i] = 0; }
            if (tid == 0) { s_npairs = 0; s_overflow = 0; }
            __syncthreads();
            // ---- phase 1: slot + rank for every sub-tile ----
            for (int s0 = warp * 32; s0 < AL; s0 += THREADS) {
                const int s = s0 + lane;
                uint32_t lo = 0, n = 0;
                if (s < AL) {
                    const uint32_t inf = __ldg(info + s);
                    lo = inf & 0x0FFFFFFFu;
                    n = inf >> 28;
                }
                // fetch the first four descriptors together (one L2 round trip)
                uint2 d[4];
                #pragma unroll
                for (int k = 0; k < 4; ++k)
                    d[k] = (k < (int)n) ? __ldg(sub_desc + lo + k) : make_uint2(0, 0);
                uint32_t mine = n;
                if (R > 1) {
                    mine = 0;
                    for (uint32_t k = 0; k < n; ++k) {
                        const uint32_t blk = (k == 0) ? d[0].x : (k == 1) ? d[1].x : (k == 2) ? d[2].x
                                           : (k == 3) ? d[3].x : __ldg(sub_desc + lo + k).x;
                        mine += ((part_block(blk) & (R - 1)) == r);
                    }
                }
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
                for (uint32_t k = 0; k < n; ++k) {
                    uint2 dsc;
                    if (k == 0) dsc = d[0]; else if (k == 1) dsc = d[1];
                    else if (k == 2) dsc = d[2]; else if (k == 3) dsc = d[3];
                    else dsc = __ldg(sub_desc + lo + k);
                    const uint32_t blk = dsc.x;
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
                    const uint32_t rank = atomicAdd(&s_cnt[h], 1u);
                    GCP_DASSERT(h < (uint32_t)T && rank < 0x10000u);
                    s_pay[idx] = dsc.y;
                    s_sr[idx] = (h << 16) | rank;
                    ++idx;
                }
            }
            __syncthreads();
            overflow = (s_overflow != 0);
            if (overflow) { __syncthreads(); break; }
            const uint32_t npairs = s_npairs;
            // ---- exclusive scan of the slot counts ----
            {
                const int per = T / THREADS;
                uint32_t loc = 0;
                for (int k = 0; k < per; ++k) loc += s_cnt[k * THREADS + tid];
                uint32_t inc = loc;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                uint32_t woff = 0;
                for (int w2 = 0; w2 < warp; ++w2) woff += s_wsum[w2];
                uint32_t run = woff + inc - loc;
                for (int k = 0; k < per; ++k) {
                    const uint32_t c = s_cnt[k * THREADS + tid];
                    s_cnt[k * THREADS + tid] = (c << 16) | run;
                    run += c;
                }
            }
            __syncthreads();
            // ---- place through registers so the sorted array can reuse s_pay ----
            uint32_t ent[K2M_PPT], pos[K2M_PPT];
            #pragma unroll
            for (int k = 0; k < K2M_PPT; ++k) {
                const uint32_t i = tid + k * THREADS;
                pos[k] = 0xFFFFFFFFu;
                if (i < npairs) {
                    const uint32_t sr = s_sr[i];
                    const uint32_t cb = s_cnt[sr >> 16];
                    const uint32_t rank = sr & 0xFFFFu;
                    pos[k] = (cb & 0xFFFFu) + rank;
                    ent[k] = (s_pay[i] << 1) | (rank + 1 == (cb >> 16) ? 1u : 0u);
                }
            }
            __syncthreads();
            #pragma unroll
            for (int k = 0; k < K2M_PPT; ++k)
                if (pos[k] != 0xFFFFFFFFu) s_pay[pos[k]] = ent[k];
            __syncthreads();
            // ---- phase 2: each warp takes an equal share of the flat array, snapped to
            //      bucket boundaries (a range starts right after a "last" entry) ----
            uint32_t bounds[2];
            #pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t w = warp + e;
                uint32_t b = (uint32_t)(((uint64_t)npairs * w) / NW);
                if (w == 0) b = 0;
                else if (w >= (uint32_t)NW) b = npairs;
                else {
                    // first position >= b whose predecessor is a "last" entry
                    uint32_t c = b;
                    for (;;) {
                        const uint32_t pp = c + lane;       // predecessor index pp - 1
                        const bool is_start = (pp >= npairs) ||
                                              (pp == 0) || ((s_pay[pp - 1] & 1u) != 0);
                        const uint32_t m = __ballot_sync(FULL, is_start);
                        if (m) { c += __ffs(m) - 1; break; }
                        c += 32;
                    }
                    b = min(c, npairs);
                }
                bounds[e] = b;
            }
            // acc = (acc & keep) | v : keep = 0 right after a bucket was closed (one LOP3 per
            // word, no clearing); a bucket is closed by a warp-uniform branch on its last entry.
            uint4 acc = make_uint4(0, 0, 0, 0);
            uint32_t keep = 0;
            const uint4* tile_l = tile + l7;
            uint32_t qq = q1, bl = below1;
            // opaque copies: stop ptxas from re-deriving the lane constants inside the loop
            asm volatile("" : "+l"(tile_l), "+r"(qq), "+r"(bl));
            for (uint32_t c0 = bounds[0]; c0 < bounds[1]; c0 += 32) {
                uint32_t my = 0;
                if (c0 + lane < bounds[1]) my = s_pay[c0 + lane];
                const uint32_t lastm = __ballot_sync(FULL, my & 1u);     // warp-uniform
                const uint32_t nn = min(32u, bounds[1] - c0);
                for (uint32_t k = 0; k < nn; k += 4) {
                    uint32_t p[4];
                    uint4 v[4];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) p[j] = __shfl_sync(FULL, my, k + j);
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j] = make_uint4(0, 0, 0, 0);
                        if ((p[j] >> qq) & 1u)
                            v[j] = __ldg(tile_l + ((((p[j] >> 5) + __popc(p[j] & bl))) << 3));
                    }
                    const uint32_t lm = (lastm >> k) & 15u;
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc.x = (acc.x & keep) | v[j].x;
                        acc.y = (acc.y & keep) | v[j].y;
                        acc.z = (acc.z & keep) | v[j].z;
                        acc.w = (acc.w & keep) | v[j].w;
                        keep = 0xFFFFFFFFu;
                        if (lm & (1u << j)) {          // uniform: close the bucket
                            n_cov += popc128(acc);
                            keep = 0;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (!overflow) break;
        R <<= 1;
        if (R > 4096) { failed = true; break; }
    }

    if (tid == 0) s_total = 0;
    __syncthreads();
    uint32_t v = n_cov;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0 && v) atomicAdd(&s_total, v);
    __syncthreads();
    if (tid < 8) {
        const uint32_t n = failed ? 0u : s_total;
        cov[(size_t)org * 8 + tid] = (tid == 0 || tid == 4) ? n : 0u;
    }
    if (tid == 0 && failed && flags) flags[(size_t)org * 4 + 3] |= 2;
}

int launch_coverage_masked(gcp_ctx* ctx, const uint32_t* step_info, uint32_t P, uint32_t* cov,
                           uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    if (!ctx->have_masked) return gcp_fail(ctx, -1, "masked footprint store not uploaded");
    const int AL = ctx->p.num_agents * ctx->p.max_dna_len;
    int maxp = ctx->cov_maxpairs > 0 ? ctx->cov_maxpairs : 2 * AL;
    if (maxp < 256) maxp = 256;
    if (maxp > K2M_PPT * K2M_THREADS) maxp = K2M_PPT * K2M_THREADS;
    int T = ctx->cov_table;
    if (T <= 0) { T = K2M_THREADS; while (T < maxp / 5) T <<= 1; }
    if (T > 32768) T = 32768;
    const size_t smem = (size_t)T * 8 + (size_t)maxp * 8;
    if (smem > (size_t)ctx->max_smem_optin)
        return gcp_fail(ctx, -3, "masked coverage kernel needs %zu B shared memory", smem);
    GCP_CUDA(ctx, cudaFuncSetAttribute(k_coverage_masked,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(ctx, 1, st);
    uint32_t R0 = 1;                      // expected ~1.1 masked sub-tiles per step
    while ((size_t)R0 * maxp < (size_t)AL + AL / 2 && R0 < 4096) R0 <<= 1;
    k_coverage_masked<<<P, K2M_THREADS, smem, st>>>(step_info, P, ctx->t.m_sub_desc, ctx->t.m_tile,
                                                    AL, T, maxp, R0, cov, flags);
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
//   Most keys are unique, so instead of sorting everything the keys go through a
//   shared-memory hash table (64-bit CAS); only members of groups with count >= 2 are
//   linked per slot, and every member finds its successor (next larger position in the
//   group) by walking its group's list - the "consecutive ascending gap" test of the
//   stable-argsort canonical order without a sort.
// ============================================================================
constexpr int K3_THREADS = 256;
constexpr uint32_t K3_NIL = 0xFFFFFFFFu;

__global__ void __launch_bounds__(K3_THREADS)
k_loop_closures(const int32_t* __restrict__ dna, uint32_t P, int A, int L, int T3 /*slots > A*L*/,
                gcp_params prm, gcp_map_consts mc,
                const double* __restrict__ dist, const double* __restrict__ turn,
                const uint32_t* __restrict__ cov,
                int32_t* __restrict__ lc_out, double* __restrict__ obj,
                double* __restrict__ neg_cov_out, uint8_t* __restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem[];
    // per slot: key (u64; after the count pass reused as {head, owners}), count|flags
    unsigned long long* s_hkey = reinterpret_cast<unsigned long long*>(smem);     // [T3]
    uint32_t* s_hcnt = reinterpret_cast<uint32_t*>(s_hkey + T3);                   // [T3]
    int32_t*  s_seq  = reinterpret_cast<int32_t*>(s_hcnt + T3);                    // [A*L]
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_seq + A * L);                 // [A*L]
    uint16_t* s_next = s_slot + A * L;                                             // [A*L]
    uint8_t*  s_own  = reinterpret_cast<uint8_t*>(s_next + A * L);                 // [A*L]
    __shared__ int s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    __shared__ int s_lc[GCP_MAX_AGENTS * GCP_MAX_AGENTS];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = K3_THREADS / 32;
    const int32_t* d = dna + (size_t)org * A * L;

    for (int i = tid; i < A * A; i += K3_THREADS) s_lc[i] = 0;
    for (int i = tid; i < T3; i += K3_THREADS) { s_hkey[i] = 0ull; s_hcnt[i] = 0u; }
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
    // ---- pass 1: hash every key, count group sizes ----
    for (int i = tid; i < n; i += K3_THREADS) {
        const unsigned long long f = (uint32_t)s_seq[i] & 0xFFFFFFu;
        const unsigned long long g = (uint32_t)s_seq[(i + 1 == n) ? 0 : i + 1] & 0xFFFFFFu;
        const unsigned long long key = ((f << 24) | g) + 1ull;              // never 0
        // fast range reduction of the top 32 hash bits onto [0, T3)
        uint32_t h = __umulhi((uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32), (uint32_t)T3);
        for (;;) {
            const unsigned long long cur = ((volatile unsigned long long*)s_hkey)[h];
            if (cur == key) break;
            if (cur == 0ull) {
                const unsigned long long old = atomicCAS(&s_hkey[h], 0ull, key);
                if (old == 0ull || old == key) break;
            }
            h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;    // T3 > n: always terminates
        }
        atomicAdd(&s_hcnt[h], 1u);
        s_slot[i] = (uint16_t)h;
    }
    __syncthreads();
    // ---- pass 2: slots with a duplicate group become {head = NIL, owners = 0} ----
    for (int h = tid; h < T3; h += K3_THREADS)
        if (s_hcnt[h] >= 2u) s_hkey[h] = (unsigned long long)K3_NIL;      // lo = head, hi = owners
    __syncthreads();
    uint32_t* s_w = reinterpret_cast<uint32_t*>(s_hkey);                  // [2*h] head, [2*h+1] owners
    for (int i = tid; i < n; i += K3_THREADS) {
        const uint32_t h = s_slot[i];
        if ((s_hcnt[h] & 0xFFFFu) < 2u) continue;
        const uint8_t o = s_own[i];
        s_next[i] = (uint16_t)atomicExch(&s_w[2 * h], (uint32_t)i);       // NIL -> 0xFFFF
        atomicOr(&s_w[2 * h + 1], 1u << (o & 31));
        if (o & 0x80) atomicOr(&s_hcnt[h], 1u << 16);                     // group touches a split (F10)
    }
    __syncthreads();
    // ---- pass 3: successor gap test (ascending consecutive positions) ----
    const int thresh = prm.solo_sep_thresh;
    for (int i = tid; i < n; i += K3_THREADS) {
        const uint32_t h = s_slot[i];
        const uint32_t c = s_hcnt[h];
        if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
        uint32_t succ = 0xFFFFFFFFu;
        for (uint32_t j = s_w[2 * h]; j != 0xFFFFu && j != K3_NIL; j = s_next[j])
            if (j > (uint32_t)i && j < succ) succ = j;
        if (succ != 0xFFFFFFFFu && (int)(succ - (uint32_t)i) > thresh)
            atomicOr(&s_hcnt[h], 1u << 17);
    }
    __syncthreads();
    // ---- pass 4: one thread per duplicate group updates lc_mat ----
    for (int h = tid; h < T3; h += K3_THREADS) {
        const uint32_t c = s_hcnt[h];
        if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
        const uint32_t owners = s_w[2 * h + 1];
        if (__popc(owners) == 1) {
            if (c & (1u << 17)) { const int a = __ffs(owners) - 1; atomicAdd(&s_lc[a * A + a], 1); }
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

    // ---- constraints + objectives (geneticalgorithm.py:177-201), one warp cooperatively ----
    if (tid < 32)
        finish_objectives(s_lc, A, prm, mc, cov + (size_t)org * 8, dist + (size_t)org * A,
                          turn + (size_t)org * A, org, obj, neg_cov_out, flags);
}

int launch_loop_closures(gcp_ctx* ctx, const int32_t* dna, uint32_t P, const double* dist,
                         const double* turn, const uint32_t* cov, int32_t* lc, double* obj,
                         double* neg_cov, uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    const int A = ctx->p.num_agents, L = ctx->p.max_dna_len;
    int T3 = (A * L * 3) / 2 + 64;            // load factor <= 2/3, never full; ids fit 16 bits
    const size_t smem = (size_t)T3 * 12 + (size_t)A * L * 9 + 16;
    if (T3 > 65535 || smem > (size_t)ctx->max_smem_optin || ctx->force_lc_big)   // 1: radix sort, 2: partitioned hash
        return launch_loop_closures_big(ctx, dna, P, dist, turn, cov, lc, obj, neg_cov, flags, st);
    GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(ctx, 2, st);
    k_loop_closures<<<P, K3_THREADS, smem, st>>>(dna, P, A, L, T3, ctx->p, ctx->mc, dist, turn,
                                                 cov, lc, obj, neg_cov, flags);
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
