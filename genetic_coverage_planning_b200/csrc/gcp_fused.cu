// Fused evaluation kernel (objectives-only fast path): Organism.calcObj for one organism per
// CTA with the chromosome and the quarter-item range of every step held in shared memory, so the
// DNA is read from HBM exactly once and nothing intermediate is written back.
//
//   phase A  path costs        mappy.py:341-352   (edge lookup, per-step costs, revisited edges dropped)
//   phase B  wall coverage     mappy.py:353-364   (coverage_quarters: counting sort of quarter items
//                                                  by (map block, quarter), flat union + popcount)
//   phase C  loop closures     mappy.py:372-437   (shared-memory hash of consecutive gene pairs, one
//                                                  record per duplicate group)
//   phase D  constraints + objectives  geneticalgorithm.py:177-201
//
// Worlds with at most 65,534 waypoints run the COMPACT variant: 16-bit genes and sequence in
// shared memory, 32-bit pair keys, 40 registers -> six resident CTAs per SM instead of five, and
// phase C starts without a barrier while other warps still stream their B2 ranges.
//
// Objectives-only variant: pre-masked footprint store uploaded and coverage_blend == 1 (wall coverage
// only); binary and gray maps (HAS_GRAY).  GENERAL variant: any coverage_blend and the six pixel
// counters, over the unmasked store with marker entries for the map planes (see coverage_quarters).
// The staged kernels in gcp_eval.cu are the cross-check of both (tests/test_gpu_parity.py).  Genomes
// that outgrow this kernel's shared memory run phase B stand-alone: k_cov_quarters (one CTA per
// organism, partitions one after the other) or, by default, k_cov_bucket + k_cov_part (the partitions
// of all organisms as work items of small CTAs).
//
#include "gcp_common.cuh"
#include "gcp_device.cuh"
#include <type_traits>

#ifndef KF_A1_U
#define KF_A1_U 1        // steps per thread whose table loads are issued together in phase A1
#endif
constexpr int KF_PPT = 12;               // items (quarters) staged per thread: MAXP <= 12 * threads
constexpr uint32_t KF_NIL = 0xFFFFFFFFu;

struct FusedShape {
    int A, L, AL;
    uint32_t inv_L;      // ceil(2^32 / L): s / L == __umulhi(s, inv_L) for s < 2^16
    int T, MAXP, T3;
    int gene16;          // chromosome stored as int16 in global memory
    int prefetch;        // footprint store outgrows the L2: bit 0 = B1 requests every quarter line into the L2 while it
                         // sorts the items, bit 1 = B2 requests the lines KF_PF_DIST entries ahead of its stream (PF kernels)
    uint32_t R0;
    int c_overlap;       // loop-closure tables (32-bit keys) fit in front of the sorted quarter array
    uint32_t DS;         // slots of the step-dedup hash set (power of two, 0 = skip)
    uint32_t off_info, off_x, off_cst, off_lc, off_cost;   // byte offsets into dynamic smem (off_cst relative to off_x)
};

// ============================================================================
// Wall coverage of one organism (CTA-wide): see "phase B" in the header comment of k_eval_fused.
//   info[s], s < AL: quarter-item range of step s (first | count << 26), 0 = no step.
//   n_cov: covered mask pixels with g == 0, n_wg: sum of (255 - g) over covered gray mask pixels
//   (per-thread partial sums; the caller reduces them).  s_x: T*20 + MAXP*8 + 32 bytes.
// ============================================================================
// base + idx * 16 in ONE wide multiply-add (the compiler otherwise splits the 36-bit product into
// shift / funnel-shift / mask / add / add-with-carry in front of every footprint load)
__device__ __forceinline__ const uint4* tile_at(const uint4* base, uint32_t idx)
{
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(idx), "l"(base));
    return reinterpret_cast<const uint4*>(addr);
}

// base + idx * 128: a whole quarter (8 x uint4) per index
__device__ __forceinline__ const uint4* quarter_at(const uint4* base, uint32_t idx)
{
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(addr) : "r"(idx), "l"(base));
    return reinterpret_cast<const uint4*>(addr);
}

// Per-thread partial pixel counts of one organism (the caller reduces them over the CTA).
//   objectives-only (pre-masked store): n_cov = covered mask pixels with g == 0, n_wg = sum of (255 - g)
//   over covered gray mask pixels.
//   GENERAL (unmasked store): n_cov = n_mask0 and n_wg = wg_mask as above, n_free0 = covered pixels with
//   g == 0, wg_all = sum of (255 - g) over all covered gray pixels, n_gmask = covered gray mask pixels,
//   n_all = every covered pixel (occluded ones included) - the six counters of k_coverage (gcp_eval.cu).
struct CovCounts { uint32_t n_cov, n_wg, n_free0, wg_all, n_gmask, n_all; };

// GENERAL: the items come from the UNMASKED store (GcpTables::q_info / q_desc / tile) and the sorted
// array carries, behind every (block, quarter) list, two MARKER entries {tag, block << 2 | quarter}: the
// first makes the group AND the finished union with the (g == 0) plane of that quarter, the second with
// the mask & (g == 0) plane - loaded like any other 128-B line of the stream, four loads in flight.
// Gray pixels are weighed at the second marker from the per-quarter gray lists.
#ifndef KF_GEN_MINB
#define KF_GEN_MINB 3     // resident CTAs per SM the general kernel (512 threads) is compiled for
#endif
#ifndef KF_PF_DIST_N
#define KF_PF_DIST_N 16
#endif
constexpr uint32_t KF_PF_DIST = KF_PF_DIST_N;      // B2 look-ahead of the L2 prefetch, in entries of a group's range (multiple of 8)

template <int KF_THREADS, bool HAS_GRAY, bool BUCKETS = false, bool GENERAL = false, bool PF = false>
__device__ __forceinline__ void coverage_quarters(uint32_t* info, int AL, const GcpTables& t,
                                                  int T, int MAXP, uint32_t R0, uint32_t DS, unsigned char* s_x,
                                                  CovCounts& cc, bool& failed,
                                                  uint2* rec = nullptr, uint32_t rec_cap = 0, int prefetch = 0,
                                                  bool tail_barrier = true, uint32_t pre_m = 0xFFFFFFFFu)
{
    // pre_m (BUCKETS only): `rec` already holds the pre_m item descriptors of ONE hash partition of the
    // organism (k_cov_bucket wrote them): a single pass over them, no dedup, no retry - an overflow of the
    // item store or of the block table is reported through `failed` and the caller redoes the organism.
    constexpr int KF_NW = KF_THREADS / 32;
    constexpr uint32_t NM = GENERAL ? 2u : 0u;                 // marker entries behind every list
    const uint2* const q_desc = GENERAL ? t.q_desc : t.m_q_desc;
    const uint4* const q_tile = GENERAL ? t.tile : t.m_tile;
    uint32_t n_cov = 0, n_wg = 0, n_free0 = 0, wg_all = 0, n_gmask = 0, n_all = 0;
    __shared__ uint32_t s_npairs, s_overflow, s_group;
    __shared__ uint32_t s_wsum[KF_NW];
    __shared__ uint32_t s_pcnt[64], s_pbase[65];               // item buckets per hash partition (long genomes)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_x);      // [T] block id + 1
    uint32_t* s_cq   = s_keys + T;                             // [T][4] count -> count<<16 | base
    uint32_t* s_sr   = s_cq + 4 * T;                           // [MAXP] slot<<18 | q<<16 | rank
    uint32_t* s_pay  = s_sr + MAXP;                            // [MAXP] uint4 index of the quarter, then sorted
    // (s_pay last: B2 reads nothing else, so the caller may start reusing everything in front of it
    //  as soon as it has passed the barrier in front of B2 - `tail_barrier` false)
    const int logT = 31 - __clz(T);
    const uint32_t myq = lane >> 3, l7 = lane & 7;

    // ---- B0: a walk revisits edges (18 % of the steps at 150-step walks, half of them at 512);
    //      the union does not care, so revisits are dropped: one exchange per step into a table of the
    //      info words over the first DS words of s_x - a step that finds its own item list in its slot
    //      is a revisit (DS = 0: table too small, skip).  No probing: a slot shared with another edge may
    //      let a duplicate through, which costs time, never correctness.
    if (DS) {
        uint32_t* s_set = reinterpret_cast<uint32_t*>(s_x);
        const int sh_ds = __clz(DS) + 1;                       // 32 - log2(DS)
        for (int i = tid; i < (int)DS; i += KF_THREADS) s_set[i] = 0u;
        __syncthreads();
        for (int s = tid; s < AL; s += KF_THREADS) {
            const uint32_t inf = info[s];
            if (!inf) continue;
            uint32_t h = (inf * 0x9E3779B1u) >> sh_ds;
            if (atomicExch(&s_set[h], inf) == inf) info[s] = 0u;      // see k_eval_fused phase A1
        }
        __syncthreads();
    }

    uint32_t R = R0;
    failed = false;
    for (;;) {                                   // retry loop over partition count R
        n_cov = 0; n_wg = 0; n_free0 = 0; wg_all = 0; n_gmask = 0; n_all = 0;
        bool overflow = false;
        // ---- long genomes (R > 1) with a scratch array: bucket the item descriptors by hash
        //      partition ONCE (count, prefix, scatter; lanes that hit the same partition share
        //      one shared atomic), so that every partition pass below reads its own items
        //      linearly instead of re-scanning and filtering all of them
        bool bucketed = false;
        if (BUCKETS && pre_m != 0xFFFFFFFFu) {
            if (tid == 0) { s_pbase[0] = 0; s_pbase[1] = pre_m; }
            bucketed = true;
            __syncthreads();
        }
        if (BUCKETS && pre_m == 0xFFFFFFFFu && R > 1 && R <= 64 && rec != nullptr) {
            auto scan_items = [&](auto&& fn) {
                for (int s0 = warp * 32; s0 < AL; s0 += KF_NW * 32) {
                    const int s = s0 + lane;
                    uint32_t lo = 0, n = 0;
                    if (s < AL) { const uint32_t inf = info[s]; lo = inf & 0x03FFFFFFu; n = inf >> 26; }
                    uint32_t incl = n;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    const uint32_t excl = incl - n;
                    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        int a = 0, b = 31;
                        #pragma unroll
                        for (int it = 0; it < 5; ++it) {
                            const int mid = (a + b) >> 1;
                            const uint32_t v = __shfl_sync(FULL, incl, mid);
                            if (v > j) b = mid; else a = mid + 1;
                        }
                        a = min(a, 31);
                        const uint32_t ex = __shfl_sync(FULL, excl, a), dlo = __shfl_sync(FULL, lo, a);
                        uint2 dsc = make_uint2(0, 0);
                        if (j < total) dsc = __ldg(q_desc + dlo + (j - ex));
                        fn(j < total, dsc);
                    }
                }
            };
            if (tid < 64) s_pcnt[tid] = 0;
            __syncthreads();
            scan_items([&](bool valid, uint2 dsc) {
                const uint32_t p = valid ? (part_block(dsc.x >> 2) & (R - 1)) : KF_NIL;
                const uint32_t peers = __match_any_sync(FULL, p);
                if (valid && lane == (int)(__ffs(peers) - 1)) atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
            });
            __syncthreads();
            if (tid == 0) {
                uint32_t b = 0;
                for (uint32_t r = 0; r < R; ++r) { s_pbase[r] = b; b += s_pcnt[r]; s_pcnt[r] = 0; }
                s_pbase[R] = b;
            }
            __syncthreads();
            if (s_pbase[R] <= rec_cap) {         // else: more items than the scratch holds -> scanning passes
                scan_items([&](bool valid, uint2 dsc) {
                    const uint32_t p = valid ? (part_block(dsc.x >> 2) & (R - 1)) : KF_NIL;
                    const uint32_t peers = __match_any_sync(FULL, p);
                    const int leader = __ffs(peers) - 1;
                    uint32_t base = 0;
                    if (valid && lane == leader) base = atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
                    base = __shfl_sync(FULL, base, leader);
                    GCP_DASSERT(!valid || s_pbase[p] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u)) < rec_cap);
                    if (valid) rec[s_pbase[p] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = dsc;
                });
                bucketed = true;
            }
            __syncthreads();                     // the CTA's own global writes are visible to it
        }
        for (uint32_t r = 0; r < R && !overflow; ++r) {
            for (int i = tid; i < T; i += KF_THREADS) s_keys[i] = 0;
            for (int i = tid; i < 4 * T; i += KF_THREADS) s_cq[i] = 0;
            if (tid == 0) { s_npairs = 0; s_overflow = 0; s_group = 0; }
            __syncthreads();
            // ---- B1: slot + rank for every quarter item ----
            if (BUCKETS && bucketed) {
                const uint2* rc = rec + s_pbase[r];
                const uint32_t m = s_pbase[r + 1] - s_pbase[r];
                if (m > (uint32_t)MAXP) {
                    if (tid == 0) s_overflow = 1;
                } else {
                    for (uint32_t j0 = tid; j0 < m; j0 += 4 * KF_THREADS) {
                      uint2 d4[4];                                          // four descriptor loads in flight
                      #pragma unroll
                      for (int u = 0; u < 4; ++u) { const uint32_t j = j0 + u * KF_THREADS; d4[u] = (j < m) ? rc[j] : make_uint2(0, 0); }
                      #pragma unroll
                      for (int u = 0; u < 4; ++u) {
                        const uint32_t j = j0 + u * KF_THREADS;
                        if (j >= m) break;
                        const uint2 dsc = d4[u];
                        const uint32_t blk = dsc.x >> 2, q = dsc.x & 3u;
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
                        const uint32_t rank = (probes < T) ? atomicAdd(&s_cq[4 * h + q], 1u) : 0u;
                        GCP_DASSERT(j < (uint32_t)MAXP && h < (uint32_t)T);
                        s_pay[j] = dsc.y;
                        s_sr[j] = (h << 18) | (q << 16) | rank;
                      }
                    }
                    if (tid == 0) s_npairs = m;
                }
            } else {
                const int bw = warp, nbw = KF_NW;
                for (int s0 = bw * 32; s0 < AL; s0 += nbw * 32) {
                    const int s = s0 + lane;
                    uint32_t lo = 0, n = 0;
                    if (s < AL) {
                        const uint32_t inf = info[s];
                        lo = inf & 0x03FFFFFFu;
                        n = inf >> 26;
                    }
                    if (R == 1) {
                        // Flattened: the warp's 32 steps hold `total` items; lane j of round j0 takes
                        // item j0 + j (its step is found by bisection over the lanes' prefix sums), so
                        // every lane is busy whatever the spread of items per step.
                        uint32_t incl = n;
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
                        const uint32_t excl = incl - n;
                        for (uint32_t j0 = 0; j0 < total; j0 += 32) {
                            const uint32_t j = j0 + lane;
                            int a = 0, b = 31;                        // smallest lane with incl > j
                            #pragma unroll
                            for (int it = 0; it < 5; ++it) {
                                const int mid = (a + b) >> 1;
                                const uint32_t v = __shfl_sync(FULL, incl, mid);
                                if (v > j) b = mid; else a = mid + 1;
                            }
                            a = min(a, 31);
                            const uint32_t ex = __shfl_sync(FULL, excl, a), dlo = __shfl_sync(FULL, lo, a);
                            if (j < total) {
                                const uint2 dsc = __ldg(q_desc + dlo + (j - ex));
                                // store larger than the L2 (config C5, diverse populations): start the
                                // DRAM fetch of this quarter's line now, B2 finds it in the L2
                                if (prefetch & 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(GENERAL ? quarter_at(q_tile, dsc.y) : tile_at(q_tile, dsc.y)));
                                const uint32_t blk = dsc.x >> 2, q = dsc.x & 3u;
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
                                if (probes < T) {
                                    GCP_DASSERT(h < (uint32_t)T && base + j < (uint32_t)MAXP);
                                    const uint32_t rank = atomicAdd(&s_cq[4 * h + q], 1u);
                                    GCP_DASSERT(rank < 0x10000u);
                                    s_pay[base + j] = dsc.y;
                                    s_sr[base + j] = (h << 18) | (q << 16) | rank;
                                }
                            }
                        }
                        continue;
                    }
                    // R > 1 (long genomes): same flattening; every round keeps the items of partition r
                    // and compacts them with a ballot (one shared atomic per round of 32 items)
                    uint32_t incl = n;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    const uint32_t excl = incl - n;
                    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        int a = 0, b = 31;
                        #pragma unroll
                        for (int it = 0; it < 5; ++it) {
                            const int mid = (a + b) >> 1;
                            const uint32_t v = __shfl_sync(FULL, incl, mid);
                            if (v > j) b = mid; else a = mid + 1;
                        }
                        a = min(a, 31);
                        const uint32_t ex = __shfl_sync(FULL, excl, a), dlo = __shfl_sync(FULL, lo, a);
                        uint2 dsc = make_uint2(0, 0);
                        bool mine = false;
                        if (j < total) {
                            dsc = __ldg(q_desc + dlo + (j - ex));
                            mine = (part_block(dsc.x >> 2) & (R - 1)) == r;
                        }
                        const uint32_t bal = __ballot_sync(FULL, mine);
                        if (!bal) continue;
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&s_npairs, (uint32_t)__popc(bal));
                        base = __shfl_sync(FULL, base, 0);
                        if (base + (uint32_t)__popc(bal) > (uint32_t)MAXP) {
                            if (lane == 0) s_overflow = 1;
                            break;
                        }
                        if (mine) {
                            const uint32_t idx = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                            const uint32_t blk = dsc.x >> 2, q = dsc.x & 3u;
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
                            // a failed probe still fills its entry: the pass is discarded anyway
                            const uint32_t rank = (probes < T) ? atomicAdd(&s_cq[4 * h + q], 1u) : 0u;
                            GCP_DASSERT(idx < (uint32_t)MAXP && h < (uint32_t)T);
                            s_pay[idx] = dsc.y;
                            s_sr[idx] = (h << 18) | (q << 16) | rank;
                        }
                    }
                }
            }
            __syncthreads();
            overflow = (s_overflow != 0);
            if (overflow) { __syncthreads(); break; }
            const uint32_t npairs = s_npairs;
            uint32_t n_sorted = npairs;            // entries of the sorted array (GENERAL: items + markers)
            // ---- exclusive scan of the (slot, quarter) counts ----
            //      (thread t owns counters t, t + THREADS, ...: consecutive lanes hit consecutive banks; the
            //      order in which the lists end up in the sorted array is irrelevant, only that every
            //      (slot, quarter) gets its own contiguous range)
            {
                const int per = 4 * T / KF_THREADS;
                uint32_t loc = 0;
                for (int k = 0; k < per; ++k) { const uint32_t c = s_cq[k * KF_THREADS + tid]; loc += c + (c ? NM : 0u); }
                uint32_t inc = loc;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                uint32_t woff = 0, wall = 0;
                for (int w2 = 0; w2 < KF_NW; ++w2) { if (w2 == warp) woff = wall; wall += s_wsum[w2]; }
                if (GENERAL) {
                    // the markers share the sorted array with the items: more entries than it holds -> more partitions
                    if (wall > (uint32_t)MAXP) { overflow = true; __syncthreads(); break; }
                    n_sorted = wall;
                }
                uint32_t run = woff + inc - loc;
                for (int k = 0; k < per; ++k) {
                    const uint32_t c = s_cq[k * KF_THREADS + tid];
                    s_cq[k * KF_THREADS + tid] = (c << 16) | run;
                    run += c + (c ? NM : 0u);
                }
            }
            __syncthreads();
            // ---- place through registers so the sorted array can reuse s_pay ----
            uint32_t ent[KF_PPT], pos[KF_PPT];
            #pragma unroll
            for (int k = 0; k < KF_PPT; ++k) {
                const uint32_t i = tid + k * KF_THREADS;
                pos[k] = KF_NIL;
                if (i < npairs) {
                    const uint32_t sr = s_sr[i];
                    const uint32_t cb = s_cq[sr >> 16], rank = sr & 0xFFFFu;
                    pos[k] = (cb & 0xFFFFu) + rank;
                    GCP_DASSERT((sr >> 16) < 4u * (uint32_t)T && pos[k] + ((rank + 1u == (cb >> 16)) ? NM : 0u) < n_sorted);
                    if (GENERAL) {
                        ent[k] = s_pay[i];
                        if (rank + 1u == (cb >> 16)) pos[k] |= 0x80000000u;                  // last of its list: the markers follow
                    } else {
                        ent[k] = s_pay[i] | ((rank + 1u == (cb >> 16)) ? 0x80000000u : 0u);   // last of its list
                    }
                }
            }
            __syncthreads();
            #pragma unroll
            for (int k = 0; k < KF_PPT; ++k)
                if (pos[k] != KF_NIL) {
                    if (GENERAL) {
                        const uint32_t p = pos[k] & 0x7FFFFFFFu;
                        s_pay[p] = ent[k];
                        if (pos[k] >> 31) {
                            const uint32_t sr = s_sr[tid + k * KF_THREADS];
                            const uint32_t bq = ((s_keys[sr >> 18] - 1u) << 2) | ((sr >> 16) & 3u);
                            s_pay[p + 1] = 0x80000000u | (t.g_free_q + bq);   // AND with the (g == 0) plane
                            s_pay[p + 2] = 0xC0000000u | (t.g_mask_q + bq);   // AND with mask & (g == 0), gray weights, list ends
                        }
                    } else {
                        s_pay[pos[k]] = ent[k];
                    }
                }
            __syncthreads();
            // ---- B2 (gray maps): owner warps, groups of 32 slots handed out dynamically ----
            // per-lane tile base kept in registers: the opaque asm stops ptxas from re-deriving
            // it (S2R tid + LDC of the table pointer) in front of every load
            const uint4* tile_l = q_tile + l7;
            asm volatile("" : "+l"(tile_l));
            if (GENERAL) {
                // ---- B2 (flat, general): as below, over items + markers.  An entry is a quarter of the
                //      unmasked store (OR into the registers) or a marker {tag, block << 2 | quarter}, whose
                //      128-B line is the (g == 0) plane (tag 10) or the mask & (g == 0) plane (tag 11, list end)
                //      of that quarter of the map: the group ANDs its finished union with it and counts.
                constexpr uint32_t NG = KF_NW * 4;
                const uint32_t gid = (uint32_t)warp * 4u + myq;
                const uint32_t gbit = 0xFFu << (myq * 8);
                uint32_t bnd[2];
                #pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t w = gid + e;
                    uint32_t c = (uint32_t)(((uint64_t)n_sorted * w) / NG);
                    bool done = (w == 0) || (w >= NG);
                    if (w >= NG) c = n_sorted;
                    while (__any_sync(FULL, !done)) {               // first position >= c that starts a list
                        const uint32_t pp = c + l7;
                        const bool is_start = (pp >= n_sorted) || (pp == 0) || ((s_pay[pp - 1] >> 30) == 3u);
                        const uint32_t m = __ballot_sync(FULL, is_start) & gbit;
                        if (!done) {
                            if (m) { c += (uint32_t)(__ffs(m) - 1) - myq * 8u; done = true; }
                            else c += 8;
                        }
                    }
                    bnd[e] = min(c, n_sorted);
                }
                GCP_DASSERT(bnd[0] <= bnd[1] && bnd[1] <= n_sorted);
                const uint32_t len = bnd[1] - bnd[0];
                const uint32_t steps = __reduce_max_sync(FULL, len);
                const uint32_t* pb = s_pay + bnd[0];
                uint4 acc = make_uint4(0, 0, 0, 0);
                // entry: footprint quarter (OR) or marker - its line is the plane quarter to AND the finished
                // union with; the second marker ends the list
                auto take = [&](uint32_t e, const uint4& v) {
                    if (!(e >> 31)) {
                        acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
                    } else {
                        const uint32_t c = (uint32_t)popc128(and128(acc, v));
                        if (!(e & 0x40000000u)) {
                            n_free0 += c;
                        } else {
                            n_cov += c;
                            n_all += popc128(acc);
                            if (HAS_GRAY) {
                                // gray pixels of this quarter (CSR list {pos, 255 - g, in_mask}): lane i of the
                                // group takes entry i and fetches the row word that holds its pixel by shuffle
                                const uint32_t bq = (e & 0x3FFFFFFFu) - t.g_mask_q;
                                const uint32_t ga = __ldg(t.gray_qptr + bq), gb = __ldg(t.gray_qptr + bq + 1);
                                for (uint32_t g0 = ga; g0 < gb; g0 += 8) {
                                    const uint32_t gi = g0 + l7;
                                    const uint32_t ent = (gi < gb) ? __ldg(t.gray_ent + gi) : 0u;
                                    const uint32_t pos = ent & 4095u, row = (pos >> 6) & 15u, col = pos & 63u;
                                    const int src = (int)((myq << 3) | (row >> 1));
                                    const uint32_t wx = __shfl_sync(gbit, acc.x, src), wy = __shfl_sync(gbit, acc.y, src);
                                    const uint32_t wz = __shfl_sync(gbit, acc.z, src), ww = __shfl_sync(gbit, acc.w, src);
                                    const uint32_t word = (row & 1u) ? ((col & 32u) ? ww : wz) : ((col & 32u) ? wy : wx);
                                    if (gi < gb && ((word >> (col & 31u)) & 1u)) {
                                        const uint32_t w = (ent >> 12) & 255u;
                                        wg_all += w;
                                        if ((ent >> 20) & 1u) { n_wg += w; n_gmask += 1u; }
                                    }
                                }
                            }
                            acc = make_uint4(0, 0, 0, 0);
                        }
                    }
                };
                uint32_t k = 0;
                // the ranges of a warp's four groups are equal up to the snapping: no guards up to the shortest one
                const uint32_t steps_min = __reduce_min_sync(FULL, len) & ~3u;
                for (; k < steps_min; k += 4) {
                    uint32_t e[4];
                    uint4 v[4];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        e[j] = pb[k + j];
                        v[j] = __ldg(quarter_at(tile_l, e[j] & 0x3FFFFFFFu));
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) take(e[j], v[j]);
                }
                for (; k < steps; k += 4) {
                    uint32_t e[4];
                    uint4 v[4];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        e[j] = (k + j < len) ? pb[k + j] : 0u;
                        v[j] = make_uint4(0, 0, 0, 0);
                        if (k + j < len) v[j] = __ldg(quarter_at(tile_l, e[j] & 0x3FFFFFFFu));
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) take(e[j], v[j]);
                }
            } else if (!HAS_GRAY) {
                // ---- B2 (flat): the sorted array is cut into KF_NW * 4 contiguous ranges snapped to
                //      list boundaries; every 8-lane group streams its own range (lane = one uint4 =
                //      2 rows of the quarter), ORs into registers and popcounts at each list end.
                //      All groups stay busy whatever the spread of list lengths.
                constexpr uint32_t NG = KF_NW * 4;
                const uint32_t gid = (uint32_t)warp * 4u + myq;
                const uint32_t gbit = 0xFFu << (myq * 8);
                uint32_t bnd[2];
                #pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t w = gid + e;
                    uint32_t c = (uint32_t)(((uint64_t)npairs * w) / NG);
                    bool done = (w == 0) || (w >= NG);
                    if (w >= NG) c = npairs;
                    while (__any_sync(FULL, !done)) {               // first position >= c that starts a list
                        const uint32_t pp = c + l7;
                        const bool is_start = (pp >= npairs) || (pp == 0) || (s_pay[pp - 1] >> 31);
                        const uint32_t m = __ballot_sync(FULL, is_start) & gbit;
                        if (!done) {
                            if (m) { c += (uint32_t)(__ffs(m) - 1) - myq * 8u; done = true; }
                            else c += 8;
                        }
                    }
                    bnd[e] = min(c, npairs);
                }
                GCP_DASSERT(bnd[0] <= bnd[1] && bnd[1] <= npairs);
                const uint32_t len = bnd[1] - bnd[0];
                const uint32_t steps = __reduce_max_sync(FULL, len);
                const uint32_t* pb = s_pay + bnd[0];
                uint4 acc = make_uint4(0, 0, 0, 0);
                uint32_t k = 0;
                // the ranges of a warp's four groups are equal up to the snapping: no guards up to the
                // shortest one
                const uint32_t steps_min = __reduce_min_sync(FULL, len) & ~3u;
                if (PF) {
                    // store larger than the L2: the group requests the lines of its next KF_PF_DIST entries
                    // now and keeps that distance in the loop (lane j < 4 takes entry k + KF_PF_DIST + j), so a
                    // line is on its way from DRAM ~4 iterations before its load and the lines in flight over
                    // the whole chip stay far below the L2's capacity
                    #pragma unroll
                    for (uint32_t u = 0; u < KF_PF_DIST; u += 8)
                        if (u + l7 < len) asm volatile("prefetch.global.L2 [%0];" :: "l"(tile_at(q_tile, pb[u + l7] & 0x7FFFFFFFu)));
                }
                for (; k < steps_min; k += 4) {
                    uint32_t e[4];
                    uint4 v[4];
                    if (PF) {
                        const uint32_t kk = k + KF_PF_DIST + l7;
                        if (l7 < 4 && kk < len) asm volatile("prefetch.global.L2 [%0];" :: "l"(tile_at(q_tile, pb[kk] & 0x7FFFFFFFu)));
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        e[j] = pb[k + j];
                        v[j] = __ldg(tile_at(tile_l, e[j] & 0x7FFFFFFFu));
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc.x |= v[j].x; acc.y |= v[j].y; acc.z |= v[j].z; acc.w |= v[j].w;
                        if (e[j] >> 31) { n_cov += popc128(acc); acc = make_uint4(0, 0, 0, 0); }
                    }
                }
                for (; k < steps; k += 4) {
                    uint32_t e[4];
                    uint4 v[4];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        e[j] = (k + j < len) ? pb[k + j] : 0u;
                        v[j] = make_uint4(0, 0, 0, 0);
                        GCP_DASSERT(!(k + j < len) || (e[j] & 0x7FFFFFFFu) + l7 < 0x7FFFFFFFu);
                        if (k + j < len) v[j] = __ldg(tile_at(tile_l, e[j] & 0x7FFFFFFFu));
                    }
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc.x |= v[j].x; acc.y |= v[j].y; acc.z |= v[j].z; acc.w |= v[j].w;
                        if (e[j] >> 31) { n_cov += popc128(acc); acc = make_uint4(0, 0, 0, 0); }
                    }
                }
            } else
            for (;;) {
                uint32_t g = 0;
                if (lane == 0) g = atomicAdd(&s_group, 1u);
                g = __shfl_sync(FULL, g, 0);
                if (g * 32 >= (uint32_t)T) break;
                const uint32_t used = __ballot_sync(FULL, s_keys[g * 32 + lane] != 0u);
                for (uint32_t m = used; m; m &= m - 1) {
                    const uint32_t slot = g * 32 + (__ffs(m) - 1);
                    const uint32_t cb = s_cq[4 * slot + myq];            // my quarter's list
                    const uint32_t cnt = cb >> 16;
                    const uint32_t* pb = s_pay + (cb & 0xFFFFu);
                    const uint32_t maxc = __reduce_max_sync(FULL, cnt);
                    uint4 acc = make_uint4(0, 0, 0, 0);
                    for (uint32_t k = 0; k < maxc; k += 4) {
                        // reading past cnt only fetches a neighbour's entry (s_pay is padded)
                        const uint32_t i0 = pb[k] & 0x7FFFFFFFu, i1 = pb[k + 1] & 0x7FFFFFFFu, i2 = pb[k + 2] & 0x7FFFFFFFu,
                                       i3 = pb[k + 3] & 0x7FFFFFFFu;
                        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
                        if (k < cnt) v0 = __ldg(tile_at(tile_l, i0));
                        if (k + 1 < cnt) v1 = __ldg(tile_at(tile_l, i1));
                        if (k + 2 < cnt) v2 = __ldg(tile_at(tile_l, i2));
                        if (k + 3 < cnt) v3 = __ldg(tile_at(tile_l, i3));
                        acc.x |= (v0.x | v1.x) | (v2.x | v3.x);
                        acc.y |= (v0.y | v1.y) | (v2.y | v3.y);
                        acc.z |= (v0.z | v1.z) | (v2.z | v3.z);
                        acc.w |= (v0.w | v1.w) | (v2.w | v3.w);
                    }
                    n_cov += popc128(acc);
                    if (HAS_GRAY) {
                        // gray pixels of this block (CSR list {pos, 255 - g, in_mask}): lane i takes
                        // entry i and fetches the row word that holds its pixel by shuffle
                        const uint32_t blk = s_keys[slot] - 1u;
                        const uint32_t ga = __ldg(t.gray_ptr + blk), gb = __ldg(t.gray_ptr + blk + 1);
                        for (uint32_t g0 = ga; g0 < gb; g0 += 32) {
                            const uint32_t gi = g0 + lane;
                            const uint32_t ent = (gi < gb) ? __ldg(t.gray_ent + gi) : 0u;
                            const uint32_t pos = ent & 4095u, row = pos >> 6, col = pos & 63u;
                            const int src = (int)(((row >> 4) << 3) | ((row & 15u) >> 1));
                            const uint32_t wx = __shfl_sync(FULL, acc.x, src), wy = __shfl_sync(FULL, acc.y, src);
                            const uint32_t wz = __shfl_sync(FULL, acc.z, src), ww = __shfl_sync(FULL, acc.w, src);
                            const uint32_t word = (row & 1u) ? ((col & 32u) ? ww : wz) : ((col & 32u) ? wy : wx);
                            if (gi < gb && ((ent >> 20) & 1u) && ((word >> (col & 31u)) & 1u)) {
                                n_wg += (ent >> 12) & 255u;    // weighs 255 - g ...
                                n_cov -= 1u;                    // ... instead of a full pixel (wraps, summed mod 2^32)
                            }
                        }
                    }
                }
            }
            // binary maps, one partition: B2 only reads the sorted array and registers
            if ((HAS_GRAY && !GENERAL) || tail_barrier || R > 1) __syncthreads();
        }
        if (!overflow) break;
        if (BUCKETS && pre_m != 0xFFFFFFFFu) { failed = true; break; }
        R <<= 1;
        if (R > 4096) { failed = true; break; }
    }
    cc.n_cov = n_cov; cc.n_wg = n_wg; cc.n_free0 = n_free0; cc.wg_all = wg_all; cc.n_gmask = n_gmask; cc.n_all = n_all;
}

// Compile-time copy of fused_shape() for A*L == SAL with 256 threads and default tuning: with the
// shared-memory layout known at compile time the offsets fold into the LDS / STS / ATOMS
// immediates instead of being rebuilt from kernel parameters in front of every access.
constexpr uint32_t c_al16(uint32_t x) { return (x + 15u) & ~15u; }
// COMPACT (waypoint ids fit 16 bits, N <= 65,534): the chromosome and the compacted sequence are
// held as u16, loop-closure keys take 32 bits and the item store is sized 49/24 instead of 9/4 items
// per gene - together 37.4 KB instead of 41.7 KB for 6 x 200 genes, i.e. SIX resident CTAs per SM.
template <int SAL, bool COMPACT, bool GEN = false>
struct StaticShape {
    static constexpr int AL = SAL;
    static constexpr int TH = GEN ? 512 : 256;           // CTA size this layout is for
    static constexpr int maxp_raw = GEN ? (41 * SAL) / 8 : COMPACT ? (49 * SAL) / 24 : (9 * SAL) / 4;
    static constexpr int maxp0 = maxp_raw < 256 ? 256 : (maxp_raw > KF_PPT * TH ? KF_PPT * TH : maxp_raw);
    static constexpr int MAXP = (maxp0 + 7) & ~7;
    static constexpr int T = SAL / 3 <= TH ? TH : SAL / 3 <= 512 ? 512 : SAL / 3 <= 1024 ? 1024 : SAL / 3 <= 2048 ? 2048 : SAL / 3 <= 4096 ? 4096 : 8192;
    static constexpr int T3 = (SAL * 3) / 2 + 64;
    static constexpr uint32_t DS = (uint32_t)T * 4;        // dedup hash set: 16 KiB-aligned part of the T * 20 bytes in front of s_cst
    static constexpr uint32_t R0 = (size_t)MAXP >= (size_t)SAL * (GEN ? 5 : 2) ? 1u : 0u;      // 0: not representable -> runtime shape
    static constexpr uint32_t off_cst = (uint32_t)T * 20;
    static constexpr uint32_t cov_b = (uint32_t)T * 20 + (uint32_t)MAXP * 8 + 32;
    static constexpr uint32_t a_b = off_cst + (uint32_t)SAL * 16 + GCP_MAX_AGENTS * 16;
    // loop closures: keys (4 / 8 B) + counters (4 B) per slot, gene per position (COMPACT: 2 B + 2 B of the owner masks)
    static constexpr uint32_t lc_b = c_al16((uint32_t)T3 * (COMPACT ? 8 : 12)) + (uint32_t)SAL * 4;
    // the loop-closure tables fit in front of the sorted quarter array: phase C starts while other
    // warps are still in B2
    static constexpr bool c_overlap = COMPACT && lc_b <= (uint32_t)T * 20 + (uint32_t)MAXP * 4;
    static constexpr uint32_t x_b = cov_b > lc_b ? (cov_b > a_b ? cov_b : a_b) : (lc_b > a_b ? lc_b : a_b);
    static constexpr uint32_t off_info = c_al16((uint32_t)SAL * (COMPACT ? 2 : 4));
    static constexpr uint32_t off_x = off_info + c_al16((uint32_t)SAL * 4);
    static constexpr uint32_t off_lc = off_x + c_al16(x_b);
};

// KF_THREADS = 256 (5 or 6 CTAs / SM) for reference-sized genomes, 1024 (one CTA / SM) for long ones.
// HAS_GRAY: the map has pixels with 0 < g < 255: covered gray mask pixels weigh 255 - g.
// SAL > 0: A*L == SAL and the shared-memory layout is StaticShape<SAL> (must equal fused_shape()).
// COMPACT: N <= 65,534 waypoints - see StaticShape.
// GENERAL: any coverage_blend and / or the six pixel counters (cov_out) - the union runs over the
// UNMASKED footprint store with marker entries for the map planes (see coverage_quarters); 512 threads,
// three CTAs per SM (the unmasked store carries 2.6 x the items of the pre-masked one).
template <int KF_THREADS, bool HAS_GRAY, int SAL, bool COMPACT, bool GENERAL = false, bool PF = false>
__global__ void __launch_bounds__(KF_THREADS, KF_THREADS == 256 ? ((COMPACT && !HAS_GRAY) ? 6 : 5) : KF_THREADS == 512 ? KF_GEN_MINB : 1)
k_eval_fused(const void* __restrict__ dna_v, uint32_t P, GcpTables t, FusedShape sh,
             gcp_params prm, gcp_map_consts mc,
             double* __restrict__ obj, double* __restrict__ neg_cov_out,
             double* __restrict__ dist_out, double* __restrict__ turn_out,
             int32_t* __restrict__ lc_out, uint8_t* __restrict__ flags, uint32_t* __restrict__ cov_out = nullptr)
{
    constexpr int KF_NW = KF_THREADS / 32;
    constexpr bool KEY32 = COMPACT;
    using G = typename std::conditional<COMPACT, uint16_t, int32_t>::type;       // gene in shared memory
    extern __shared__ __align__(16) unsigned char smem[];
    G*        s_dna  = reinterpret_cast<G*>(smem);                            // [AL] (COMPACT: 0xFFFF = -1, 0xFFFE = out of range)
    constexpr bool ST = SAL > 0;
    using SS = StaticShape<ST ? SAL : 1200, COMPACT, GENERAL>;
    uint32_t* s_info = reinterpret_cast<uint32_t*>(smem + (ST ? SS::off_info : sh.off_info)); // [AL] quarter-item range per step (C: slot/next)
    unsigned char* s_x = smem + (ST ? SS::off_x : sh.off_x);                  // B / C tables
    double2*  s_cst  = reinterpret_cast<double2*>(s_x + (ST ? SS::off_cst : sh.off_cst)); // [AL] {dist, turn} per step (A only; aliases s_pay/s_sr)
    int*      s_lc   = reinterpret_cast<int*>(smem + (ST ? SS::off_lc : sh.off_lc));       // [A*A]
    double*   s_cost = reinterpret_cast<double*>(smem + sh.off_cost);         // dist[A], turn[A]
    __shared__ int s_first_neg[GCP_MAX_AGENTS], s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    __shared__ int s_bad;
    __shared__ uint32_t s_total, s_totalw, s_ngroups;
    __shared__ uint32_t s_gen[4];          // GENERAL: n_free0, wg_all, n_gmask, n_all
    __shared__ uint32_t s_cov[4];
    __shared__ double s_objv[2];

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = sh.A, L = sh.L, AL = ST ? SAL : sh.AL;
    // gene i of the chromosome as the reference's integer (-1 = padding)
    auto gene = [&](int i) -> int {
        if (COMPACT) { const int v = (int)s_dna[i]; return v == 0xFFFF ? -1 : v; }
        return (int)s_dna[i];
    };

    // ------------------------------------------------------------------ load the chromosome
    // The first vectors are requested before the table clears so that their HBM latency runs
    // while the CTA zeroes its tables; they are consumed afterwards.
    const uint32_t DS = ST ? SS::DS : sh.DS;
    uint32_t* s_set = reinterpret_cast<uint32_t*>(s_x);       // step-dedup table (phase A1), in front of s_cst
    auto clear_tables = [&]() {
        if (tid == 0) { s_bad = 0; s_total = 0; s_totalw = 0; }
        if (GENERAL && tid < 4) s_gen[tid] = 0;
        for (int i = tid; i < A * A; i += KF_THREADS) s_lc[i] = 0;
        if ((DS & 3u) == 0) {
            for (int i = tid; i < (int)(DS >> 2); i += KF_THREADS) reinterpret_cast<uint4*>(s_set)[i] = make_uint4(0, 0, 0, 0);
        } else {
            for (int i = tid; i < (int)DS; i += KF_THREADS) s_set[i] = 0u;
        }
    };
    int bad_gene = 0;                      // a gene outside [-1, N): error bit 0 of flags[:, 3]
    if (COMPACT) {
        // u16 in shared memory: -1 -> 0xFFFF, anything else outside [0, N) -> 0xFFFE (>= N: never a
        // waypoint, never a step; the organism's error flag is raised)
        auto pack = [&](int g) -> uint32_t {
            const uint32_t u = (uint32_t)g;
            const bool in = u < t.N, pad = g == -1;
            bad_gene |= (in || pad) ? 0 : 2;
            return in ? u : (pad ? 0xFFFFu : 0xFFFEu);
        };
        uint16_t* s16 = reinterpret_cast<uint16_t*>(smem);
        if (sh.gene16) {                   // int16 chromosome in global memory
            const int16_t* d = reinterpret_cast<const int16_t*>(dna_v) + (size_t)org * AL;
            if ((AL & 7) == 0) {
                const int4* d8 = reinterpret_cast<const int4*>(d);
                auto put = [&](int i, const int4& v) {
                    uint4 o;
                    o.x = pack((int)(short)(v.x & 0xFFFF)) | (pack(v.x >> 16) << 16);
                    o.y = pack((int)(short)(v.y & 0xFFFF)) | (pack(v.y >> 16) << 16);
                    o.z = pack((int)(short)(v.z & 0xFFFF)) | (pack(v.z >> 16) << 16);
                    o.w = pack((int)(short)(v.w & 0xFFFF)) | (pack(v.w >> 16) << 16);
                    reinterpret_cast<uint4*>(s16)[i] = o;
                };
                const int nv = AL >> 3;
                int4 v0 = make_int4(0, 0, 0, 0);
                if (tid < nv) v0 = __ldg(d8 + tid);
                clear_tables();
                if (tid < nv) put(tid, v0);
                for (int i = tid + KF_THREADS; i < nv; i += KF_THREADS) put(i, __ldg(d8 + i));
            } else {
                clear_tables();
                for (int i = tid; i < AL; i += KF_THREADS) s16[i] = (uint16_t)pack((int)__ldg(d + i));
            }
        } else {
            const int32_t* d = reinterpret_cast<const int32_t*>(dna_v) + (size_t)org * AL;
            if ((AL & 3) == 0) {
                const int4* d4 = reinterpret_cast<const int4*>(d);
                auto put = [&](int i, const int4& v) {
                    reinterpret_cast<uint2*>(s16)[i] = make_uint2(pack(v.x) | (pack(v.y) << 16), pack(v.z) | (pack(v.w) << 16));
                };
                const int nv = AL >> 2;
                int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
                if (tid < nv) v0 = __ldg(d4 + tid);
                if (tid + KF_THREADS < nv) v1 = __ldg(d4 + tid + KF_THREADS);
                clear_tables();
                if (tid < nv) put(tid, v0);
                if (tid + KF_THREADS < nv) put(tid + KF_THREADS, v1);
                for (int i = tid + 2 * KF_THREADS; i < nv; i += KF_THREADS) put(i, __ldg(d4 + i));
            } else {
                clear_tables();
                for (int i = tid; i < AL; i += KF_THREADS) s16[i] = (uint16_t)pack(__ldg(d + i));
            }
        }
    } else {
        int32_t* s32 = reinterpret_cast<int32_t*>(smem);
        auto check = [&](int g) { if (g < -1 || (g >= 0 && (uint32_t)g >= t.N)) bad_gene = 2; };
        if (sh.gene16) {
            const int16_t* d = reinterpret_cast<const int16_t*>(dna_v) + (size_t)org * AL;
            clear_tables();
            for (int i = tid; i < AL; i += KF_THREADS) { const int g = (int)__ldg(d + i); check(g); s32[i] = g; }
        } else {
            const int32_t* d = reinterpret_cast<const int32_t*>(dna_v) + (size_t)org * AL;
            if ((AL & 3) == 0) {
                const int4* d4 = reinterpret_cast<const int4*>(d);
                int4* s4 = reinterpret_cast<int4*>(s32);
                const int nv = AL >> 2;
                int4 v0 = make_int4(-1, -1, -1, -1), v1 = v0;
                if (tid < nv) v0 = __ldg(d4 + tid);
                if (tid + KF_THREADS < nv) v1 = __ldg(d4 + tid + KF_THREADS);
                clear_tables();
                check(v0.x); check(v0.y); check(v0.z); check(v0.w);
                check(v1.x); check(v1.y); check(v1.z); check(v1.w);
                if (tid < nv) s4[tid] = v0;
                if (tid + KF_THREADS < nv) s4[tid + KF_THREADS] = v1;
                for (int i = tid + 2 * KF_THREADS; i < nv; i += KF_THREADS) {
                    const int4 v = __ldg(d4 + i);
                    check(v.x); check(v.y); check(v.z); check(v.w);
                    s4[i] = v;
                }
            } else {
                clear_tables();
                for (int i = tid; i < AL; i += KF_THREADS) { const int g = __ldg(d + i); check(g); s32[i] = g; }
            }
        }
    }
    __syncthreads();
    if (bad_gene) atomicOr(&s_bad, bad_gene);
    // first -1 per row (step rule mappy.py:343-345) and number of valid genes (:390)
    for (int a = warp; a < A; a += KF_NW) {
        int first = L, cnt = 0;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? gene(a * L + i) : 0;
            const uint32_t neg = __ballot_sync(FULL, (i < L) && g < 0);
            if (neg && first == L) first = i0 + __ffs(neg) - 1;
            cnt += __popc(__ballot_sync(FULL, (i < L) && g != -1));
        }
        if (lane == 0) { s_first_neg[a] = first; s_len[a] = cnt; }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase A1: edge lookup
    if (t.adjc16) {
        // U steps per thread at a time: all adjacency loads first, then all cost / item-range loads
        // (two dependent L2 round trips per batch instead of two per step)
        constexpr int U = KF_A1_U;
        int bad = 0;
        for (int s0 = tid; s0 < AL; s0 += U * KF_THREADS) {
            int a_[U];
            uint32_t e_[U], w_[U], nx_[U];
            uint4 q_[U];
            bool act_[U], sent_[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s = s0 + u * KF_THREADS;
                act_[u] = false; sent_[u] = false; a_[u] = 0; e_[u] = GCP_INVALID_EDGE; w_[u] = 0; nx_[u] = 0;
                q_[u] = make_uint4(0, 0, 0, 0);
                if (s < AL) {
                    const int a = (int)__umulhi((uint32_t)s, sh.inv_L);
                    const int i = s - a * L;
                    const int g = gene(s);
                    a_[u] = a;
                    if (i < min(s_first_neg[a], L - 1)) {
                        int nx = gene(s + 1);
                        sent_[u] = nx < 0;
                        if (nx < 0) nx += (int)t.N;              // numpy wrap: -1 -> N-1
                        if ((uint32_t)g < t.N && nx >= 0 && (uint32_t)nx < t.N) {
                            act_[u] = true; w_[u] = (uint32_t)g; nx_[u] = (uint32_t)nx;
                            e_[u] = __ldg(t.row_ptr + g);
                            q_[u] = __ldg(t.adjc16 + g);         // 8 x u16 columns in one 16-B load
                        }
                    }
                }
            }
            uint32_t info_[U];
            double2 cst_[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                info_[u] = 0; cst_[u] = make_double2(0.0, 0.0);
                if (act_[u]) {
                    const uint4 q = q_[u];
                    const uint32_t pat = nx_[u] * 0x00010001u;
                    const uint32_t m0 = __vcmpeq2(q.x, pat), m1 = __vcmpeq2(q.y, pat);
                    const uint32_t m2 = __vcmpeq2(q.z, pat), m3 = __vcmpeq2(q.w, pat);
                    // one bit per 16-bit lane, in slot order
                    const uint32_t hit = ((m0 & 1u) | ((m0 >> 15) & 2u)) | (((m1 & 1u) | ((m1 >> 15) & 2u)) << 2) |
                                         (((m2 & 1u) | ((m2 >> 15) & 2u)) << 4) | (((m3 & 1u) | ((m3 >> 15) & 2u)) << 6);
                    if (hit) e_[u] += (uint32_t)(__ffs(hit) - 1);
                    else { e_[u] = t.E + w_[u]; if (!sent_[u]) bad |= 1; }   // null edge (F4)
                    info_[u] = __ldg((GENERAL ? t.q_info : t.m_q_info) + e_[u]);
                    cst_[u] = __ldg(t.edge_cost + e_[u]);
                }
            }
            #pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s = s0 + u * KF_THREADS;
                if (s < AL) {
                    uint32_t info = info_[u];
                    if (DS && info) {                            // a revisited edge adds nothing to the union
                        // one exchange, no probing: a step that finds its own item list in the slot is a
                        // revisit of the step that put it there.  Dropping is an optimisation only (the
                        // union is idempotent), so a slot shared with another edge may keep a duplicate;
                        // the first arrival of every list always survives.
                        const uint32_t h = (info * 0x9E3779B1u) >> (__clz(DS) + 1);
                        if (atomicExch(&s_set[h], info) == info) info = 0u;
                    }
                    GCP_DASSERT(a_[u] < A && s + a_[u] < AL + GCP_MAX_AGENTS && (e_[u] == GCP_INVALID_EDGE || e_[u] < t.E + t.N));
                    s_cst[s + a_[u]] = cst_[u];                   // row stride L + 1: see phase A2
                    s_info[s] = info;
                }
            }
        }
        if (bad) atomicOr(&s_bad, bad);
    } else {
        int bad = 0;
        for (int s = tid; s < AL; s += KF_THREADS) {
            const int a = (int)__umulhi((uint32_t)s, sh.inv_L);
            const int i = s - a * L;
            const int g = gene(s);
            uint32_t e = GCP_INVALID_EDGE, info = 0;
            double2 cst = make_double2(0.0, 0.0);
            const int nsteps = min(s_first_neg[a], L - 1);
            if (i < nsteps) {
                const uint32_t w = (uint32_t)g;
                int nx = gene(s + 1);
                const bool sentinel = nx < 0;
                if (nx < 0) nx += (int)t.N;                  // numpy wrap: -1 -> N-1
                if (w < t.N && nx >= 0 && (uint32_t)nx < t.N) {
                    const uint32_t lo = __ldg(t.row_ptr + w);
                    e = t.E + w;                             // null edge unless found (F4)
                    bool found = false;
                    if (t.adj8) {
                        const uint4* rec = reinterpret_cast<const uint4*>(t.adj8 + (size_t)w * 8);
                        #pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const uint4 q = __ldg(rec + h);
                            if (q.x == (uint32_t)nx && !found) { e = lo + 2 * h; found = true; }
                            if (q.z == (uint32_t)nx && !found) { e = lo + 2 * h + 1; found = true; }
                        }
                    } else {
                        const uint32_t hi = __ldg(t.row_ptr + w + 1);
                        for (uint32_t q = lo; q < hi; ++q)
                            if (__ldg(t.col + q) == (uint32_t)nx) { e = q; found = true; break; }
                    }
                    if (!found && !sentinel) bad |= 1;
                    info = __ldg((GENERAL ? t.q_info : t.m_q_info) + e);
                    cst = __ldg(t.edge_cost + e);
                    if (DS && info) {                        // revisited edge: see above
                        const uint32_t h = (info * 0x9E3779B1u) >> (__clz(DS) + 1);
                        if (atomicExch(&s_set[h], info) == info) info = 0u;
                    }
                }
            }
            GCP_DASSERT(a < A && s + a < AL + GCP_MAX_AGENTS && (e == GCP_INVALID_EDGE || e < t.E + t.N));
            s_cst[s + a] = cst;                                   // row stride L + 1: see phase A2
            s_info[s] = info;
        }
        if (bad) atomicOr(&s_bad, bad);
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase B: wall coverage
    // Items are QUARTERS (16 rows x 64 px = one 128-B line) of the visited edges' pre-masked
    // sub-tiles.  B1 counting-sorts them by (map block, quarter index): an open-addressing hash
    // table maps block id -> slot, one shared atomicAdd per item returns its rank inside
    // (slot, quarter).  B2: the sorted array is cut into equal ranges snapped to list starts; an
    // 8-lane group streams its range (lane = one uint4 = 2 rows of the quarter), ORs into registers
    // and popcounts at every list end (gray maps: a warp per block, one 8-lane group per quarter).
    // No lane idles on absent quarters, no bitmap in shared memory, no bitmap atomics.
    // phase A2 (lanes 0..A-1 of warp 0) runs while the other warps start clearing B's tables
    if (warp == 0 && lane < A) {
        // phase A2: sequential fp64 adds in step order per agent (mappy.py:351-352),
        // bit-identical to the reference's `+=`; the per-step costs were staged by A1
        // rows are L + 1 entries apart: with a stride of L * 16 B (a multiple of 128 B for L = 200)
        // the A lanes would all hit the same banks
        const double2* ca = s_cst + lane * (L + 1);
        const int nsteps = min(s_first_neg[lane], L - 1);
        double acc_d = 0.0, acc_t = 0.0;
        int i = 0;
        for (; i + 8 <= nsteps; i += 8) {
            double2 c[8];
            #pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = ca[i + k];
            #pragma unroll
            for (int k = 0; k < 8; ++k) { acc_d = __dadd_rn(acc_d, c[k].x); acc_t = __dadd_rn(acc_t, c[k].y); }
        }
        for (; i < nsteps; ++i) { acc_d = __dadd_rn(acc_d, ca[i].x); acc_t = __dadd_rn(acc_t, ca[i].y); }
        s_cost[lane] = acc_d;
        s_cost[A + lane] = acc_t;
        if (dist_out) dist_out[(size_t)org * A + lane] = acc_d;
        if (turn_out) turn_out[(size_t)org * A + lane] = acc_t;
    }
    CovCounts cc;
    bool failed;
    // 32-bit loop-closure keys and a flat B2 (binary map, or the general kernel): phase C's tables lie in
    // front of the sorted quarter array, so no barrier separates B2 from C (warps that finish their B2
    // range early go on)
    const bool c_overlap = COMPACT && (!HAS_GRAY || GENERAL) && (ST ? SS::c_overlap : (sh.c_overlap != 0));
    coverage_quarters<KF_THREADS, HAS_GRAY, false, GENERAL, PF>(s_info, AL, t, ST ? SS::T : sh.T, ST ? SS::MAXP : sh.MAXP,
                                            ST ? 1u : sh.R0, 0u, s_x, cc, failed, nullptr, 0,
                                            sh.prefetch, !c_overlap);
    {
        uint32_t v = cc.n_cov, vw = cc.n_wg;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(FULL, v, o); vw += __shfl_xor_sync(FULL, vw, o); }
        if (lane == 0 && v) atomicAdd(&s_total, v);
        if (HAS_GRAY && lane == 0 && vw) atomicAdd(&s_totalw, vw);
        if (GENERAL) {
            uint32_t g4[4] = {cc.n_free0, cc.wg_all, cc.n_gmask, cc.n_all};
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!HAS_GRAY && (k == 1 || k == 2)) continue;
                uint32_t x = g4[k];
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
                if (lane == 0 && x) atomicAdd(&s_gen[k], x);
            }
        }
    }
    if (!c_overlap) __syncthreads();      // B's tables are dead from here

    // ------------------------------------------------------------------ phase C: loop closures
    // Consecutive gene pairs of the compacted sequence are hashed (key -> slot, members counted); the
    // second arrival of a key opens a GROUP (duplicates are the minority: ~80 groups for ~900 pairs),
    // whose record {slot << 16 | list head, owner mask} lives where the chromosome was (dead after
    // the compaction).  Members chain themselves with one exchange, test their successor gap, and
    // one thread per group updates lc_mat - no pass over the hash slots after the clear.
    //   s_hcnt[h]: members (15 bits) | touches a row split << 15 (F10) | solo gap found << 16 | group id + 1 << 17
    using HK = typename std::conditional<KEY32, uint32_t, unsigned long long>::type;
    const int T3 = ST ? SS::T3 : sh.T3;
    HK*       s_hkey = reinterpret_cast<HK*>(s_x);                                 // [T3]
    uint32_t* s_hcnt = reinterpret_cast<uint32_t*>(s_hkey + T3);                   // [T3]
    G*        s_seq  = reinterpret_cast<G*>(s_x + (((size_t)T3 * (sizeof(HK) + 4) + 15) & ~(size_t)15)); // [AL]
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_info);                        // [AL]
    uint16_t* s_next = s_slot + AL;                                                // [AL]
    uint32_t* g_head = reinterpret_cast<uint32_t*>(s_dna);                         // [AL/2] slot << 16 | newest member
    // [AL/2] owner mask: behind the heads where the chromosome took 4 bytes per gene, else behind s_seq
    uint32_t* g_own  = COMPACT ? reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_seq) + (((size_t)AL * 2 + 3) & ~(size_t)3))
                               : g_head + (AL >> 1);
    for (int i = tid; i < T3; i += KF_THREADS) { s_hkey[i] = (HK)0; s_hcnt[i] = 0u; }
    if (tid == 0) {
        int b = 0;
        for (int a = 0; a < A; ++a) { s_base[a] = b; b += s_len[a]; }
        s_base[A] = b;
        s_ngroups = 0;
    }
    __syncthreads();
    const int n = s_base[A];
    // compaction: seq[pos] = gene (the owner of a position follows from s_base when it is needed)
    for (int a = warp; a < A; a += KF_NW) {
        int run = s_base[a];
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? gene(a * L + i) : -1;
            const bool v = g != -1;
            const uint32_t bal = __ballot_sync(FULL, v);
            if (v) {
                const int pos = run + __popc(bal & ((1u << lane) - 1u));
                GCP_DASSERT(pos >= 0 && pos < AL);
                s_seq[pos] = (G)g;
            }
            run += __popc(bal);
        }
    }
    __syncthreads();                       // the chromosome is dead from here: group records take its place
    for (int i = tid; i < (AL >> 1); i += KF_THREADS) { g_head[i] = KF_NIL; g_own[i] = 0u; }
    auto lc_key = [&](int i, HK& key, uint32_t& h) {
        const uint32_t f = (uint32_t)s_seq[i], g = (uint32_t)s_seq[(i + 1 == n) ? 0 : i + 1];
        if (KEY32) {
            key = (HK)((((f & 0xFFFFu) << 16) | (g & 0xFFFFu)) + 1u);      // ids < N <= 65,535: never 0
            h = __umulhi((uint32_t)key * 0x9E3779B1u, (uint32_t)T3);
        } else {
            key = (HK)(((unsigned long long)(f & 0xFFFFFFu) << 24 | (g & 0xFFFFFFu)) + 1ull);
            h = __umulhi((uint32_t)(((unsigned long long)key * 0x9E3779B97F4A7C15ull) >> 32), (uint32_t)T3);
        }
    };
    // pass 1: hash every key, count the members, open a group at the second arrival
    for (int i = tid; i < n; i += KF_THREADS) {
        HK key;
        uint32_t h;
        lc_key(i, key, h);
        for (;;) {
            const HK cur = ((volatile HK*)s_hkey)[h];
            if (cur == key) break;
            if (cur == (HK)0) {
                const HK old = atomicCAS(&s_hkey[h], (HK)0, key);
                if (old == (HK)0 || old == key) break;
            }
            h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;    // T3 > n: always terminates
        }
        GCP_DASSERT(h < (uint32_t)T3 && i < AL);
        if ((atomicAdd(&s_hcnt[h], 1u) & 0x7FFFu) == 1u) {
            const uint32_t gid = atomicAdd(&s_ngroups, 1u);
            GCP_DASSERT(gid < (uint32_t)(AL >> 1));
            atomicAdd(&s_hcnt[h], (gid + 1u) << 17);
        }
        s_slot[i] = (uint16_t)h;
    }
    __syncthreads();
    // pass 2: members of duplicate groups chain themselves and leave their owner bit
    for (int i = tid; i < n; i += KF_THREADS) {
        const uint32_t h = s_slot[i];
        const uint32_t c = s_hcnt[h];
        if ((c & 0x7FFFu) < 2u) continue;
        const uint32_t gid = (c >> 17) - 1u;
        const int a = owner_row(s_base, A, i);
        GCP_DASSERT(gid < (uint32_t)(AL >> 1));
        s_next[i] = (uint16_t)atomicExch(&g_head[gid], (h << 16) | (uint32_t)i);       // NIL -> 0xFFFF
        atomicOr(&g_own[gid], 1u << a);
        if (a >= 1 && i == s_base[a]) atomicOr(&s_hcnt[h], 1u << 15);       // first gene of a row: group touches a split (F10)
    }
    __syncthreads();
    // phase D, first half: -coverage and travel cost need nothing from phase C - one thread of the last
    // warp (no groups to update in pass 4) computes them here instead of at the tail of the CTA
    if (tid == KF_THREADS - 32) {
        s_cov[0] = failed ? 0u : s_total; s_cov[1] = (GENERAL && !failed) ? s_gen[0] : 0u;
        s_cov[2] = failed ? 0u : s_totalw; s_cov[3] = (GENERAL && !failed) ? s_gen[1] : 0u;
        double nc, tr;
        objective_values(A, prm, mc, s_cov, s_cost, s_cost + A, nc, tr);
        s_objv[0] = nc; s_objv[1] = tr;
    }
    // pass 3: successor gap test (ascending consecutive positions)
    const int thresh = prm.solo_sep_thresh;
    for (int i = tid; i < n; i += KF_THREADS) {
        const uint32_t h = s_slot[i];
        const uint32_t c = s_hcnt[h];
        if ((c & 0x7FFFu) < 2u || (c & (1u << 15))) continue;
        uint32_t succ = 0xFFFFFFFFu;
        for (uint32_t j = g_head[(c >> 17) - 1u] & 0xFFFFu; j != 0xFFFFu; j = s_next[j])
            if (j > (uint32_t)i && j < succ) succ = j;
        if (succ != 0xFFFFFFFFu && (int)(succ - (uint32_t)i) > thresh)
            atomicOr(&s_hcnt[h], 1u << 16);
    }
    __syncthreads();
    // pass 4: one thread per duplicate group updates lc_mat
    for (int gid = tid; gid < (int)s_ngroups; gid += KF_THREADS) {
        const uint32_t c = s_hcnt[g_head[gid] >> 16];
        if (c & (1u << 15)) continue;
        const uint32_t owners = g_own[gid];
        if (__popc(owners) == 1) {
            if (c & (1u << 16)) { const int a = __ffs(owners) - 1; atomicAdd(&s_lc[a * A + a], 1); }
        } else {
            GCP_DASSERT(owners < (A >= 32 ? 0xFFFFFFFFu : (1u << A)) || A >= 32);
            for (uint32_t mx = owners; mx; mx &= mx - 1) {
                const int x = __ffs(mx) - 1;
                for (uint32_t my = mx & (mx - 1); my; my &= my - 1)
                    atomicAdd(&s_lc[x * A + (__ffs(my) - 1)], 1);
            }
        }
    }
    __syncthreads();
    if (lc_out)
        for (int i = tid; i < A * A; i += KF_THREADS) lc_out[(size_t)org * A * A + i] = s_lc[i];

    // the six pixel counters of the general path, as k_coverage writes them (gcp_eval.cu)
    if (GENERAL && cov_out && tid < 8) {
        uint32_t v = 0;
        if (!failed) {
            if (tid == 0) v = s_total;
            else if (tid == 1) v = s_gen[0];
            else if (tid == 2) v = s_totalw;
            else if (tid == 3) v = s_gen[1];
            else if (tid == 4) v = s_total + s_gen[2];
            else if (tid == 5) v = s_gen[3];
        }
        cov_out[(size_t)org * 8 + tid] = v;
    }
    // ------------------------------------------------------------------ phase D: constraints (warp 0)
    if (warp == 0) {
        int ncomp;
        const bool feasible = constraints_warp(s_lc, A, prm, ncomp);
        if (lane == 0) {
            const double neg_cov = s_objv[0];
            if (neg_cov_out) neg_cov_out[org] = neg_cov;
            if (obj) {
                obj[(size_t)org * 2 + 0] = feasible ? neg_cov : 0.0;
                obj[(size_t)org * 2 + 1] = s_objv[1];
            }
            if (flags) {
                flags[(size_t)org * 4 + 0] = feasible ? 1 : 0;
                flags[(size_t)org * 4 + 1] = (s_bad & 1) ? 0 : 1;                       // traversable
                flags[(size_t)org * 4 + 2] = (uint8_t)ncomp;
                flags[(size_t)org * 4 + 3] = (uint8_t)(((s_bad & 2) ? 1 : 0) | (failed ? 2 : 0));
            }
        }
    }
}

// ============================================================================
// Stand-alone wall coverage for genomes that outgrow the fused kernel's shared memory (staged
// path: k_path_costs -> k_cov_quarters -> loop closures).  step_qinfo: dev u32 [P][A*L], per
// step the quarter-item range of its edge (GcpTables::m_q_info), 0 where no step is executed.
// cov[P][8] = {n_mask0, 0, wg_mask, 0, n_mask0, 0, 0, 0}.
// ============================================================================
template <int KF_THREADS, bool HAS_GRAY>
__global__ void __launch_bounds__(KF_THREADS, KF_THREADS == 256 ? 5 : KF_THREADS == 512 ? 2 : 1)
k_cov_quarters(uint32_t* step_qinfo, uint32_t P, GcpTables t, int AL, int T, int MAXP,
               uint32_t R0, uint32_t DS, uint2* rec_all, uint32_t rec_cap,
               uint32_t* __restrict__ cov, uint8_t* __restrict__ flags,
               const uint8_t* __restrict__ fix = nullptr, const uint32_t* __restrict__ acc = nullptr)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_total, s_totalw;
    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31;
    // split path (k_cov_bucket -> k_cov_part): the partition CTAs have summed this organism's counts
    // unless one of them overflowed (fix[org]) - then it is redone here, in one CTA, exactly
    if (fix && !fix[org]) {
        if (tid < 8) cov[(size_t)org * 8 + tid] = (tid == 0 || tid == 4) ? acc[(size_t)org * 2] : (tid == 2 ? acc[(size_t)org * 2 + 1] : 0u);
        return;
    }
    if (tid == 0) { s_total = 0; s_totalw = 0; }
    CovCounts cc;
    bool failed;
    coverage_quarters<KF_THREADS, HAS_GRAY, true>(step_qinfo + (size_t)org * AL, AL, t, T, MAXP, R0, DS, smem,
                                                  cc, failed,
                                                  rec_all ? rec_all + (size_t)org * rec_cap : nullptr, rec_cap);
    uint32_t v = cc.n_cov, vw = cc.n_wg;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(FULL, v, o); vw += __shfl_xor_sync(FULL, vw, o); }
    if (lane == 0 && v) atomicAdd(&s_total, v);
    if (HAS_GRAY && lane == 0 && vw) atomicAdd(&s_totalw, vw);
    __syncthreads();
    if (tid < 8) {
        const uint32_t n = failed ? 0u : s_total, w = failed ? 0u : s_totalw;
        cov[(size_t)org * 8 + tid] = (tid == 0 || tid == 4) ? n : (tid == 2 ? w : 0u);
    }
    if (tid == 0 && failed && flags) flags[(size_t)org * 4 + 3] |= 2;
}


// ============================================================================
// Long genomes, split form.  The monolithic k_cov_quarters keeps one 1,024-thread CTA per organism (one
// per SM) busy with R sequential hash partitions, each a chain of short barrier-separated phases: issue
// slots ~50 % busy.  Here the partitions of ALL organisms become independent work items for small CTAs
// (256 threads, six per SM - the regime of the fused kernel's phase B):
//   k_cov_bucket  one CTA per organism: revisited steps dropped, items counted, K = 2^k partitions chosen
//                 for ~cap items each, item descriptors scattered by partition into the organism's stretch of
//                 the scratch array (partition r owns the r-th K-th of it); {K, items per partition} to hdr, K work
//                 items {organism, partition} to the list
//   k_cov_part    persistent CTAs take work items off the list: one coverage_quarters pass over the
//                 bucket (counting sort by (block, quarter), flat union, popcount), counts added to acc[org]
//   k_cov_quarters(fix, acc)  writes cov[] from acc; an organism whose bucket overflowed is redone monolithically
// ============================================================================
constexpr int CB_THREADS = 1024;
constexpr uint32_t CB_MAXK = 64;
constexpr uint32_t CB_HDR = CB_MAXK + 2;         // words per organism: K, items of partition 0 .. K-1

__global__ void __launch_bounds__(CB_THREADS, 1)
k_cov_bucket(uint32_t* step_qinfo, uint32_t P, GcpTables t, int AL, uint32_t DS, uint32_t cap, uint32_t K_fixed,
             uint2* __restrict__ rec_all, uint32_t rec_cap, uint32_t* __restrict__ hdr_all,
             uint32_t* __restrict__ work /* [0] = items listed, [1] = ticket, [4 + 4 i ..] = item i: {organism, first record, records, 0} */,
             uint8_t* __restrict__ fix)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NW = CB_THREADS / 32;
    __shared__ uint32_t s_pcnt[CB_MAXK], s_wsum[NW], s_K, s_wbase;
    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* info = step_qinfo + (size_t)org * AL;
    uint2* rec = rec_all + (size_t)org * rec_cap;
    uint32_t* hdr = hdr_all + (size_t)org * CB_HDR;
    // revisited edges add nothing to the union (see coverage_quarters B0).  K_fixed != 0 (the launcher's choice
    // from the genome length): the exchange filter runs inside the ONE scatter pass below.  K_fixed == 0: a
    // first pass filters and counts the surviving items, K is chosen per organism for <= cap items each.
    uint32_t* s_set = reinterpret_cast<uint32_t*>(smem);
    const int sh_ds = DS ? __clz(DS) + 1 : 0;
    if (tid < (int)CB_MAXK) s_pcnt[tid] = 0;
    if (DS)
        for (int i = tid; i < (int)(DS >> 2); i += CB_THREADS) reinterpret_cast<uint4*>(s_set)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const bool inline_dedup = K_fixed != 0 && DS != 0;
    uint32_t K = K_fixed;
    if (K_fixed == 0) {
        uint32_t mine = 0;
        if (DS) {
            for (int s = tid; s < AL; s += CB_THREADS) {
                const uint32_t inf = info[s];
                if (!inf) continue;
                const uint32_t h = (inf * 0x9E3779B1u) >> sh_ds;
                if (atomicExch(&s_set[h], inf) == inf) info[s] = 0u;
                else mine += inf >> 26;
            }
        } else {
            for (int s = tid; s < AL; s += CB_THREADS) mine += info[s] >> 26;
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(FULL, mine, o);
        if (lane == 0) s_wsum[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
            for (int w2 = 0; w2 < NW; ++w2) total += s_wsum[w2];
            uint32_t Kc = 1;
            while ((uint64_t)Kc * cap < total && Kc < CB_MAXK) Kc <<= 1;
            if (total > rec_cap) Kc = 0;                  // more items than the scratch holds: monolithic pass
            s_K = Kc;
        }
        __syncthreads();
        K = s_K;
        if (K == 0) {
            if (tid == 0) { hdr[0] = 0; fix[org] = 1; }
            return;
        }
    }
    auto scan_items = [&](auto&& fn) {
        for (int s0 = warp * 32; s0 < AL; s0 += NW * 32) {
            const int s = s0 + lane;
            uint32_t lo = 0, n = 0;
            if (s < AL) {
                uint32_t inf = info[s];
                if (inline_dedup && inf) {                // a revisit finds its own item list in its slot
                    const uint32_t h = (inf * 0x9E3779B1u) >> sh_ds;
                    if (atomicExch(&s_set[h], inf) == inf) inf = 0u;
                }
                lo = inf & 0x03FFFFFFu; n = inf >> 26;
            }
            uint32_t incl = n;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            const uint32_t excl = incl - n;
            for (uint32_t j0 = 0; j0 < total; j0 += 32) {
                const uint32_t j = j0 + lane;
                int a = 0, b = 31;
                #pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const int mid = (a + b) >> 1;
                    const uint32_t v = __shfl_sync(FULL, incl, mid);
                    if (v > j) b = mid; else a = mid + 1;
                }
                a = min(a, 31);
                const uint32_t ex = __shfl_sync(FULL, excl, a), dlo = __shfl_sync(FULL, lo, a);
                uint2 dsc = make_uint2(0, 0);
                if (j < total) dsc = __ldg(t.m_q_desc + dlo + (j - ex));
                fn(j < total, dsc);
            }
        }
    };
    // ONE pass: partition r owns the stretch [r * region, (r + 1) * region) of the organism's scratch, an item
    // takes the next free place of its partition (lanes of a warp that hit the same partition share one
    // shared-memory atomic).  A partition that outgrows its stretch could not be sorted by a partition CTA
    // either (region >= its item store whenever K >= 2): the organism is redone monolithically.
    const uint32_t region = rec_cap / K;
    scan_items([&](bool valid, uint2 dsc) {
        const uint32_t p = valid ? (part_block(dsc.x >> 2) & (K - 1)) : KF_NIL;
        const uint32_t peers = __match_any_sync(FULL, p);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) base = atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
        base = __shfl_sync(FULL, base, leader);
        const uint32_t at = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (valid && at < region) rec[(size_t)p * region + at] = dsc;
    });
    __syncthreads();
    if (tid == 0) {
        bool over = false;
        hdr[0] = K;
        for (uint32_t r = 0; r < K; ++r) { hdr[1 + r] = s_pcnt[r]; over |= s_pcnt[r] > region; }
        if (over) { hdr[0] = 0; fix[org] = 1; }
        else s_wbase = atomicAdd(&work[0], K);
        s_K = over ? 0u : K;
    }
    __syncthreads();
    if (s_K && tid < (int)K)
        reinterpret_cast<uint4*>(work + 4)[s_wbase + tid] = make_uint4(org, (uint32_t)tid * region, s_pcnt[tid], 0u);
}

template <bool HAS_GRAY>
__global__ void __launch_bounds__(256, 6)
k_cov_part(GcpTables t, int T, int MAXP, uint2* __restrict__ rec_all, uint32_t rec_cap,
           uint32_t* __restrict__ work, uint32_t* __restrict__ acc, uint8_t* __restrict__ fix)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_total, s_totalw;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t n_work = work[0];
    const uint4* items = reinterpret_cast<const uint4*>(work + 4);
    // items are taken off the list with a ticket: their cost varies by a factor of two with the organism
    // (K is a power of two), and dealing them out statically left the slowest of the 888 CTAs 10 % behind
    // the mean (measured: coverage 2.30 -> 2.53 ms per 4,096 organisms of C4-32)
    __shared__ uint4 s_item;
    for (;;) {
        __syncthreads();                                  // the previous item's readers are done
        if (tid == 0) {
            const uint32_t w = atomicAdd(&work[1], 1u);
            s_item = (w < n_work) ? items[w] : make_uint4(0, 0, 0xFFFFFFFFu, 0);
            s_total = 0; s_totalw = 0;
        }
        __syncthreads();
        const uint32_t org = s_item.x, b0 = s_item.y, m = s_item.z;
        if (m == 0xFFFFFFFFu) break;
        if (m == 0) continue;
        CovCounts cc;
        bool failed;
        coverage_quarters<256, HAS_GRAY, true>(nullptr, 0, t, T, MAXP, 1u, 0u, smem, cc, failed,
                                               rec_all + (size_t)org * rec_cap + b0, 0u, 0, true, m);
        uint32_t v = cc.n_cov, vw = cc.n_wg;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(FULL, v, o); vw += __shfl_xor_sync(FULL, vw, o); }
        if (lane == 0 && v) atomicAdd(&s_total, v);
        if (HAS_GRAY && lane == 0 && vw) atomicAdd(&s_totalw, vw);
        __syncthreads();
        if (tid == 0) {
            if (failed) fix[org] = 1;
            else {
                if (s_total) atomicAdd(&acc[(size_t)org * 2], s_total);
                if (HAS_GRAY && s_totalw) atomicAdd(&acc[(size_t)org * 2 + 1], s_totalw);
            }
        }
    }
}

int launch_coverage_quarters(gcp_ctx* ctx, uint32_t* step_qinfo, uint32_t P, uint32_t* cov,
                             uint8_t* flags, cudaStream_t st)
{
    if (P == 0) return 0;
    if (!ctx->have_masked || !ctx->t.m_q_info) return gcp_fail(ctx, -1, "quarter-item tables not uploaded");
    const int AL = ctx->p.num_agents * ctx->p.max_dna_len;
    int threads = AL > 2048 ? 1024 : 256;
    if (ctx->cov_threads == 256 || ctx->cov_threads == 512 || ctx->cov_threads == 1024) threads = ctx->cov_threads;
    int maxp = ctx->cov_maxpairs > 0 ? ctx->cov_maxpairs : (9 * AL) / 4;
    if (maxp < 256) maxp = 256;
    if (maxp > KF_PPT * threads) maxp = KF_PPT * threads;
    maxp = (maxp + 7) & ~7;
    uint32_t R0 = 1;
    while ((size_t)R0 * maxp < (size_t)AL * 2 && R0 < 4096) R0 <<= 1;
    int T = ctx->cov_table;
    if (T <= 0) { T = threads; while (T < (AL / 3) / (int)R0 && T < 4096) T <<= 1; }
    const size_t smem = (size_t)T * 20 + (size_t)maxp * 8 + 32;
    if (smem > (size_t)ctx->max_smem_optin)
        return gcp_fail(ctx, -3, "coverage kernel needs %zu B shared memory", smem);
    // step dedup (B0): hash set over the whole table area; skipped if it cannot stay under 80 % full
    uint32_t DS = 1;
    while ((size_t)DS * 2 * 4 <= (size_t)T * 20 + (size_t)maxp * 8) DS <<= 1;
    if ((size_t)DS * 4 < (size_t)AL * 5 || ctx->cov_dedup == 0) DS = 0;
    // long genomes (R0 > 1): scratch for the item buckets, 3 * A*L descriptors per organism, at most
    // 2 GiB (organisms go through in chunks); an organism with more items falls back to scanning passes
    uint2* rec = nullptr;
    uint32_t rec_cap = 0, chunk = P;
    // split form (k_cov_bucket -> k_cov_part -> k_cov_quarters as the writer / fallback): partitions of all
    // organisms as work items for 256-thread CTAs
    const bool split = R0 > 1 && ctx->cov_buckets && ctx->cov_split && threads == 1024;
    size_t off_hdr = 0, off_work = 0, off_acc = 0;
    if (R0 > 1 && ctx->cov_buckets) {
        rec_cap = 3u * (uint32_t)AL;
        chunk = (uint32_t)(((size_t)2 << 30) / ((size_t)rec_cap * 8));
        if (chunk < 1) chunk = 1;
        if (chunk > P) chunk = P;
        size_t need = (size_t)chunk * rec_cap * 8;
        if (split) {
            off_hdr = (need + 255) & ~(size_t)255;
            off_work = (off_hdr + (size_t)chunk * CB_HDR * 4 + 255) & ~(size_t)255;
            off_acc = (off_work + (4 + (size_t)chunk * CB_MAXK * 4) * 4 + 255) & ~(size_t)255;
            need = off_acc + (size_t)chunk * 8 + chunk;          // acc[chunk][2] u32, fix[chunk] u8
        }
        if (ctx->cov_scratch_bytes_s[ctx->scratch_slot] < need) {
            if (ctx->cov_scratch_s[ctx->scratch_slot]) cudaFree(ctx->cov_scratch_s[ctx->scratch_slot]);
            ctx->cov_scratch_s[ctx->scratch_slot] = nullptr; ctx->cov_scratch_bytes_s[ctx->scratch_slot] = 0;
            GCP_CUDA(ctx, cudaMalloc(&ctx->cov_scratch_s[ctx->scratch_slot], need));
            ctx->cov_scratch_bytes_s[ctx->scratch_slot] = need;
        }
        rec = reinterpret_cast<uint2*>(ctx->cov_scratch_s[ctx->scratch_slot]);
    }
    StageTimer tm(ctx, 1, st);
    const bool gray = ctx->n_gray > 0;
    if (split) {
        char* base = reinterpret_cast<char*>(ctx->cov_scratch_s[ctx->scratch_slot]);
        uint32_t* hdr = reinterpret_cast<uint32_t*>(base + off_hdr);
        uint32_t* work = reinterpret_cast<uint32_t*>(base + off_work);
        uint32_t* acc = reinterpret_cast<uint32_t*>(base + off_acc);
        uint8_t* fix = reinterpret_cast<uint8_t*>(base + off_acc + (size_t)chunk * 8);
        // items aimed at per partition: few large work items beat many small ones (C4-32, coverage per 4,096
        // organisms: 2.91 / 2.32 / 2.15 ms at <= 1,024 / 2,048 / 2,600 items, i.e. K = 32 / 16 / 8); 2,600 + the
        // ~10 % imbalance of the block hash stays inside the partition CTA's store of 3,072
        const uint32_t CAP = ctx->cov_split_cap > 0 ? (uint32_t)ctx->cov_split_cap : 2600u;
        // cov_split_k (tuning): K fixed for all organisms - the bucket kernel then filters and scatters in ONE
        // pass (-2 %), but an organism with more items than expected overflows into the monolithic fallback.
        // Default 0: the kernel counts the surviving items first and chooses K per organism.
        uint32_t KFIX = 0;
        if (ctx->cov_split_k > 0) { KFIX = 1; while (KFIX < CB_MAXK && KFIX < (uint32_t)ctx->cov_split_k) KFIX <<= 1; }
        constexpr int TP = 512, MAXPP = 3072;              // block slots / item store of a partition CTA
        const size_t smem_p = (size_t)TP * 20 + (size_t)MAXPP * 8 + 32;
        // dedup table of the bucket kernel: the CTA's whole shared memory is free for it
        uint32_t DSb = 1024;
        while (DSb < 32768u && (size_t)DSb * 2 < (size_t)AL * 3) DSb <<= 1;
        if (ctx->cov_dedup == 0) DSb = 0;
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DSb * 4 + 16)));
        if (gray) {
            GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_part<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_quarters<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else {
            GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_part<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_quarters<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        const int grid_p = ctx->num_sms * 6;
        for (uint32_t o0 = 0; o0 < P; o0 += chunk) {
            const uint32_t n_org = (P - o0 < chunk) ? (P - o0) : chunk;
            uint32_t* qi = step_qinfo + (size_t)o0 * AL;
            GCP_CUDA(ctx, cudaMemsetAsync(work, 0, 16, st));
            GCP_CUDA(ctx, cudaMemsetAsync(acc, 0, (size_t)chunk * 8 + chunk, st));
            k_cov_bucket<<<n_org, CB_THREADS, DSb * 4 + 16, st>>>(qi, n_org, ctx->t, AL, DSb, CAP, KFIX, rec, rec_cap, hdr, work, fix);
            if (gray) {
                k_cov_part<true><<<grid_p, 256, smem_p, st>>>(ctx->t, TP, MAXPP, rec, rec_cap, work, acc, fix);
                k_cov_quarters<1024, true><<<n_org, 1024, smem, st>>>(qi, n_org, ctx->t, AL, T, maxp, R0, DS, rec, rec_cap,
                                                                      cov + (size_t)o0 * 8, flags ? flags + (size_t)o0 * 4 : nullptr, fix, acc);
            } else {
                k_cov_part<false><<<grid_p, 256, smem_p, st>>>(ctx->t, TP, MAXPP, rec, rec_cap, work, acc, fix);
                k_cov_quarters<1024, false><<<n_org, 1024, smem, st>>>(qi, n_org, ctx->t, AL, T, maxp, R0, DS, rec, rec_cap,
                                                                       cov + (size_t)o0 * 8, flags ? flags + (size_t)o0 * 4 : nullptr, fix, acc);
            }
            ctx->launches += 3;
        }
        GCP_CUDA(ctx, cudaGetLastError());
        return 0;
    }
#define GCP_LAUNCH_COVQ(TH, GRAY)                                                                      \
    do {                                                                                               \
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_cov_quarters<TH, GRAY>,                                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        for (uint32_t o0 = 0; o0 < P; o0 += chunk) {                                                   \
            const uint32_t n_org = (P - o0 < chunk) ? (P - o0) : chunk;                                \
            k_cov_quarters<TH, GRAY><<<n_org, TH, smem, st>>>(step_qinfo + (size_t)o0 * AL, n_org, ctx->t, AL, T, \
                                                              maxp, R0, DS, rec, rec_cap,               \
                                                              cov + (size_t)o0 * 8,                    \
                                                              flags ? flags + (size_t)o0 * 4 : nullptr); \
            ctx->launches++;                                                                           \
        }                                                                                              \
    } while (0)
    if (threads == 256) { if (gray) GCP_LAUNCH_COVQ(256, true); else GCP_LAUNCH_COVQ(256, false); }
    else if (threads == 512) { if (gray) GCP_LAUNCH_COVQ(512, true); else GCP_LAUNCH_COVQ(512, false); }
    else                { if (gray) GCP_LAUNCH_COVQ(1024, true); else GCP_LAUNCH_COVQ(1024, false); }
#undef GCP_LAUNCH_COVQ
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}

static inline uint32_t al16(uint32_t x) { return (x + 15u) & ~15u; }

static size_t fused_shape(gcp_ctx* ctx, FusedShape& sh, int threads, bool general);
static int fused_threads(gcp_ctx* ctx, FusedShape& sh, size_t* smem, bool general);

// does the fused kernel's shared-memory footprint fit this (A, L)?
bool eval_fused_fits(gcp_ctx* ctx, bool general)
{
    FusedShape sh;
    size_t smem;
    if (general && (!ctx->t.q_info || !ctx->t.gray_qptr || !ctx->t.g_mask_q)) return false;
    return fused_threads(ctx, sh, &smem, general) != 0;
}

// 256 threads when at least 3 CTAs fit an SM, else 1024 threads (one CTA per SM, 4x the item
// store, so long genomes are not split into hash partitions); 0 if neither fits.
// General kernel: 512 threads when at least 2 CTAs fit, else 1024.
static int fused_threads(gcp_ctx* ctx, FusedShape& sh, size_t* smem, bool general)
{
    if (general) {
        *smem = fused_shape(ctx, sh, 512, true);
        if (*smem != 0 && *smem * 2 + 2 * 1024 <= (size_t)ctx->max_smem_optin) return 512;
        *smem = fused_shape(ctx, sh, 1024, true);
        if (*smem != 0 && *smem <= (size_t)ctx->max_smem_optin) return 1024;
        return 0;
    }
    *smem = fused_shape(ctx, sh, 256, false);
    if (*smem != 0 && *smem * 3 + 3 * 1024 <= (size_t)ctx->max_smem_optin) return 256;
    *smem = fused_shape(ctx, sh, 1024, false);
    if (*smem != 0 && *smem <= (size_t)ctx->max_smem_optin) return 1024;
    *smem = fused_shape(ctx, sh, 256, false);
    if (*smem != 0 && *smem <= (size_t)ctx->max_smem_optin) return 256;
    return 0;
}

static bool fused_compact(gcp_ctx* ctx) { return ctx->t.N <= 65534u; }

static size_t fused_shape(gcp_ctx* ctx, FusedShape& sh, int KF_THREADS, bool general)
{
    const bool compact = fused_compact(ctx);      // u16 genes in shared memory, 32-bit loop-closure keys
    sh.A = ctx->p.num_agents; sh.L = ctx->p.max_dna_len; sh.AL = sh.A * sh.L;
    sh.inv_L = (uint32_t)((0x100000000ull + (uint64_t)sh.L - 1) / (uint64_t)sh.L);
    // expected ~1.4 masked sub-tiles x ~1.6 quarters per executed step (C3: 1,940 +- 80 items for 6 x 150
    // steps before the revisited edges are dropped); an organism with more is redone in 2, 4, ... partitions
    int maxp = ctx->cov_maxpairs > 0 ? ctx->cov_maxpairs : (compact ? (49 * sh.AL) / 24 : (9 * sh.AL) / 4);
    // general kernel: the unmasked store carries ~4.9 quarters per edge (C3: ~3,700 items per organism
    // after the dedup of 6 x 150 steps) plus two markers per (block, quarter) list (~1,100)
    if (general && ctx->cov_maxpairs <= 0) maxp = (41 * sh.AL) / 8;
    if (maxp < 256) maxp = 256;
    if (maxp > KF_PPT * KF_THREADS) maxp = KF_PPT * KF_THREADS;
    maxp = (maxp + 7) & ~7;
    int T = ctx->cov_table;
    if (T <= 0) { T = KF_THREADS; while (T < sh.AL / 3) T <<= 1; }
    if (T > 8192) T = 8192;
    sh.T = T; sh.MAXP = maxp;
    sh.T3 = (sh.AL * 3) / 2 + 64;
    if (sh.T3 > 65535 || sh.AL > 32767) return 0;                  // 16-bit slots, 15-bit member counts
    sh.DS = ctx->cov_dedup ? (uint32_t)T * 4 : 0u;                   // step-dedup hash set in front of s_cst
    sh.R0 = 1;
    while ((size_t)sh.R0 * maxp < (size_t)sh.AL * (general ? 5 : 2) && sh.R0 < 4096) sh.R0 <<= 1;
    sh.off_cst = (uint32_t)T * 20;                                    // over s_sr / s_pay
    const uint32_t cov_b = (uint32_t)T * 20 + (uint32_t)maxp * 8 + 32;  // + pad for B2's read-ahead
    const uint32_t a_b = sh.off_cst + (uint32_t)sh.AL * 16 + GCP_MAX_AGENTS * 16;    // + one pad entry per row
    const uint32_t lc_b = al16((uint32_t)sh.T3 * (compact ? 8 : 12)) + (uint32_t)sh.AL * 4;    // keys + counters, genes (+ owner masks)
    sh.c_overlap = (compact && sh.R0 == 1 && lc_b <= (uint32_t)T * 20 + (uint32_t)maxp * 4) ? 1 : 0;
    uint32_t x_b = cov_b > lc_b ? cov_b : lc_b;
    if (a_b > x_b) x_b = a_b;
    sh.off_info = al16((uint32_t)sh.AL * (compact ? 2 : 4));
    sh.off_x = sh.off_info + al16((uint32_t)sh.AL * 4);
    sh.off_lc = sh.off_x + al16(x_b);
    sh.off_cost = sh.off_lc + al16((uint32_t)sh.A * sh.A * 4);
    return sh.off_cost + (size_t)sh.A * 16;
}

template <class SS>
static bool shape_is(const FusedShape& sh)
{
    return SS::R0 == 1 && sh.R0 == 1 && sh.T == SS::T && sh.MAXP == SS::MAXP && sh.T3 == SS::T3 &&
           sh.off_info == SS::off_info && sh.off_x == SS::off_x && sh.off_cst == SS::off_cst &&
           sh.off_lc == SS::off_lc && sh.DS == SS::DS && (sh.c_overlap != 0) == SS::c_overlap;
}

int launch_eval_fused(gcp_ctx* ctx, const void* dna, int gene16, uint32_t P, const gcp_eval_out* out,
                      cudaStream_t st, bool general)
{
    if (P == 0) return 0;
    if (!general && !ctx->have_masked) return gcp_fail(ctx, -1, "masked footprint store not uploaded");
    if (general && (!ctx->t.q_info || !ctx->t.gray_qptr || !ctx->t.g_mask_q)) return gcp_fail(ctx, -1, "quarter-item tables of the unmasked store not uploaded");
    FusedShape sh;
    size_t smem;
    const int threads = fused_threads(ctx, sh, &smem, general);
    sh.gene16 = gene16;
    if (threads == 0)
        return gcp_fail(ctx, -3, "fused eval kernel needs %zu B shared memory (A*L too large)", smem);
    // a footprint store that outgrows the L2: B2 requests its lines KF_PF_DIST entries ahead of its stream
    // (C5, organisms spread over the whole map: 6.05 -> 5.53 ms per 65,536; requesting every line while B1
    // sorts - bit 0 - gains nothing: 740 resident CTAs x 200 KB of lines is more than the L2 holds)
    sh.prefetch = ctx->cov_prefetch >= 0 ? ctx->cov_prefetch
                                         : ((size_t)(general ? ctx->n_quarters : ctx->m_n_quarters) * 128 > ((size_t)96 << 20) ? 2 : 0);
    StageTimer tm(ctx, 1, st);
    if (general) {
        // any coverage_blend, all six pixel counters: one kernel over the unmasked store
#define GCP_LAUNCH_GENERAL(TH, GRAY, SAL, K32)                                                         \
    do {                                                                                               \
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_eval_fused<TH, GRAY, SAL, K32, true>,                     \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_eval_fused<TH, GRAY, SAL, K32, true><<<P, TH, smem, st>>>(dna, P, ctx->t, sh, ctx->p, ctx->mc, out->obj, \
                                                         out->neg_cov, out->dist, out->turn,           \
                                                         out->lc_mat, out->flags, out->cov);           \
    } while (0)
        const bool k32 = fused_compact(ctx), g = ctx->n_gray > 0;
        // 6 x 200 genes (the reference's six-agent shape): compile-time shared-memory layout
        const bool st1200 = threads == 512 && sh.AL == 1200 &&
                            (k32 ? shape_is<StaticShape<1200, true, true>>(sh) : shape_is<StaticShape<1200, false, true>>(sh));
        if (st1200) {
            if (g) { if (k32) GCP_LAUNCH_GENERAL(512, true, 1200, true); else GCP_LAUNCH_GENERAL(512, true, 1200, false); }
            else   { if (k32) GCP_LAUNCH_GENERAL(512, false, 1200, true); else GCP_LAUNCH_GENERAL(512, false, 1200, false); }
        } else if (threads == 512) {
            if (g) { if (k32) GCP_LAUNCH_GENERAL(512, true, 0, true); else GCP_LAUNCH_GENERAL(512, true, 0, false); }
            else   { if (k32) GCP_LAUNCH_GENERAL(512, false, 0, true); else GCP_LAUNCH_GENERAL(512, false, 0, false); }
        } else {
            if (g) { if (k32) GCP_LAUNCH_GENERAL(1024, true, 0, true); else GCP_LAUNCH_GENERAL(1024, true, 0, false); }
            else   { if (k32) GCP_LAUNCH_GENERAL(1024, false, 0, true); else GCP_LAUNCH_GENERAL(1024, false, 0, false); }
        }
#undef GCP_LAUNCH_GENERAL
        ctx->launches++;
        GCP_CUDA(ctx, cudaGetLastError());
        return 0;
    }
#define GCP_LAUNCH_FUSED_K(TH, GRAY, SAL, K32)                                                         \
    do {                                                                                               \
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_eval_fused<TH, GRAY, SAL, K32>,                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_eval_fused<TH, GRAY, SAL, K32><<<P, TH, smem, st>>>(dna, P, ctx->t, sh, ctx->p, ctx->mc, out->obj, \
                                                         out->neg_cov, out->dist, out->turn,           \
                                                         out->lc_mat, out->flags);                     \
    } while (0)
    // waypoint ids fit 16 bits: compact shared-memory layout (StaticShape), six CTAs per SM
    const bool key32 = fused_compact(ctx);
#define GCP_LAUNCH_FUSED(TH, GRAY, SAL)                                                                \
    do { if (key32) GCP_LAUNCH_FUSED_K(TH, GRAY, SAL, true); else GCP_LAUNCH_FUSED_K(TH, GRAY, SAL, false); } while (0)
    const bool gray = ctx->n_gray > 0;
    // the genome shapes of the reference's driver (pather_driver.py:122-158: max_dna_len 250 / 200 /
    // 150 / 125 / 125 / 200 for 1..6 agents) get the compile-time layout
#define GCP_TRY_STATIC(SAL)                                                                            \
    if (!done && threads == 256 && sh.AL == SAL &&                                                     \
        (key32 ? shape_is<StaticShape<SAL, true>>(sh) : shape_is<StaticShape<SAL, false>>(sh))) {      \
        if (gray) GCP_LAUNCH_FUSED(256, true, SAL); else GCP_LAUNCH_FUSED(256, false, SAL);           \
        done = true;                                                                                   \
    }
    bool done = false;
    // footprint store beyond the L2 (config C5): the variant whose B2 prefetches ahead of its own stream
    if ((sh.prefetch & 2) && !gray && threads == 256 && sh.AL == 1200 &&
        (key32 ? shape_is<StaticShape<1200, true>>(sh) : shape_is<StaticShape<1200, false>>(sh))) {
#define GCP_LAUNCH_PF(K32)                                                                             \
    do {                                                                                               \
        GCP_CUDA(ctx, cudaFuncSetAttribute(k_eval_fused<256, false, 1200, K32, false, true>,           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_eval_fused<256, false, 1200, K32, false, true><<<P, 256, smem, st>>>(dna, P, ctx->t, sh, ctx->p, ctx->mc, \
                                                         out->obj, out->neg_cov, out->dist, out->turn, \
                                                         out->lc_mat, out->flags);                     \
    } while (0)
        if (key32) GCP_LAUNCH_PF(true); else GCP_LAUNCH_PF(false);
#undef GCP_LAUNCH_PF
        done = true;
    }
    GCP_TRY_STATIC(1200)
    GCP_TRY_STATIC(625)
    GCP_TRY_STATIC(500)
    GCP_TRY_STATIC(450)
    GCP_TRY_STATIC(400)
    GCP_TRY_STATIC(250)
#undef GCP_TRY_STATIC
    if (!done) {
        if (threads == 256) { if (gray) GCP_LAUNCH_FUSED(256, true, 0); else GCP_LAUNCH_FUSED(256, false, 0); }
        else                { if (gray) GCP_LAUNCH_FUSED(1024, true, 0); else GCP_LAUNCH_FUSED(1024, false, 0); }
    }
#undef GCP_LAUNCH_FUSED
#undef GCP_LAUNCH_FUSED_K
    ctx->launches++;
    GCP_CUDA(ctx, cudaGetLastError());
    return 0;
}
