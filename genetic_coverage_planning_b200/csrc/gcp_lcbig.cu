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
#include <type_traits>
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
             double* __restrict__ neg_cov_out, uint8_t* __restrict__ flags, const uint8_t* __restrict__ fix = nullptr)
{
    __shared__ int s_lc[GCP_MAX_AGENTS * GCP_MAX_AGENTS];
    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    if (fix && fix[org]) return;              // redone (and finished) by the monolithic kernel
    for (int i = threadIdx.x; i < A * A; i += 32) s_lc[i] = lc[(size_t)org * A * A + i];
    __syncwarp();
    finish_objectives(s_lc, A, prm, mc, cov + (size_t)org * 8, dist + (size_t)org * A,
                          turn + (size_t)org * A, org, obj, neg_cov_out, flags);
}

// ---------------------------------------------------------------------------------------------
// Hash-partitioned shared-memory variant (default for long genomes): one 1024-thread CTA per
// organism.  The duplicate-edge hash table of k_loop_closures is kept, but only for the keys of
// ONE of R hash partitions at a time, so that it fits shared memory for any A*L <= 65,535.
//
// The positions are BUCKETED by partition first: the compacted gene sequence is built in shared
// memory (in the space the tables use later), one scan counts the keys per partition, a second
// scan scatters {key, position} records into the organism's stretch of a global scratch array
// (written and read back by the same CTA: it stays in L2).  Every partition then makes its three
// passes over its own records only - linear, coalesced, key ready - instead of re-scanning all
// positions, recomputing and filtering their keys, R times.  The owner agent of a position is
// found from the row bases instead of a per-position array.
// ---------------------------------------------------------------------------------------------
constexpr int LP_THREADS = 1024;
constexpr int LP_NW = LP_THREADS / 32;
constexpr uint32_t LP_NIL = 0xFFFFFFFFu;
constexpr int LP_MAX_PARTS = 64;

template <typename SEQ>
__device__ __forceinline__ unsigned long long lp_key48(const SEQ* sq, int i, int n)
{
    const unsigned long long f = (uint32_t)sq[i] & 0xFFFFFFu;
    const unsigned long long g = (uint32_t)sq[(i + 1 == n) ? 0 : i + 1] & 0xFFFFFFu;
    return (f << 24) | g;
}
__device__ __forceinline__ uint32_t lp_hash(unsigned long long key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32); }

// SEQ16: node ids fit 16 bits -> the compacted sequence takes 2 bytes per gene (out-of-range genes
// are reported by k_path_costs; here they are truncated).
template <bool SEQ16>
__global__ void __launch_bounds__(LP_THREADS, 1)
k_loop_closures_part(const int32_t* __restrict__ dna, uint32_t P, int A, int L, int T3 /*slots per partition*/,
                     uint32_t R /*pow2 partitions*/, gcp_params prm, gcp_map_consts mc,
                     unsigned long long* __restrict__ rec_all /*[P][A*L] scratch: key48 << 16 | position*/,
                     const double* __restrict__ dist, const double* __restrict__ turn,
                     const uint32_t* __restrict__ cov, int32_t* __restrict__ lc_out,
                     double* __restrict__ obj, double* __restrict__ neg_cov_out, uint8_t* __restrict__ flags,
                     const uint8_t* __restrict__ fix = nullptr)
{
    typedef typename std::conditional<SEQ16, uint16_t, int32_t>::type seq_t;
    if (fix && !fix[blockIdx.x]) return;      // split form: only organisms whose bucket overflowed come here
    extern __shared__ __align__(16) unsigned char smem[];
    const int AL = A * L;
    const size_t tab_b = (size_t)T3 * 20, seq_b = (size_t)AL * sizeof(seq_t);
    const size_t x_b = ((tab_b > seq_b ? tab_b : seq_b) + 15) & ~(size_t)15;
    unsigned long long* s_hkey = reinterpret_cast<unsigned long long*>(smem);       // [T3]
    uint32_t* s_hcnt = reinterpret_cast<uint32_t*>(s_hkey + T3);                     // [T3] count | flags
    uint32_t* s_head = s_hcnt + T3;                                                  // [T3] list head
    uint32_t* s_ownm = s_head + T3;                                                  // [T3] owner mask
    seq_t*    s_seq  = reinterpret_cast<seq_t*>(smem);                               // [A*L] until the buckets are written
    uint16_t* s_next = reinterpret_cast<uint16_t*>(smem + x_b);                      // [A*L]
    int*      s_lc   = reinterpret_cast<int*>(s_next + ((AL + 1) & ~1));             // [A*A]
    __shared__ int s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    __shared__ uint32_t s_pcnt[LP_MAX_PARTS], s_pbase[LP_MAX_PARTS + 1];
    __shared__ int s_fail;

    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* d = dna + (size_t)org * AL;
    unsigned long long* rec = rec_all + (size_t)org * AL;

    if (tid == 0) s_fail = 0;
    if (tid < LP_MAX_PARTS) s_pcnt[tid] = 0;
    for (int i = tid; i < A * A; i += LP_THREADS) s_lc[i] = 0;
    for (int a = warp; a < A; a += LP_NW) {
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
    for (int a = warp; a < A; a += LP_NW) {                      // compaction (row-major, mappy.py:390)
        int run = s_base[a];
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? d[a * L + i] : -1;
            const bool v = g != -1;
            const uint32_t bal = __ballot_sync(FULL, v);
            if (v) s_seq[run + __popc(bal & ((1u << lane) - 1u))] = (seq_t)g;
            run += __popc(bal);
        }
    }
    __syncthreads();
    // bucket the positions by hash partition: count, prefix, scatter (lanes of a warp that hit
    // the same partition share one shared-memory atomic)
    for (int i0 = warp * 32; i0 < n; i0 += LP_THREADS) {
        const int i = i0 + lane;
        const uint32_t p = (i < n) ? (lp_hash(lp_key48(s_seq, i, n) + 1ull) & (R - 1)) : LP_NIL;
        const uint32_t peers = __match_any_sync(FULL, p);
        if (i < n && lane == (int)(__ffs(peers) - 1)) atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t b = 0;
        for (uint32_t r = 0; r < R; ++r) { s_pbase[r] = b; b += s_pcnt[r]; s_pcnt[r] = 0; }
        s_pbase[R] = b;
    }
    __syncthreads();
    for (int i0 = warp * 32; i0 < n; i0 += LP_THREADS) {
        const int i = i0 + lane;
        unsigned long long key = 0;
        uint32_t p = LP_NIL;
        if (i < n) { key = lp_key48(s_seq, i, n); p = lp_hash(key + 1ull) & (R - 1); }
        const uint32_t peers = __match_any_sync(FULL, p);
        uint32_t base = 0;
        const int leader = __ffs(peers) - 1;
        if (i < n && lane == leader) base = atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
        base = __shfl_sync(FULL, base, leader);
        GCP_DASSERT(i >= n || s_pbase[p] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u)) < (uint32_t)n);
        if (i < n)
            rec[s_pbase[p] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (key << 16) | (unsigned long long)i;
    }
    __syncthreads();                                             // the CTA's own global writes are visible to it; s_seq is dead

    const int thresh = prm.solo_sep_thresh;
    for (uint32_t r = 0; r < R; ++r) {
        for (int h = tid; h < T3; h += LP_THREADS) { s_hkey[h] = 0ull; s_hcnt[h] = 0u; s_head[h] = LP_NIL; s_ownm[h] = 0u; }
        __syncthreads();
        const unsigned long long* rc = rec + s_pbase[r];
        const int m = (int)(s_pbase[r + 1] - s_pbase[r]);
        // pass 1: count the keys of this partition
        for (int e0 = tid; e0 < m; e0 += 4 * LP_THREADS) {
          unsigned long long v4[4];                                                // four record loads in flight
          #pragma unroll
          for (int u = 0; u < 4; ++u) { const int e = e0 + u * LP_THREADS; v4[u] = (e < m) ? rc[e] : 0ull; }
          #pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (e0 + u * LP_THREADS >= m) break;
            const unsigned long long key = (v4[u] >> 16) + 1ull;                   // never 0
            uint32_t h = __umulhi(lp_hash(key), (uint32_t)T3);
            int probes = 0;
            for (;;) {
                const unsigned long long cur = ((volatile unsigned long long*)s_hkey)[h];
                if (cur == key) break;
                if (cur == 0ull) {
                    const unsigned long long old = atomicCAS(&s_hkey[h], 0ull, key);
                    if (old == 0ull || old == key) break;
                }
                h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;
                if (++probes >= T3) { s_fail = 1; break; }       // partition far larger than expected
            }
            if (probes < T3) atomicAdd(&s_hcnt[h], 1u);
          }
        }
        __syncthreads();
        if (s_fail) break;                                       // uniform: reported through flags[3]
        // pass 2: members of duplicate groups link themselves, record owner and split flag
        for (int e0 = tid; e0 < m; e0 += 4 * LP_THREADS) {
          unsigned long long v4[4];
          #pragma unroll
          for (int u = 0; u < 4; ++u) { const int e = e0 + u * LP_THREADS; v4[u] = (e < m) ? rc[e] : 0ull; }
          #pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (e0 + u * LP_THREADS >= m) break;
            const unsigned long long v = v4[u];
            const unsigned long long key = (v >> 16) + 1ull;
            const int i = (int)(v & 0xFFFFull);
            uint32_t h = __umulhi(lp_hash(key), (uint32_t)T3);
            while (s_hkey[h] != key) h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;
            if ((s_hcnt[h] & 0xFFFFu) < 2u) continue;
            const int a = owner_row(s_base, A, i);
            s_next[i] = (uint16_t)atomicExch(&s_head[h], (uint32_t)i);      // NIL -> 0xFFFF
            atomicOr(&s_ownm[h], 1u << a);
            if (a >= 1 && i == s_base[a]) atomicOr(&s_hcnt[h], 1u << 16);   // touches a path split (F10)
          }
        }
        __syncthreads();
        // pass 3: successor gap test (ascending consecutive positions)
        for (int e0 = tid; e0 < m; e0 += 4 * LP_THREADS) {
          unsigned long long v4[4];
          #pragma unroll
          for (int u = 0; u < 4; ++u) { const int e = e0 + u * LP_THREADS; v4[u] = (e < m) ? rc[e] : 0ull; }
          #pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (e0 + u * LP_THREADS >= m) break;
            const unsigned long long v = v4[u];
            const unsigned long long key = (v >> 16) + 1ull;
            const int i = (int)(v & 0xFFFFull);
            uint32_t h = __umulhi(lp_hash(key), (uint32_t)T3);
            while (s_hkey[h] != key) h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;
            const uint32_t c = s_hcnt[h];
            if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
            uint32_t succ = 0xFFFFFFFFu;
            for (uint32_t j = s_head[h]; j != 0xFFFFu && j != LP_NIL; j = s_next[j])
                if (j > (uint32_t)i && j < succ) succ = j;
            if (succ != 0xFFFFFFFFu && (int)(succ - (uint32_t)i) > thresh) atomicOr(&s_hcnt[h], 1u << 17);
          }
        }
        __syncthreads();
        // pass 4: one thread per duplicate group updates lc_mat
        for (int h = tid; h < T3; h += LP_THREADS) {
            const uint32_t c = s_hcnt[h];
            if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
            const uint32_t owners = s_ownm[h];
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
    }
    if (lc_out)
        for (int i = tid; i < A * A; i += LP_THREADS) lc_out[(size_t)org * A * A + i] = s_lc[i];
    if (tid < 32) {                       // one warp, cooperatively
        finish_objectives(s_lc, A, prm, mc, cov + (size_t)org * 8, dist + (size_t)org * A,
                          turn + (size_t)org * A, org, obj, neg_cov_out, flags);
        if (tid == 0 && s_fail && flags) flags[(size_t)org * 4 + 3] |= 4;
    }
}


// ---------------------------------------------------------------------------------------------
// Split form of the hash-partitioned kernel: the K partitions of ALL organisms become independent work
// items for small CTAs (256 threads, several per SM) instead of K sequential passes inside one
// 1,024-thread CTA per organism (one per SM, issue slots ~50 % busy between its barriers).
//   k_lc_bucket  one CTA per organism: compaction, {key, position} records scattered by key partition -
//                partition r owns the r-th K-th of the organism's scratch stretch; {K, n, row bases, records
//                per partition} to hdr, K work items to the list (an overflowing partition: fix[org])
//   k_lc_part    persistent CTAs take work items: the four passes of the monolithic kernel over ONE bucket,
//                chains by record index, lc_mat contributions added to the organism's matrix in global memory
//   k_lcb_finish constraints + objectives; k_loop_closures_part(fix) redoes flagged organisms
// ---------------------------------------------------------------------------------------------
constexpr uint32_t LS_MAXK = 64;
constexpr int LS_THREADS = 256;
constexpr uint32_t LS_HDR = 2 + GCP_MAX_AGENTS + 1 + LS_MAXK;      // K, n, base[A + 1], records per partition

template <bool SEQ16>
__global__ void __launch_bounds__(LP_THREADS, 1)
k_lc_bucket(const int32_t* __restrict__ dna, uint32_t P, int A, int L, uint32_t K, uint32_t region,
            unsigned long long* __restrict__ rec_all, uint32_t* __restrict__ hdr_all,
            uint32_t* __restrict__ work /* [0] = items listed, [4 + 4 i ..] = item i: {organism, partition, records, K} */,
            uint8_t* __restrict__ fix)
{
    typedef typename std::conditional<SEQ16, uint16_t, int32_t>::type seq_t;
    extern __shared__ __align__(16) unsigned char smem[];
    seq_t* s_seq = reinterpret_cast<seq_t*>(smem);                                   // [A*L]
    __shared__ int s_len[GCP_MAX_AGENTS], s_base[GCP_MAX_AGENTS + 1];
    __shared__ uint32_t s_pcnt[LS_MAXK], s_wbase, s_ok;
    const uint32_t org = blockIdx.x;
    if (org >= P) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AL = A * L;
    const int32_t* d = dna + (size_t)org * AL;
    unsigned long long* rec = rec_all + (size_t)org * K * region;
    uint32_t* hdr = hdr_all + (size_t)org * LS_HDR;
    if (tid < (int)LS_MAXK) s_pcnt[tid] = 0;
    for (int a = warp; a < A; a += LP_NW) {
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
    for (int a = warp; a < A; a += LP_NW) {                      // compaction (row-major, mappy.py:390)
        int run = s_base[a];
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int g = (i < L) ? d[a * L + i] : -1;
            const bool v = g != -1;
            const uint32_t bal = __ballot_sync(FULL, v);
            if (v) s_seq[run + __popc(bal & ((1u << lane) - 1u))] = (seq_t)g;
            run += __popc(bal);
        }
    }
    __syncthreads();
    for (int i0 = warp * 32; i0 < n; i0 += LP_THREADS) {
        const int i = i0 + lane;
        unsigned long long key = 0;
        uint32_t p = LP_NIL;
        if (i < n) { key = lp_key48(s_seq, i, n); p = lp_hash(key + 1ull) & (K - 1); }
        const uint32_t peers = __match_any_sync(FULL, p);
        uint32_t base = 0;
        const int leader = __ffs(peers) - 1;
        if (i < n && lane == leader) base = atomicAdd(&s_pcnt[p], (uint32_t)__popc(peers));
        base = __shfl_sync(FULL, base, leader);
        const uint32_t at = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (i < n && at < region) rec[(size_t)p * region + at] = (key << 16) | (unsigned long long)i;
    }
    __syncthreads();
    if (tid == 0) {
        bool over = false;
        for (uint32_t r = 0; r < K; ++r) over |= s_pcnt[r] > region;
        hdr[0] = K; hdr[1] = (uint32_t)n;
        s_ok = over ? 0u : 1u;
        if (over) fix[org] = 1;
        else if (n > 0) s_wbase = atomicAdd(&work[0], K);
    }
    __syncthreads();
    if (tid <= A) hdr[2 + tid] = (uint32_t)s_base[tid];
    if (tid < (int)K) hdr[2 + GCP_MAX_AGENTS + 1 + tid] = s_pcnt[tid];
    if (s_ok && n > 0 && tid < (int)K)
        reinterpret_cast<uint4*>(work + 4)[s_wbase + tid] = make_uint4(org, (uint32_t)tid, s_pcnt[tid], K);
}

// KEY32: waypoint ids fit 16 bits - the table key is (from << 16 | to) + 1 in 32 bits (16 B per slot, 32-bit CAS)
template <bool KEY32>
__global__ void __launch_bounds__(LS_THREADS, KEY32 ? 6 : 5)
k_lc_part(int A, int T3max, uint32_t region, int thresh, const unsigned long long* __restrict__ rec_all,
          const uint32_t* __restrict__ hdr_all, uint32_t* __restrict__ work, int32_t* __restrict__ lc_acc /*[P][A][A]*/)
{
    using HK = typename std::conditional<KEY32, uint32_t, unsigned long long>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    HK* s_hkey = reinterpret_cast<HK*>(smem);                                         // [T3max]
    uint32_t* s_hcnt = reinterpret_cast<uint32_t*>(s_hkey + T3max);                   // [T3max] members | split << 16 | gap << 17
    uint32_t* s_head = s_hcnt + T3max;                                                // [T3max] newest member (record index)
    uint32_t* s_ownm = s_head + T3max;                                                // [T3max] owner mask
    int*      s_lc   = reinterpret_cast<int*>(s_ownm + T3max);                        // [A*A]
    uint16_t* s_pos  = reinterpret_cast<uint16_t*>(s_lc + A * A);                     // [region] position of record e
    uint16_t* s_nxt  = s_pos + region;                                                // [region] next member of e's group
    uint16_t* s_slot = s_nxt + region;                                                // [region] hash slot of e's key
    __shared__ int s_base[GCP_MAX_AGENTS + 1];
    const int tid = threadIdx.x;
    const uint32_t n_work = work[0];
    const uint4* items = reinterpret_cast<const uint4*>(work + 4);
    // items are dealt out statically (CTA b takes b, b + grid, ...) and the next item's descriptor is loaded
    // while the current one is processed: no ticket, one dependent L2 round trip (the records) per item
    uint32_t w = blockIdx.x;
    uint4 cur = (w < n_work) ? items[w] : make_uint4(0, 0, 0, 0);
    for (; w < n_work; w += gridDim.x) {
        const uint4 nxt = (w + gridDim.x < n_work) ? items[w + gridDim.x] : make_uint4(0, 0, 0, 0);
        const uint32_t org = cur.x, r = cur.y, K = cur.w;
        const int m = (int)cur.z;
        cur = nxt;
        if (m == 0) continue;
        const uint32_t* hdr = hdr_all + (size_t)org * LS_HDR;
        __syncthreads();                                  // the previous item's readers are done
        const unsigned long long* rc = rec_all + ((size_t)org * K + r) * region;
        // the arrays are laid out for the largest bucket (T3max slots); this bucket uses 1.5 slots per record,
        // so clearing and the slot scan of pass 4 cost what the bucket holds, not what it may hold
        const int T3 = min(T3max, max(64, m + (m >> 1) + 8));
        for (int h = tid; h < T3; h += LS_THREADS) { s_hkey[h] = (HK)0; s_hcnt[h] = 0u; s_head[h] = LP_NIL; s_ownm[h] = 0u; }
        for (int i = tid; i < A * A; i += LS_THREADS) s_lc[i] = 0;
        if (tid <= A) s_base[tid] = (int)hdr[2 + tid];
        __syncthreads();
        // pass 1: count the keys of this partition
        for (int e0 = tid; e0 < m; e0 += 4 * LS_THREADS) {
            unsigned long long v4[4];                                                // four record loads in flight
            #pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * LS_THREADS; v4[u] = (e < m) ? rc[e] : 0ull; }
            #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * LS_THREADS;
                if (e >= m) break;
                const unsigned long long k48 = v4[u] >> 16;
                HK key;                                                              // never 0
                uint32_t h;
                if (KEY32) {
                    key = (HK)(((uint32_t)(k48 >> 24) << 16 | (uint32_t)(k48 & 0xFFFFull)) + 1u);
                    h = __umulhi((uint32_t)key * 0x9E3779B1u, (uint32_t)T3);
                } else {
                    key = (HK)(k48 + 1ull);
                    h = __umulhi(lp_hash(k48 + 1ull), (uint32_t)T3);
                }
                for (;;) {                                                           // T3 > region >= m: always terminates
                    const HK cur = ((volatile HK*)s_hkey)[h];
                    if (cur == key) break;
                    if (cur == (HK)0) {
                        const HK old = atomicCAS(&s_hkey[h], (HK)0, key);
                        if (old == (HK)0 || old == key) break;
                    }
                    h = (h + 1 == (uint32_t)T3) ? 0u : h + 1;
                }
                atomicAdd(&s_hcnt[h], 1u);
                GCP_DASSERT(e < (int)region && h < (uint32_t)T3);
                s_pos[e] = (uint16_t)(v4[u] & 0xFFFFull);
                s_slot[e] = (uint16_t)h;
            }
        }
        __syncthreads();
        // pass 2: members of duplicate groups link themselves, record owner and split flag
        for (int e = tid; e < m; e += LS_THREADS) {
            const uint32_t h = s_slot[e];
            if ((s_hcnt[h] & 0xFFFFu) < 2u) continue;
            const int i = (int)s_pos[e];
            const int a = owner_row(s_base, A, i);
            s_nxt[e] = (uint16_t)atomicExch(&s_head[h], (uint32_t)e);               // NIL -> 0xFFFF
            atomicOr(&s_ownm[h], 1u << a);
            if (a >= 1 && i == s_base[a]) atomicOr(&s_hcnt[h], 1u << 16);            // touches a path split (F10)
        }
        __syncthreads();
        // pass 3: successor gap test (ascending consecutive positions)
        for (int e = tid; e < m; e += LS_THREADS) {
            const uint32_t h = s_slot[e];
            const uint32_t c = s_hcnt[h];
            if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
            const uint32_t i = s_pos[e];
            uint32_t succ = 0xFFFFFFFFu;
            for (uint32_t j = s_head[h] & 0xFFFFu; j != 0xFFFFu; j = s_nxt[j]) {
                const uint32_t pj = s_pos[j];
                if (pj > i && pj < succ) succ = pj;
            }
            if (succ != 0xFFFFFFFFu && (int)(succ - i) > thresh) atomicOr(&s_hcnt[h], 1u << 17);
        }
        __syncthreads();
        // pass 4: one thread per duplicate group updates lc_mat
        for (int h = tid; h < T3; h += LS_THREADS) {
            const uint32_t c = s_hcnt[h];
            if ((c & 0xFFFFu) < 2u || (c & (1u << 16))) continue;
            const uint32_t owners = s_ownm[h];
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
        int32_t* out = lc_acc + (size_t)org * A * A;
        for (int i = tid; i < A * A; i += LS_THREADS) { const int v = s_lc[i]; if (v) atomicAdd(&out[i], v); }
    }
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
    if (ctx->force_lc_big != 1) {
        // hash-partitioned shared-memory kernel: pick the smallest R whose tables fit; keeping the
        // sequence in shared memory (16-bit node ids) only when that does not cost a partition
        const bool seq16 = ctx->t.N <= 65536;
        const size_t seq_b = AL * (seq16 ? 2 : 4);
        const size_t fixed = ((AL + 1) & ~(size_t)1) * 2 + (size_t)A * A * 4 + 64;
        uint32_t R = 0; int T3 = 0; size_t smem = 0;
        for (uint32_t Rt = 1; Rt <= (uint32_t)LP_MAX_PARTS; Rt <<= 1) {
            if ((int)Rt < ctx->lc_min_parts) continue;
            const int T3t = (int)((AL * 3) / (2 * Rt)) + 64 + (Rt > 1 ? (int)(AL / (4 * Rt)) : 0);   // slack: uneven partitions
            const size_t tab = (size_t)T3t * 20;
            const size_t need_s = (((tab > seq_b ? tab : seq_b) + 15) & ~(size_t)15) + fixed;
            if (need_s <= (size_t)ctx->max_smem_optin) { R = Rt; T3 = T3t; smem = need_s; break; }
        }
        if (R > 1 && ctx->lc_split) {
            // split form: partitions as work items of 256-thread CTAs (see k_lc_bucket)
            const size_t keys_per = ctx->lc_split_keys > 0 ? (size_t)ctx->lc_split_keys : 640;   // genes aimed at per partition
            uint32_t K = 2;
            while (K < LS_MAXK && (size_t)K * keys_per < AL) K <<= 1;
            const size_t slack = ctx->lc_split_slack > 0 ? (size_t)ctx->lc_split_slack : 150;
            uint32_t region = (uint32_t)((AL * slack / 100 + K - 1) / K);   // 1.5 x the mean bucket (sigma of a bucket: ~5 % of it)
            region = (region + 7u) & ~7u;
            const int T3p = (int)(region * 3 / 2) + 64;
            const bool key32 = ctx->t.N <= 65534u;
            const size_t smem_p = (size_t)T3p * (key32 ? 16 : 20) + (size_t)A * A * 4 + (size_t)region * 6 + 16;
            const size_t smem_b = seq_b + 16;
            if (region < 65000u && smem_p * 2 <= (size_t)ctx->max_smem_optin && smem_b <= (size_t)ctx->max_smem_optin) {
                // records per organism: K stretches of the split form, and at least the A*L the monolithic fallback indexes
                const size_t rec_per = (size_t)K * region > AL ? (size_t)K * region : AL;
                uint32_t chunk = (uint32_t)(((size_t)2 << 30) / (rec_per * 8));
                if (chunk < 1) chunk = 1;
                if (chunk > P) chunk = P;
                const size_t off_hdr = al256((size_t)chunk * rec_per * 8);
                const size_t off_work = off_hdr + al256((size_t)chunk * LS_HDR * 4);
                const size_t off_fix = off_work + al256((4 + (size_t)chunk * K * 4) * 4);
                const size_t off_lc = off_fix + al256(chunk);
                const size_t need = off_lc + (lc_out ? 0 : al256((size_t)chunk * A * A * 4));
                if (ctx->lcb_scratch_bytes_s[ctx->scratch_slot] < need) {
                    if (ctx->lcb_scratch_s[ctx->scratch_slot]) cudaFree(ctx->lcb_scratch_s[ctx->scratch_slot]);
                    ctx->lcb_scratch_s[ctx->scratch_slot] = nullptr; ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = 0;
                    GCP_CUDA(ctx, cudaMalloc(&ctx->lcb_scratch_s[ctx->scratch_slot], need));
                    ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = need;
                }
                char* base = reinterpret_cast<char*>(ctx->lcb_scratch_s[ctx->scratch_slot]);
                unsigned long long* rec = reinterpret_cast<unsigned long long*>(base);
                uint32_t* hdr = reinterpret_cast<uint32_t*>(base + off_hdr);
                uint32_t* work = reinterpret_cast<uint32_t*>(base + off_work);
                uint8_t* fix = reinterpret_cast<uint8_t*>(base + off_fix);
                StageTimer tm(ctx, 2, st);
                if (seq16) {
                    GCP_CUDA(ctx, cudaFuncSetAttribute(k_lc_bucket<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
                    GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures_part<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                } else {
                    GCP_CUDA(ctx, cudaFuncSetAttribute(k_lc_bucket<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
                    GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures_part<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                }
                if (key32) GCP_CUDA(ctx, cudaFuncSetAttribute(k_lc_part<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
                else       GCP_CUDA(ctx, cudaFuncSetAttribute(k_lc_part<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
                int per_sm = (int)((size_t)ctx->max_smem_optin / (smem_p + 1024));
                if (per_sm > (key32 ? 6 : 5)) per_sm = key32 ? 6 : 5;
                if (per_sm < 1) per_sm = 1;
                for (uint32_t o0 = 0; o0 < P; o0 += chunk) {
                    const uint32_t n_org = (P - o0 < chunk) ? (P - o0) : chunk;
                    const int32_t* dna_c = dna + (size_t)o0 * AL;
                    int32_t* lc_c = lc_out ? lc_out + (size_t)o0 * A * A : reinterpret_cast<int32_t*>(base + off_lc);
                    double* obj_c = obj ? obj + (size_t)o0 * 2 : nullptr;
                    double* neg_c = neg_cov ? neg_cov + (size_t)o0 : nullptr;
                    uint8_t* flg_c = flags ? flags + (size_t)o0 * 4 : nullptr;
                    GCP_CUDA(ctx, cudaMemsetAsync(work, 0, 16, st));
                    GCP_CUDA(ctx, cudaMemsetAsync(fix, 0, n_org, st));
                    GCP_CUDA(ctx, cudaMemsetAsync(lc_c, 0, (size_t)n_org * A * A * 4, st));
                    if (seq16) k_lc_bucket<true><<<n_org, LP_THREADS, smem_b, st>>>(dna_c, n_org, A, L, K, region, rec, hdr, work, fix);
                    else       k_lc_bucket<false><<<n_org, LP_THREADS, smem_b, st>>>(dna_c, n_org, A, L, K, region, rec, hdr, work, fix);
                    if (key32) k_lc_part<true><<<ctx->num_sms * per_sm, LS_THREADS, smem_p, st>>>(A, T3p, region, ctx->p.solo_sep_thresh, rec, hdr, work, lc_c);
                    else       k_lc_part<false><<<ctx->num_sms * per_sm, LS_THREADS, smem_p, st>>>(A, T3p, region, ctx->p.solo_sep_thresh, rec, hdr, work, lc_c);
                    k_lcb_finish<<<n_org, 32, 0, st>>>(n_org, A, ctx->p, ctx->mc, lc_c, dist + (size_t)o0 * A, turn + (size_t)o0 * A,
                                                       cov + (size_t)o0 * 8, obj_c, neg_c, flg_c, fix);
                    // organisms whose bucket overflowed (fix): the monolithic kernel redoes and finishes them.  Its record
                    // scratch is the split form's stretch of the organism (2 * A*L records >= the A*L it needs).
                    if (seq16)
                        k_loop_closures_part<true><<<n_org, LP_THREADS, smem, st>>>(dna_c, n_org, A, L, T3, R, ctx->p, ctx->mc, rec,
                                                                                    dist + (size_t)o0 * A, turn + (size_t)o0 * A, cov + (size_t)o0 * 8,
                                                                                    lc_out ? lc_c : nullptr, obj_c, neg_c, flg_c, fix);
                    else
                        k_loop_closures_part<false><<<n_org, LP_THREADS, smem, st>>>(dna_c, n_org, A, L, T3, R, ctx->p, ctx->mc, rec,
                                                                                     dist + (size_t)o0 * A, turn + (size_t)o0 * A, cov + (size_t)o0 * 8,
                                                                                     lc_out ? lc_c : nullptr, obj_c, neg_c, flg_c, fix);
                    ctx->launches += 4;
                }
                GCP_CUDA(ctx, cudaGetLastError());
                return 0;
            }
        }
        if (R != 0) {
            // record scratch: 8 bytes per gene, at most 2 GiB (organisms go through in chunks)
            uint32_t chunk = (uint32_t)(((size_t)2 << 30) / (AL * 8));
            if (chunk < 1) chunk = 1;
            if (chunk > P) chunk = P;
            const size_t need = al256((size_t)chunk * AL * 8);
            if (ctx->lcb_scratch_bytes_s[ctx->scratch_slot] < need) {
                if (ctx->lcb_scratch_s[ctx->scratch_slot]) cudaFree(ctx->lcb_scratch_s[ctx->scratch_slot]);
                ctx->lcb_scratch_s[ctx->scratch_slot] = nullptr; ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = 0;
                GCP_CUDA(ctx, cudaMalloc(&ctx->lcb_scratch_s[ctx->scratch_slot], need));
                ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = need;
            }
            StageTimer tm(ctx, 2, st);
            unsigned long long* rec = reinterpret_cast<unsigned long long*>(ctx->lcb_scratch_s[ctx->scratch_slot]);
            if (seq16) GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures_part<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else       GCP_CUDA(ctx, cudaFuncSetAttribute(k_loop_closures_part<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (uint32_t o0 = 0; o0 < P; o0 += chunk) {
                const uint32_t n_org = (P - o0 < chunk) ? (P - o0) : chunk;
                const int32_t* dna_c = dna + (size_t)o0 * AL;
                const double* dist_c = dist + (size_t)o0 * A;
                const double* turn_c = turn + (size_t)o0 * A;
                const uint32_t* cov_c = cov + (size_t)o0 * 8;
                int32_t* lc_c = lc_out ? lc_out + (size_t)o0 * A * A : nullptr;
                double* obj_c = obj ? obj + (size_t)o0 * 2 : nullptr;
                double* neg_c = neg_cov ? neg_cov + (size_t)o0 : nullptr;
                uint8_t* flg_c = flags ? flags + (size_t)o0 * 4 : nullptr;
                if (seq16)
                    k_loop_closures_part<true><<<n_org, LP_THREADS, smem, st>>>(dna_c, n_org, A, L, T3, R, ctx->p, ctx->mc, rec,
                                                                                dist_c, turn_c, cov_c, lc_c, obj_c, neg_c, flg_c);
                else
                    k_loop_closures_part<false><<<n_org, LP_THREADS, smem, st>>>(dna_c, n_org, A, L, T3, R, ctx->p, ctx->mc, rec,
                                                                                 dist_c, turn_c, cov_c, lc_c, obj_c, neg_c, flg_c);
                ctx->launches++;
            }
            GCP_CUDA(ctx, cudaGetLastError());
            return 0;
        }
    }
    const int nb = bits_for(ctx->t.N - 1);
    // chunk of organisms sorted at once: <= 16 Mi items, organism index must fit the key
    uint32_t chunk = (uint32_t)((16u << 20) / AL);
    if (chunk < 1) chunk = 1;
    if (chunk > P) chunk = P;
    while (bits_for(chunk - 1) + 2 * nb > 63 && chunk > 1) chunk >>= 1;
    const size_t M = (size_t)chunk * AL;
    const size_t need = 2 * al256(M * 8) + 2 * al256(M * 4) + al256(M * 4) +
                        al256(radix_hist_bytes((uint32_t)M)) + al256((size_t)P * A * A * 4) + 256;
    if (ctx->lcb_scratch_bytes_s[ctx->scratch_slot] < need) {
        if (ctx->lcb_scratch_s[ctx->scratch_slot]) cudaFree(ctx->lcb_scratch_s[ctx->scratch_slot]);
        ctx->lcb_scratch_s[ctx->scratch_slot] = nullptr; ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = 0;
        GCP_CUDA(ctx, cudaMalloc(&ctx->lcb_scratch_s[ctx->scratch_slot], need));
        ctx->lcb_scratch_bytes_s[ctx->scratch_slot] = need;
    }
    char* p = reinterpret_cast<char*>(ctx->lcb_scratch_s[ctx->scratch_slot]);
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
