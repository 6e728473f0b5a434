// Shared declarations of the B200 fitness engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/gcp.h"

#define GCP_NUM_STAGES 8

// Debug build (libgcp_debug.so, -DGCP_BOUNDS_CHECK): every index the kernels form into shared
// memory or scratch is checked before use; a violation traps the kernel with file:line (device
// assert -> cudaErrorAssert on the host).  Compiled out of the release library.
#ifdef GCP_BOUNDS_CHECK
#include <cassert>
#define GCP_DASSERT(cond) assert(cond)
#else
#define GCP_DASSERT(cond) ((void)0)
#endif

// Device-side view of the static lookup tables (uploaded once, read-only afterwards).
struct GcpTables {
    // graph (pathmaker.py:157-173)
    uint32_t N, E;
    const uint32_t* row_ptr;   // [N+1]
    const uint32_t* col;       // [E]
    const int2*     xy;        // [N] (row, col)
    const uint8_t*  xover;     // [N]
    // edges (mappy.py:274-328): ids [0,E) real, [E,E+N) null edge of node id-E
    const double2*  edge_cost; // {dist, turn}
    const double*   edge_theta;
    const uint4*    adjc16;    // [N] 8 x u16 out-neighbour ids (0xFFFF = empty); NULL if N > 65535 or degree > 8
    const uint2*    adj8;      // [N][8] {col, float bits of theta}, col = ~0 for empty; NULL if max degree > 8
    const uint32_t* sub_info;  // [E+N] first sub-tile (28 bits) | count << 28
    const uint2*    sub_desc;  // [S] {block id, (offset128<<4)|qmask}
    const uint4*    tile;      // quarters: 8 x uint4 = 16 rows of 64 px
    const uint32_t* q_info;    // [E+N] first quarter item (26 bits) | count << 26 over the UNMASKED store (general fused kernel)
    const uint2*    q_desc;    // {block id << 2 | quarter index, index of the 128-B quarter in tile}
    // general fused kernel: the (g == 0) and mask & (g == 0) planes are appended to the tile allocation, so
    // that ONE 30-bit quarter index addresses footprint quarters and plane quarters alike (0 = not built)
    uint32_t g_free_q, g_mask_q;
    // second store: footprints pre-ANDed with buffer_mask (optional)
    const uint32_t* m_sub_info;
    const uint2*    m_sub_desc;
    const uint4*    m_tile;
    const uint32_t* m_q_info;  // [E+N] first quarter item (26 bits) | count << 26   (fused kernel)
    const uint2*    m_q_desc;  // {block id << 2 | quarter index, uint4 index of the quarter in m_tile}
    // map planes, block-major: 32 x uint4 per block (lane l = rows 2l, 2l+1)
    uint32_t n_blocks;
    const uint4*    blk_mask0; // buffer_mask & (g==0)
    const uint4*    blk_free0; // (g==0)
    const uint32_t* gray_ptr;  // [n_blocks+1]
    const uint32_t* gray_qptr; // [4*n_blocks+1] the same entries per (block, quarter): a block's entries are ordered by quarter
    const uint32_t* gray_ent;  // pos(12) | (255-g)<<12 | in_mask<<20
};

struct gcp_ctx {
    int device = -1;
    std::string err;
    GcpTables t{};
    gcp_map_consts mc{};
    gcp_params p{};
    gcp_vary_params vary{};
    bool have_vary = false;
    bool have_graph = false, have_edges = false, have_map = false, have_params = false;
    bool have_masked = false;
    uint64_t n_sub = 0, n_quarters = 0, n_gray = 0;
    uint32_t max_degree = 0;
    std::vector<uint32_t> h_row_ptr, h_col;   // host copies until the edges are uploaded
    uint64_t launches = 0;
    int max_smem_optin = 0;
    int num_sms = 0;
    // owned device allocations
    void* owned[32] = {};
    int n_owned = 0;
    // population shards: own allocations exported over CUDA IPC, and peers' shards opened here
    std::vector<void*> shard_own, shard_peer;
    // timing
    int timing = 0;
    cudaEvent_t ev_beg[GCP_NUM_STAGES] = {}, ev_end[GCP_NUM_STAGES] = {};
    bool ev_valid[GCP_NUM_STAGES] = {};
    // host-eval staging (gcp_eval_host)
    void* stage_dev = nullptr; size_t stage_bytes = 0;
    // int32 host chromosomes narrowed to int16 by host threads on the way (gcp_host.cu)
    void* host_pool = nullptr;
    void* host_stage16 = nullptr; size_t host_stage16_bytes = 0;
    void* hstage_dev[3] = {}; size_t hstage_dev_bytes = 0;
    int host_narrow = -1;    // gcp_eval_host narrows int32 genes to int16 on the host when N <= 32767 (fused path): 1 always, 0 never,
                             // -1 auto = one rank on this host and >= 6 worker threads (gcp_host.cu: host_narrow_auto)
    int host_threads = 0;    // worker threads of the narrowing (0 = auto: cores / LOCAL_WORLD_SIZE - 1, at most 12)
    void* vary_scratch = nullptr; size_t vary_scratch_bytes = 0;   // mutation row list (gcp_vary.cu)
    // long-genome scratch, one set per pipeline slot: gcp_eval_host overlaps neighbouring chunks on
    // three streams, and a chunk's kernels must not overwrite records another chunk still reads
    void* lcb_scratch_s[3] = {}; size_t lcb_scratch_bytes_s[3] = {};   // sort buffers / position records (gcp_lcbig.cu)
    void* cov_scratch_s[3] = {}; size_t cov_scratch_bytes_s[3] = {};   // item buckets of long genomes (gcp_fused.cu)
    int scratch_slot = 0;           // slot the launchers use (0 outside gcp_eval_host)
    uint32_t vary_org_offset = 0;   // population shards: global id of local organism 0 (Philox keys)
    int force_lc_big = 0;    // testing: always take the sort-based loop-closure path
    cudaStream_t hstream[3] = {};
    // coverage kernel tuning (0 = auto)
    int cov_table = 0, cov_maxpairs = 0;
    int cov_buckets = 1;     // 1: long genomes bucket their items by hash partition once (else: R scanning passes)
    int cov_dedup = 1;       // 1: revisited edges are dropped before the coverage union (phase B0)
    int cov_split = 1;       // 1: long genomes run their hash partitions as work items of small CTAs (k_cov_bucket / k_cov_part)
    int cov_split_cap = 0;   // tuning / testing: items aimed at per partition (0 = 2,600); a huge value overflows every partition CTA -> fallback path
    int lc_split_keys = 0;   // tuning: genes aimed at per loop-closure partition (0 = 640)
    int cov_split_k = 0;     // tuning: partitions per organism fixed (power of two) instead of chosen per organism
    int lc_split_slack = 0;  // testing: stretch of a loop-closure bucket in % of the mean bucket (0 = 150); < 100 overflows -> fallback path
    int cov_threads = 0;     // tuning: CTA size of k_cov_quarters (256 / 512 / 1024; 0 = auto)
    int lc_min_parts = 0;    // tuning: at least this many hash partitions in k_loop_closures_part
    int lc_split = 1;        // 1: long genomes run their loop-closure partitions as work items of small CTAs (k_lc_bucket / k_lc_part)
    int force_general = 0;   // testing: never take the pre-masked fast path
    int cov_detail = 1;      // 1: all six counters; 0: only what the objectives need
    int fused_eval = 1;      // 1: objectives-only fast path runs as ONE kernel (gcp_fused.cu)
    int cov_prefetch = -1;   // -1 auto (store > 96 MiB: 2), bit 0: L2 prefetch of the footprint lines in phase B1, bit 1: look-ahead in B2
    uint64_t m_n_quarters = 0;
    int rank_algo = 0;       // 0 auto, 1 O(N^2) dominance tiles, 2 sort-based (gcp_rank.cu)
    int sort_onesweep = 1;   // 1: radix sorts of <= 2^20 pairs run one launch per pass (decoupled look-back)
};

int gcp_fail(gcp_ctx* ctx, int code, const char* fmt, ...);
#define GCP_CUDA(ctx, call)                                                        \
    do { cudaError_t e__ = (call);                                                 \
         if (e__ != cudaSuccess)                                                   \
             return gcp_fail(ctx, -2, "%s failed: %s (%s:%d)", #call,             \
                             cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

struct StageTimer {
    gcp_ctx* c; int s; cudaStream_t st;
    StageTimer(gcp_ctx* ctx, int stage, cudaStream_t stream) : c(ctx), s(stage), st(stream) {
        if (c->timing) cudaEventRecord(c->ev_beg[s], st);
    }
    ~StageTimer() {
        if (c->timing) { cudaEventRecord(c->ev_end[s], st); c->ev_valid[s] = true; }
    }
};

// kernel launchers (defined in the .cu files)
int launch_path_costs(gcp_ctx*, const int32_t* dna, uint32_t P, uint32_t* edge_ids,
                      uint32_t* step_info, double* dist, double* turn, uint8_t* flags,
                      cudaStream_t, int quarter_info = 0);
int launch_coverage_quarters(gcp_ctx*, uint32_t* step_qinfo /* duplicates are zeroed in place */, uint32_t P, uint32_t* cov,
                             uint8_t* flags, cudaStream_t);
int launch_coverage_masked(gcp_ctx*, const uint32_t* step_info, uint32_t P, uint32_t* cov,
                           uint8_t* flags, cudaStream_t);
int launch_coverage(gcp_ctx*, const uint32_t* edge_ids, uint32_t P, uint32_t* cov,
                    uint8_t* flags, cudaStream_t);
int launch_loop_closures(gcp_ctx*, const int32_t* dna, uint32_t P, const double* dist,
                         const double* turn, const uint32_t* cov, int32_t* lc, double* obj,
                         double* neg_cov, uint8_t* flags, cudaStream_t);
int launch_eval_fused(gcp_ctx*, const void* dna, int gene16, uint32_t P, const gcp_eval_out* out,
                      cudaStream_t, bool general = false);
bool eval_fused_fits(gcp_ctx*, bool general = false);
int eval_host_narrowed(gcp_ctx*, const int32_t* dna_host, uint32_t P, double* obj_host, uint8_t* flags_host);
void gcp_host_free(gcp_ctx*);
bool host_narrow_auto();
int launch_loop_closures_big(gcp_ctx*, const int32_t* dna, uint32_t P, const double* dist,
                             const double* turn, const uint32_t* cov, int32_t* lc, double* obj,
                             double* neg_cov, uint8_t* flags, cudaStream_t);
int launch_maximin(gcp_ctx*, const double* obj, uint32_t n, double constr, uint32_t r0,
                   uint32_t r1, double* obj_sc, double* fit, void* scratch, size_t sb,
                   cudaStream_t);
int launch_arena_order(gcp_ctx*, const double* fit, uint32_t n, int32_t* order,
                       uint8_t* nondom, void* scratch, size_t sb, cudaStream_t);

int launch_low_var_select(gcp_ctx*, const double* fit, uint32_t M, double gamma, double r0,
                          double* scratch, int32_t* out, cudaStream_t);
int launch_crossover(gcp_ctx*, const int32_t* parents, const int32_t* strong, uint32_t P,
                     uint64_t seed, uint32_t gen, int32_t* children, cudaStream_t);
int launch_mutate(gcp_ctx*, int32_t* dna, uint32_t P, uint64_t seed, uint32_t gen, uint32_t salt,
                  cudaStream_t);
int launch_mutate_genes(gcp_ctx*, void* dna, int gene_bytes, uint32_t P, uint32_t org_offset, uint64_t seed,
                        uint32_t gen, uint32_t salt, cudaStream_t);
int launch_crossover_shards(gcp_ctx*, const gcp_shards* parents, int gene_bytes, const int32_t* strong,
                            uint32_t P, uint32_t pair_begin, uint32_t n_pairs, uint64_t seed, uint32_t gen,
                            void* children, cudaStream_t);
int launch_gather_shards(gcp_ctx*, const int32_t* order, uint32_t P, uint32_t k_begin, uint32_t n_local,
                         const gcp_shards* parents, const gcp_shards* children, int gene_bytes,
                         const double* glad_obj, const double* glad_fit, void* out_dna, double* out_obj,
                         double* out_fit, cudaStream_t);
int launch_peer_barrier(gcp_ctx*, const gcp_shards* flags, uint32_t rank, uint32_t epoch, uint32_t* timed_out,
                        cudaStream_t);
int launch_random_walks(gcp_ctx*, int32_t* dna, uint32_t P, const int32_t* start_idx_dev,
                        int path_len, uint64_t seed, uint32_t batch, cudaStream_t);
int launch_gather_survivors(gcp_ctx*, const int32_t* order, uint32_t P, const int32_t* par_dna,
                            const int32_t* kid_dna, const double* glad_obj, const double* glad_fit,
                            int32_t* out_dna, double* out_obj, double* out_fit, cudaStream_t);
