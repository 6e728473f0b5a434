/*
 * gcp.h - C ABI of the B200 fitness engine for genetic_coverage_planning.
 *
 * The reference has no FFI layer: its seam is the Python class API
 * (geneticalgorithm.py / mappy.py).  Each entry point below names the reference
 * function(s) it replaces (file:line into /root/reference/src); the Python facade in
 * genetic_coverage_planning_b200/ binds them with ctypes.
 *
 * Conventions
 *   - C linkage, POD arguments, no C++/torch types.
 *   - every function returns 0 on success, <0 on error; gcp_last_error(ctx) gives the
 *     text (owned by ctx, valid until the next call on that ctx).
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g.
 *     torch.Tensor.data_ptr()); "host" pointers are plain host memory.
 *   - work is enqueued on the caller's stream (void* = cudaStream_t, NULL = default
 *     stream); no hidden synchronisation except in the *_host convenience calls and
 *     the gcp_upload_* calls, which are synchronous.
 *   - the static lookup tables (uploaded once) live in device memory owned by ctx.
 *   - one host thread per ctx at a time; one ctx per GPU.
 *   - there is NO CPU fallback: without a CUDA device gcp_create fails.
 */
#ifndef GCP_H
#define GCP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCP_ABI_VERSION 2
#define GCP_MAX_AGENTS 32
#define GCP_INVALID_EDGE 0xFFFFFFFFu
#define GCP_MAX_SHARDS 16

typedef struct gcp_ctx gcp_ctx;

/* Scalar parameters of the path.  Sources: pather_driver.py:161-200 dicts. */
typedef struct gcp_params {
    int32_t num_agents;        /* A  <= GCP_MAX_AGENTS         gen_params['num_agents']   */
    int32_t max_dna_len;       /* L                             org_params['max_dna_len']  */
    int32_t solo_sep_thresh;   /* mappy.py:413                  map_params['solo_sep_thresh'] */
    int32_t min_solo_lcs;      /* geneticalgorithm.py:198       org_params['min_solo_lcs'] */
    int32_t min_comb_lcs;      /* geneticalgorithm.py:199       org_params['min_comb_lcs'] */
    int32_t reserved0;
    double  flight_time_scale; /* geneticalgorithm.py:196       org_params['flight_time_scale'] */
    double  coverage_blend;    /* mappy.py:364                  map_params['coverage_blend'] */
    double  one_minus_blend;   /* (1 - blend) as the host language evaluates it */
} gcp_params;

/* Variation parameters.  Sources: org_params / path_params of pather_driver.py:161-191. */
typedef struct gcp_vary_params {
    int32_t min_dna_len;              /* org_params['min_dna_len']              geneticalgorithm.py:205 */
    int32_t crossover_time_thresh;    /* org_params['crossover_time_thresh']    :330 */
    int32_t num_muterpolations;       /* org_params['num_muterpolations']       :275 (<= 64) */
    int32_t muterpolation_srch_dist;  /* org_params['muterpolation_srch_dist']  :283 */
    int32_t path_memory;              /* path_params['path_memory']             pathmaker.py:199 (<= 16) */
    int32_t reserved0;
    double  crossover_prob;           /* :217 */
    double  mutation_prob;            /* :160 */
    double  muterpolate_prob;         /* :164 */
    double  muterpolation_sub_prob;   /* :285 */
    double  log_scale_weight;         /* path_params['log_scale_weight']        pathmaker.py:182 */
} gcp_vary_params;

/* Constants of the map planes (computed by the host packer with the reference's own
 * numpy/cv2 expressions: mappy.py:25,36-39,337). */
typedef struct gcp_map_consts {
    int32_t  height, width;    /* pixels */
    int32_t  nbx, nby;         /* 64x64 blocks per row / column */
    int64_t  mask_size;        /* draw_map[buffer_mask].size          mappy.py:361 */
    int64_t  gsum_mask;        /* sum over buffer_mask of gray level g (0..255) */
    int64_t  map_size;         /* draw_map.size                        mappy.py:358 */
    double   num_occluded;     /* np.sum(img)                          mappy.py:25  */
} gcp_map_consts;

/* Per-organism outputs of gcp_eval; any pointer may be NULL (that output is skipped). */
typedef struct gcp_eval_out {
    double*   obj;        /* dev [P][2]   Organism.calcObj -> [-coverage | 0, travel_cost]  geneticalgorithm.py:177-201 */
    double*   neg_cov;    /* dev [P]      -coverage before the feasibility zeroing          mappy.py:364-370 */
    double*   dist;       /* dev [P][A]   travel_cost per agent                             mappy.py:351 */
    double*   turn;       /* dev [P][A]   turning_cost per agent                            mappy.py:352 */
    uint32_t* cov;        /* dev [P][8]   {n_mask0, n_free0, wg_mask, wg_all, n_cov_mask, n_cov_all, 0, 0} pixel counts */
    int32_t*  lc_mat;     /* dev [P][A][A] Mappy.getLoopClosures                            mappy.py:386-437 */
    uint8_t*  flags;      /* dev [P][4]   {feasible, traversable, n_components, error}      geneticalgorithm.py:188-200 */
} gcp_eval_out;

/* ---- life cycle ---------------------------------------------------------------- */
int         gcp_abi_version(void);
int         gcp_create(gcp_ctx** out, int device);
void        gcp_destroy(gcp_ctx* ctx);
const char* gcp_last_error(gcp_ctx* ctx);   /* ctx may be NULL: error of a failed gcp_create */

/* ---- one-time table upload (synchronous; host pointers) ---------------------------
 * Replaces: PathMaker._graph/_XY/_crossover_cand (pathmaker.py:157-173) and
 * Mappy._traverse_frustum/_angles/_dists (mappy.py:274-328) as seen by the hot path. */
int gcp_upload_graph(gcp_ctx* ctx, uint32_t n_nodes, uint32_t n_edges,
                     const uint32_t* row_ptr /*[N+1]*/, const uint32_t* col /*[E]*/,
                     const int32_t* xy /*[N][2] (row,col)*/, const uint8_t* xover_cand /*[N]*/);
int gcp_upload_edges(gcp_ctx* ctx,
                     const double* edge_cost   /*[E+N][2] {dist, turn}; ids >= E are null edges*/,
                     const double* edge_theta  /*[E] heading, radians*/,
                     const uint32_t* sub_ptr   /*[E+N+1]*/,
                     const uint32_t* sub_block /*[S] block id of each sub-tile*/,
                     const uint32_t* sub_payload /*[S] (offset128<<4)|qmask*/,
                     uint64_t n_sub,
                     const uint64_t* tile_data /*[n_quarters][16] rows of 64 px*/,
                     uint64_t n_quarters);
/* Optional second footprint store: every sub-tile ANDed with buffer_mask (mappy.py:337).
 * Union and AND commute, so popcount(union of masked tiles) == the reference's
 * wall-coverage pixel count (mappy.py:361).  Enables the objectives-only fast kernel
 * (binary map, coverage_blend == 1).  Same layout as the arguments of gcp_upload_edges;
 * offsets must be < 2^27. */
int gcp_upload_masked_footprints(gcp_ctx* ctx, const uint32_t* sub_ptr /*[E+N+1]*/,
                                 const uint32_t* sub_block, const uint32_t* sub_payload,
                                 uint64_t n_sub, const uint64_t* tile_data, uint64_t n_quarters);
int gcp_upload_map(gcp_ctx* ctx, const gcp_map_consts* consts,
                   const uint64_t* blk_mask0 /*[nb][64]*/, const uint64_t* blk_free0 /*[nb][64]*/,
                   const uint32_t* gray_ptr /*[nb+1]*/, const uint32_t* gray_ent /*[G]*/,
                   uint64_t n_gray);
int gcp_set_params(gcp_ctx* ctx, const gcp_params* p);

/* ---- evaluation: Organism.calcObj for P organisms ---------------------------------
 * Replaces: `for o in organisms: o.calcObj()` = Mappy.getCoverage (mappy.py:330-370)
 * + Mappy.getLoopClosures (mappy.py:386-437) + constraints (geneticalgorithm.py:177-201).
 * dna: dev int32 [P][A][L], -1 padded (Organism._dna, geneticalgorithm.py:335-344).
 *
 * Dispatch: with the pre-masked footprints uploaded, coverage_blend == 1 and out->cov == NULL the
 * whole evaluation is ONE kernel (k_eval_fused: binary and gray maps, genomes up to ~7,000 genes).
 * Any other request - coverage_blend < 1 (mappy.py:364), the six pixel counters of out->cov - runs the
 * GENERAL variant of the same kernel over the unmasked footprints (genomes up to ~6,000 genes).
 * Longer genomes, and every request under gcp_set_option("fused_eval", 0), run the staged kernels
 * (loop closures switch to a hash-partitioned / radix-sort formulation when A*L outgrows shared
 * memory).  scratch_dev must hold gcp_eval_scratch_bytes(n_org) bytes whenever the staged kernels
 * may run (the fused kernels do not touch it).  Results are identical on every route. */
size_t gcp_eval_scratch_bytes(gcp_ctx* ctx, uint32_t n_org);
int    gcp_eval(gcp_ctx* ctx, const int32_t* dna_dev, uint32_t n_org,
                const gcp_eval_out* out, void* scratch_dev, size_t scratch_bytes,
                void* stream);
/* Same evaluation from int16 genes (dev int16 [P][A][L], -1 padded): halves the chromosome
 * bytes moved over PCIe / HBM.  Needs N <= 32767 waypoints and the fused objectives-only path
 * (pre-masked footprints uploaded, coverage_blend == 1, out->cov == NULL). */
int    gcp_eval16(gcp_ctx* ctx, const int16_t* dna_dev, uint32_t n_org, const gcp_eval_out* out,
                  void* stream);
/* 1 if gcp_eval16 can run with the tables / parameters uploaded so far, else 0. */
int    gcp_eval16_available(gcp_ctx* ctx);
/* 1 if gcp_eval_host would narrow its int32 genes to int16 on the host with the current tables,
 * parameters and options (see "host_narrow" below), else 0. */
int    gcp_host_narrow_active(gcp_ctx* ctx);
/* Individual stages (same buffers; used by tests / profiling). edge_ids: dev u32 [P][A][L];
 * step_info (optional, needs the masked store): dev u32 [P][A][L], per step the masked
 * sub-tile range of its edge (first | count<<28), 0 where no step is executed. */
int gcp_path_costs(gcp_ctx* ctx, const int32_t* dna_dev, uint32_t n_org, uint32_t* edge_ids_dev,
                   uint32_t* step_info_dev, double* dist_dev, double* turn_dev,
                   uint8_t* flags_dev, void* stream);
int gcp_coverage(gcp_ctx* ctx, const uint32_t* edge_ids_dev, uint32_t n_org,
                 uint32_t* cov_dev, uint8_t* flags_dev, void* stream);
/* objectives-only wall coverage from the masked store: cov[P][8] = {n,0,0,0,n,0,0,0}. */
int gcp_coverage_masked(gcp_ctx* ctx, const uint32_t* step_info_dev, uint32_t n_org,
                        uint32_t* cov_dev, uint8_t* flags_dev, void* stream);
int gcp_loop_closures(gcp_ctx* ctx, const int32_t* dna_dev, uint32_t n_org,
                      const double* dist_dev, const double* turn_dev, const uint32_t* cov_dev,
                      int32_t* lc_mat_dev, double* obj_dev, double* neg_cov_dev,
                      uint8_t* flags_dev, void* stream);

/* Host-buffer convenience (the end-to-end path): copies dna H2D in chunks, evaluates,
 * copies obj [P][2] and flags [P][4] back; synchronous.  Pinned host memory is used as
 * is; pageable memory works but is slower. */
int gcp_eval_host(gcp_ctx* ctx, const int32_t* dna_host, uint32_t n_org,
                  double* obj_host, uint8_t* flags_host);
int gcp_eval_host16(gcp_ctx* ctx, const int16_t* dna_host, uint32_t n_org,
                    double* obj_host, uint8_t* flags_host);

/* ---- ranking: GeneticalGorithm.rackNstack + arena ordering -------------------------
 * Replaces geneticalgorithm.py:86-123 (maximin fitness) and gori_tools.py:232-236
 * (argsort; canonical stable order by (fit, index)).
 * coverage_constr = (1-alpha)*c0 + alpha*cf evaluated by the host (geneticalgorithm.py:98-100).
 * Rows [row_begin, row_end) of fit are computed (multi-GPU row sharding); pass 0, n.
 * obj: dev [n][2]; obj_sc (optional): dev [n][2]; fit: dev [n]. */
size_t gcp_rank_scratch_bytes(gcp_ctx* ctx, uint32_t n);
int gcp_maximin(gcp_ctx* ctx, const double* obj_dev, uint32_t n, double coverage_constr,
                uint32_t row_begin, uint32_t row_end, double* obj_sc_dev, double* fit_dev,
                void* scratch_dev, size_t scratch_bytes, void* stream);
/* order[k] = index of the k-th best (ascending fit, ties by index); nondom[i] = fit[i] < 0. */
int gcp_arena_order(gcp_ctx* ctx, const double* fit_dev, uint32_t n, int32_t* order_dev,
                    uint8_t* nondom_dev, void* scratch_dev, size_t scratch_bytes, void* stream);

/* ---- selection + variation (Philox streams keyed by seed / generation / organism / agent)
 * Replaces got.lowVarSample (gori_tools.py:207-229), Organism.crossover + matchWaypt +
 * padMinDNA (geneticalgorithm.py:203-249,311-333), the constructor's mutation / muterpolate /
 * pruneUTurns / addTelomere (:156-175,251-357), PathMaker.makeMeAPath (pathmaker.py:188-213)
 * and the arena truncation (geneticalgorithm.py:78-84). */
int gcp_set_vary_params(gcp_ctx* ctx, const gcp_vary_params* p);
/* out_idx[m] = index of the m-th resampled parent; r0 = the single U(0, 1/M) draw;
 * scratch: dev double [M + 2 * ceil(M / 1024)]. */
int gcp_low_var_select(gcp_ctx* ctx, const double* fit_dev, uint32_t m, double gamma, double r0,
                       double* scratch_dev, int32_t* out_idx_dev, void* stream);
/* the r0 that gcp_select_vary derives from (seed, generation) for a population of m. */
double gcp_low_var_r0(uint64_t seed, uint32_t generation, uint32_t m);
/* pair k = (strong[k], strong[k + P/2]) -> children 2k, 2k+1.  dna buffers: dev int32 [P][A][L]. */
int gcp_crossover(gcp_ctx* ctx, const int32_t* parents_dna_dev, const int32_t* strong_idx_dev,
                  uint32_t n_org, uint64_t seed, uint32_t generation, int32_t* children_dna_dev,
                  void* stream);
/* in place: mutation (prob), muterpolate (prob), pruneUTurns, addTelomere; `salt` separates
 * independent uses within one generation.  Rows must be a valid prefix followed by -1 padding
 * (what every producer in this API and Organism.addTelomere emit). */
int gcp_mutate(gcp_ctx* ctx, int32_t* dna_dev, uint32_t n_org, uint64_t seed, uint32_t generation,
               uint32_t salt, void* stream);
/* select + crossover + mutate in one call (the reference's child construction loop,
 * geneticalgorithm.py:59-68).  scratch: dev, >= 12 * n_org + 16 * ceil(n_org / 1024) bytes
 * (n_org * 16 + 64 always suffices). */
int gcp_select_vary(gcp_ctx* ctx, const int32_t* parents_dna_dev, const double* parent_fit_dev,
                    uint32_t n_org, double gamma, uint64_t seed, uint32_t generation,
                    int32_t* children_dna_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
/* heading-biased random walks of path_len steps from start_idx[a] (firstGeneration,
 * geneticalgorithm.py:44-48).  start_idx_dev: dev int32 [A]. */
int gcp_random_walks(gcp_ctx* ctx, int32_t* dna_dev, uint32_t n_org, const int32_t* start_idx_dev,
                     int32_t path_len, uint64_t seed, uint32_t batch, void* stream);
/* survivors k < P of the arena: out[k] = gladiator[order[k]], gladiators = parents ++ children;
 * glad_obj [2P][2], glad_fit [2P]. */
int gcp_gather_survivors(gcp_ctx* ctx, const int32_t* order_dev, uint32_t n_org,
                         const int32_t* parents_dna_dev, const int32_t* children_dna_dev,
                         const double* glad_obj_dev, const double* glad_fit_dev,
                         int32_t* out_dna_dev, double* out_obj_dev, double* out_fit_dev, void* stream);

/* ---- population shards over the GPUs of one box ------------------------------
 * The reference keeps ONE population (geneticalgorithm.py:28-32,55-84).  Here rank r of `world`
 * holds parents [r*per, (r+1)*per) and builds / evaluates children of the same range; objective
 * and fitness records (24 B / organism) are replicated; chromosomes never leave the GPU that
 * produced them except when another rank needs them as a crossover parent or as a survivor of
 * its own range - and then the kernel reads them from the owner's memory over NVLink (peer
 * memory mapped with CUDA IPC).  There is no chromosome all-gather.
 *
 * gcp_shards describes one logical [world * per_shard][A][L] gene array: base[w] is shard w as
 * mapped into THIS process (own shard: the pointer gcp_shard_alloc returned, peers: gcp_shard_open).
 * world == 1 is a plain array.  Genes are int16 (gene_bytes 2, needs N <= 32767 waypoints) or
 * int32 (gene_bytes 4), -1 padded. */
typedef struct gcp_shards {
    uint32_t    world;                 /* number of shards, <= GCP_MAX_SHARDS */
    uint32_t    per_shard;             /* organisms per shard */
    const void* base[GCP_MAX_SHARDS];  /* dev */
} gcp_shards;
/* device memory that other processes of the box can map: cudaMalloc + cudaIpcGetMemHandle.
 * ipc_handle: host, 64 bytes (cudaIpcMemHandle_t), to be sent to the peers by the caller
 * (e.g. torch.distributed all-gather).  Freed by gcp_shard_free or gcp_destroy. */
int gcp_shard_alloc(gcp_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t ipc_handle[64]);
int gcp_shard_free(gcp_ctx* ctx, void* dev_ptr);
/* map a peer's shard (same or another GPU of the box; peer access is enabled on demand). */
int gcp_shard_open(gcp_ctx* ctx, const uint8_t ipc_handle[64], void** dev_ptr);
int gcp_shard_close(gcp_ctx* ctx, void* dev_ptr);
/* Organism.crossover for pairs [pair_begin, pair_begin + n_pairs) of the n_org_global / 2 pairs
 * (geneticalgorithm.py:64-68): pair k = (strong[k], strong[k + n_org_global / 2]), parents read
 * through the shard table, children 2 (k - pair_begin), +1 written to children_dev
 * [2 * n_pairs][A][L].  Philox streams are keyed by the GLOBAL pair id: any sharding gives the
 * children a single GPU would build.  strong_idx_dev: dev int32 [n_org_global]. */
int gcp_crossover_shards(gcp_ctx* ctx, const gcp_shards* parents, int gene_bytes,
                         const int32_t* strong_idx_dev, uint32_t n_org_global, uint32_t pair_begin,
                         uint32_t n_pairs, uint64_t seed, uint32_t generation, void* children_dev,
                         void* stream);
/* gcp_mutate on int16 or int32 genes; org_offset = global id of local organism 0 (Philox keys). */
int gcp_mutate_genes(gcp_ctx* ctx, void* dna_dev, int gene_bytes, uint32_t n_org, uint32_t org_offset,
                     uint64_t seed, uint32_t generation, uint32_t salt, void* stream);
/* arena truncation (geneticalgorithm.py:83-84) for survivors [k_begin, k_begin + n_local): out_dna
 * [n_local][A][L] = gladiator[order[k]], gladiators = parents ++ children, both sharded (a peer's
 * organism is read over NVLink).  out_obj [n_org_global][2] / out_fit [n_org_global] (optional, both
 * or neither) receive the records of ALL survivors. */
int gcp_gather_shards(gcp_ctx* ctx, const int32_t* order_dev, uint32_t n_org_global, uint32_t k_begin,
                      uint32_t n_local, const gcp_shards* parents, const gcp_shards* children,
                      int gene_bytes, const double* glad_obj_dev, const double* glad_fit_dev,
                      void* out_dna_dev, double* out_obj_dev, double* out_fit_dev, void* stream);
/* Rendezvous of the ranks on the device, without a host round trip or a collective library: every
 * rank owns `world` u32 flags in shard memory; the kernel stores `epoch` into slot `rank` of every
 * peer's flags (system-scope release) and waits until all of its own slots hold >= epoch.  Work
 * enqueued after it on this stream sees everything the peers wrote before their barrier.  A peer
 * that does not arrive within ~2 s makes the kernel give up and set *timed_out_dev (dev u32). */
int gcp_peer_barrier(gcp_ctx* ctx, const gcp_shards* flags, uint32_t rank, uint32_t epoch,
                     uint32_t* timed_out_dev, void* stream);

/* ---- tuning / testing knobs (defaults are right for production) ---------------------
 * "rank_algo": 0 auto, 1 O(N^2) dominance-matrix tiles, 2 sort-based O(N log N) (both exact);
 * "host_narrow": gcp_eval_host narrows int32 genes to int16 with host threads on the way to the GPU
 * when the waypoint ids fit (N <= 32767) - half the PCIe bytes, same results: 1 always, 0 never, -1
 * (default) when it pays, i.e. one rank on this host (LOCAL_WORLD_SIZE) and at least six worker
 * threads (gcp_host_narrow_active tells); "host_threads": worker threads of that conversion
 * (0 = auto: cores / LOCAL_WORLD_SIZE - 1, at most 12);
 * "sort_onesweep": 1 (default) radix sorts of <= 2^20 pairs take one launch per pass (decoupled
 * look-back), 0 = histogram / scan / scatter launches per pass (same order, the cross-check);
 * "force_general": never take the pre-masked coverage fast path; "fused_eval": 0 = staged kernels;
 * "force_lc_big": long-genome loop closures even for short genomes (1 radix sort, 2 partitioned hash); "vary_org_offset": global id of
 * local organism 0 when a rank varies only its shard of the children (keys of the Philox streams);
 * "cov_table", "cov_maxpairs", "cov_threads": shared-memory sizing / CTA size of the coverage kernels
 * (0 = auto); "cov_dedup": 0 = keep revisited edges in the coverage union (same result, more work);
 * "cov_buckets": 0 = long genomes re-scan their items per hash partition instead of bucketing them
 * once; "cov_split": 0 = the hash partitions of a long genome run one after the other in the organism's
 * CTA instead of as work items of small CTAs ("lc_split": the same switch for the loop closures;
 * "cov_split_cap" (items aimed at per partition, 0 = 2,600) / "cov_split_k" (partitions per organism fixed
 * instead of chosen per organism) / "lc_split_keys" (genes aimed at per loop-closure partition, 0 = 640) /
 * "lc_split_slack": tuning and testing - e.g. force the overflow fallback of the split kernels); "lc_min_parts": at least this many hash partitions in the long-genome loop-closure kernel;
 * "cov_prefetch": -1 auto (2 when the footprint store is larger than the L2, else 0) / bit 0: L2 prefetch
 * of the footprint lines while the fused kernel sorts its items / bit 1: the union stream requests its
 * lines 16 entries ahead (the one that helps on config C5).
 * Unknown names return an error. */
int gcp_set_option(gcp_ctx* ctx, const char* name, int64_t value);

/* ---- introspection ---------------------------------------------------------------- */
/* kernel launch counter since gcp_create (for bench.py's gpu_launches). */
uint64_t gcp_launch_count(gcp_ctx* ctx);
/* time of the most recent call of each stage measured with CUDA events on the launch
 * stream; only recorded while enabled (adds two event records per stage). */
int gcp_set_timing(gcp_ctx* ctx, int enabled);
/* stage: 0 path_costs, 1 coverage (or the whole fused evaluation), 2 loop_closures, 3 maximin,
 * 4 arena_order, 5 select, 6 crossover, 7 mutate (walks + muterpolate + prune).
 * Synchronises on the stage's end event. Returns milliseconds, <0 if not recorded. */
float gcp_stage_ms(gcp_ctx* ctx, int stage);

/* ---- setup path: footprint polygons ------------------------------------
 * Replaces Mappy.computeFrustums (mappy.py:274-328): the (num_rays + 2)-point footprint polygon
 * of every directed edge, from the obstacle pixels within max_view of the source waypoint.
 * Needs no context (it runs before any table exists); errors are read with gcp_last_error(NULL).
 * All pointers are HOST pointers; the call is synchronous.
 *   occ[H][W]            != 0 where Mappy._img != 0 (mappy.py:278)
 *   xy[N][2]             waypoint (row, col) = Mappy._all_waypoints
 *   row_ptr[N+1], col[E] CSR of the traversability graph (graph[from, to] == 1)
 *   ray_angles[R]        np.arange(R) * (view_angle / R) - view_angle / 2          (:279-280)
 *   ray_sin/ray_cos[R]   np.sin / np.cos of ray_angles                             (:314)
 *   edge_theta[E]        np.arctan2(d_col, d_row) of the edge                      (:286)
 *   edge_cos/edge_sin[E] np.cos / np.sin of edge_theta (got.getRot2D, gori_tools.py:184-186)
 *   poly[E][R+2][2]      out: np.around(Rot.dot(pts)).astype(int32).T              (:320)
 *   ambiguous[E]         out: 1 where a floating-point decision fell within eps of a boundary
 *                        (bearing vs ray boundary / field of view, coordinate vs x.5): the
 *                        caller recomputes those edges with the reference's own expressions
 *   kernel_ms            out (optional): device time of the kernel                          */
typedef struct gcp_frustum_params {
    int32_t num_rays;         /* R <= 62 */
    int32_t window;           /* half-size in px of the obstacle window: ceil(max_view / scale) + 1 */
    double  scale;            /* m per px */
    double  max_view;         /* m */
    double  max_view_px;      /* max_view / scale */
    double  half_view;        /* view_angle / 2 (< pi) */
    double  near_last[2];     /* {sin(ray_angles[R-1]) * min_view, cos(ray_angles[R-1]) * min_view}  (:315-316) */
    double  near_first[2];    /* {sin(ray_angles[0]) * min_view, cos(ray_angles[0]) * min_view}      (:317-318) */
    double  eps_angle;        /* e.g. 1e-11 rad */
    double  eps_round;        /* e.g. 1e-9 px */
} gcp_frustum_params;
int gcp_compute_frustums(int device, const uint8_t* occ, uint32_t H, uint32_t W, const int32_t* xy,
                         const uint32_t* row_ptr, const uint32_t* col, uint32_t N, uint32_t E,
                         const gcp_frustum_params* fp, const double* ray_angles, const double* ray_sin,
                         const double* ray_cos, const double* edge_theta, const double* edge_cos,
                         const double* edge_sin, int32_t* poly, uint8_t* ambiguous, float* kernel_ms);


/* ---- setup path: traversability graph -----------------------------------
 * Replaces PathMaker.computeTraversableGraph (pathmaker.py:119-155) and the per-edge
 * Mappy.lineCollisionCheck (mappy.py:124-159).  No context; HOST pointers; synchronous; errors
 * through gcp_last_error(NULL).
 *   occ[H][W]              != 0 where Mappy._img != 0
 *   xy[N][2]               waypoint (row, col) = PathMaker._XY
 *   cell_ptr / cell_items  waypoints bucketed into square grid cells of `cell` px, row-major cells
 *                          ((H + cell - 1) / cell rows of (W + cell - 1) / cell), ascending waypoint
 *                          index inside a cell; cell >= max_dist / scale
 *   nbr[N][8]              out: per 45-degree sector the surviving neighbour, -1 if none
 *                          (graph[i, nbr[i][s]] = 1)
 *   ambiguous[N]           out: 1 where a bearing fell within eps_angle of a sector boundary: the
 *                          caller recomputes those rows with the reference's own expressions   */
typedef struct gcp_graph_params {
    double  scale;            /* m per px (Mappy._scale) */
    double  max_dist;         /* path_params['max_traverse_dist'], m (pathmaker.py:147) */
    int32_t num_dilations;    /* int((hall_width / 2 / 3) / scale)  (pathmaker.py:153, mappy.py:136) */
    int32_t cell;             /* grid cell size, px */
    double  sectors[8];       /* (pi / 4) * [-7/2, -5/2, ..., 7/2]  (pathmaker.py:125-127) */
    double  eps_angle;        /* e.g. 1e-11 rad */
} gcp_graph_params;
int gcp_build_graph(int device, const uint8_t* occ, uint32_t H, uint32_t W, const int32_t* xy, uint32_t N,
                    const uint32_t* cell_ptr, const uint32_t* cell_items, const gcp_graph_params* gp,
                    int32_t* nbr, uint8_t* ambiguous, float* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* GCP_H */
