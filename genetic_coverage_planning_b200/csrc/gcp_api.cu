// C ABI (include/gcp.h): context, table upload, stage dispatch.  Host code only.
#include "gcp_common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

size_t rank_scratch_bytes(gcp_ctx* ctx, uint32_t n);

static std::string g_create_err;

int gcp_fail(gcp_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

extern "C" {

int gcp_abi_version(void) { return GCP_ABI_VERSION; }

const char* gcp_last_error(gcp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int gcp_create(gcp_ctx** out, int device)
{
    if (!out) return gcp_fail(nullptr, -1, "gcp_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return gcp_fail(nullptr, -2, "gcp_create: no CUDA device (%s); this engine has no CPU path",
                        cudaGetErrorString(e));
    if (device < 0 || device >= n) return gcp_fail(nullptr, -1, "gcp_create: bad device %d", device);
    gcp_ctx* c = new (std::nothrow) gcp_ctx();
    if (!c) return gcp_fail(nullptr, -4, "out of host memory");
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        delete c;
        return gcp_fail(nullptr, -2, "cudaSetDevice: %s", cudaGetErrorString(e));
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (int s = 0; s < GCP_NUM_STAGES; ++s) {
        cudaEventCreate(&c->ev_beg[s]);
        cudaEventCreate(&c->ev_end[s]);
    }
    *out = c;
    return 0;
}

void gcp_destroy(gcp_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (void* q : c->shard_peer) cudaIpcCloseMemHandle(q);
    for (void* q : c->shard_own) cudaFree(q);
    for (int i = 0; i < c->n_owned; ++i) cudaFree(c->owned[i]);
    if (c->stage_dev) cudaFree(c->stage_dev);
    gcp_host_free(c);
    if (c->vary_scratch) cudaFree(c->vary_scratch);
    for (int s = 0; s < 3; ++s) {
        if (c->lcb_scratch_s[s]) cudaFree(c->lcb_scratch_s[s]);
        if (c->cov_scratch_s[s]) cudaFree(c->cov_scratch_s[s]);
    }
    for (int s = 0; s < 3; ++s) if (c->hstream[s]) cudaStreamDestroy(c->hstream[s]);
    for (int s = 0; s < GCP_NUM_STAGES; ++s) { cudaEventDestroy(c->ev_beg[s]); cudaEventDestroy(c->ev_end[s]); }
    delete c;
}

} // extern "C"

template <typename T>
static int upload(gcp_ctx* c, const T** dst, const void* src, size_t count)
{
    void* p = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    GCP_CUDA(c, cudaMalloc(&p, bytes));
    if (c->n_owned >= 32) return gcp_fail(c, -4, "too many table uploads (upload each table once)");
    c->owned[c->n_owned++] = p;
    if (count) GCP_CUDA(c, cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *dst = reinterpret_cast<const T*>(p);
    return 0;
}

// Flat per-edge lists of QUARTER items {block << 2 | q, index} for the fused kernels (index: quarter
// number << idx_shift; 3 = uint4 units).  false if an edge has more than 63 items or the lists
// outgrow 26 bits.
static bool build_quarter_lists(const uint32_t* sub_ptr, const uint32_t* sub_block, const uint32_t* sub_payload,
                                size_t EN, uint64_t n_sub, int idx_shift, std::vector<uint32_t>& qinfo,
                                std::vector<uint2>& qdesc)
{
    qinfo.assign(EN, 0u);
    qdesc.clear();
    qdesc.reserve((size_t)n_sub * 2);
    for (size_t e = 0; e < EN; ++e) {
        if (qdesc.size() & 1) qdesc.push_back(make_uint2(0u, 0u));   // lists start 16-B aligned: items are read in pairs
        const size_t first = qdesc.size();
        for (uint32_t k = sub_ptr[e]; k < sub_ptr[e + 1]; ++k) {
            uint32_t off = sub_payload[k] >> 4;
            for (uint32_t q = 0; q < 4; ++q)
                if ((sub_payload[k] >> q) & 1u) qdesc.push_back(make_uint2((sub_block[k] << 2) | q, (off++) << idx_shift));
        }
        const size_t n = qdesc.size() - first;
        if (n > 63 || first >= (1u << 26)) return false;
        qinfo[e] = (uint32_t)first | ((uint32_t)n << 26);
    }
    qdesc.push_back(make_uint2(0u, 0u));                             // pair reads may touch one item past a list
    qdesc.push_back(make_uint2(0u, 0u));
    qdesc.push_back(make_uint2(0u, 0u));
    return true;
}

// General fused kernel: footprint quarters, (g == 0) plane and mask & (g == 0) plane in ONE allocation
// (the map planes are 32 x uint4 = four 128-B quarters per block, like the tiles), so that the kernel's
// sorted stream addresses all three with one 30-bit quarter index.  Replaces the tile allocation.
static int build_general_store(gcp_ctx* c)
{
    c->t.g_free_q = 0; c->t.g_mask_q = 0;
    if (!c->have_edges || !c->have_map || !c->t.q_info) return 0;
    const size_t nq = c->n_quarters, nb = c->t.n_blocks;
    if (nq == 0 || nq + 8 * nb >= (1ull << 30)) return 0;            // the staged kernels remain
    char* p = nullptr;
    GCP_CUDA(c, cudaMalloc(&p, (nq + 8 * nb) * 128));
    GCP_CUDA(c, cudaMemcpy(p, c->t.tile, nq * 128, cudaMemcpyDeviceToDevice));
    GCP_CUDA(c, cudaMemcpy(p + nq * 128, c->t.blk_free0, nb * 512, cudaMemcpyDeviceToDevice));
    GCP_CUDA(c, cudaMemcpy(p + (nq + 4 * nb) * 128, c->t.blk_mask0, nb * 512, cudaMemcpyDeviceToDevice));
    for (int i = 0; i < c->n_owned; ++i)
        if (c->owned[i] == (void*)c->t.tile) { cudaFree(c->owned[i]); c->owned[i] = p; }
    c->t.tile = reinterpret_cast<const uint4*>(p);
    c->t.g_free_q = (uint32_t)nq;
    c->t.g_mask_q = (uint32_t)(nq + 4 * nb);
    return 0;
}

extern "C" {

int gcp_upload_graph(gcp_ctx* c, uint32_t N, uint32_t E, const uint32_t* row_ptr,
                     const uint32_t* col, const int32_t* xy, const uint8_t* xover)
{
    if (!c) return -1;
    if (c->have_graph) return gcp_fail(c, -1, "graph already uploaded (create a new ctx)");
    if (N == 0 || N >= (1u << 24)) return gcp_fail(c, -1, "need 0 < N < 2^24 waypoints, got %u", N);
    if (row_ptr[N] != E) return gcp_fail(c, -1, "row_ptr[N]=%u != E=%u", row_ptr[N], E);
    GCP_CUDA(c, cudaSetDevice(c->device));
    int r;
    if ((r = upload(c, &c->t.row_ptr, row_ptr, (size_t)N + 1))) return r;
    if ((r = upload(c, &c->t.col, col, E))) return r;
    if ((r = upload(c, &c->t.xy, xy, (size_t)N))) return r;           // int2 = 2 x int32
    if ((r = upload(c, &c->t.xover, xover, N))) return r;
    c->t.N = N; c->t.E = E;
    c->h_row_ptr.assign(row_ptr, row_ptr + N + 1);
    c->h_col.assign(col, col + E);
    c->max_degree = 0;
    for (uint32_t i = 0; i < N; ++i) {
        if (row_ptr[i + 1] < row_ptr[i]) return gcp_fail(c, -1, "row_ptr not monotone at %u", i);
        const uint32_t d = row_ptr[i + 1] - row_ptr[i];
        if (d > c->max_degree) c->max_degree = d;
    }
    c->have_graph = true;
    return 0;
}

int gcp_upload_edges(gcp_ctx* c, const double* edge_cost, const double* edge_theta,
                     const uint32_t* sub_ptr, const uint32_t* sub_block,
                     const uint32_t* sub_payload, uint64_t n_sub, const uint64_t* tile_data,
                     uint64_t n_quarters)
{
    if (!c) return -1;
    if (!c->have_graph) return gcp_fail(c, -1, "upload the graph first");
    if (c->have_edges) return gcp_fail(c, -1, "edges already uploaded");
    const size_t EN = (size_t)c->t.E + c->t.N;
    if (sub_ptr[EN] != n_sub) return gcp_fail(c, -1, "sub_ptr[E+N] != n_sub");
    if (n_quarters >= (1ull << 28)) return gcp_fail(c, -1, "footprint store too large");
    GCP_CUDA(c, cudaSetDevice(c->device));
    int r;
    if ((r = upload(c, &c->t.edge_cost, edge_cost, EN))) return r;     // double2
    if ((r = upload(c, &c->t.edge_theta, edge_theta, c->t.E))) return r;
    c->t.adj8 = nullptr;
    if (c->max_degree <= 8) {   // packed adjacency for the random walks: one 64-B record per node
        std::vector<uint2> adj((size_t)c->t.N * 8, make_uint2(0xFFFFFFFFu, 0u));
        for (uint32_t i = 0; i < c->t.N; ++i)
            for (uint32_t e = c->h_row_ptr[i]; e < c->h_row_ptr[i + 1]; ++e) {
                const float th = (float)edge_theta[e];
                uint32_t bits; memcpy(&bits, &th, 4);
                adj[(size_t)i * 8 + (e - c->h_row_ptr[i])] = make_uint2(c->h_col[e], bits);
            }
        if ((r = upload(c, &c->t.adj8, adj.data(), adj.size()))) return r;
    }
    c->t.adjc16 = nullptr;
    if (c->max_degree <= 8 && c->t.N <= 65535) {   // edge lookup of the evaluation: 8 x u16 columns = one 16-B load
        std::vector<uint16_t> a16((size_t)c->t.N * 8, (uint16_t)0xFFFF);
        for (uint32_t i = 0; i < c->t.N; ++i)
            for (uint32_t e = c->h_row_ptr[i]; e < c->h_row_ptr[i + 1]; ++e)
                a16[(size_t)i * 8 + (e - c->h_row_ptr[i])] = (uint16_t)c->h_col[e];
        const uint4* dev = nullptr;
        if ((r = upload(c, &dev, a16.data(), a16.size() / 8))) return r;      // uint4 = 8 x u16
        c->t.adjc16 = dev;
    }
    c->h_row_ptr.clear(); c->h_row_ptr.shrink_to_fit();
    c->h_col.clear(); c->h_col.shrink_to_fit();
    {   // device layout: one u32 per edge (first | count<<28) and one uint2 per sub-tile
        if (n_sub >= (1ull << 28)) return gcp_fail(c, -1, "too many sub-tiles");
        std::vector<uint32_t> info(EN);
        for (size_t e = 0; e < EN; ++e) {
            const uint32_t n = sub_ptr[e + 1] - sub_ptr[e];
            if (n > 15) return gcp_fail(c, -1, "edge %zu has %u sub-tiles (max 15)", e, n);
            info[e] = sub_ptr[e] | (n << 28);
        }
        std::vector<uint2> desc(n_sub);
        for (size_t k = 0; k < n_sub; ++k) desc[k] = make_uint2(sub_block[k], sub_payload[k]);
        if ((r = upload(c, &c->t.sub_info, info.data(), EN))) return r;
        if ((r = upload(c, &c->t.sub_desc, desc.data(), n_sub))) return r;
    }
    if ((r = upload(c, &c->t.tile, tile_data, (size_t)n_quarters * 8))) return r;   // uint4
    {   // quarter items of the unmasked store: the general fused kernel (blend < 1, pixel counters)
        std::vector<uint32_t> qinfo;
        std::vector<uint2> qdesc;
        c->t.q_info = nullptr; c->t.q_desc = nullptr;
        if (build_quarter_lists(sub_ptr, sub_block, sub_payload, EN, n_sub, 0, qinfo, qdesc)) {
            if ((r = upload(c, &c->t.q_info, qinfo.data(), EN))) return r;
            if ((r = upload(c, &c->t.q_desc, qdesc.data(), qdesc.size()))) return r;
        }
    }
    c->n_sub = n_sub; c->n_quarters = n_quarters;
    c->have_edges = true;
    return build_general_store(c);
}

int gcp_upload_masked_footprints(gcp_ctx* c, const uint32_t* sub_ptr, const uint32_t* sub_block,
                                 const uint32_t* sub_payload, uint64_t n_sub,
                                 const uint64_t* tile_data, uint64_t n_quarters)
{
    if (!c) return -1;
    if (!c->have_graph) return gcp_fail(c, -1, "upload the graph first");
    if (c->have_masked) return gcp_fail(c, -1, "masked footprints already uploaded");
    const size_t EN = (size_t)c->t.E + c->t.N;
    if (sub_ptr[EN] != n_sub) return gcp_fail(c, -1, "masked sub_ptr[E+N] != n_sub");
    if (n_quarters >= (1ull << 27) || n_sub >= (1ull << 28))
        return gcp_fail(c, -1, "masked footprint store too large");
    GCP_CUDA(c, cudaSetDevice(c->device));
    std::vector<uint32_t> info(EN);
    for (size_t e = 0; e < EN; ++e) {
        const uint32_t n = sub_ptr[e + 1] - sub_ptr[e];
        if (n > 15) return gcp_fail(c, -1, "edge %zu has %u masked sub-tiles (max 15)", e, n);
        info[e] = sub_ptr[e] | (n << 28);
    }
    std::vector<uint2> desc(n_sub);
    for (size_t k = 0; k < n_sub; ++k) desc[k] = make_uint2(sub_block[k], sub_payload[k]);
    int r;
    if ((r = upload(c, &c->t.m_sub_info, info.data(), EN))) return r;
    if ((r = upload(c, &c->t.m_sub_desc, desc.data(), n_sub))) return r;
    if ((r = upload(c, &c->t.m_tile, tile_data, (size_t)n_quarters * 8))) return r;
    {   // flat per-edge lists of QUARTER items for the fused kernel: {block<<2 | q, uint4 index}
        std::vector<uint32_t> qinfo;
        std::vector<uint2> qdesc;
        c->t.m_q_info = nullptr; c->t.m_q_desc = nullptr;
        if (build_quarter_lists(sub_ptr, sub_block, sub_payload, EN, n_sub, 3, qinfo, qdesc)) {
            if ((r = upload(c, &c->t.m_q_info, qinfo.data(), EN))) return r;
            if ((r = upload(c, &c->t.m_q_desc, qdesc.data(), qdesc.size()))) return r;
        }
    }
    c->m_n_quarters = n_quarters;
    c->have_masked = true;
    return 0;
}

int gcp_upload_map(gcp_ctx* c, const gcp_map_consts* mc, const uint64_t* blk_mask0,
                   const uint64_t* blk_free0, const uint32_t* gray_ptr, const uint32_t* gray_ent,
                   uint64_t n_gray)
{
    if (!c || !mc) return -1;
    if (c->have_map) return gcp_fail(c, -1, "map already uploaded");
    const size_t nb = (size_t)mc->nbx * mc->nby;
    if (nb == 0 || nb >= (1u << 28)) return gcp_fail(c, -1, "bad block grid %d x %d", mc->nbx, mc->nby);
    if (gray_ptr[nb] != n_gray) return gcp_fail(c, -1, "gray_ptr[nb] != n_gray");
    GCP_CUDA(c, cudaSetDevice(c->device));
    int r;
    if ((r = upload(c, &c->t.blk_mask0, blk_mask0, nb * 32))) return r;    // uint4
    if ((r = upload(c, &c->t.blk_free0, blk_free0, nb * 32))) return r;
    if ((r = upload(c, &c->t.gray_ptr, gray_ptr, nb + 1))) return r;
    {   // a block's gray entries ordered by quarter (stable), plus the CSR over (block, quarter): the flat
        // union of the general fused kernel applies the gray weights at the end of every quarter list
        std::vector<uint32_t> ent(gray_ent, gray_ent + n_gray), qptr(nb * 4 + 1, 0u);
        for (size_t b = 0; b < nb; ++b) {
            uint32_t cnt[4] = {0, 0, 0, 0};
            for (uint32_t g = gray_ptr[b]; g < gray_ptr[b + 1]; ++g) cnt[((gray_ent[g] & 4095u) >> 6) >> 4]++;
            uint32_t at[4], run = gray_ptr[b];
            for (int q = 0; q < 4; ++q) { qptr[b * 4 + q] = run; at[q] = run; run += cnt[q]; }
            for (uint32_t g = gray_ptr[b]; g < gray_ptr[b + 1]; ++g) ent[at[((gray_ent[g] & 4095u) >> 6) >> 4]++] = gray_ent[g];
        }
        qptr[nb * 4] = (uint32_t)n_gray;
        if ((r = upload(c, &c->t.gray_qptr, qptr.data(), qptr.size()))) return r;
        if ((r = upload(c, &c->t.gray_ent, ent.data(), n_gray))) return r;
    }
    c->t.n_blocks = (uint32_t)nb;
    c->mc = *mc;
    c->n_gray = n_gray;
    c->have_map = true;
    return build_general_store(c);
}

int gcp_set_params(gcp_ctx* c, const gcp_params* p)
{
    if (!c || !p) return -1;
    if (p->num_agents < 1 || p->num_agents > GCP_MAX_AGENTS)
        return gcp_fail(c, -1, "num_agents must be in [1,%d]", GCP_MAX_AGENTS);
    if (p->max_dna_len < 2) return gcp_fail(c, -1, "max_dna_len must be >= 2");
    if ((int64_t)p->num_agents * p->max_dna_len > 65535)
        return gcp_fail(c, -1, "num_agents*max_dna_len must be <= 65535");
    c->p = *p;
    c->have_params = true;
    return 0;
}

static int ready(gcp_ctx* c)
{
    if (!c) return -1;
    if (!(c->have_graph && c->have_edges && c->have_map && c->have_params))
        return gcp_fail(c, -1, "tables/params not uploaded (graph=%d edges=%d map=%d params=%d)",
                        c->have_graph, c->have_edges, c->have_map, c->have_params);
    return 0;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t gcp_eval_scratch_bytes(gcp_ctx* c, uint32_t P)
{
    if (!c || !c->have_params) return 0;
    const size_t A = c->p.num_agents, L = c->p.max_dna_len;
    return 2 * align256(P * A * L * 4) + 2 * align256(P * A * 8) + align256((size_t)P * 32) +
           align256((size_t)P * 4) + 256;
}

int gcp_path_costs(gcp_ctx* c, const int32_t* dna, uint32_t P, uint32_t* edge_ids,
                   uint32_t* step_info, double* dist, double* turn, uint8_t* flags, void* stream)
{
    int r = ready(c); if (r) return r;
    return launch_path_costs(c, dna, P, edge_ids, step_info, dist, turn, flags,
                             (cudaStream_t)stream);
}

int gcp_coverage_masked(gcp_ctx* c, const uint32_t* step_info, uint32_t P, uint32_t* cov,
                        uint8_t* flags, void* stream)
{
    int r = ready(c); if (r) return r;
    return launch_coverage_masked(c, step_info, P, cov, flags, (cudaStream_t)stream);
}

int gcp_coverage(gcp_ctx* c, const uint32_t* edge_ids, uint32_t P, uint32_t* cov, uint8_t* flags,
                 void* stream)
{
    int r = ready(c); if (r) return r;
    c->cov_detail = 1;
    return launch_coverage(c, edge_ids, P, cov, flags, (cudaStream_t)stream);
}

int gcp_loop_closures(gcp_ctx* c, const int32_t* dna, uint32_t P, const double* dist,
                      const double* turn, const uint32_t* cov, int32_t* lc, double* obj,
                      double* neg_cov, uint8_t* flags, void* stream)
{
    int r = ready(c); if (r) return r;
    return launch_loop_closures(c, dna, P, dist, turn, cov, lc, obj, neg_cov, flags,
                                (cudaStream_t)stream);
}

int gcp_eval(gcp_ctx* c, const int32_t* dna, uint32_t P, const gcp_eval_out* out, void* scratch,
             size_t scratch_bytes, void* stream)
{
    int r = ready(c); if (r) return r;
    if (!out) return gcp_fail(c, -1, "gcp_eval: out is NULL");
    if (P == 0) return 0;
    // objectives-only fast path: binary map, wall coverage only -> pre-masked footprint store
    const bool fast = c->have_masked && c->p.coverage_blend == 1.0 && !out->cov && !c->force_general;
    // staged objectives-only path: quarter-granular kernel (any map) when its tables exist, else
    // the sub-tile kernel (binary maps only)
    const bool fast_q = fast && c->t.m_q_info != nullptr;
    const bool fast_staged = fast && (fast_q || c->n_gray == 0);
    if (fast && c->fused_eval && c->t.m_q_info && eval_fused_fits(c))   // one kernel, no scratch (gcp_fused.cu)
        return launch_eval_fused(c, dna, 0, P, out, (cudaStream_t)stream);
    // any coverage_blend / all six pixel counters: the same kernel over the unmasked store
    if (c->fused_eval && eval_fused_fits(c, true))
        return launch_eval_fused(c, dna, 0, P, out, (cudaStream_t)stream, true);
    if (scratch_bytes < gcp_eval_scratch_bytes(c, P) || !scratch)
        return gcp_fail(c, -1, "gcp_eval: scratch too small (%zu < %zu)", scratch_bytes,
                        gcp_eval_scratch_bytes(c, P));
    const size_t A = c->p.num_agents, L = c->p.max_dna_len;
    char* s = reinterpret_cast<char*>(scratch);
    uint32_t* edge_ids = reinterpret_cast<uint32_t*>(s); s += align256(P * A * L * 4);
    uint32_t* step_info = reinterpret_cast<uint32_t*>(s); s += align256(P * A * L * 4);
    double* dist = reinterpret_cast<double*>(s);         s += align256(P * A * 8);
    double* turn = reinterpret_cast<double*>(s);         s += align256(P * A * 8);
    uint32_t* cov = reinterpret_cast<uint32_t*>(s);      s += align256((size_t)P * 32);
    uint8_t* flags = reinterpret_cast<uint8_t*>(s);
    if (out->dist) dist = out->dist;
    if (out->turn) turn = out->turn;
    if (out->cov) cov = out->cov;
    if (out->flags) flags = out->flags;
    cudaStream_t st = (cudaStream_t)stream;
    if ((r = launch_path_costs(c, dna, P, fast_staged ? nullptr : edge_ids, fast_staged ? step_info : nullptr,
                               dist, turn, flags, st, fast_q ? 1 : 0))) return r;
    if (fast_q) {
        if ((r = launch_coverage_quarters(c, step_info, P, cov, flags, st))) return r;
    } else if (fast_staged) {
        if ((r = launch_coverage_masked(c, step_info, P, cov, flags, st))) return r;
    } else {
        c->cov_detail = out->cov ? 1 : 0;   // objectives-only mode skips unused pixel counters
        if ((r = launch_coverage(c, edge_ids, P, cov, flags, st))) return r;
    }
    if ((r = launch_loop_closures(c, dna, P, dist, turn, cov, out->lc_mat, out->obj, out->neg_cov,
                                  flags, st))) return r;
    return 0;
}

static bool fast_path(gcp_ctx* c, const gcp_eval_out* out)
{
    return c->have_masked && c->p.coverage_blend == 1.0 && !out->cov &&
           !c->force_general && c->fused_eval && c->t.m_q_info && eval_fused_fits(c);
}

int gcp_eval16(gcp_ctx* c, const int16_t* dna, uint32_t P, const gcp_eval_out* out, void* stream)
{
    int r = ready(c); if (r) return r;
    if (!out) return gcp_fail(c, -1, "gcp_eval16: out is NULL");
    if (P == 0) return 0;
    if (c->t.N > 32767) return gcp_fail(c, -1, "gcp_eval16: int16 genes need N <= 32767 waypoints (N=%u)", c->t.N);
    if (!fast_path(c, out))
        return gcp_fail(c, -1, "gcp_eval16: only the fused objectives path takes int16 genes "
                               "(binary map, coverage_blend == 1, no pixel counters)");
    return launch_eval_fused(c, dna, 1, P, out, (cudaStream_t)stream);
}

int gcp_eval16_available(gcp_ctx* c)
{
    if (!c || ready(c)) return 0;
    gcp_eval_out probe;
    memset(&probe, 0, sizeof(probe));
    return (c->t.N <= 32767 && fast_path(c, &probe)) ? 1 : 0;
}

static int eval_host_impl(gcp_ctx* c, const void* dna_host, int gene16, uint32_t P, double* obj_host,
                          uint8_t* flags_host);

static bool narrow_on_host(gcp_ctx* c)
{
    return c->t.N <= 32767u &&
           (c->host_narrow > 0 || (c->host_narrow < 0 && (c->host_threads > 0 || host_narrow_auto())));
}

int gcp_host_narrow_active(gcp_ctx* c)
{
    if (!c || ready(c)) return 0;
    gcp_eval_out probe;
    memset(&probe, 0, sizeof(probe));
    return (fast_path(c, &probe) && narrow_on_host(c)) ? 1 : 0;
}

int gcp_eval_host(gcp_ctx* c, const int32_t* dna_host, uint32_t P, double* obj_host,
                  uint8_t* flags_host)
{
    return eval_host_impl(c, dna_host, 0, P, obj_host, flags_host);
}

int gcp_eval_host16(gcp_ctx* c, const int16_t* dna_host, uint32_t P, double* obj_host,
                    uint8_t* flags_host)
{
    if (c && c->have_graph && c->t.N > 32767)
        return gcp_fail(c, -1, "gcp_eval_host16: int16 genes need N <= 32767 waypoints");
    return eval_host_impl(c, dna_host, 1, P, obj_host, flags_host);
}

static int eval_host_impl(gcp_ctx* c, const void* dna_host, int gene16, uint32_t P, double* obj_host,
                          uint8_t* flags_host)
{
    int r = ready(c); if (r) return r;
    if (P == 0) return 0;
    GCP_CUDA(c, cudaSetDevice(c->device));
    const size_t A = c->p.num_agents, L = c->p.max_dna_len;
    // chunks so that copy-in, kernels and copy-out of neighbouring chunks overlap: three staging
    // buffers on three streams; the first chunks are small so that the GPU starts computing after a
    // quarter of a chunk's copy time instead of a whole one
    constexpr int NBUF = 3;
    uint32_t chunk = 4096;
    if (chunk > P) chunk = P;
    const size_t gsz = gene16 ? 2 : 4;
    gcp_eval_out probe;
    memset(&probe, 0, sizeof(probe));
    const bool fused = fast_path(c, &probe);
    if (gene16 && !fused)
        return gcp_fail(c, -1, "gcp_eval_host16: only the fused objectives path takes int16 genes");
    // int32 genes of a world whose ids fit int16: host threads narrow them on the way (gcp_host.cu)
    if (!gene16 && fused && narrow_on_host(c))
        return eval_host_narrowed(c, reinterpret_cast<const int32_t*>(dna_host), P, obj_host, flags_host);
    const size_t dna_b = align256((size_t)chunk * A * L * gsz);
    const size_t scr_b = fused ? 256 : align256(gcp_eval_scratch_bytes(c, chunk));
    const size_t obj_b = align256((size_t)chunk * 16);
    const size_t flg_b = align256((size_t)chunk * 4);
    const size_t per = dna_b + scr_b + obj_b + flg_b;
    if (c->stage_bytes < NBUF * per) {
        if (c->stage_dev) cudaFree(c->stage_dev);
        c->stage_dev = nullptr; c->stage_bytes = 0;
        GCP_CUDA(c, cudaMalloc(&c->stage_dev, NBUF * per));
        c->stage_bytes = NBUF * per;
    }
    for (int s = 0; s < NBUF; ++s)
        if (!c->hstream[s]) GCP_CUDA(c, cudaStreamCreateWithFlags(&c->hstream[s], cudaStreamNonBlocking));
    int k = 0;
    for (uint32_t o0 = 0; o0 < P; ++k) {
        uint32_t n = (k == 0) ? chunk / 4 : (k == 1) ? chunk / 2 : chunk;      // ramp-up
        if (n < 1) n = 1;
        if (n > P - o0) n = P - o0;
        cudaStream_t st = c->hstream[k % NBUF];
        char* base = reinterpret_cast<char*>(c->stage_dev) + (size_t)(k % NBUF) * per;
        void* d_dna = base;
        void* d_scr = base + dna_b;
        double* d_obj = reinterpret_cast<double*>(base + dna_b + scr_b);
        uint8_t* d_flg = reinterpret_cast<uint8_t*>(base + dna_b + scr_b + obj_b);
        GCP_CUDA(c, cudaMemcpyAsync(d_dna, reinterpret_cast<const char*>(dna_host) + (size_t)o0 * A * L * gsz,
                                    (size_t)n * A * L * gsz, cudaMemcpyHostToDevice, st));
        gcp_eval_out out;
        memset(&out, 0, sizeof(out));
        out.obj = d_obj;
        out.flags = d_flg;
        c->scratch_slot = k % NBUF;               // long-genome scratch of this pipeline slot
        if (gene16) r = gcp_eval16(c, reinterpret_cast<const int16_t*>(d_dna), n, &out, st);
        else r = gcp_eval(c, reinterpret_cast<const int32_t*>(d_dna), n, &out, d_scr, scr_b, st);
        c->scratch_slot = 0;
        if (r) return r;
        if (obj_host)
            GCP_CUDA(c, cudaMemcpyAsync(obj_host + (size_t)o0 * 2, d_obj, (size_t)n * 16,
                                        cudaMemcpyDeviceToHost, st));
        if (flags_host)
            GCP_CUDA(c, cudaMemcpyAsync(flags_host + (size_t)o0 * 4, d_flg, (size_t)n * 4,
                                        cudaMemcpyDeviceToHost, st));
        o0 += n;
    }
    for (int s = 0; s < NBUF; ++s) GCP_CUDA(c, cudaStreamSynchronize(c->hstream[s]));
    return 0;
}

size_t gcp_rank_scratch_bytes(gcp_ctx* c, uint32_t n) { return c ? rank_scratch_bytes(c, n) : 0; }

int gcp_maximin(gcp_ctx* c, const double* obj, uint32_t n, double constr, uint32_t r0, uint32_t r1,
                double* obj_sc, double* fit, void* scratch, size_t sb, void* stream)
{
    if (!c) return -1;
    return launch_maximin(c, obj, n, constr, r0, r1, obj_sc, fit, scratch, sb, (cudaStream_t)stream);
}

int gcp_arena_order(gcp_ctx* c, const double* fit, uint32_t n, int32_t* order, uint8_t* nondom,
                    void* scratch, size_t sb, void* stream)
{
    if (!c) return -1;
    return launch_arena_order(c, fit, n, order, nondom, scratch, sb, (cudaStream_t)stream);
}

int gcp_set_vary_params(gcp_ctx* c, const gcp_vary_params* p)
{
    if (!c || !p) return -1;
    if (p->num_muterpolations < 0 || p->num_muterpolations > 64)
        return gcp_fail(c, -1, "num_muterpolations must be in [0,64]");
    if (p->path_memory < 0 || p->path_memory > 16) return gcp_fail(c, -1, "path_memory must be in [0,16]");
    if (!(p->log_scale_weight > 0)) return gcp_fail(c, -1, "log_scale_weight must be > 0");
    if (c->have_graph && c->max_degree > 32)
        return gcp_fail(c, -3, "variation kernels support out-degree <= 32, graph has %u", c->max_degree);
    c->vary = *p;
    c->have_vary = true;
    return 0;
}

static int vary_ready(gcp_ctx* c)
{
    int r = ready(c); if (r) return r;
    if (!c->have_vary) return gcp_fail(c, -1, "gcp_set_vary_params not called");
    return 0;
}

int gcp_low_var_select(gcp_ctx* c, const double* fit, uint32_t m, double gamma, double r0,
                       double* scratch, int32_t* out, void* stream)
{
    if (!c) return -1;
    return launch_low_var_select(c, fit, m, gamma, r0, scratch, out, (cudaStream_t)stream);
}

int gcp_crossover(gcp_ctx* c, const int32_t* parents, const int32_t* strong, uint32_t P,
                  uint64_t seed, uint32_t gen, int32_t* children, void* stream)
{
    int r = vary_ready(c); if (r) return r;
    return launch_crossover(c, parents, strong, P, seed, gen, children, (cudaStream_t)stream);
}

int gcp_mutate(gcp_ctx* c, int32_t* dna, uint32_t P, uint64_t seed, uint32_t gen, uint32_t salt,
               void* stream)
{
    int r = vary_ready(c); if (r) return r;
    return launch_mutate(c, dna, P, seed, gen, salt, (cudaStream_t)stream);
}

double gcp_low_var_r0(uint64_t seed, uint32_t gen, uint32_t m)
{
    // the single U(0, 1/M) draw of lowVarSample (gori_tools.py:217), from the seed/generation
    uint64_t z = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(gen + 1));
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    return ((double)(z >> 11) * (1.0 / 9007199254740992.0)) / (double)(m ? m : 1);
}

int gcp_select_vary(gcp_ctx* c, const int32_t* parents, const double* parent_fit, uint32_t P,
                    double gamma, uint64_t seed, uint32_t gen, int32_t* children, void* scratch,
                    size_t scratch_bytes, void* stream)
{
    int r = vary_ready(c); if (r) return r;
    if (P == 0) return 0;
    const size_t nt = ((size_t)P + 1023) / 1024;
    if (!scratch || scratch_bytes < (size_t)P * 12 + nt * 16)
        return gcp_fail(c, -1, "select_vary: scratch too small (need 12 * n_org + 16 * ceil(n_org / 1024) bytes)");
    double* cum = reinterpret_cast<double*>(scratch);
    int32_t* strong = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(scratch) + ((size_t)P + 2 * nt) * 8);
    const double r0 = gcp_low_var_r0(seed, gen, P);
    cudaStream_t st = (cudaStream_t)stream;
    if ((r = launch_low_var_select(c, parent_fit, P, gamma, r0, cum, strong, st))) return r;
    if ((r = launch_crossover(c, parents, strong, P, seed, gen, children, st))) return r;
    return launch_mutate(c, children, P, seed, gen, 0, st);
}

int gcp_random_walks(gcp_ctx* c, int32_t* dna, uint32_t P, const int32_t* start_idx, int32_t path_len,
                     uint64_t seed, uint32_t batch, void* stream)
{
    int r = vary_ready(c); if (r) return r;
    return launch_random_walks(c, dna, P, start_idx, path_len, seed, batch, (cudaStream_t)stream);
}

int gcp_gather_survivors(gcp_ctx* c, const int32_t* order, uint32_t P, const int32_t* par_dna,
                         const int32_t* kid_dna, const double* glad_obj, const double* glad_fit,
                         int32_t* out_dna, double* out_obj, double* out_fit, void* stream)
{
    int r = ready(c); if (r) return r;
    return launch_gather_survivors(c, order, P, par_dna, kid_dna, glad_obj, glad_fit, out_dna,
                                   out_obj, out_fit, (cudaStream_t)stream);
}

// ---- population shards: CUDA IPC plumbing + the sharded variation / truncation entry points ----
int gcp_shard_alloc(gcp_ctx* c, size_t bytes, void** dev_ptr, uint8_t ipc_handle[64])
{
    if (!c || !dev_ptr || !ipc_handle) return -1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    *dev_ptr = nullptr;
    GCP_CUDA(c, cudaSetDevice(c->device));
    void* p = nullptr;
    GCP_CUDA(c, cudaMalloc(&p, bytes ? bytes : 256));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return gcp_fail(c, -2, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, 64);
    c->shard_own.push_back(p);
    *dev_ptr = p;
    return 0;
}

static bool take(std::vector<void*>& v, void* p)
{
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == p) { v.erase(v.begin() + i); return true; }
    return false;
}

int gcp_shard_free(gcp_ctx* c, void* p)
{
    if (!c) return -1;
    if (!take(c->shard_own, p)) return gcp_fail(c, -1, "gcp_shard_free: not a shard of this context");
    GCP_CUDA(c, cudaSetDevice(c->device));
    GCP_CUDA(c, cudaFree(p));
    return 0;
}

int gcp_shard_open(gcp_ctx* c, const uint8_t ipc_handle[64], void** dev_ptr)
{
    if (!c || !dev_ptr || !ipc_handle) return -1;
    *dev_ptr = nullptr;
    GCP_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
        return gcp_fail(c, -2, "cudaIpcOpenMemHandle failed: %s (peer shards need GPUs with peer access "
                               "in one box, and a different process than the owner)", cudaGetErrorString(e));
    c->shard_peer.push_back(p);
    *dev_ptr = p;
    return 0;
}

int gcp_shard_close(gcp_ctx* c, void* p)
{
    if (!c) return -1;
    if (!take(c->shard_peer, p)) return gcp_fail(c, -1, "gcp_shard_close: not a peer shard of this context");
    GCP_CUDA(c, cudaSetDevice(c->device));
    GCP_CUDA(c, cudaIpcCloseMemHandle(p));
    return 0;
}

int gcp_crossover_shards(gcp_ctx* c, const gcp_shards* parents, int gene_bytes, const int32_t* strong,
                         uint32_t P, uint32_t pair_begin, uint32_t n_pairs, uint64_t seed, uint32_t gen,
                         void* children, void* stream)
{
    int r = vary_ready(c); if (r) return r;
    return launch_crossover_shards(c, parents, gene_bytes, strong, P, pair_begin, n_pairs, seed, gen,
                                   children, (cudaStream_t)stream);
}

int gcp_mutate_genes(gcp_ctx* c, void* dna, int gene_bytes, uint32_t P, uint32_t org_offset, uint64_t seed,
                     uint32_t gen, uint32_t salt, void* stream)
{
    int r = vary_ready(c); if (r) return r;
    return launch_mutate_genes(c, dna, gene_bytes, P, org_offset, seed, gen, salt, (cudaStream_t)stream);
}

int gcp_gather_shards(gcp_ctx* c, const int32_t* order, uint32_t P, uint32_t k_begin, uint32_t n_local,
                      const gcp_shards* parents, const gcp_shards* children, int gene_bytes,
                      const double* glad_obj, const double* glad_fit, void* out_dna, double* out_obj,
                      double* out_fit, void* stream)
{
    int r = ready(c); if (r) return r;
    if ((out_obj == nullptr) != (out_fit == nullptr))
        return gcp_fail(c, -1, "gcp_gather_shards: out_obj and out_fit go together");
    return launch_gather_shards(c, order, P, k_begin, n_local, parents, children, gene_bytes, glad_obj,
                                glad_fit, out_dna, out_obj, out_fit, (cudaStream_t)stream);
}

int gcp_peer_barrier(gcp_ctx* c, const gcp_shards* flags, uint32_t rank, uint32_t epoch,
                     uint32_t* timed_out, void* stream)
{
    if (!c) return -1;
    return launch_peer_barrier(c, flags, rank, epoch, timed_out, (cudaStream_t)stream);
}

int gcp_set_option(gcp_ctx* c, const char* name, int64_t value)
{
    if (!c || !name) return -1;
    if (!strcmp(name, "rank_algo")) c->rank_algo = (int)value;
    else if (!strcmp(name, "sort_onesweep")) c->sort_onesweep = (int)value;
    else if (!strcmp(name, "host_narrow")) c->host_narrow = (int)value;
    else if (!strcmp(name, "host_threads")) { if (value < 0 || value > 64) return gcp_fail(c, -1, "host_threads must be in [0,64]"); c->host_threads = (int)value; }
    else if (!strcmp(name, "force_general")) c->force_general = (int)value;
    else if (!strcmp(name, "fused_eval")) c->fused_eval = (int)value;
    else if (!strcmp(name, "force_lc_big")) c->force_lc_big = (int)value;
    else if (!strcmp(name, "vary_org_offset")) {
        if (value < 0 || (value & 1)) return gcp_fail(c, -1, "vary_org_offset must be even and >= 0");
        c->vary_org_offset = (uint32_t)value;
    }
    else if (!strcmp(name, "cov_table")) c->cov_table = (int)value;
    else if (!strcmp(name, "cov_maxpairs")) c->cov_maxpairs = (int)value;
    else if (!strcmp(name, "lc_min_parts")) c->lc_min_parts = (int)value;
    else if (!strcmp(name, "cov_threads")) c->cov_threads = (int)value;
    else if (!strcmp(name, "cov_dedup")) c->cov_dedup = (int)value;
    else if (!strcmp(name, "cov_buckets")) c->cov_buckets = (int)value;
    else if (!strcmp(name, "cov_prefetch")) c->cov_prefetch = (int)value;
    else if (!strcmp(name, "cov_split")) c->cov_split = (int)value;
    else if (!strcmp(name, "lc_split")) c->lc_split = (int)value;
    else if (!strcmp(name, "cov_split_cap")) c->cov_split_cap = (int)value;
    else if (!strcmp(name, "cov_split_k")) c->cov_split_k = (int)value;
    else if (!strcmp(name, "lc_split_keys")) c->lc_split_keys = (int)value;
    else if (!strcmp(name, "lc_split_slack")) c->lc_split_slack = (int)value;
    else return gcp_fail(c, -1, "gcp_set_option: unknown option '%s'", name);
    return 0;
}

uint64_t gcp_launch_count(gcp_ctx* c) { return c ? c->launches : 0; }

int gcp_set_timing(gcp_ctx* c, int enabled)
{
    if (!c) return -1;
    c->timing = enabled ? 1 : 0;
    if (!enabled) for (int s = 0; s < GCP_NUM_STAGES; ++s) c->ev_valid[s] = false;
    return 0;
}

float gcp_stage_ms(gcp_ctx* c, int stage)
{
    if (!c || stage < 0 || stage >= GCP_NUM_STAGES || !c->ev_valid[stage]) return -1.0f;
    if (cudaEventSynchronize(c->ev_end[stage]) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev_beg[stage], c->ev_end[stage]) != cudaSuccess) return -1.0f;
    return ms;
}

} // extern "C"
