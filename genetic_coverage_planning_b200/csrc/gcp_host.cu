// Host side of gcp_eval_host for int32 chromosomes on worlds whose waypoint ids fit int16:
// a small pool of host threads narrows the population to int16 (saturating, so a gene outside
// [-1, N) stays outside and is still reported), chunk by chunk into a pinned staging copy, while the
// caller's thread ships finished chunks to the GPU.  The PCIe link carries half the bytes; the
// conversion is part of the call (and of bench.py's timed end-to-end region).
#include "gcp_common.cuh"
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <emmintrin.h>
#include <functional>
#include <mutex>
#include <thread>

struct GcpHostPool {
    int T = 0;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::function<void(int, int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;

    explicit GcpHostPool(int n) : T(n)
    {
        for (int i = 0; i < T; ++i) th.emplace_back([this, i] { worker(i); });
    }
    ~GcpHostPool()
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_go.notify_all();
        for (auto& t : th) t.join();
    }
    void worker(int i)
    {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int, int)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            f(i, T);
            { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv_done.notify_all(); }
        }
    }
    // starts f(thread, T) on every worker and returns at once; wait() joins them
    void start(std::function<void(int, int)> f)
    {
        { std::lock_guard<std::mutex> lk(m); job = std::move(f); pending = T; ++gen; }
        cv_go.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

// dst[i] = saturate_int16(src[i]) (SSE2, the x86-64 baseline)
static void narrow_i32_i16(const int32_t* src, int16_t* dst, size_t n)
{
    size_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 8));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 12));
        if (aligned) {                          // the staging copy is write-only here: streaming stores
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi32(a, b));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm_packs_epi32(c, d));
        } else {
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_packs_epi32(a, b));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm_packs_epi32(c, d));
        }
    }
    for (; i < n; ++i) {
        const int32_t g = src[i];
        dst[i] = (int16_t)(g > 32767 ? 32767 : g < -32768 ? -32768 : g);
    }
    _mm_sfence();
}

static int host_threads_auto()
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    int t = (int)hw / ranks - 1;                 // leave the caller's thread a core
    if (t > 12) t = 12;                          // 12 threads: 3.8 ms for 315 MB (8: 4.4 ms, 15: 3.8 ms; profiles/exp_e2e)
    if (t < 1) t = 1;
    return t;
}

// Does narrowing on the host pay?  One rank per host with at least six worker threads: 315 MB of int32 genes
// take 6.0 ms over PCIe as they are, 5.2 / 4.4 / 3.8 ms narrowed by 6 / 8 / 12 threads (4 threads: 6.9 ms).
// With several ranks per host it does not: the DMA engines of N GPUs together read host memory faster
// (8 ranks: 181 GB/s) than the host's cores convert it (~135 GB/s of reads + writes on the pool's 32-vCPU VM;
// 8 ranks x 3 threads: 3.3 ms against 1.7 ms with int32 on the wire).
bool host_narrow_auto()
{
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    return ranks == 1 && host_threads_auto() >= 6;
}

void gcp_host_free(gcp_ctx* c)
{
    delete reinterpret_cast<GcpHostPool*>(c->host_pool);
    c->host_pool = nullptr;
    if (c->host_stage16) cudaFreeHost(c->host_stage16);
    c->host_stage16 = nullptr; c->host_stage16_bytes = 0;
    for (int s = 0; s < 3; ++s) if (c->hstage_dev[s]) { cudaFree(c->hstage_dev[s]); c->hstage_dev[s] = nullptr; }
    c->hstage_dev_bytes = 0;
}

// int32 host chromosomes -> objectives / flags on the host, through int16 on the wire.
int eval_host_narrowed(gcp_ctx* c, const int32_t* dna_host, uint32_t P, double* obj_host, uint8_t* flags_host)
{
    const size_t AL = (size_t)c->p.num_agents * c->p.max_dna_len;
    const int want_T = c->host_threads > 0 ? c->host_threads : host_threads_auto();
    GcpHostPool* pool = reinterpret_cast<GcpHostPool*>(c->host_pool);
    if (!pool || pool->T != want_T) {
        delete pool;
        pool = new GcpHostPool(want_T);
        c->host_pool = pool;
    }
    const size_t need = (size_t)P * AL * 2;
    if (c->host_stage16_bytes < need) {
        if (c->host_stage16) cudaFreeHost(c->host_stage16);
        c->host_stage16 = nullptr; c->host_stage16_bytes = 0;
        GCP_CUDA(c, cudaHostAlloc(&c->host_stage16, need, cudaHostAllocDefault));
        c->host_stage16_bytes = need;
    }
    int16_t* stage = reinterpret_cast<int16_t*>(c->host_stage16);
    constexpr int NBUF = 3;
    const uint32_t chunk = P < 4096u ? P : 4096u;
    const size_t dna_b = ((size_t)chunk * AL * 2 + 255) & ~(size_t)255;
    const size_t obj_b = ((size_t)chunk * 16 + 255) & ~(size_t)255;
    const size_t per = dna_b + obj_b + (((size_t)chunk * 4 + 255) & ~(size_t)255);
    if (c->hstage_dev_bytes < per) {
        for (int s = 0; s < NBUF; ++s) {
            if (c->hstage_dev[s]) cudaFree(c->hstage_dev[s]);
            c->hstage_dev[s] = nullptr;
            GCP_CUDA(c, cudaMalloc(&c->hstage_dev[s], per));
        }
        c->hstage_dev_bytes = per;
    }
    for (int s = 0; s < NBUF; ++s)
        if (!c->hstream[s]) GCP_CUDA(c, cudaStreamCreateWithFlags(&c->hstream[s], cudaStreamNonBlocking));
    // chunk table (ramped first chunks: the GPU starts after a quarter of a chunk's conversion)
    std::vector<uint32_t> beg;
    for (uint32_t o = 0, k = 0; o < P; ++k) {
        uint32_t n = k == 0 ? chunk / 4 : k == 1 ? chunk / 2 : chunk;
        if (n < 1) n = 1;
        if (n > P - o) n = P - o;
        beg.push_back(o);
        o += n;
    }
    beg.push_back(P);
    const size_t nchunks = beg.size() - 1;
    // work units: granules of GRAN genes handed out IN ORDER from one atomic counter, so the chunks are
    // finished front to back by whichever threads are free - a worker that loses its core for a while
    // holds back one granule, not its share of every chunk
    constexpr size_t GRAN = 32768;
    std::vector<size_t> gbeg(nchunks + 1, 0);
    for (size_t k = 0; k < nchunks; ++k)
        gbeg[k + 1] = gbeg[k] + (((size_t)(beg[k + 1] - beg[k]) * AL + GRAN - 1) / GRAN);
    const size_t ngran = gbeg[nchunks];
    std::vector<std::atomic<int>> done(nchunks);
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<size_t> next{0};
    pool->start([&](int, int) {
        size_t k = 0;
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= ngran) break;
            while (i >= gbeg[k + 1]) ++k;                                   // i only grows
            const size_t g0 = (size_t)beg[k] * AL, g1 = (size_t)beg[k + 1] * AL;
            const size_t a = g0 + (i - gbeg[k]) * GRAN, b = a + GRAN;
            narrow_i32_i16(dna_host + a, stage + a, (b < g1 ? b : g1) - a);
            done[k].fetch_add(1, std::memory_order_release);
        }
    });
    int rc = 0;
    for (size_t k = 0; k < nchunks && rc == 0; ++k) {
        while (done[k].load(std::memory_order_acquire) < (int)(gbeg[k + 1] - gbeg[k])) _mm_pause();
        const uint32_t o0 = beg[k], n = beg[k + 1] - beg[k];
        cudaStream_t st = c->hstream[k % NBUF];
        char* base = reinterpret_cast<char*>(c->hstage_dev[k % NBUF]);
        int16_t* d_dna = reinterpret_cast<int16_t*>(base);
        double* d_obj = reinterpret_cast<double*>(base + dna_b);
        uint8_t* d_flg = reinterpret_cast<uint8_t*>(base + dna_b + obj_b);
        cudaError_t e = cudaMemcpyAsync(d_dna, stage + (size_t)o0 * AL, (size_t)n * AL * 2, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = gcp_fail(c, -2, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e)); break; }
        gcp_eval_out out;
        memset(&out, 0, sizeof(out));
        out.obj = d_obj;
        out.flags = d_flg;
        rc = gcp_eval16(c, d_dna, n, &out, st);
        if (rc) break;
        if (obj_host) cudaMemcpyAsync(obj_host + (size_t)o0 * 2, d_obj, (size_t)n * 16, cudaMemcpyDeviceToHost, st);
        if (flags_host) cudaMemcpyAsync(flags_host + (size_t)o0 * 4, d_flg, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    }
    pool->wait();                               // the workers hold references to this frame
    for (int s = 0; s < NBUF; ++s) {
        cudaError_t e = cudaStreamSynchronize(c->hstream[s]);
        if (e != cudaSuccess && rc == 0) rc = gcp_fail(c, -2, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    }
    return rc;
}
