// Setup path N4: PathMaker.computeTraversableGraph (pathmaker.py:119-155) with
// Mappy.lineCollisionCheck (mappy.py:124-159) on the device.
//
// Reference: dense N x N displacement / distance / bearing tables; per waypoint and 45-degree
// sector the NEAREST other waypoint (np.argmin: lowest index on ties) is a candidate if it is
// closer than max_traverse_dist; a candidate edge survives if no pixel of the map dilated by
// int(safety_buffer / 3 / scale) 3x3 passes lies within 1 px of the infinite line through the two
// waypoints while its row OR its column is strictly between the end points.  The reference
// re-dilates the whole map and scans all its obstacle pixels for every candidate edge.
//
// Here: k_dilate builds the dilated occupancy once; k_graph_candidates (thread per waypoint)
// searches the 3 x 3 grid cells around the waypoint (cell = max distance, lists sorted by waypoint
// index on the host) for the nearest waypoint per sector - only waypoints closer than the
// maximum can ever become edges, so the window is exact; k_graph_collide (warp per candidate)
// scans the part of the dilated map that can satisfy the reference's test.  Decisions are taken
// on the same values as numpy: squared distances and the line form are integers; sqrt, the
// multiplication by scale and the division are correctly rounded; the sector of an integer
// displacement can only be mis-binned if its bearing lies within rounding error of an odd multiple
// of pi/8, which is flagged (ambiguous) and redone on the host (graphs.py).
#include "gcp_common.cuh"
#include <cmath>

struct GrDev {
    const uint8_t* occ; uint8_t* dil; uint32_t H, W;
    const int2* xy; uint32_t N;
    const uint32_t* cell_ptr; const uint32_t* cell_items; int ncx, ncy;
    int32_t* nbr; uint8_t* amb;
    gcp_graph_params gp;
};

__global__ void k_dilate(GrDev d)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= d.W || y >= d.H) return;
    const int n = d.gp.num_dilations;                  // n passes of a 3x3 max = one (2n+1)^2 max; the
    uint8_t v = 0;                                     // border does not contribute (cv2 default)
    for (int dy = -n; dy <= n && !v; ++dy) {
        const int yy = (int)y + dy;
        if (yy < 0 || yy >= (int)d.H) continue;
        for (int dx = -n; dx <= n; ++dx) {
            const int xx = (int)x + dx;
            if (xx < 0 || xx >= (int)d.W) continue;
            if (d.occ[(size_t)yy * d.W + xx]) { v = 1; break; }
        }
    }
    d.dil[(size_t)y * d.W + x] = v;
}

__global__ void __launch_bounds__(128)
k_graph_candidates(GrDev d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N) return;
    const int2 p = d.xy[i];
    const int cell = d.gp.cell;
    const int cy = min(max(p.x / cell, 0), d.ncy - 1), cx = min(max(p.y / cell, 0), d.ncx - 1);
    long long best_d2[8];
    int best_j[8];
    #pragma unroll
    for (int s = 0; s < 8; ++s) { best_d2[s] = 0x7FFFFFFFFFFFFFFFll; best_j[s] = -1; }
    bool amb = false;
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, d.ncy - 1); ++yy)
        for (int xx = max(cx - 1, 0); xx <= min(cx + 1, d.ncx - 1); ++xx) {
            const uint32_t c = (uint32_t)yy * d.ncx + xx;
            for (uint32_t k = d.cell_ptr[c]; k < d.cell_ptr[c + 1]; ++k) {
                const uint32_t j = d.cell_items[k];
                if (j == i) continue;                                       // pathmaker.py:139-140 (self sits in sector 4)
                const int2 q = d.xy[j];
                const long long dr = (long long)q.x - p.x, dc = (long long)q.y - p.y;   // :121-122
                const long long d2 = dr * dr + dc * dc;
                const double dist = __dmul_rn(__dsqrt_rn((double)d2), d.gp.scale);       // :123
                if (!(dist < d.gp.max_dist)) continue;                                    // :147
                const double ang = atan2((double)dc, (double)dr);                         // :124
                int s = 0;
                #pragma unroll
                for (int b = 0; b < 8; ++b) {                                             // np.digitize, :131
                    s += (d.gp.sectors[b] <= ang);
                    if (fabs(ang - d.gp.sectors[b]) < d.gp.eps_angle) amb = true;
                }
                if (s == 8) s = 0;                                                         // :132
                #pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (b == s && (d2 < best_d2[b] || (d2 == best_d2[b] && (int)j < best_j[b]))) {   // np.argmin: first minimum
                        best_d2[b] = d2; best_j[b] = (int)j;
                    }
            }
        }
    #pragma unroll
    for (int s = 0; s < 8; ++s) d.nbr[(size_t)i * 8 + s] = best_j[s];
    d.amb[i] = amb ? 1 : 0;
}

// one warp per (waypoint, sector) candidate: Mappy.lineCollisionCheck (mappy.py:124-159)
__global__ void __launch_bounds__(256)
k_graph_collide(GrDev d)
{
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= d.N * 8u) return;
    const int j = d.nbr[wid];
    if (j < 0) return;
    const int2 p = d.xy[wid >> 3], q = d.xy[j];
    const long long x1 = p.x, y1 = p.y, x2 = q.x, y2 = q.y;
    const long long a = y2 - y1, b = x2 - x1, c = x2 * y1 - y2 * x1;       // :127-133
    bool hit = false;
    if (a == 0 && b == 0) {
        hit = true;                                                         // :134-135 returns False: no edge
    } else {
        const double len = __dsqrt_rn((double)(a * a + b * b));
        // a pixel whose row is strictly between the end points and that is within 1 px of the line
        // can lie up to len / |b| columns beside the segment (|b| >= 2 there), and vice versa
        const int m = (int)(len * 0.5) + 2;
        const int r_lo = max((int)min(x1, x2) - m, 0), r_hi = min((int)max(x1, x2) + m, (int)d.H - 1);
        const int c_lo = max((int)min(y1, y2) - m, 0), c_hi = min((int)max(y1, y2) + m, (int)d.W - 1);
        const int wdt = c_hi - c_lo + 1, cnt = (r_hi - r_lo + 1) * wdt;
        for (int t = lane; t < cnt && !hit; t += 32) {
            const int r = r_lo + t / wdt, cc = c_lo + t % wdt;
            if (!d.dil[(size_t)r * d.W + cc]) continue;
            const bool row_out = (r <= x2 && r <= x1) || (r >= x2 && r >= x1);          // :145-151
            const bool col_out = (cc <= y2 && cc <= y1) || (cc >= y2 && cc >= y1);
            if (row_out && col_out) continue;
            long long num = a * r - b * cc + c;
            if (num < 0) num = -num;
            if (__ddiv_rn((double)num, len) <= 1.0) hit = true;                          // :142, :154
        }
    }
    if (__any_sync(0xFFFFFFFFu, hit) && lane == 0) d.nbr[wid] = -1;
}

template <class T>
static cudaError_t up(T** dst, const T* src, size_t n)
{
    cudaError_t e = cudaMalloc(dst, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return n ? cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
}

extern "C" int gcp_build_graph(int device, const uint8_t* occ_h, uint32_t H, uint32_t W, const int32_t* xy_h,
                               uint32_t N, const uint32_t* cell_ptr_h, const uint32_t* cell_items_h,
                               const gcp_graph_params* gp, int32_t* nbr_h, uint8_t* ambiguous_h, float* kernel_ms)
{
    if (!occ_h || !xy_h || !cell_ptr_h || !cell_items_h || !gp || !nbr_h || !ambiguous_h)
        return gcp_fail(nullptr, -1, "gcp_build_graph: NULL argument");
    if (H == 0 || W == 0 || gp->cell < 1 || gp->num_dilations < 0 || !(gp->scale > 0.0) || !(gp->max_dist > 0.0))
        return gcp_fail(nullptr, -1, "gcp_build_graph: bad map size / cell / scale / max_dist");
    if ((double)gp->cell * gp->scale < gp->max_dist)
        return gcp_fail(nullptr, -1, "gcp_build_graph: grid cell (%d px) smaller than max_dist / scale", gp->cell);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return gcp_fail(nullptr, -2, "gcp_build_graph: no CUDA device; this engine has no CPU path");
    if (device < 0 || device >= ndev) return gcp_fail(nullptr, -1, "gcp_build_graph: bad device %d", device);
    GCP_CUDA(nullptr, cudaSetDevice(device));
    GrDev d{};
    d.ncx = (int)((W + gp->cell - 1) / gp->cell); d.ncy = (int)((H + gp->cell - 1) / gp->cell);
    const size_t ncell = (size_t)d.ncx * d.ncy;
    if (cell_ptr_h[ncell] != N) return gcp_fail(nullptr, -1, "gcp_build_graph: cell_ptr[last] != N");
    uint8_t *occ = nullptr, *dil = nullptr, *amb = nullptr;
    int2* xy = nullptr; uint32_t *cp = nullptr, *ci = nullptr; int32_t* nbr = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc_out = 0;
    auto cleanup = [&]() {
        cudaFree(occ); cudaFree(dil); cudaFree(amb); cudaFree(xy); cudaFree(cp); cudaFree(ci); cudaFree(nbr);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define GR_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                       \
        rc_out = gcp_fail(nullptr, -2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        cleanup(); return rc_out; } } while (0)
    GR_CUDA(up(&occ, occ_h, (size_t)H * W));
    GR_CUDA(cudaMalloc(&dil, (size_t)H * W));
    GR_CUDA(up(&xy, reinterpret_cast<const int2*>(xy_h), (size_t)N));
    GR_CUDA(up(&cp, cell_ptr_h, ncell + 1));
    GR_CUDA(up(&ci, cell_items_h, (size_t)N));
    GR_CUDA(cudaMalloc(&nbr, (size_t)(N ? N : 1) * 8 * sizeof(int32_t)));
    GR_CUDA(cudaMalloc(&amb, N ? N : 1));
    d.occ = occ; d.dil = dil; d.H = H; d.W = W; d.xy = xy; d.N = N; d.cell_ptr = cp; d.cell_items = ci;
    d.nbr = nbr; d.amb = amb; d.gp = *gp;
    GR_CUDA(cudaEventCreate(&e0));
    GR_CUDA(cudaEventCreate(&e1));
    GR_CUDA(cudaEventRecord(e0, 0));
    k_dilate<<<dim3((W + 255) / 256, H), 256>>>(d);
    GR_CUDA(cudaGetLastError());
    if (N) {
        k_graph_candidates<<<(N + 127) / 128, 128>>>(d);
        GR_CUDA(cudaGetLastError());
        k_graph_collide<<<(unsigned)(((size_t)N * 8 * 32 + 255) / 256), 256>>>(d);
        GR_CUDA(cudaGetLastError());
    }
    GR_CUDA(cudaEventRecord(e1, 0));
    GR_CUDA(cudaEventSynchronize(e1));
    if (kernel_ms) GR_CUDA(cudaEventElapsedTime(kernel_ms, e0, e1));
    if (N) {
        GR_CUDA(cudaMemcpy(nbr_h, nbr, (size_t)N * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost));
        GR_CUDA(cudaMemcpy(ambiguous_h, amb, N, cudaMemcpyDeviceToHost));
    }
#undef GR_CUDA
    cleanup();
    return 0;
}
