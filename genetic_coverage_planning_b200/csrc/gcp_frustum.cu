// Setup path N1: Mappy.computeFrustums (mappy.py:274-328) on the device.
//
// For every directed edge (from, to) the reference scans ALL obstacle pixels of the map, keeps
// those closer than max_view to the source waypoint and inside the camera's field of view when
// looking along the edge, bins them into num_rays rays (np.digitize) and keeps the nearest hit
// per ray; the ray end points plus two near points, rotated into map axes and rounded, are the
// 17-point footprint polygon that cv2.fillPoly rasterises later (mappy.py:353-355).
//
// Here: one CTA per source waypoint.  Only the (2*window+1)^2 pixels around the waypoint can pass
// the max_view filter, so the CTA gathers them once into shared memory as {distance, bearing}
// and every out-edge of the waypoint re-bins that list.  All decisions the reference takes in
// floating point are taken on the same fp64 values here:
//   * distances: (o*scale - w*scale), squares, sum, sqrt - all correctly rounded in both
//     implementations (no FMA contraction: explicit __dmul_rn / __dadd_rn)
//   * bearings: atan2 is NOT correctly rounded (CUDA <= 2 ulp, libm <= 1 ulp), so every
//     comparison of a relative bearing against a ray boundary or the field-of-view limit that
//     falls within eps_angle marks the edge AMBIGUOUS; so does a rotated coordinate within
//     eps_round of a rounding boundary (numpy's 2x2 dot may or may not fuse the multiply-add).
//     The host recomputes ambiguous edges with the reference's own numpy expressions
//     (frustums.py); on the worlds of this repo none are flagged.
//   * per-edge heading, its cosine / sine, the ray angles and their sines / cosines are computed
//     on the host by numpy (same expressions as the reference) and passed in.
#include "gcp_common.cuh"
#include <cmath>

constexpr int FR_THREADS = 256;
constexpr int FR_MAX_RAYS = 62;

struct FrDev {
    const uint8_t* occ; uint32_t H, W;
    const int2* xy; const uint32_t* row_ptr; const uint32_t* col; uint32_t N;
    const double* ray_angles; const double* ray_sin; const double* ray_cos;
    const double* e_theta; const double* e_cos; const double* e_sin;
    int2* poly; uint8_t* amb;
    gcp_frustum_params fp;
    uint32_t cap;
};

__global__ void __launch_bounds__(FR_THREADS)
k_frustums(FrDev d)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double2* s_list = reinterpret_cast<double2*>(smem);            // [cap] {distance, bearing}
    __shared__ unsigned long long s_min[FR_MAX_RAYS + 2];
    __shared__ double s_ray[FR_MAX_RAYS + 2];
    __shared__ uint32_t s_n, s_amb;
    const uint32_t frm = blockIdx.x;
    if (frm >= d.N) return;
    const uint32_t lo = d.row_ptr[frm], hi = d.row_ptr[frm + 1];
    if (lo == hi) return;
    const int tid = threadIdx.x;
    const int R = d.fp.num_rays;
    if (tid == 0) s_n = 0;
    if (tid < R) s_ray[tid] = d.ray_angles[tid];
    __syncthreads();

    // ---- gather: obstacle pixels of the window closer than max_view (mappy.py:278,290-296)
    const int2 w = d.xy[frm];
    const double scale = d.fp.scale;
    const double wr = __dmul_rn((double)w.x, scale), wc = __dmul_rn((double)w.y, scale);
    const int win = d.fp.window, side = 2 * win + 1;
    for (int p = tid; p < side * side; p += FR_THREADS) {
        const int rr = w.x - win + p / side, cc = w.y - win + p % side;
        if (rr < 0 || cc < 0 || rr >= (int)d.H || cc >= (int)d.W) continue;
        if (!d.occ[(size_t)rr * d.W + cc]) continue;
        const double d0 = __dsub_rn(__dmul_rn((double)rr, scale), wr);
        const double d1 = __dsub_rn(__dmul_rn((double)cc, scale), wc);
        const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
        if (!(dist < d.fp.max_view)) continue;
        const uint32_t k = atomicAdd(&s_n, 1u);
        if (k < d.cap) s_list[k] = make_double2(dist, atan2(d1, d0));
    }
    __syncthreads();
    const uint32_t n = min(s_n, d.cap);
    const bool overflow = s_n > d.cap;                              // cannot happen: cap = lattice points in the disc
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    const double half = d.fp.half_view, eps = d.fp.eps_angle;

    for (uint32_t e = lo; e < hi; ++e) {
        if (tid < R) s_min[tid] = 0x7FF0000000000000ull;            // +inf
        if (tid == 0) s_amb = overflow ? 1u : 0u;
        __syncthreads();
        const double theta = d.e_theta[e];
        uint32_t amb = 0;
        for (uint32_t i = tid; i < n; i += FR_THREADS) {
            const double2 it = s_list[i];
            double rel = __dsub_rn(it.y, theta);                                 // :298
            if (rel < -PI) rel = __dadd_rn(rel, TWO_PI);                         // :300
            if (rel > PI) rel = __dsub_rn(rel, TWO_PI);                          // :301
            if (fabs(rel - half) < eps || fabs(rel + half) < eps) amb = 1;
            if (!(rel < half && rel > -half)) continue;                          // :303
            int bin = 0;                                                         // np.digitize(rel, ray_angles), :306
            for (int k = 0; k < R; ++k) {
                const double b = s_ray[k];
                bin += (b <= rel);
                if (k > 0 && fabs(rel - b) < eps) amb = 1;
            }
            if (bin >= 1) atomicMin(&s_min[bin - 1], (unsigned long long)__double_as_longlong(it.x));   // :310
        }
        if (amb) atomicOr(&s_amb, 1u);
        __syncthreads();
        if (tid < R + 2) {
            double p0, p1;
            if (tid < R) {
                const unsigned long long m = s_min[tid];
                const double z0 = (m == 0x7FF0000000000000ull) ? d.fp.max_view_px
                                                                : __ddiv_rn(__longlong_as_double((long long)m), scale);
                p0 = __dmul_rn(d.ray_sin[tid], z0);                              // :314
                p1 = __dmul_rn(d.ray_cos[tid], z0);
            } else if (tid == R) { p0 = d.fp.near_last[0]; p1 = d.fp.near_last[1]; }   // :315-316
            else { p0 = d.fp.near_first[0]; p1 = d.fp.near_first[1]; }                 // :317-318
            const double c = d.e_cos[e], s = d.e_sin[e];
            // Rot = getRot2D(theta).T = [[c, s], [-s, c]]  (gori_tools.py:184-186, mappy.py:319-320)
            const double x = __dadd_rn(__dmul_rn(c, p0), __dmul_rn(s, p1));
            const double y = __dadd_rn(__dmul_rn(-s, p0), __dmul_rn(c, p1));
            const double fx = fabs(x - floor(x) - 0.5), fy = fabs(y - floor(y) - 0.5);
            if (fx < d.fp.eps_round || fy < d.fp.eps_round) atomicOr(&s_amb, 1u);
            d.poly[(size_t)e * (R + 2) + tid] = make_int2((int)rint(x), (int)rint(y));   // np.around: half to even
        }
        __syncthreads();
        if (tid == 0) d.amb[e] = (uint8_t)s_amb;
    }
}

template <class T>
static cudaError_t up(T** dst, const T* src, size_t n)
{
    cudaError_t e = cudaMalloc(dst, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return n ? cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
}

extern "C" int gcp_compute_frustums(int device, const uint8_t* occ_h, uint32_t H, uint32_t W,
                                    const int32_t* xy_h, const uint32_t* row_ptr_h, const uint32_t* col_h,
                                    uint32_t N, uint32_t E, const gcp_frustum_params* fp,
                                    const double* ray_angles_h, const double* ray_sin_h, const double* ray_cos_h,
                                    const double* edge_theta_h, const double* edge_cos_h, const double* edge_sin_h,
                                    int32_t* poly_h, uint8_t* ambiguous_h, float* kernel_ms)
{
    if (!occ_h || !xy_h || !row_ptr_h || !col_h || !fp || !ray_angles_h || !ray_sin_h || !ray_cos_h ||
        !edge_theta_h || !edge_cos_h || !edge_sin_h || !poly_h || !ambiguous_h)
        return gcp_fail(nullptr, -1, "gcp_compute_frustums: NULL argument");
    if (fp->num_rays < 1 || fp->num_rays > FR_MAX_RAYS)
        return gcp_fail(nullptr, -1, "gcp_compute_frustums: num_rays %d not in [1, %d]", fp->num_rays, FR_MAX_RAYS);
    if (fp->window < 1 || !(fp->scale > 0.0) || !(fp->max_view > 0.0))
        return gcp_fail(nullptr, -1, "gcp_compute_frustums: bad window / scale / max_view");
    if (!(fp->half_view > 0.0 && fp->half_view < 3.0))
        return gcp_fail(nullptr, -1, "gcp_compute_frustums: need 0 < view_angle / 2 < 3 rad");
    if (row_ptr_h[N] != E) return gcp_fail(nullptr, -1, "gcp_compute_frustums: row_ptr[N] != E");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return gcp_fail(nullptr, -2, "gcp_compute_frustums: no CUDA device; this engine has no CPU path");
    if (device < 0 || device >= ndev) return gcp_fail(nullptr, -1, "gcp_compute_frustums: bad device %d", device);
    GCP_CUDA(nullptr, cudaSetDevice(device));

    // capacity of the per-waypoint list: lattice points of the window inside the max_view disc
    // (one pixel of slack for the rounding of o*scale - w*scale)
    uint32_t cap = 0;
    {
        const double r = fp->max_view / fp->scale + 1.0;
        for (int a = -fp->window; a <= fp->window; ++a)
            for (int b = -fp->window; b <= fp->window; ++b)
                cap += ((double)a * a + (double)b * b <= r * r);
    }
    int max_smem = 0;
    GCP_CUDA(nullptr, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t smem = (size_t)cap * sizeof(double2);
    if (smem + 2048 > (size_t)max_smem)
        return gcp_fail(nullptr, -3, "gcp_compute_frustums: %u obstacle pixels per view disc need %zu B of shared memory (max %d)",
                        cap, smem, max_smem);

    FrDev d{};
    uint8_t* occ = nullptr; int2* xy = nullptr; uint32_t *rp = nullptr, *cl = nullptr;
    double *ra = nullptr, *rs = nullptr, *rc = nullptr, *et = nullptr, *ec = nullptr, *es = nullptr;
    int2* poly = nullptr; uint8_t* amb = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc_out = 0;
    auto cleanup = [&]() {
        cudaFree(occ); cudaFree(xy); cudaFree(rp); cudaFree(cl); cudaFree(ra); cudaFree(rs); cudaFree(rc);
        cudaFree(et); cudaFree(ec); cudaFree(es); cudaFree(poly); cudaFree(amb);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define FR_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                       \
        rc_out = gcp_fail(nullptr, -2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        cleanup(); return rc_out; } } while (0)
    const size_t npoly = (size_t)E * (fp->num_rays + 2);
    FR_CUDA(up(&occ, occ_h, (size_t)H * W));
    FR_CUDA(up(&xy, reinterpret_cast<const int2*>(xy_h), (size_t)N));
    FR_CUDA(up(&rp, row_ptr_h, (size_t)N + 1));
    FR_CUDA(up(&cl, col_h, (size_t)E));
    FR_CUDA(up(&ra, ray_angles_h, (size_t)fp->num_rays));
    FR_CUDA(up(&rs, ray_sin_h, (size_t)fp->num_rays));
    FR_CUDA(up(&rc, ray_cos_h, (size_t)fp->num_rays));
    FR_CUDA(up(&et, edge_theta_h, (size_t)E));
    FR_CUDA(up(&ec, edge_cos_h, (size_t)E));
    FR_CUDA(up(&es, edge_sin_h, (size_t)E));
    FR_CUDA(cudaMalloc(&poly, (npoly ? npoly : 1) * sizeof(int2)));
    FR_CUDA(cudaMalloc(&amb, E ? E : 1));
    FR_CUDA(cudaMemset(amb, 0, E ? E : 1));
    d.occ = occ; d.H = H; d.W = W; d.xy = xy; d.row_ptr = rp; d.col = cl; d.N = N;
    d.ray_angles = ra; d.ray_sin = rs; d.ray_cos = rc; d.e_theta = et; d.e_cos = ec; d.e_sin = es;
    d.poly = poly; d.amb = amb; d.fp = *fp; d.cap = cap;
    FR_CUDA(cudaFuncSetAttribute(k_frustums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FR_CUDA(cudaEventCreate(&e0));
    FR_CUDA(cudaEventCreate(&e1));
    FR_CUDA(cudaEventRecord(e0, 0));
    if (N) k_frustums<<<N, FR_THREADS, smem, 0>>>(d);
    FR_CUDA(cudaGetLastError());
    FR_CUDA(cudaEventRecord(e1, 0));
    FR_CUDA(cudaEventSynchronize(e1));
    if (kernel_ms) FR_CUDA(cudaEventElapsedTime(kernel_ms, e0, e1));
    if (npoly) FR_CUDA(cudaMemcpy(poly_h, poly, npoly * sizeof(int2), cudaMemcpyDeviceToHost));
    if (E) FR_CUDA(cudaMemcpy(ambiguous_h, amb, E, cudaMemcpyDeviceToHost));
#undef FR_CUDA
    cleanup();
    return 0;
}
