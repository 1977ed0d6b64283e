// mst_kernels.cu - batched solves of the spanning-tree initialisation for sm_100a:
//   weighted similarity (Procrustes/Umeyama) registration between two pointmaps of one frame,
//   Weiszfeld focal estimation, and log-depth initialisation.  C ABI in include/a3r_ga.h.
// All reductions accumulate in fp64 (one-off work; the B200 fp64 pipe is idle otherwise).
#include <cuda_runtime.h>
#include <math.h>

#include "a3r_ga.h"
#include "a3r_common.cuh"

namespace {

__device__ __forceinline__ float conf_untrf(float w, int mode) {   // inverse of reference commons.py:53-64
    switch (mode) {
        case 0: return expf(w);
        case 1: return w * w;
        case 2: return w + 1.0f;
        default: return w;
    }
}

// ------------------------------------------------------------------------------------------
// Procrustes: per problem 17 weighted moment sums -> sums[k][20]
//   [0] sum w, [1..3] sum w x, [4..6] sum w y, [7..15] sum w y x^T (row-major y_r x_c), [16] sum w |x|^2
// ------------------------------------------------------------------------------------------
constexpr int PR_SUMS = 20;
constexpr int PR_USED = 17;

__global__ void __launch_bounds__(256) k_procrustes_moments(const float* __restrict__ predw, int64_t pix_stride,
                                                            const int32_t* __restrict__ xs, const int32_t* __restrict__ ys,
                                                            const int32_t* __restrict__ npix, int conf_mode,
                                                            double* __restrict__ sums) {
    __shared__ double s_red[32];
    const int k = blockIdx.x;
    const int P = npix[k];
    const float4* X = reinterpret_cast<const float4*>(predw) + (int64_t)xs[k] * pix_stride;
    const float4* Y = reinterpret_cast<const float4*>(predw) + (int64_t)ys[k] * pix_stride;
    double a[PR_USED];
#pragma unroll
    for (int i = 0; i < PR_USED; ++i) a[i] = 0.0;
    for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix < P; pix += gridDim.y * blockDim.x) {
        const float4 x = ld_stream4(X + pix), y = ld_stream4(Y + pix);
        const double w = (double)conf_untrf(x.w, conf_mode);
        const double x0 = x.x, x1 = x.y, x2 = x.z, y0 = y.x, y1 = y.y, y2 = y.z;
        a[0] += w;
        a[1] += w * x0; a[2] += w * x1; a[3] += w * x2;
        a[4] += w * y0; a[5] += w * y1; a[6] += w * y2;
        const double wy0 = w * y0, wy1 = w * y1, wy2 = w * y2;
        a[7] += wy0 * x0; a[8] += wy0 * x1; a[9] += wy0 * x2;
        a[10] += wy1 * x0; a[11] += wy1 * x1; a[12] += wy1 * x2;
        a[13] += wy2 * x0; a[14] += wy2 * x1; a[15] += wy2 * x2;
        a[16] += w * (x0 * x0 + x1 * x1 + x2 * x2);
    }
#pragma unroll
    for (int i = 0; i < PR_USED; ++i) {
        const double v = block_sum(a[i], s_red);
        if (threadIdx.x == 0) atomicAdd(&sums[(int64_t)k * PR_SUMS + i], v);
    }
}

// one-sided Jacobi SVD of a 3x3 matrix: M = U diag(sig) V^T, sig sorted descending
__device__ void svd3(const double M[9], double U[9], double sig[3], double V[9]) {
    double A[9];
    for (int i = 0; i < 9; ++i) { A[i] = M[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double alpha = 0, beta = 0, gamma = 0;
            for (int r = 0; r < 3; ++r) {
                alpha += A[r * 3 + p] * A[r * 3 + p];
                beta += A[r * 3 + q] * A[r * 3 + q];
                gamma += A[r * 3 + p] * A[r * 3 + q];
            }
            const double lim = 1e-30 + 1e-16 * sqrt(alpha * beta);
            if (fabs(gamma) <= lim) continue;
            off = fmax(off, fabs(gamma) / sqrt(alpha * beta + 1e-300));
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
            for (int r = 0; r < 3; ++r) {
                const double ap = A[r * 3 + p], aq = A[r * 3 + q];
                A[r * 3 + p] = c * ap - s * aq;
                A[r * 3 + q] = s * ap + c * aq;
                const double vp = V[r * 3 + p], vq = V[r * 3 + q];
                V[r * 3 + p] = c * vp - s * vq;
                V[r * 3 + q] = s * vp + c * vq;
            }
        }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    // sort columns by descending singular value
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2 - a; ++b)
            if (sig[b] < sig[b + 1]) {
                double t = sig[b]; sig[b] = sig[b + 1]; sig[b + 1] = t;
                for (int r = 0; r < 3; ++r) {
                    t = A[r * 3 + b]; A[r * 3 + b] = A[r * 3 + b + 1]; A[r * 3 + b + 1] = t;
                    t = V[r * 3 + b]; V[r * 3 + b] = V[r * 3 + b + 1]; V[r * 3 + b + 1] = t;
                }
            }
    for (int j = 0; j < 3; ++j) {
        const double inv = sig[j] > 1e-300 ? 1.0 / sig[j] : 0.0;
        for (int r = 0; r < 3; ++r) U[r * 3 + j] = A[r * 3 + j] * inv;
    }
    if (sig[2] <= 1e-14 * sig[0]) {   // rank-deficient: complete U with the cross product of its first two columns
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

__device__ __forceinline__ double det3(const double A[9]) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// weighted Umeyama solve from the moment sums (roma.rigid_points_registration semantics, see
// oracle/roma_shim.py): out[k] = s, R (9), t (3)
__global__ void k_procrustes_solve(const double* __restrict__ sums, int count, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const double* a = sums + (int64_t)k * PR_SUMS;
    const double n = a[0];
    double xm[3], ym[3], M[9];
    for (int i = 0; i < 3; ++i) { xm[i] = a[1 + i] / n; ym[i] = a[4 + i] / n; }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[r * 3 + c] = a[7 + r * 3 + c] - n * ym[r] * xm[c];
    const double var = a[16] - n * (xm[0] * xm[0] + xm[1] * xm[1] + xm[2] * xm[2]);
    double U[9], sig[3], V[9];
    svd3(M, U, sig, V);
    const double d = det3(U) * det3(V) < 0 ? -1.0 : 1.0;
    double R[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            R[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + d * U[r * 3 + 2] * V[c * 3 + 2];
    const double s = (sig[0] + sig[1] + d * sig[2]) / var;
    double* o = out + (int64_t)k * 16;
    o[0] = s;
    for (int i = 0; i < 9; ++i) o[1 + i] = R[i];
    for (int r = 0; r < 3; ++r) o[10 + r] = ym[r] - s * (R[r * 3] * xm[0] + R[r * 3 + 1] * xm[1] + R[r * 3 + 2] * xm[2]);
    o[13] = o[14] = o[15] = 0.0;
}

// ------------------------------------------------------------------------------------------
// Weiszfeld focal (reference post_process.py:36-58): pass 0 = closed-form L2, passes 1..10 = IRLS
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_focal_pass(const float* __restrict__ predw, int64_t pix_stride,
                                                    const int32_t* __restrict__ sl, const int32_t* __restrict__ hw,
                                                    const double* __restrict__ focal, int first, double* __restrict__ sums) {
    __shared__ double s_red[32];
    const int k = blockIdx.x;
    const int H = hw[2 * k], W = hw[2 * k + 1], P = H * W;
    const float4* X = reinterpret_cast<const float4*>(predw) + (int64_t)sl[k] * pix_stride;
    const float f = first ? 0.f : (float)focal[k];
    const float ppx = 0.5f * W, ppy = 0.5f * H;
    double num = 0.0, den = 0.0;
    for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix < P; pix += gridDim.y * blockDim.x) {
        const float4 x = ld_stream4(X + pix);
        const int v = pix / W, u = pix - v * W;
        const float pu = (float)u - ppx, pv = (float)v - ppy;
        float ax = x.x / x.z, ay = x.y / x.z;
        if (!isfinite(ax)) ax = 0.f;          // nan_to_num(posinf=0, neginf=0)
        if (!isfinite(ay)) ay = 0.f;
        const float dpx = ax * pu + ay * pv;
        const float dxx = ax * ax + ay * ay;
        float w = 1.f;
        if (!first) {
            const float ex = pu - f * ax, ey = pv - f * ay;
            w = 1.0f / fmaxf(sqrtf(ex * ex + ey * ey), 1e-8f);
        }
        num += (double)(w * dpx);
        den += (double)(w * dxx);
    }
    num = block_sum(num, s_red);
    den = block_sum(den, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[2 * k], num);
        atomicAdd(&sums[2 * k + 1], den);
    }
}

__global__ void k_focal_finish(const double* __restrict__ sums, int count, double* __restrict__ focal) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) focal[k] = sums[2 * k] / sums[2 * k + 1];
}

// ------------------------------------------------------------------------------------------
// log-depth planes (reference init_im_poses.py:138-142, optimizer.py:430-436)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_init_depth(const float* __restrict__ predw, int64_t pix_stride,
                                                    const int32_t* __restrict__ src, const int32_t* __restrict__ npix,
                                                    const double* __restrict__ rows, float* __restrict__ logdepth) {
    const int k = blockIdx.y;
    const int P = npix[k];
    const float4* X = reinterpret_cast<const float4*>(predw) + (int64_t)src[k] * pix_stride;
    const float r0 = (float)rows[4 * k], r1 = (float)rows[4 * k + 1], r2 = (float)rows[4 * k + 2], r3 = (float)rows[4 * k + 3];
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        const float4 x = ld_stream4(X + pix);
        const float d = fmaf(r0, x.x, fmaf(r1, x.y, fmaf(r2, x.z, r3)));
        float l = logf(d);
        if (isnan(l) || l == -INFINITY) l = 0.f;        // depth.log().nan_to_num(neginf=0)
        else if (l == INFINITY) l = 3.4028234663852886e38f;
        logdepth[(int64_t)k * pix_stride + pix] = l;
    }
}


// ---- RANSAC-PnP: the O(points) half of one hypothesis test ----------------------------------------------------------
// cv::solvePnPRansac (reference init_im_poses.py:498: the frames the spanning tree leaves without a pose; two-view cameras of
// the self-estimated masks; init='known_poses') spends its time re-projecting ALL points for every hypothesis
// (PnPRansacCallback::computeError -> projectPoints, ~8 ms per hypothesis on 147 k points).  The sampling, the 5-point EPnP
// solves and the final SQPNP refit stay on the host, in OpenCV's own routines (cloud_opt_flow/pnp_ransac.py); this kernel is
// the inlier test of one hypothesis, operation by operation what OpenCV computes - fp64 products and sums in its order with
// no contraction into FMAs, the projection rounded to fp32, the squared error in fp32, compared against the squared
// threshold in fp64 - so that the inlier sets (and with them the refit) are the ones cv::solvePnPRansac finds.
struct PnpHyp { double R[9], t[3], fx, fy, cx, cy, thr2; };

__global__ void __launch_bounds__(256) k_pnp_inliers(const float* __restrict__ pts, const uint8_t* __restrict__ mask, int W,
                                                     int64_t P, const PnpHyp h, uint8_t* __restrict__ flags,
                                                     int* __restrict__ count) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (p < P && mask[p]) {
        const double X = (double)pts[3 * p], Y = (double)pts[3 * p + 1], Z = (double)pts[3 * p + 2];
        const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h.R[0], X), __dmul_rn(h.R[1], Y)), __dmul_rn(h.R[2], Z)), h.t[0]);
        const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h.R[3], X), __dmul_rn(h.R[4], Y)), __dmul_rn(h.R[5], Z)), h.t[1]);
        const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h.R[6], X), __dmul_rn(h.R[7], Y)), __dmul_rn(h.R[8], Z)), h.t[2]);
        const double iz = z != 0.0 ? __ddiv_rn(1.0, z) : 1.0;
        const float pu = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(x, iz), h.fx), h.cx));
        const float pv = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(y, iz), h.fy), h.cy));
        const int v = (int)(p / W), u = (int)(p - (int64_t)v * W);
        const float dx = __fsub_rn((float)u, pu), dy = __fsub_rn((float)v, pv);
        const float err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        in = (double)err <= h.thr2;
    }
    if (flags && p < P) flags[p] = in ? 1 : 0;
    const int n = __syncthreads_count(in);
    if (threadIdx.x == 0 && n) atomicAdd(count, n);
}

// world points of a frame from its raw slot plane: pts = X * A[:,0] + Y * A[:,1] + Z * A[:,2] + t in fp32, every product and sum
// rounded on its own, in the order numpy evaluates the same expression on the host (init_im_poses.py: the similarity of the
// spanning tree applied next to the PnP call that consumes it) - the RANSAC-PnP sees the same fp32 points either way
struct SimF { float A[9], t[3]; };
__global__ void __launch_bounds__(256) k_similarity_points(const float4* __restrict__ plane, int64_t P, const SimF s,
                                                           float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float4 x = plane[p];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out[3 * p + c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x.x, s.A[3 * c]), __fmul_rn(x.y, s.A[3 * c + 1])),
                                             __fmul_rn(x.z, s.A[3 * c + 2])), s.t[c]);
}

// the sampled points of one hypothesis: row idx[k] of the solver's list = pixel sel[idx[k]] -> (x, y, z, pixel)
struct Idx64 { int32_t v[64]; };
__global__ void k_pnp_gather(const float* __restrict__ pts, const int64_t* __restrict__ sel, const Idx64 idx, int n,
                             float* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= n) return;
    const int64_t pix = sel[idx.v[k]];
    out[4 * k] = pts[3 * pix]; out[4 * k + 1] = pts[3 * pix + 1]; out[4 * k + 2] = pts[3 * pix + 2];
    out[4 * k + 3] = (float)pix;           // < 2^24 (plan.py): exact
}

}  // namespace

// confidence clean-up of frame i against all other frames (reference base_opt.py:587-623)
// pts3d holds the frames [i0, ...) only (row i - i0 is frame i): a frame-window shard cleans its own frames against the
// depth maps / confidences of ALL frames (gathered by the caller)
__global__ void __launch_bounds__(256) k_clean_frame(const float* __restrict__ pts3d, const float* __restrict__ depth,
                                                     const float* __restrict__ w2c, const float* __restrict__ intr,
                                                     const int32_t* __restrict__ hw, int n, int i, int i0, int64_t pix_stride,
                                                     float tol, float bad_conf, float* conf) {
    const int P = hw[2 * i] * hw[2 * i + 1];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gridDim.x * blockDim.x) {
        const float* X = pts3d + ((int64_t)(i - i0) * pix_stride + q) * 3;
        const float x = X[0], y = X[1], z = X[2];
        float c = conf[(int64_t)i * pix_stride + q];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const float* M = w2c + 16 * j;
            // geotrf(cams[j], pts3d): R x + t, then geotrf(K[j], ., norm=1, ncol=2): (K p)_xy / (K p)_z, rounded half-to-even
            const float px = M[0] * x + M[1] * y + M[2] * z + M[3];
            const float py = M[4] * x + M[5] * y + M[6] * z + M[7];
            const float pz = M[8] * x + M[9] * y + M[10] * z + M[11];
            const float* K = intr + 4 * j;
            const float uu = (K[0] * px + K[2] * pz) / pz, vv = (K[1] * py + K[3] * pz) / pz;
            const long long u = (long long)rintf(uu), v = (long long)rintf(vv);
            const int Hj = hw[2 * j], Wj = hw[2 * j + 1];
            if (!(pz > 0.f) || u < 0 || u >= Wj || v < 0 || v >= Hj) continue;       // NaN coordinates fail the range test
            const int64_t o = (int64_t)j * pix_stride + v * Wj + u;
            if (pz < (1.0f - tol) * depth[o] && c < conf[o]) c = fminf(c, bad_conf);
        }
        conf[(int64_t)i * pix_stride + q] = c;
    }
}

extern "C" {

int a3r_clean_pointcloud_range(const float* pts3d, const float* depth, const float* w2c, const float* intr, const int32_t* hw,
                               int32_t n, int32_t i0, int32_t i1, int64_t pix_stride, float tol, float bad_conf, float* conf,
                               void* stream) {
    if (!pts3d || !depth || !w2c || !intr || !hw || !conf || n <= 0 || i0 < 0 || i1 > n || i0 > i1)
        return a3r_fail(A3R_E_ARG, "a3r_clean_pointcloud: bad arguments");
    if (!(tol >= 0.f && tol < 1.f)) return a3r_fail(A3R_E_ARG, "a3r_clean_pointcloud: tol must be in [0, 1)");
    for (int i = i0; i < i1; ++i)        // frame i reads the confidences frames < i ended with: one launch per frame, in order
        k_clean_frame<<<592, 256, 0, (cudaStream_t)stream>>>(pts3d, depth, w2c, intr, hw, n, i, i0, pix_stride, tol, bad_conf, conf);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_clean_pointcloud(const float* pts3d, const float* depth, const float* w2c, const float* intr, const int32_t* hw,
                         int32_t n, int64_t pix_stride, float tol, float bad_conf, float* conf, void* stream) {
    return a3r_clean_pointcloud_range(pts3d, depth, w2c, intr, hw, n, 0, n, pix_stride, tol, bad_conf, conf, stream);
}


int a3r_procrustes_batched(const float* predw, int64_t pix_stride, const int32_t* xs, const int32_t* ys,
                           const int32_t* npix, int32_t count, int32_t conf_mode, double* sums, double* out,
                           void* stream) {
    if (count <= 0) return A3R_OK;
    if (!predw || !xs || !ys || !npix || !sums || !out) return a3r_fail(A3R_E_ARG, "a3r_procrustes_batched: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    A3R_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * PR_SUMS * (size_t)count, st));
    const int chunks = count >= 592 ? 8 : (count >= 148 ? 16 : 64);
    k_procrustes_moments<<<dim3(count, chunks), 256, 0, st>>>(predw, pix_stride, xs, ys, npix, conf_mode, sums);
    A3R_CUDA(cudaGetLastError());
    k_procrustes_solve<<<(count + 63) / 64, 64, 0, st>>>(sums, count, out);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_weiszfeld_focal_batched(const float* predw, int64_t pix_stride, const int32_t* sl, const int32_t* hw,
                                int32_t count, double* sums, double* out_focal, void* stream) {
    if (count <= 0) return A3R_OK;
    if (!predw || !sl || !hw || !sums || !out_focal) return a3r_fail(A3R_E_ARG, "a3r_weiszfeld_focal_batched: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = count >= 592 ? 8 : (count >= 148 ? 16 : 64);
    for (int pass = 0; pass <= 10; ++pass) {
        A3R_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)count, st));
        k_focal_pass<<<dim3(count, chunks), 256, 0, st>>>(predw, pix_stride, sl, hw, out_focal, pass == 0, sums);
        A3R_CUDA(cudaGetLastError());
        k_focal_finish<<<(count + 127) / 128, 128, 0, st>>>(sums, count, out_focal);
        A3R_CUDA(cudaGetLastError());
    }
    return A3R_OK;
}

int a3r_init_depth(const float* predw, int64_t pix_stride, const int32_t* src, const int32_t* npix,
                   const double* rows, int32_t n_local_frames, float* logdepth, void* stream) {
    if (n_local_frames <= 0) return A3R_OK;
    if (!predw || !src || !npix || !rows || !logdepth) return a3r_fail(A3R_E_ARG, "a3r_init_depth: NULL pointer");
    k_init_depth<<<dim3(64, n_local_frames), 256, 0, (cudaStream_t)stream>>>(predw, pix_stride, src, npix, rows, logdepth);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_pnp_inliers(const float* pts, const uint8_t* mask, int32_t W, int64_t P, const double* hyp, uint8_t* flags,
                    int32_t* count, void* stream) {
    if (P <= 0) return A3R_OK;
    if (!pts || !mask || !hyp || !count || W <= 0) return a3r_fail(A3R_E_ARG, "a3r_pnp_inliers: NULL pointer or W <= 0");
    cudaStream_t st = (cudaStream_t)stream;
    PnpHyp h;
    for (int k = 0; k < 9; ++k) h.R[k] = hyp[k];
    for (int k = 0; k < 3; ++k) h.t[k] = hyp[9 + k];
    h.fx = hyp[12]; h.fy = hyp[13]; h.cx = hyp[14]; h.cy = hyp[15]; h.thr2 = hyp[16];
    A3R_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    k_pnp_inliers<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pts, mask, W, P, h, flags, count);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_similarity_points(const float* plane, int64_t P, const float* sim, float* out, void* stream) {
    if (P <= 0) return A3R_OK;
    if (!plane || !sim || !out) return a3r_fail(A3R_E_ARG, "a3r_similarity_points: NULL pointer");
    SimF s;
    for (int k = 0; k < 9; ++k) s.A[k] = sim[k];
    for (int k = 0; k < 3; ++k) s.t[k] = sim[9 + k];
    k_similarity_points<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(plane), P, s, out);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_pnp_gather(const float* pts, const int64_t* sel, const int32_t* idx, int32_t n, float* out, void* stream) {
    if (n <= 0) return A3R_OK;
    if (!pts || !sel || !idx || !out || n > 64) return a3r_fail(A3R_E_ARG, "a3r_pnp_gather: NULL pointer or more than 64 rows");
    Idx64 v;
    for (int k = 0; k < 64; ++k) v.v[k] = k < n ? idx[k] : 0;
    k_pnp_gather<<<1, 64, 0, (cudaStream_t)stream>>>(pts, sel, v, n, out);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

}  // extern "C"
