// motion_kernels.cu - the self-estimated dynamic masks of the flow term for sm_100a
// (reference optimizer.py:156-235 get_motion_mask_from_pairs; ego flow = goem_opt.py:195-236 warp_by_disp).
// C ABI in include/a3r_ga.h.
//
// An "item" is one side of a symmetric pair: the optical flow plane of a slot is compared with the ego flow that the
// pair's two-view cameras (pair_viewer.py:21-78, solved on the host) induce on the slot's frame.  Everything an item
// reads lives on ONE frame's pixel grid, so the pre-step shards by frame window like the hot loop, with no exchange.
//   pass 1  per item: min and max of the error map (ordered-integer atomics; the error is non-negative) -> minmax[item] =
//           (~bits(min), bits(max))
//   pass 2  per local frame and pixel: mean over the frame's items of (err - min) / (max - min) > threshold
// The error map (E x P floats) is recomputed in pass 2 instead of being parked in HBM: 24 B per pixel-item are read twice,
// nothing is written but the mask bytes.
#include <cuda_runtime.h>
#include <math.h>

#include "a3r_ga.h"
#include "a3r_common.cuh"

namespace {

constexpr int CAM = 16;      // floats per item: A (9), b (3), depth row (4)

__device__ __forceinline__ float motion_err(const float4* __restrict__ xw, const float2* __restrict__ fl, const float* __restrict__ c,
                                            int pix, int W) {
    const int v = pix / W, u = pix - v * W;
    const float x = (float)u, y = (float)v;
    const float4 X = xw[pix];
    const float2 F = fl[pix];
    const float depth = c[12] * X.x + c[13] * X.y + c[14] * X.z + c[15];
    const float disp = 1.0f / (depth + 1e-6f);                       // optimizer.py:213: 1 / (depth + 1e-6)
    const float tx = c[0] * x + c[1] * y + c[2] + disp * c[9];       // goem_opt.py:230-231: H (x,y,1) + disp K relt
    const float ty = c[3] * x + c[4] * y + c[5] + disp * c[10];
    const float tz = c[6] * x + c[7] * y + c[8] + disp * c[11];
    const float den = tz + 1e-6f;                                    // goem_opt.py:232
    const float ex = tx / den - x - F.x, ey = ty / den - y - F.y;
    return sqrtf(ex * ex + ey * ey);                                 // optimizer.py:218-219 torch.norm over the 2 channels
}

__global__ void __launch_bounds__(256) k_motion_minmax(const float* __restrict__ predw, const float* __restrict__ flow, int64_t pix_stride,
                                                       const int32_t* __restrict__ item_slot, const int32_t* __restrict__ item_dslot,
                                                       const int32_t* __restrict__ item_hw, const float* __restrict__ item_cam,
                                                       uint32_t* __restrict__ minmax) {
    __shared__ float s_lo[8], s_hi[8];
    const int k = blockIdx.x;
    const int W = item_hw[2 * k + 1], P = item_hw[2 * k] * W;
    const float4* xw = reinterpret_cast<const float4*>(predw) + (int64_t)item_dslot[k] * pix_stride;
    const float2* fl = reinterpret_cast<const float2*>(flow) + (int64_t)item_slot[k] * pix_stride;
    __shared__ float c[CAM];
    if (threadIdx.x < CAM) c[threadIdx.x] = item_cam[(int64_t)k * CAM + threadIdx.x];
    __syncthreads();
    float lo = INFINITY, hi = 0.0f;
    for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix < P; pix += gridDim.y * blockDim.x) {
        const float e = motion_err(xw, fl, c, pix, W);
        lo = fminf(lo, e); hi = fmaxf(hi, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
        // non-negative floats order like their bit patterns; the minimum is kept complemented so that both columns are
        // running maxima over a zero-initialised buffer
        atomicMax(&minmax[2 * k], ~__float_as_uint(lo));
        atomicMax(&minmax[2 * k + 1], __float_as_uint(hi));
    }
}

__global__ void __launch_bounds__(256) k_motion_masks(const float* __restrict__ predw, const float* __restrict__ flow, int64_t pix_stride,
                                                      const int32_t* __restrict__ item_slot, const int32_t* __restrict__ item_dslot,
                                                      const float* __restrict__ item_cam, const uint32_t* __restrict__ minmax,
                                                      const int32_t* __restrict__ frame_item_begin, const int32_t* __restrict__ frame_items,
                                                      const int32_t* __restrict__ lframe_hw, float thre,
                                                      uint8_t* __restrict__ static_mask, uint8_t* __restrict__ dynamic_mask,
                                                      float* __restrict__ mean_out) {
    const int lf = blockIdx.y;
    const int W = lframe_hw[2 * lf + 1], P = lframe_hw[2 * lf] * W;
    const int b = frame_item_begin[lf], e = frame_item_begin[lf + 1];
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        float sum = 0.0f;
        for (int q = b; q < e; ++q) {                 // the reference's stack order: pairs in edge order (optimizer.py:228-235)
            const int k = frame_items[q];
            const float4* xw = reinterpret_cast<const float4*>(predw) + (int64_t)item_dslot[k] * pix_stride;
            const float2* fl = reinterpret_cast<const float2*>(flow) + (int64_t)item_slot[k] * pix_stride;
            const float lo = __uint_as_float(~minmax[2 * k]), hi = __uint_as_float(minmax[2 * k + 1]);
            const float err = motion_err(xw, fl, item_cam + (int64_t)k * CAM, pix, W);
            sum += (err - lo) / (hi - lo);            // optimizer.py:221-222
        }
        const float mean = e > b ? sum / (float)(e - b) : 0.0f;
        const bool dyn = mean > thre;                 // optimizer.py:235
        const int64_t o = (int64_t)lf * pix_stride + pix;
        static_mask[o] = dyn ? 0 : 1;
        dynamic_mask[o] = dyn ? 1 : 0;
        if (mean_out) mean_out[o] = mean;
    }
}

// ---- flow validity maps (reference goem_opt.py:575-618 OccMask, called by get_flow, optimizer.py:146-147) ----
// valid = |(fwd + bwd(p + fwd)).x + (...).y| < th  and  p + fwd inside the image; bwd sampled bilinearly with zero padding
// through torch.grid_sample's align_corners=True coordinate round trip (x -> x / ((W-1)/2) - 1 -> ((g + 1) / 2) (W-1)).
__global__ void __launch_bounds__(256) k_occ_mask(const float* __restrict__ flow, int64_t pix_stride, const int32_t* __restrict__ fwd_slot,
                                                  const int32_t* __restrict__ bwd_slot, int H, int W, float th, uint8_t* __restrict__ valid) {
    const int k = blockIdx.x, P = H * W;
    const float2* fw = reinterpret_cast<const float2*>(flow) + (int64_t)fwd_slot[k] * pix_stride;
    const float2* bw = reinterpret_cast<const float2*>(flow) + (int64_t)bwd_slot[k] * pix_stride;
    const float hx = (float)((double)(W - 1) / 2.0), hy = (float)((double)(H - 1) / 2.0);
    for (int pix = blockIdx.y * blockDim.x + threadIdx.x; pix < P; pix += gridDim.y * blockDim.x) {
        const int v = pix / W, u = pix - v * W;
        const float2 f = fw[pix];
        const float x = (float)u + f.x, y = (float)v + f.y;
        const bool inside = !(x < 0.f || x > (float)(W - 1) || y < 0.f || y > (float)(H - 1));
        const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(__fdiv_rn(x, hx), 1.f), 1.f), 2.f), (float)(W - 1));
        const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(__fdiv_rn(y, hy), 1.f), 1.f), 2.f), (float)(H - 1));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
        float sx = 0.f, sy = 0.f;
        if (fabsf(fx0) < 1e9f && fabsf(fy0) < 1e9f) {               // NaN / huge flows sample nothing
            const int x0 = (int)fx0, y0 = (int)fy0;
            const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};            // nw, ne, sw, se
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) {
                    const float2 b = bw[yy * W + xx];
                    sx += b.x * wgt[t];
                    sy += b.y * wgt[t];
                }
            }
        }
        const float inc = fabsf((sx + f.x) + (sy + f.y));
        valid[(int64_t)k * P + pix] = (inc < th && inside) ? 1 : 0;
    }
}

}  // namespace

extern "C" {

int a3r_occ_masks(const float* flow, int64_t pix_stride, const int32_t* fwd_slot, const int32_t* bwd_slot, int32_t n_items, int32_t H,
                  int32_t W, float th, uint8_t* valid, void* stream) {
    if (n_items < 0 || H < 2 || W < 2 || (int64_t)H * W > pix_stride)
        return a3r_fail(A3R_E_ARG, "a3r_occ_masks: n_items %d, image %dx%d, pix_stride %lld", n_items, H, W, (long long)pix_stride);
    if (n_items == 0) return A3R_OK;
    if (!flow || !fwd_slot || !bwd_slot || !valid) return a3r_fail(A3R_E_ARG, "a3r_occ_masks: NULL argument");
    k_occ_mask<<<dim3((unsigned)n_items, 32), 256, 0, (cudaStream_t)stream>>>(flow, pix_stride, fwd_slot, bwd_slot, H, W, th, valid);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_motion_minmax(const float* predw, const float* flow, int64_t pix_stride, const int32_t* item_slot, const int32_t* item_dslot,
                      const int32_t* item_hw, const float* item_cam, int32_t n_items, uint32_t* minmax, void* stream) {
    if (n_items < 0) return a3r_fail(A3R_E_ARG, "a3r_motion_minmax: n_items %d", n_items);
    if (n_items == 0) return A3R_OK;
    if (!predw || !flow || !item_slot || !item_dslot || !item_hw || !item_cam || !minmax)
        return a3r_fail(A3R_E_ARG, "a3r_motion_minmax: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    A3R_CUDA(cudaMemsetAsync(minmax, 0, sizeof(uint32_t) * 2 * (size_t)n_items, st));     // (~min bits, max bits) = (0, 0)
    k_motion_minmax<<<dim3((unsigned)n_items, 32), 256, 0, st>>>(predw, flow, pix_stride, item_slot, item_dslot, item_hw, item_cam, minmax);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_motion_masks(const float* predw, const float* flow, int64_t pix_stride, const int32_t* item_slot, const int32_t* item_dslot,
                     const float* item_cam, const uint32_t* minmax, const int32_t* frame_item_begin, const int32_t* frame_items,
                     const int32_t* lframe_hw, int32_t n_local_frames, float thre, uint8_t* static_mask, uint8_t* dynamic_mask,
                     float* mean_out, void* stream) {
    if (n_local_frames < 0) return a3r_fail(A3R_E_ARG, "a3r_motion_masks: n_local_frames %d", n_local_frames);
    if (n_local_frames == 0) return A3R_OK;
    if (!predw || !flow || !item_slot || !item_dslot || !item_cam || !minmax || !frame_item_begin || !frame_items || !lframe_hw ||
        !static_mask || !dynamic_mask)
        return a3r_fail(A3R_E_ARG, "a3r_motion_masks: NULL argument");
    k_motion_masks<<<dim3(64, (unsigned)n_local_frames), 256, 0, (cudaStream_t)stream>>>(
        predw, flow, pix_stride, item_slot, item_dslot, item_cam, minmax, frame_item_begin, frame_items, lframe_hw, thre,
        static_mask, dynamic_mask, mean_out);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

}  // extern "C"
