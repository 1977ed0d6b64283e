// ga_kernels.cu - hot loop of the global alignment for sm_100a: parameter maps, the fused
// forward+backward scan over pixel-slots, the small-parameter chain + Adam, the depth Adam.
// C ABI in include/a3r_ga.h.  No tensor cores on purpose: the scan is HBM/issue bound.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "a3r_ga.h"
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include "ga_scan_v4.cuh"
#include <stdlib.h>

using namespace a3r;

// ==========================================================================================
// K0: parameter maps  (reference base_opt.py:195-247, optimizer.py:398-428)
// ==========================================================================================
// body shared by the stand-alone launch and the tail of the fused update kernel (256 threads per CTA in both, so the
// mean of the pair log-scales is the same number, bit for bit, wherever it is evaluated)
__device__ __forceinline__ void prepare_body(const A3rProblem& p, const int gid, const int gstride, double* s_red, double* s_mean) {
    const int tid = threadIdx.x;
    const int E = p.n_edges, N = p.n_frames;
    const bool shared_focal = p.flags & A3R_F_SHARED_FOCAL;
    const bool norm_scale = p.flags & A3R_F_NORM_PW_SCALE;

    // every CTA recomputes mean_e(log scale) in the same order -> identical on all CTAs
    double part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) part += (double)p.pw_poses[8 * e + 7];
    part = block_sum(part, s_red);
    if (tid == 0) *s_mean = part / (double)E;
    __syncthreads();
    const double log_norm = norm_scale ? (log((double)p.base_scale) - *s_mean) : 0.0;

    for (int n = gid; n < N; n += gstride) {
        FrameD F;
        load_frame(p.im_poses, p.im_focals, p.im_pp, p.frame_hw, n, shared_focal, (double)p.focal_break, F);
        derive_frame_par(F, p.frame_par + FP_STRIDE * n);
    }
    for (int e = gid; e < E; e += gstride)
        derive_edge_par(p.pw_poses + 8 * e, p.pw_adaptors + 2 * e, log_norm, norm_scale, (double)p.pw_break, p.edge_par + EP_STRIDE * e);
    if (p.flags & A3R_F_HAS_FLOW) {
        for (int gs = gid; gs < 2 * E; gs += gstride) {
            const int e = gs >> 1, side = gs & 1;
            const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
            FrameD S, T;
            load_frame(p.im_poses, p.im_focals, p.im_pp, p.frame_hw, side ? j : i, shared_focal, (double)p.focal_break, S);
            load_frame(p.im_poses, p.im_focals, p.im_pp, p.frame_hw, side ? i : j, shared_focal, (double)p.focal_break, T);
            derive_slot_par(S, T, p.slot_par + SP_STRIDE * gs);
        }
    }
}

__global__ void __launch_bounds__(256) k_prepare(const A3rProblem p) {
    __shared__ double s_red[32];
    __shared__ double s_mean;
    prepare_body(p, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, s_red, &s_mean);
}

// ==========================================================================================
// K1 (v1): fused forward + backward over pixel-slots, frame-major, thread-owns-pixels.
//   CTA = 256 threads, tile = 1024 pixels of one work frame (4 pixels per thread, strided by 256 so
//   every warp load is 32 consecutive float4 = 512 B).  Per slot: 27 per-thread partial sums ->
//   warp shuffle reduce -> per-warp fp64 shared accumulators -> fp64 global atomics once per
//   (CTA, work frame).
// ==========================================================================================
constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_PPT = 4;                       // pixels per thread per tile
constexpr int K1_TILE = K1_THREADS * K1_PPT;    // 1024
constexpr int K1_MAXS = 32;                     // slots per work frame

struct K1Shared {
    SlotConst slot[K1_MAXS];
    int gslot[K1_MAXS];
    float fp[FP_STRIDE];
    double fred[K1_WARPS][FACC_N];
};

template <bool FLOW, bool L2D>
__global__ void __launch_bounds__(K1_THREADS, 2) k_accumulate_v1(const A3rProblem p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_wacc = reinterpret_cast<double*>(smem_raw);                      // [K1_WARPS][K1_MAXS][ACC_STRIDE]
    K1Shared& sh = *reinterpret_cast<K1Shared*>(smem_raw + sizeof(double) * K1_WARPS * K1_MAXS * ACC_STRIDE);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = p.cta_tile_begin[blockIdx.x], t1 = p.cta_tile_begin[blockIdx.x + 1];
    if (t0 >= t1) return;

    int wf = 0;
    while (t0 >= p.wf_tile_begin[wf + 1]) ++wf;
    const float thre_eff = flow_thre_eff(p.pxl_thre);
    int cur_wf = -1, lf = 0, n = 0, P = 0, W = 1, deg = 0, s_begin = 0, accum = 0;
    float facc[FACC_N];
#pragma unroll
    for (int k = 0; k < FACC_N; ++k) facc[k] = 0.f;

    auto flush = [&]() {
        __syncthreads();
        for (int idx = tid; idx < deg * ACC_N; idx += K1_THREADS) {
            const int s = idx / ACC_N, i = idx - s * ACC_N;
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < K1_WARPS; ++w) sum += s_wacc[(w * K1_MAXS + s) * ACC_STRIDE + i];
            if (sum != 0.0) atomicAdd(&p.acc_slot[(int64_t)sh.gslot[s] * ACC_STRIDE + i], sum);
        }
#pragma unroll
        for (int k = 0; k < FACC_N; ++k) {
            const float v = warp_sum(facc[k]);
            if (lane == 0) sh.fred[warp][k] = (double)v;
            facc[k] = 0.f;
        }
        __syncthreads();
        if (tid < FACC_N) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < K1_WARPS; ++w) sum += sh.fred[w][tid];
            atomicAdd(&p.acc_frame[(int64_t)n * FACC_N + tid], sum);
        }
    };

    for (int tile = t0; tile < t1; ++tile) {
        while (tile >= p.wf_tile_begin[wf + 1]) ++wf;
        if (wf != cur_wf) {
            if (cur_wf >= 0) flush();
            cur_wf = wf;
            lf = p.wf_lframe[wf];
            n = p.lframe_id[lf];
            W = p.frame_hw[2 * n + 1];
            P = p.frame_hw[2 * n] * W;
            s_begin = p.wf_slot_begin[wf];
            deg = p.wf_slot_begin[wf + 1] - s_begin;
            accum = p.wf_accum[wf];
            __syncthreads();
            if (tid < FP_STRIDE) sh.fp[tid] = p.frame_par[FP_STRIDE * n + tid];
            if (tid < deg) {
                const int gs = p.lslot_gslot[s_begin + tid];
                sh.gslot[tid] = gs;
                SlotConst& c = sh.slot[tid];
                const float* ep = p.edge_par + EP_STRIDE * (gs >> 1);
#pragma unroll
                for (int k = 0; k < 9; ++k) c.nM[k] = -ep[EP_M + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) c.st[k] = ep[EP_ST + k];
                if (FLOW) {
                    const float* sp = p.slot_par + SP_STRIDE * gs;
#pragma unroll
                    for (int k = 0; k < 9; ++k) c.A[k] = sp[SP_A + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) c.b[k] = sp[SP_B + k];
                }
                c.wn = (gs & 1) ? p.inv_area_j : p.inv_area_i;
            }
            for (int idx = tid; idx < K1_WARPS * K1_MAXS * ACC_STRIDE; idx += K1_THREADS) s_wacc[idx] = 0.0;
            __syncthreads();
        }
        const int p0 = (tile - p.wf_tile_begin[wf]) * K1_TILE;
        const int64_t pix_base = (int64_t)lf * p.pix_stride;

        // ---- per-pixel state ----
        float zx[K1_PPT], zy[K1_PPT], zz[K1_PPT], pu[K1_PPT], pv[K1_PPT], disp[K1_PPT];
        float gsx[K1_PPT], gsy[K1_PPT], gsz[K1_PPT], gfi[K1_PPT], gfj[K1_PPT];
        bool valid[K1_PPT], stat[K1_PPT];
#pragma unroll
        for (int j = 0; j < K1_PPT; ++j) {
            const int pix = p0 + tid + j * K1_THREADS;
            valid[j] = pix < P;
            const float ld = valid[j] ? p.logdepth[pix_base + pix] : 0.f;
            const int v = pix / W, u = pix - v * W;
            PixState s;
            pixel_state(sh.fp, ld, u, v, s);
            zx[j] = s.zx; zy[j] = s.zy; zz[j] = s.zz; pu[j] = s.u; pv[j] = s.v; disp[j] = s.disp;
            gsx[j] = gsy[j] = gsz[j] = gfi[j] = gfj[j] = 0.f;
            stat[j] = false;
            if (FLOW) stat[j] = valid[j] && (p.static_mask[pix_base + pix] != 0);
        }

        // ---- slots ----
        for (int s = 0; s < deg; ++s) {
            const SlotConst c = sh.slot[s];
            const bool side_j = sh.gslot[s] & 1;
            const float4* xw = reinterpret_cast<const float4*>(p.predw) + (int64_t)(s_begin + s) * p.pix_stride + p0 + tid;
            float4 X[K1_PPT];
            float2 F[K1_PPT];
#pragma unroll
            for (int j = 0; j < K1_PPT; ++j) X[j] = valid[j] ? ld_stream4(xw + j * K1_THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (FLOW) {
                const float2* fl = reinterpret_cast<const float2*>(p.flow) + (int64_t)(s_begin + s) * p.pix_stride + p0 + tid;
#pragma unroll
                for (int j = 0; j < K1_PPT; ++j) F[j] = stat[j] ? ld_stream2(fl + j * K1_THREADS) : make_float2(0.f, 0.f);
            }
            float acc[ACC_N];
#pragma unroll
            for (int k = 0; k < ACC_N; ++k) acc[k] = 0.f;
#pragma unroll
            for (int j = 0; j < K1_PPT; ++j) {
                PixState st;
                st.zx = zx[j]; st.zy = zy[j]; st.zz = zz[j]; st.u = pu[j]; st.v = pv[j]; st.disp = disp[j];
                float gx, gy, gz;
                align_px<L2D>(c, X[j].x, X[j].y, X[j].z, X[j].w, st, acc, gx, gy, gz);
                gsx[j] += gx; gsy[j] += gy; gsz[j] += gz;
                if (FLOW) {
                    if (stat[j]) {
                        const float gd = flow_px(c, st, F[j].x, F[j].y, thre_eff, true, acc);
                        if (side_j) gfj[j] += gd; else gfi[j] += gd;
                    }
                }
            }
            double* wa = s_wacc + (warp * K1_MAXS + s) * ACC_STRIDE;
#pragma unroll
            for (int k = 0; k < (FLOW ? ACC_N : ACC_GA); ++k) {
                const float v = warp_sum(acc[k]);
                if (lane == 0) wa[k] += (double)v;
            }
        }

        // ---- per-pixel gradient planes + per-frame sums ----
#pragma unroll
        for (int j = 0; j < K1_PPT; ++j) {
            if (!valid[j]) continue;
            const int pix = p0 + tid + j * K1_THREADS;
            const float ld = p.logdepth[pix_base + pix];
            const int v = pix / W, u = pix - v * W;
            PixState s;
            pixel_state(sh.fp, ld, u, v, s);
            const float ga = pixel_epilogue(sh.fp, s, gsx[j], gsy[j], gsz[j], facc);
            if (accum) {
                atomicAdd(&p.g_align[pix_base + pix], ga);
                if (FLOW) {
                    atomicAdd(&p.g_flow_i[pix_base + pix], gfi[j] * s.dfac);
                    atomicAdd(&p.g_flow_j[pix_base + pix], gfj[j] * s.dfac);
                }
            } else {
                p.g_align[pix_base + pix] = ga;
                if (FLOW) {
                    p.g_flow_i[pix_base + pix] = gfi[j] * s.dfac;
                    p.g_flow_j[pix_base + pix] = gfj[j] * s.dfac;
                }
            }
        }
    }
    flush();
}

// ==========================================================================================
// K2: loss assembly, gate, temporal term, parameter-chain gradients, Adam on the small parameters.
//     O(E + N) small fp64 blocks, split at its two global dependencies into four tiny launches:
//     scalars/gate -> camera chains (atomics) -> gradients + Adam -> shared focal.
// ==========================================================================================
__device__ __forceinline__ void scatter_pair(const PairGrad& g, double* fgS, double* fgT) {
    for (int k = 0; k < 15; ++k) {
        if (g.s[k] != 0.0) atomicAdd(&fgS[k], g.s[k]);
        if (g.t[k] != 0.0) atomicAdd(&fgT[k], g.t[k]);
    }
}

// extra scratch scalars of the update step (A3rProblem.scal)
constexpr int S_GSUM = 8;      // sum_e dL/dlog s_e  (mean coupling of the pair scales)
constexpr int S_GPHI = 9;      // sum_n dL/dphi_n    (shared focal)

// depth prior: 4 sums per local frame (grid: chunks x local frames), fp64 atomics into prior_acc
__global__ void __launch_bounds__(256) k_depth_prior(const A3rProblem p) {
    __shared__ double s_red[32];
    const int lf = blockIdx.y, n = p.lframe_id[lf];
    const int P = p.frame_hw[2 * n] * p.frame_hw[2 * n + 1];
    const int64_t base = (int64_t)lf * p.pix_stride;
    double a[PR_N] = {0, 0, 0, 0};
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gridDim.x * blockDim.x) {
        const float d = prior_delta(p.logdepth[base + q], p.logdepth_init[base + q]);
        const float w = p.static_mask[base + q] ? 1.0f : 2.0f;          // 1 + dynamic_mask
        a[PR_D] += d; a[PR_W] += w; a[PR_WD] += w * d; a[PR_WDD] += w * d * d;
    }
    for (int k = 0; k < PR_N; ++k) {
        const double v = block_sum(a[k], s_red);
        if (threadIdx.x == 0) atomicAdd(&p.prior_acc[(int64_t)n * PR_N + k], v);
    }
}

// ==========================================================================================
// K2: loss, gate, every parameter chain, Adam on the small parameters and the parameter maps of the NEXT iteration in
// ONE launch: a thread-block cluster of 8 CTAs x 256 threads whose phases are separated by the hardware cluster barrier
// (barrier.cluster arrive.release / wait.acquire orders the global-memory traffic of the phases), instead of five
// launches + a memset separated by kernel boundaries.  This chain is pure latency (a few thousand items of fp64
// arithmetic, instruction fetch of run-once code, dependent L2 round trips) and sits on the critical path between two
// scans, so the kernel is organised around doing every expensive thing once:
//   * the fp64 camera of every frame (quaternion -> R, signed_expm1 translations, exp focal: FrameD) is evaluated ONCE
//     per CTA into shared memory and every chain reads it from there; once more after the update for the maps;
//   * the shared focal's Adam step is evaluated redundantly by every CTA from values it read before anyone writes
//     them, so it needs no phase (and no barrier) of its own;
//   * frame_grad is zeroed in the tail for the next call (the entry points zero it once before the first).
//   phase A  cameras -> shared memory | loss sums + gate: every CTA in the same order, CTA 0 publishes p.scal |
//            chains that end in a camera (fp64 atomics into frame_grad) + dL/dlog s_e
//   phase B  gradients of pw_poses / im_poses (/ per-frame focal, pp) + torch.optim.Adam
//   phase C  the shared focal | cameras of the updated parameters -> shared memory | parameter maps (= k_prepare of the
//            next iteration) | zero frame_grad
// ==========================================================================================
constexpr int K2_CLUSTER = 8;
constexpr int K2_THREADS = 256;
constexpr int K2_MAX_FRAMES = 1280;        // cameras in shared memory: 160 B each (200 KB of the 227 KB opt-in limit); more: global scratch

// diagnostic phase trace of K2 (a3r_debug_k2_trace): globaltimer stamps of thread 0 of every CTA at the phase boundaries
__device__ unsigned long long g_k2_trace[K2_CLUSTER * 8];
__device__ int g_k2_trace_on = 0;
__device__ __forceinline__ void k2_stamp(int slot) {
    if (g_k2_trace_on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_k2_trace[blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// the four flow sums (numerator / per-channel mask count of both sides) over all 2E slots in a fixed 256-thread order:
// K2 and the depth-plane Adam kernel K3 both evaluate them (K3 must not wait for K2) and must agree bit for bit.
// Element k = 2 gs + which of the (FNUM, FCNT) pairs; with a stride of 256 a thread only ever sees category k & 3 =
// (side, which), so it carries ONE running sum, all loads are independent, and one shuffle tree + one block barrier
// finish all four sums.  out = num_i, cnt_i, num_j, cnt_j.  `red`: >= 32 doubles of shared memory.
__device__ __forceinline__ void flow_sums_256(const A3rProblem& p, double* red, double out[4]) {
    static_assert(ACC_FCNT == ACC_FNUM + 1, "flow numerator and count are adjacent in a slot's accumulator row");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = 4 * p.n_edges;
    double acc = 0.0;
#pragma unroll 4
    for (int k = tid; k < total; k += 256) acc += p.acc_slot[(int64_t)(k >> 1) * ACC_STRIDE + ACC_FNUM + (k & 1)];
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (lane < 4) red[warp * 4 + lane] = acc;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += red[w * 4 + c];
        out[c] = r;
    }
    __syncthreads();
}

// three CTA-wide sums with one block barrier (256 threads; fixed order, every thread gets the results)
__device__ __forceinline__ void block_sum3_256(double& a, double& b, double& c, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { red[warp * 3] = a; red[warp * 3 + 1] = b; red[warp * 3 + 2] = c; }
    __syncthreads();
    double ra = 0.0, rb = 0.0, rc = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ra += red[w * 3]; rb += red[w * 3 + 1]; rc += red[w * 3 + 2]; }
    __syncthreads();
    a = ra; b = rb; c = rc;
}

__device__ __forceinline__ void adam_vec(float* prm, float* m, float* v, const float* g, int cnt,
                                         const A3rProblem& p, const A3rAdamStep& a) {
    for (int k = 0; k < cnt; ++k)
        adam_update(prm[k], m[k], v[k], g[k], p.adam_beta1, p.adam_beta2, p.adam_eps, a.step_size, a.bc2_sqrt);
}

// every frame's fp64 camera into shared memory; `focals` may point at a shared-memory copy of the shared focal
// (one copy of this code, called twice; plain arguments so that the parameter block is not spilled to pass a reference)
__device__ __noinline__ void fill_cameras(const float* im_poses, const float* focals, const float* im_pp, const int32_t* frame_hw,
                                          int N, bool shared_focal, double focal_break, FrameD* fr) {
    for (int n = threadIdx.x; n < N; n += K2_THREADS)
        load_frame(im_poses, focals, im_pp, frame_hw, n, shared_focal, focal_break, fr[n]);
    __syncthreads();
}

__global__ void __cluster_dims__(K2_CLUSTER, 1, 1) __launch_bounds__(K2_THREADS, 1)
k_small(const A3rProblem p, const uint32_t mode, const A3rAdamStep adam, const uint32_t prepare_next, FrameD* const cam_scratch) {
    extern __shared__ __align__(16) unsigned char k2_smem[];
    // cameras: shared memory up to K2_MAX_FRAMES frames; beyond that a library-owned global scratch every CTA fills with the
    // same values (identical writes; a CTA reads what it wrote itself after its own block barrier)
    FrameD* fr = cam_scratch ? cam_scratch : reinterpret_cast<FrameD*>(k2_smem);
    __shared__ double s_red[32];
    __shared__ double s_sc[2];
    __shared__ double s_mean;
    __shared__ float s_foc[3];                 // the shared focal: parameter, m, v as they were when the kernel started
    const int tid = threadIdx.x;
    const int vt = blockIdx.x * K2_THREADS + tid, nvt = gridDim.x * K2_THREADS;      // the grid is one cluster
    const int N = p.n_frames, E = p.n_edges;
    const bool shared_focal = p.flags & A3R_F_SHARED_FOCAL;
    const bool norm_scale = p.flags & A3R_F_NORM_PW_SCALE;
    const bool flow_active = (mode & A3R_M_FLOW_ACTIVE) && (p.flags & A3R_F_HAS_FLOW) && p.flow_loss_weight > 0.f;
    const bool apply = mode & A3R_M_APPLY_ADAM;
    const double fb = (double)p.focal_break;
    const double tsw = (double)p.temporal_smoothing_weight, tw = (double)p.translation_weight;

    // ---------------- phase A: cameras, loss assembly + gate (reference optimizer.py:538-557), chains ----------------
    k2_stamp(0);
    if (shared_focal && tid == 0) { s_foc[0] = p.im_focals[0]; s_foc[1] = p.im_focals_m[0]; s_foc[2] = p.im_focals_v[0]; }
    fill_cameras(p.im_poses, p.im_focals, p.im_pp, p.frame_hw, N, shared_focal, fb, fr);
    k2_stamp(1);
    {
        double l_align = 0, l_temp = 0, l_prior = 0;
        if (p.depth_regularize_weight > 0.f)
            for (int n = tid; n < N; n += K2_THREADS)
                l_prior += prior_frame(p.prior_acc + (int64_t)n * PR_N, (double)p.frame_hw[2 * n] * p.frame_hw[2 * n + 1]).loss;
        for (int gs = tid; gs < 2 * E; gs += K2_THREADS) l_align += p.acc_slot[(int64_t)gs * ACC_STRIDE + ACC_LOSS];
        if (tsw > 0)
            for (int n = tid; n < N - 1; n += K2_THREADS) l_temp += temporal_pair(fr[n + 1], fr[n], tsw, tw, nullptr);
        double fs[4];
        flow_sums_256(p, s_red, fs);
        block_sum3_256(l_align, l_temp, l_prior, s_red);
        l_prior /= (double)N;
        if (tid == 0) {
            LossScalars r = assemble_loss(l_align, l_temp, fs[0], fs[1], fs[2], fs[3], flow_active,
                                          (double)p.flow_loss_weight, (double)p.flow_loss_thre, tsw);
            s_sc[0] = r.sc_i; s_sc[1] = r.sc_j;
            if (blockIdx.x == 0) {
                r.loss += (double)p.depth_regularize_weight * l_prior;           // optimizer.py:554-557
                p.scal[A3R_S_PRIOR] = l_prior;
                p.scal[A3R_S_LOSS] = r.loss; p.scal[A3R_S_ALIGN] = l_align; p.scal[A3R_S_TEMPORAL] = l_temp;
                p.scal[A3R_S_FLOW] = r.floss; p.scal[A3R_S_FLOW_DROPPED] = r.dropped;
                p.scal[A3R_S_SC_I] = r.sc_i; p.scal[A3R_S_SC_J] = r.sc_j;
                if (r.dropped != 0) p.scal[A3R_S_FLAG_STICKY] = 1.0;
                if (p.loss_hist && adam.iter >= 0) p.loss_hist[adam.iter] = (float)r.loss;
            }
        }
        __syncthreads();
    }
    k2_stamp(2);
    //   chain items: [0,N) alignment part of frame n | [N, N+2E) flow chain of a slot | [.., +N-1) temporal pair | [.., +E) edge
    //   (frame_grad, scal[S_GSUM] and scal[S_GPHI] are zero on entry: zeroed by the previous call's tail / the entry point)
    {
        const double sc_i = s_sc[0], sc_j = s_sc[1];
        const int total = N + 2 * E + (N - 1) + E;
        for (int it = vt; it < total; it += nvt) {
            if (it < N) {
                const int n = it;
                double fg[FG_STRIDE];
                frame_align_grad(fr[n], p.acc_frame + (int64_t)n * FACC_N, fg);
                double* dst = p.frame_grad + (int64_t)n * FG_STRIDE;
                for (int k = 0; k < 15; ++k) atomicAdd(&dst[k], fg[k]);
            } else if (it < N + 2 * E) {
                if (!(flow_active && (sc_i != 0 || sc_j != 0))) continue;
                const int gs = it - N, e = gs >> 1, side = gs & 1;
                const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
                const int sI = side ? j : i, tI = side ? i : j;
                PairGrad g;
                slot_flow_chain(fr[sI], fr[tI], p.acc_slot + (int64_t)gs * ACC_STRIDE, side ? sc_j : sc_i, g);
                scatter_pair(g, p.frame_grad + (int64_t)sI * FG_STRIDE, p.frame_grad + (int64_t)tI * FG_STRIDE);
            } else if (it < N + 2 * E + N - 1) {
                if (!(tsw > 0)) continue;
                const int n = it - N - 2 * E;
                PairGrad g;
                temporal_pair(fr[n + 1], fr[n], tsw, tw, &g);
                scatter_pair(g, p.frame_grad + (int64_t)(n + 1) * FG_STRIDE, p.frame_grad + (int64_t)n * FG_STRIDE);
            } else {
                const int e = it - N - 2 * E - (N - 1);
                const double* a0 = p.acc_slot + (int64_t)(2 * e) * ACC_STRIDE;
                const double g = edge_logscale_grad(p.edge_par + EP_STRIDE * e, a0, a0 + ACC_STRIDE);
                p.edge_scratch[e] = g;
                atomicAdd(&p.scal[S_GSUM], g);
            }
        }
    }
    k2_stamp(3);
    cluster_barrier();
    k2_stamp(4);

    // ---------------- phase B: gradients of the small parameters + torch.optim.Adam ----------------
    //   items: [0,E) pairwise pose | [E, E+N) image pose
    {
        const double gmean = norm_scale ? p.scal[S_GSUM] / (double)E : 0.0;
        for (int it = vt; it < E + N; it += nvt) {
            if (it < E) {
                const int e = it;
                const double* a0 = p.acc_slot + (int64_t)(2 * e) * ACC_STRIDE;
                float g[8], ga[2];
                edge_pose_grad(p.pw_poses + 8 * e, p.edge_par + EP_STRIDE * e, a0, a0 + ACC_STRIDE, p.edge_scratch[e] - gmean,
                               p.flags & A3R_F_TRAIN_ADAPTORS, norm_scale, (double)p.pw_break, g, ga);
                for (int k = 0; k < 8; ++k) p.g_pw_poses[8 * e + k] = g[k];
                p.g_pw_adaptors[2 * e] = ga[0]; p.g_pw_adaptors[2 * e + 1] = ga[1];
                if (apply && (p.flags & A3R_F_TRAIN_PW_POSES)) adam_vec(p.pw_poses + 8 * e, p.pw_poses_m + 8 * e, p.pw_poses_v + 8 * e, g, 8, p, adam);
                if (apply && (p.flags & A3R_F_TRAIN_ADAPTORS)) adam_vec(p.pw_adaptors + 2 * e, p.pw_adaptors_m + 2 * e, p.pw_adaptors_v + 2 * e, ga, 2, p, adam);
            } else {
                const int n = it - E;
                float g[7], gpp[2];
                double gphi;
                frame_pose_grad(p.im_poses + 7 * n, fr[n], p.frame_grad + (int64_t)n * FG_STRIDE, fb, p.flags & A3R_F_TRAIN_PP, g, &gphi, gpp);
                for (int k = 0; k < 7; ++k) p.g_im_poses[7 * n + k] = g[k];
                p.g_im_pp[2 * n] = gpp[0]; p.g_im_pp[2 * n + 1] = gpp[1];
                if (shared_focal) atomicAdd(&p.scal[S_GPHI], gphi);     // applied in phase C
                else {
                    float gf = (float)gphi;
                    p.g_im_focals[n] = gf;
                    if (apply && (p.flags & A3R_F_TRAIN_FOCALS)) adam_vec(p.im_focals + n, p.im_focals_m + n, p.im_focals_v + n, &gf, 1, p, adam);
                }
                if (apply && (p.flags & A3R_F_TRAIN_PP)) adam_vec(p.im_pp + 2 * n, p.im_pp_m + 2 * n, p.im_pp_v + 2 * n, gpp, 2, p, adam);
                if (apply && (p.flags & A3R_F_TRAIN_IM_POSES)) adam_vec(p.im_poses + 7 * n, p.im_poses_m + 7 * n, p.im_poses_v + 7 * n, g, 7, p, adam);
            }
        }
    }
    k2_stamp(5);
    cluster_barrier();
    k2_stamp(6);

    // ---------------- phase C: the shared focal, the parameter maps of the next iteration, zero the scratch ----------------
    // Every CTA steps the shared focal from the values it saved in phase A (nobody has written them yet) and the complete
    // sum of phase B: the same number on every CTA; CTA 0 stores it.
    if (shared_focal && tid == 0) {
        float gf = (float)p.scal[S_GPHI];
        if (apply && (p.flags & A3R_F_TRAIN_FOCALS)) adam_vec(&s_foc[0], &s_foc[1], &s_foc[2], &gf, 1, p, adam);
        if (blockIdx.x == 0) {
            p.g_im_focals[0] = gf;
            p.im_focals[0] = s_foc[0]; p.im_focals_m[0] = s_foc[1]; p.im_focals_v[0] = s_foc[2];
        }
    }
    __syncthreads();
    if (prepare_next) {       // reference base_opt.py:195-247, optimizer.py:398-428 of the updated parameters
        const float* focals = shared_focal ? s_foc : p.im_focals;
        fill_cameras(p.im_poses, focals, p.im_pp, p.frame_hw, N, shared_focal, fb, fr);
        double part = 0.0;    // every CTA recomputes mean_e(log scale) in k_prepare's order -> the same number everywhere
        for (int e = tid; e < E; e += K2_THREADS) part += (double)p.pw_poses[8 * e + 7];
        part = block_sum(part, s_red);
        if (tid == 0) s_mean = part / (double)E;
        __syncthreads();
        const double log_norm = norm_scale ? (log((double)p.base_scale) - s_mean) : 0.0;
        for (int n = vt; n < N; n += nvt) derive_frame_par(fr[n], p.frame_par + FP_STRIDE * n);
        for (int e = vt; e < E; e += nvt)
            derive_edge_par(p.pw_poses + 8 * e, p.pw_adaptors + 2 * e, log_norm, norm_scale, (double)p.pw_break, p.edge_par + EP_STRIDE * e);
        if (p.flags & A3R_F_HAS_FLOW) {
            for (int gs = vt; gs < 2 * E; gs += nvt) {
                const int e = gs >> 1, side = gs & 1;
                const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
                derive_slot_par(fr[side ? j : i], fr[side ? i : j], p.slot_par + SP_STRIDE * gs);
            }
        }
    }
    for (int i = vt; i < N * FG_STRIDE; i += nvt) p.frame_grad[i] = 0.0;
    if (vt == 0) { p.scal[S_GSUM] = 0.0; p.scal[S_GPHI] = 0.0; }
    k2_stamp(7);
}

// ==========================================================================================
// K3: Adam on the log-depth planes (or, without A3R_M_APPLY_ADAM, total depth gradient -> g_align)
// ==========================================================================================
template <bool FLOW>
__global__ void __launch_bounds__(256, 4) k_update_depth(const A3rProblem p, const uint32_t mode, const A3rAdamStep adam) {
    const int64_t total4 = (int64_t)p.n_local_frames * p.pix_stride / 4;
    const bool prior = p.depth_regularize_weight > 0.f;
    // the flow normalisers and the gate, evaluated here exactly as K2 evaluates them: K3 starts right after the scan,
    // beside K2, instead of waiting for K2's scalars
    float sc_i = 0.f, sc_j = 0.f;
    if (FLOW) {
        __shared__ double s_red[32];
        __shared__ float s_sc[2];
        double fs[4];
        flow_sums_256(p, s_red, fs);
        if (threadIdx.x == 0) {
            const LossScalars r = assemble_loss(0.0, 0.0, fs[0], fs[1], fs[2], fs[3], true, (double)p.flow_loss_weight,
                                                (double)p.flow_loss_thre, 0.0);
            s_sc[0] = (float)r.sc_i; s_sc[1] = (float)r.sc_j;
        }
        __syncthreads();
        sc_i = s_sc[0]; sc_j = s_sc[1];
    }
    const bool apply = mode & A3R_M_APPLY_ADAM;
    float4* ld = reinterpret_cast<float4*>(p.logdepth);
    float4* mm = reinterpret_cast<float4*>(p.logdepth_m);
    float4* vv = reinterpret_cast<float4*>(p.logdepth_v);
    float4* ga = reinterpret_cast<float4*>(p.g_align);
    const float4* gi = reinterpret_cast<const float4*>(p.g_flow_i);
    const float4* gj = reinterpret_cast<const float4*>(p.g_flow_j);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = ga[i];
        if (FLOW) {
            const float4 a = gi[i], b = gj[i];
            g.x += sc_i * a.x + sc_j * b.x; g.y += sc_i * a.y + sc_j * b.y;
            g.z += sc_i * a.z + sc_j * b.z; g.w += sc_i * a.w + sc_j * b.w;
        }
        if (prior) {               // d (weight * prior) / d logdepth: per-frame scalars from the 4 sums, then per pixel
            const int64_t e0 = i * 4;
            const int lf = (int)(e0 / p.pix_stride);
            const int q0 = (int)(e0 - (int64_t)lf * p.pix_stride);
            const int n = p.lframe_id[lf];
            const int P = p.frame_hw[2 * n] * p.frame_hw[2 * n + 1];
            if (q0 < P) {
                const PriorFrame pf = prior_frame(p.prior_acc + (int64_t)n * PR_N, (double)P);
                const float coef = (float)(2.0 * (double)p.depth_regularize_weight / ((double)p.n_frames * (double)P));
                const float4 l = ld[i];
                const float4 l0 = reinterpret_cast<const float4*>(p.logdepth_init)[i];
                const uchar4 sm = reinterpret_cast<const uchar4*>(p.static_mask)[i];
                const float s = (float)pf.s, mb = (float)pf.mean_b;
                g.x += prior_grad(l.x, l0.x, sm.x ? 1.f : 2.f, s, mb, coef);
                if (q0 + 1 < P) g.y += prior_grad(l.y, l0.y, sm.y ? 1.f : 2.f, s, mb, coef);
                if (q0 + 2 < P) g.z += prior_grad(l.z, l0.z, sm.z ? 1.f : 2.f, s, mb, coef);
                if (q0 + 3 < P) g.w += prior_grad(l.w, l0.w, sm.w ? 1.f : 2.f, s, mb, coef);
            }
        }
        if (apply) {
            float4 l = ld[i], m = mm[i], v = vv[i];
            adam_update(l.x, m.x, v.x, g.x, p.adam_beta1, p.adam_beta2, p.adam_eps, adam.step_size, adam.bc2_sqrt);
            adam_update(l.y, m.y, v.y, g.y, p.adam_beta1, p.adam_beta2, p.adam_eps, adam.step_size, adam.bc2_sqrt);
            adam_update(l.z, m.z, v.z, g.z, p.adam_beta1, p.adam_beta2, p.adam_eps, adam.step_size, adam.bc2_sqrt);
            adam_update(l.w, m.w, v.w, g.w, p.adam_beta1, p.adam_beta2, p.adam_eps, adam.step_size, adam.bc2_sqrt);
            ld[i] = l; mm[i] = m; vv[i] = v;
        } else {
            ga[i] = g;
        }
    }
}

// ==========================================================================================
// read-out: world pointmaps (reference optimizer.py:468-478)
// ==========================================================================================
__global__ void __launch_bounds__(256) k_unproject(const A3rProblem p, float* __restrict__ out) {
    const int lf = blockIdx.y;
    const int n = p.lframe_id[lf];
    const int W = p.frame_hw[2 * n + 1], P = p.frame_hw[2 * n] * W;
    __shared__ float fp[FP_STRIDE];
    if (threadIdx.x < FP_STRIDE) fp[threadIdx.x] = p.frame_par[FP_STRIDE * n + threadIdx.x];
    __syncthreads();
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
        const int v = pix / W, u = pix - v * W;
        PixState s;
        pixel_state(fp, p.logdepth[(int64_t)lf * p.pix_stride + pix], u, v, s);
        float* o = out + ((int64_t)lf * p.pix_stride + pix) * 3;
        o[0] = s.zx; o[1] = s.zy; o[2] = s.zz;
    }
}

// ==========================================================================================
// one-off packing
// ==========================================================================================
__global__ void __launch_bounds__(256) k_pack_pred(const float* __restrict__ pts, const float* __restrict__ conf, int64_t P,
                                                   const int32_t* __restrict__ dst_slot, const int32_t* __restrict__ dst_lframe,
                                                   int64_t pix_stride, int conf_mode, float* __restrict__ predw,
                                                   float* __restrict__ im_conf, double* __restrict__ conf_sum) {
    __shared__ double s_red[32];
    const int k = blockIdx.y;
    const int slot = dst_slot[k], lf = dst_lframe[k];
    const float* sp = pts + (int64_t)k * P * 3;
    const float* sc = conf + (int64_t)k * P;
    float4* dp = reinterpret_cast<float4*>(predw) + (int64_t)slot * pix_stride;
    double csum = 0.0;
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += (int64_t)gridDim.x * blockDim.x) {
        const float c = sc[pix];
        float w;
        switch (conf_mode) {                   // reference commons.py:53-64
            case 0: w = logf(c); break;
            case 1: w = sqrtf(c); break;
            case 2: w = c - 1.0f; break;
            default: w = c; break;
        }
        dp[pix] = make_float4(sp[3 * pix], sp[3 * pix + 1], sp[3 * pix + 2], w);
        if (im_conf && c > 0.f) atomicMax(reinterpret_cast<int*>(im_conf + (int64_t)lf * pix_stride + pix), __float_as_int(c));
        csum += (double)c;
    }
    csum = block_sum(csum, s_red);
    if (threadIdx.x == 0 && conf_sum) atomicAdd(&conf_sum[slot], csum);
}

__global__ void __launch_bounds__(256) k_pack_flow(const float* __restrict__ src, int64_t P, const int32_t* __restrict__ dst_slot,
                                                   int64_t pix_stride, float* __restrict__ dst) {
    const int k = blockIdx.y;
    const float* fu = src + (int64_t)k * 2 * P;
    const float* fv = fu + P;
    float2* d = reinterpret_cast<float2*>(dst) + (int64_t)dst_slot[k] * pix_stride;
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += (int64_t)gridDim.x * blockDim.x)
        d[pix] = make_float2(fu[pix], fv[pix]);
}

// ==========================================================================================
// C ABI
// ==========================================================================================
// Scan kernel variant.  v4 (ga_scan_v4.cuh: independent warps, private cp.async rings, shared-memory slot table,
// half-warp transpose reduction) is the product path.  A3R_SCAN=v1 selects the simple thread-owns-pixels kernel above:
// the on-device cross-check of v4 in the GPU tests (same per-element math header, different organisation).
static int scan_variant() {
    static int v = [] {
        const char* e = getenv("A3R_SCAN");
        return (e && e[0] == 'v' && e[1] == '1') ? 1 : 4;
    }();
    return v;
}
static int scan_tile() { return scan_variant() == 1 ? K1_TILE : V4_T; }
static int scan_slots() { return scan_variant() == 1 ? K1_MAXS : V4_SLOTS; }

static int check_problem(const A3rProblem* p) {
    if (!p) return a3r_fail(A3R_E_ARG, "problem is NULL");
    if (p->tile_pixels != scan_tile()) return a3r_fail(A3R_E_ARG, "tile_pixels %d != a3r_tile_pixels() %d", p->tile_pixels, scan_tile());
    if (p->pix_stride % scan_tile() != 0) return a3r_fail(A3R_E_ARG, "pix_stride must be a multiple of a3r_tile_pixels()");
    if (p->n_frames <= 0 || p->n_edges <= 0 || p->n_local_frames < 0 || p->n_ctas <= 0)
        return a3r_fail(A3R_E_ARG, "bad sizes N=%d E=%d local=%d ctas=%d", p->n_frames, p->n_edges, p->n_local_frames, p->n_ctas);
    if (p->pix_stride % 4 != 0) return a3r_fail(A3R_E_ARG, "pix_stride must be a multiple of 4");
    if (p->acc_frame != p->acc_slot + (int64_t)2 * p->n_edges * A3R_SLOT_ACC)
        return a3r_fail(A3R_E_ARG, "acc_frame must directly follow acc_slot");
    return A3R_OK;
}

template <bool FLOW, bool L2D>
static int launch_accumulate(const A3rProblem* p, cudaStream_t st) {
    if (scan_variant() == 1) {
        const size_t smem = sizeof(double) * K1_WARPS * K1_MAXS * ACC_STRIDE + sizeof(K1Shared);
        static bool configured = false;   // per template instantiation
        if (!configured) {
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v1<FLOW, L2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = true;
        }
        k_accumulate_v1<FLOW, L2D><<<p->n_ctas, K1_THREADS, smem, st>>>(*p);
    } else {
        if (p->n_ctas % V4_WARPS != 0) return a3r_fail(A3R_E_ARG, "n_ctas must be a multiple of %d for the v4 scan", V4_WARPS);
        // copy engine of the private rings (A3R_V4_COPY=tma|ldgsts forces one; default: per-lane cp.async)
        static const int forced = [] {
            const char* e = getenv("A3R_V4_COPY");
            return (!e || !e[0]) ? 0 : (e[0] == 't' ? 1 : 2);
        }();
        const bool tma = forced == 1;      // measured on C2: per-lane cp.async 0.760 / 0.474 ms (flow / align), TMA 0.843 / 0.533 ms
        static bool configured = false;
        if (!configured) {
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            configured = true;
        }
        if (tma) k_accumulate_v4<FLOW, L2D, true><<<p->n_ctas / V4_WARPS, V4_THREADS, V4_SMEM, st>>>(*p);
        else k_accumulate_v4<FLOW, L2D, false><<<p->n_ctas / V4_WARPS, V4_THREADS, V4_SMEM, st>>>(*p);
    }
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

extern "C" {

int a3r_abi_version(void) { return A3R_ABI_VERSION; }
const char* a3r_last_error(void) { return a3r_error_buffer(); }
int a3r_tile_pixels(void) { return scan_tile(); }
int a3r_max_slots(void) { return scan_slots(); }

int a3r_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    // three attribute queries (microseconds); cudaGetDeviceProperties takes tens of milliseconds on these boxes
    int sm = 0, major = 0, minor = 0;
    A3R_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    A3R_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    A3R_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (sm_count) *sm_count = sm;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return A3R_OK;
}

int a3r_max_ctas(int device) {
    int sm = 0;
    if (a3r_device_info(device, &sm, nullptr, nullptr) != A3R_OK) return A3R_E_CUDA;
    // v1: 2 resident CTAs per SM; v4: every warp has its own range of the plan
    const int v = scan_variant();
    return v == 1 ? sm * 2 : sm * V4_WARPS;
}

int a3r_pack_pred(const float* pts, const float* conf, int32_t n, int64_t P, const int32_t* dst_slot,
                  const int32_t* dst_lframe, int64_t pix_stride, int32_t conf_mode, float* predw,
                  float* im_conf, double* conf_sum, void* stream) {
    if (n <= 0) return A3R_OK;
    if (!pts || !conf || !dst_slot || !dst_lframe || !predw) return a3r_fail(A3R_E_ARG, "a3r_pack_pred: NULL pointer");
    dim3 grid((unsigned)((P + 1023) / 1024 < 1 ? 1 : (P + 1023) / 1024), (unsigned)n);
    if (grid.x > 64) grid.x = 64;
    k_pack_pred<<<grid, 256, 0, (cudaStream_t)stream>>>(pts, conf, P, dst_slot, dst_lframe, pix_stride, conf_mode, predw, im_conf, conf_sum);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_pack_flow(const float* flow_src, int32_t n, int64_t P, const int32_t* dst_slot, int64_t pix_stride,
                  float* flow_dst, void* stream) {
    if (n <= 0) return A3R_OK;
    if (!flow_src || !dst_slot || !flow_dst) return a3r_fail(A3R_E_ARG, "a3r_pack_flow: NULL pointer");
    dim3 grid((unsigned)((P + 1023) / 1024 < 1 ? 1 : (P + 1023) / 1024), (unsigned)n);
    if (grid.x > 64) grid.x = 64;
    k_pack_flow<<<grid, 256, 0, (cudaStream_t)stream>>>(flow_src, P, dst_slot, pix_stride, flow_dst);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_ga_prepare(const A3rProblem* p, uint32_t mode, void* stream) {
    (void)mode;
    if (int rc = check_problem(p)) return rc;
    int work = 2 * p->n_edges > p->n_frames ? 2 * p->n_edges : p->n_frames;
    int grid = (work + 255) / 256;
    if (grid > 148) grid = 148;
    k_prepare<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

int a3r_ga_accumulate(const A3rProblem* p, uint32_t mode, void* stream) {
    if (int rc = check_problem(p)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool flow = (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
    if (flow && (!p->flow || !p->static_mask)) return a3r_fail(A3R_E_ARG, "flow active but flow/static_mask planes are NULL");
    const bool prior = p->depth_regularize_weight > 0.f;
    if (prior && (!p->logdepth_init || !p->static_mask || !p->prior_acc))
        return a3r_fail(A3R_E_ARG, "depth_regularize_weight > 0 needs logdepth_init, static_mask and prior_acc");
    if (prior && p->prior_acc != p->acc_frame + (int64_t)p->n_frames * A3R_FRAME_ACC)
        return a3r_fail(A3R_E_ARG, "prior_acc must directly follow acc_frame");
    const size_t acc_bytes = sizeof(double) * ((size_t)2 * p->n_edges * A3R_SLOT_ACC + (size_t)p->n_frames * (A3R_FRAME_ACC + (prior ? A3R_PRIOR_ACC : 0)));
    A3R_CUDA(cudaMemsetAsync(p->acc_slot, 0, acc_bytes, st));
    if (prior && p->n_local_frames > 0) {
        k_depth_prior<<<dim3(16, (unsigned)p->n_local_frames), 256, 0, st>>>(*p);
        A3R_CUDA(cudaGetLastError());
    }
    if (p->n_work_frames > p->n_local_frames) {     // split frames add their gradients atomically
        const size_t plane = sizeof(float) * (size_t)p->n_local_frames * p->pix_stride;
        A3R_CUDA(cudaMemsetAsync(p->g_align, 0, plane, st));
        if (flow) {
            A3R_CUDA(cudaMemsetAsync(p->g_flow_i, 0, plane, st));
            A3R_CUDA(cudaMemsetAsync(p->g_flow_j, 0, plane, st));
        }
    }
    if (p->n_tiles == 0) return A3R_OK;
    const bool l2 = p->flags & A3R_F_DIST_L2;
    if (flow) return l2 ? launch_accumulate<true, true>(p, st) : launch_accumulate<true, false>(p, st);
    return l2 ? launch_accumulate<false, true>(p, st) : launch_accumulate<false, false>(p, st);
}

// The depth-plane Adam step (K3, HBM-bound, ~38 us on C2) needs only the flow normalisers, which it evaluates itself from
// the slot accumulators; K2 (latency-bound) never touches the depth planes.  They run side by side: K3 on a library-owned
// auxiliary stream, forked right after the scan (or the exchange) and joined before the next scan.
struct AuxLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static int aux_lane(AuxLane** out) {
    static AuxLane lanes[64];
    int dev = 0;
    A3R_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return a3r_fail(A3R_E_ARG, "device index %d out of range", dev);
    AuxLane& l = lanes[dev];
    if (!l.stream) {
        A3R_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        A3R_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
        A3R_CUDA(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
    }
    *out = &l;
    return A3R_OK;
}

// K2 reads frame_grad / scal[S_GSUM] / scal[S_GPHI] as zero and leaves them zero: once per entry point, before the first launch
static int zero_small_scratch(const A3rProblem* p, cudaStream_t st) {
    A3R_CUDA(cudaMemsetAsync(p->frame_grad, 0, sizeof(double) * FG_STRIDE * (size_t)p->n_frames, st));
    A3R_CUDA(cudaMemsetAsync(p->scal + S_GSUM, 0, sizeof(double) * 2, st));
    return A3R_OK;
}

// cameras of problems with more than K2_MAX_FRAMES frames: one lazily grown global scratch per device (never shrinks)
struct CamScratch { void* ptr = nullptr; size_t bytes = 0; };
static int camera_scratch(int dev, size_t bytes, void** out) {
    static CamScratch scratch[64];
    CamScratch& c = scratch[dev];
    if (c.bytes < bytes) {
        if (c.ptr) { A3R_CUDA(cudaDeviceSynchronize()); A3R_CUDA(cudaFree(c.ptr)); c.ptr = nullptr; c.bytes = 0; }
        A3R_CUDA(cudaMalloc(&c.ptr, bytes));
        c.bytes = bytes;
    }
    *out = c.ptr;
    return A3R_OK;
}

static int launch_small(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, cudaStream_t st, bool prepare_next) {
    static bool attr_set[64] = {};
    static int max_smem_frames = -1;
    int dev = 0;
    A3R_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return a3r_fail(A3R_E_ARG, "device index %d out of range", dev);
    if (!attr_set[dev]) {
        A3R_CUDA(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FrameD) * K2_MAX_FRAMES)));
        attr_set[dev] = true;
    }
    if (max_smem_frames < 0) {        // A3R_K2_SMEM_FRAMES: lower the switch-over point (tests exercise the global-scratch path)
        const char* e = getenv("A3R_K2_SMEM_FRAMES");
        max_smem_frames = e ? atoi(e) : K2_MAX_FRAMES;
        if (max_smem_frames > K2_MAX_FRAMES || max_smem_frames < 0) max_smem_frames = K2_MAX_FRAMES;
    }
    void* scratch = nullptr;
    size_t smem = sizeof(FrameD) * (size_t)p->n_frames;
    if (p->n_frames > max_smem_frames) {
        if (int rc = camera_scratch(dev, sizeof(FrameD) * (size_t)p->n_frames, &scratch)) return rc;
        smem = 0;
    }
    k_small<<<K2_CLUSTER, K2_THREADS, smem, st>>>(*p, mode, *adam, prepare_next ? 1u : 0u, static_cast<FrameD*>(scratch));
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

// update without the final join: *pending_join is set when K3 is still running on the auxiliary stream.
// prepare_next: K2's tail writes the parameter maps of the updated parameters (the next iteration's k_prepare).
static int update_forked(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, cudaStream_t st, AuxLane* lane, bool* pending_join,
                         bool prepare_next) {
    *pending_join = false;
    const bool apply = mode & A3R_M_APPLY_ADAM;
    if (p->n_local_frames > 0 && (!apply || (p->flags & A3R_F_TRAIN_DEPTH))) {
        A3R_CUDA(cudaEventRecord(lane->fork, st));
        A3R_CUDA(cudaStreamWaitEvent(lane->stream, lane->fork, 0));
        const bool flow = (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
        const int64_t total4 = (int64_t)p->n_local_frames * p->pix_stride / 4;
        int64_t grid = (total4 + 255) / 256;
        if (grid > 148 * 4) grid = 148 * 4;              // one resident wave (62 registers: 4 CTAs per SM)
        if (flow) k_update_depth<true><<<(unsigned)grid, 256, 0, lane->stream>>>(*p, mode, *adam);
        else k_update_depth<false><<<(unsigned)grid, 256, 0, lane->stream>>>(*p, mode, *adam);
        A3R_CUDA(cudaGetLastError());
        A3R_CUDA(cudaEventRecord(lane->join, lane->stream));
        *pending_join = true;
    }
    if (int rc = launch_small(p, mode, adam, st, prepare_next)) return rc;
    return A3R_OK;
}

int a3r_ga_update(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, void* stream) {
    if (int rc = check_problem(p)) return rc;
    if (!adam) return a3r_fail(A3R_E_ARG, "adam is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    AuxLane* lane = nullptr;
    if (int rc = aux_lane(&lane)) return rc;
    bool pending = false;
    if (int rc = zero_small_scratch(p, st)) return rc;
    if (int rc = update_forked(p, mode, adam, st, lane, &pending, false)) return rc;
    if (pending) A3R_CUDA(cudaStreamWaitEvent(st, lane->join, 0));      // everything is ordered in the caller's stream on return
    return A3R_OK;
}

int a3r_ga_run(const A3rProblem* p, int32_t count, const uint32_t* modes, const A3rAdamStep* steps, void* stream) {
    if (count < 0 || !modes || !steps) return a3r_fail(A3R_E_ARG, "a3r_ga_run: bad arguments");
    if (int rc = check_problem(p)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    AuxLane* lane = nullptr;
    if (int rc = aux_lane(&lane)) return rc;
    bool pending = false;
    if (int rc = zero_small_scratch(p, st)) return rc;
    if (count > 0)
        if (int rc = a3r_ga_prepare(p, modes[0], stream)) return rc;                 // later ones: the tail of K2
    for (int32_t k = 0; k < count; ++k) {
        if (pending) A3R_CUDA(cudaStreamWaitEvent(st, lane->join, 0));               // the scan reads the depth planes
        if (int rc = a3r_ga_accumulate(p, modes[k], stream)) return rc;
        if (int rc = update_forked(p, modes[k], &steps[k], st, lane, &pending, k + 1 < count)) return rc;
    }
    if (pending) A3R_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
    return A3R_OK;
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU exchange over peer memory
// ------------------------------------------------------------------------------------------------------------
constexpr size_t PEER_FLAG_STRIDE = 128;
constexpr size_t PEER_HEADER = A3R_MAX_PEERS * PEER_FLAG_STRIDE;

size_t a3r_peer_region_bytes(int64_t n_acc) { return PEER_HEADER + 2 * sizeof(double) * (size_t)n_acc; }

int a3r_peer_alloc(int64_t n_acc, void** region, unsigned char handle[64]) {
    if (n_acc <= 0 || !region || !handle) return a3r_fail(A3R_E_ARG, "a3r_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* ptr = nullptr;
    const size_t bytes = a3r_peer_region_bytes(n_acc);
    A3R_CUDA(cudaMalloc(&ptr, bytes));
    A3R_CUDA(cudaMemset(ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    A3R_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    *region = ptr;
    return A3R_OK;
}
int a3r_peer_open(const unsigned char handle[64], void** region) {
    if (!handle || !region) return a3r_fail(A3R_E_ARG, "a3r_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    A3R_CUDA(cudaIpcOpenMemHandle(region, h, cudaIpcMemLazyEnablePeerAccess));
    return A3R_OK;
}
int a3r_peer_close(void* region) { if (region) A3R_CUDA(cudaIpcCloseMemHandle(region)); return A3R_OK; }
int a3r_peer_free(void* region) { if (region) A3R_CUDA(cudaFree(region)); return A3R_OK; }

}  // extern "C"

__device__ __forceinline__ void st_release_sys(uint32_t* q, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(q), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* q) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* q) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(q) : "memory");
    return v;
}

// signal + wait + reduce in one launch (every CTA is resident: grid <= SM count)
__global__ void __launch_bounds__(256) k_exchange(const A3rPeers peers, const uint32_t tag, double* __restrict__ out) {
    const int W = peers.world, me = peers.rank;
    if (blockIdx.x == 0 && threadIdx.x < W && (int)threadIdx.x != me) {
        // the scan of this iteration completed before this kernel started (stream order): publish its sums
        __threadfence_system();
        uint32_t* flag = reinterpret_cast<uint32_t*>(static_cast<char*>(peers.base[threadIdx.x]) + (size_t)me * PEER_FLAG_STRIDE);
        st_release_sys(flag, tag);
    }
    if (threadIdx.x < W && (int)threadIdx.x != me) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(static_cast<const char*>(peers.base[me]) + (size_t)threadIdx.x * PEER_FLAG_STRIDE);
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(flag) - tag) < 0) {
            if (++spins > (1u << 26)) __trap();          // a peer that never arrives must fail loudly, not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    const size_t off = PEER_HEADER + (size_t)(tag & 1u) * sizeof(double) * (size_t)peers.n_acc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < peers.n_acc; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < W; ++r)                       // rank order: the same sum, bit for bit, on every rank
            s += ld_peer_f64(reinterpret_cast<const double*>(static_cast<const char*>(peers.base[r]) + off) + i);
        out[i] = s;
    }
}

extern "C" {

int a3r_ga_run_peers(const A3rProblem* p, const A3rPeers* peers, uint32_t tag0, int32_t count, const uint32_t* modes,
                     const A3rAdamStep* steps, void* stream) {
    if (count < 0 || !modes || !steps || !peers) return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: bad arguments");
    if (int rc = check_problem(p)) return rc;
    if (peers->world < 1 || peers->world > A3R_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world)
        return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: world %d rank %d", peers->world, peers->rank);
    const bool prior = p->depth_regularize_weight > 0.f;
    const int64_t n_acc = (int64_t)2 * p->n_edges * A3R_SLOT_ACC + (int64_t)p->n_frames * (A3R_FRAME_ACC + (prior ? A3R_PRIOR_ACC : 0));
    if (peers->n_acc != n_acc) return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: n_acc %lld != %lld", (long long)peers->n_acc, (long long)n_acc);
    cudaStream_t st = (cudaStream_t)stream;
    AuxLane* lane = nullptr;
    if (int rc = aux_lane(&lane)) return rc;
    int sm = 148;
    { int dev = 0; A3R_CUDA(cudaGetDevice(&dev)); A3R_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev)); }
    int grid = (int)((n_acc + 255) / 256);
    if (grid > sm) grid = sm;
    bool pending = false;
    if (int rc = zero_small_scratch(p, st)) return rc;
    for (int32_t k = 0; k < count; ++k) {
        const uint32_t tag = tag0 + 1u + (uint32_t)k;
        A3rProblem q = *p;                               // the scan adds into this rank's shared partial-sum buffer
        double* part = reinterpret_cast<double*>(static_cast<char*>(peers->base[peers->rank]) + PEER_HEADER) + (size_t)(tag & 1u) * (size_t)n_acc;
        q.acc_slot = part;
        q.acc_frame = part + (int64_t)2 * p->n_edges * A3R_SLOT_ACC;
        if (prior) q.prior_acc = q.acc_frame + (int64_t)p->n_frames * A3R_FRAME_ACC;
        if (k == 0)
            if (int rc = a3r_ga_prepare(p, modes[k], stream)) return rc;             // later ones: the tail of K2
        if (pending) A3R_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
        if (int rc = a3r_ga_accumulate(&q, modes[k], stream)) return rc;
        k_exchange<<<grid, 256, 0, st>>>(*peers, tag, p->acc_slot);
        A3R_CUDA(cudaGetLastError());
        if (int rc = update_forked(p, modes[k], &steps[k], st, lane, &pending, k + 1 < count)) return rc;
    }
    if (pending) A3R_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
    return A3R_OK;
}

// diagnostic: switch the K2 phase trace on / off and read the stamps of the last launch (nanoseconds of %globaltimer;
// [cta][8]: start, cameras, loss + gate, chains, barrier 1, phase B, barrier 2, end).  Synchronises the device.
int a3r_debug_k2_trace(int enable, unsigned long long* out) {
    A3R_CUDA(cudaDeviceSynchronize());
    if (out) A3R_CUDA(cudaMemcpyFromSymbol(out, g_k2_trace, sizeof(unsigned long long) * K2_CLUSTER * 8));
    A3R_CUDA(cudaMemcpyToSymbol(g_k2_trace_on, &enable, sizeof(int)));
    return A3R_OK;
}

int a3r_ga_unproject(const A3rProblem* p, float* pts3d, void* stream) {
    if (int rc = check_problem(p)) return rc;
    if (!pts3d) return a3r_fail(A3R_E_ARG, "pts3d is NULL");
    if (p->n_local_frames == 0) return A3R_OK;
    dim3 grid(64, (unsigned)p->n_local_frames);
    k_unproject<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, pts3d);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

}  // extern "C"
