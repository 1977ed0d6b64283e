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
#include "ga_scan_v5.cuh"
#include "ga_small.cuh"     // K0 (parameter maps) and K2 (loss, gate, chains, small Adam, exchange)
#include <stdlib.h>

using namespace a3r;

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
            if (sum != 0.0) {
                atomicAdd(&p.acc_slot[(int64_t)sh.gslot[s] * ACC_STRIDE + i], sum);
                // the whole-problem sums the gate and the loss need (acc_glob)
                if (i == ACC_LOSS) atomicAdd(&p.acc_glob[GL_LOSS], sum);
                else if (i == ACC_FNUM || i == ACC_FCNT) atomicAdd(&p.acc_glob[((sh.gslot[s] & 1) ? GL_FNUM_J : GL_FNUM_I) + (i - ACC_FNUM)], sum);
            }
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
// K3: Adam on the log-depth planes (or, without A3R_M_APPLY_ADAM, total depth gradient -> g_align)
// ==========================================================================================
template <bool FLOW>
__global__ void __launch_bounds__(256, 4) k_update_depth(const A3rProblem p, const uint32_t mode, const A3rAdamStep adam) {
    const int64_t total4 = (int64_t)p.n_local_frames * p.pix_stride / 4;
    const bool prior = p.depth_regularize_weight > 0.f;
    // the flow normalisers and the gate as K2 published them (K3 runs after K2 on the same stream)
    float sc_i = 0.f, sc_j = 0.f;
    if (FLOW) { sc_i = (float)p.scal[A3R_S_SC_I]; sc_j = (float)p.scal[A3R_S_SC_J]; }
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
        if (e && e[0] == 'v' && e[1] == '1') return 1;
        if (e && e[0] == 'v' && e[1] == '5') return 5;
        return 4;
    }();
    return v;
}
static int scan_tile() { const int v = scan_variant(); return v == 1 ? K1_TILE : v == 5 ? V5_T : V4_T; }
static int scan_slots() { const int v = scan_variant(); return v == 1 ? K1_MAXS : v == 5 ? V5_SLOTS : V4_SLOTS; }

static int check_problem(const A3rProblem* p) {
    if (!p) return a3r_fail(A3R_E_ARG, "problem is NULL");
    if (p->tile_pixels != scan_tile()) return a3r_fail(A3R_E_ARG, "tile_pixels %d != a3r_tile_pixels() %d", p->tile_pixels, scan_tile());
    if (p->pix_stride % scan_tile() != 0) return a3r_fail(A3R_E_ARG, "pix_stride must be a multiple of a3r_tile_pixels()");
    if (p->n_frames <= 0 || p->n_edges <= 0 || p->n_local_frames < 0 || p->n_ctas <= 0)
        return a3r_fail(A3R_E_ARG, "bad sizes N=%d E=%d local=%d ctas=%d", p->n_frames, p->n_edges, p->n_local_frames, p->n_ctas);
    if (p->pix_stride % 4 != 0) return a3r_fail(A3R_E_ARG, "pix_stride must be a multiple of 4");
    if (p->acc_frame != p->acc_slot + (int64_t)2 * p->n_edges * A3R_SLOT_ACC)
        return a3r_fail(A3R_E_ARG, "acc_frame must directly follow acc_slot");
    {
        const bool prior = p->depth_regularize_weight > 0.f;
        if (prior && p->prior_acc != p->acc_frame + (int64_t)p->n_frames * A3R_FRAME_ACC)
            return a3r_fail(A3R_E_ARG, "prior_acc must directly follow acc_frame");
        if (p->acc_glob != p->acc_frame + (int64_t)p->n_frames * (A3R_FRAME_ACC + (prior ? A3R_PRIOR_ACC : 0)))
            return a3r_fail(A3R_E_ARG, "acc_glob must close the accumulator block (after prior_acc / acc_frame)");
    }
    if (!p->pair_scratch || !p->cam_cache || !p->inc_begin || !p->inc_items || !p->sync_ctr || !p->frame_grad || !p->edge_scratch || !p->scal)
        return a3r_fail(A3R_E_ARG, "a scratch buffer of the update step is NULL (pair_scratch, cam_cache, inc_begin, inc_items, sync_ctr, ...)");
    return A3R_OK;
}

// cudaFuncSetAttribute acts on the CURRENT device's copy of a kernel: a process that drives several GPUs (rank-emulated runs,
// a caller with one problem per device) must opt every device in, once - one flag per (template instantiation, device)
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("A3R_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

// fold: the scan's tile prologue applies the previous iteration's depth Adam step (prev_adam, prev_flow); NULL: plain scan
template <bool FLOW, bool L2D>
static int launch_accumulate(const A3rProblem* p, cudaStream_t st, const A3rAdamStep* prev_adam, bool prev_flow) {
    const A3rAdamStep none = {0.f, 1.f, -1, 0};
    const bool pdl = pdl_enabled();
    if (scan_variant() == 5) {
        if (prev_adam) return a3r_fail(A3R_E_UNSUPPORTED, "the v5 scan has no folded depth step");
        if (p->n_ctas % V5_WARPS != 0) return a3r_fail(A3R_E_ARG, "n_ctas must be a multiple of %d for the v5 scan", V5_WARPS);
        constexpr size_t smem = FLOW ? V5_SMEM_FLOW : V5_SMEM_ALIGN;
        static PerDeviceOnce configured;
        if (configured.need())
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v5<FLOW, L2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        A3R_CUDA(launch_ex(k_accumulate_v5<FLOW, L2D>, dim3(p->n_ctas / V5_WARPS), dim3(V5_THREADS), smem, st, pdl, *p));
    } else if (scan_variant() == 1) {
        if (prev_adam) return a3r_fail(A3R_E_UNSUPPORTED, "the v1 scan has no folded depth step");
        const size_t smem = sizeof(double) * K1_WARPS * K1_MAXS * ACC_STRIDE + sizeof(K1Shared);
        static PerDeviceOnce configured;   // per template instantiation
        if (configured.need())
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v1<FLOW, L2D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_accumulate_v1<FLOW, L2D><<<p->n_ctas, K1_THREADS, smem, st>>>(*p);
    } else {
        if (p->n_ctas % V4_WARPS != 0) return a3r_fail(A3R_E_ARG, "n_ctas must be a multiple of %d for the v4 scan", V4_WARPS);
        // copy engine of the private rings (A3R_V4_COPY=tma|ldgsts forces one; default: per-lane cp.async)
        static const int forced = [] {
            const char* e = getenv("A3R_V4_COPY");
            return (!e || !e[0]) ? 0 : (e[0] == 't' ? 1 : 2);
        }();
        const bool tma = forced == 1;      // measured on C2: per-lane cp.async 0.760 / 0.474 ms (flow / align), TMA 0.843 / 0.533 ms
        static PerDeviceOnce configured;
        if (configured.need()) {
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            if (FLOW) {
                A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, false, false, FLOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
                A3R_CUDA(cudaFuncSetAttribute(k_accumulate_v4<FLOW, L2D, false, true, FLOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM));
            }
        }
        const unsigned grid = p->n_ctas / V4_WARPS;
        // row tiles (every frame's width a multiple of the tile): the flow arithmetic with v as a per-tile constant
        static const bool rowt_off = [] { const char* e = getenv("A3R_V4_ROWT"); return e && e[0] == '0'; }();
        const bool rowt = FLOW && !tma && !rowt_off && (p->flags & A3R_F_ROW_TILES);
        if (rowt && prev_adam) A3R_CUDA(launch_ex(k_accumulate_v4<FLOW, L2D, false, true, FLOW>, dim3(grid), dim3(V4_THREADS), V4_SMEM, st, pdl, *p, *prev_adam, prev_flow ? 1u : 0u));
        else if (rowt) A3R_CUDA(launch_ex(k_accumulate_v4<FLOW, L2D, false, false, FLOW>, dim3(grid), dim3(V4_THREADS), V4_SMEM, st, pdl, *p, none, 0u));
        else if (prev_adam) A3R_CUDA(launch_ex(k_accumulate_v4<FLOW, L2D, false, true>, dim3(grid), dim3(V4_THREADS), V4_SMEM, st, pdl, *p, *prev_adam, prev_flow ? 1u : 0u));
        else if (tma) k_accumulate_v4<FLOW, L2D, true, false><<<grid, V4_THREADS, V4_SMEM, st>>>(*p, none, 0u);
        else A3R_CUDA(launch_ex(k_accumulate_v4<FLOW, L2D, false, false>, dim3(grid), dim3(V4_THREADS), V4_SMEM, st, pdl, *p, none, 0u));
    }
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

// the depth step of an iteration can ride in the NEXT scan's tile prologue when every pixel belongs to exactly one warp
// (no split work frames), the depth planes are trained, the depth prior is off and the product scan (v4, cp.async) runs
static bool depth_step_folds(const A3rProblem* p) {
    static const bool off = [] { const char* e = getenv("A3R_FOLD_DEPTH"); return e && e[0] == '0'; }();
    static const bool tma = [] { const char* e = getenv("A3R_V4_COPY"); return e && e[0] == 't'; }();
    return !off && !tma && scan_variant() == 4 && p->n_work_frames == p->n_local_frames && (p->flags & A3R_F_TRAIN_DEPTH) &&
           !(p->depth_regularize_weight > 0.f) && p->n_tiles > 0;
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

/* v5 scan: align warps per CTA when the flow term is on (cta_tile_begin then holds three partitions, ga_scan_v5.cuh); 0 otherwise */
int a3r_scan_align_warps(void) { return scan_variant() == 5 ? V5_NA : 0; }

int a3r_max_ctas(int device) {
    int sm = 0;
    if (a3r_device_info(device, &sm, nullptr, nullptr) != A3R_OK) return A3R_E_CUDA;
    // v1: 2 resident CTAs per SM; v4: every warp has its own range of the plan
    const int v = scan_variant();
    return v == 1 ? sm * 2 : v == 5 ? sm * V5_WARPS : sm * V4_WARPS;
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
    const int work = 3 * p->n_edges + p->n_frames;
    int grid = (work + K2_THREADS - 1) / K2_THREADS;
    if (grid > 148) grid = 148;
    k_prepare<<<grid, K2_THREADS, 0, (cudaStream_t)stream>>>(*p);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

// doubles in one accumulator block: acc_slot | acc_frame | prior_acc | acc_glob
static int64_t acc_block_doubles(const A3rProblem* p) {
    const bool prior = p->depth_regularize_weight > 0.f;
    return (int64_t)2 * p->n_edges * A3R_SLOT_ACC + (int64_t)p->n_frames * (A3R_FRAME_ACC + (prior ? A3R_PRIOR_ACC : 0)) + A3R_GLOB_ACC;
}

// a copy of the descriptor whose accumulator pointers address `block`
static A3rProblem with_block(const A3rProblem* p, double* block) {
    A3rProblem q = *p;
    const bool prior = p->depth_regularize_weight > 0.f;
    q.acc_slot = block;
    q.acc_frame = block + (int64_t)2 * p->n_edges * A3R_SLOT_ACC;
    q.prior_acc = prior ? q.acc_frame + (int64_t)p->n_frames * A3R_FRAME_ACC : p->prior_acc;
    q.acc_glob = q.acc_frame + (int64_t)p->n_frames * (A3R_FRAME_ACC + (prior ? A3R_PRIOR_ACC : 0));
    return q;
}

// the scan proper; the accumulator block it adds into must be zero (zero_first: a memset ahead of it)
static int accumulate(const A3rProblem* p, uint32_t mode, cudaStream_t st, bool zero_first, const A3rAdamStep* prev_adam = nullptr,
                      bool prev_flow = false) {
    const bool flow = (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
    if (flow && (!p->flow || !p->static_mask)) return a3r_fail(A3R_E_ARG, "flow active but flow/static_mask planes are NULL");
    const bool prior = p->depth_regularize_weight > 0.f;
    if (prior && (!p->logdepth_init || !p->static_mask || !p->prior_acc))
        return a3r_fail(A3R_E_ARG, "depth_regularize_weight > 0 needs logdepth_init, static_mask and prior_acc");
    if (zero_first) A3R_CUDA(cudaMemsetAsync(p->acc_slot, 0, sizeof(double) * (size_t)acc_block_doubles(p), st));
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
    if (flow) return l2 ? launch_accumulate<true, true>(p, st, prev_adam, prev_flow) : launch_accumulate<true, false>(p, st, prev_adam, prev_flow);
    return l2 ? launch_accumulate<false, true>(p, st, prev_adam, prev_flow) : launch_accumulate<false, false>(p, st, prev_adam, prev_flow);
}

static bool flow_on(const A3rProblem* p, uint32_t mode) {
    return (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
}

int a3r_ga_accumulate(const A3rProblem* p, uint32_t mode, void* stream) {
    if (int rc = check_problem(p)) return rc;
    return accumulate(p, mode, (cudaStream_t)stream, true);
}

int a3r_ga_accumulate_folded(const A3rProblem* p, uint32_t mode, const A3rAdamStep* prev_adam, uint32_t prev_mode, void* stream) {
    if (int rc = check_problem(p)) return rc;
    if (!prev_adam) return a3r_fail(A3R_E_ARG, "prev_adam is NULL");
    if (!depth_step_folds(p)) return a3r_fail(A3R_E_UNSUPPORTED, "the depth step does not fold on this problem (split work frames, depth "
                                              "prior, frozen depth maps, A3R_FOLD_DEPTH=0 or a non-default scan)");
    return accumulate(p, mode, (cudaStream_t)stream, true, prev_adam, flow_on(p, prev_mode));
}

// K2: one launch, grid sized by the problem (every CTA resident: the kernel's grid barrier needs it)
static int launch_small(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, cudaStream_t st, bool prepare_next, const AccSrc& src) {
    int dev = 0, sm = 148;
    A3R_CUDA(cudaGetDevice(&dev));
    A3R_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    const int64_t items = (int64_t)3 * p->n_edges + p->n_frames;
    int grid = (int)((items + K2_THREADS - 1) / K2_THREADS);
    if (grid < 8) grid = 8;
    if (grid > sm) grid = sm;
    static const bool late = [] { const char* e = getenv("A3R_K2_TRIGGER"); return e && e[0] == 'l'; }();
    A3R_CUDA(launch_ex(k_small, dim3(grid), dim3(K2_THREADS), 0, st, pdl_enabled(), *p, mode, *adam, (prepare_next ? 1u : 0u) | (late ? 2u : 0u), src));
    return A3R_OK;
}

// K3: Adam on the log-depth planes (or, without A3R_M_APPLY_ADAM, the total depth gradient), after K2 on the same stream
static int launch_depth(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, cudaStream_t st) {
    const bool apply = mode & A3R_M_APPLY_ADAM;
    if (p->n_local_frames <= 0 || (apply && !(p->flags & A3R_F_TRAIN_DEPTH))) return A3R_OK;
    const bool flow = (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
    const int64_t total4 = (int64_t)p->n_local_frames * p->pix_stride / 4;
    int64_t grid = (total4 + 255) / 256;
    if (grid > 148 * 4) grid = 148 * 4;              // one resident wave (62 registers: 4 CTAs per SM)
    if (flow) k_update_depth<true><<<(unsigned)grid, 256, 0, st>>>(*p, mode, *adam);
    else k_update_depth<false><<<(unsigned)grid, 256, 0, st>>>(*p, mode, *adam);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

static AccSrc local_source(const A3rProblem* q, double* zero_block) {
    AccSrc src;
    memset(&src, 0, sizeof(src));
    src.world = 1;
    src.base[0] = q->acc_slot;
    src.zero_block = zero_block;
    src.n_acc = acc_block_doubles(q);
    return src;
}

int a3r_ga_update(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam, void* stream) {
    if (int rc = check_problem(p)) return rc;
    if (!adam) return a3r_fail(A3R_E_ARG, "adam is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    A3R_CUDA(cudaMemsetAsync(p->sync_ctr, 0, sizeof(uint32_t) * 4, st));
    if (int rc = launch_small(p, mode, adam, st, false, local_source(p, nullptr))) return rc;
    return launch_depth(p, mode, adam, st);
}

int a3r_ga_run(const A3rProblem* p, int32_t count, const uint32_t* modes, const A3rAdamStep* steps, void* stream) {
    if (count < 0 || !modes || !steps) return a3r_fail(A3R_E_ARG, "a3r_ga_run: bad arguments");
    if (int rc = check_problem(p)) return rc;
    if (count == 0) return A3R_OK;
    cudaStream_t st = (cudaStream_t)stream;
    A3R_CUDA(cudaMemsetAsync(p->sync_ctr, 0, sizeof(uint32_t) * 4, st));
    if (int rc = a3r_ga_prepare(p, modes[0], stream)) return rc;                     // later ones: the tail of K2
    // two accumulator blocks alternate by iteration parity: K2 of iteration k zeroes the block iteration k+1 adds into
    const size_t block_bytes = sizeof(double) * (size_t)acc_block_doubles(p);
    double* block[2] = {p->acc_slot, p->acc_alt ? p->acc_alt : p->acc_slot};
    A3R_CUDA(cudaMemsetAsync(block[0], 0, block_bytes, st));
    if (p->acc_alt) A3R_CUDA(cudaMemsetAsync(block[1], 0, block_bytes, st));
    // the depth Adam step of iteration k rides in the tile prologue of scan k+1 (the last one runs stand-alone)
    bool pending_depth = false;
    for (int32_t k = 0; k < count; ++k) {
        const A3rProblem q = with_block(p, block[k & 1]);
        const bool fold = pending_depth;
        if (int rc = accumulate(&q, modes[k], st, !p->acc_alt && k > 0, fold ? &steps[k - 1] : nullptr, fold && flow_on(p, modes[k - 1]))) return rc;
        if (int rc = launch_small(&q, modes[k], &steps[k], st, k + 1 < count, local_source(&q, p->acc_alt ? block[(k + 1) & 1] : nullptr))) return rc;
        pending_depth = (modes[k] & A3R_M_APPLY_ADAM) && k + 1 < count && (modes[k + 1] & A3R_M_APPLY_ADAM) && depth_step_folds(p);
        if (!pending_depth)
            if (int rc = launch_depth(&q, modes[k], &steps[k], st)) return rc;
    }
    return A3R_OK;
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU exchange over peer memory
// ------------------------------------------------------------------------------------------------------------
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

int a3r_ga_run_peers(const A3rProblem* p, const A3rPeers* peers, uint32_t tag0, int32_t count, const uint32_t* modes,
                     const A3rAdamStep* steps, void* stream) {
    if (count < 0 || !modes || !steps || !peers) return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: bad arguments");
    if (int rc = check_problem(p)) return rc;
    if (peers->world < 1 || peers->world > A3R_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world)
        return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: world %d rank %d", peers->world, peers->rank);
    const int64_t n_acc = acc_block_doubles(p);
    if (peers->n_acc != n_acc) return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: n_acc %lld != %lld", (long long)peers->n_acc, (long long)n_acc);
    if (peers->win[0] != 0 || peers->win[peers->world] != p->n_frames)
        return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: the frame windows must cover frames 0 .. %d", p->n_frames);
    for (int r = 0; r < peers->world; ++r)
        if (!peers->base[r] || peers->win[r] > peers->win[r + 1]) return a3r_fail(A3R_E_ARG, "a3r_ga_run_peers: bad region / window of rank %d", r);
    if (count == 0) return A3R_OK;
    cudaStream_t st = (cudaStream_t)stream;
    A3R_CUDA(cudaMemsetAsync(p->sync_ctr, 0, sizeof(uint32_t) * 4, st));
    if (int rc = a3r_ga_prepare(p, modes[0], stream)) return rc;                     // later ones: the tail of K2
    auto block_of = [&](int r, uint32_t parity) {
        return reinterpret_cast<double*>(static_cast<char*>(peers->base[r]) + PEER_HEADER) + (size_t)parity * (size_t)n_acc;
    };
    // Both blocks of this rank's region start from zero.  A peer may still be reading this rank's block of the previous
    // call's last iteration (parity tag0 & 1): it is zeroed by this call's first K2 - which runs after the peers'
    // flags of this call's first iteration - and only the block the first scan adds into is cleared here.
    A3R_CUDA(cudaMemsetAsync(block_of(peers->rank, (tag0 + 1u) & 1u), 0, sizeof(double) * (size_t)n_acc, st));
    bool pending_depth = false;
    for (int32_t k = 0; k < count; ++k) {
        const uint32_t tag = tag0 + 1u + (uint32_t)k;
        const A3rProblem q = with_block(p, block_of(peers->rank, tag & 1u));      // the scan adds into this rank's shared block
        const bool fold = pending_depth;
        if (int rc = accumulate(&q, modes[k], st, false, fold ? &steps[k - 1] : nullptr, fold && flow_on(p, modes[k - 1]))) return rc;
        AccSrc src;
        memset(&src, 0, sizeof(src));
        src.world = peers->world;
        src.rank = peers->rank;
        for (int r = 0; r <= peers->world; ++r) src.win[r] = peers->win[r];
        for (int r = 0; r < peers->world; ++r) { src.base[r] = block_of(r, tag & 1u); src.region[r] = peers->base[r]; }
        src.zero_block = block_of(peers->rank, (tag + 1u) & 1u);
        src.n_acc = n_acc;
        src.tag = tag;
        if (int rc = launch_small(&q, modes[k], &steps[k], st, k + 1 < count, src)) return rc;
        pending_depth = (modes[k] & A3R_M_APPLY_ADAM) && k + 1 < count && (modes[k + 1] & A3R_M_APPLY_ADAM) && depth_step_folds(p);
        if (!pending_depth)
            if (int rc = launch_depth(&q, modes[k], &steps[k], st)) return rc;
    }
    return A3R_OK;
}

// diagnostic: switch the K2 phase trace on / off and read the stamps of the last launch (nanoseconds of %globaltimer;
// [cta 0..7][8]: start, gate, phase A1, barrier 1, phase A2, barrier 2, phase C, end).  Synchronises the device.
int a3r_debug_k2_trace(int enable, unsigned long long* out) {
    A3R_CUDA(cudaDeviceSynchronize());
    {       // enable >= 2: stamp only the launch of iteration enable - 2 (a mid-loop launch also writes the next parameter maps)
        const int it = enable >= 2 ? enable - 2 : -1;
        A3R_CUDA(cudaMemcpyToSymbol(g_k2_trace_iter, &it, sizeof(int)));
    }
    if (out) A3R_CUDA(cudaMemcpyFromSymbol(out, g_k2_trace, sizeof(unsigned long long) * 8 * 8));
    A3R_CUDA(cudaMemcpyToSymbol(g_k2_trace_on, &enable, sizeof(int)));
    return A3R_OK;
}

// diagnostic: the last 64 scan / update launches as CTA 0 saw them: rows (kernel id 1 = scan, 2 = update; entry; return of the
// programmatic-launch wait; exit) in %globaltimer nanoseconds, oldest first in out[0..n); returns n in *count.  Synchronises.
int a3r_debug_timeline(unsigned long long* out, int* count) {
    A3R_CUDA(cudaDeviceSynchronize());
    unsigned long long ring[64 * 4];
    unsigned int seq = 0;
    A3R_CUDA(cudaMemcpyFromSymbol(ring, g_timeline, sizeof(ring)));
    A3R_CUDA(cudaMemcpyFromSymbol(&seq, g_timeline_seq, sizeof(seq)));
    const unsigned int n = seq < 64u ? seq : 64u;
    for (unsigned int k = 0; k < n; ++k) memcpy(out + 4 * k, ring + 4 * ((seq - n + k) & 63u), 4 * sizeof(unsigned long long));
    if (count) *count = (int)n;
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
