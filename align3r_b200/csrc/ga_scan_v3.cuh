// ga_scan_v3.cuh - K1 (v3): the fused forward+backward scan with NO cross-warp hand-offs.
//
// Every warp is an independent stream: it owns a contiguous range of 128-pixel tiles (4 pixels per lane, state
// in registers), and for each tile walks the slots living on that frame.  The (X,Y,Z,w) + flow chunk of every
// (tile, slot) - 3 KB - lands in a private 3-stage shared-memory ring, three chunks ahead of the one being
// computed, through one of two copy engines (template parameter TMA):
//   TMA  = true : lane 0 issues cp.async.bulk (TMA bulk copies, 2 KB + 1 KB), completion on the warp's own mbarrier;
//   TMA  = false: every lane issues cp.async (LDGSTS) for exactly the pixels it will read - completion is per lane
//                 (cp.async.wait_group): no mbarrier, no elected issuer, no cross-lane hand-off at all.
// Measured on C2 (B200): flow-active scan 0.957 ms (TMA) vs 0.963 ms (LDGSTS); align-only 0.609 vs 0.550 ms - the
// TMA elect / uniform-register issue sequence + try_wait loop is ~80 warp instructions per chunk, which matters when a
// chunk is only ~330 instructions of work.  The launcher picks TMA for the flow-active scan, LDGSTS otherwise.
// Per (tile, slot) the 27 per-lane partial sums are combined with a transposing butterfly (31 shuffles: lane l
// ends up holding the warp total of sum l) and added to the warp's private shared accumulators; fp64 global
// atomics happen once per (warp, work frame).  16 independent warps per SM hide each other's latencies; there
// is no barrier, no producer/consumer role and no wake-up latency on any critical path.
#pragma once
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include "ga_scan_v2.cuh"     // PTX helpers (mbarrier, bulk copy, raw shared-memory accesses)

namespace a3r {

#ifndef A3R_V3_T
#define A3R_V3_T 128
#endif
constexpr int V3_T = A3R_V3_T;                // pixels per tile (V3_T / 32 per lane)
constexpr int V3_PPL = V3_T / 32;
#ifndef A3R_V3_WARPS
#define A3R_V3_WARPS 12
#endif
constexpr int V3_WARPS = A3R_V3_WARPS;
constexpr int V3_THREADS = V3_WARPS * 32;
#ifndef A3R_V3_STAGES
#define A3R_V3_STAGES 3
#endif
constexpr int V3_STAGES = A3R_V3_STAGES;
constexpr int V3_SLOTS = 24;                  // max slots per work frame
constexpr int V3_TR = 28;                     // row stride (floats) of the transpose buffer: 28/4 odd -> float4 rows of
                                              // 8 consecutive lanes fall into distinct 16-byte bank groups

struct V3Smem {
    float4 xw[V3_WARPS][V3_STAGES][V3_T];     // 98304 B
    float2 fl[V3_WARPS][V3_STAGES][V3_T];     // 49152 B
    float wacc[V3_WARPS][V3_SLOTS][32];       // per-warp running sums (column l = sum l)
    float tr[V3_WARPS][32 * V3_TR + 4];       // per-warp transpose buffer of the cross-lane reduction
    unsigned long long full[V3_WARPS][V3_STAGES];   // TMA completion barriers (TMA = true only)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Transposing butterfly: in: a[0..31] per lane (entries >= NV are structurally zero); out: the warp-wide total of
// entry `lane`.  log2(32) exchange steps, the vector halves at every step: 16+8+4+2+1 = 31 shuffles.
template <int NV>
__device__ __forceinline__ float warp_transpose_sum(float (&a)[32], int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const bool upper = lane & d;
#pragma unroll
        for (int k = 0; k < d; ++k) {
            // after the previous steps entry k of the current vector stands for original entries k, k+2d, ... folded;
            // entries whose whole index class is >= NV are zero: skip their arithmetic at compile time
            const bool lo_live = k < NV, hi_live = (k + d) < NV;
            if (!lo_live && !hi_live) continue;
            const float lo = a[k], hi = hi_live ? a[k + d] : 0.f;
            const float send = upper ? lo : hi;
            const float keep = upper ? hi : lo;
            a[k] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
    return a[0];
}

template <bool FLOW, bool L2D, bool TMA>
__global__ void __launch_bounds__(V3_THREADS, 1) k_accumulate_v3(const A3rProblem p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V3Smem& sm = *reinterpret_cast<V3Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * V3_WARPS + warp;                 // this warp's range in the plan ("CTA" = warp here)
    const int t0 = p.cta_tile_begin[gw], t1 = p.cta_tile_begin[gw + 1];
    if (t0 >= t1) return;
    constexpr int NV = FLOW ? ACC_N : ACC_GA;
    const float thre_eff = flow_thre_eff(p.pxl_thre);
    const float4* xw_g = reinterpret_cast<const float4*>(p.predw);
    const float2* fl_g = reinterpret_cast<const float2*>(p.flow);

    const uint32_t xw_s = smem_u32(&sm.xw[warp][0][0]);
    const uint32_t fl_s = smem_u32(&sm.fl[warp][0][0]);
    const uint32_t tr_s = smem_u32(&sm.tr[warp][0]);
    const uint32_t bar_s = smem_u32(&sm.full[warp][0]);
    constexpr uint32_t kBytes = V3_T * 16 + (FLOW ? V3_T * 8 : 0);
    float* wacc = &sm.wacc[warp][0][0];

    if (TMA && lane == 0) {
#pragma unroll
        for (int s = 0; s < V3_STAGES; ++s) mbar_init(&sm.full[warp][s], 1);
        fence_mbar_init();
    }
    for (int k = lane; k < V3_SLOTS * 32; k += 32) wacc[k] = 0.f;
    __syncwarp();

    // ---- load cursor (used by lane 0): next chunk = (l_tile, l_s) ----
    int l_tile = t0, l_wf = 0, l_wf_end = -1, l_s = 0, l_deg = 0, l_sbegin = 0, l_tile0 = 0;
    uint32_t l_stage = 0;
#define A3R_V3_ISSUE()                                                                                           \
    do {                                                                                                         \
        if (l_tile < t1) {                                                                                       \
            if (l_tile >= l_wf_end) {                                                                            \
                while (l_tile >= p.wf_tile_begin[l_wf + 1]) ++l_wf;                                              \
                l_wf_end = p.wf_tile_begin[l_wf + 1];                                                            \
                l_tile0 = p.wf_tile_begin[l_wf];                                                                 \
                l_sbegin = p.wf_slot_begin[l_wf];                                                                \
                l_deg = p.wf_slot_begin[l_wf + 1] - l_sbegin;                                                    \
            }                                                                                                    \
            const int64_t off = (int64_t)(l_sbegin + l_s) * p.pix_stride + (int64_t)(l_tile - l_tile0) * V3_T;   \
            if (TMA) {                                                                                           \
                const uint32_t bar_l = bar_s + l_stage * 8;                                                      \
                mbar_expect_tx_a(bar_l, kBytes);                                                                 \
                tma_load_1d_a(xw_s + l_stage * (V3_T * 16), xw_g + off, V3_T * 16, bar_l);                       \
                if (FLOW) tma_load_1d_a(fl_s + l_stage * (V3_T * 8), fl_g + off, V3_T * 8, bar_l);               \
            } else {                                                                                             \
                const uint32_t xd = xw_s + l_stage * (V3_T * 16) + lane * 16;                                    \
                const uint32_t fd = fl_s + l_stage * (V3_T * 8) + lane * 8;                                      \
                _Pragma("unroll") for (int j = 0; j < V3_PPL; ++j) {                                             \
                    cp_async16(xd + j * 512, xw_g + off + lane + j * 32);                                        \
                    if (FLOW) cp_async8(fd + j * 256, fl_g + off + lane + j * 32);                               \
                }                                                                                                \
            }                                                                                                    \
            l_stage = l_stage == V3_STAGES - 1 ? 0u : l_stage + 1;                                               \
            if (++l_s >= l_deg) { l_s = 0; ++l_tile; }                                                           \
        }                                                                                                        \
        if (!TMA) cp_async_commit();   /* one group per chunk slot, empty past the end: uniform wait_group accounting */ \
    } while (0)

    if (!TMA || lane == 0) {          // TMA: one elected issuer; LDGSTS: every lane copies its own pixels
#pragma unroll
        for (int k = 0; k < V3_STAGES; ++k) A3R_V3_ISSUE();
    }

    // ---- compute cursor ----
    int cw = 0, cw_end = -1, lf = 0, n = 0, P = 0, W = 1, deg = 0, s_begin = 0, accum = 0, tile0 = 0;
    uint32_t stage = 0, phase = 0;              // ring position of the chunk being consumed; TMA: parity bit per stage
    float facc[FACC_N];
#pragma unroll
    for (int k = 0; k < FACC_N; ++k) facc[k] = 0.f;

    auto flush = [&]() {                        // this warp's share of a work frame is complete
        __syncwarp();
        for (int s = 0; s < deg; ++s) {
            const float v = wacc[s * 32 + lane];
            wacc[s * 32 + lane] = 0.f;
            if (lane < NV && v != 0.f) atomicAdd(&p.acc_slot[(int64_t)p.lslot_gslot[s_begin + s] * ACC_STRIDE + lane], (double)v);
        }
#pragma unroll
        for (int k = 0; k < FACC_N; ++k) {
            const float v = warp_sum(facc[k]);
            if (lane == k) atomicAdd(&p.acc_frame[(int64_t)n * FACC_N + k], (double)v);
            facc[k] = 0.f;
        }
        __syncwarp();
    };

    for (int tile = t0; tile < t1; ++tile) {
        if (tile >= cw_end) {
            if (cw_end >= 0) flush();
            while (tile >= p.wf_tile_begin[cw + 1]) ++cw;
            cw_end = p.wf_tile_begin[cw + 1];
            tile0 = p.wf_tile_begin[cw];
            lf = p.wf_lframe[cw];
            n = p.lframe_id[lf];
            W = p.frame_hw[2 * n + 1];
            P = p.frame_hw[2 * n] * W;
            s_begin = p.wf_slot_begin[cw];
            deg = p.wf_slot_begin[cw + 1] - s_begin;
            accum = p.wf_accum[cw];
        }
        const int p0 = (tile - tile0) * V3_T;
        const int64_t pix_base = (int64_t)lf * p.pix_stride + p0;
        const float* fp = p.frame_par + FP_STRIDE * n;

        // ---- per-pixel state (4 pixels per lane) ----
        float zx[V3_PPL], zy[V3_PPL], zz[V3_PPL], pu[V3_PPL], pv[V3_PPL], sdisp[V3_PPL], ldv[V3_PPL];
        float gsx[V3_PPL], gsy[V3_PPL], gsz[V3_PPL], gfi[V3_PPL], gfj[V3_PPL];
        {
            float fpr[FP_STRIDE];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(fp) + k);
                fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
            }
#pragma unroll
            for (int j = 0; j < V3_PPL; ++j) {
                const int pl = j * 32 + lane, pix = p0 + pl;
                const bool valid = pix < P;
                ldv[j] = valid ? p.logdepth[pix_base + pl] : 0.f;
                bool stat = false;
                if (FLOW) stat = valid && p.static_mask[pix_base + pl] != 0;
                const int v = pix / W, u = pix - v * W;
                PixState s;
                pixel_state(fpr, ldv[j], u, v, s);
                zx[j] = s.zx; zy[j] = s.zy; zz[j] = s.zz; pu[j] = s.u; pv[j] = s.v;
                sdisp[j] = stat ? s.disp : -s.disp;
                gsx[j] = gsy[j] = gsz[j] = gfi[j] = gfj[j] = 0.f;
            }
        }

        // ---- slots ----
        for (int s = 0; s < deg; ++s) {
            const int gs = p.lslot_gslot[s_begin + s];
            SlotConst c;
            {
                const float4* ep = reinterpret_cast<const float4*>(p.edge_par + EP_STRIDE * (gs >> 1));
                const float4 e0 = __ldg(ep), e1 = __ldg(ep + 1), e2 = __ldg(ep + 2);
                c.nM[0] = -e0.x; c.nM[1] = -e0.y; c.nM[2] = -e0.z; c.nM[3] = -e0.w;      // (negations fold into
                c.nM[4] = -e1.x; c.nM[5] = -e1.y; c.nM[6] = -e1.z; c.nM[7] = -e1.w;      //  the FFMA operand modifiers)
                c.nM[8] = -e2.x; c.st[0] = e2.y; c.st[1] = e2.z; c.st[2] = e2.w;
                if (FLOW) {
                    const float4* sp = reinterpret_cast<const float4*>(p.slot_par + SP_STRIDE * gs);
                    const float4 q0 = __ldg(sp), q1 = __ldg(sp + 1), q2 = __ldg(sp + 2);
                    c.A[0] = q0.x; c.A[1] = q0.y; c.A[2] = q0.z; c.A[3] = q0.w;
                    c.A[4] = q1.x; c.A[5] = q1.y; c.A[6] = q1.z; c.A[7] = q1.w;
                    c.A[8] = q2.x; c.b[0] = q2.y; c.b[1] = q2.z; c.b[2] = q2.w;
                }
                c.wn = (gs & 1) ? p.inv_area_j : p.inv_area_i;
            }
            const bool side_j = gs & 1;
            float acc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.f;

            if (TMA) {
                mbar_wait_a(bar_s + stage * 8, (phase >> stage) & 1);
                phase ^= 1u << stage;
            } else {
                cp_async_wait<V3_STAGES - 1>();       // this lane's copies of the oldest chunk have landed
            }
            const uint32_t xa = xw_s + stage * (V3_T * 16) + lane * 16;
            const uint32_t fa = fl_s + stage * (V3_T * 8) + lane * 8;
#pragma unroll
            for (int j = 0; j < V3_PPL; ++j) {
                const float4 X = lds128(xa + j * 512);
                PixState st;
                st.zx = zx[j]; st.zy = zy[j]; st.zz = zz[j]; st.u = pu[j]; st.v = pv[j]; st.disp = fabsf(sdisp[j]);
                float gx, gy, gz;
                align_px<L2D>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                gsx[j] += gx; gsy[j] += gy; gsz[j] += gz;
                if (FLOW) {
                    const float2 F = lds64(fa + j * 256);
                    const float gd = flow_px(c, st, F.x, F.y, thre_eff, sdisp[j] > 0.f, acc);
                    if (side_j) gfj[j] += gd; else gfi[j] += gd;
                }
            }
            // the chunk is consumed: refill its stage (LDGSTS: a lane only ever reads what it copied itself)
            if (TMA) {
                __syncwarp();
                if (lane == 0) A3R_V3_ISSUE();
            } else {
                A3R_V3_ISSUE();
            }
            stage = stage == V3_STAGES - 1 ? 0u : stage + 1;
            // 27 (13) per-lane partial sums -> warp totals through a shared-memory transpose: every lane stores its
            // row as float4, then lane l adds up column l (4 independent chains) into the warp's running sums
            {
                const uint32_t row = tr_s + lane * (V3_TR * 4);
#pragma unroll
                for (int q = 0; q < (NV + 3) / 4; ++q) sts128(row + q * 16, acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
                __syncwarp();
                const uint32_t col = tr_s + lane * 4;
                float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    t0 += lds32(col + (r + 0) * (V3_TR * 4));
                    t1 += lds32(col + (r + 1) * (V3_TR * 4));
                    t2 += lds32(col + (r + 2) * (V3_TR * 4));
                    t3 += lds32(col + (r + 3) * (V3_TR * 4));
                }
                wacc[s * 32 + lane] += (t0 + t1) + (t2 + t3);
                __syncwarp();
            }
        }

        // ---- per-pixel gradient planes + per-frame sums ----
        {
            float fpr[FP_STRIDE];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(fp) + k);
                fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
            }
#pragma unroll
            for (int j = 0; j < V3_PPL; ++j) {
                const int pl = j * 32 + lane, pix = p0 + pl;
                if (pix >= P) continue;
                const int v = pix / W, u = pix - v * W;
                PixState s;
                pixel_state(fpr, ldv[j], u, v, s);
                const float ga = pixel_epilogue(fpr, s, gsx[j], gsy[j], gsz[j], facc);
                const int64_t o = pix_base + pl;
                if (accum) {
                    atomicAdd(&p.g_align[o], ga);
                    if (FLOW) { atomicAdd(&p.g_flow_i[o], gfi[j] * s.dfac); atomicAdd(&p.g_flow_j[o], gfj[j] * s.dfac); }
                } else {
                    p.g_align[o] = ga;
                    if (FLOW) { p.g_flow_i[o] = gfi[j] * s.dfac; p.g_flow_j[o] = gfj[j] * s.dfac; }
                }
            }
        }
    }
    flush();
#undef A3R_V3_ISSUE
}

}  // namespace a3r
