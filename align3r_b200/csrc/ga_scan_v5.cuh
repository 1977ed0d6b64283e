// ga_scan_v5.cuh - K1 (v5): the fused forward+backward scan with CLASS-SPECIALISED warps.
//
// v4 (ga_scan_v4.cuh) gives every warp the whole job: 4 pixels per lane, per pixel the state of BOTH loss terms (world point
// z and the running sum of dL/dz for the alignment term; pixel coordinates, disparity and the running d loss / d disp for the
// flow term) = 11 registers per pixel, 127 registers per thread at 16 warps per SM.  Its per-chunk overhead (slot constants,
// ring bookkeeping, the transpose reduction of the 27 partial sums: ~160 of 560 warp instructions per 128 pixel-slots) could
// be halved by 8 pixels per lane, but that state does not fit 16 warps' registers, and with fewer warps the issue
// utilisation falls as fast as the instruction count (measured: DESIGN.md section 8).
//
// The two loss terms never meet inside the scan: the alignment term reads (X, w) and the world point, the flow term reads
// the flow vector, the pixel coordinates and the disparity; they write different gradient planes and different accumulator
// entries.  v5 therefore splits them between warps: an ALIGN warp (class A) keeps 6 registers per pixel, a FLOW warp
// (class F) 5 - both walk 8 pixels per lane (256-pixel tiles) at 16 warps per SM.  Each class has its own contiguous tile
// ranges (its own partition of the SAME tile list, balanced by its own cost), its own ring (A: float4 chunks of 4 KB,
// F: float2 chunks of 2 KB) and its own 13 / 14 partial sums; a frame's slots are walked by both classes independently.
// With the flow term off every warp is an ALIGN warp (a third partition).
//
// Per warp everything else is v4's organisation: independent streams, private cp.async rings, the slot table staged per
// work frame, the transpose reduction per chunk, running sums in shared memory, fp64 atomics per (warp, work frame) with
// the CTA-level flush of the last work frame.  The depth Adam step is NOT folded into this scan (a flow warp would need the
// log-depths an align warp of another SM updates): K3 runs after K2.
#pragma once
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include "ptx_helpers.cuh"
#include "ga_scan_v4.cuh"     // sts32 / sts64 / pix_to_uv

namespace a3r {

constexpr int V5_T = 256;                      // pixels per tile
constexpr int V5_PPL = V5_T / 32;              // 8 per lane
constexpr int V5_WARPS = 16;
#ifndef A3R_V5_NA
#define A3R_V5_NA 7
#endif
constexpr int V5_NA = A3R_V5_NA;               // align warps per CTA when the flow term is on (the other 16 - NA are flow warps)
constexpr int V5_NF = V5_WARPS - V5_NA;
#ifndef A3R_V5_STAGES
#define A3R_V5_STAGES 3
#endif
constexpr int V5_STAGES_FLOW = A3R_V5_STAGES;  // ring depth when the flow term is on (7 align + 9 flow blocks: 3 stages fit)
constexpr int V5_STAGES_ALIGN = 2;             // ... when every warp is an align warp (16 of the larger blocks)
constexpr int V5_SLOTS = 20;                   // max slots per work frame
constexpr int V5_ROW = 16;                     // floats per row of the slot table / the running sums
constexpr int V5_TRROW = 20;                   // floats per row of the transpose buffer (80 B: conflict-free 128-bit stores)
constexpr int V5_THREADS = V5_WARPS * 32;

// byte offsets inside a warp's private block; B = bytes per pixel of the class's ring (16: align, 8: flow)
template <int B, int STAGES>
struct V5Layout {
    static constexpr uint32_t RING = 0;
    static constexpr uint32_t STAB = RING + STAGES * V5_T * B;                     // [SLOTS][16]
    static constexpr uint32_t WACC = STAB + V5_SLOTS * V5_ROW * 4;                 // [SLOTS + 1][16] (+ frame row)
    static constexpr uint32_t TR = WACC + (V5_SLOTS + 1) * V5_ROW * 4;             // [32][20] + 16 floats of slack
    static constexpr uint32_t FP = TR + (32 * V5_TRROW + 16) * 4;                  // [16] frame block
    static constexpr uint32_t BYTES = ((FP + 16 * 4 + 127) / 128) * 128;
};
// the CTA's shared memory: flow term on: NA align blocks + NF flow blocks; off: 16 align blocks
constexpr size_t V5_SMEM_FLOW = (size_t)V5Layout<16, V5_STAGES_FLOW>::BYTES * V5_NA + (size_t)V5Layout<8, V5_STAGES_FLOW>::BYTES * V5_NF;
constexpr size_t V5_SMEM_ALIGN = (size_t)V5Layout<16, V5_STAGES_ALIGN>::BYTES * V5_WARPS;
static_assert(V5_SMEM_FLOW <= 232448 - 512 && V5_SMEM_ALIGN <= 232448 - 512, "v5 scan: shared memory of one CTA exceeds the 227 KB opt-in limit");

// partitions of the tile list packed behind each other in A3rProblem.cta_tile_begin (S = n_ctas / 16 CTAs):
//   [0, 16 S]                 every warp an align warp (flow term off)
//   [16 S + 1, ... + NA S]    align warps (flow term on)
//   [... + 1, ... + NF S]     flow warps
__host__ __device__ inline int v5_part_offset(int S, int which) {
    return which == 0 ? 0 : which == 1 ? 16 * S + 1 : 16 * S + 1 + V5_NA * S + 1;
}

// Reduce the per-lane rows (NQ float4 each) through the warp's transpose buffer; lanes 0..7 of both halves return the warp
// totals of columns (2 lane, 2 lane + 1).
template <int NQ>
__device__ __forceinline__ float2 v5_transpose_reduce(uint32_t tr_s, int lane, const float* a) {
    const uint32_t row = tr_s + lane * (V5_TRROW * 4);
#pragma unroll
    for (int q = 0; q < NQ; ++q) sts128(row + q * 16, a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
    __syncwarp();
    const uint32_t col = tr_s + (lane >> 4) * (16 * V5_TRROW * 4) + (lane & 7) * 8;
    float2 t0 = make_float2(0.f, 0.f), t1 = t0;
#pragma unroll
    for (int r = 0; r < 16; r += 2) {
        const float2 x0 = lds64(col + r * (V5_TRROW * 4));
        const float2 x1 = lds64(col + (r + 1) * (V5_TRROW * 4));
        t0.x += x0.x; t0.y += x0.y;
        t1.x += x1.x; t1.y += x1.y;
    }
    t0.x += t1.x; t0.y += t1.y;
    t0.x += __shfl_xor_sync(0xffffffffu, t0.x, 16);
    t0.y += __shfl_xor_sync(0xffffffffu, t0.y, 16);
    return t0;
}

// what a warp ends on (for the CTA-level flush of the last work frame)
struct V5Last { int cw, n, s_begin, deg; };

// --------------------------------------------------------------------------------------------------------------------
// one warp's stream; FLOWCLASS false: alignment term (ring of float4 (X, w)), true: flow term (ring of float2 flow vectors)
// --------------------------------------------------------------------------------------------------------------------
template <bool FLOWCLASS, bool L2D, int V5_STAGES>
__device__ __forceinline__ V5Last v5_warp(const A3rProblem& p, const uint32_t wb, const int t0, const int t1, const int lane) {
    using L = V5Layout<FLOWCLASS ? 8 : 16, V5_STAGES>;
    constexpr int B = FLOWCLASS ? 8 : 16;
    constexpr int NV = FLOWCLASS ? 14 : 13;            // partial sums per slot: ACC_GA.. (flow) / ACC_SG..ACC_LOSS (align)
    constexpr int NQ = 4;
    constexpr int COL0 = FLOWCLASS ? ACC_GA : 0;       // first accumulator entry of this class
    constexpr uint32_t kStage = V5_T * B;
    const float thre_eff = flow_thre_eff(p.pxl_thre);
    const char* src_g = FLOWCLASS ? reinterpret_cast<const char*>(p.flow) : reinterpret_cast<const char*>(p.predw);
    const int64_t stride = p.pix_stride;
    V5Last last{-1, 0, 0, 0};

    for (int k = lane; k < (V5_SLOTS + 1) * V5_ROW; k += 32) sts32(wb + L::WACC + k * 4, 0.f);
    __syncwarp();

    // ---- copy cursor ----
    int l_tile = t0, l_wf = 0, l_wf_end = -1, l_s = 0, l_deg = 1;
    int64_t l_off = 0;
    auto issue = [&](uint32_t stage_off) {
        if (l_tile < t1) {
            if (l_tile >= l_wf_end) {
                while (l_tile >= p.wf_tile_begin[l_wf + 1]) ++l_wf;
                l_wf_end = p.wf_tile_begin[l_wf + 1];
                const int sb = p.wf_slot_begin[l_wf];
                l_deg = p.wf_slot_begin[l_wf + 1] - sb;
                l_off = (int64_t)sb * stride + (int64_t)(l_tile - p.wf_tile_begin[l_wf]) * V5_T;
                l_s = 0;
            }
            const uint32_t d = wb + L::RING + stage_off + lane * B;
            const char* g = src_g + (l_off + lane) * B;
#pragma unroll
            for (int j = 0; j < V5_PPL; ++j) {
                if (FLOWCLASS) cp_async8(d + j * (32 * B), g + j * (32 * B));
                else cp_async16(d + j * (32 * B), g + j * (32 * B));
            }
            if (++l_s >= l_deg) { l_s = 0; ++l_tile; l_off += V5_T - (int64_t)(l_deg - 1) * stride; }
            else l_off += stride;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int k = 0; k < V5_STAGES; ++k) issue(k * kStage);
    pdl_wait();           // everything above reads constant inputs only; from here on: K2's maps, the zeroed accumulator block

    int cw = 0, cw_end = -1, lf = 0, n = 0, P = 0, W = 1, deg = 0, n_i = 0, s_begin = 0, accum = 0, tile0 = 0;
    float invW = 1.f;
    uint32_t stage_off = 0;

    auto flush = [&]() {                        // this warp's share of a work frame is complete (mid-range: the last one is the CTA's)
        __syncwarp();
        float gl_i = 0.f, gl_j = 0.f;
        for (int s = 0; s < deg; ++s) {
            const uint32_t a = wb + L::WACC + (s * V5_ROW + lane) * 4;
            if (lane < NV) {
                const float v = lds32(a);
                sts32(a, 0.f);
                const int gs = p.lslot_gslot[s_begin + s];
                if (v != 0.f) atomicAdd(&p.acc_slot[(int64_t)gs * ACC_STRIDE + COL0 + lane], (double)v);
                if (gs & 1) gl_j += v; else gl_i += v;
            }
        }
        if (!FLOWCLASS && lane == ACC_LOSS) atomicAdd(&p.acc_glob[GL_LOSS], (double)gl_i + (double)gl_j);
        if (FLOWCLASS && (lane == ACC_FNUM - ACC_GA || lane == ACC_FCNT - ACC_GA)) {
            atomicAdd(&p.acc_glob[GL_FNUM_I + (lane - (ACC_FNUM - ACC_GA))], (double)gl_i);
            atomicAdd(&p.acc_glob[GL_FNUM_J + (lane - (ACC_FNUM - ACC_GA))], (double)gl_j);
        }
        if (!FLOWCLASS && lane < FACC_N) {
            const uint32_t a = wb + L::WACC + (V5_SLOTS * V5_ROW + lane) * 4;
            const float v = lds32(a);
            sts32(a, 0.f);
            atomicAdd(&p.acc_frame[(int64_t)n * FACC_N + lane], (double)v);
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
            invW = 1.0f / (float)W;
            s_begin = p.wf_slot_begin[cw];
            deg = p.wf_slot_begin[cw + 1] - s_begin;
            accum = p.wf_accum[cw];
            if (lane < 4) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.frame_par + FP_STRIDE * n) + lane);
                sts128(wb + L::FP + lane * 16, v4.x, v4.y, v4.z, v4.w);
            }
            // slot table of this class: align -M (9), s t (3), 1/area ; flow A (9), b (3)
            for (int k = lane; k < deg * 4; k += 32) {
                const int s = k >> 2, q = k & 3;
                const int gs = p.lslot_gslot[s_begin + s];
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < 3) {
                    if (FLOWCLASS) v4 = __ldg(reinterpret_cast<const float4*>(p.slot_par + SP_STRIDE * gs) + q);
                    else {
                        v4 = __ldg(reinterpret_cast<const float4*>(p.edge_par + EP_STRIDE * (gs >> 1)) + q);
                        v4.x = -v4.x;
                        if (q < 2) { v4.y = -v4.y; v4.z = -v4.z; v4.w = -v4.w; }
                    }
                } else {
                    v4.x = (gs & 1) ? p.inv_area_j : p.inv_area_i;
                }
                sts128(wb + L::STAB + (s * V5_ROW + q * 4) * 4, v4.x, v4.y, v4.z, v4.w);
            }
            {
                const bool is_i = lane < deg && (p.lslot_gslot[s_begin + (lane < deg ? lane : 0)] & 1) == 0;
                const unsigned mi = __ballot_sync(0xffffffffu, is_i);
                n_i = __popc(mi);
                if (FLOWCLASS && mi != (n_i >= 32 ? 0xffffffffu : (1u << n_i) - 1u)) __trap();   // plan contract: side-i slots first
            }
            __syncwarp();
        }
        const int p0 = (tile - tile0) * V5_T;
        const int64_t pix_base = (int64_t)lf * stride + p0;

        if (!FLOWCLASS) {
            // ================= alignment term: world points in, dL/dz summed over the frame's slots =================
            float zx[V5_PPL], zy[V5_PPL], zz[V5_PPL], gsx[V5_PPL], gsy[V5_PPL], gsz[V5_PPL];
            {
                float fpr[FP_STRIDE];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v4 = lds128(wb + L::FP + k * 16);
                    fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
                }
#pragma unroll
                for (int j = 0; j < V5_PPL; ++j) {
                    const int pl = j * 32 + lane, pix = p0 + pl;
                    const float ld = pix < P ? p.logdepth[pix_base + pl] : 0.f;
                    int u, v;
                    pix_to_uv(pix, W, invW, u, v);
                    PixState s;
                    pixel_state(fpr, ld, u, v, s);
                    zx[j] = s.zx; zy[j] = s.zy; zz[j] = s.zz;
                    gsx[j] = gsy[j] = gsz[j] = 0.f;
                }
            }
#pragma unroll 1
            for (int s = 0; s < deg; ++s) {
                const uint32_t tab = wb + L::STAB + s * (V5_ROW * 4);
                cp_async_wait<V5_STAGES - 1>();
                float acc[4 * NQ];
#pragma unroll
                for (int k = 0; k < 4 * NQ; ++k) acc[k] = 0.f;
                const uint32_t xa = wb + L::RING + stage_off + lane * 16;
                SlotConst c;
                {
                    const float4 q0 = lds128(tab), q1 = lds128(tab + 16), q2 = lds128(tab + 32), q3 = lds128(tab + 48);
                    c.nM[0] = q0.x; c.nM[1] = q0.y; c.nM[2] = q0.z; c.nM[3] = q0.w;
                    c.nM[4] = q1.x; c.nM[5] = q1.y; c.nM[6] = q1.z; c.nM[7] = q1.w;
                    c.nM[8] = q2.x; c.st[0] = q2.y; c.st[1] = q2.z; c.st[2] = q2.w;
                    c.wn = q3.x;
                }
#pragma unroll
                for (int j = 0; j < V5_PPL; ++j) {
                    const float4 X = lds128(xa + j * 512);
                    PixState st;
                    st.zx = zx[j]; st.zy = zy[j]; st.zz = zz[j];
                    float gx, gy, gz;
                    align_px<L2D>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                    gsx[j] += gx; gsy[j] += gy; gsz[j] += gz;
                }
                issue(stage_off);
                stage_off = stage_off == (V5_STAGES - 1) * kStage ? 0u : stage_off + kStage;
                const float2 t = v5_transpose_reduce<NQ>(wb + L::TR, lane, acc);
                if (lane < (NV + 1) / 2) {
                    const uint32_t a = wb + L::WACC + s * (V5_ROW * 4) + lane * 8;
                    const float2 o = lds64(a);
                    sts64(a, o.x + t.x, o.y + t.y);
                }
                __syncwarp();
            }
            // ---- alignment gradient plane + per-frame sums ----
            {
                float fpr[FP_STRIDE];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v4 = lds128(wb + L::FP + k * 16);
                    fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
                }
                float facc[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) facc[k] = 0.f;
#pragma unroll
                for (int j = 0; j < V5_PPL; ++j) {
                    const int pl = j * 32 + lane, pix = p0 + pl;
                    if (pix >= P) continue;
                    int u, v;
                    pix_to_uv(pix, W, invW, u, v);
                    PixState s;
                    pixel_state(fpr, p.logdepth[pix_base + pl], u, v, s);      // (re-read: cheaper than 24 more registers)
                    const float ga = pixel_epilogue(fpr, s, gsx[j], gsy[j], gsz[j], facc);
                    const int64_t o = pix_base + pl;
                    if (accum) atomicAdd(&p.g_align[o], ga); else p.g_align[o] = ga;
                }
                const float2 t = v5_transpose_reduce<3>(wb + L::TR, lane, facc);
                if (lane < FACC_N / 2) {
                    const uint32_t a = wb + L::WACC + V5_SLOTS * (V5_ROW * 4) + lane * 8;
                    const float2 o = lds64(a);
                    sts64(a, o.x + t.x, o.y + t.y);
                }
                __syncwarp();
            }
        } else {
            // ================= flow term: pixel coordinates + disparity in, d loss / d disp per flow class out =================
            float pu[V5_PPL], pv[V5_PPL], sdisp[V5_PPL], gf[V5_PPL];
            {
#pragma unroll
                for (int j = 0; j < V5_PPL; ++j) {
                    const int pl = j * 32 + lane, pix = p0 + pl;
                    const bool valid = pix < P;
                    const float ld = valid ? p.logdepth[pix_base + pl] : 0.f;
                    const bool stat = valid && p.static_mask[pix_base + pl] != 0;
                    int u, v;
                    pix_to_uv(pix, W, invW, u, v);
                    const float disp = fast_rcp(expf(ld) + 1e-6f);            // as pixel_state
                    pu[j] = (float)u; pv[j] = (float)v;
                    sdisp[j] = stat ? disp : -disp;
                    gf[j] = 0.f;
                }
            }
#pragma unroll 1
            for (int cls = 0; cls < 2; ++cls) {
                const int s_lo = cls ? n_i : 0, s_hi = cls ? deg : n_i;
#pragma unroll 1
                for (int s = s_lo; s < s_hi; ++s) {
                    const uint32_t tab = wb + L::STAB + s * (V5_ROW * 4);
                    cp_async_wait<V5_STAGES - 1>();
                    float acc[4 * NQ];
#pragma unroll
                    for (int k = 0; k < 4 * NQ; ++k) acc[k] = 0.f;
                    const uint32_t fa = wb + L::RING + stage_off + lane * 8;
                    SlotConst c;
                    {
                        const float4 q0 = lds128(tab), q1 = lds128(tab + 16), q2 = lds128(tab + 32);
                        c.A[0] = q0.x; c.A[1] = q0.y; c.A[2] = q0.z; c.A[3] = q0.w;
                        c.A[4] = q1.x; c.A[5] = q1.y; c.A[6] = q1.z; c.A[7] = q1.w;
                        c.A[8] = q2.x; c.b[0] = q2.y; c.b[1] = q2.z; c.b[2] = q2.w;
                    }
#pragma unroll
                    for (int j = 0; j < V5_PPL; ++j) {
                        const float2 F = lds64(fa + j * 256);
                        PixState st;
                        st.u = pu[j]; st.v = pv[j]; st.disp = fabsf(sdisp[j]);
                        gf[j] = flow_px<ACC_GA>(c, st, F.x, F.y, thre_eff, sdisp[j] > 0.f, acc, gf[j]);
                    }
                    issue(stage_off);
                    stage_off = stage_off == (V5_STAGES - 1) * kStage ? 0u : stage_off + kStage;
                    const float2 t = v5_transpose_reduce<NQ>(wb + L::TR, lane, acc);
                    if (lane < (NV + 1) / 2) {
                        const uint32_t a = wb + L::WACC + s * (V5_ROW * 4) + lane * 8;
                        const float2 o = lds64(a);
                        sts64(a, o.x + t.x, o.y + t.y);
                    }
                    __syncwarp();
                }
                // this class's d loss / d logdepth plane (un-normalised): d disp / d logdepth = -d disp^2
                float* plane = cls ? p.g_flow_j : p.g_flow_i;
#pragma unroll
                for (int j = 0; j < V5_PPL; ++j) {
                    const int pl = j * 32 + lane;
                    if (p0 + pl < P) {
                        const float disp = fabsf(sdisp[j]);
                        const float d = expf(p.logdepth[pix_base + pl]);            // (re-read: cheaper than 8 more registers)
                        const float v = gf[j] * (-d * disp * disp);
                        if (accum) atomicAdd(&plane[pix_base + pl], v); else plane[pix_base + pl] = v;
                    }
                    gf[j] = 0.f;
                }
            }
        }
    }
    last.cw = cw; last.n = n; last.s_begin = s_begin; last.deg = deg;
    return last;
}

// --------------------------------------------------------------------------------------------------------------------
// the kernel: FLOW false: 16 align warps per CTA (partition 0); true: NA align warps (partition 1) + NF flow warps (partition 2)
// --------------------------------------------------------------------------------------------------------------------
template <bool FLOW, bool L2D>
__global__ void __launch_bounds__(V5_THREADS, 1) k_accumulate_v5(const A3rProblem p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_last[V5_WARPS][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = gridDim.x;
    const bool flow_warp = FLOW && warp >= V5_NA;
    int gw, off;
    if (!FLOW) { gw = blockIdx.x * V5_WARPS + warp; off = v5_part_offset(S, 0); }
    else if (!flow_warp) { gw = blockIdx.x * V5_NA + warp; off = v5_part_offset(S, 1); }
    else { gw = blockIdx.x * V5_NF + (warp - V5_NA); off = v5_part_offset(S, 2); }
    const int t0 = p.cta_tile_begin[off + gw], t1 = p.cta_tile_begin[off + gw + 1];
    constexpr int STG = FLOW ? V5_STAGES_FLOW : V5_STAGES_ALIGN;
    using V5A = V5Layout<16, STG>;
    using V5F = V5Layout<8, STG>;
    // blocks: align warps first (the larger ones), flow warps behind them
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t wb = flow_warp ? base + V5_NA * V5A::BYTES + (warp - V5_NA) * V5F::BYTES : base + warp * V5A::BYTES;
    const TimelineSlot tl = timeline_enter(1);
    V5Last last{-1, 0, 0, 0};
    if (t0 < t1) {
        if (flow_warp) last = v5_warp<true, L2D, STG>(p, wb, t0, t1, lane);
        else last = v5_warp<false, L2D, STG>(p, wb, t0, t1, lane);
    }
    timeline_mark(tl, 2);

    // ---- final flush per CTA and class: the warps of a class that end on the same work frame add their running sums in
    //      shared memory (fixed warp order, fp64) and issue one atomic per value ----
    if (lane == 0) { s_last[warp][0] = last.cw; s_last[warp][1] = last.n; s_last[warp][2] = last.s_begin; s_last[warp][3] = last.deg; }
    __syncthreads();
    if (last.cw >= 0) {
        const int w_lo = flow_warp ? V5_NA : 0, w_hi = (FLOW && !flow_warp) ? V5_NA : V5_WARPS;
        unsigned members = 0;
        for (int w = w_lo; w < w_hi; ++w) members |= (s_last[w][0] == last.cw) ? (1u << w) : 0u;
        const int size = __popc(members), rank = __popc(members & ((1u << warp) - 1u));
        const int nv = flow_warp ? 14 : 13, col0 = flow_warp ? (int)ACC_GA : 0;
        const uint32_t wacc = flow_warp ? V5F::WACC : V5A::WACC;
        auto block_of = [&](int w) {
            return (FLOW && w >= V5_NA) ? base + V5_NA * V5A::BYTES + (w - V5_NA) * V5F::BYTES : base + w * V5A::BYTES;
        };
        double gl_i = 0.0, gl_j = 0.0;
        for (int s = rank; s < last.deg; s += size) {
            if (lane < nv) {
                double v = 0.0;
                for (int w = w_lo; w < w_hi; ++w)
                    if (members >> w & 1u) v += (double)lds32(block_of(w) + wacc + (s * V5_ROW + lane) * 4);
                const int gs = p.lslot_gslot[last.s_begin + s];
                if (v != 0.0) atomicAdd(&p.acc_slot[(int64_t)gs * ACC_STRIDE + col0 + lane], v);
                if (gs & 1) gl_j += v; else gl_i += v;
            }
        }
        if (!flow_warp && lane == ACC_LOSS && gl_i + gl_j != 0.0) atomicAdd(&p.acc_glob[GL_LOSS], gl_i + gl_j);
        if (flow_warp && (lane == ACC_FNUM - ACC_GA || lane == ACC_FCNT - ACC_GA)) {
            if (gl_i != 0.0) atomicAdd(&p.acc_glob[GL_FNUM_I + (lane - (ACC_FNUM - ACC_GA))], gl_i);
            if (gl_j != 0.0) atomicAdd(&p.acc_glob[GL_FNUM_J + (lane - (ACC_FNUM - ACC_GA))], gl_j);
        }
        if (!flow_warp && rank == size - 1 && lane < FACC_N) {
            double v = 0.0;
            for (int w = w_lo; w < w_hi; ++w)
                if (members >> w & 1u) v += (double)lds32(block_of(w) + wacc + (V5_SLOTS * V5_ROW + lane) * 4);
            atomicAdd(&p.acc_frame[(int64_t)last.n * FACC_N + lane], v);
        }
    }
    timeline_mark(tl, 3);
}

}  // namespace a3r
