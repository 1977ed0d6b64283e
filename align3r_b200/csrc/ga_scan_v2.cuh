// ga_scan_v2.cuh - K1 (v2): the fused forward+backward scan as a persistent, warp-specialised,
// TMA-fed kernel for sm_100a.  One CTA per SM, 24 warps, no CTA-wide barrier in steady state.
//
//   slot warps 0..19   warp w owns slot w of the current work frame for the CTA's whole share of that frame,
//                      so its 27 partial sums (sum g, sum g X^T, loss, flow G_A, G_b, ...) live in registers:
//                      no cross-lane reduction per tile.  Per tile it waits for its chunk (2 KB packed
//                      (X,Y,Z,w) + 1 KB flow, landed by TMA in shared memory), evaluates residual + gradient
//                      of 128 pixels (4 per lane) and writes the per-pixel gradient (g, g_q.b) back IN PLACE
//                      over the inputs it just consumed.
//   pixel warps 20..23 one pixel per thread per tile.  phase A: per-pixel state (depth, camera point, world
//                      point, disparity) computed once for all 20 slots, published in a 2-deep buffer.
//                      phase C (one tile behind): sums the 20 in-place gradients of its pixel, writes the
//                      per-pixel gradient planes, keeps the per-frame sums in registers.
//                      Lane 0 of warp 20 is also the producer: cp.async.bulk (TMA bulk copy) of every
//                      (tile, slot) chunk into a 3-stage ring, each chunk completing on its own mbarrier;
//                      up to 184 KB of loads in flight per SM.
//   Hand-offs are mbarriers only: full[stage][slot] (TMA -> slot warp), pix_ready / pix_free (pixel state),
//   slots_done[stage] (slot warps -> phase C).  A stage is refilled after phase C of its tile.
//   fp32 register partial sums are drained into fp64 shared accumulators every V2_DRAIN tiles and into
//   fp64 global atomics once per (CTA, work frame).
#pragma once
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include <stddef.h>

namespace a3r {

#ifndef A3R_V2_T
#define A3R_V2_T 128
#endif
#ifndef A3R_V2_STAGES
#define A3R_V2_STAGES 3
#endif
constexpr int V2_T = A3R_V2_T;                 // pixels per tile (multiple of 32)
constexpr int V2_SLOTS = 20;                   // slot warps == max slots per work frame
constexpr int V2_STAGES = A3R_V2_STAGES;       // ring depth (tiles): one being computed, one held by the lagging
                                               // phase C, V2_STAGES - 2 in flight from HBM
constexpr int V2_AHEAD = V2_STAGES - 2;        // prefetch distance in tiles
constexpr int V2_PIXW = V2_T / 32;             // pixel warps
constexpr int V2_THREADS = (V2_SLOTS + V2_PIXW) * 32;   // 768
constexpr int V2_DRAIN = 32;                   // tiles between fp32 -> fp64 drains (128 pixels per lane)

struct V2Smem {
    float4 xw[V2_STAGES][V2_SLOTS][V2_T];      // 122880 B  TMA destination, overwritten in place by (g, g_q.b)
    float2 fl[V2_STAGES][V2_SLOTS][V2_T];      //  61440 B
    float4 pixA[2][V2_T];                      // zx zy zz +-disp (negative: pixel is dynamic / outside the frame)
    float2 pixB[2][V2_T];                      // u v
    double dacc[V2_SLOTS][ACC_STRIDE];         // fp64 accumulators of the slot warps
    unsigned long long full[V2_STAGES][V2_SLOTS];
    unsigned long long slots_done[V2_STAGES];
    unsigned long long stage_free[V2_STAGES];
    unsigned long long pix_ready[2];
    unsigned long long pix_free[2];
    float fp[2][FP_STRIDE];                    // frame blocks of the two work frames phases A and C may straddle
    float side_j[2][V2_SLOTS];
};

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();       // a lost arrival must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// raw shared-memory accesses: one base register per tile + compile-time immediates per pixel
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t a, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();     // a lost arrival must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pixel_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(V2_PIXW * 32) : "memory"); }

// optional per-phase timestamps of CTA 0 (tuning builds only: -DA3R_V2_TRACE): edge_scratch[(tile*12)+k]
#ifdef A3R_V2_TRACE
#define A3R_TRACE(WHO, I, K)                                                                         \
    do {                                                                                             \
        if (blockIdx.x == A3R_V2_TRACE && (WHO) && (I) >= 100 && (I) < 136)                                      \
            p.edge_scratch[((I) - 100) * 12 + (K)] = __longlong_as_double(clock64() - trace_t0);                 \
    } while (0)
#else
#define A3R_TRACE(WHO, I, K) do { } while (0)
#endif

// walks the work-frame table: position of a tile.  The boundary of the current work frame is cached in a
// register so the steady state costs one compare per tile.
struct WfCursor {
    int wf = 0, end = -1;
    __device__ __forceinline__ bool seek(const A3rProblem& p, int tile) {     // true when the work frame changed
        if (tile < end) return false;
        while (tile >= p.wf_tile_begin[wf + 1]) ++wf;
        end = p.wf_tile_begin[wf + 1];
        return true;
    }
};

template <bool FLOW, bool L2D>
__global__ void __launch_bounds__(V2_THREADS, 1) k_accumulate_v2(const A3rProblem p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V2Smem& sm = *reinterpret_cast<V2Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = p.cta_tile_begin[blockIdx.x], t1 = p.cta_tile_begin[blockIdx.x + 1];
    const int nt = t1 - t0;
    if (nt <= 0) return;
#ifdef A3R_V2_TRACE
    const long long trace_t0 = clock64();
#endif

    if (tid == 0) {
        for (int b = 0; b < V2_STAGES; ++b) {
            for (int s = 0; s < V2_SLOTS; ++s) mbar_init(&sm.full[b][s], 1);
            mbar_init(&sm.slots_done[b], V2_SLOTS);
            mbar_init(&sm.stage_free[b], V2_PIXW);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.pix_ready[b], V2_PIXW);
            mbar_init(&sm.pix_free[b], V2_SLOTS);
        }
        fence_mbar_init();
    }
    for (int idx = tid; idx < V2_SLOTS * ACC_STRIDE; idx += V2_THREADS) (&sm.dacc[0][0])[idx] = 0.0;
    __syncthreads();

    if (warp >= V2_SLOTS) {
        // =====================================================================================
        // pixel warps: phase A of tile i, then phase C of tile i-1
        // =====================================================================================
        const int ptid = tid - V2_SLOTS * 32;            // pixel of the tile owned by this thread
        float facc[FACC_N];
#pragma unroll
        for (int k = 0; k < FACC_N; ++k) facc[k] = 0.f;
        WfCursor cur_a;
        int wf_a = -1, wfpar_a = 1;          // phase A's work frame and which fp[]/side_j[] copy holds it
        int lf_a = 0, n_a = 0, P_a = 0, W_a = 1, deg_a = 0, accum_a = 0, tile0_a = 0;
        // phase C state (lags one tile): what phase A had for the previous tile
        PixState ps_c;
        bool valid_c = false;
        int wfpar_c = 0, deg_c = 0, accum_c = 0, n_c = 0;
        int64_t out_c = 0;
        // software prefetch of the next tile's log-depth / mask (global latency off the critical path).  Within a
        // work frame the next tile is V2_T pixels on; only a work-frame change walks the tables.
        float ld_next = 0.f;
        unsigned char mk_next = 0;
        int pf_wf = 0, pf_end = -1, pf_pix = 0, pf_P = 0;
        int64_t pf_base = 0;
        auto prefetch = [&](int i) {
            if (i >= nt) return;
            const int tile = t0 + i;
            if (tile >= pf_end) {
                while (tile >= p.wf_tile_begin[pf_wf + 1]) ++pf_wf;
                pf_end = p.wf_tile_begin[pf_wf + 1];
                const int lf = p.wf_lframe[pf_wf], n = p.lframe_id[lf];
                pf_P = p.frame_hw[2 * n] * p.frame_hw[2 * n + 1];
                pf_pix = (tile - p.wf_tile_begin[pf_wf]) * V2_T + ptid;
                pf_base = (int64_t)lf * p.pix_stride;
            }
            ld_next = pf_pix < pf_P ? p.logdepth[pf_base + pf_pix] : 0.f;
            if (FLOW) mk_next = pf_pix < pf_P ? p.static_mask[pf_base + pf_pix] : 0;
            pf_pix += V2_T;
        };
        auto flush_frame = [&](int n) {
#pragma unroll
            for (int k = 0; k < FACC_N; ++k) {
                const float v = warp_sum(facc[k]);
                if (lane == k) atomicAdd(&p.acc_frame[(int64_t)n * FACC_N + k], (double)v);
                facc[k] = 0.f;
            }
        };
        prefetch(0);

        for (int i = 0; i <= nt; ++i) {
            PixState ps_a;
            bool valid_a = false;
            int64_t out_a = 0;
            if (i < nt) {
                // ---------------- phase A of tile i ----------------
                const int tile = t0 + i, pb = i & 1;
                A3R_TRACE(ptid == 0, i, 6);
                const float ld = ld_next;
                const unsigned char mk = mk_next;
                if (cur_a.seek(p, tile)) {
                    wf_a = cur_a.wf;
                    wfpar_a ^= 1;
                    lf_a = p.wf_lframe[wf_a];
                    n_a = p.lframe_id[lf_a];
                    W_a = p.frame_hw[2 * n_a + 1];
                    P_a = p.frame_hw[2 * n_a] * W_a;
                    const int s_begin = p.wf_slot_begin[wf_a];
                    deg_a = p.wf_slot_begin[wf_a + 1] - s_begin;
                    accum_a = p.wf_accum[wf_a];
                    tile0_a = p.wf_tile_begin[wf_a];
                    // fp[wfpar_a] was last read by phase C two work frames ago: every pixel warp is past it
                    pixel_warps_sync();
                    if (ptid < FP_STRIDE) sm.fp[wfpar_a][ptid] = p.frame_par[FP_STRIDE * n_a + ptid];
                    if (ptid < deg_a) sm.side_j[wfpar_a][ptid] = (p.lslot_gslot[s_begin + ptid] & 1) ? 1.f : 0.f;
                    pixel_warps_sync();
                }
                prefetch(i + 1);
                const int pix = (tile - tile0_a) * V2_T + ptid;
                valid_a = pix < P_a;
                out_a = (int64_t)lf_a * p.pix_stride + pix;
                const int v = pix / W_a, u = pix - v * W_a;
                pixel_state(sm.fp[wfpar_a], ld, u, v, ps_a);
                const bool stat = FLOW && valid_a && mk != 0;
                if (i >= 2) mbar_wait(&sm.pix_free[pb], ((i >> 1) - 1) & 1);      // slot warps are done with tile i-2's state
                sm.pixA[pb][ptid] = make_float4(ps_a.zx, ps_a.zy, ps_a.zz, stat ? ps_a.disp : -ps_a.disp);
                sm.pixB[pb][ptid] = make_float2(ps_a.u, ps_a.v);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.pix_ready[pb]);
                A3R_TRACE(ptid == 0, i, 7);
            }
            if (i >= 1) {
                // ---------------- phase C of tile i-1 ----------------
                const int j = i - 1, stage = j % V2_STAGES;
                mbar_wait(&sm.slots_done[stage], (j / V2_STAGES) & 1);
                A3R_TRACE(ptid == 0, j, 8);
                float gsx = 0.f, gsy = 0.f, gsz = 0.f, gfi = 0.f, gfj = 0.f;
#pragma unroll 4
                for (int s = 0; s < deg_c; ++s) {
                    const float4 g = sm.xw[stage][s][ptid];
                    gsx += g.x; gsy += g.y; gsz += g.z;
                    if (FLOW) {
                        const float sj = sm.side_j[wfpar_c][s];
                        gfj = fmaf(g.w, sj, gfj);
                        gfi = fmaf(g.w, 1.f - sj, gfi);
                    }
                }
                // the stage may be refilled (by the slot warps' own TMA loads) once all pixel warps have read it
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.stage_free[stage]);
                A3R_TRACE(ptid == 0, j, 9);
                const float ga = pixel_epilogue(sm.fp[wfpar_c], ps_c, gsx, gsy, gsz, facc);
                if (valid_c) {
                    if (accum_c) {
                        atomicAdd(&p.g_align[out_c], ga);
                        if (FLOW) { atomicAdd(&p.g_flow_i[out_c], gfi * ps_c.dfac); atomicAdd(&p.g_flow_j[out_c], gfj * ps_c.dfac); }
                    } else {
                        p.g_align[out_c] = ga;
                        if (FLOW) { p.g_flow_i[out_c] = gfi * ps_c.dfac; p.g_flow_j[out_c] = gfj * ps_c.dfac; }
                    }
                }
                // end of this CTA's share of a frame: hand the per-frame sums over
                if (i == nt || n_a != n_c) flush_frame(n_c);
                A3R_TRACE(ptid == 0, j, 10);
            }
            // what phase A just computed becomes phase C's tile in the next iteration
            ps_c = ps_a; valid_c = valid_a; wfpar_c = wfpar_a; deg_c = deg_a; accum_c = accum_a; n_c = n_a; out_c = out_a;
        }
        return;
    }

    // =========================================================================================
    // slot warps: own TMA loads + phase B
    // =========================================================================================
    float acc[ACC_N];
#pragma unroll
    for (int k = 0; k < ACC_N; ++k) acc[k] = 0.f;
    SlotConst c;
    int my_gs = -1;
    const float4* xw_g = reinterpret_cast<const float4*>(p.predw);
    const float2* fl_g = reinterpret_cast<const float2*>(p.flow);
    constexpr uint32_t kBytes = V2_T * 16 + (FLOW ? V2_T * 8 : 0);
    constexpr int kAcc = FLOW ? ACC_N : ACC_GA;
    const float thre_eff = flow_thre_eff(p.pxl_thre);

    // shared-memory addresses (32-bit, shared state space): everything is base + stage/parity stride + immediates
    // (re-derived from one base register per tile: a handful of integer ops, no persistent registers)
    const uint32_t sm0 = smem_u32(&sm);
#define xw_w (sm0 + (uint32_t)offsetof(V2Smem, xw) + warp * (V2_T * 16))        /* this warp's chunk in stage 0 */
#define fl_w (sm0 + (uint32_t)offsetof(V2Smem, fl) + warp * (V2_T * 8))
#define full_w (sm0 + (uint32_t)offsetof(V2Smem, full) + warp * 8)
#define pixA_l (sm0 + (uint32_t)offsetof(V2Smem, pixA) + lane * 16)
#define pixB_l (sm0 + (uint32_t)offsetof(V2Smem, pixB) + lane * 8)
#define bar_done (sm0 + (uint32_t)offsetof(V2Smem, slots_done))
#define bar_free (sm0 + (uint32_t)offsetof(V2Smem, stage_free))
#define bar_ready (sm0 + (uint32_t)offsetof(V2Smem, pix_ready))
#define bar_pfree (sm0 + (uint32_t)offsetof(V2Smem, pix_free))
    constexpr uint32_t kXwStage = V2_SLOTS * V2_T * 16, kFlStage = V2_SLOTS * V2_T * 8, kFullStage = V2_SLOTS * 8;

    // running state: stage = i % 3, its successor, parities as bits that flip when their ring wraps
    int cw = 0, cw_end = -1;                       // work frame of the tile being computed
    uint32_t stage = 0, nstage = V2_AHEAD % V2_STAGES;   // stage of tile i and of tile i + V2_AHEAD
    uint32_t full_phase = 0;                       // parity bit per stage of this warp's own full[] barriers
    uint32_t free_phase = 0;                       // parity bit per stage of stage_free[] (consumed by lane 0)

    // the tile being loaded is at most a few work frames ahead of the one being computed: the issuing lane
    // re-derives its position from the tables (cached loads, one lane, off the other lanes' critical path)
#define A3R_V2_ISSUE(I, ST)                                                                                      \
    do {                                                                                                         \
        const int tile_l = t0 + (I);                                                                             \
        int lw = cw;                                                                                             \
        while (tile_l >= p.wf_tile_begin[lw + 1]) ++lw;                                                          \
        const int lw_begin = p.wf_tile_begin[lw], lw_sbegin = p.wf_slot_begin[lw];                               \
        const int lw_deg = p.wf_slot_begin[lw + 1] - lw_sbegin;                                                  \
        if (warp < lw_deg) {                                                                                     \
            const int64_t off = (int64_t)(lw_sbegin + warp) * p.pix_stride + (int64_t)(tile_l - lw_begin) * V2_T; \
            const uint32_t bar_l = full_w + (ST) * kFullStage;                                                   \
            mbar_expect_tx_a(bar_l, kBytes);                                                                     \
            tma_load_1d_a(xw_w + (ST) * kXwStage, xw_g + off, V2_T * 16, bar_l);                                 \
            if (FLOW) tma_load_1d_a(fl_w + (ST) * kFlStage, fl_g + off, V2_T * 8, bar_l);                        \
        }                                                                                                        \
    } while (0)

#define A3R_V2_DRAIN()                                                                                           \
    do {                                                                                                         \
        if (my_gs >= 0) {                                                                                        \
            _Pragma("unroll") for (int k = 0; k < kAcc; ++k) {                                                   \
                const float v = warp_sum(acc[k]);                                                                \
                if (lane == k) sm.dacc[warp][k] += (double)v;                                                    \
                acc[k] = 0.f;                                                                                    \
            }                                                                                                    \
        }                                                                                                        \
    } while (0)

#define A3R_V2_FLUSH()                                                                                           \
    do {                                                                                                         \
        A3R_V2_DRAIN();                                                                                          \
        __syncwarp();                                                                                            \
        if (my_gs >= 0 && lane < kAcc) {                                                                         \
            const double v = sm.dacc[warp][lane];                                                                \
            if (v != 0.0) atomicAdd(&p.acc_slot[(int64_t)my_gs * ACC_STRIDE + lane], v);                         \
            sm.dacc[warp][lane] = 0.0;                                                                           \
        }                                                                                                        \
        __syncwarp();                                                                                            \
    } while (0)

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < V2_AHEAD; ++k)
            if (k < nt) A3R_V2_ISSUE(k, (uint32_t)k);
    }

    for (int i = 0; i < nt; ++i) {
        const int tile = t0 + i;
        const uint32_t pb = i & 1;
        A3R_TRACE(tid == 0, i, 0);
        if (tile >= cw_end) {
            if (cw_end >= 0) A3R_V2_FLUSH();
            while (tile >= p.wf_tile_begin[cw + 1]) ++cw;
            cw_end = p.wf_tile_begin[cw + 1];
            const int s_begin = p.wf_slot_begin[cw];
            const int deg = p.wf_slot_begin[cw + 1] - s_begin;
            my_gs = -1;
            if (warp < deg) {
                my_gs = p.lslot_gslot[s_begin + warp];
                const float* ep = p.edge_par + EP_STRIDE * (my_gs >> 1);
#pragma unroll
                for (int k = 0; k < 9; ++k) c.nM[k] = -ep[EP_M + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) c.st[k] = ep[EP_ST + k];
                if (FLOW) {
                    const float* sp = p.slot_par + SP_STRIDE * my_gs;
#pragma unroll
                    for (int k = 0; k < 9; ++k) c.A[k] = sp[SP_A + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) c.b[k] = sp[SP_B + k];
                }
                c.wn = (my_gs & 1) ? p.inv_area_j : p.inv_area_i;
            }
        }
        // prefetch: tile i+AHEAD goes into the stage tile i-2 used; phase C of tile i-2 ran during B(i-1), so this
        // wait is already satisfied in steady state.  (B(i) and the lagging phase C(i-1) hold two other stages.)
        // The proxy fence orders this warp's earlier generic-proxy writes to that stage before the TMA write.
        if (lane == 0 && i + V2_AHEAD < nt) {
            if (i >= 2) {
                mbar_wait_a(bar_free + nstage * 8, (free_phase >> nstage) & 1);
                free_phase ^= 1u << nstage;
            }
            fence_proxy_async();
            A3R_V2_ISSUE(i + V2_AHEAD, nstage);
        }
        __syncwarp();
        A3R_TRACE(tid == 0, i, 1);
        mbar_wait_a(bar_ready + pb * 8, (i >> 1) & 1);
        A3R_TRACE(tid == 0, i, 2);
        if (my_gs >= 0) {
            mbar_wait_a(full_w + stage * kFullStage, (full_phase >> stage) & 1);
            A3R_TRACE(tid == 0, i, 3);
            full_phase ^= 1u << stage;
            const uint32_t xa = xw_w + stage * kXwStage + lane * 16;
            const uint32_t fa = fl_w + stage * kFlStage + lane * 8;
            const uint32_t pa = pixA_l + pb * (V2_T * 16);
            const uint32_t pq = pixB_l + pb * (V2_T * 8);
#ifdef A3R_EXP_CHEAP_B
#pragma unroll
            for (int j = 0; j < 1; ++j) {
#else
#pragma unroll
            for (int j = 0; j < V2_T / 32; ++j) {
#endif
                const float4 X = lds128(xa + j * 512);
                const float4 A = lds128(pa + j * 512);
                PixState st;
                st.zx = A.x; st.zy = A.y; st.zz = A.z; st.disp = fabsf(A.w);
                float gx, gy, gz, gd = 0.f;
                align_px<L2D>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                if (FLOW) {
                    const float2 B = lds64(pq + j * 256);
                    const float2 F = lds64(fa + j * 256);
                    st.u = B.x; st.v = B.y;
                    gd = flow_px(c, st, F.x, F.y, thre_eff, A.w > 0.f, acc);
                }
                sts128(xa + j * 512, gx, gy, gz, gd);
            }
        }
        __syncwarp();
        A3R_TRACE(tid == 0, i, 4);
        if (lane == 0) {
            mbar_arrive_a(bar_done + stage * 8);
            mbar_arrive_a(bar_pfree + pb * 8);
        }
        A3R_TRACE(tid == 0, i, 5);
        if ((i & (V2_DRAIN - 1)) == V2_DRAIN - 1) A3R_V2_DRAIN();
        stage = stage == V2_STAGES - 1 ? 0u : stage + 1;
        nstage = nstage == V2_STAGES - 1 ? 0u : nstage + 1;
    }
    A3R_V2_FLUSH();
#undef A3R_V2_ISSUE
#undef A3R_V2_DRAIN
#undef A3R_V2_FLUSH
#undef xw_w
#undef fl_w
#undef full_w
#undef pixA_l
#undef pixB_l
#undef bar_done
#undef bar_free
#undef bar_ready
#undef bar_pfree
}

}  // namespace a3r
