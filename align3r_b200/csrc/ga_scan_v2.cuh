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

namespace a3r {

constexpr int V2_T = 128;                      // pixels per tile
constexpr int V2_SLOTS = 20;                   // slot warps == max slots per work frame
constexpr int V2_STAGES = 3;                   // ring depth (tiles)
constexpr int V2_PIXW = V2_T / 32;             // pixel warps
constexpr int V2_THREADS = (V2_SLOTS + V2_PIXW) * 32;   // 768
constexpr int V2_DRAIN = 32;                   // tiles between fp32 -> fp64 drains (128 pixels per lane)

struct V2Smem {
    float4 xw[V2_STAGES][V2_SLOTS][V2_T];      // 122880 B  TMA destination, overwritten in place by (g, g_q.b)
    float2 fl[V2_STAGES][V2_SLOTS][V2_T];      //  61440 B
    float4 pixA[2][V2_T];                      // zx zy zz disp (disp < 0: pixel is dynamic / outside the frame)
    float2 pixB[2][V2_T];                      // u v
    double dacc[V2_SLOTS][ACC_STRIDE];         // fp64 accumulators of the slot warps
    unsigned long long full[V2_STAGES][V2_SLOTS];
    unsigned long long slots_done[V2_STAGES];
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pixel_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(V2_PIXW * 32) : "memory"); }

// walks the work-frame table: position of a tile
struct WfCursor {
    int wf = 0;
    __device__ __forceinline__ void seek(const A3rProblem& p, int tile) {
        while (tile >= p.wf_tile_begin[wf + 1]) ++wf;
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

    if (tid == 0) {
        for (int b = 0; b < V2_STAGES; ++b) {
            for (int s = 0; s < V2_SLOTS; ++s) mbar_init(&sm.full[b][s], 1);
            mbar_init(&sm.slots_done[b], V2_SLOTS);
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
        // pixel warps (+ producer)
        // =====================================================================================
        const int ptid = tid - V2_SLOTS * 32;            // pixel of the tile owned by this thread
        const bool is_producer = (ptid == 0);
        const float4* xw_g = reinterpret_cast<const float4*>(p.predw);
        const float2* fl_g = reinterpret_cast<const float2*>(p.flow);
        const uint32_t bytes = V2_T * 16 + (FLOW ? V2_T * 8 : 0);
        WfCursor cur_load;
        auto issue_tile = [&](int i) {                    // producer: all chunks of local tile i -> stage i % 3
            const int tile = t0 + i, stage = i % V2_STAGES;
            cur_load.seek(p, tile);
            const int s_begin = p.wf_slot_begin[cur_load.wf], deg = p.wf_slot_begin[cur_load.wf + 1] - s_begin;
            const int64_t p0 = (int64_t)(tile - p.wf_tile_begin[cur_load.wf]) * V2_T;
            for (int s = 0; s < deg; ++s) {
                const int64_t off = (int64_t)(s_begin + s) * p.pix_stride + p0;
                mbar_arrive_expect_tx(&sm.full[stage][s], bytes);
                tma_load_1d(&sm.xw[stage][s][0], xw_g + off, V2_T * 16, &sm.full[stage][s]);
                if (FLOW) tma_load_1d(&sm.fl[stage][s][0], fl_g + off, V2_T * 8, &sm.full[stage][s]);
            }
        };
        if (is_producer)
            for (int i = 0; i < V2_STAGES && i < nt; ++i) issue_tile(i);

        float facc[FACC_N];
#pragma unroll
        for (int k = 0; k < FACC_N; ++k) facc[k] = 0.f;
        WfCursor cur_a;
        int wf_a = -1, wfpar_a = 1;          // phase A's work frame and which fp[]/side_j[] copy holds it
        int lf_a = 0, n_a = 0, P_a = 0, W_a = 1;
        // phase C state (lags one tile): the values phase A had for the previous tile
        PixState ps_c;
        bool valid_c = false;
        int wfpar_c = 0, deg_c = 0, accum_c = 0, n_c = 0;
        int64_t out_c = 0;

        auto flush_frame = [&](int n) {
#pragma unroll
            for (int k = 0; k < FACC_N; ++k) {
                const float v = warp_sum(facc[k]);
                if (lane == k) atomicAdd(&p.acc_frame[(int64_t)n * FACC_N + k], (double)v);
                facc[k] = 0.f;
            }
        };

        for (int i = 0; i <= nt; ++i) {
            PixState ps_a;
            bool valid_a = false;
            int deg_a = 0, accum_a = 0;
            int64_t out_a = 0;
            if (i < nt) {
                // ---------------- phase A of tile i ----------------
                const int tile = t0 + i, pb = i & 1;
                cur_a.seek(p, tile);
                if (cur_a.wf != wf_a) {
                    wf_a = cur_a.wf;
                    wfpar_a ^= 1;
                    lf_a = p.wf_lframe[wf_a];
                    n_a = p.lframe_id[lf_a];
                    W_a = p.frame_hw[2 * n_a + 1];
                    P_a = p.frame_hw[2 * n_a] * W_a;
                    // fp[wfpar_a] was last read by phase C of a tile two work frames ago: all pixel warps are past it
                    pixel_warps_sync();
                    const int s_begin = p.wf_slot_begin[wf_a], deg = p.wf_slot_begin[wf_a + 1] - s_begin;
                    if (ptid < FP_STRIDE) sm.fp[wfpar_a][ptid] = p.frame_par[FP_STRIDE * n_a + ptid];
                    if (ptid < deg) sm.side_j[wfpar_a][ptid] = (p.lslot_gslot[s_begin + ptid] & 1) ? 1.f : 0.f;
                    pixel_warps_sync();
                }
                deg_a = p.wf_slot_begin[wf_a + 1] - p.wf_slot_begin[wf_a];
                accum_a = p.wf_accum[wf_a];
                const int p0 = (tile - p.wf_tile_begin[wf_a]) * V2_T;
                const int pix = p0 + ptid;
                valid_a = pix < P_a;
                out_a = (int64_t)lf_a * p.pix_stride + pix;
                const float ld = valid_a ? p.logdepth[out_a] : 0.f;
                const int v = pix / W_a, u = pix - v * W_a;
                pixel_state(sm.fp[wfpar_a], ld, u, v, ps_a);
                bool stat = false;
                if (FLOW) stat = valid_a && (p.static_mask[out_a] != 0);
                if (i >= 2) mbar_wait(&sm.pix_free[pb], ((i >> 1) - 1) & 1);      // slot warps are done with tile i-2's state
                sm.pixA[pb][ptid] = make_float4(ps_a.zx, ps_a.zy, ps_a.zz, stat ? ps_a.disp : -1.f);
                sm.pixB[pb][ptid] = make_float2(ps_a.u, ps_a.v);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.pix_ready[pb]);
            }
            if (i >= 1) {
                // ---------------- phase C of tile i-1 ----------------
                const int j = i - 1, stage = j % V2_STAGES;
                mbar_wait(&sm.slots_done[stage], (j / V2_STAGES) & 1);
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
                // the stage is free once all pixel warps have read it; then the producer refills it
                fence_proxy_async();
                pixel_warps_sync();
                if (is_producer && j + V2_STAGES < nt) issue_tile(j + V2_STAGES);
                // end of this CTA's share of a frame: hand the per-frame sums over
                if (i == nt || n_a != n_c) flush_frame(n_c);
            }
            // what phase A just computed becomes phase C's tile in the next iteration
            ps_c = ps_a; valid_c = valid_a; wfpar_c = wfpar_a; deg_c = deg_a; accum_c = accum_a; n_c = n_a; out_c = out_a;
        }
        return;
    }

    // =========================================================================================
    // slot warps
    // =========================================================================================
    float acc[ACC_N];
#pragma unroll
    for (int k = 0; k < ACC_N; ++k) acc[k] = 0.f;
    SlotConst c;
    int my_gs = -1, cur_wf = -1, deg = 0, since_drain = 0;
    uint32_t full_phase = 0;                       // one parity bit per stage, for this warp's own full[] barriers
    WfCursor cur;

    auto drain = [&]() {            // fp32 register partial sums -> this warp's fp64 shared accumulators
        if (my_gs >= 0) {
#pragma unroll
            for (int k = 0; k < (FLOW ? ACC_N : ACC_GA); ++k) {
                const float v = warp_sum(acc[k]);
                if (lane == k) sm.dacc[warp][k] += (double)v;
                acc[k] = 0.f;
            }
        }
        since_drain = 0;
    };
    auto flush = [&]() {            // end of this CTA's share of a work frame
        drain();
        __syncwarp();
        if (my_gs >= 0 && lane < (FLOW ? ACC_N : ACC_GA)) {
            const double v = sm.dacc[warp][lane];
            if (v != 0.0) atomicAdd(&p.acc_slot[(int64_t)my_gs * ACC_STRIDE + lane], v);
            sm.dacc[warp][lane] = 0.0;
        }
        __syncwarp();
    };

    for (int i = 0; i < nt; ++i) {
        const int tile = t0 + i, stage = i % V2_STAGES, pb = i & 1;
        cur.seek(p, tile);
        if (cur.wf != cur_wf) {
            if (cur_wf >= 0) flush();
            cur_wf = cur.wf;
            const int s_begin = p.wf_slot_begin[cur_wf];
            deg = p.wf_slot_begin[cur_wf + 1] - s_begin;
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
        mbar_wait(&sm.pix_ready[pb], (i >> 1) & 1);
        if (my_gs >= 0) {
            mbar_wait(&sm.full[stage][warp], (full_phase >> stage) & 1);
            full_phase ^= 1u << stage;
#pragma unroll
            for (int j = 0; j < V2_T / 32; ++j) {
                const int pix = j * 32 + lane;
                const float4 X = sm.xw[stage][warp][pix];
                const float4 A = sm.pixA[pb][pix];
                PixState st;
                st.zx = A.x; st.zy = A.y; st.zz = A.z; st.disp = A.w;
                float gx = 0.f, gy = 0.f, gz = 0.f, gd = 0.f;
                align_px<L2D>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                if (FLOW) {
                    if (A.w > 0.f) {
                        const float2 B = sm.pixB[pb][pix];
                        const float2 F = sm.fl[stage][warp][pix];
                        st.u = B.x; st.v = B.y;
                        gd = flow_px(c, st, F.x, F.y, p.pxl_thre, acc);
                    }
                }
                sm.xw[stage][warp][pix] = make_float4(gx, gy, gz, gd);
            }
            fence_proxy_async();          // these generic-proxy writes precede the next TMA write of this stage
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&sm.slots_done[stage]);
            mbar_arrive(&sm.pix_free[pb]);
        }
        if (++since_drain >= V2_DRAIN) drain();
    }
    flush();
}

}  // namespace a3r
