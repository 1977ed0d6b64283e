// ga_scan_v4.cuh - K1 (v4): the fused forward+backward scan (product path).
//
// Every warp is an independent stream: it owns a contiguous range of 128-pixel tiles (4 pixels per lane,
// per-pixel state in registers) and walks, per tile, the slots living on that frame; the (X,Y,Z,w) + flow chunk of
// every (tile, slot) - 3 KB - lands in the warp's private shared-memory ring STAGES chunks ahead of the one being
// computed (per-lane cp.async - a lane only ever reads what it copied itself, completion is cp.async.wait_group; or TMA
// bulk copies issued by an elected lane, completion on the warp's own mbarriers).  No cross-warp hand-off exists.
//
// Measured on C2 (B200, profiles/): 16 warps x 2 stages x 127 registers.  Round 1 / first half of round 2: 0.721 ms
// flow-active (71 % of the measured HBM peak), 0.424 ms align-only (81 %); 12 warps x 3 stages x 149 registers: 0.766 /
// 0.473 ms; 13 and 14 warps are slower than 12 (uneven warps per scheduler); TMA copies 0.843 / 0.533 ms (elect +
// uniform-register issue sequence and try_wait loop per chunk).  The kernel is bound by instruction issue first: warps per
// scheduler matter more than prefetch depth, and every instruction taken out of the chunk loop comes back as time - the
// row-tile form (ROWT below) with the leaner loop runs 0.618 ms flow-active (83 %), 0.633 ms with the folded depth step
// (85 % of its 3.54 GB), at 420 executed instructions per (tile, slot) pass instead of 488; issue slots, the shared-memory
// pipe (the 27-column transpose is half of its wavefronts) and DRAM now sit within ten points of each other
// (profiles/r02b_scan_rowtile_ncu_full.txt).
//
// History (git log, profiles/r01_scan_v{1,2,3,3c}_*): v1 = thread owns pixels, shuffle reduction (33 % of the HBM roofline);
// v2 = warp owns slot + pixel warps + TMA producer with mbarrier hand-offs (37 %: hand-off latency); v3 = this
// organisation at 12 warps with TMA rings, 630 warp instructions per chunk (4 pixel-slots per lane) of which ~210 were
// not arithmetic: re-materialised addresses, a dependent global-load chain for the slot constants, the copy-issue
// sequence and the cross-lane reduction (54 %).  v4 removes most of that:
//   * everything a warp owns lives in ONE contiguous shared-memory block (ring, slot table, running sums, transpose
//     buffer, barriers): a single base register + compile-time offsets;
//   * the constants of the frame's slots (-M, s t, flow homography, 1/area, global slot) are staged once per
//     (warp, work frame) into a shared-memory slot table: 7 LDS.128 per chunk instead of an index load, 64-bit
//     address arithmetic and 7 dependent LDG.128;
//   * the copy cursor is a running 64-bit pixel offset (one add per chunk);
//   * the 27 per-lane partial sums are reduced through the transpose buffer by half-warps: lane (h, c) adds rows
//     16h..16h+15 of column pair c with LDS.64 (16 loads instead of 32), two shuffles join the halves;
//   * the 12 per-frame sums go through the same transpose once per tile instead of living in registers;
//   * a frame's side-i slots are walked before its side-j slots (plan order): one running flow-gradient register
//     per pixel, parked in its plane when the class is complete;
//   * leaner per-pixel-slot arithmetic (ga_math.cuh: min-form smooth L1, predicated mask sums, folded dot products).
#pragma once
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include "ptx_helpers.cuh"    // mbarrier, bulk copy, cp.async, raw shared-memory accesses
#include "ga_small.cuh"       // (the diagnostic launch timeline lives there)

namespace a3r {

#ifndef A3R_V4_TILE
#define A3R_V4_TILE 128
#endif
constexpr int V4_T = A3R_V4_TILE;             // pixels per tile (128: 4 per lane; 256: 8 per lane - needs A3R_V4_WARPS <= 8)
constexpr int V4_PPL = V4_T / 32;             // pixels per lane
static_assert(V4_T == 128 || V4_T == 192 || V4_T == 256, "tile of 128, 192 or 256 pixels");
constexpr uint32_t V4_PLANE = V4_T * 4;       // bytes of one per-frame plane's share of a tile (folded depth step)
#ifndef A3R_V4_WARPS
#define A3R_V4_WARPS 16
#endif
constexpr int V4_WARPS = A3R_V4_WARPS;
constexpr int V4_THREADS = V4_WARPS * 32;
#ifndef A3R_V4_STAGES
#define A3R_V4_STAGES 2
#endif
constexpr int V4_STAGES = A3R_V4_STAGES;
#ifndef A3R_V4_EVICT_FIRST
#define A3R_V4_EVICT_FIRST 0
#endif
#ifndef A3R_V4_SLOTS
#define A3R_V4_SLOTS 20
#endif
constexpr int V4_SLOTS = A3R_V4_SLOTS;        // max slots per work frame
constexpr int V4_ROW = 28;                    // floats per row of the slot table / running sums / transpose buffer

// byte offsets inside a warp's private block
constexpr uint32_t V4_XW = 0;                                             // [STAGES][T] float4
constexpr uint32_t V4_FL = V4_XW + V4_STAGES * V4_T * 16;                 // [STAGES][T] float2
constexpr uint32_t V4_STAB = V4_FL + V4_STAGES * V4_T * 8;                // [SLOTS][28] slot table
constexpr uint32_t V4_WACC = V4_STAB + V4_SLOTS * V4_ROW * 4;             // [SLOTS + 1][28] running sums (+ frame row)
constexpr uint32_t V4_TR = V4_WACC + (V4_SLOTS + 1) * V4_ROW * 4;         // [32][28] + 4 transpose buffer
constexpr uint32_t V4_FP = V4_TR + (32 * V4_ROW + 4) * 4;                 // [16] frame block
constexpr uint32_t V4_BAR = V4_FP + 16 * 4;                               // [STAGES] mbarriers
constexpr uint32_t V4_WARP_BYTES = ((V4_BAR + V4_STAGES * 8 + 127) / 128) * 128;
constexpr size_t V4_SMEM = (size_t)V4_WARP_BYTES * V4_WARPS;
static_assert(V4_SMEM <= 232448, "v4 scan: shared memory of one CTA exceeds the 227 KB opt-in limit");

// slot table row: q0..q2 = -M (9), s t (3); q3..q5 = A (9), b (3); q6 = 1/area, side (not read by the scan: the class loop knows both)
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }

// exact pix -> (u, v) for pix < 2^24 without an integer division: the float quotient is off by at most one
__device__ __forceinline__ void pix_to_uv(int pix, int W, float invW, int& u, int& v) {
    v = __float2int_rz(__int2float_rn(pix) * invW);
    u = pix - v * W;
    if (u < 0) { u += W; --v; }
    else if (u >= W) { u -= W; ++v; }
}

// Reduce the per-lane rows (NQ float4 each) through the warp's transpose buffer; returns in lanes 0..13 the warp
// totals of columns (2 lane, 2 lane + 1) (all lanes of both halves hold them; lanes & 15 >= 14 hold garbage).
template <int NQ>
__device__ __forceinline__ float2 transpose_reduce(uint32_t tr_s, int lane, const float* a) {
    const uint32_t row = tr_s + lane * (V4_ROW * 4);
#pragma unroll
    for (int q = 0; q < NQ; ++q) sts128(row + q * 16, a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
    __syncwarp();
    const uint32_t col = tr_s + (lane >> 4) * (16 * V4_ROW * 4) + (lane & 15) * 8;
    float2 t0 = lds64(col), t1 = lds64(col + V4_ROW * 4);          // (the sums start with the first rows: no 0 + x)
#pragma unroll
    for (int r = 2; r < 16; r += 2) {
        const float2 x0 = lds64(col + r * (V4_ROW * 4));
        const float2 x1 = lds64(col + (r + 1) * (V4_ROW * 4));
        t0.x += x0.x; t0.y += x0.y;
        t1.x += x1.x; t1.y += x1.y;
    }
    t0.x += t1.x; t0.y += t1.y;
    t0.x += __shfl_xor_sync(0xffffffffu, t0.x, 16);
    t0.y += __shfl_xor_sync(0xffffffffu, t0.y, 16);
    return t0;
}

// ADAM: the tile prologue first applies the PREVIOUS iteration's Adam step to the tile's log-depths (reference
// base_opt.py:581 for the depth parameter; the stand-alone K3 otherwise): the six planes it needs (log-depth, m, v, and the
// three parked gradient planes = 24 B per pixel = exactly one 3 KB ring chunk per tile) ride through the warp's cp.async
// ring as a pseudo-slot ahead of the tile's first real slot, so their DRAM latency is hidden behind the previous tile's
// last slots like any other chunk's.  prev_adam / prev_flow: the Adam scalars and the flow state of that iteration; the
// flow normalisers and the gate are read from scal, where that iteration's K2 left them.
// ROWT: the caller promises (A3R_F_ROW_TILES; checked per work frame) that every frame's width is a multiple of the tile, so a
// tile lies in one image row: v is a per-tile constant, no pixel of a tile is out of the frame (ga_math.cuh: flow_px_scan).
template <bool FLOW, bool L2D, bool TMA, bool ADAM, bool ROWT = false>
__global__ void __launch_bounds__(V4_THREADS, 1) k_accumulate_v4(const A3rProblem p, const A3rAdamStep prev_adam, const uint32_t prev_flow) {
    static_assert(!(TMA && ADAM), "the folded depth step rides on the per-lane cp.async ring");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    asm volatile("" : "+r"(lane));           // opaque: a register, not an S2R (25 cycles) wherever an address needs it
    const int gw = blockIdx.x * V4_WARPS + warp;                 // this warp's range in the plan ("CTA" = warp here)
    const int t0 = p.cta_tile_begin[gw], t1 = p.cta_tile_begin[gw + 1];
    const TimelineSlot tl = timeline_enter(1);
    constexpr int NV = FLOW ? ACC_N : ACC_GA;
    constexpr int NQ = (NV + 3) / 4;
    // the LAST work frame of every warp is flushed by the CTA together (below): what each warp ends on
    __shared__ int s_last[V4_WARPS][4];          // work frame (-1: none), global frame, first local slot, slot count
    int f_cw = -1, f_n = 0, f_sbegin = 0, f_deg = 0;
    do {
    if (t0 >= t1) break;
    const float thre_eff = flow_thre_eff(p.pxl_thre);
    const float4* xw_g = reinterpret_cast<const float4*>(p.predw);
    const float2* fl_g = reinterpret_cast<const float2*>(p.flow);
    const int64_t stride = p.pix_stride;

    uint32_t wb = smem_u32(smem_raw) + warp * V4_WARP_BYTES;               // the warp's private block
    asm volatile("" : "+r"(wb));             // opaque: kept in a register instead of re-derived from %tid in every chunk
    constexpr uint32_t kBytes = V4_T * 16 + (FLOW ? V4_T * 8 : 0);

    if (TMA && lane == 0) {
#pragma unroll
        for (int s = 0; s < V4_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wb + V4_BAR + s * 8), "r"(1) : "memory");
        fence_mbar_init();
    }
    for (int k = lane; k < (V4_SLOTS + 1) * V4_ROW; k += 32) sts32(wb + V4_WACC + k * 4, 0.f);
    __syncwarp();

#if A3R_V4_EVICT_FIRST
    const uint64_t stream_policy = l2_policy_evict_first();      // predw / flow are read once per scan
#endif
    // ---- copy cursor: l_left real slots of tile l_tile are still to be copied, the next one from pixel offset l_off of the
    //      slot-major planes; with ADAM every tile starts with the pseudo-slot (the tile's six per-frame planes at pixel offset
    //      l_pix).  Both offsets carry the lane (l_off + lane, l_pix + 4 lane).  The common case - the next slot of the same
    //      tile - is one test, the copies and two additions. ----
    int l_tile = t0 - 1, l_wf = 0, l_wf_end = -1, l_left = 0, l_deg = 0;
    int64_t l_off = 0, l_pix = 0;
    auto copy_chunk = [&](uint32_t stage_off) {
        if (TMA) {
            if (lane == 0) {
                const uint32_t bar = wb + V4_BAR + (stage_off / (V4_T * 16)) * 8;
                mbar_expect_tx_a(bar, kBytes);
                tma_load_1d_a(wb + V4_XW + stage_off, xw_g + l_off, V4_T * 16, bar);        // (lane 0: l_off is the chunk's)
                if (FLOW) tma_load_1d_a(wb + V4_FL + (stage_off >> 1), fl_g + l_off, V4_T * 8, bar);
            }
        } else {
            const uint32_t xd = wb + V4_XW + stage_off + lane * 16;
            const uint32_t fd = wb + V4_FL + (stage_off >> 1) + lane * 8;
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j) {
#if A3R_V4_EVICT_FIRST
                cp_async16_hint(xd + j * 512, xw_g + l_off + j * 32, stream_policy);
                if (FLOW) cp_async8_hint(fd + j * 256, fl_g + l_off + j * 32, stream_policy);
#else
                cp_async16(xd + j * 512, xw_g + l_off + j * 32);
                if (FLOW) cp_async8(fd + j * 256, fl_g + l_off + j * 32);
#endif
            }
        }
        l_off += stride;
        --l_left;
    };
    auto issue = [&](uint32_t stage_off) {      // stage_off = stage * (V4_T * 16)
        if (l_left > 0) copy_chunk(stage_off);
        else if (l_tile + 1 < t1) {             // the next tile
            ++l_tile;
            if (l_tile >= l_wf_end) {
                while (l_tile >= p.wf_tile_begin[l_wf + 1]) ++l_wf;
                l_wf_end = p.wf_tile_begin[l_wf + 1];
                const int sb = p.wf_slot_begin[l_wf];
                l_deg = p.wf_slot_begin[l_wf + 1] - sb;
                const int64_t in_frame = (int64_t)(l_tile - p.wf_tile_begin[l_wf]) * V4_T;
                l_off = (int64_t)sb * stride + in_frame + lane;
                if (ADAM) l_pix = (int64_t)p.wf_lframe[l_wf] * stride + in_frame + 4 * lane;
            } else {
                l_off += V4_T - (int64_t)l_deg * stride;
                if (ADAM) l_pix += V4_T;
            }
            l_left = l_deg;
            if (ADAM) {                         // lane L copies pixels 4L .. 4L+3 (+128 ...) of every plane (16 B each)
#pragma unroll
                for (int h = 0; h < (V4_T + 127) / 128; ++h) {
                    if (4 * lane + h * 128 >= V4_T) break;
                    const uint32_t xd = wb + V4_XW + stage_off + lane * 16 + h * 512;
                    const int64_t po = l_pix + h * 128;
                    cp_async16(xd, p.logdepth + po);
                    cp_async16(xd + V4_PLANE, p.logdepth_m + po);
                    cp_async16(xd + 2 * V4_PLANE, p.logdepth_v + po);
                    cp_async16(xd + 3 * V4_PLANE, p.g_align + po);
                    if (prev_flow) {
                        const uint32_t fd = wb + V4_FL + (stage_off >> 1) + lane * 16 + h * 512;
                        cp_async16(fd, p.g_flow_i + po);
                        cp_async16(fd + V4_PLANE, p.g_flow_j + po);
                    }
                }
            } else copy_chunk(stage_off);
        }
        if (!TMA) cp_async_commit();   // one group per chunk slot, empty past the end: uniform wait_group accounting
    };
#pragma unroll
    for (int k = 0; k < V4_STAGES; ++k) issue(k * (V4_T * 16));
    // Everything above reads only what no kernel of this iteration writes (the plan, the packed inputs and - ADAM - the
    // per-pixel planes the previous SCAN left): with a programmatic launch it overlaps the tail of K2.  From here on the
    // kernel reads K2's results (parameter maps, flow normalisers) and adds into the accumulator block K2 zeroed.
    pdl_wait();
    timeline_mark(tl, 2);
    // (No early griddepcontrol.launch_dependents here: with it K2's CTAs are resident before the scan's last CTA exits and its
    // launch latency disappears, but they land on the first SMs the scan frees - neighbours - and K2 itself runs 6-8 us longer:
    // the iteration gets slower, 135 vs 128 us on a 7-frame shard, 808 vs 792 us on C2, tools/timeline.py.)
    float sc_i = 0.f, sc_j = 0.f;
    if (ADAM && prev_flow) { sc_i = (float)p.scal[A3R_S_SC_I]; sc_j = (float)p.scal[A3R_S_SC_J]; }

    // ---- compute cursor ----
    int cw = 0, cw_end = -1, lf = 0, n = 0, P = 0, W = 1, deg = 0, n_i = 0, s_begin = 0, accum = 0, tile0 = 0;
    float invW = 1.f;
    uint32_t stage_off = 0, phase = 0;          // ring position of the chunk being consumed; TMA: parity bit per stage

    auto flush = [&]() {                        // this warp's share of a work frame is complete
        __syncwarp();
#ifdef A3R_V4_NOFLUSH                          // diagnostic build: what the flush atomics cost (results are wrong)
        return;
#endif
        float gl_i = 0.f, gl_j = 0.f;           // lanes ACC_LOSS / ACC_FNUM / ACC_FCNT: the whole-problem sums (acc_glob), per side
        for (int s = 0; s < deg; ++s) {
            const uint32_t a = wb + V4_WACC + (s * V4_ROW + lane) * 4;
            if (lane < NV) {
                const float v = lds32(a);
                sts32(a, 0.f);
                const int gs = p.lslot_gslot[s_begin + s];
                if (v != 0.f) atomicAdd(&p.acc_slot[(int64_t)gs * ACC_STRIDE + lane], (double)v);
                if (gs & 1) gl_j += v; else gl_i += v;
            }
        }
        if (lane == ACC_LOSS) atomicAdd(&p.acc_glob[GL_LOSS], (double)gl_i + (double)gl_j);
        if (FLOW && (lane == ACC_FNUM || lane == ACC_FCNT)) {
            atomicAdd(&p.acc_glob[GL_FNUM_I + (lane - ACC_FNUM)], (double)gl_i);
            atomicAdd(&p.acc_glob[GL_FNUM_J + (lane - ACC_FNUM)], (double)gl_j);
        }
        if (lane < FACC_N) {
            const uint32_t a = wb + V4_WACC + (V4_SLOTS * V4_ROW + lane) * 4;
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
            if (ROWT && (W % V4_T) != 0) __trap();               // the caller's promise behind A3R_F_ROW_TILES
            invW = 1.0f / (float)W;
            s_begin = p.wf_slot_begin[cw];
            deg = p.wf_slot_begin[cw + 1] - s_begin;
            accum = p.wf_accum[cw];
            // stage the frame block and the slot table of this work frame
            if (lane < 4) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.frame_par + FP_STRIDE * n) + lane);
                sts128(wb + V4_FP + lane * 16, v4.x, v4.y, v4.z, v4.w);
            }
            for (int k = lane; k < deg * 7; k += 32) {
                const int s = k / 7, q = k - s * 7;
                const int gs = p.lslot_gslot[s_begin + s];
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < 3) {
                    v4 = __ldg(reinterpret_cast<const float4*>(p.edge_par + EP_STRIDE * (gs >> 1)) + q);
                    v4.x = -v4.x;                                     // -M (9 entries), s t (3) stay positive
                    if (q < 2) { v4.y = -v4.y; v4.z = -v4.z; v4.w = -v4.w; }
                } else if (q < 6) {
                    if (FLOW) v4 = __ldg(reinterpret_cast<const float4*>(p.slot_par + SP_STRIDE * gs) + (q - 3));
                    if (FLOW && ROWT && q == 5) v4.x += 1e-6f;    // A[8] + the 1e-6 of the projection's denominator (flow_px_scan)
                } else {
                    v4.x = (gs & 1) ? p.inv_area_j : p.inv_area_i;
                    v4.y = (gs & 1) ? 1.0f : 0.0f;
                }
                sts128(wb + V4_STAB + (s * V4_ROW + q * 4) * 4, v4.x, v4.y, v4.z, v4.w);
            }
            {
                const bool is_i = lane < deg && (p.lslot_gslot[s_begin + (lane < deg ? lane : 0)] & 1) == 0;
                const unsigned mi = __ballot_sync(0xffffffffu, is_i);
                n_i = __popc(mi);
                if (mi != (n_i >= 32 ? 0xffffffffu : (1u << n_i) - 1u)) __trap();   // plan contract: side-i slots first
            }
            __syncwarp();
        }
        const int p0 = (tile - tile0) * V4_T;
        const int64_t pix_base = (int64_t)lf * stride + p0;

        // ---- the previous iteration's Adam step on this tile's log-depths (the pseudo-slot of the ring) ----
        float ldv[V4_PPL];
        // the tile's static-mask bytes: loaded ahead of the depth step, whose arithmetic covers their latency
        uint8_t smask[V4_PPL];
        if (FLOW) {
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j)
                smask[j] = (ROWT || p0 + j * 32 + lane < P) ? p.static_mask[pix_base + j * 32 + lane] : (uint8_t)0;
        }
        if (ADAM) {
            cp_async_wait<V4_STAGES - 1>();       // this lane's copies of the pseudo-slot have landed (it reads only its own)
#pragma unroll
            for (int h = 0; h < (V4_T + 127) / 128; ++h) {
                if (4 * lane + h * 128 >= V4_T) break;
                const uint32_t xa = wb + V4_XW + stage_off + lane * 16 + h * 512;
                float4 l = lds128(xa), m = lds128(xa + V4_PLANE), v = lds128(xa + 2 * V4_PLANE), g = lds128(xa + 3 * V4_PLANE);
                if (prev_flow) {
                    const uint32_t fa = wb + V4_FL + (stage_off >> 1) + lane * 16 + h * 512;
                    const float4 a = lds128(fa), b = lds128(fa + V4_PLANE);
                    g.x += sc_i * a.x + sc_j * b.x; g.y += sc_i * a.y + sc_j * b.y;
                    g.z += sc_i * a.z + sc_j * b.z; g.w += sc_i * a.w + sc_j * b.w;
                }
                adam_update(l.x, m.x, v.x, g.x, p.adam_beta1, p.adam_beta2, p.adam_eps, prev_adam.step_size, prev_adam.bc2_sqrt);
                adam_update(l.y, m.y, v.y, g.y, p.adam_beta1, p.adam_beta2, p.adam_eps, prev_adam.step_size, prev_adam.bc2_sqrt);
                adam_update(l.z, m.z, v.z, g.z, p.adam_beta1, p.adam_beta2, p.adam_eps, prev_adam.step_size, prev_adam.bc2_sqrt);
                adam_update(l.w, m.w, v.w, g.w, p.adam_beta1, p.adam_beta2, p.adam_eps, prev_adam.step_size, prev_adam.bc2_sqrt);
                const int64_t po = pix_base + 4 * lane + h * 128;
                *reinterpret_cast<float4*>(p.logdepth + po) = l;
                *reinterpret_cast<float4*>(p.logdepth_m + po) = m;
                *reinterpret_cast<float4*>(p.logdepth_v + po) = v;
                sts128(xa, l.x, l.y, l.z, l.w);   // the main loop owns pixels lane, lane + 32, ...: hand the new values over
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j) ldv[j] = lds32(wb + V4_XW + stage_off + (j * 32 + lane) * 4);
            __syncwarp();                         // every lane has read before the stage is refilled
            issue(stage_off);
            stage_off = stage_off == (V4_STAGES - 1) * (V4_T * 16) ? 0u : stage_off + V4_T * 16;
        }

        // ---- per-pixel state (4 pixels per lane) ----
        // ROWT: v is the tile's constant, the static mask is the pixel's own flow threshold (thre_eff where static, -1
        // elsewhere); general form: the mask is the sign of disp (the arithmetic of flow_px, as the v1 / v5 scans use it)
        float zx[V4_PPL], zy[V4_PPL], zz[V4_PPL], pu[V4_PPL], pv[ROWT ? 1 : V4_PPL], disp[V4_PPL], dep[V4_PPL];
        float thr[ROWT ? V4_PPL : 1];
        float gsx[V4_PPL], gsy[V4_PPL], gsz[V4_PPL];
        float gf[V4_PPL];                        // d loss / d disp of the flow class (side i, then side j) being walked
        {
            float fpr[FP_STRIDE];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v4 = lds128(wb + V4_FP + k * 16);
                fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
            }
            int u0 = 0, v0 = 0;
            if (ROWT) pix_to_uv(p0 + lane, W, invW, u0, v0);
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j) {
                const int pl = j * 32 + lane, pix = p0 + pl;
                const bool valid = ROWT || pix < P;
                const float ld = valid ? (ADAM ? ldv[j] : p.logdepth[pix_base + pl]) : 0.f;
                bool stat = false;
                if (FLOW) stat = smask[j] != 0;
                int u, v;
                if (ROWT) { u = u0 + 32 * j; v = v0; }
                else pix_to_uv(pix, W, invW, u, v);
                PixState s;
                pixel_state(fpr, ld, u, v, s);
                zx[j] = s.zx; zy[j] = s.zy; zz[j] = s.zz; pu[j] = s.u; pv[ROWT ? 0 : j] = s.v; dep[j] = s.cz;
                if (ROWT) { disp[j] = s.disp; thr[j] = stat ? thre_eff : -1.0f; }
                else disp[j] = stat ? s.disp : -s.disp;
                gsx[j] = gsy[j] = gsz[j] = gf[j] = 0.f;
            }
        }

        // ---- slots: the side-i slots of the frame, then its side-j slots (the plan orders them so; checked when the
        //      slot table is staged).  The two flow classes are normalised by different global counts that are only
        //      known after the scan, so each class parks its own per-pixel plane as soon as it is complete. ----
#pragma unroll 1
        for (int cls = 0; cls < 2; ++cls) {
        const int s_lo = cls ? n_i : 0, s_hi = cls ? deg : n_i;
        const float wn = cls ? p.inv_area_j : p.inv_area_i;     // 1 / total area of the class (warp-uniform: not a table load)
#pragma unroll 1
        for (int s = s_lo; s < s_hi; ++s) {
            const uint32_t tab = wb + V4_STAB + s * (V4_ROW * 4);
            if (TMA) {
                mbar_wait_a(wb + V4_BAR + (stage_off / (V4_T * 16)) * 8, (phase >> (stage_off / (V4_T * 16))) & 1);
                phase ^= 1u << (stage_off / (V4_T * 16));
            } else {
                cp_async_wait<V4_STAGES - 1>();       // this lane's copies of the oldest chunk have landed
            }
            float acc[4 * NQ];
#pragma unroll
            for (int k = 0; k < 4 * NQ; ++k) acc[k] = 0.f;
            const uint32_t xa = wb + V4_XW + stage_off + lane * 16;
            const uint32_t fa = wb + V4_FL + (stage_off >> 1) + lane * 8;
            {
                SlotConst c;
                const float4 q0 = lds128(tab), q1 = lds128(tab + 16), q2 = lds128(tab + 32);
                c.nM[0] = q0.x; c.nM[1] = q0.y; c.nM[2] = q0.z; c.nM[3] = q0.w;
                c.nM[4] = q1.x; c.nM[5] = q1.y; c.nM[6] = q1.z; c.nM[7] = q1.w;
                c.nM[8] = q2.x; c.st[0] = q2.y; c.st[1] = q2.z; c.st[2] = q2.w;
                c.wn = wn;
#pragma unroll
                for (int j = 0; j < V4_PPL; ++j) {
                    const float4 X = lds128(xa + j * 512);
                    PixState st;
                    st.zx = zx[j]; st.zy = zy[j]; st.zz = zz[j];
                    float gx, gy, gz;
                    if (j == 0) align_px<L2D, true>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                    else align_px<L2D>(c, X.x, X.y, X.z, X.w, st, acc, gx, gy, gz);
                    gsx[j] += gx; gsy[j] += gy; gsz[j] += gz;
                }
            }
            if (FLOW) {
                SlotConst c;
                const float4 q3 = lds128(tab + 48), q4 = lds128(tab + 64), q5 = lds128(tab + 80);
                c.A[0] = q3.x; c.A[1] = q3.y; c.A[2] = q3.z; c.A[3] = q3.w;
                c.A[4] = q4.x; c.A[5] = q4.y; c.A[6] = q4.z; c.A[7] = q4.w;
                c.A[8] = q5.x; c.b[0] = q5.y; c.b[1] = q5.z; c.b[2] = q5.w;
                if (ROWT) flow_row_begin(c, pv[0]);
#pragma unroll
                for (int j = 0; j < V4_PPL; ++j) {
                    const float2 F = lds64(fa + j * 256);
                    if (ROWT && j == 0) gf[j] = flow_px_scan<0, true, true>(c, pu[j], pv[0], disp[j], F.x, F.y, thr[j], acc, gf[j]);
                    else if (ROWT) gf[j] = flow_px_scan<0, true>(c, pu[j], pv[0], disp[j], F.x, F.y, thr[j], acc, gf[j]);
                    else {
                        PixState st;
                        st.u = pu[j]; st.v = pv[ROWT ? 0 : j]; st.disp = fabsf(disp[j]);
                        gf[j] = flow_px(c, st, F.x, F.y, thre_eff, disp[j] > 0.f, acc, gf[j]);
                    }
                }
                if (ROWT) flow_row_finish<0>(pv[0], acc);
            }
            // the chunk is consumed: refill its stage (LDGSTS: a lane only ever reads what it copied itself)
            if (TMA) __syncwarp();
            issue(stage_off);
            stage_off = stage_off == (V4_STAGES - 1) * (V4_T * 16) ? 0u : stage_off + V4_T * 16;
            // 27 (13) per-lane partial sums -> warp totals -> the warp's running sums of this slot
            {
                const float2 t = transpose_reduce<NQ>(wb + V4_TR, lane, acc);
                if (lane < (NV + 1) / 2) {
                    const uint32_t a = wb + V4_WACC + s * (V4_ROW * 4) + lane * 8;
                    const float2 o = lds64(a);
                    sts64(a, o.x + t.x, o.y + t.y);
                }
                __syncwarp();
            }
        }
        if (FLOW) {                                   // this class's d loss / d logdepth plane (un-normalised)
            float* plane = cls ? p.g_flow_j : p.g_flow_i;
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j) {
                const int pl = j * 32 + lane;
                if (ROWT || p0 + pl < P) {
                    const float v = gf[j] * (-dep[j] * disp[j] * disp[j]);      // (general form: the sign squares away)
                    if (accum) atomicAdd(&plane[pix_base + pl], v); else plane[pix_base + pl] = v;
                }
                gf[j] = 0.f;
            }
        }
        }

        // ---- alignment gradient plane + per-frame sums ----
        {
            float fpr[FP_STRIDE];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v4 = lds128(wb + V4_FP + k * 16);
                fpr[4 * k] = v4.x; fpr[4 * k + 1] = v4.y; fpr[4 * k + 2] = v4.z; fpr[4 * k + 3] = v4.w;
            }
            float facc[FACC_N];
#pragma unroll
            for (int k = 0; k < FACC_N; ++k) facc[k] = 0.f;
#pragma unroll
            for (int j = 0; j < V4_PPL; ++j) {
                const int pl = j * 32 + lane, pix = p0 + pl;
                if (!ROWT && pix >= P) continue;
                PixState s;
                const float d = dep[j];
                s.cx = d * (pu[j] - fpr[FP_CX]) * fpr[FP_INVF];
                s.cy = d * (pv[ROWT ? 0 : j] - fpr[FP_CY]) * fpr[FP_INVF];
                s.cz = d;
                const float ga = pixel_epilogue(fpr, s, gsx[j], gsy[j], gsz[j], facc);
                const int64_t o = pix_base + pl;
                if (accum) atomicAdd(&p.g_align[o], ga); else p.g_align[o] = ga;
            }
            const float2 t = transpose_reduce<FACC_N / 4>(wb + V4_TR, lane, facc);
            if (lane < FACC_N / 2) {
                const uint32_t a = wb + V4_WACC + V4_SLOTS * (V4_ROW * 4) + lane * 8;
                const float2 o = lds64(a);
                sts64(a, o.x + t.x, o.y + t.y);
            }
            __syncwarp();
        }
    }
    f_cw = cw; f_n = n; f_sbegin = s_begin; f_deg = deg;
    } while (0);

    // ---- final flush, per CTA: the 16 warps of a CTA walk adjacent tile ranges, so most of them end on the same work
    //      frame - such a group adds its members' running sums in shared memory (fixed warp order, fp64) and issues ONE
    //      atomic per value instead of one per warp (2 M fp64 atomics per scan on C2 otherwise, all landing at the kernel's
    //      end on a few thousand addresses: 17 us of 720; 7 us of 112 on a 1/8 shard) ----
#ifndef A3R_V4_NOFLUSH
    if (lane == 0) { s_last[warp][0] = f_cw; s_last[warp][1] = f_n; s_last[warp][2] = f_sbegin; s_last[warp][3] = f_deg; }
    __syncthreads();
    if (f_cw >= 0) {
        // the group = the warps of this CTA that end on work frame f_cw; its members share the frame's slots out among
        // themselves (slot s goes to member s mod size), so the flush is as parallel as the per-warp one it replaces
        unsigned members = 0;
        for (int w = 0; w < V4_WARPS; ++w) members |= (s_last[w][0] == f_cw) ? (1u << w) : 0u;
        const int size = __popc(members), rank = __popc(members & ((1u << warp) - 1u));
        const uint32_t base = smem_u32(smem_raw) + V4_WACC + lane * 4;
        double gl_i = 0.0, gl_j = 0.0;              // lanes ACC_LOSS / ACC_FNUM / ACC_FCNT: whole-problem sums (acc_glob), per side
        for (int s = rank; s < f_deg; s += size) {
            if (lane < NV) {
                double v = 0.0;
                for (int w = 0; w < V4_WARPS; ++w)      // fixed warp order: reproducible sums
                    if (members >> w & 1u) v += (double)lds32(base + w * V4_WARP_BYTES + s * (V4_ROW * 4));
                const int gs = p.lslot_gslot[f_sbegin + s];
                if (v != 0.0) atomicAdd(&p.acc_slot[(int64_t)gs * ACC_STRIDE + lane], v);
                if (gs & 1) gl_j += v; else gl_i += v;
            }
        }
        if (lane == ACC_LOSS && gl_i + gl_j != 0.0) atomicAdd(&p.acc_glob[GL_LOSS], gl_i + gl_j);
        if (FLOW && (lane == ACC_FNUM || lane == ACC_FCNT)) {
            if (gl_i != 0.0) atomicAdd(&p.acc_glob[GL_FNUM_I + (lane - ACC_FNUM)], gl_i);
            if (gl_j != 0.0) atomicAdd(&p.acc_glob[GL_FNUM_J + (lane - ACC_FNUM)], gl_j);
        }
        if (rank == size - 1 && lane < FACC_N) {      // the frame row: the member with the fewest slots
            double v = 0.0;
            for (int w = 0; w < V4_WARPS; ++w)
                if (members >> w & 1u) v += (double)lds32(base + w * V4_WARP_BYTES + V4_SLOTS * (V4_ROW * 4));
            atomicAdd(&p.acc_frame[(int64_t)f_n * FACC_N + lane], v);
        }
    }
#endif
    timeline_mark(tl, 3);
}

}  // namespace a3r
