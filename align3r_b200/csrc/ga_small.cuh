// ga_small.cuh - K2: everything between two scans that is not per-pixel, in ONE launch whose grid grows with the
// problem: loss assembly + gate (reference optimizer.py:536-557), the temporal term (:559-572), every parameter chain
// (hand-derived backward of base_opt.py:195-247 / optimizer.py:398-428 / goem_opt.py:149-154), torch.optim.Adam on the
// small parameters (base_opt.py:524,581), the parameter maps of the NEXT iteration (k_prepare's job), and - on frame-window
// shards - the exchange of the accumulator block (every value is read from the ONE rank that owns it, over NVLink).
//
// Organisation (round 2; round 1's version was one 8-CTA cluster, fp64 atomics into per-frame rows, 7432 SASS instructions
// of straight-line code, 0.5 ms per iteration on the 1000-frame graph):
//   * grid = ceil(items / 256) CTAs (8 ... SM count), phases separated by a grid barrier (one counter in global memory;
//     every CTA is resident: the grid never exceeds the SM count and nothing else runs beside it);
//   * no atomics and no order-dependent sums: the contribution of every flow slot / temporal pair to its two cameras is
//     written to its own row of pair_scratch, a WARP per frame then gathers the rows of the frame's incidence list in a
//     fixed order - gradients, and therefore the replicated parameters of all ranks, are reproducible bit for bit;
//   * every global sum that feeds a parameter (mean of d loss / d log-scale, the shared focal's gradient, the mean
//     log-scale of the updated pairs) is recomputed by every CTA in one fixed order instead of being exchanged;
//   * the fp64 cameras of the current parameters live in cam_cache (written by k_prepare / by the previous launch's
//     tail): no quaternion / expm1 / exp is evaluated twice;
//   * the whole-problem sums the gate needs arrive with the accumulator block (acc_glob, added up by the scan's flush):
//     no pass over 2E slots before the first chain can start;
//   * the other accumulator block is zeroed here for the scan after next: no memset between the kernels;
//   * heavy math lives in a handful of __noinline__ functions (one copy of the code): the kernel is latency-bound and
//     run-once, instruction fetch of straight-line fp64 code was a sixth of its time.
//   phase A1  items: flow slots (chain -> pair_scratch) | pairs (gradient of q, t + Adam; d loss / d log-scale -> scratch)
//             | temporal pairs (loss + chain -> pair_scratch)
//   phase A2  mean of d loss / d log-scale (every CTA) | log-scale gradient + Adam | frames: gather + pose / focal /
//             principal-point gradient + Adam (a warp per frame)
//   phase C   loss (CTA 0) | shared focal (every CTA) | next maps: cameras -> cam_cache, frame_par, edge_par, slot_par |
//             zero the other accumulator block
#pragma once
#include "ga_math.cuh"
#include "a3r_common.cuh"
#include "a3r_ga.h"
#include <stddef.h>

namespace a3r {

constexpr int K2_THREADS = 256;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int PS_STRIDE = 32;                  // pair_scratch row: source half [0,16), target half [16,32)
constexpr int CAM_STRIDE = 20;                 // FrameD as doubles
static_assert(sizeof(FrameD) == CAM_STRIDE * sizeof(double) && offsetof(FrameD, f) == 12 * sizeof(double), "cam_cache rows hold a FrameD");
constexpr size_t PEER_FLAG_STRIDE = 128;
constexpr size_t PEER_HEADER = A3R_MAX_PEERS * PEER_FLAG_STRIDE;

// where the accumulators of this iteration live
struct AccSrc {
    int32_t world, rank;
    int32_t win[A3R_MAX_PEERS + 1];            // frame windows (world > 1)
    const double* base[A3R_MAX_PEERS];         // every rank's accumulator block of this parity (world == 1: base[0])
    double* zero_block;                        // own block of the other parity: zeroed for the scan after next, or NULL
    int64_t n_acc;
    void* region[A3R_MAX_PEERS];               // world > 1: every rank's exchange region (flag lines first)
    uint32_t tag;
};

__device__ __forceinline__ void st_release_sys(uint32_t* q, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(q), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* q) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* q) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
// accumulators: written by another kernel (the scan's fp64 atomics land in L2), possibly by another GPU
__device__ __forceinline__ double ld_acc(const double* q, bool sys) {
    if (sys) {
        double v;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(q) : "memory");
        return v;
    }
    return __ldcg(q);
}
// scratch written by other CTAs of this launch before the last grid barrier (L2 loads; the barrier's acquire orders them)
__device__ __forceinline__ double ld_cg(const double* q) { return __ldcg(q); }
__device__ __forceinline__ float ld_cg(const float* q) { return __ldcg(q); }

// diagnostic phase trace (a3r_debug_k2_trace): globaltimer stamps of thread 0 of the first 8 CTAs at the phase boundaries
__device__ unsigned long long g_k2_trace[8 * 8];
__device__ int g_k2_trace_on = 0;
__device__ int g_k2_trace_iter = -1;           // >= 0: stamp only the launch of this iteration (A3rAdamStep.iter)
__device__ __forceinline__ void k2_stamp(int slot, int iter = -1) {
    if (g_k2_trace_on && threadIdx.x == 0 && blockIdx.x < 8 && (g_k2_trace_iter < 0 || g_k2_trace_iter == iter)) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_k2_trace[blockIdx.x * 8 + slot] = t;
    }
}

// diagnostic launch timeline (a3r_debug_timeline): CTA 0 of every scan / update launch records (kernel id, entry, the moment
// its programmatic-launch wait returned, exit) in a 64-entry ring
__device__ unsigned long long g_timeline[64 * 4];
__device__ unsigned int g_timeline_seq = 0;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct TimelineSlot { unsigned long long* row; };
__device__ __forceinline__ TimelineSlot timeline_enter(int kernel_id) {
    TimelineSlot t{nullptr};
    if (g_k2_trace_on && blockIdx.x == 0 && threadIdx.x == 0) {
        t.row = g_timeline + (atomicAdd(&g_timeline_seq, 1u) & 63u) * 4;
        t.row[0] = (unsigned long long)kernel_id; t.row[1] = global_ns(); t.row[2] = 0; t.row[3] = 0;
    }
    return t;
}
__device__ __forceinline__ void timeline_mark(const TimelineSlot& t, int k) { if (t.row) t.row[k] = global_ns(); }

// grid barrier: every CTA of the launch is resident (grid <= SM count, nothing runs beside it)
__device__ __forceinline__ void grid_barrier(uint32_t* ctr, uint32_t target) {
    __syncthreads();                           // the CTA's writes precede thread 0's release (bar.sync is morally strong)
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");      // arrive: no round trip
        uint32_t seen, spins = 0;
        do {        // poll at L2 without an L1 invalidation per probe (ld.acquire costs one); acquire once, on the way out
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            if (++spins > (1u << 27)) __trap();      // a CTA that never arrives (the grid is not co-resident) must fail loudly, not hang
        } while (seen < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

__device__ __forceinline__ int owner_of(const AccSrc& a, int frame) {
    int r = 0;
    while (r + 1 < a.world && frame >= a.win[r + 1]) ++r;
    return r;
}

__device__ __forceinline__ void cam_load(const double* cam_cache, int n, FrameD& F) {
    const double2* q = reinterpret_cast<const double2*>(cam_cache + (int64_t)n * CAM_STRIDE);
    double2* d = reinterpret_cast<double2*>(&F);
#pragma unroll
    for (int k = 0; k < CAM_STRIDE / 2; ++k) d[k] = q[k];
}
__device__ __forceinline__ void cam_store(double* cam_cache, int n, const FrameD& F) {
    double2* q = reinterpret_cast<double2*>(cam_cache + (int64_t)n * CAM_STRIDE);
    const double2* d = reinterpret_cast<const double2*>(&F);
#pragma unroll
    for (int k = 0; k < CAM_STRIDE / 2; ++k) q[k] = d[k];
}

// one copy of each heavy body (the callers are run-once code: size matters more than call overhead)
__device__ __noinline__ void nl_load_frame(const float* im_poses, const float* im_focals, const float* im_pp, const int32_t* frame_hw,
                                           int n, bool shared_focal, double focal_break, FrameD* F) {
    load_frame(im_poses, im_focals, im_pp, frame_hw, n, shared_focal, focal_break, *F);
}
__device__ __noinline__ void nl_derive_slot_par(const FrameD* S, const FrameD* T, float* o) { derive_slot_par(*S, *T, o); }

// torch.optim.Adam on ONE entry (every caller spreads the entries of a parameter over threads: no serial loops of
// dependent load -> sqrt -> divide -> store round trips)
__device__ __noinline__ void adam_one(float* prm, float* m, float* v, float g, float b1, float b2, float eps, float step_size, float bc2_sqrt) {
    float pv = *prm, mv = *m, vv = *v;
    adam_update(pv, mv, vv, g, b1, b2, eps, step_size, bc2_sqrt);
    *prm = pv; *m = mv; *v = vv;
}
__device__ __forceinline__ void adam_entry(float* prm, float* m, float* v, float g, const A3rProblem& p, const A3rAdamStep& a) {
    adam_one(prm, m, v, g, p.adam_beta1, p.adam_beta2, p.adam_eps, a.step_size, a.bc2_sqrt);
}

// parameter maps of the current parameters (k_prepare, and the tail of K2 for the next iteration): items
// [0, 2E) flow homographies (only with flow planes) | [2E, 3E) pair blocks | [3E, 3E+N) cameras + frame blocks.
// `focals`: the focal parameters to read (K2 hands in its updated copy of the shared focal).  Every CTA recomputes
// mean_e(log scale) in one fixed order (256 threads) -> the same number everywhere.
__device__ __forceinline__ void prepare_items(const A3rProblem& p, const float* focals, const int gid, const int gstride,
                                              double* s_red, double* s_mean) {
    const int tid = threadIdx.x;
    const int E = p.n_edges, N = p.n_frames;
    const bool shared_focal = p.flags & A3R_F_SHARED_FOCAL;
    const bool norm_scale = p.flags & A3R_F_NORM_PW_SCALE;
    const double fb = (double)p.focal_break;
    double part = 0.0;
    for (int e = tid; e < E; e += K2_THREADS) part += (double)ld_cg(p.pw_poses + 8 * e + 7);
    part = block_sum(part, s_red);
    if (tid == 0) *s_mean = part / (double)E;
    __syncthreads();
    const double log_norm = norm_scale ? (log((double)p.base_scale) - *s_mean) : 0.0;
    const int nS = (p.flags & A3R_F_HAS_FLOW) ? 2 * E : 0;
    const int total = nS + E + N;
#pragma unroll 1
    for (int it = gid; it < total; it += gstride) {
        if (it < nS) {
            const int gs = it, e = gs >> 1, side = gs & 1;
            const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
            FrameD S, T;
            nl_load_frame(p.im_poses, focals, p.im_pp, p.frame_hw, side ? j : i, shared_focal, fb, &S);
            nl_load_frame(p.im_poses, focals, p.im_pp, p.frame_hw, side ? i : j, shared_focal, fb, &T);
            nl_derive_slot_par(&S, &T, p.slot_par + SP_STRIDE * gs);
        } else if (it < nS + E) {
            const int e = it - nS;
            derive_edge_par(p.pw_poses + 8 * e, p.pw_adaptors + 2 * e, log_norm, norm_scale, (double)p.pw_break, p.edge_par + EP_STRIDE * e);
        } else {
            const int n = it - nS - E;
            FrameD F;
            nl_load_frame(p.im_poses, focals, p.im_pp, p.frame_hw, n, shared_focal, fb, &F);
            derive_frame_par(F, p.frame_par + FP_STRIDE * n);
            cam_store(p.cam_cache, n, F);
        }
    }
}

__global__ void __launch_bounds__(K2_THREADS) k_prepare(const A3rProblem p) {
    __shared__ double s_red[32];
    __shared__ double s_mean;
    prepare_items(p, p.im_focals, blockIdx.x * K2_THREADS + threadIdx.x, gridDim.x * K2_THREADS, s_red, &s_mean);
}

__global__ void __launch_bounds__(K2_THREADS, 1)
k_small(const A3rProblem p, const uint32_t mode, const A3rAdamStep adam, const uint32_t prepare_next, const AccSrc src) {
    __shared__ double s_red[32];
    __shared__ double s_gl[8];
    __shared__ double s_sc[4];                 // sc_i, sc_j, dropped, flow loss
    __shared__ double s_mean;
    __shared__ float s_foc[3];                 // the shared focal: parameter, m, v as they were when the kernel started
    __shared__ double s_fg[K2_WARPS * FG_STRIDE];
    __shared__ double s_tr[K2_WARPS * 32 * FG_STRIDE];            // 32 KB: the frame gather's transpose buffers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int vt = blockIdx.x * K2_THREADS + tid, nvt = G * K2_THREADS;
    const int N = p.n_frames, E = p.n_edges;
    const bool shared_focal = p.flags & A3R_F_SHARED_FOCAL;
    const bool norm_scale = p.flags & A3R_F_NORM_PW_SCALE;
    const bool flow_active = (mode & A3R_M_FLOW_ACTIVE) && (p.flags & A3R_F_HAS_FLOW) && p.flow_loss_weight > 0.f;
    const bool apply = mode & A3R_M_APPLY_ADAM;
    const bool prior = p.depth_regularize_weight > 0.f;
    const bool sys = src.world > 1;
    const double fb = (double)p.focal_break;
    const double tsw = (double)p.temporal_smoothing_weight, tw = (double)p.translation_weight;
    // offsets inside an accumulator block
    const int64_t off_frame = (int64_t)2 * E * ACC_STRIDE;
    const int64_t off_prior = off_frame + (int64_t)N * FACC_N;
    const int64_t off_glob = off_prior + (prior ? (int64_t)N * PR_N : 0);

    const TimelineSlot tl = timeline_enter(2);
    pdl_wait();             // launched while the scan was still running: from here on its sums are complete and visible
    timeline_mark(tl, 2);
    if (!(prepare_next & 2u)) pdl_trigger();   // the next scan's CTAs may take the idle SMs and prefetch their first tiles (they wait for this grid)
    k2_stamp(0, adam.iter);
    // ---------------- exchange: publish "my sums of this iteration are complete", wait for the peers' ----------------
    if (sys) {
        if (blockIdx.x == 0 && tid < src.world && tid != src.rank) {
            __threadfence_system();            // the scan completed before this kernel started (stream order)
            st_release_sys(reinterpret_cast<uint32_t*>(static_cast<char*>(src.region[tid]) + (size_t)src.rank * PEER_FLAG_STRIDE), src.tag);
        }
        if (tid < src.world && tid != src.rank) {
            const uint32_t* flag = reinterpret_cast<const uint32_t*>(static_cast<const char*>(src.region[src.rank]) + (size_t)tid * PEER_FLAG_STRIDE);
            uint32_t spins = 0;
            while ((int32_t)(ld_acquire_sys(flag) - src.tag) < 0) {
                if (++spins > (1u << 26)) __trap();          // a peer that never arrives must fail loudly, not hang the GPU
                __nanosleep(64);
            }
        }
        __syncthreads();
    }
    // ---------------- whole-problem sums -> gate (reference optimizer.py:538-542), the same on every CTA ----------------
    if (shared_focal && tid == 0) { s_foc[0] = p.im_focals[0]; s_foc[1] = p.im_focals_m[0]; s_foc[2] = p.im_focals_v[0]; }
    if (tid < 5) {
        double s = 0.0;
        for (int r = 0; r < src.world; ++r) s += ld_acc(src.base[r] + off_glob + tid, sys);      // rank order
        s_gl[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        const LossScalars r = assemble_loss(0.0, 0.0, s_gl[GL_FNUM_I], s_gl[GL_FCNT_I], s_gl[GL_FNUM_J], s_gl[GL_FCNT_J], flow_active,
                                            (double)p.flow_loss_weight, (double)p.flow_loss_thre, 0.0);
        s_sc[0] = r.sc_i; s_sc[1] = r.sc_j; s_sc[2] = r.dropped; s_sc[3] = r.floss;
    }
    __syncthreads();
    const double sc_i = s_sc[0], sc_j = s_sc[1];
    const bool chains = flow_active && (sc_i != 0.0 || sc_j != 0.0);
    k2_stamp(1, adam.iter);

    // ---------------- phase A1 ----------------
    {
        const int nS = chains ? 2 * E : 0;
        const int nT = tsw > 0 ? N - 1 : 0;
        const int total = nS + E + nT;
#pragma unroll 1
        for (int it = vt; it < total; it += nvt) {
            if (it < nS) {                                       // flow chain of a slot -> its two cameras
                const int gs = it, e = gs >> 1, side = gs & 1;
                const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
                const int sI = side ? j : i, tI = side ? i : j;
                FrameD S, T;
                cam_load(p.cam_cache, sI, S);
                cam_load(p.cam_cache, tI, T);
                const double* a = src.base[sys ? owner_of(src, sI) : 0] + (int64_t)gs * ACC_STRIDE;
                double row[ACC_STRIDE];
#pragma unroll
                for (int k = ACC_GA; k < ACC_GB + 3; ++k) row[k] = ld_acc(a + k, sys);
                PairGrad g;
                slot_flow_chain(S, T, row, side ? sc_j : sc_i, g);
                double* o = p.pair_scratch + (int64_t)gs * PS_STRIDE;
#pragma unroll
                for (int k = 0; k < 15; ++k) { o[k] = g.s[k]; o[16 + k] = g.t[k]; }
            } else if (it < nS + E) {                            // a pair: everything but its log-scale
                const int e = it - nS;
                const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
                const double* q0 = src.base[sys ? owner_of(src, i) : 0] + (int64_t)(2 * e) * ACC_STRIDE;
                const double* q1 = src.base[sys ? owner_of(src, j) : 0] + (int64_t)(2 * e + 1) * ACC_STRIDE;
                double a0[12], a1[12];                           // sum g (3), sum g X^T (9) of the two slots
#pragma unroll
                for (int k = 0; k < 12; ++k) { a0[k] = ld_acc(q0 + k, sys); a1[k] = ld_acc(q1 + k, sys); }
                const float* ep = p.edge_par + EP_STRIDE * e;
                p.edge_scratch[e] = edge_logscale_grad(ep, a0, a1);
                float g[8], ga[2];
                edge_pose_grad(p.pw_poses + 8 * e, ep, a0, a1, 0.0, p.flags & A3R_F_TRAIN_ADAPTORS, norm_scale, (double)p.pw_break, g, ga);
                for (int k = 0; k < 7; ++k) p.g_pw_poses[8 * e + k] = g[k];       // Adam: phase A2, one thread per entry
                p.g_pw_adaptors[2 * e] = ga[0]; p.g_pw_adaptors[2 * e + 1] = ga[1];
            } else {                                             // temporal pair (S = n+1, T = n)
                const int n = it - nS - E;
                FrameD S, T;
                cam_load(p.cam_cache, n + 1, S);
                cam_load(p.cam_cache, n, T);
                PairGrad g;
                const double l = temporal_pair(S, T, tsw, tw, &g);
                p.frame_grad[(int64_t)n * FG_STRIDE + 15] = l;
                double* o = p.pair_scratch + (int64_t)(2 * E + n) * PS_STRIDE;
#pragma unroll
                for (int k = 0; k < 15; ++k) { o[k] = g.s[k]; o[16 + k] = g.t[k]; }
            }
        }
    }
    k2_stamp(2, adam.iter);
    grid_barrier(p.sync_ctr, (uint32_t)G);
    k2_stamp(3, adam.iter);

    // ---------------- phase A2 ----------------
    {
        // d loss / d log s_e = G_e - mean_e G (base_opt.py:230-241): the mean in one fixed order on every CTA
        double gsum = 0.0;
        if (norm_scale)
            for (int e = tid; e < E; e += K2_THREADS) gsum += ld_cg(p.edge_scratch + e);
        gsum = block_sum(gsum, s_red);
        const double gmean = norm_scale ? gsum / (double)E : 0.0;

        // frames (the longest dependent chain of the phase): a warp gathers the rows of the frame's incidence list.  Lane L
        // takes rows L, L+32, ... (all loads of all lanes in flight at once), the 32 x 15 partial sums cross the warp
        // through shared memory, lane j adds column j in lane order: a fixed order, reproducible sums, no shuffles; lane 0
        // finishes the camera, lanes 0-9 step the frame's ten parameter entries.
        const int gw = blockIdx.x * K2_WARPS + warp, nw = G * K2_WARPS;
        double* my_tr = s_tr + warp * (32 * FG_STRIDE);
        double* my_fg = s_fg + warp * FG_STRIDE;
#pragma unroll 1
        for (int n = gw; n < N; n += nw) {
            const int b = p.inc_begin[n], c = p.inc_begin[n + 1];
            // everything the end of the item needs is requested now: the L2 round trips overlap the gather's
            FrameD F;
            double af[FACC_N], ex[3];
            float* e_prm = nullptr; float* e_m = nullptr; float* e_v = nullptr;       // this lane's parameter entry (lanes 0-9)
            float e_p = 0.f, e_mm = 0.f, e_vv = 0.f;
            if (apply && lane < 10) {
                if (lane < 7) { if (p.flags & A3R_F_TRAIN_IM_POSES) { e_prm = p.im_poses + 7 * n + lane; e_m = p.im_poses_m + 7 * n + lane; e_v = p.im_poses_v + 7 * n + lane; } }
                else if (lane < 9) { if (p.flags & A3R_F_TRAIN_PP) { e_prm = p.im_pp + 2 * n + lane - 7; e_m = p.im_pp_m + 2 * n + lane - 7; e_v = p.im_pp_v + 2 * n + lane - 7; } }
                else if (!shared_focal && (p.flags & A3R_F_TRAIN_FOCALS)) { e_prm = p.im_focals + n; e_m = p.im_focals_m + n; e_v = p.im_focals_v + n; }
                if (e_prm) { e_p = *e_prm; e_mm = *e_m; e_vv = *e_v; }
            }
            if (lane == 0) {
                cam_load(p.cam_cache, n, F);
                const double* qa = src.base[sys ? owner_of(src, n) : 0] + off_frame + (int64_t)n * FACC_N;
#pragma unroll
                for (int k = 0; k < FACC_N; ++k) af[k] = ld_acc(qa + k, sys);
                for (int k = 0; k < 3; ++k) ex[k] = signed_expm1_grad(p.im_poses[7 * n + 4 + k]);
            }
            double part[15];
#pragma unroll
            for (int j = 0; j < 15; ++j) part[j] = 0.0;
#pragma unroll 1
            for (int k = b + lane; k < c; k += 32) {
                const int code = p.inc_items[k], row = code >> 1;
                if (row < 2 * E ? chains : (tsw > 0)) {
                    const double* q = p.pair_scratch + (int64_t)row * PS_STRIDE + (code & 1) * 16;
#pragma unroll
                    for (int j = 0; j < 15; ++j) part[j] += ld_cg(q + j);
                }
            }
#pragma unroll
            for (int j = 0; j < 15; ++j) my_tr[lane * FG_STRIDE + j] = part[j];
            __syncwarp();
            if (lane < 15) {
                double col = 0.0;
#pragma unroll
                for (int l = 0; l < 32; ++l) col += my_tr[l * FG_STRIDE + lane];
                my_fg[lane] = col;
            }
            __syncwarp();
            if (lane == 0) {
                double fa[FG_STRIDE];
                frame_align_grad(F, af, fa);
                double fg[FG_STRIDE];
#pragma unroll
                for (int k = 0; k < 15; ++k) fg[k] = my_fg[k] + fa[k];
                fg[15] = 0.0;
                float g[7], gpp[2];
                double gphi;
                frame_pose_grad_pre(F, fg, fb, p.flags & A3R_F_TRAIN_PP, ex, g, &gphi, gpp);
                for (int k = 0; k < 7; ++k) p.g_im_poses[7 * n + k] = g[k];
                p.g_im_pp[2 * n] = gpp[0]; p.g_im_pp[2 * n + 1] = gpp[1];
                if (shared_focal) p.frame_grad[(int64_t)n * FG_STRIDE + FG_F] = gphi;      // summed in phase C
                else p.g_im_focals[n] = (float)gphi;
                float* hand = reinterpret_cast<float*>(my_fg);        // hand the 10 gradients to the lanes that step them
                for (int k = 0; k < 7; ++k) hand[k] = g[k];
                hand[7] = gpp[0]; hand[8] = gpp[1]; hand[9] = (float)gphi;
            }
            __syncwarp();
            if (e_prm) {                     // lanes 0-6: pose entries, 7-8: principal point, 9: per-frame focal
                adam_update(e_p, e_mm, e_vv, reinterpret_cast<const float*>(my_fg)[lane], p.adam_beta1, p.adam_beta2, p.adam_eps,
                            adam.step_size, adam.bc2_sqrt);
                *e_prm = e_p; *e_m = e_mm; *e_v = e_vv;
            }
            __syncwarp();
            // the camera of the frame's UPDATED parameters -> cam_cache (what the next iteration's chains and this launch's
            // phase C read; the shared focal, only known after phase C's sum, is patched in there).  A1 is done with the old row.
            if ((prepare_next & 1u) && lane == 0) {
                FrameD F2;
                load_frame(p.im_poses, p.im_focals, p.im_pp, p.frame_hw, n, shared_focal, fb, F2);
                cam_store(p.cam_cache, n, F2);
            }
            __syncwarp();
        }

        // the pair parameters, one thread per entry (four entries of a thread in flight at once): entries 0-6 of a pair
        // carry the gradient phase A1 wrote, entry 7 (log-scale) G_e - mean G
        const bool step_pw = apply && (p.flags & A3R_F_TRAIN_PW_POSES);
#pragma unroll 1
        for (int base = vt; base < 8 * E; base += 4 * nvt) {
            float g[4], pv[4], mv[4], vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * nvt;
                if (idx < 8 * E) {
                    g[u] = (idx & 7) == 7 ? (float)(ld_cg(p.edge_scratch + (idx >> 3)) - gmean) : ld_cg(p.g_pw_poses + idx);
                    if (step_pw) { pv[u] = p.pw_poses[idx]; mv[u] = p.pw_poses_m[idx]; vv[u] = p.pw_poses_v[idx]; }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * nvt;
                if (idx < 8 * E) {
                    if ((idx & 7) == 7) p.g_pw_poses[idx] = g[u];
                    if (step_pw) {
                        adam_update(pv[u], mv[u], vv[u], g[u], p.adam_beta1, p.adam_beta2, p.adam_eps, adam.step_size, adam.bc2_sqrt);
                        p.pw_poses[idx] = pv[u]; p.pw_poses_m[idx] = mv[u]; p.pw_poses_v[idx] = vv[u];
                    }
                }
            }
        }
        if (apply && (p.flags & A3R_F_TRAIN_ADAPTORS)) {
#pragma unroll 1
            for (int idx = vt; idx < 2 * E; idx += nvt)
                adam_entry(p.pw_adaptors + idx, p.pw_adaptors_m + idx, p.pw_adaptors_v + idx, ld_cg(p.g_pw_adaptors + idx), p, adam);
        }
    }
    k2_stamp(4, adam.iter);
    grid_barrier(p.sync_ctr, (uint32_t)(2 * G));
    k2_stamp(5, adam.iter);
    if (prepare_next & 2u) pdl_trigger();      // (A3R_K2_TRIGGER=late: the next scan's prologue overlaps phase C only)

    // ---------------- phase C ----------------
    // the shared focal: every CTA sums the frames' contributions in one fixed order and steps its own copy of the
    // parameter (read before anybody could write it); CTA 0 stores
    if (shared_focal) {
        double part = 0.0;
        for (int n = tid; n < N; n += K2_THREADS) part += ld_cg(p.frame_grad + (int64_t)n * FG_STRIDE + FG_F);
        part = block_sum(part, s_red);
        if (tid == 0) {
            float gf = (float)part;
            if (apply && (p.flags & A3R_F_TRAIN_FOCALS)) adam_entry(&s_foc[0], &s_foc[1], &s_foc[2], gf, p, adam);
            if (blockIdx.x == 0) {
                p.g_im_focals[0] = gf;
                p.im_focals[0] = s_foc[0]; p.im_focals_m[0] = s_foc[1]; p.im_focals_v[0] = s_foc[2];
            }
        }
        __syncthreads();
    }
    // the loss of this evaluation (optimizer.py:554-557): CTA 0
    if (blockIdx.x == 0) {
        double l_temp = 0.0, l_prior = 0.0;
        if (tsw > 0)
            for (int n = tid; n < N - 1; n += K2_THREADS) l_temp += ld_cg(p.frame_grad + (int64_t)n * FG_STRIDE + 15);
        if (prior)
            for (int n = tid; n < N; n += K2_THREADS) {
                double a[PR_N];
                const double* q = src.base[sys ? owner_of(src, n) : 0] + off_prior + (int64_t)n * PR_N;
                for (int k = 0; k < PR_N; ++k) a[k] = ld_acc(q + k, sys);
                l_prior += prior_frame(a, (double)p.frame_hw[2 * n] * p.frame_hw[2 * n + 1]).loss;
            }
        l_temp = block_sum(l_temp, s_red);
        __syncthreads();
        l_prior = block_sum(l_prior, s_red);
        if (tid == 0) {
            l_prior /= (double)N;
            LossScalars r = assemble_loss(s_gl[GL_LOSS], l_temp, s_gl[GL_FNUM_I], s_gl[GL_FCNT_I], s_gl[GL_FNUM_J], s_gl[GL_FCNT_J],
                                          flow_active, (double)p.flow_loss_weight, (double)p.flow_loss_thre, tsw);
            r.loss += (double)p.depth_regularize_weight * l_prior;
            p.scal[A3R_S_PRIOR] = l_prior;
            p.scal[A3R_S_LOSS] = r.loss; p.scal[A3R_S_ALIGN] = s_gl[GL_LOSS]; p.scal[A3R_S_TEMPORAL] = l_temp;
            p.scal[A3R_S_FLOW] = r.floss; p.scal[A3R_S_FLOW_DROPPED] = r.dropped;
            p.scal[A3R_S_SC_I] = r.sc_i; p.scal[A3R_S_SC_J] = r.sc_j;
            if (r.dropped != 0) p.scal[A3R_S_FLAG_STICKY] = 1.0;
            if (p.loss_hist && adam.iter >= 0) p.loss_hist[adam.iter] = (float)r.loss;
        }
        __syncthreads();
    }
    if (prepare_next & 1u) {        // reference base_opt.py:195-247, optimizer.py:398-428 of the updated parameters
        // (k_prepare's job, from the cameras phase A2 cached: no quaternion / expm1 is evaluated again)
        double part = 0.0;
        for (int e = tid; e < E; e += K2_THREADS) part += (double)ld_cg(p.pw_poses + 8 * e + 7);
        part = block_sum(part, s_red);
        if (tid == 0) { s_mean = part / (double)E; s_gl[7] = shared_focal ? a3r_exp((double)s_foc[0] / fb) : 0.0; }
        __syncthreads();
        const double log_norm = norm_scale ? (log((double)p.base_scale) - s_mean) : 0.0;
        const double f_shared = s_gl[7];
        const int nS = (p.flags & A3R_F_HAS_FLOW) ? 2 * E : 0;
        const int total = nS + E + N;
#pragma unroll 1
        for (int it = vt; it < total; it += nvt) {
            if (it < nS) {
                const int gs = it, e = gs >> 1, side = gs & 1;
                const int i = p.edge_ij[2 * e], j = p.edge_ij[2 * e + 1];
                FrameD S, T;
                cam_load(p.cam_cache, side ? j : i, S);
                cam_load(p.cam_cache, side ? i : j, T);
                if (shared_focal) { S.f = f_shared; T.f = f_shared; }
                nl_derive_slot_par(&S, &T, p.slot_par + SP_STRIDE * gs);
            } else if (it < nS + E) {
                const int e = it - nS;
                derive_edge_par(p.pw_poses + 8 * e, p.pw_adaptors + 2 * e, log_norm, norm_scale, (double)p.pw_break, p.edge_par + EP_STRIDE * e);
            } else {
                const int n = it - nS - E;
                FrameD F;
                cam_load(p.cam_cache, n, F);
                if (shared_focal) {
                    F.f = f_shared;
                    p.cam_cache[(int64_t)n * CAM_STRIDE + 12] = f_shared;       // FrameD::f (the other fields are final since A2)
                }
                derive_frame_par(F, p.frame_par + FP_STRIDE * n);
            }
        }
    }
    if (src.zero_block)
        for (int64_t i = vt; i < src.n_acc; i += nvt) src.zero_block[i] = 0.0;
    k2_stamp(6, adam.iter);
    // the last CTA out re-arms the barrier counters for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(p.sync_ctr + 1, 1u) == (uint32_t)G - 1) {
            p.sync_ctr[0] = 0u;
            p.sync_ctr[1] = 0u;
            __threadfence();
        }
    }
    k2_stamp(7, adam.iter);
    timeline_mark(tl, 3);
}

}  // namespace a3r
