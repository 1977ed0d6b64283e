/*
 * a3r_ga.h - C ABI of the B200-native global-alignment kernels (liba3r_ga.so).
 *
 * Drop-in boundary.  The reference has no FFI seam on this path: the seam is the Python class
 * dust3r.cloud_opt_flow.PointCloudOptimizer (reference dust3r/cloud_opt_flow/optimizer.py:30,
 * base_opt.py:44).  align3r_b200.cloud_opt_flow.PointCloudOptimizer keeps that class API and hands
 * raw device pointers of PyTorch-owned storage to the entry points below through ctypes.  Each entry
 * point names the reference code it replaces.
 *
 * Contract
 *  - plain C: no torch / C++ types in any signature; every pointer inside A3rProblem is a DEVICE
 *    pointer unless its comment says "host"; the caller keeps the memory alive during the call.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no call synchronises
 *    the device except a3r_device_info.
 *  - returns 0 on success, a negative A3R_E_* code on failure; a3r_last_error() gives the message of
 *    the calling thread's last failure.
 *  - no internal threads, no hidden allocations: all scratch lives in buffers listed in A3rProblem.
 *  - there is no CPU path.  Without a CUDA device every compute entry point fails with A3R_E_CUDA.
 *
 * Index conventions
 *  N frames, E directed edges e = (i, j).  Every edge has two "slots" (edge-sides): global slot
 *  gs = 2*e + side, side 0 = the i-side (pred1 pointmap, lives on frame i's pixel grid), side 1 = the
 *  j-side (pred2 pointmap, lives on frame j's grid).  A device owns a set of "local frames" (all of
 *  them on one GPU; a contiguous frame window per GPU otherwise) and the slots living on them, stored
 *  frame-major: the local slots of local frame k are [lframe_slot_begin[k], lframe_slot_begin[k+1]).
 *  Per-pixel planes are padded to `pix_stride` pixels per frame / per slot, pixel p = v*W + u.
 */
#ifndef A3R_GA_H
#define A3R_GA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3R_ABI_VERSION 7

/* error codes */
#define A3R_OK 0
#define A3R_E_ARG (-1)
#define A3R_E_CUDA (-2)
#define A3R_E_UNSUPPORTED (-3)

/* A3rProblem.flags */
#define A3R_F_SHARED_FOCAL   (1u << 0)  /* im_focals has one entry used by all frames (optimizer.py:57-59) */
#define A3R_F_NORM_PW_SCALE  (1u << 1)  /* base_opt.py:230-236 */
#define A3R_F_DIST_L2        (1u << 2)  /* commons.py:67-68; default is 'l1' = weighted L2 norm (:71-72) */
#define A3R_F_TRAIN_PW_POSES (1u << 3)
#define A3R_F_TRAIN_DEPTH    (1u << 4)
#define A3R_F_TRAIN_IM_POSES (1u << 5)
#define A3R_F_TRAIN_FOCALS   (1u << 6)
#define A3R_F_TRAIN_PP       (1u << 7)  /* optimize_pp (optimizer.py:63-64) */
#define A3R_F_TRAIN_ADAPTORS (1u << 8)  /* allow_pw_adaptors (base_opt.py:113-114) */
#define A3R_F_HAS_FLOW       (1u << 9)  /* flow planes + dynamic masks are present */
#define A3R_F_ROW_TILES      (1u << 10) /* the caller's promise that every local frame's width is a multiple of tile_pixels (a tile
                                           lies in one image row): selects the scan specialised for it; the kernel traps on a frame
                                           that breaks the promise */

/* a3r_ga_* `mode` argument */
#define A3R_M_FLOW_ACTIVE (1u << 0)     /* epoch >= num_total_iter*flow_loss_start_epoch (optimizer.py:520) */
#define A3R_M_APPLY_ADAM  (1u << 1)     /* update parameters; otherwise only loss + gradients are produced */

/* number of fp64 accumulators per slot / per frame produced by a3r_ga_accumulate */
#define A3R_SLOT_ACC 28
#define A3R_FRAME_ACC 12
#define A3R_PRIOR_ACC 4      /* per frame: sum d, sum w, sum w d, sum w d^2 of the depth prior (d = log-depth - its initial value) */
#define A3R_GLOB_ACC 8       /* whole-problem sums the scan adds up beside the per-slot ones: alignment loss, flow numerator and
                                per-channel mask count of side i, the same of side j (5 used) */
#define A3R_SCAL 16

/* layout of A3rProblem.scal (fp64[A3R_SCAL]) after a3r_ga_update */
#define A3R_S_LOSS 0        /* total loss of this evaluation (optimizer.py:554-557) */
#define A3R_S_ALIGN 1       /* li + lj */
#define A3R_S_TEMPORAL 2    /* un-weighted temporal term */
#define A3R_S_FLOW 3        /* un-weighted flow_loss_i + flow_loss_j (NaN when inactive) */
#define A3R_S_FLOW_DROPPED 4 /* 1 when the gate dropped the flow term in this evaluation (optimizer.py:540-542) */
#define A3R_S_SC_I 5        /* flow_loss_weight / sum(per_pixel_mask_i), 0 when inactive or dropped */
#define A3R_S_SC_J 6
#define A3R_S_FLAG_STICKY 7 /* flow_loss_flag: set once the gate has tripped at least once */
#define A3R_S_PRIOR 10       /* un-weighted depth prior loss (0 when off) */

typedef struct A3rProblem {
    /* ---- sizes ---- */
    int32_t n_frames;        /* N (global) */
    int32_t n_edges;         /* E (global, directed) */
    int32_t n_local_frames;  /* frames whose pixel planes live on this device */
    int32_t n_local_slots;   /* slots living on those frames */
    int32_t n_work_frames;   /* >= n_local_frames: a frame with more than a3r_max_slots() slots is split */
    int32_t n_tiles;         /* sum over work frames of ceil(P / tile_pixels) */
    int32_t n_ctas;          /* persistent CTAs of the accumulate kernel (multiple of the SM count) */
    int32_t tile_pixels;     /* must equal a3r_tile_pixels() */
    uint32_t flags;          /* A3R_F_* */
    int64_t pix_stride;      /* padded pixels per frame plane and per slot plane (max area) */

    /* ---- hyper-parameters (reference optimizer.py:36-38, base_opt.py:61-71) ---- */
    float focal_break, pw_break, base_scale;
    float flow_loss_weight, flow_loss_thre, pxl_thre;
    float temporal_smoothing_weight, translation_weight;
    float inv_area_i, inv_area_j;     /* 1/total_area_i, 1/total_area_j (optimizer.py:91-92, GLOBAL sums) */
    float adam_beta1, adam_beta2, adam_eps;
    float depth_regularize_weight;    /* optimizer.py:546-552; needs logdepth_init + static_mask + prior_acc when > 0 */

    /* ---- global graph, replicated on every device ---- */
    const int32_t *edge_ij;           /* [E][2] */
    const int32_t *frame_hw;          /* [N][2] (H, W) */

    /* ---- small parameters + Adam moments (fp32, replicated) ---- */
    float *pw_poses, *pw_poses_m, *pw_poses_v;          /* [E][8]  qx qy qz qw tx~ ty~ tz~ log-scale */
    float *pw_adaptors, *pw_adaptors_m, *pw_adaptors_v; /* [E][2] */
    float *im_poses, *im_poses_m, *im_poses_v;          /* [N][7] */
    float *im_focals, *im_focals_m, *im_focals_v;       /* [1] if shared else [N]: focal_break*log f */
    float *im_pp, *im_pp_m, *im_pp_v;                   /* [N][2] */
    /* gradients of the last evaluation (always written by a3r_ga_update) */
    float *g_pw_poses, *g_pw_adaptors, *g_im_poses, *g_im_focals, *g_im_pp;

    /* ---- derived per-iteration blocks written by a3r_ga_prepare (fp32) ---- */
    float *frame_par;                 /* [N][16]  R(9) t(3) f cx cy - */
    float *edge_par;                  /* [E][16]  s*R*diag(a) (9), s*t (3), s, a(3) */
    float *slot_par;                  /* [2E][16] flow homography A (9), b (3), - */

    /* ---- local shard description ---- */
    const int32_t *lframe_id;         /* [n_local_frames] global frame index */
    const int32_t *wf_lframe;         /* [n_work_frames] local frame a work frame belongs to */
    const int32_t *wf_slot_begin;     /* [n_work_frames+1] local slots of a work frame (<= a3r_max_slots()) */
    const int32_t *wf_tile_begin;     /* [n_work_frames+1] */
    const int32_t *wf_accum;          /* [n_work_frames] 0: store per-pixel gradients, 1: atomically add them
                                         (frames split over several work frames; planes are zeroed first) */
    const int32_t *lslot_gslot;       /* [n_local_slots] global slot id 2*e+side */
    const int32_t *cta_tile_begin;    /* [n_ctas+1] balanced contiguous tile ranges */

    /* ---- local constant inputs ---- */
    const float *predw;               /* [n_local_slots][pix_stride][4]  X Y Z weight(=conf_trf(conf)) */
    const float *flow;                /* [n_local_slots][pix_stride][2]  or NULL */
    const uint8_t *static_mask;       /* [n_local_frames][pix_stride]  1 = ~dynamic_mask, or NULL */

    /* ---- local per-pixel state ---- */
    float *logdepth, *logdepth_m, *logdepth_v;   /* [n_local_frames][pix_stride] */
    float *g_align, *g_flow_i, *g_flow_j;        /* [n_local_frames][pix_stride] per-pixel gradient planes */

    /* ---- reduction buffers (fp64): ONE contiguous "accumulator block"
     *      acc_slot [2E][28] | acc_frame [N][12] | prior_acc [N][4] (only with the depth prior) | acc_glob [8].
     *      Multi-GPU: all-reduce(sum) the block between a3r_ga_accumulate and a3r_ga_update (or a3r_ga_run_peers) ---- */
    double *acc_slot;                 /* [2E][A3R_SLOT_ACC] */
    double *acc_frame;                /* [N][A3R_FRAME_ACC] */
    double *frame_grad;               /* [N][16] scratch of a3r_ga_update (per-frame focal gradient, temporal pair loss) */
    double *edge_scratch;             /* [E] scratch of a3r_ga_update (d loss / d log-scale of every pair) */
    double *scal;                     /* [A3R_SCAL] */
    float *loss_hist;                 /* [>= niter] loss of every iteration, or NULL */

    /* ---- depth prior (reference optimizer.py:546-552, goem_opt.py:15-36), only read when depth_regularize_weight > 0 ---- */
    const float *logdepth_init;       /* [n_local_frames][pix_stride] log(clamp(initial depth, 1e-6)) */
    double *prior_acc;                /* [N][A3R_PRIOR_ACC], directly after acc_frame (rides on the same all-reduce) */

    /* ---- added with ABI 7 ---- */
    double *acc_glob;                 /* [A3R_GLOB_ACC] last part of the accumulator block (after prior_acc / acc_frame) */
    double *acc_alt;                  /* a second accumulator block of the same layout, or NULL.  With it a3r_ga_run
                                         alternates the two by iteration parity: the update kernel of iteration k zeroes the
                                         block iteration k+1 accumulates into (no memset between the kernels) */
    double *pair_scratch;             /* [2E + N][32] scratch of a3r_ga_update: the gradient contribution of every flow slot /
                                         temporal pair to its source frame ([0,16)) and to its target frame ([16,32)) */
    double *cam_cache;                /* [N][20] fp64 cameras (R 9, t 3, f, cx, cy, unit quaternion 4, 1/|q|) of the CURRENT
                                         parameters: written by a3r_ga_prepare and by the update kernel for the next iteration */
    const int32_t *inc_begin;         /* [N+1] incidence lists: rows of pair_scratch that touch frame n ... */
    const int32_t *inc_items;         /* ... as 2*row + part (part 0: the row's source half, 1: its target half).  Row gs < 2E:
                                         flow slot gs (source = the frame it lives on, target = the pair's other frame);
                                         row 2E + n: temporal pair (source n+1, target n) */
    uint32_t *sync_ctr;               /* [4] zero-initialised counters of the update kernel's grid barrier */
} A3rProblem;

/* Adam scalars of one step, computed on the host in double like torch/optim/adam.py:457-547 */
typedef struct A3rAdamStep {
    float step_size;        /* lr / (1 - beta1^t) */
    float bc2_sqrt;         /* sqrt(1 - beta2^t) */
    int32_t iter;           /* index into loss_hist */
    int32_t _pad;
} A3rAdamStep;

/* ---- library ---- */
int a3r_abi_version(void);
const char *a3r_last_error(void);
int a3r_tile_pixels(void);              /* pixels per tile the accumulate kernel was built for */
int a3r_max_slots(void);                /* most slots one work frame may hold */
int a3r_max_ctas(int device);           /* recommended n_ctas = SM count * resident CTAs per SM */
int a3r_scan_align_warps(void);         /* 0, or (class-specialised scan, A3R_SCAN=v5) the number NA of align warps among a CTA's
                                           16: cta_tile_begin then holds three partitions of the tile list behind each other -
                                           n_ctas + 1 ranges for all-align launches, NA/16 n_ctas + 1 for the align warps and
                                           (16-NA)/16 n_ctas + 1 for the flow warps of flow-on launches */
int a3r_device_info(int device, int *sm_count, int *cc_major, int *cc_minor);

/* ---- one-off packing (replaces the ParameterStack buffers of reference optimizer.py:80-92,
 *      _compute_img_conf base_opt.py:187-193 and compute_edge_scores commons.py:21-31) ----
 * pts [n][P][3], conf [n][P] raw network outputs of n edge-sides (device, fp32); dst_slot[n] (device int32)
 * local slot index of each; conf_mode 0=log 1=sqrt 2=m1 3=id.  Writes predw, max-merges raw conf into
 * im_conf[lframe][pix_stride] (local frame index of each row in dst_lframe[n], device int32) and adds sum(conf) into
 * conf_sum[dst_slot] (fp64, device). */
int a3r_pack_pred(const float *pts, const float *conf, int32_t n, int64_t P, const int32_t *dst_slot,
                  const int32_t *dst_lframe, int64_t pix_stride, int32_t conf_mode, float *predw,
                  float *im_conf, double *conf_sum, void *stream);
/* flow [n][2][P] (reference get_flow layout, optimizer.py:118-154) -> flow plane [slot][pix_stride][2]; dst_slot device int32 */
int a3r_pack_flow(const float *flow_src, int32_t n, int64_t P, const int32_t *dst_slot, int64_t pix_stride,
                  float *flow_dst, void *stream);

/* ---- hot loop ---- */
/* parameter maps: reference base_opt.py:195-247, optimizer.py:398-428 -> frame_par/edge_par/slot_par */
int a3r_ga_prepare(const A3rProblem *p, uint32_t mode, void *stream);
/* fused forward + backward over all local pixel-slots: reference optimizer.py:500-544 and its autograd
 * backward.  Zeroes then fills acc_slot/acc_frame, writes g_align/g_flow_i/g_flow_j. */
int a3r_ga_accumulate(const A3rProblem *p, uint32_t mode, void *stream);
/* The scan as a3r_ga_run launches it from the second iteration on: the tile prologue first applies the PREVIOUS iteration's
 * Adam step to the log-depth planes (prev_adam, prev_mode: that iteration's step scalars and mode; the flow normalisers are
 * read from scal, where that iteration's a3r_ga_update left them), then scans.  Exposed for measurements and tests (the loop
 * entry points drive it themselves); A3R_E_UNSUPPORTED when the depth step does not fold on this problem. */
int a3r_ga_accumulate_folded(const A3rProblem *p, uint32_t mode, const A3rAdamStep *prev_adam, uint32_t prev_mode, void *stream);
/* loss assembly + gate (optimizer.py:538-557), temporal term (:559-572), parameter-chain gradients,
 * torch.optim.Adam step (base_opt.py:524,581) on the small parameters and on the depth planes. */
int a3r_ga_update(const A3rProblem *p, uint32_t mode, const A3rAdamStep *adam, void *stream);
/* prepare + accumulate + update, `count` consecutive iterations (single-device loop of
 * global_alignment_loop, base_opt.py:513-547).  steps/modes are host arrays of length count. */
int a3r_ga_run(const A3rProblem *p, int32_t count, const uint32_t *modes, const A3rAdamStep *steps, void *stream);

/* ---- handle layer: the export set SURVEY.md section 8(b) names, over the stateless entry points above ----
 * A handle is a validated copy of the descriptor + the Adam step count (torch's state['step']); it owns no device memory
 * (every pointer in the descriptor stays the caller's) and belongs to the device that was current at creation.  One device
 * holding every frame only: a frame-window shard uses the three-phase calls with an all-reduce, or a3r_ga_run_peers. */
typedef struct A3rHandle A3rHandle;
int a3r_ga_create(const A3rProblem *desc, A3rHandle **out);
int a3r_ga_destroy(A3rHandle *h);
const A3rProblem *a3r_ga_problem(const A3rHandle *h);
/* a3r_pack_pred (+ a3r_pack_flow when `flow` [n][2][P] is not NULL) of n edge-side rows into the descriptor's predw / flow
 * planes; arguments as for a3r_pack_pred. */
int a3r_ga_pack_inputs(A3rHandle *h, const float *pts, const float *conf, const float *flow, int32_t n, int64_t P,
                       const int32_t *dst_slot, const int32_t *dst_lframe, int32_t conf_mode, float *im_conf, double *conf_sum,
                       void *stream);
/* a fresh torch.optim.Adam (reference base_opt.py:524): zero every moment buffer of the descriptor, step count 0 */
int a3r_ga_reset_adam(A3rHandle *h, void *stream);
/* one iteration of global_alignment_iter (base_opt.py:550-583) at learning rate lr: parameter maps, scan, loss + gate +
 * chains + Adam.  flags: A3R_M_FLOW_ACTIVE or 0.  The loss lands in scal[A3R_S_LOSS] and loss_hist[iter] (when not NULL).
 * The bias corrections use the handle's step count and the decimal betas the fp32 fields stand for, like torch. */
int a3r_ga_step(A3rHandle *h, int32_t iter, int32_t niter, float lr, uint32_t flags, void *stream);
/* loss + every gradient of the current parameters, no update (reference forward + loss.backward()); the fp64 loss is
 * also copied to loss_out (device or pinned host, may be NULL) on the stream. */
int a3r_ga_forward(A3rHandle *h, uint32_t flags, double *loss_out, void *stream);
/* where the gradients of the last a3r_ga_forward live (device pointers into the descriptor's buffers; test hook) */
typedef struct A3rGrads {
    const float *pw_poses, *pw_adaptors, *im_poses, *im_focals, *im_pp;   /* shapes of the parameters */
    const float *logdepth;                                                 /* [n_local_frames][pix_stride] */
    const double *scal;                                                    /* [A3R_SCAL], A3R_S_* */
} A3rGrads;
int a3r_ga_grads(const A3rHandle *h, A3rGrads *out);
/* compute_edge_scores (reference commons.py:21-31): scores[e] = fp32(mean conf_i) * fp32(mean conf_j) from the per-slot
 * confidence sums a3r_pack_pred accumulated, indexed by GLOBAL slot 2e+side (conf_sum [2E] fp64, device); edge_ij [E][2],
 * frame_hw [N][2] device int32; scores [E] fp64 (device). */
int a3r_mst_edge_scores(const double *conf_sum, const int32_t *edge_ij, const int32_t *frame_hw, int32_t n_edges,
                        double *scores, void *stream);

/* ---- multi-GPU exchange over peer memory (NVLink / NVSwitch; one process per GPU on one node) ----
 * The only thing that crosses GPUs is the fp64 accumulator block (acc_slot | acc_frame | prior_acc).  Instead of a
 * library all-reduce between two kernels, every rank's scan writes its partial sums into a CUDA-IPC shared region of its
 * own memory, and ONE kernel per iteration (a) flags "my sums of iteration t are complete" into every peer's region,
 * (b) waits for the peers' flags, (c) reads all ranks' partial sums over NVLink and adds them in rank order (bit-identical
 * on every rank) into the private block the update kernels read.  Two partial-sum buffers alternate by iteration parity,
 * so no second barrier is needed (a rank cannot overwrite a buffer a peer may still be reading: it only reaches that
 * buffer again after the peer flagged the iteration in between).
 * Region layout (bytes): [A3R_MAX_PEERS flags, 128 B apart][partial sums, parity 0][partial sums, parity 1]. */
#define A3R_MAX_PEERS 8
typedef struct A3rPeers {
    int32_t world, rank;
    int64_t n_acc;                          /* doubles in the accumulator block */
    void *base[A3R_MAX_PEERS];              /* every rank's region mapped into this process (base[rank] = own) */
    int32_t win[A3R_MAX_PEERS + 1];         /* frame windows: rank r owns frames win[r] <= n < win[r+1] */
} A3rPeers;
size_t a3r_peer_region_bytes(int64_t n_acc);
/* own region: cudaMalloc + zero + cudaIpcGetMemHandle (64-byte handle to hand to the other ranks) */
int a3r_peer_alloc(int64_t n_acc, void **region, unsigned char handle[64]);
int a3r_peer_open(const unsigned char handle[64], void **region);       /* a peer's region (cudaIpcOpenMemHandle) */
int a3r_peer_close(void *region);
int a3r_peer_free(void *region);
/* a3r_ga_run for a frame-window sharded problem: prepare, scan into the shared partial-sum buffer, exchange kernel,
 * update.  `tag0`: iteration tags tag0+1 .. tag0+count must be larger than any tag used before on these regions and
 * identical on all ranks (they run the same iterations in lockstep). */
int a3r_ga_run_peers(const A3rProblem *p, const A3rPeers *peers, uint32_t tag0, int32_t count, const uint32_t *modes,
                     const A3rAdamStep *steps, void *stream);

/* diagnostic (no reference counterpart): phase trace of the small-parameter kernel.  enable != 0 switches the stamps on for
 * the following launches; out (host, 64 entries, may be NULL) receives the %globaltimer stamps [cta][8] of the last
 * traced launch.  Synchronises the device. */
int a3r_debug_k2_trace(int enable, unsigned long long *out);
/* diagnostic: while the trace is enabled, CTA 0 of every scan (id 1) / update (id 2) launch records (id, entry, return of its
 * programmatic-launch wait, exit) in %globaltimer ns; out (host, 64 x 4 entries) receives the last <= 64 launches oldest first,
 * *count how many.  Synchronises the device. */
int a3r_debug_timeline(unsigned long long *out, int *count);

/* ---- read-out (reference optimizer.py:468-478 depth_to_pts3d) ----
 * world points [n_local_frames][pix_stride][3] from logdepth + frame_par (call a3r_ga_prepare first) */
int a3r_ga_unproject(const A3rProblem *p, float *pts3d, void *stream);

/* ---- post-step: confidence clean-up (reference base_opt.py:293-303, 587-623 clean_pointcloud) ----
 * For i = 0..n-1 in order, every point of frame i is projected into every other camera j (in order); where it lands
 * in front of j's depth map ((1-tol) factor) while being less confident than the pixel it lands on, its confidence is
 * clipped to bad_conf.  One launch per frame i (a pixel's j-loop is sequential, frames are sequential, pixels of one
 * frame are independent).  pts3d [n][pix_stride][3] world points, depth [n][pix_stride], w2c [n][16] row-major
 * world-to-camera, intr [n][4] = fx fy cx cy, hw [n][2] (device int32), conf [n][pix_stride] in/out. */
int a3r_clean_pointcloud(const float *pts3d, const float *depth, const float *w2c, const float *intr, const int32_t *hw,
                         int32_t n, int64_t pix_stride, float tol, float bad_conf, float *conf, void *stream);
/* the same for frames i0 <= i < i1 only (a frame-window shard): pts3d holds just those frames ([i1-i0][pix_stride][3], row
 * i - i0), depth / conf / w2c / intr / hw cover all n frames (gathered by the caller; frames < i0 must already hold their
 * final confidences - the windows are processed in rank order). */
int a3r_clean_pointcloud_range(const float *pts3d, const float *depth, const float *w2c, const float *intr, const int32_t *hw,
                               int32_t n, int32_t i0, int32_t i1, int64_t pix_stride, float tol, float bad_conf, float *conf,
                               void *stream);

/* ---- pre-step: self-estimated dynamic masks (reference optimizer.py:156-235 get_motion_mask_from_pairs, use_self_mask=True;
 *      ego flow = goem_opt.py:195-236 warp_by_disp) ----
 * An item is one side of a symmetric pair: the flow plane of local slot item_slot[k] is compared, on that slot's frame, with
 * the ego flow induced by the pair's two-view cameras (pair_viewer.py:21-78; PnP solved on the host).  item_cam [k][16] fp32:
 * homography H = K_t relR K_s^-1 (9, row-major), K_t relt (3), depth row (4): depth = row . (X, Y, Z, 1) of the points in
 * local slot item_dslot[k] (the same frame's grid).  item_hw [k][2] = (H, W).  All index arrays are device int32.
 * a3r_motion_minmax: per item min / max of the error map |ego - flow|_2 -> minmax [k][2] = (~bits(min), bits(max)).
 * a3r_motion_masks: per local frame and pixel the mean over the frame's items (frame_items[frame_item_begin[f] ..
 * frame_item_begin[f+1]), in the reference's stacking order) of (err - min) / (max - min), compared with `thre`:
 * dynamic_mask [n_local_frames][pix_stride] (1 = dynamic), static_mask = its complement (A3rProblem.static_mask),
 * mean_out the mean itself (optional, may be NULL). */
int a3r_motion_minmax(const float *predw, const float *flow, int64_t pix_stride, const int32_t *item_slot,
                      const int32_t *item_dslot, const int32_t *item_hw, const float *item_cam, int32_t n_items,
                      uint32_t *minmax, void *stream);
int a3r_motion_masks(const float *predw, const float *flow, int64_t pix_stride, const int32_t *item_slot,
                     const int32_t *item_dslot, const float *item_cam, const uint32_t *minmax,
                     const int32_t *frame_item_begin, const int32_t *frame_items, const int32_t *lframe_hw,
                     int32_t n_local_frames, float thre, uint8_t *static_mask, uint8_t *dynamic_mask, float *mean_out,
                     void *stream);

/* ---- pre-step: flow validity maps (reference goem_opt.py:575-618 OccMask, computed by get_flow, optimizer.py:146-147, and
 *      stored as flow_valid_mask_i / _j; the loss never reads them) ----
 * Item k: forward flow = the flow plane of local slot fwd_slot[k], backward flow = the plane of bwd_slot[k] (the edge's other
 * side; uniform H x W).  valid [n_items][H*W] = 1 where the forward-backward inconsistency |sum over both channels of
 * (fwd + bwd sampled at p + fwd)| is below th (3 in the reference) and p + fwd stays inside the image.  Index arrays are
 * device int32. */
int a3r_occ_masks(const float *flow, int64_t pix_stride, const int32_t *fwd_slot, const int32_t *bwd_slot, int32_t n_items,
                  int32_t H, int32_t W, float th, uint8_t *valid, void *stream);

/* ---- MST initialisation solves (reference init_im_poses.py:446-465, post_process.py:36-58) ---- */
/* count weighted similarity registrations y ~ s R x + t between two slots living on the same frame:
 * x = slot xs[k] (weights = raw conf = conf_untrf(w)), y = slot ys[k]; P[k] valid pixels.
 * out [count][16] fp64: s, R(9 row-major), t(3), -.  xs/ys/P are device int32 arrays.
 * sums [count][20] fp64 scratch. */
int a3r_procrustes_batched(const float *predw, int64_t pix_stride, const int32_t *xs, const int32_t *ys,
                           const int32_t *npix, int32_t count, int32_t conf_mode, double *sums, double *out,
                           void *stream);
/* count Weiszfeld focal estimates (closed-form L2 start + 10 IRLS steps) of slots sl[k] with image size
 * hw[k] = (H, W) and principal point (W/2, H/2).  out_focal [count] fp64; sums [count][2] fp64 scratch. */
int a3r_weiszfeld_focal_batched(const float *predw, int64_t pix_stride, const int32_t *sl, const int32_t *hw,
                                int32_t count, double *sums, double *out_focal, void *stream);
/* log-depth planes from transformed source slots (reference init_im_poses.py:138-142,
 * optimizer.py:430-436): for local frame k, depth = row[k] . (X_src, 1) with row [k][4] fp64 and the
 * source slot src[k]; writes logdepth (log, nan/-inf -> 0). */
int a3r_init_depth(const float *predw, int64_t pix_stride, const int32_t *src, const int32_t *npix,
                   const double *rows, int32_t n_local_frames, float *logdepth, void *stream);

/* The inlier test of ONE RANSAC-PnP hypothesis over all points (reference init_im_poses.py:498 -> cv::solvePnPRansac, whose
 * PnPRansacCallback::computeError re-projects every point for every hypothesis on the host).  pts [P][3] fp32 world points of
 * a frame (pixel p = v * W + u is observed at (u, v)), mask [P] bytes (the points handed to the solver), hyp = 17 HOST doubles:
 * world-to-camera R (9, row-major), t (3), fx, fy, cx, cy, squared re-projection threshold.  Writes flags [P] (1 = masked-in
 * inlier; NULL: the count only) and *count (device int32).  Operation by operation OpenCV's arithmetic (fp64 without FMA contraction, projection
 * rounded to fp32, fp32 squared error, fp64 comparison): the inlier sets are cv::solvePnPRansac's. */
int a3r_pnp_inliers(const float *pts, const uint8_t *mask, int32_t W, int64_t P, const double *hyp, uint8_t *flags,
                    int32_t *count, void *stream);

/* World points of a frame from a raw slot plane [P][4] (X Y Z w): out [P][3] = X * A[:,0] + Y * A[:,1] + Z * A[:,2] + t in fp32,
 * every operation rounded on its own in numpy's order (the similarity of the spanning tree, reference init_im_poses.py:272-279
 * applies it before fast_pnp).  sim = 12 HOST floats: A (9, row-major), t (3). */
int a3r_similarity_points(const float *plane, int64_t P, const float *sim, float *out, void *stream);
/* The n <= 64 sampled points of a batch of RANSAC hypotheses: row idx[k] (HOST int32) of the solver's point list, whose k-th
 * entry is pixel sel[k] (device int64, ascending), -> out [n][4] = x, y, z, pixel. */
int a3r_pnp_gather(const float *pts, const int64_t *sel, const int32_t *idx, int32_t n, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* A3R_GA_H */
