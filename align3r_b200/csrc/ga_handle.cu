// ga_handle.cu - handle-style layer of the C ABI (include/a3r_ga.h, "handle layer"): the export set SURVEY.md
// section 8(b) names for this path - create / destroy / pack_inputs / step / forward / grads / mst_edge_scores - written
// over the stateless entry points of ga_kernels.cu.  Host code only, except the E-thread edge-score kernel; nothing here
// is on the hot loop's critical path (the Python class drives a3r_ga_run / a3r_ga_run_peers directly).
#include <math.h>
#include <new>

#include "a3r_ga.h"
#include "a3r_common.cuh"

struct A3rHandle {
    A3rProblem problem;      // the caller's descriptor (every pointer in it stays the caller's memory)
    int device;              // device that was current at creation: every later call must run on it
    int64_t adam_t;          // Adam steps taken since creation / a3r_ga_reset_adam (torch's state['step'])
};

namespace {

// The descriptor carries the betas as fp32.  torch computes the bias corrections in python doubles from the decimal
// value the user wrote (0.9, base_opt.py:524); an fp32 holds 7 significant digits, so snapping to a 1e-7 grid gives
// that double back and the step scalars equal the ones the Python side computes (align3r_b200/problem.py:adam_schedule).
double decimal_of(float beta) { return round((double)beta * 1e7) / 1e7; }

int check_handle(const A3rHandle* h, const char* who) {
    if (!h) return a3r_fail(A3R_E_ARG, "%s: handle is NULL", who);
    int dev = -1;
    A3R_CUDA(cudaGetDevice(&dev));
    if (dev != h->device) return a3r_fail(A3R_E_ARG, "%s: handle belongs to device %d, current device is %d", who, h->device, dev);
    return A3R_OK;
}

int check_unsharded(const A3rHandle* h, const char* who) {
    const A3rProblem& p = h->problem;
    if (p.n_local_frames != p.n_frames || p.n_local_slots != 2 * p.n_edges)
        return a3r_fail(A3R_E_UNSUPPORTED, "%s: the handle layer drives one device holding every frame; a frame-window shard "
                        "exchanges its accumulator block between a3r_ga_accumulate and a3r_ga_update (or uses a3r_ga_run_peers)", who);
    return A3R_OK;
}

__global__ void __launch_bounds__(256) k_edge_scores(const double* __restrict__ conf_sum, const int32_t* __restrict__ edge_ij,
                                                     const int32_t* __restrict__ frame_hw, int n_edges, double* __restrict__ scores) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int i = edge_ij[2 * e], j = edge_ij[2 * e + 1];
    const double area_i = (double)frame_hw[2 * i] * (double)frame_hw[2 * i + 1];
    const double area_j = (double)frame_hw[2 * j] * (double)frame_hw[2 * j + 1];
    const float mi = (float)(conf_sum[2 * e] / area_i), mj = (float)(conf_sum[2 * e + 1] / area_j);     // fp32 means, fp32 product
    scores[e] = (double)__fmul_rn(mi, mj);
}

}  // namespace

extern "C" {

int a3r_ga_create(const A3rProblem* desc, A3rHandle** out) {
    if (!desc || !out) return a3r_fail(A3R_E_ARG, "a3r_ga_create: bad arguments");
    *out = nullptr;
    if (desc->tile_pixels != a3r_tile_pixels())
        return a3r_fail(A3R_E_ARG, "a3r_ga_create: tile_pixels %d != a3r_tile_pixels() %d", desc->tile_pixels, a3r_tile_pixels());
    if (desc->n_frames <= 0 || desc->n_edges <= 0 || desc->n_local_frames < 0 || desc->n_ctas <= 0 || desc->pix_stride <= 0 ||
        desc->pix_stride % desc->tile_pixels != 0)
        return a3r_fail(A3R_E_ARG, "a3r_ga_create: bad sizes N=%d E=%d local=%d ctas=%d pix_stride=%lld", desc->n_frames, desc->n_edges,
                        desc->n_local_frames, desc->n_ctas, (long long)desc->pix_stride);
    if (!desc->edge_ij || !desc->frame_hw || !desc->pw_poses || !desc->im_poses || !desc->im_focals || !desc->im_pp || !desc->pw_adaptors ||
        !desc->frame_par || !desc->edge_par || !desc->slot_par || !desc->acc_slot || !desc->acc_frame || !desc->frame_grad ||
        !desc->edge_scratch || !desc->scal)
        return a3r_fail(A3R_E_ARG, "a3r_ga_create: a required pointer of the descriptor is NULL");
    if (desc->n_local_frames > 0 && (!desc->predw || !desc->logdepth || !desc->g_align || !desc->lframe_id || !desc->cta_tile_begin))
        return a3r_fail(A3R_E_ARG, "a3r_ga_create: a per-pixel plane or plan table of the descriptor is NULL");
    int dev = -1;
    A3R_CUDA(cudaGetDevice(&dev));
    A3rHandle* h = new (std::nothrow) A3rHandle;
    if (!h) return a3r_fail(A3R_E_ARG, "a3r_ga_create: out of host memory");
    h->problem = *desc;
    h->device = dev;
    h->adam_t = 0;
    *out = h;
    return A3R_OK;
}

int a3r_ga_destroy(A3rHandle* h) {
    delete h;                      // owns nothing on the device: every buffer is the caller's
    return A3R_OK;
}

const A3rProblem* a3r_ga_problem(const A3rHandle* h) { return h ? &h->problem : nullptr; }

int a3r_ga_pack_inputs(A3rHandle* h, const float* pts, const float* conf, const float* flow, int32_t n, int64_t P,
                       const int32_t* dst_slot, const int32_t* dst_lframe, int32_t conf_mode, float* im_conf, double* conf_sum,
                       void* stream) {
    if (int rc = check_handle(h, "a3r_ga_pack_inputs")) return rc;
    const A3rProblem& p = h->problem;
    if (n < 0 || P <= 0 || P > p.pix_stride || !dst_slot || !dst_lframe || (n > 0 && (!pts || !conf)))
        return a3r_fail(A3R_E_ARG, "a3r_ga_pack_inputs: bad arguments (n=%d, P=%lld, pix_stride=%lld)", n, (long long)P, (long long)p.pix_stride);
    if (flow && !p.flow) return a3r_fail(A3R_E_ARG, "a3r_ga_pack_inputs: flow rows given but the descriptor has no flow plane");
    if (n == 0) return A3R_OK;
    // the planes are constant inputs of the loop (const in the descriptor); packing is the one place that writes them
    if (int rc = a3r_pack_pred(pts, conf, n, P, dst_slot, dst_lframe, p.pix_stride, conf_mode, const_cast<float*>(p.predw), im_conf,
                               conf_sum, stream))
        return rc;
    if (flow)
        if (int rc = a3r_pack_flow(flow, n, P, dst_slot, p.pix_stride, const_cast<float*>(p.flow), stream)) return rc;
    return A3R_OK;
}

int a3r_ga_reset_adam(A3rHandle* h, void* stream) {
    if (int rc = check_handle(h, "a3r_ga_reset_adam")) return rc;
    const A3rProblem& p = h->problem;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)p.n_edges, N = (size_t)p.n_frames, nf = (p.flags & A3R_F_SHARED_FOCAL) ? 1 : N;
    struct { float* ptr; size_t count; } bufs[] = {
        {p.pw_poses_m, E * 8}, {p.pw_poses_v, E * 8}, {p.pw_adaptors_m, E * 2}, {p.pw_adaptors_v, E * 2},
        {p.im_poses_m, N * 7}, {p.im_poses_v, N * 7}, {p.im_focals_m, nf}, {p.im_focals_v, nf}, {p.im_pp_m, N * 2}, {p.im_pp_v, N * 2},
        {p.logdepth_m, (size_t)p.n_local_frames * (size_t)p.pix_stride}, {p.logdepth_v, (size_t)p.n_local_frames * (size_t)p.pix_stride}};
    for (auto& b : bufs)
        if (b.ptr && b.count) A3R_CUDA(cudaMemsetAsync(b.ptr, 0, sizeof(float) * b.count, st));
    h->adam_t = 0;
    return A3R_OK;
}

int a3r_ga_step(A3rHandle* h, int32_t iter, int32_t niter, float lr, uint32_t flags, void* stream) {
    if (int rc = check_handle(h, "a3r_ga_step")) return rc;
    if (int rc = check_unsharded(h, "a3r_ga_step")) return rc;
    if (niter <= 0 || iter < 0 || iter >= niter) return a3r_fail(A3R_E_ARG, "a3r_ga_step: iteration %d of %d", iter, niter);
    if (!(lr >= 0.f)) return a3r_fail(A3R_E_ARG, "a3r_ga_step: bad learning rate");
    const A3rProblem& p = h->problem;
    const double t = (double)(h->adam_t + 1);
    const double b1 = decimal_of(p.adam_beta1), b2 = decimal_of(p.adam_beta2);
    A3rAdamStep s;
    s.step_size = (float)((double)lr / (1.0 - pow(b1, t)));      // torch/optim/adam.py: lr / bias_correction1
    s.bc2_sqrt = (float)sqrt(1.0 - pow(b2, t));
    s.iter = p.loss_hist ? iter : -1;
    s._pad = 0;
    const uint32_t mode = (flags & A3R_M_FLOW_ACTIVE) | A3R_M_APPLY_ADAM;
    if (int rc = a3r_ga_run(&p, 1, &mode, &s, stream)) return rc;
    h->adam_t += 1;
    return A3R_OK;
}

int a3r_ga_forward(A3rHandle* h, uint32_t flags, double* loss_out, void* stream) {
    if (int rc = check_handle(h, "a3r_ga_forward")) return rc;
    if (int rc = check_unsharded(h, "a3r_ga_forward")) return rc;
    const A3rProblem& p = h->problem;
    const uint32_t mode = flags & A3R_M_FLOW_ACTIVE;                 // no A3R_M_APPLY_ADAM: loss + gradients only
    A3rAdamStep s;
    s.step_size = 0.f; s.bc2_sqrt = 1.f; s.iter = -1; s._pad = 0;
    if (int rc = a3r_ga_prepare(&p, mode, stream)) return rc;
    if (int rc = a3r_ga_accumulate(&p, mode, stream)) return rc;
    if (int rc = a3r_ga_update(&p, mode, &s, stream)) return rc;
    if (loss_out) A3R_CUDA(cudaMemcpyAsync(loss_out, p.scal + A3R_S_LOSS, sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
    return A3R_OK;
}

int a3r_ga_grads(const A3rHandle* h, A3rGrads* out) {
    if (!h || !out) return a3r_fail(A3R_E_ARG, "a3r_ga_grads: bad arguments");
    const A3rProblem& p = h->problem;
    out->pw_poses = p.g_pw_poses;
    out->pw_adaptors = p.g_pw_adaptors;
    out->im_poses = p.g_im_poses;
    out->im_focals = p.g_im_focals;
    out->im_pp = p.g_im_pp;
    out->logdepth = p.g_align;      // after a3r_ga_forward: the whole d loss / d log-depth (alignment + flow + prior)
    out->scal = p.scal;
    return A3R_OK;
}

int a3r_mst_edge_scores(const double* conf_sum, const int32_t* edge_ij, const int32_t* frame_hw, int32_t n_edges, double* scores,
                        void* stream) {
    if (n_edges < 0 || (n_edges > 0 && (!conf_sum || !edge_ij || !frame_hw || !scores)))
        return a3r_fail(A3R_E_ARG, "a3r_mst_edge_scores: bad arguments");
    if (n_edges == 0) return A3R_OK;
    k_edge_scores<<<(n_edges + 255) / 256, 256, 0, (cudaStream_t)stream>>>(conf_sum, edge_ij, frame_hw, n_edges, scores);
    A3R_CUDA(cudaGetLastError());
    return A3R_OK;
}

}  // extern "C"
