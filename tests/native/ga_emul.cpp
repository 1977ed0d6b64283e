// ga_emul.cpp - TEST HARNESS ONLY.  Serial host re-enactment of the three hot-loop kernels
// (prepare / accumulate / update) that calls the very same per-element math the CUDA kernels use
// (align3r_b200/csrc/ga_math.cuh), on an A3rProblem whose pointers are HOST pointers.
// It lets the `not gpu` test-suite check the hand-derived backward, the parameter chains, the gate
// and the Adam arithmetic against the autograd oracle without a GPU.  It is built by
// tests/test_math_emulation.py with g++ into tests/native/_build/ and is never imported, linked or
// shipped by the align3r_b200 package: the product has no CPU path.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "a3r_ga.h"
#include "ga_math.cuh"

using namespace a3r;

static bool flow_on(const A3rProblem* p, uint32_t mode) {
    return (mode & A3R_M_FLOW_ACTIVE) && (p->flags & A3R_F_HAS_FLOW) && p->flow_loss_weight > 0.f;
}

extern "C" int emul_prepare(const A3rProblem* p) {
    const int E = p->n_edges, N = p->n_frames;
    const bool shared_focal = p->flags & A3R_F_SHARED_FOCAL, norm_scale = p->flags & A3R_F_NORM_PW_SCALE;
    double mean = 0;
    for (int e = 0; e < E; ++e) mean += (double)p->pw_poses[8 * e + 7];
    mean /= E;
    const double log_norm = norm_scale ? log((double)p->base_scale) - mean : 0.0;
    for (int n = 0; n < N; ++n) {
        FrameD F;
        load_frame(p->im_poses, p->im_focals, p->im_pp, p->frame_hw, n, shared_focal, p->focal_break, F);
        derive_frame_par(F, p->frame_par + FP_STRIDE * n);
    }
    for (int e = 0; e < E; ++e)
        derive_edge_par(p->pw_poses + 8 * e, p->pw_adaptors + 2 * e, log_norm, norm_scale, p->pw_break, p->edge_par + EP_STRIDE * e);
    if (p->flags & A3R_F_HAS_FLOW)
        for (int gs = 0; gs < 2 * E; ++gs) {
            const int e = gs >> 1, side = gs & 1, i = p->edge_ij[2 * e], j = p->edge_ij[2 * e + 1];
            FrameD S, T;
            load_frame(p->im_poses, p->im_focals, p->im_pp, p->frame_hw, side ? j : i, shared_focal, p->focal_break, S);
            load_frame(p->im_poses, p->im_focals, p->im_pp, p->frame_hw, side ? i : j, shared_focal, p->focal_break, T);
            derive_slot_par(S, T, p->slot_par + SP_STRIDE * gs);
        }
    return 0;
}

// which form of the flow pixel-slot emul_accumulate evaluates: 0 = flow_px (the v1 / v5 cross-check scans), 1 = flow_px_scan
// (the product scan's general form), 2 = its row-tile form (v as the chunk's constant; a "chunk" here is one pixel)
static int g_flow_form = 0;
extern "C" void emul_set_flow_form(int form) { g_flow_form = form; }

extern "C" int emul_accumulate(const A3rProblem* p, uint32_t mode) {
    const int E = p->n_edges, N = p->n_frames;
    const bool flow = flow_on(p, mode), l2 = p->flags & A3R_F_DIST_L2;
    memset(p->acc_slot, 0, sizeof(double) * 2 * E * ACC_STRIDE);
    memset(p->acc_frame, 0, sizeof(double) * N * FACC_N);
    const size_t plane = (size_t)p->n_local_frames * p->pix_stride;
    memset(p->g_align, 0, sizeof(float) * plane);
    if (flow) {
        memset(p->g_flow_i, 0, sizeof(float) * plane);
        memset(p->g_flow_j, 0, sizeof(float) * plane);
    }
    for (int wf = 0; wf < p->n_work_frames; ++wf) {
        const int lf = p->wf_lframe[wf], n = p->lframe_id[lf];
        const int W = p->frame_hw[2 * n + 1], P = p->frame_hw[2 * n] * W;
        const int s0 = p->wf_slot_begin[wf], s1 = p->wf_slot_begin[wf + 1];
        const float* fp = p->frame_par + FP_STRIDE * n;
        const int64_t base = (int64_t)lf * p->pix_stride;
        for (int pix = 0; pix < P; ++pix) {
            PixState st;
            pixel_state(fp, p->logdepth[base + pix], pix % W, pix / W, st);
            const bool stat = flow && p->static_mask[base + pix] != 0;
            float gsx = 0, gsy = 0, gsz = 0, gfi = 0, gfj = 0;
            for (int s = s0; s < s1; ++s) {
                const int gs = p->lslot_gslot[s];
                SlotConst c;
                const float* ep = p->edge_par + EP_STRIDE * (gs >> 1);
                for (int k = 0; k < 9; ++k) c.nM[k] = -ep[EP_M + k];
                for (int k = 0; k < 3; ++k) c.st[k] = ep[EP_ST + k];
                const float* sp = p->slot_par + SP_STRIDE * gs;
                for (int k = 0; k < 9; ++k) c.A[k] = sp[SP_A + k];
                for (int k = 0; k < 3; ++k) c.b[k] = sp[SP_B + k];
                c.wn = (gs & 1) ? p->inv_area_j : p->inv_area_i;
                const float* x = p->predw + ((int64_t)s * p->pix_stride + pix) * 4;
                float acc[ACC_STRIDE];
                for (int k = 0; k < ACC_STRIDE; ++k) acc[k] = 0.f;
                float gx, gy, gz;
                if (l2) align_px<true>(c, x[0], x[1], x[2], x[3], st, acc, gx, gy, gz);
                else align_px<false>(c, x[0], x[1], x[2], x[3], st, acc, gx, gy, gz);
                gsx += gx; gsy += gy; gsz += gz;
                if (flow) {         // evaluated for every pixel, masked branch-free like the kernels
                    const float* f = p->flow + ((int64_t)s * p->pix_stride + pix) * 2;
                    const float thre_eff = flow_thre_eff(p->pxl_thre);
                    float gd;
                    if (g_flow_form == 0) gd = flow_px(c, st, f[0], f[1], thre_eff, stat, acc);
                    else {
                        c.A[8] += 1e-6f;                       // as the product scan stages its slot table
                        const float thre_pix = stat ? thre_eff : -1.0f;
                        if (g_flow_form == 1) gd = flow_px_scan<0, false>(c, st.u, st.v, st.disp, f[0], f[1], thre_pix, acc, 0.0f);
                        else {
                            flow_row_begin(c, st.v);
                            gd = flow_px_scan<0, true>(c, st.u, st.v, st.disp, f[0], f[1], thre_pix, acc, 0.0f);
                            flow_row_finish<0>(st.v, acc);
                        }
                    }
                    if (gs & 1) gfj += gd; else gfi += gd;
                }
                double* a = p->acc_slot + (int64_t)gs * ACC_STRIDE;
                for (int k = 0; k < ACC_N; ++k) a[k] += (double)acc[k];
            }
            float fa32[FACC_N];
            for (int k = 0; k < FACC_N; ++k) fa32[k] = 0.f;
            p->g_align[base + pix] += pixel_epilogue(fp, st, gsx, gsy, gsz, fa32);
            if (flow) {
                p->g_flow_i[base + pix] += gfi * st.dfac;
                p->g_flow_j[base + pix] += gfj * st.dfac;
            }
            double* fa = p->acc_frame + (int64_t)n * FACC_N;
            for (int k = 0; k < FACC_N; ++k) fa[k] += (double)fa32[k];
        }
    }
    return 0;
}

static void adam_vec(float* prm, float* m, float* v, const float* g, int cnt, const A3rProblem* p, const A3rAdamStep* a) {
    for (int k = 0; k < cnt; ++k) adam_update(prm[k], m[k], v[k], g[k], p->adam_beta1, p->adam_beta2, p->adam_eps, a->step_size, a->bc2_sqrt);
}

extern "C" int emul_update(const A3rProblem* p, uint32_t mode, const A3rAdamStep* adam) {
    const int E = p->n_edges, N = p->n_frames;
    const bool shared_focal = p->flags & A3R_F_SHARED_FOCAL, norm_scale = p->flags & A3R_F_NORM_PW_SCALE;
    const bool flow = flow_on(p, mode), apply = mode & A3R_M_APPLY_ADAM;
    const double fb = p->focal_break, tsw = p->temporal_smoothing_weight, tw = p->translation_weight;
    auto frame = [&](int n, FrameD& F) { load_frame(p->im_poses, p->im_focals, p->im_pp, p->frame_hw, n, shared_focal, fb, F); };

    double l_align = 0, num_i = 0, cnt_i = 0, num_j = 0, cnt_j = 0, l_temp = 0;
    for (int gs = 0; gs < 2 * E; ++gs) {
        const double* a = p->acc_slot + (int64_t)gs * ACC_STRIDE;
        l_align += a[ACC_LOSS];
        if (gs & 1) { num_j += a[ACC_FNUM]; cnt_j += a[ACC_FCNT]; } else { num_i += a[ACC_FNUM]; cnt_i += a[ACC_FCNT]; }
    }
    if (tsw > 0)
        for (int n = 0; n < N - 1; ++n) {
            FrameD T, S;
            frame(n, T); frame(n + 1, S);
            l_temp += temporal_pair(S, T, tsw, tw, nullptr);
        }
    const LossScalars r = assemble_loss(l_align, l_temp, num_i, cnt_i, num_j, cnt_j, flow, p->flow_loss_weight, p->flow_loss_thre, tsw);
    p->scal[A3R_S_LOSS] = r.loss; p->scal[A3R_S_ALIGN] = l_align; p->scal[A3R_S_TEMPORAL] = l_temp;
    p->scal[A3R_S_FLOW] = r.floss; p->scal[A3R_S_FLOW_DROPPED] = r.dropped; p->scal[A3R_S_SC_I] = r.sc_i; p->scal[A3R_S_SC_J] = r.sc_j;
    if (r.dropped != 0) p->scal[A3R_S_FLAG_STICKY] = 1.0;
    if (p->loss_hist && adam->iter >= 0) p->loss_hist[adam->iter] = (float)r.loss;

    for (int n = 0; n < N; ++n) {
        FrameD F;
        frame(n, F);
        frame_align_grad(F, p->acc_frame + (int64_t)n * FACC_N, p->frame_grad + (int64_t)n * FG_STRIDE);
    }
    auto scatter = [&](const PairGrad& g, int sI, int tI) {
        for (int k = 0; k < 15; ++k) {
            p->frame_grad[(int64_t)sI * FG_STRIDE + k] += g.s[k];
            p->frame_grad[(int64_t)tI * FG_STRIDE + k] += g.t[k];
        }
    };
    if (flow && (r.sc_i != 0 || r.sc_j != 0))
        for (int gs = 0; gs < 2 * E; ++gs) {
            const int e = gs >> 1, side = gs & 1, i = p->edge_ij[2 * e], j = p->edge_ij[2 * e + 1];
            const int sI = side ? j : i, tI = side ? i : j;
            FrameD S, T;
            frame(sI, S); frame(tI, T);
            PairGrad g;
            slot_flow_chain(S, T, p->acc_slot + (int64_t)gs * ACC_STRIDE, side ? r.sc_j : r.sc_i, g);
            scatter(g, sI, tI);
        }
    if (tsw > 0)
        for (int n = 0; n < N - 1; ++n) {
            FrameD T, S;
            frame(n, T); frame(n + 1, S);
            PairGrad g;
            temporal_pair(S, T, tsw, tw, &g);
            scatter(g, n + 1, n);
        }
    double gsum = 0;
    for (int e = 0; e < E; ++e) {
        const double* a0 = p->acc_slot + (int64_t)(2 * e) * ACC_STRIDE;
        p->edge_scratch[e] = edge_logscale_grad(p->edge_par + EP_STRIDE * e, a0, a0 + ACC_STRIDE);
        gsum += p->edge_scratch[e];
    }
    const double gmean = norm_scale ? gsum / E : 0.0;
    for (int e = 0; e < E; ++e) {
        const double* a0 = p->acc_slot + (int64_t)(2 * e) * ACC_STRIDE;
        float g[8], ga[2];
        edge_pose_grad(p->pw_poses + 8 * e, p->edge_par + EP_STRIDE * e, a0, a0 + ACC_STRIDE, p->edge_scratch[e] - gmean,
                       p->flags & A3R_F_TRAIN_ADAPTORS, norm_scale, p->pw_break, g, ga);
        for (int k = 0; k < 8; ++k) p->g_pw_poses[8 * e + k] = g[k];
        p->g_pw_adaptors[2 * e] = ga[0]; p->g_pw_adaptors[2 * e + 1] = ga[1];
        if (apply && (p->flags & A3R_F_TRAIN_PW_POSES)) adam_vec(p->pw_poses + 8 * e, p->pw_poses_m + 8 * e, p->pw_poses_v + 8 * e, g, 8, p, adam);
        if (apply && (p->flags & A3R_F_TRAIN_ADAPTORS)) adam_vec(p->pw_adaptors + 2 * e, p->pw_adaptors_m + 2 * e, p->pw_adaptors_v + 2 * e, ga, 2, p, adam);
    }
    // gradients of all frames first (they read the shared focal), then the updates
    std::vector<float> g7(7 * N), gpp(2 * N);
    std::vector<double> gphi(N);
    for (int n = 0; n < N; ++n) {
        FrameD F;
        frame(n, F);
        frame_pose_grad(p->im_poses + 7 * n, F, p->frame_grad + (int64_t)n * FG_STRIDE, fb, p->flags & A3R_F_TRAIN_PP, &g7[7 * n], &gphi[n], &gpp[2 * n]);
    }
    double gphi_sum = 0;
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < 7; ++k) p->g_im_poses[7 * n + k] = g7[7 * n + k];
        p->g_im_pp[2 * n] = gpp[2 * n]; p->g_im_pp[2 * n + 1] = gpp[2 * n + 1];
        if (shared_focal) gphi_sum += gphi[n];
        else {
            float gf = (float)gphi[n];
            p->g_im_focals[n] = gf;
            if (apply && (p->flags & A3R_F_TRAIN_FOCALS)) adam_vec(p->im_focals + n, p->im_focals_m + n, p->im_focals_v + n, &gf, 1, p, adam);
        }
        if (apply && (p->flags & A3R_F_TRAIN_PP)) adam_vec(p->im_pp + 2 * n, p->im_pp_m + 2 * n, p->im_pp_v + 2 * n, &gpp[2 * n], 2, p, adam);
        if (apply && (p->flags & A3R_F_TRAIN_IM_POSES)) adam_vec(p->im_poses + 7 * n, p->im_poses_m + 7 * n, p->im_poses_v + 7 * n, &g7[7 * n], 7, p, adam);
    }
    if (shared_focal) {
        float gf = (float)gphi_sum;
        p->g_im_focals[0] = gf;
        if (apply && (p->flags & A3R_F_TRAIN_FOCALS)) adam_vec(p->im_focals, p->im_focals_m, p->im_focals_v, &gf, 1, p, adam);
    }
    // depth planes
    if (p->n_local_frames > 0 && (!apply || (p->flags & A3R_F_TRAIN_DEPTH))) {
        const int64_t total = (int64_t)p->n_local_frames * p->pix_stride;
        const float sci = (float)r.sc_i, scj = (float)r.sc_j;
        for (int64_t i = 0; i < total; ++i) {
            float g = p->g_align[i];
            if (flow) g += sci * p->g_flow_i[i] + scj * p->g_flow_j[i];
            if (apply) adam_update(p->logdepth[i], p->logdepth_m[i], p->logdepth_v[i], g, p->adam_beta1, p->adam_beta2, p->adam_eps, adam->step_size, adam->bc2_sqrt);
            else p->g_align[i] = g;
        }
    }
    return 0;
}
