// ga_math.cuh - per-element math of the global-alignment kernels, shared by every kernel in
// ga_kernels.cu.  Everything is __host__ __device__ so that tests/native/ga_emul.cpp can exercise
// exactly the same arithmetic on the CPU against the oracle (test harness only - there is no CPU
// product path).
//
// Notation: edge e=(i,j); slot = edge-side; x = a (.) X; y = s (R_e x + t_e);
// c = d ((u-cx)/f, (v-cy)/f, 1), d = exp(l); z = R_n c + t_n; r = z - y; rho = |r|.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define A3R_HD __host__ __device__ __forceinline__
#else
#define A3R_HD inline
#endif

namespace a3r {

// fp64 exp / expm1: ~100-instruction inline sequences on the device.  The small-parameter kernel is run-once,
// latency-bound code whose instruction fetch matters: one out-of-line copy of each instead of two dozen inlined ones.
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ double a3r_exp(double x) { return exp(x); }
static __device__ __noinline__ double a3r_expm1(double x) { return expm1(x); }
#else
inline double a3r_exp(double x) { return exp(x); }
inline double a3r_expm1(double x) { return expm1(x); }
#endif

// accumulator slots (A3R_SLOT_ACC = 28 per edge-side)
enum : int {
    ACC_SG = 0,      // sum g                (3)
    ACC_SGX = 3,     // sum g X^T            (9, row-major r*3+c)
    ACC_LOSS = 12,   // sum w rho / area     (1)
    ACC_GA = 13,     // sum g_q (u,v,1)^T    (9)
    ACC_GB = 22,     // sum disp * g_q       (3)
    ACC_FNUM = 25,   // sum raw * ppm        (1)
    ACC_FCNT = 26,   // sum ppm              (1)
    ACC_N = 27,
    ACC_STRIDE = 28
};
// whole-problem sums the scan adds up beside the per-slot ones (A3R_GLOB_ACC = 8, 5 used): the gate and the loss need them
// before anything else can start
enum : int { GL_LOSS = 0, GL_FNUM_I = 1, GL_FCNT_I = 2, GL_FNUM_J = 3, GL_FCNT_J = 4 };
// per-frame accumulators (A3R_FRAME_ACC = 12), in the CAMERA frame: with h = R^T gsum (gsum = sum over the
// pixel's slots of dL/dz): sum h (3), sum h c^T (9).  The focal / principal-point gradients are single entries
// of sum h c^T (no cancellation between separately rounded sums).
enum : int { FACC_G = 0, FACC_GC = 3, FACC_N = 12 };
// frame_grad scratch rows (16 doubles): dL/dR (9), dL/dt (3), dL/df, dL/dcx, dL/dcy
enum : int { FG_R = 0, FG_T = 9, FG_F = 12, FG_CX = 13, FG_CY = 14, FG_STRIDE = 16 };
// frame_par rows (16 floats): R (9), t (3), f, cx, cy, 1/f
enum : int { FP_R = 0, FP_T = 9, FP_F = 12, FP_CX = 13, FP_CY = 14, FP_INVF = 15, FP_STRIDE = 16 };
// edge_par rows (16 floats): M = s R diag(a) (9), s t (3), s, a (3)
enum : int { EP_M = 0, EP_ST = 9, EP_S = 12, EP_A = 13, EP_STRIDE = 16 };
// slot_par rows (16 floats): A (9), b (3)
enum : int { SP_A = 0, SP_B = 9, SP_STRIDE = 16 };

// single-MUFU reciprocal square root / reciprocal (flush-to-zero, ~1 ulp): the scan is issue-bound, the
// IEEE-denormal fix-up sequences of rsqrtf()/__fdividef() cost 4-5 extra instructions per call
A3R_HD float fast_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
A3R_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// sign(e) min(|e|, 1) = d smooth_l1(e) / d e in one instruction (FMNMX.XORSIGN.ABS: the sign of the result is the xor of the
// operands' signs); its magnitude is the min(|e|, 1) of the loss's min form
A3R_HD float clamp_unit_signed(float e) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(e), "f"(1.0f));
    return r;
#else
    return copysignf(fminf(fabsf(e), 1.0f), e);
#endif
}

// ------------------------------------------------------------------------------------------
// parameter maps (fp64 in, rounded to fp32 by the caller)
// ------------------------------------------------------------------------------------------
// unit quaternion (x,y,z,w) -> rotation, the roma/scipy convention used by reference base_opt.py:206
A3R_HD void quat_to_rot(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

A3R_HD double signed_expm1_d(double x) {   // reference commons.py:83-85
    const double a = a3r_expm1(fabs(x));
    return x > 0 ? a : (x < 0 ? -a : 0.0);
}

// pose vector [qx qy qz qw tx~ ty~ tz~] -> normalised quaternion, R, t  (reference base_opt.py:202-207)
A3R_HD void pose_to_Rt(const float* pv, double qn[4], double* inv_norm, double R[9], double t[3]) {
    double q[4] = {(double)pv[0], (double)pv[1], (double)pv[2], (double)pv[3]};
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double inv = 1.0 / n;
    for (int k = 0; k < 4; ++k) qn[k] = q[k] * inv;
    *inv_norm = inv;
    quat_to_rot(qn, R);
    for (int k = 0; k < 3; ++k) t[k] = signed_expm1_d((double)pv[4 + k]);
}

// dL/dq (raw, un-normalised quaternion) from dL/dR; chain of quat_to_rot and q/|q|
A3R_HD void rot_grad_to_quat(const double G[9], const double qn[4], double inv_norm, double gq[4]) {
    const double x = qn[0], y = qn[1], z = qn[2], w = qn[3];
    double h[4];
    // <G, dR/dx> etc. (SURVEY.md 8(a) chain block)
    h[0] = 2 * (y * G[1] + z * G[2] + y * G[3] - 2 * x * G[4] - w * G[5] + z * G[6] + w * G[7] - 2 * x * G[8]);
    h[1] = 2 * (-2 * y * G[0] + x * G[1] + w * G[2] + x * G[3] + z * G[5] - w * G[6] + z * G[7] - 2 * y * G[8]);
    h[2] = 2 * (-2 * z * G[0] - w * G[1] + x * G[2] + w * G[3] - 2 * z * G[4] + y * G[5] + x * G[6] + y * G[7]);
    h[3] = 2 * (-z * G[1] + y * G[2] + z * G[3] - x * G[5] - y * G[6] + x * G[7]);
    const double dot = h[0] * x + h[1] * y + h[2] * z + h[3] * w;
    for (int k = 0; k < 4; ++k) gq[k] = (h[k] - dot * qn[k]) * inv_norm;
}

// d/dt~ of sign(t~) expm1(|t~|): exp(|t~|), and exactly 0 at t~ == 0 (torch.sign(0) = 0)
A3R_HD double signed_expm1_grad(float tt) { return tt != 0.0f ? a3r_exp(fabs((double)tt)) : 0.0; }

A3R_HD void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
A3R_HD void mat3_tmul(const double A[9], const double B[9], double C[9]) {   // A^T B
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
A3R_HD void mat3_mult(const double A[9], const double B[9], double C[9]) {   // A B^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c * 3] + A[r * 3 + 1] * B[c * 3 + 1] + A[r * 3 + 2] * B[c * 3 + 2];
}

struct FrameD {   // fp64 view of one frame's camera
    double R[9], t[3], f, cx, cy;
    double qn[4], inv_norm;
};

A3R_HD void load_frame(const float* im_poses, const float* im_focals, const float* im_pp, const int32_t* frame_hw,
                       int n, bool shared_focal, double focal_break, FrameD& F) {
    pose_to_Rt(im_poses + 7 * n, F.qn, &F.inv_norm, F.R, F.t);
    const double phi = (double)im_focals[shared_focal ? 0 : n];
    F.f = a3r_exp(phi / focal_break);                                    // reference optimizer.py:398-403
    F.cx = 0.5 * frame_hw[2 * n + 1] + 10.0 * (double)im_pp[2 * n];     // optimizer.py:76,415-416
    F.cy = 0.5 * frame_hw[2 * n] + 10.0 * (double)im_pp[2 * n + 1];
}

// flow homography of slot (src -> tgt): A = K_t R_t^T R_s K_s^-1, b = K_t R_t^T (t_s - t_t)
// (reference goem_opt.py:149-154, 195-236)
A3R_HD void flow_homography(const FrameD& S, const FrameD& T, double A[9], double b[3],
                            double relR[9], double relt[3]) {
    mat3_tmul(T.R, S.R, relR);
    const double dt[3] = {S.t[0] - T.t[0], S.t[1] - T.t[1], S.t[2] - T.t[2]};
    for (int r = 0; r < 3; ++r) relt[r] = T.R[r] * dt[0] + T.R[3 + r] * dt[1] + T.R[6 + r] * dt[2];
    // K_t * relR
    double KR[9];
    for (int c = 0; c < 3; ++c) {
        KR[c] = T.f * relR[c] + T.cx * relR[6 + c];
        KR[3 + c] = T.f * relR[3 + c] + T.cy * relR[6 + c];
        KR[6 + c] = relR[6 + c];
    }
    // (K_t relR) * K_s^-1,  K_s^-1 = [[1/f,0,-cx/f],[0,1/f,-cy/f],[0,0,1]]
    const double inf = 1.0 / S.f;
    for (int r = 0; r < 3; ++r) {
        A[r * 3] = KR[r * 3] * inf;
        A[r * 3 + 1] = KR[r * 3 + 1] * inf;
        A[r * 3 + 2] = KR[r * 3 + 2] - (KR[r * 3] * S.cx + KR[r * 3 + 1] * S.cy) * inf;
    }
    b[0] = T.f * relt[0] + T.cx * relt[2];
    b[1] = T.f * relt[1] + T.cy * relt[2];
    b[2] = relt[2];
}

// ------------------------------------------------------------------------------------------
// derived per-iteration blocks (bodies of the prepare kernel)
// ------------------------------------------------------------------------------------------
A3R_HD void derive_frame_par(const FrameD& F, float* o) {
    for (int k = 0; k < 9; ++k) o[FP_R + k] = (float)F.R[k];
    for (int k = 0; k < 3; ++k) o[FP_T + k] = (float)F.t[k];
    o[FP_F] = (float)F.f; o[FP_CX] = (float)F.cx; o[FP_CY] = (float)F.cy; o[FP_INVF] = (float)(1.0 / F.f);
}

// pairwise pose block: M = s R diag(a), s t, s, a with s = exp(s~ + log_norm)
// (reference base_opt.py:195-200, 230-247); log_norm = log(base_scale) - mean_e s~ or 0
A3R_HD void derive_edge_par(const float* pv, const float* adaptor, double log_norm, bool norm_scale, double pw_break, float* o) {
    double qn[4], inv_norm, R[9], t[3];
    pose_to_Rt(pv, qn, &inv_norm, R, t);
    const double s = a3r_exp((double)pv[7] + log_norm);
    double ad[3] = {(double)adaptor[0], (double)adaptor[0], (double)adaptor[1]};
    if (norm_scale) {
        const double m = (ad[0] + ad[1] + ad[2]) / 3.0;
        ad[0] -= m; ad[1] -= m; ad[2] -= m;
    }
    for (int c = 0; c < 3; ++c) {
        const double a = a3r_exp(ad[c] / pw_break);
        o[EP_A + c] = (float)a;
        for (int r = 0; r < 3; ++r) o[EP_M + r * 3 + c] = (float)(s * R[r * 3 + c] * a);
    }
    for (int k = 0; k < 3; ++k) o[EP_ST + k] = (float)(s * t[k]);
    o[EP_S] = (float)s;
}

A3R_HD void derive_slot_par(const FrameD& S, const FrameD& T, float* o) {
    double A[9], b[3], relR[9], relt[3];
    flow_homography(S, T, A, b, relR, relt);
    for (int k = 0; k < 9; ++k) o[SP_A + k] = (float)A[k];
    for (int k = 0; k < 3; ++k) o[SP_B + k] = (float)b[k];
    o[12] = o[13] = o[14] = o[15] = 0.f;
}

// ------------------------------------------------------------------------------------------
// parameter-chain gradients (fp64): bodies of the small-parameter update kernel
// ------------------------------------------------------------------------------------------
// contribution of one relative-pose term to the two frames it couples: rows laid out like frame_grad
struct PairGrad {
    double s[FG_STRIDE];     // added to the source frame's row
    double t[FG_STRIDE];     // added to the target frame's row
};

// chain of relR = R_t^T R_s, relt = R_t^T (t_s - t_t)
A3R_HD void rel_chain(const FrameD& S, const FrameD& T, const double GrelR[9], const double Grelt[3], PairGrad& o) {
    double tmp[9];
    mat3_mul(T.R, GrelR, tmp);                 // dL/dR_s = R_t G_relR
    for (int k = 0; k < 9; ++k) o.s[FG_R + k] += tmp[k];
    mat3_mult(S.R, GrelR, tmp);                // dL/dR_t = R_s G_relR^T + (t_s - t_t) G_relt^T
    const double dt[3] = {S.t[0] - T.t[0], S.t[1] - T.t[1], S.t[2] - T.t[2]};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o.t[FG_R + r * 3 + c] += tmp[r * 3 + c] + dt[r] * Grelt[c];
    for (int r = 0; r < 3; ++r) {
        const double g = T.R[r * 3] * Grelt[0] + T.R[r * 3 + 1] * Grelt[1] + T.R[r * 3 + 2] * Grelt[2];
        o.s[FG_T + r] += g;
        o.t[FG_T + r] -= g;
    }
}

// per-frame gradient of the alignment term from the frame accumulators (sum h, sum h c^T; h = R^T gsum):
// z = R c + t  =>  dL/dR = R (sum h c^T), dL/dt = R sum h;  c = d ((u-cx)/f, (v-cy)/f, 1)  =>
// dL/df = -(H00 + H11)/f, dL/dcx = -H02/f, dL/dcy = -H12/f
A3R_HD void frame_align_grad(const FrameD& F, const double* a, double* fg) {
    const double* H = a + FACC_GC;
    for (int m = 0; m < 3; ++m) {
        for (int k = 0; k < 3; ++k) fg[FG_R + m * 3 + k] = F.R[m * 3] * H[k] + F.R[m * 3 + 1] * H[3 + k] + F.R[m * 3 + 2] * H[6 + k];
        fg[FG_T + m] = F.R[m * 3] * a[FACC_G] + F.R[m * 3 + 1] * a[FACC_G + 1] + F.R[m * 3 + 2] * a[FACC_G + 2];
    }
    fg[FG_F] = -(H[0] + H[4]) / F.f; fg[FG_CX] = -H[2] / F.f; fg[FG_CY] = -H[5] / F.f; fg[15] = 0;
}

// flow chain of one slot (src S -> tgt T): G_A, G_b (already scaled by flow_weight / count) ->
// gradients w.r.t. both cameras (SURVEY.md 8(a) "flow side")
A3R_HD void slot_flow_chain(const FrameD& S, const FrameD& T, const double* acc, double sc, PairGrad& o) {
    for (int k = 0; k < FG_STRIDE; ++k) o.s[k] = o.t[k] = 0.0;
    double A[9], b[3], relR[9], relt[3];
    flow_homography(S, T, A, b, relR, relt);
    double GA[9], Gb[3];
    for (int k = 0; k < 9; ++k) GA[k] = sc * acc[ACC_GA + k];
    for (int k = 0; k < 3; ++k) Gb[k] = sc * acc[ACC_GB + k];
    const double Kt[9] = {T.f, 0, T.cx, 0, T.f, T.cy, 0, 0, 1};
    const double Ksi[9] = {1 / S.f, 0, -S.cx / S.f, 0, 1 / S.f, -S.cy / S.f, 0, 0, 1};
    double KtG[9], GrelR[9], Grelt[3], Wm[9], GKt[9], KtR[9], GKsi[9];
    mat3_tmul(Kt, GA, KtG);                // K_t^T G_A
    mat3_mult(KtG, Ksi, GrelR);            // ... K_s^-T
    for (int r = 0; r < 3; ++r) Grelt[r] = Kt[r] * Gb[0] + Kt[3 + r] * Gb[1] + Kt[6 + r] * Gb[2];
    mat3_mul(relR, Ksi, Wm);               // W = relR K_s^-1
    mat3_mult(GA, Wm, GKt);                // dL/dK_t = G_A W^T + G_b relt^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) GKt[r * 3 + c] += Gb[r] * relt[c];
    mat3_mul(Kt, relR, KtR);
    mat3_tmul(KtR, GA, GKsi);              // dL/dK_s^-1 = (K_t relR)^T G_A
    o.t[FG_F] += GKt[0] + GKt[4];
    o.t[FG_CX] += GKt[2];
    o.t[FG_CY] += GKt[5];
    const double if2 = 1.0 / (S.f * S.f);
    o.s[FG_F] += (-GKsi[0] - GKsi[4] + S.cx * GKsi[2] + S.cy * GKsi[5]) * if2;
    o.s[FG_CX] += -GKsi[2] / S.f;
    o.s[FG_CY] += -GKsi[5] / S.f;
    rel_chain(S, T, GrelR, Grelt, o);
}

// temporal smoothness between frame T = n and S = n+1 (reference optimizer.py:559-572):
// returns the un-weighted loss, fills the weighted gradient contributions
A3R_HD double temporal_pair(const FrameD& S, const FrameD& T, double tsw, double tw, PairGrad* o) {
    double A[9], b[3], relR[9], relt[3];
    flow_homography(S, T, A, b, relR, relt);
    double D[9];
    for (int k = 0; k < 9; ++k) D[k] = relR[k];
    D[0] -= 1; D[4] -= 1; D[8] -= 1;
    double nr = 0;
    for (int k = 0; k < 9; ++k) nr += D[k] * D[k];
    nr = sqrt(nr);
    const double ntn = sqrt(relt[0] * relt[0] + relt[1] * relt[1] + relt[2] * relt[2]);
    if (o) {
        for (int k = 0; k < FG_STRIDE; ++k) o->s[k] = o->t[k] = 0.0;
        double GrelR[9], Grelt[3];
        for (int k = 0; k < 9; ++k) GrelR[k] = nr > 0 ? tsw * D[k] / nr : 0.0;
        for (int k = 0; k < 3; ++k) Grelt[k] = ntn > 0 ? tsw * tw * relt[k] / ntn : 0.0;
        rel_chain(S, T, GrelR, Grelt, *o);
    }
    return nr + tw * ntn;
}

// d L / d log s_e = - sum g . y  from the two slots of edge e  (y = M X + s t)
A3R_HD double edge_logscale_grad(const float* ep, const double* a0, const double* a1) {
    double g = 0;
    for (int k = 0; k < 9; ++k) g += (double)ep[EP_M + k] * (a0[ACC_SGX + k] + a1[ACC_SGX + k]);
    for (int k = 0; k < 3; ++k) g += (double)ep[EP_ST + k] * (a0[ACC_SG + k] + a1[ACC_SG + k]);
    return -g;
}

// gradient of one pairwise pose vector (8) and its adaptor pair (2)
A3R_HD void edge_pose_grad(const float* pv, const float* ep, const double* a0, const double* a1, double g_logscale,
                           bool train_adaptors, bool norm_scale, double pw_break, float g[8], float ga[2]) {
    double qn[4], inv_norm, R[9], t[3];
    pose_to_Rt(pv, qn, &inv_norm, R, t);
    const double s = (double)ep[EP_S];
    double GR[9], gq[4];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) GR[r * 3 + c] = -s * (a0[ACC_SGX + r * 3 + c] + a1[ACC_SGX + r * 3 + c]) * (double)ep[EP_A + c];
    rot_grad_to_quat(GR, qn, inv_norm, gq);
    for (int k = 0; k < 4; ++k) g[k] = (float)gq[k];
    for (int k = 0; k < 3; ++k) g[4 + k] = (float)(-s * (a0[ACC_SG + k] + a1[ACC_SG + k]) * signed_expm1_grad(pv[4 + k]));
    g[7] = (float)g_logscale;
    ga[0] = ga[1] = 0.f;
    if (train_adaptors) {                           // reference base_opt.py:195-200
        double h[3];
        for (int c = 0; c < 3; ++c) {
            double d = 0;
            for (int r = 0; r < 3; ++r) d += R[r * 3 + c] * (a0[ACC_SGX + r * 3 + c] + a1[ACC_SGX + r * 3 + c]);
            h[c] = -s * d * (double)ep[EP_A + c] / pw_break;        // dL/da_c * da_c/d(adapt_c)
        }
        if (norm_scale) {
            const double m = (h[0] + h[1] + h[2]) / 3.0;
            h[0] -= m; h[1] -= m; h[2] -= m;
        }
        ga[0] = (float)(h[0] + h[1]);
        ga[1] = (float)h[2];
    }
}

// gradient of one image pose vector (7), its log-focal and principal point from its frame_grad row;
// ex[k] = signed_expm1_grad(pv[4 + k]) (the caller may have evaluated the three exponentials ahead of time)
A3R_HD void frame_pose_grad_pre(const FrameD& F, const double* fg, double focal_break, bool train_pp, const double ex[3],
                                float g[7], double* gphi, float gpp[2]) {
    double gq[4];
    rot_grad_to_quat(fg + FG_R, F.qn, F.inv_norm, gq);
    for (int k = 0; k < 4; ++k) g[k] = (float)gq[k];
    for (int k = 0; k < 3; ++k) g[4 + k] = (float)(fg[FG_T + k] * ex[k]);
    *gphi = fg[FG_F] * F.f / focal_break;                       // f = exp(phi / focal_break)
    gpp[0] = train_pp ? (float)(10.0 * fg[FG_CX]) : 0.f;         // pp = (W/2,H/2) + 10 im_pp
    gpp[1] = train_pp ? (float)(10.0 * fg[FG_CY]) : 0.f;
}
A3R_HD void frame_pose_grad(const float* pv, const FrameD& F, const double* fg, double focal_break, bool train_pp,
                            float g[7], double* gphi, float gpp[2]) {
    const double ex[3] = {signed_expm1_grad(pv[4]), signed_expm1_grad(pv[5]), signed_expm1_grad(pv[6])};
    frame_pose_grad_pre(F, fg, focal_break, train_pp, ex, g, gphi, gpp);
}

// loss assembly + gate (reference optimizer.py:536-557)
struct LossScalars { double loss, floss, dropped, sc_i, sc_j; };
A3R_HD LossScalars assemble_loss(double l_align, double l_temp, double num_i, double cnt_i, double num_j, double cnt_j,
                                 bool flow_active, double flow_weight, double flow_thre, double tsw) {
    LossScalars r;
    r.floss = nan(""); r.dropped = 0; r.sc_i = 0; r.sc_j = 0;
    if (flow_active) {
        r.floss = num_i / cnt_i + num_j / cnt_j;
        if (r.floss > flow_thre && flow_thre > 0) r.dropped = 1;
        else { r.sc_i = flow_weight / cnt_i; r.sc_j = flow_weight / cnt_j; }
    }
    r.loss = l_align;
    if (tsw > 0) r.loss += tsw * l_temp;
    if (flow_active && r.dropped == 0) r.loss += flow_weight * r.floss;
    return r;
}

// ------------------------------------------------------------------------------------------
// per-pixel / per-pixel-slot math of the fused forward+backward (fp32)
// ------------------------------------------------------------------------------------------
struct PixState {
    float zx, zy, zz;     // world point z = R c + t
    float cx, cy, cz;     // camera point c (cz = depth)
    float u, v;           // pixel coordinates as floats
    float disp;           // 1 / (d + 1e-6)            (reference optimizer.py:530)
    float dfac;           // d disp / d logdepth = -d / (d + 1e-6)^2
};

A3R_HD void pixel_state(const float* fp, float logd, int u, int v, PixState& s) {
    const float d = expf(logd);                                        // optimizer.py:462-466
    s.u = (float)u;
    s.v = (float)v;
    // optimizer.py:574-581 computes d (u - cx) / f; here the division is a multiplication by the fp64-rounded 1/f
    // of the prepare kernel (<= 1 ulp apart), and 1/(d + 1e-6) is a single-MUFU reciprocal on the device
    s.cx = d * (s.u - fp[FP_CX]) * fp[FP_INVF];
    s.cy = d * (s.v - fp[FP_CY]) * fp[FP_INVF];
    s.cz = d;
    s.zx = fp[FP_R + 0] * s.cx + fp[FP_R + 1] * s.cy + fp[FP_R + 2] * s.cz + fp[FP_T + 0];
    s.zy = fp[FP_R + 3] * s.cx + fp[FP_R + 4] * s.cy + fp[FP_R + 5] * s.cz + fp[FP_T + 1];
    s.zz = fp[FP_R + 6] * s.cx + fp[FP_R + 7] * s.cy + fp[FP_R + 8] * s.cz + fp[FP_T + 2];
    s.disp = fast_rcp(d + 1e-6f);
    s.dfac = -d * s.disp * s.disp;
}

// per-pixel epilogue of the scan: h = R^T gsum; d L_align / d logdepth = gsum . (R c) = h . c; frame sums
A3R_HD float pixel_epilogue(const float* fp, const PixState& s, float gsx, float gsy, float gsz, float* facc) {
    const float hx = fp[FP_R + 0] * gsx + fp[FP_R + 3] * gsy + fp[FP_R + 6] * gsz;
    const float hy = fp[FP_R + 1] * gsx + fp[FP_R + 4] * gsy + fp[FP_R + 7] * gsz;
    const float hz = fp[FP_R + 2] * gsx + fp[FP_R + 5] * gsy + fp[FP_R + 8] * gsz;
    facc[FACC_G + 0] += hx; facc[FACC_G + 1] += hy; facc[FACC_G + 2] += hz;
    facc[FACC_GC + 0] = fmaf(hx, s.cx, facc[FACC_GC + 0]); facc[FACC_GC + 1] = fmaf(hx, s.cy, facc[FACC_GC + 1]); facc[FACC_GC + 2] = fmaf(hx, s.cz, facc[FACC_GC + 2]);
    facc[FACC_GC + 3] = fmaf(hy, s.cx, facc[FACC_GC + 3]); facc[FACC_GC + 4] = fmaf(hy, s.cy, facc[FACC_GC + 4]); facc[FACC_GC + 5] = fmaf(hy, s.cz, facc[FACC_GC + 5]);
    facc[FACC_GC + 6] = fmaf(hz, s.cx, facc[FACC_GC + 6]); facc[FACC_GC + 7] = fmaf(hz, s.cy, facc[FACC_GC + 7]); facc[FACC_GC + 8] = fmaf(hz, s.cz, facc[FACC_GC + 8]);
    return hx * s.cx + hy * s.cy + hz * s.cz;
}

// constants of one slot as the accumulate kernel wants them
struct SlotConst {
    float nM[9];     // -(s R diag(a))
    float st[3];     // s t
    float A[9];      // flow homography
    float b[3];
    float wn;        // 1 / total_area of this side's class
};

// alignment residual of one pixel-slot: updates acc[ACC_SG..ACC_LOSS], returns the pixel-slot's g = dL/dz
// FIRST: the chunk's first pixel - the sums START with its terms (acc is not read: no adds onto freshly zeroed registers)
template <bool L2DIST, bool FIRST = false>
A3R_HD void align_px(const SlotConst& c, float X0, float X1, float X2, float w, const PixState& s,
                     float* acc, float& gx, float& gy, float& gz) {
    const float w2 = w * c.wn;
    float rx = s.zx - c.st[0], ry = s.zy - c.st[1], rz = s.zz - c.st[2];
    rx = fmaf(c.nM[0], X0, fmaf(c.nM[1], X1, fmaf(c.nM[2], X2, rx)));
    ry = fmaf(c.nM[3], X0, fmaf(c.nM[4], X1, fmaf(c.nM[5], X2, ry)));
    rz = fmaf(c.nM[6], X0, fmaf(c.nM[7], X1, fmaf(c.nM[8], X2, rz)));
    // 'l1': the 1e-30 keeps rsqrt finite at rho = 0 (zero (sub)gradient there, as torch.norm) without a clamp of its own;
    // it is below half an ulp of r2 unless rz^2 < 1e-23
    const float r2 = fmaf(rx, rx, fmaf(ry, ry, L2DIST ? rz * rz : fmaf(rz, rz, 1e-30f)));
    float k;
    if (L2DIST) {                     // reference commons.py:67-68
        acc[ACC_LOSS] = FIRST ? w2 * r2 : fmaf(w2, r2, acc[ACC_LOSS]);
        k = 2.0f * w2;
    } else {                          // reference commons.py:71-72 ('l1' = un-squared L2 norm)
        const float inv = fast_rsqrt(r2);
        k = w2 * inv;
        acc[ACC_LOSS] = FIRST ? k * r2 : fmaf(k, r2, acc[ACC_LOSS]);     // w rho = (w / rho) rho^2
    }
    gx = k * rx; gy = k * ry; gz = k * rz;
    if (FIRST) {
        acc[ACC_SG + 0] = gx; acc[ACC_SG + 1] = gy; acc[ACC_SG + 2] = gz;
        acc[ACC_SGX + 0] = gx * X0; acc[ACC_SGX + 1] = gx * X1; acc[ACC_SGX + 2] = gx * X2;
        acc[ACC_SGX + 3] = gy * X0; acc[ACC_SGX + 4] = gy * X1; acc[ACC_SGX + 5] = gy * X2;
        acc[ACC_SGX + 6] = gz * X0; acc[ACC_SGX + 7] = gz * X1; acc[ACC_SGX + 8] = gz * X2;
        return;
    }
    acc[ACC_SG + 0] += gx; acc[ACC_SG + 1] += gy; acc[ACC_SG + 2] += gz;
    acc[ACC_SGX + 0] = fmaf(gx, X0, acc[ACC_SGX + 0]); acc[ACC_SGX + 1] = fmaf(gx, X1, acc[ACC_SGX + 1]); acc[ACC_SGX + 2] = fmaf(gx, X2, acc[ACC_SGX + 2]);
    acc[ACC_SGX + 3] = fmaf(gy, X0, acc[ACC_SGX + 3]); acc[ACC_SGX + 4] = fmaf(gy, X1, acc[ACC_SGX + 4]); acc[ACC_SGX + 5] = fmaf(gy, X2, acc[ACC_SGX + 5]);
    acc[ACC_SGX + 6] = fmaf(gz, X0, acc[ACC_SGX + 6]); acc[ACC_SGX + 7] = fmaf(gz, X1, acc[ACC_SGX + 7]); acc[ACC_SGX + 8] = fmaf(gz, X2, acc[ACC_SGX + 8]);
}

// flow-projection residual of one pixel-slot: updates acc[ACC_GA..ACC_FCNT], returns g_q . b (the un-scaled
// d loss / d disp).  Reference optimizer.py:18-24,531-537, goem_opt.py:221-236.
//   stat      false for dynamic / out-of-frame pixels: they are evaluated but contribute exactly nothing
//             (branch-free: the mask enters through the per-pixel-mask select)
//   thre_eff  pxl_thre, or +inf when pxl_thre <= 0 (per_pixel_mask = mask, optimizer.py:20-23)
A3R_HD float flow_thre_eff(float pxl_thre) { return pxl_thre > 0.0f ? pxl_thre : INFINITY; }

// Returns gf0 + g_q . b (gf0: the running d loss / d disp of this pixel, so that the dot product folds into FMAs).
// BASE: index of acc[0] in the accumulator row (0: a whole row; ACC_GA: an array that holds only the 14 flow sums).
template <int BASE = 0>
A3R_HD float flow_px(const SlotConst& c, const PixState& s, float fu, float fv, float thre_eff, bool stat, float* acc,
                     float gf0 = 0.0f) {
    const float qx = fmaf(c.A[0], s.u, fmaf(c.A[1], s.v, fmaf(s.disp, c.b[0], c.A[2])));
    const float qy = fmaf(c.A[3], s.u, fmaf(c.A[4], s.v, fmaf(s.disp, c.b[1], c.A[5])));
    const float qz = fmaf(c.A[6], s.u, fmaf(c.A[7], s.v, fmaf(s.disp, c.b[2], c.A[8])));
    const float inv = fast_rcp(qz + 1e-6f);
    const float up = qx * inv, vp = qy * inv;
    const float eu = (up - s.u) - fu, ev = (vp - s.v) - fv;
    // smooth L1, beta = 1, in min form: m = min(|e|, 1); loss = m (|e| - m/2); d loss / d e = sign(e) m
    const float au = fabsf(eu), av = fabsf(ev);
    const float su = clamp_unit_signed(eu), sv = clamp_unit_signed(ev);
    const float mu = fabsf(su), mv = fabsf(sv);
    const float ru = mu * fmaf(-0.5f, mu, au);
    const float rv = mv * fmaf(-0.5f, mv, av);
    const bool pu = stat && ru < thre_eff;                            // per_pixel_mask (NaN residuals are out)
    const bool pv = stat && rv < thre_eff;
    if (pu) { acc[ACC_FNUM - BASE] += ru; acc[ACC_FCNT - BASE] += 1.0f; }
    if (pv) { acc[ACC_FNUM - BASE] += rv; acc[ACC_FCNT - BASE] += 1.0f; }
    const float geu = pu ? su : 0.0f;
    const float gev = pv ? sv : 0.0f;
    const float gqx = geu * inv, gqy = gev * inv;
    const float gqz = -fmaf(gqx, up, gqy * vp);
    acc[ACC_GA - BASE + 0] = fmaf(gqx, s.u, acc[ACC_GA - BASE + 0]); acc[ACC_GA - BASE + 1] = fmaf(gqx, s.v, acc[ACC_GA - BASE + 1]); acc[ACC_GA - BASE + 2] += gqx;
    acc[ACC_GA - BASE + 3] = fmaf(gqy, s.u, acc[ACC_GA - BASE + 3]); acc[ACC_GA - BASE + 4] = fmaf(gqy, s.v, acc[ACC_GA - BASE + 4]); acc[ACC_GA - BASE + 5] += gqy;
    acc[ACC_GA - BASE + 6] = fmaf(gqz, s.u, acc[ACC_GA - BASE + 6]); acc[ACC_GA - BASE + 7] = fmaf(gqz, s.v, acc[ACC_GA - BASE + 7]); acc[ACC_GA - BASE + 8] += gqz;
    acc[ACC_GB - BASE + 0] = fmaf(s.disp, gqx, acc[ACC_GB - BASE + 0]);
    acc[ACC_GB - BASE + 1] = fmaf(s.disp, gqy, acc[ACC_GB - BASE + 1]);
    acc[ACC_GB - BASE + 2] = fmaf(s.disp, gqz, acc[ACC_GB - BASE + 2]);
    return fmaf(gqx, c.b[0], fmaf(gqy, c.b[1], fmaf(gqz, c.b[2], gf0)));
}

// The same pixel-slot as the product scan (ga_scan_v4.cuh) evaluates it.  Differences to flow_px, all of them in what is
// folded where, none in the function computed:
//   * c.A[8] already carries the + 1e-6 of the projection's denominator (added when the slot table is staged);
//   * the static mask comes as the pixel's own threshold thre_pix = static ? thre_eff : -1 (residuals are >= 0);
//   * ROWT (every tile lies in one image row: v is the same for all pixels of a (tile, slot) chunk): c.A[2], c.A[5], c.A[8]
//     hold the row constants A[3r+1] v + A[3r+2], and the v-column of sum g_q (u,v,1)^T is NOT accumulated here - the
//     caller sets it to v * sum g_q once per chunk (flow_row_finish).
// FIRST: as in align_px - the chunk's first pixel starts the sums.
template <int BASE, bool ROWT, bool FIRST = false>
A3R_HD float flow_px_scan(const SlotConst& c, float u, float v, float disp, float fu, float fv, float thre_pix, float* acc,
                          float gf0) {
    float qx, qy, qz;
    if (ROWT) {
        qx = fmaf(c.A[0], u, fmaf(disp, c.b[0], c.A[2]));
        qy = fmaf(c.A[3], u, fmaf(disp, c.b[1], c.A[5]));
        qz = fmaf(c.A[6], u, fmaf(disp, c.b[2], c.A[8]));
    } else {
        qx = fmaf(c.A[0], u, fmaf(c.A[1], v, fmaf(disp, c.b[0], c.A[2])));
        qy = fmaf(c.A[3], u, fmaf(c.A[4], v, fmaf(disp, c.b[1], c.A[5])));
        qz = fmaf(c.A[6], u, fmaf(c.A[7], v, fmaf(disp, c.b[2], c.A[8])));
    }
    const float inv = fast_rcp(qz);
    const float up = qx * inv, vp = qy * inv;
    const float eu = (up - u) - fu, ev = (vp - v) - fv;
    const float au = fabsf(eu), av = fabsf(ev);
    const float su = clamp_unit_signed(eu), sv = clamp_unit_signed(ev);
    const float mu = fabsf(su), mv = fabsf(sv);
    const float ru = mu * fmaf(-0.5f, mu, au);
    const float rv = mv * fmaf(-0.5f, mv, av);
    const bool pu = ru < thre_pix;                                    // per_pixel_mask (NaN residuals are out)
    const bool pv = rv < thre_pix;
    if (FIRST) { acc[ACC_FNUM - BASE] = pu ? ru : 0.0f; acc[ACC_FCNT - BASE] = pu ? 1.0f : 0.0f; }
    else if (pu) { acc[ACC_FNUM - BASE] += ru; acc[ACC_FCNT - BASE] += 1.0f; }
    if (pv) { acc[ACC_FNUM - BASE] += rv; acc[ACC_FCNT - BASE] += 1.0f; }
    const float geu = pu ? su : 0.0f;
    const float gev = pv ? sv : 0.0f;
    const float gqx = geu * inv, gqy = gev * inv;
    const float gqz = -fmaf(gqx, up, gqy * vp);
    if (FIRST) {
        acc[ACC_GA - BASE + 0] = gqx * u; acc[ACC_GA - BASE + 2] = gqx;
        acc[ACC_GA - BASE + 3] = gqy * u; acc[ACC_GA - BASE + 5] = gqy;
        acc[ACC_GA - BASE + 6] = gqz * u; acc[ACC_GA - BASE + 8] = gqz;
        if (!ROWT) { acc[ACC_GA - BASE + 1] = gqx * v; acc[ACC_GA - BASE + 4] = gqy * v; acc[ACC_GA - BASE + 7] = gqz * v; }
        acc[ACC_GB - BASE + 0] = disp * gqx; acc[ACC_GB - BASE + 1] = disp * gqy; acc[ACC_GB - BASE + 2] = disp * gqz;
        return fmaf(gqx, c.b[0], fmaf(gqy, c.b[1], fmaf(gqz, c.b[2], gf0)));
    }
    acc[ACC_GA - BASE + 0] = fmaf(gqx, u, acc[ACC_GA - BASE + 0]); acc[ACC_GA - BASE + 2] += gqx;
    acc[ACC_GA - BASE + 3] = fmaf(gqy, u, acc[ACC_GA - BASE + 3]); acc[ACC_GA - BASE + 5] += gqy;
    acc[ACC_GA - BASE + 6] = fmaf(gqz, u, acc[ACC_GA - BASE + 6]); acc[ACC_GA - BASE + 8] += gqz;
    if (!ROWT) {
        acc[ACC_GA - BASE + 1] = fmaf(gqx, v, acc[ACC_GA - BASE + 1]);
        acc[ACC_GA - BASE + 4] = fmaf(gqy, v, acc[ACC_GA - BASE + 4]);
        acc[ACC_GA - BASE + 7] = fmaf(gqz, v, acc[ACC_GA - BASE + 7]);
    }
    acc[ACC_GB - BASE + 0] = fmaf(disp, gqx, acc[ACC_GB - BASE + 0]);
    acc[ACC_GB - BASE + 1] = fmaf(disp, gqy, acc[ACC_GB - BASE + 1]);
    acc[ACC_GB - BASE + 2] = fmaf(disp, gqz, acc[ACC_GB - BASE + 2]);
    return fmaf(gqx, c.b[0], fmaf(gqy, c.b[1], fmaf(gqz, c.b[2], gf0)));
}
// ROWT: the row constants of a chunk, and the v-column once the chunk's pixels are in
A3R_HD void flow_row_begin(SlotConst& c, float v) {
    c.A[2] = fmaf(c.A[1], v, c.A[2]); c.A[5] = fmaf(c.A[4], v, c.A[5]); c.A[8] = fmaf(c.A[7], v, c.A[8]);
}
template <int BASE>
A3R_HD void flow_row_finish(float v, float* acc) {
    acc[ACC_GA - BASE + 1] = v * acc[ACC_GA - BASE + 2];
    acc[ACC_GA - BASE + 4] = v * acc[ACC_GA - BASE + 5];
    acc[ACC_GA - BASE + 7] = v * acc[ACC_GA - BASE + 8];
}

// ------------------------------------------------------------------------------------------
// depth prior (reference optimizer.py:546-552, goem_opt.py:15-36): scale-invariant deviation of the log-depth from its
// initial value, pixel weight 1 + dynamic_mask.  Per frame with d_p = l_p - l0_p (both clamped at log 1e-6),
// s = -mean d:  loss_n = (1/P) sum w (d + s)^2 = (S_wdd + 2 s S_wd + s^2 S_w) / P;  the total is the mean over frames.
// d loss / d l_q = (2 / (N P)) [ w_q (d_q + s) - (S_wd + s S_w) / P ]   (0 where l_q sits on the clamp)
// ------------------------------------------------------------------------------------------
enum : int { PR_D = 0, PR_W = 1, PR_WD = 2, PR_WDD = 3, PR_N = 4 };
#define A3R_LOG_EPS (-13.815510557964274f)      /* log(1e-6): goem_opt.py clamps both depths at eps = 1e-6 */

A3R_HD float prior_delta(float logd, float logd_init) { return fmaxf(logd, A3R_LOG_EPS) - logd_init; }

struct PriorFrame { double s, mean_b, loss; };
A3R_HD PriorFrame prior_frame(const double* a, double P) {
    PriorFrame r;
    r.s = -a[PR_D] / P;
    r.mean_b = (a[PR_WD] + r.s * a[PR_W]) / P;
    r.loss = (a[PR_WDD] + 2.0 * r.s * a[PR_WD] + r.s * r.s * a[PR_W]) / P;
    return r;
}
A3R_HD float prior_grad(float logd, float logd_init, float w, float s, float mean_b, float coef) {
    return logd > A3R_LOG_EPS ? coef * (w * (prior_delta(logd, logd_init) + s) - mean_b) : 0.0f;
}

// ------------------------------------------------------------------------------------------
// torch.optim.Adam, single-tensor math (torch/optim/adam.py:457-547), fp32
// ------------------------------------------------------------------------------------------
A3R_HD void adam_update(float& p, float& m, float& v, float g, float beta1, float beta2, float eps,
                        float step_size, float bc2_sqrt) {
    m = m + (1.0f - beta1) * (g - m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (1.0f - beta2) * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);                  // param.addcdiv_(exp_avg, denom, value=-step_size)
}

}  // namespace a3r
