/*
 * TEST INFRASTRUCTURE ONLY -- the product path (cpi_b200/) never loads, links or calls this file.
 *
 * Plain-C, dense, literal CPU restatement of the hot path of rpng/cpi (paths relative to
 * /root/reference/cpi_compare/src).  Every function cites the reference lines it follows.  It deliberately does
 * NOT use the block-sparse formulation of the CUDA kernels: the 15x15 / 21x21 products are full dense loops, the
 * way Eigen evaluates them, so that it is an independent check of the kernels' algebra.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (a) oracle/_ref/libcpi_ref.so = the
 * unmodified reference compiled in place (oracle/ref_shim.cpp) when that library is present, and (b) the committed
 * golden vectors under tests/golden/ that were generated from that library by tests/golden/make_golden.py.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, to stay close to the reference's SSE2 code).
 */
#include "cpi_oracle.h"
#include "../include/cpi_b200.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small dense helpers (column-major) ---------- */
#define E(M, ld, i, j) ((M)[(i) + (j) * (ld)])

static void mat_zero(double* A, int n) { memset(A, 0, sizeof(double) * (size_t)n); }
static void mat_eye(double* A, int n) { mat_zero(A, n * n); for (int i = 0; i < n; i++) E(A, n, i, i) = 1.0; }
/* C(m x n) = A(m x k) * B(k x n) */
static void gemm(int m, int n, int k, const double* A, const double* B, double* C) {
    for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int l = 0; l < k; l++) s += E(A, m, i, l) * E(B, k, l, j);
        E(C, m, i, j) = s;
    }
}
/* C(m x n) = A(m x k) * B(n x k)^T */
static void gemm_nt(int m, int n, int k, const double* A, const double* B, double* C) {
    for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int l = 0; l < k; l++) s += E(A, m, i, l) * E(B, n, j, l);
        E(C, m, i, j) = s;
    }
}
static void transpose(int m, int n, const double* A, double* At) {
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) E(At, n, j, i) = E(A, m, i, j);
}
/* dst block (r0,c0) of size br x bc in a matrix with leading dim ld  <-  s * B */
static void set_block(double* M, int ld, int r0, int c0, int br, int bc, const double* B, double s) {
    for (int j = 0; j < bc; j++) for (int i = 0; i < br; i++) E(M, ld, r0 + i, c0 + j) = s * E(B, br, i, j);
}
static void get_block(const double* M, int ld, int r0, int c0, int br, int bc, double* B) {
    for (int j = 0; j < bc; j++) for (int i = 0; i < br; i++) E(B, br, i, j) = E(M, ld, r0 + i, c0 + j);
}
static void m3_mul(const double* A, const double* B, double* C) { double T[9]; gemm(3, 3, 3, A, B, T); memcpy(C, T, sizeof T); }
static void m3_vec(const double* A, const double* v, double* o) {
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = E(A, 3, i, 0) * v[0] + E(A, 3, i, 1) * v[1] + E(A, 3, i, 2) * v[2];
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static double norm3(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static double norm4(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]); }

/* ------------------------------------------------------------------ utils/quat_ops.h ----------------------------- */

/* skew_x, quat_ops.h:92-98 */
static void skew_x(const double* w, double* M) {
    E(M, 3, 0, 0) = 0;     E(M, 3, 0, 1) = -w[2]; E(M, 3, 0, 2) = w[1];
    E(M, 3, 1, 0) = w[2];  E(M, 3, 1, 1) = 0;     E(M, 3, 1, 2) = -w[0];
    E(M, 3, 2, 0) = -w[1]; E(M, 3, 2, 1) = w[0];  E(M, 3, 2, 2) = 0;
}

/* rot_2_quat, quat_ops.h:45-86 : four-case largest-diagonal extraction, force w >= 0, normalise */
void oracle_rot_2_quat(const double* rot, double* q) {
    double r00 = E(rot, 3, 0, 0), r11 = E(rot, 3, 1, 1), r22 = E(rot, 3, 2, 2);
    double T = r00 + r11 + r22;
    if (r00 >= T && r00 >= r11 && r00 >= r22) {
        q[0] = sqrt((1 + (2 * r00) - T) / 4);
        q[1] = (1 / (4 * q[0])) * (E(rot, 3, 0, 1) + E(rot, 3, 1, 0));
        q[2] = (1 / (4 * q[0])) * (E(rot, 3, 0, 2) + E(rot, 3, 2, 0));
        q[3] = (1 / (4 * q[0])) * (E(rot, 3, 1, 2) - E(rot, 3, 2, 1));
    } else if (r11 >= T && r11 >= r00 && r11 >= r22) {
        q[1] = sqrt((1 + (2 * r11) - T) / 4);
        q[0] = (1 / (4 * q[1])) * (E(rot, 3, 0, 1) + E(rot, 3, 1, 0));
        q[2] = (1 / (4 * q[1])) * (E(rot, 3, 1, 2) + E(rot, 3, 2, 1));
        q[3] = (1 / (4 * q[1])) * (E(rot, 3, 2, 0) - E(rot, 3, 0, 2));
    } else if (r22 >= T && r22 >= r00 && r22 >= r11) {
        q[2] = sqrt((1 + (2 * r22) - T) / 4);
        q[0] = (1 / (4 * q[2])) * (E(rot, 3, 0, 2) + E(rot, 3, 2, 0));
        q[1] = (1 / (4 * q[2])) * (E(rot, 3, 1, 2) + E(rot, 3, 2, 1));
        q[3] = (1 / (4 * q[2])) * (E(rot, 3, 0, 1) - E(rot, 3, 1, 0));
    } else {
        q[3] = sqrt((1 + T) / 4);
        q[0] = (1 / (4 * q[3])) * (E(rot, 3, 1, 2) - E(rot, 3, 2, 1));
        q[1] = (1 / (4 * q[3])) * (E(rot, 3, 2, 0) - E(rot, 3, 0, 2));
        q[2] = (1 / (4 * q[3])) * (E(rot, 3, 0, 1) - E(rot, 3, 1, 0));
    }
    if (q[3] < 0) { for (int i = 0; i < 4; i++) q[i] = -q[i]; }
    double n = norm4(q);
    for (int i = 0; i < 4; i++) q[i] = q[i] / n;
}

/* quat_2_Rot, quat_ops.h:104-109 : (2 q4^2 - 1) I - 2 q4 [q_v x] + 2 q_v q_v^T */
void oracle_quat_2_Rot(const double* q, double* R) {
    double qx[9]; skew_x(q, qx);
    double s = 2 * pow(q[3], 2) - 1;
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++)
        E(R, 3, i, j) = s * (i == j ? 1.0 : 0.0) - 2 * q[3] * E(qx, 3, i, j) + 2 * q[i] * q[j];
}

/* quat_multiply, quat_ops.h:115-128 : JPL product q (x) p, force w >= 0, normalise */
void oracle_quat_multiply(const double* q, const double* p, double* out) {
    double Qm[16], qx[9], t[4];
    skew_x(q, qx);
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) E(Qm, 4, i, j) = q[3] * (i == j ? 1.0 : 0.0) - E(qx, 3, i, j);
    for (int i = 0; i < 3; i++) { E(Qm, 4, i, 3) = q[i]; E(Qm, 4, 3, i) = -q[i]; }
    E(Qm, 4, 3, 3) = q[3];
    gemm(4, 1, 4, Qm, p, t);
    if (t[3] < 0) { for (int i = 0; i < 4; i++) t[i] *= -1; }
    double n = norm4(t);
    for (int i = 0; i < 4; i++) out[i] = t[i] / n;
}

/* Exp, quat_ops.h:145-162 : I + sin(th)/th [w x] + (1-cos th)/th^2 [w x]^2 ; identity iff th == 0 */
void oracle_Exp(const double* w, double* R) {
    double wx[9], wx2[9];
    skew_x(w, wx);
    double th = norm3(w);
    if (th == 0) { mat_eye(R, 3); return; }
    gemm(3, 3, 3, wx, wx, wx2);
    double a = sin(th) / th, b = (1 - cos(th)) / pow(th, 2);
    for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0 ? 1.0 : 0.0) + a * wx[k] + b * wx2[k];
}

/* Inv, quat_ops.h:190-197 */
static void quat_inv(const double* q, double* o) { o[0] = -q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = q[3]; }

/* ------------------------------------------------------------------ cpi/CpiBase.h -------------------------------- */

typedef struct {
    int imu_avg, stj;                          /* CpiBase.h:95, CpiV2.h:58 */
    double DT, alpha[3], beta[3], q[4], R[9];  /* CpiBase.h:99-103 */
    double J_q[9], J_a[9], J_b[9], H_a[9], H_b[9];  /* :106-110 */
    double O_a[9], O_b[9];                     /* CpiV2.h:62-63 */
    double b_w[3], b_a[3], q_lin[4], grav[3];  /* CpiBase.h:113-118 */
    double Qc[144];                            /* :121 */
    double P[225];                             /* :124 */
    double Pbig[441], D[441];                  /* CpiV2.h:46-49 */
} cpi_t;

/* ctor CpiBase.h:52-66 + setLinearizationPoints :73-80 + member initialisers :99-124 */
static void cpi_init(cpi_t* c, const double* sig, const double* lin, int imu_avg, int stj) {
    memset(c, 0, sizeof *c);
    c->imu_avg = imu_avg; c->stj = stj;
    for (int b = 0; b < 4; b++) for (int i = 0; i < 3; i++) E(c->Qc, 12, 3 * b + i, 3 * b + i) = pow(sig[b], 2);
    mat_eye(c->R, 3);
    c->q[3] = 1.0;                 /* uninitialised in the reference; see include/cpi_b200.h */
    mat_eye(c->D, 21);             /* CpiV2.h:49 */
    memcpy(c->b_w, lin, 24); memcpy(c->b_a, lin + 3, 24); memcpy(c->q_lin, lin + 6, 32); memcpy(c->grav, lin + 10, 24);
}

/* Closed-form scalars shared by both models.  CpiV1.h:132-142 & 196-238  ==  CpiV2.h:158-168 & 231-274 */
typedef struct { double f1, f2, f3, f4, df1, df2, df3, df4; } coef_t;
static void coefficients(int small_w, double dt, double mag_w, double w_dt, double s, double c, coef_t* k) {
    if (small_w) {
        k->f1 = -(pow(dt, 3) / 3); k->f2 = (pow(dt, 4) / 8); k->f3 = -(pow(dt, 2) / 2); k->f4 = (pow(dt, 3) / 6);
        k->df1 = -(pow(dt, 5) / 15); k->df2 = (pow(dt, 6) / 72); k->df3 = -(pow(dt, 4) / 12); k->df4 = (pow(dt, 5) / 60);
    } else {
        k->f1 = (w_dt * c - s) / (pow(mag_w, 3));
        k->f2 = (pow(w_dt, 2) - 2 * c - 2 * w_dt * s + 2) / (2 * pow(mag_w, 4));
        k->f3 = -(1 - c) / pow(mag_w, 2);
        k->f4 = (w_dt - s) / pow(mag_w, 3);
        k->df1 = (pow(w_dt, 2) * s - 3 * s + 3 * w_dt * c) / pow(mag_w, 5);
        k->df2 = (pow(w_dt, 2) - 4 * c - 4 * w_dt * s + pow(w_dt, 2) * c + 4) / (pow(mag_w, 6));
        k->df3 = (2 * (c - 1) + w_dt * s) / (pow(mag_w, 4));
        k->df4 = (2 * w_dt + w_dt * c - 3 * s) / (pow(mag_w, 5));
    }
}

/* I - a*W + b*W2  (the three rotation expressions CpiV1.h:119-120, 267-268; CpiV2.h:138-139, 321-322) */
static void rot_from(double a, double b, const double* W, const double* W2, double* R) {
    for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0 ? 1.0 : 0.0) - a * W[k] + b * W2[k];
}

/* Dense Lyapunov right-hand side  F P + P F^T + G Qc G^T  for n = 15 or 21  (CpiV1.h:291, CpiV2.h:348) */
static void lyap_rhs(int n, const double* F, const double* G, const double* Qc, const double* P, double* out) {
    double FP[441], PFt[441], GQ[21 * 12], GQGt[441];
    gemm(n, n, n, F, P, FP);
    gemm_nt(n, n, n, P, F, PFt);
    gemm(n, 12, 12, G, Qc, GQ);
    gemm_nt(n, n, 12, GQ, G, GQGt);
    for (int k = 0; k < n * n; k++) out[k] = FP[k] + PFt[k] + GQGt[k];
}

/* F and G for one RK4 stage.  CpiV1.h:276-288 (n=15); CpiV2.h:330-344 (n=21, with the two gravity blocks) */
static void build_FG(int n, const double* w_x, const double* a_x, const double* Rstar, const double* R_old,
                     const double* RGk, const double* grav, double* F, double* G) {
    double RsT[9], t[9], eye[9];
    mat_zero(F, n * n); mat_zero(G, n * 12); mat_eye(eye, 3);
    transpose(3, 3, Rstar, RsT);
    set_block(F, n, 0, 0, 3, 3, w_x, -1.0);
    set_block(F, n, 0, 3, 3, 3, eye, -1.0);
    gemm(3, 3, 3, RsT, a_x, t); set_block(F, n, 6, 0, 3, 3, t, -1.0);
    set_block(F, n, 6, 9, 3, 3, RsT, -1.0);
    set_block(F, n, 12, 6, 3, 3, eye, 1.0);
    if (n == 21) {
        double g_k[3], g_tau[3], sk[9], t2[9];
        m3_vec(RGk, grav, g_k);                       /* R_G_to_k * grav */
        m3_vec(R_old, g_k, g_tau);                    /* R_k2tau * R_G_to_k * grav  (always the OLD R_k2tau) */
        skew_x(g_tau, sk); gemm(3, 3, 3, RsT, sk, t); set_block(F, n, 6, 15, 3, 3, t, -1.0);        /* CpiV2.h:335 */
        skew_x(g_k, sk); gemm(3, 3, 3, RsT, R_old, t2); gemm(3, 3, 3, t2, sk, t); set_block(F, n, 6, 18, 3, 3, t, -1.0); /* :336 */
    }
    set_block(G, n, 0, 0, 3, 3, eye, -1.0);
    set_block(G, n, 3, 3, 3, 3, eye, 1.0);
    set_block(G, n, 6, 6, 3, 3, RsT, -1.0);
    set_block(G, n, 9, 9, 3, 3, eye, 1.0);
}

/* The analytic mean / bias-Jacobian part that the two models share textually.
 * model 1: CpiV1.h:119-259.  model 2: CpiV2.h:138-305 (adds O_a/O_b and the gravity terms, incl. the sign slip at :296-297).
 * Outputs R1 (= R_k2tau1) and leaves c->R untouched (the caller commits it after the covariance, CpiV1.h:357). */
static void means_and_jacobians(cpi_t* c, int model, double dt, const double* w_hat, double* a_hat,
                                const double* a_m_1, double mag_w, double w_dt, int small_w, double sin_wt, double cos_wt,
                                const double* w_x, const double* w_tx, const double* w_x_2, double* R1) {
    double eye[9], Rstep[9], R1T[9];
    mat_eye(eye, 3);
    double dt_2 = pow(dt, 2);
    if (small_w) rot_from(dt, pow(dt, 2) / 2, w_x, w_x_2, Rstep);
    else rot_from(sin_wt / mag_w, (1.0 - cos_wt) / (pow(mag_w, 2.0)), w_x, w_x_2, Rstep);
    gemm(3, 3, 3, Rstep, c->R, R1);
    transpose(3, 3, R1, R1T);

    double RGk[9], g_k[3] = {0, 0, 0}, g_tau[3] = {0, 0, 0};
    if (model == 2) {
        oracle_quat_2_Rot(c->q_lin, RGk);
        if (c->imu_avg) {            /* CpiV2.h:146-149 : average the LOCAL acceleration using the NEW rotation */
            double t[3], u[3];
            m3_vec(RGk, c->grav, t); m3_vec(R1, t, u);
            for (int i = 0; i < 3; i++) { a_hat[i] += a_m_1[i] - c->b_a[i] - u[i]; a_hat[i] = 0.5 * a_hat[i]; }
        }
        m3_vec(RGk, c->grav, g_k);   /* :202 */
        m3_vec(c->R, g_k, g_tau);    /* :277 */
    }

    coef_t k; coefficients(small_w, dt, mag_w, w_dt, sin_wt, cos_wt, &k);
    double alpha_arg[9], Beta_arg[9], H_al[9], H_be[9], t3[3];
    for (int i = 0; i < 9; i++) {
        alpha_arg[i] = ((dt_2 / 2.0) * eye[i] + k.f1 * w_x[i] + k.f2 * w_x_2[i]);
        Beta_arg[i] = (dt * eye[i] + k.f3 * w_x[i] + k.f4 * w_x_2[i]);
    }
    gemm(3, 3, 3, R1T, alpha_arg, H_al);
    gemm(3, 3, 3, R1T, Beta_arg, H_be);
    m3_vec(H_al, a_hat, t3);
    for (int i = 0; i < 3; i++) c->alpha[i] += c->beta[i] * dt + t3[i];     /* old beta */
    m3_vec(H_be, a_hat, t3);
    for (int i = 0; i < 3; i++) c->beta[i] += t3[i];

    /* right Jacobian and J_q */
    double J_r[9], wtx2[9], J_save[9], t9[9];
    gemm(3, 3, 3, w_tx, w_tx, wtx2);
    if (small_w) for (int i = 0; i < 9; i++) J_r[i] = eye[i] - .5 * w_tx[i] + (1.0 / 6.0) * wtx2[i];
    else for (int i = 0; i < 9; i++)
        J_r[i] = eye[i] - ((1 - cos_wt) / (pow((w_dt), 2.0))) * w_tx[i] + ((w_dt - sin_wt) / (pow(w_dt, 3.0))) * wtx2[i];
    memcpy(J_save, c->J_q, sizeof J_save);
    gemm(3, 3, 3, Rstep, c->J_q, t9);
    for (int i = 0; i < 9; i++) c->J_q[i] = t9[i] + J_r[i] * dt;

    for (int i = 0; i < 9; i++) c->H_a[i] -= H_al[i];
    for (int i = 0; i < 9; i++) c->H_a[i] += dt * c->H_b[i];      /* old H_b */
    for (int i = 0; i < 9; i++) c->H_b[i] -= H_be[i];

    if (model == 2) {                /* CpiV2.h:203-205 */
        double sk[9], t1[9], t2[9];
        skew_x(g_k, sk);
        for (int i = 0; i < 9; i++) c->O_a[i] += dt * c->O_b[i];
        gemm(3, 3, 3, H_al, c->R, t1); gemm(3, 3, 3, t1, sk, t2);
        for (int i = 0; i < 9; i++) c->O_a[i] += -t2[i];
        gemm(3, 3, 3, H_be, c->R, t1); gemm(3, 3, 3, t1, sk, t2);
        for (int i = 0; i < 9; i++) c->O_b[i] += -t2[i];
    }

    for (int i = 0; i < 9; i++) c->J_a[i] += c->J_b[i] * dt;      /* old J_b */
    for (int col = 0; col < 3; col++) {
        double e[3] = {0, 0, 0}, ex[9], Jq_e[3], sk[9], dR[9], inner[9], exw[9], wex[9], M[9], v[3];
        e[col] = 1.0; skew_x(e, ex);
        m3_vec(c->J_q, e, Jq_e);                /* NEW J_q */
        skew_x(Jq_e, sk); gemm(3, 3, 3, R1T, sk, dR);
        for (int i = 0; i < 9; i++) dR[i] = -dR[i];
        gemm(3, 3, 3, ex, w_x, exw); gemm(3, 3, 3, w_x, ex, wex);
        double wi = w_hat[col];
        /* J_a column */
        for (int i = 0; i < 9; i++) inner[i] = (wi * k.df1) * w_x[i] - k.f1 * ex[i] + (wi * k.df2) * w_x_2[i] - k.f2 * (exw[i] + wex[i]);
        gemm(3, 3, 3, dR, alpha_arg, M); gemm(3, 3, 3, R1T, inner, t9);
        for (int i = 0; i < 9; i++) M[i] += t9[i];
        m3_vec(M, a_hat, v);
        if (model == 2) {
            double Js_e[3], sk2[9], t1[9], u[3];
            m3_vec(J_save, e, Js_e); skew_x(Js_e, sk2); gemm(3, 3, 3, H_al, sk2, t1); m3_vec(t1, g_tau, u);
            for (int i = 0; i < 3; i++) v[i] = v[i] - u[i];                          /* CpiV2.h:285, 289, 293 */
        }
        for (int i = 0; i < 3; i++) E(c->J_a, 3, i, col) += v[i];
        /* J_b column */
        for (int i = 0; i < 9; i++) inner[i] = (wi * k.df3) * w_x[i] - k.f3 * ex[i] + (wi * k.df4) * w_x_2[i] - k.f4 * (exw[i] + wex[i]);
        gemm(3, 3, 3, dR, Beta_arg, M); gemm(3, 3, 3, R1T, inner, t9);
        for (int i = 0; i < 9; i++) M[i] += t9[i];
        m3_vec(M, a_hat, v);
        if (model == 2) {
            double Js_e[3], sk2[9], t1[9], u[3];
            m3_vec(J_save, e, Js_e); skew_x(Js_e, sk2); gemm(3, 3, 3, H_be, sk2, t1); m3_vec(t1, g_tau, u);
            /* CpiV2.h:296-297 reads "a_hat-\n-H_be*..." : a double minus, i.e. PLUS, for column 0 only */
            for (int i = 0; i < 3; i++) v[i] = (col == 0) ? v[i] + u[i] : v[i] - u[i];
        }
        for (int i = 0; i < 3; i++) E(c->J_b, 3, i, col) += v[i];
    }
}

/* CpiV1::feed_IMU, CpiV1.h:62-361 */
static void feed_imu_v1(cpi_t* c, double dt, const double* w_m_0, const double* a_m_0, const double* w_m_1, const double* a_m_1) {
    c->DT += dt;                                             /* :69 */
    if (dt == 0) return;                                     /* :72-74 */
    double w_hat[3], a_hat[3], w_hatdt[3];
    for (int i = 0; i < 3; i++) { w_hat[i] = w_m_0[i] - c->b_w[i]; a_hat[i] = a_m_0[i] - c->b_a[i]; }
    if (c->imu_avg) for (int i = 0; i < 3; i++) {            /* :81-86 */
        w_hat[i] += w_m_1[i] - c->b_w[i]; w_hat[i] = 0.5 * w_hat[i];
        a_hat[i] += a_m_1[i] - c->b_a[i]; a_hat[i] = .5 * a_hat[i];
    }
    for (int i = 0; i < 3; i++) w_hatdt[i] = w_hat[i] * dt;
    double mag_w = norm3(w_hat), w_dt = mag_w * dt;
    int small_w = (mag_w < 0.008726646);                     /* :101 */
    double cos_wt = cos(w_dt), sin_wt = sin(w_dt);
    double w_x[9], a_x[9], w_tx[9], w_x_2[9], R1[9];
    skew_x(w_hat, w_x); skew_x(a_hat, a_x); skew_x(w_hatdt, w_tx); gemm(3, 3, 3, w_x, w_x, w_x_2);

    means_and_jacobians(c, 1, dt, w_hat, a_hat, a_m_1, mag_w, w_dt, small_w, sin_wt, cos_wt, w_x, w_tx, w_x_2, R1);

    /* covariance, RK4 on the Lyapunov ODE, :267-353 */
    double R_mid[9], t9[9];
    if (small_w) rot_from(.5 * dt, pow(.5 * dt, 2) / 2, w_x, w_x_2, t9);
    else rot_from(sin(mag_w * .5 * dt) / mag_w, (1.0 - cos(mag_w * .5 * dt)) / (pow(mag_w, 2.0)), w_x, w_x_2, t9);
    gemm(3, 3, 3, t9, c->R, R_mid);

    double F[225], G[180], k1[225], k2[225], k3[225], k4[225], Ps[225];
    build_FG(15, w_x, a_x, c->R, 0, 0, 0, F, G);  lyap_rhs(15, F, G, c->Qc, c->P, k1);
    build_FG(15, w_x, a_x, R_mid, 0, 0, 0, F, G);
    for (int i = 0; i < 225; i++) Ps[i] = c->P[i] + k1[i] * dt / 2.0;
    lyap_rhs(15, F, G, c->Qc, Ps, k2);
    for (int i = 0; i < 225; i++) Ps[i] = c->P[i] + k2[i] * dt / 2.0;
    lyap_rhs(15, F, G, c->Qc, Ps, k3);
    build_FG(15, w_x, a_x, R1, 0, 0, 0, F, G);
    for (int i = 0; i < 225; i++) Ps[i] = c->P[i] + k3[i] * dt;
    lyap_rhs(15, F, G, c->Qc, Ps, k4);
    for (int i = 0; i < 225; i++) c->P[i] += (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    transpose(15, 15, c->P, Ps);
    for (int i = 0; i < 225; i++) c->P[i] = 0.5 * (c->P[i] + Ps[i]);         /* :353 */

    memcpy(c->R, R1, sizeof c->R);                                            /* :357 */
    oracle_rot_2_quat(c->R, c->q);                                            /* :358 */
}

/* CpiV2::feed_IMU, CpiV2.h:84-467 */
static void feed_imu_v2(cpi_t* c, double dt, const double* w_m_0, const double* a_m_0, const double* w_m_1, const double* a_m_1) {
    c->DT += dt;
    if (dt == 0) return;
    double RGk[9], w_hat[3], a_hat[3], w_hatdt[3], t3[3], u3[3];
    oracle_quat_2_Rot(c->q_lin, RGk);
    m3_vec(RGk, c->grav, t3); m3_vec(c->R, t3, u3);          /* R_k2tau*quat_2_Rot(q_k_lin)*grav, :99 */
    for (int i = 0; i < 3; i++) { w_hat[i] = w_m_0[i] - c->b_w[i]; a_hat[i] = a_m_0[i] - c->b_a[i] - u3[i]; }
    if (c->imu_avg) for (int i = 0; i < 3; i++) { w_hat[i] += w_m_1[i] - c->b_w[i]; w_hat[i] = 0.5 * w_hat[i]; }
    for (int i = 0; i < 3; i++) w_hatdt[i] = w_hat[i] * dt;
    double mag_w = norm3(w_hat), w_dt = mag_w * dt;
    int small_w = (mag_w < 0.008726646);
    double cos_wt = cos(w_dt), sin_wt = sin(w_dt);
    double w_x[9], a_x[9], w_tx[9], w_x_2[9], R1[9];
    skew_x(w_hat, w_x); skew_x(w_hatdt, w_tx); gemm(3, 3, 3, w_x, w_x, w_x_2);

    means_and_jacobians(c, 2, dt, w_hat, a_hat, a_m_1, mag_w, w_dt, small_w, sin_wt, cos_wt, w_x, w_tx, w_x_2, R1);
    skew_x(a_hat, a_x);                                      /* :150 (after the optional averaging) */

    double dt_mid = dt / 2.0, w_dt_mid = mag_w * dt_mid, R_mid[9], t9[9];
    if (small_w) rot_from(dt_mid, pow(dt_mid, 2) / 2, w_x, w_x_2, t9);
    else rot_from(sin(w_dt_mid) / mag_w, (1.0 - cos(w_dt_mid)) / (pow(mag_w, 2.0)), w_x, w_x_2, t9);
    gemm(3, 3, 3, t9, c->R, R_mid);

    enum { N = 21, NN = 441 };
    double F1[NN], F2[NN], F4[NN], G[N * 12], I21[NN];
    double Pd1[NN], Pd2[NN], Pd3[NN], Pd4[NN], Ph1[NN], Ph2[NN], Ph3[NN], Ph4[NN], Ps[NN], Phs[NN];
    mat_eye(I21, N);
    /* k1 :330-348 */
    build_FG(N, w_x, a_x, c->R, c->R, RGk, c->grav, F1, G);
    memcpy(Ph1, F1, sizeof Ph1);
    lyap_rhs(N, F1, G, c->Qc, c->Pbig, Pd1);
    /* k2 :354-374 */
    build_FG(N, w_x, a_x, R_mid, c->R, RGk, c->grav, F2, G);
    for (int i = 0; i < NN; i++) { Phs[i] = I21[i] + Ph1[i] * dt_mid; Ps[i] = c->Pbig[i] + Pd1[i] * dt_mid; }
    gemm(N, N, N, F2, Phs, Ph2);
    lyap_rhs(N, F2, G, c->Qc, Ps, Pd2);
    /* k3 :381-388 */
    for (int i = 0; i < NN; i++) { Phs[i] = I21[i] + Ph2[i] * dt_mid; Ps[i] = c->Pbig[i] + Pd2[i] * dt_mid; }
    gemm(N, N, N, F2, Phs, Ph3);
    lyap_rhs(N, F2, G, c->Qc, Ps, Pd3);
    /* k4 :394-414 */
    build_FG(N, w_x, a_x, R1, c->R, RGk, c->grav, F4, G);
    for (int i = 0; i < NN; i++) { Phs[i] = I21[i] + Ph3[i] * dt; Ps[i] = c->Pbig[i] + Pd3[i] * dt; }
    gemm(N, N, N, F4, Phs, Ph4);
    lyap_rhs(N, F4, G, c->Qc, Ps, Pd4);
    /* combine :421-426 */
    double Phi[NN];
    for (int i = 0; i < NN; i++) c->Pbig[i] += (dt / 6.0) * (Pd1[i] + 2.0 * Pd2[i] + 2.0 * Pd3[i] + Pd4[i]);
    transpose(N, N, c->Pbig, Ps);
    for (int i = 0; i < NN; i++) c->Pbig[i] = 0.5 * (c->Pbig[i] + Ps[i]);
    for (int i = 0; i < NN; i++) Phi[i] = I21[i] + (dt / 6.0) * (Ph1[i] + 2.0 * Ph2[i] + 2.0 * Ph3[i] + Ph4[i]);
    /* clone + marginalise :436-446 */
    double B[NN], T1[NN], T2[NN];
    mat_eye(B, N);
    for (int i = 0; i < 3; i++) { E(B, N, 15 + i, 15 + i) = 0.0; E(B, N, 15 + i, i) = 1.0; }
    gemm(N, N, N, B, c->Pbig, T1); gemm_nt(N, N, N, T1, B, T2);
    transpose(N, N, T2, T1);
    for (int i = 0; i < NN; i++) c->Pbig[i] = 0.5 * (T2[i] + T1[i]);
    gemm(N, N, N, Phi, c->D, T1); gemm(N, N, N, B, T1, T2);     /* B_k * Phi * Discrete_J_b */
    memcpy(c->D, T2, sizeof T2);
    get_block(c->Pbig, N, 0, 0, 15, 15, c->P);
    if (c->stj) {                                               /* :450-458 */
        get_block(c->D, N, 0, 3, 3, 3, c->J_q); for (int i = 0; i < 9; i++) c->J_q[i] = -c->J_q[i];
        get_block(c->D, N, 12, 3, 3, 3, c->J_a); get_block(c->D, N, 6, 3, 3, 3, c->J_b);
        get_block(c->D, N, 12, 9, 3, 3, c->H_a); get_block(c->D, N, 6, 9, 3, 3, c->H_b);
        get_block(c->D, N, 12, 18, 3, 3, c->O_a); get_block(c->D, N, 6, 18, 3, 3, c->O_b);
    }
    memcpy(c->R, R1, sizeof c->R);
    oracle_rot_2_quat(c->R, c->q);
}

static void window(int model, const double* s, int64_t entries, const double* lin, const double* sig, int flags, double* rec) {
    int avg = (flags & CPI_FLAG_IMU_AVG) != 0;
    int64_t steps = avg ? (entries > 0 ? entries - 1 : 0) : entries;
    cpi_t* c = (cpi_t*)malloc(sizeof(cpi_t));
    cpi_init(c, sig, lin, avg, (flags & CPI_FLAG_ANALYTIC_JACOBIANS) == 0);
    for (int64_t i = 0; i < steps; i++) {       /* driver loop: GraphSolver_IMU.cpp:50-69 */
        const double* e0 = s + i * CPI_SAMPLE_DOUBLES;
        const double* e1 = avg ? e0 + CPI_SAMPLE_DOUBLES : e0;
        if (model == 1) feed_imu_v1(c, e0[6], e0, e0 + 3, e1, e1 + 3);
        else feed_imu_v2(c, e0[6], e0, e0 + 3, e1, e1 + 3);
    }
    memcpy(rec + CPI_REC_Q, c->q, 32); memcpy(rec + CPI_REC_R, c->R, 72);
    memcpy(rec + CPI_REC_ALPHA, c->alpha, 24); memcpy(rec + CPI_REC_BETA, c->beta, 24);
    rec[CPI_REC_DT] = c->DT;
    memcpy(rec + CPI_REC_JQ, c->J_q, 72); memcpy(rec + CPI_REC_JA, c->J_a, 72); memcpy(rec + CPI_REC_JB, c->J_b, 72);
    memcpy(rec + CPI_REC_HA, c->H_a, 72); memcpy(rec + CPI_REC_HB, c->H_b, 72);
    memcpy(rec + CPI_REC_P, c->P, 225 * 8);
    if (model == 2) { memcpy(rec + CPI_REC_OA, c->O_a, 72); memcpy(rec + CPI_REC_OB, c->O_b, 72); }
    free(c);
}

/* ------------------------------------------------------------------ gtsam/ImuFactorCPIv1.cpp, ImuFactorCPIv2.cpp -- */

/* w*I - skew(v) (sign = -1) or w*I + skew(v) (sign = +1) */
static void quat_mat(const double* q, double sign, double* M) {
    double sk[9]; skew_x(q, sk);
    for (int i = 0; i < 9; i++) M[i] = q[3] * (i % 4 == 0 ? 1.0 : 0.0) + sign * sk[i];
}

/* ImuFactorCPIv1::evaluateError (ImuFactorCPIv1.cpp:37-208); model 2: ImuFactorCPIv2.cpp:38-212 */
static void factor_eval(int model, const double* xi, const double* xj, const double* r, const double* l,
                        double* err, double* H1, double* H2) {
    const double *q_GtoK = xi, *bg_K = xi + 4, *v_K = xi + 7, *ba_K = xi + 10, *p_K = xi + 13;
    const double *q_GtoK1 = xj, *bg_K1 = xj + 4, *v_K1 = xj + 7, *ba_K1 = xj + 10, *p_K1 = xj + 13;
    const double *bg_lin = l, *ba_lin = l + 3, *q_K_lin = l + 6, *grav = l + 10;
    const double *q_KtoK1 = r + CPI_REC_Q, *alpha = r + CPI_REC_ALPHA, *beta = r + CPI_REC_BETA;
    const double *J_q = r + CPI_REC_JQ, *J_alpha = r + CPI_REC_JA, *J_beta = r + CPI_REC_JB;
    const double *H_alpha = r + CPI_REC_HA, *H_beta = r + CPI_REC_HB;
    const double *O_alpha = r + CPI_REC_OA, *O_beta = r + CPI_REC_OB;
    double deltatime = r[CPI_REC_DT];

    double dbg[3], dba[3], t3[3], ExpB[9], q_b[4], qi[4], q_n[4], q_rminus[4], q_r[4], q_m[4];
    for (int i = 0; i < 3; i++) { dbg[i] = bg_K[i] - bg_lin[i]; dba[i] = ba_K[i] - ba_lin[i]; }
    m3_vec(J_q, dbg, t3); for (int i = 0; i < 3; i++) t3[i] = -t3[i];
    oracle_Exp(t3, ExpB); oracle_rot_2_quat(ExpB, q_b);                       /* :57-58 */
    quat_inv(q_GtoK, qi); oracle_quat_multiply(q_GtoK1, qi, q_n);             /* :61 */
    quat_inv(q_KtoK1, qi); oracle_quat_multiply(q_n, qi, q_rminus);           /* :62 */
    oracle_quat_multiply(q_rminus, q_b, q_r);                                 /* :63 */
    quat_inv(q_b, qi); oracle_quat_multiply(qi, q_KtoK1, q_m);                /* :64 */

    double q_kR[4] = {0, 0, 0, 1}, dthk[3] = {0, 0, 0};
    if (model == 2) {                                                         /* v2 :68-69 */
        quat_inv(q_K_lin, qi); oracle_quat_multiply(q_GtoK, qi, q_kR);
        for (int i = 0; i < 3; i++) dthk[i] = 2 * q_kR[i];
    }
    double Rk[9], pa[3], pb[3], alphahat[3], betahat[3], u[3], w[3], o[3];
    oracle_quat_2_Rot(q_GtoK, Rk);
    for (int i = 0; i < 3; i++) {
        if (model == 1) {                                                     /* v1 :70, :72 */
            pa[i] = p_K1[i] - p_K[i] - v_K[i] * deltatime + 0.5 * grav[i] * pow(deltatime, 2);
            pb[i] = v_K1[i] - v_K[i] + grav[i] * deltatime;
        } else {                                                              /* v2 :72, :74 */
            pa[i] = p_K1[i] - p_K[i] - v_K[i] * deltatime;
            pb[i] = v_K1[i] - v_K[i];
        }
    }
    m3_vec(Rk, pa, alphahat); m3_vec(J_alpha, dbg, u); m3_vec(H_alpha, dba, w);
    if (model == 2) m3_vec(O_alpha, dthk, o); else o[0] = o[1] = o[2] = 0;
    for (int i = 0; i < 3; i++) alphahat[i] = model == 2 ? alphahat[i] - u[i] - w[i] - o[i] : alphahat[i] - u[i] - w[i];
    m3_vec(Rk, pb, betahat); m3_vec(J_beta, dbg, u); m3_vec(H_beta, dba, w);
    if (model == 2) m3_vec(O_beta, dthk, o);
    for (int i = 0; i < 3; i++) betahat[i] = model == 2 ? betahat[i] - u[i] - w[i] - o[i] : betahat[i] - u[i] - w[i];

    for (int i = 0; i < 3; i++) {                                             /* :84-88 */
        err[i] = 2 * q_r[i];
        err[3 + i] = bg_K1[i] - bg_K[i];
        err[6 + i] = betahat[i] - beta[i];
        err[9 + i] = ba_K1[i] - ba_K[i];
        err[12 + i] = alphahat[i] - alpha[i];
    }
    double eye[9]; mat_eye(eye, 3);
    if (H1) {                                                                 /* :98-154 */
        double Hi[225], A[9], Bm[9], AB[9], blk[9], sk[9], v3[3];
        mat_zero(Hi, 225);
        quat_mat(q_n, -1.0, A); quat_mat(q_m, -1.0, Bm); gemm(3, 3, 3, A, Bm, AB);
        for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) E(blk, 3, i, j) = -(E(AB, 3, i, j) + q_n[i] * q_m[j]);
        set_block(Hi, 15, 0, 0, 3, 3, blk, 1.0);                              /* :109-111 */
        m3_vec(Rk, pb, v3); skew_x(v3, sk);
        if (model == 2) { double qm[9], t[9]; quat_mat(q_kR, +1.0, qm); gemm(3, 3, 3, O_beta, qm, t); for (int i = 0; i < 9; i++) sk[i] -= t[i]; }
        set_block(Hi, 15, 6, 0, 3, 3, sk, 1.0);                               /* :113 / v2 :115-116 */
        m3_vec(Rk, pa, v3); skew_x(v3, sk);
        if (model == 2) { double qm[9], t[9]; quat_mat(q_kR, +1.0, qm); gemm(3, 3, 3, O_alpha, qm, t); for (int i = 0; i < 9; i++) sk[i] -= t[i]; }
        set_block(Hi, 15, 12, 0, 3, 3, sk, 1.0);                              /* :115 / v2 :118-119 */
        quat_mat(q_rminus, -1.0, A); gemm(3, 3, 3, A, J_q, blk);
        set_block(Hi, 15, 0, 3, 3, 3, blk, 1.0);                              /* :119 */
        set_block(Hi, 15, 3, 3, 3, 3, eye, -1.0);                             /* :121 */
        set_block(Hi, 15, 6, 3, 3, 3, J_beta, -1.0);                          /* :123 */
        set_block(Hi, 15, 12, 3, 3, 3, J_alpha, -1.0);                        /* :125 */
        set_block(Hi, 15, 6, 6, 3, 3, Rk, -1.0);                              /* :129 */
        set_block(Hi, 15, 12, 6, 3, 3, Rk, -deltatime);                       /* :131 */
        set_block(Hi, 15, 6, 9, 3, 3, H_beta, -1.0);                          /* :135 */
        set_block(Hi, 15, 9, 9, 3, 3, eye, -1.0);                             /* :137 */
        set_block(Hi, 15, 12, 9, 3, 3, H_alpha, -1.0);                        /* :139 */
        set_block(Hi, 15, 12, 12, 3, 3, Rk, -1.0);                            /* :143 */
        memcpy(H1, Hi, sizeof Hi);
    }
    if (H2) {                                                                 /* :158-198 */
        double Hj[225], A[9];
        mat_zero(Hj, 225);
        quat_mat(q_r, +1.0, A); set_block(Hj, 15, 0, 0, 3, 3, A, 1.0);        /* :169 */
        set_block(Hj, 15, 3, 3, 3, 3, eye, 1.0);
        set_block(Hj, 15, 6, 6, 3, 3, Rk, 1.0);
        set_block(Hj, 15, 9, 9, 3, 3, eye, 1.0);
        set_block(Hj, 15, 12, 12, 3, 3, Rk, 1.0);
        memcpy(H2, Hj, sizeof Hj);
    }
}

/* getpredictedstate_v1 / _v2, solvers/GraphSolver_IMU.cpp:263-307 */
static void predict_state(int model, const double* x, const double* r, const double* l, double* o) {
    const double *q_GtoK = x, *bg_K = x + 4, *v_K = x + 7, *ba_K = x + 10, *p_K = x + 13, *grav = l + 10;
    double dt = r[CPI_REC_DT], qi[4], Rinv[9], rb[3], ra[3];
    oracle_quat_multiply(r + CPI_REC_Q, q_GtoK, o);
    quat_inv(q_GtoK, qi); oracle_quat_2_Rot(qi, Rinv);
    m3_vec(Rinv, r + CPI_REC_BETA, rb); m3_vec(Rinv, r + CPI_REC_ALPHA, ra);
    for (int i = 0; i < 3; i++) {
        o[4 + i] = bg_K[i]; o[10 + i] = ba_K[i];
        if (model == 1) {
            o[7 + i] = v_K[i] - grav[i] * dt + rb[i];
            o[13 + i] = p_K[i] + v_K[i] * dt - 0.5 * grav[i] * pow(dt, 2) + ra[i];
        } else {
            o[7 + i] = v_K[i] + rb[i];
            o[13 + i] = p_K[i] + v_K[i] * dt + ra[i];
        }
    }
}

/* JPLNavState::retract, gtsam/JPLNavState.cpp:37-71 */
static void retract(const double* x, const double* xi, double* o) {
    double n = norm3(xi), dq[4];
    for (int i = 0; i < 3; i++) dq[i] = ((sin(n / 2) / n)) * xi[i];
    dq[3] = cos(n / 2);
    double nn = norm4(dq);
    for (int i = 0; i < 4; i++) dq[i] = dq[i] / nn;
    if (dq[3] < 0) for (int i = 0; i < 4; i++) dq[i] = -dq[i];
    if (isnan(norm4(dq))) { dq[0] = dq[1] = dq[2] = 0; dq[3] = 1.0; }
    oracle_quat_multiply(dq, x, o);
    for (int i = 0; i < 12; i++) o[4 + i] = x[4 + i] + xi[3 + i];
}

/* ------------------------------------------------------------------ batch drivers -------------------------------- */
typedef struct {
    int kind, model, flags, t, nt; int64_t n;
    const int64_t *offsets, *idx_i, *idx_j; int64_t ns;
    const double *samples, *lin, *sig, *states, *records;
    double *out, *e, *H1, *H2;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    int64_t lo = j->n * j->t / j->nt, hi = j->n * (j->t + 1) / j->nt;
    int rd = j->model == 1 ? CPI_REC_V1_DOUBLES : CPI_REC_V2_DOUBLES;
    int avg = (j->flags & CPI_FLAG_IMU_AVG) != 0;
    for (int64_t w = lo; w < hi; w++) {
        if (j->kind == 0) {
            int64_t o0 = j->offsets ? j->offsets[w] : w * (j->ns + avg);
            int64_t o1 = j->offsets ? j->offsets[w + 1] : o0 + j->ns + avg;
            window(j->model, j->samples + o0 * CPI_SAMPLE_DOUBLES, o1 - o0, j->lin + w * CPI_LIN_DOUBLES, j->sig, j->flags, j->out + w * (int64_t)rd);
        } else {
            int64_t a = j->idx_i ? j->idx_i[w] : w, b = j->idx_j ? j->idx_j[w] : w + 1;
            factor_eval(j->model, j->states + a * CPI_STATE_DOUBLES, j->states + b * CPI_STATE_DOUBLES,
                        j->records + w * (int64_t)rd, j->lin + w * CPI_LIN_DOUBLES, j->e + w * 15,
                        j->H1 ? j->H1 + w * 225 : 0, j->H2 ? j->H2 + w * 225 : 0);
        }
    }
    return 0;
}

static void run(job_t* proto, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    job_t jobs[256]; pthread_t th[256];
    for (int t = 0; t < nthreads; t++) { jobs[t] = *proto; jobs[t].t = t; jobs[t].nt = nthreads; }
    if (nthreads == 1) { worker(&jobs[0]); return; }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], 0, worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
}

int oracle_cpi_preintegrate(int model, int64_t n_windows, const int64_t* offsets, int64_t ns_uniform,
                            const double* samples, const double* lin, const double* sigmas, int flags,
                            double* out, int nthreads) {
    if (model != 1 && model != 2) return -1;
    job_t j; memset(&j, 0, sizeof j);
    j.kind = 0; j.model = model; j.flags = flags; j.n = n_windows; j.offsets = offsets; j.ns = ns_uniform;
    j.samples = samples; j.lin = lin; j.sig = sigmas; j.out = out;
    run(&j, nthreads);
    return 0;
}

int oracle_imu_factor_eval(int model, int64_t n, const double* states, const int64_t* idx_i, const int64_t* idx_j,
                           const double* records, const double* lin, double* e, double* H1, double* H2, int nthreads) {
    if (model != 1 && model != 2) return -1;
    job_t j; memset(&j, 0, sizeof j);
    j.kind = 1; j.model = model; j.n = n; j.states = states; j.idx_i = idx_i; j.idx_j = idx_j;
    j.records = records; j.lin = lin; j.e = e; j.H1 = H1; j.H2 = H2;
    run(&j, nthreads);
    return 0;
}

int oracle_predict_state(int model, int64_t n, const double* states_k, const double* records, const double* lin, double* states_k1) {
    if (model != 1 && model != 2) return -1;
    int rd = model == 1 ? CPI_REC_V1_DOUBLES : CPI_REC_V2_DOUBLES;
    for (int64_t i = 0; i < n; i++)
        predict_state(model, states_k + i * CPI_STATE_DOUBLES, records + i * (int64_t)rd, lin + i * CPI_LIN_DOUBLES, states_k1 + i * CPI_STATE_DOUBLES);
    return 0;
}

int oracle_retract(int64_t n, const double* states, const double* xi, double* out) {
    for (int64_t i = 0; i < n; i++) retract(states + i * CPI_STATE_DOUBLES, xi + i * 15, out + i * CPI_STATE_DOUBLES);
    return 0;
}
