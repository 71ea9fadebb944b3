// mpc_common.cuh -- kernel parameters and the model-structure algebra shared by the solver kernels.
//
// Replaces (B200-native, FP64) what acados assembles around src/simulation/robot_ocp_problem.py:195:
//   * discretisation/sensitivities come from csrc/model_generated.cuh (generated from the
//     expression graph of src/models/robot_model.py);
//   * the Gauss-Newton QP of robot_ocp_problem.py:59-122 is never formed
//     densely: the products with [B A] below exploit the sensitivity pattern the generator found.
#pragma once
#include <stdint.h>
#include "model_generated.cuh"

#define B2_NX 5
#define B2_NU 2
#define B2_NV 7
#define B2_MAXM 32

// variable order per node v = [u_a, u_alpha, x, y, psi, v, omega]
// box-constraint slots [u_a, u_alpha, x, y, v, omega] -> index in v
__host__ __device__ __forceinline__ int b2_slot2v(int s) { return s < 4 ? s : s + 1; }

// compile-time check that the hand-written sparse algebra below matches the generated pattern
__host__ __device__ constexpr bool b2_streq(const char* a, const char* b)
{
    return (*a == *b) && (*a == '\0' || b2_streq(a + 1, b + 1));
}
static_assert(b2_streq(MODEL_GL4_PATTERN, "10vvvvv/01vvvvv/0010c0c/00010c0/000010c"),
              "sensitivity pattern changed: regenerate and update b2_ba_* in mpc_common.cuh");
static_assert(b2_streq(MODEL_RK4_PATTERN, "10vvvvv/01vvvvv/0010c0c/00010c0/000010c"),
              "sensitivity pattern changed: regenerate and update b2_ba_* in mpc_common.cuh");
static_assert(MODEL_GL4_NSV == 10 && MODEL_RK4_NSV == 10, "unexpected number of variable sensitivities");
// sv order (generated): A02 A03 A04 B00 B01 A12 A13 A14 B10 B11
#define SV_A02 0
#define SV_A03 1
#define SV_A04 2
#define SV_B00 3
#define SV_B01 4
#define SV_A12 5
#define SV_A13 6
#define SV_A14 7
#define SV_B10 8
#define SV_B11 9

struct B2Params {
    // dims
    int32_t N, M, B, Bm;         // horizon, obstacles, batch in this launch, row stride (= max_batch)
    int32_t integrator, qp_iter_max, pred_corr, cond_pred_corr, itref_corr_max;
    int32_t nct;                 // padded constraints per node = 12 + 4M
    int32_t nlp_iter_max;        // full-SQP mode: QP limit, the four residual tolerances, line search
    int32_t nlp_globalization;   // 0 FIXED_STEP, 1 MERIT_BACKTRACKING
    double nlp_tol[4];
    double ls_alpha_min, ls_alpha_red, ls_eps;
    int32_t slots_per_stage;
    double dt, dt2h;             // h, h*h/2
    double Hd[7], Hde[7];        // Hessian diagonals (stage / terminal), LM included
    double Wg[7], Wge[7];        // gradient weights: g = Wg * (y - yref) in v order (psi entry 0)
    double lbx[4], ubx[4], lbu[2], ubu[2];
    double lb6[6], ub6[6];       // the same bounds in box-slot order [u_a, u_alpha, x, y, v, omega]
    double r2, lh, uh, zu, Zu;
    double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max;
    double reg_prim, lam_min, t_min, tau_min, thr0;
    // inputs (SoA, stride Bm)
    const double *X0, *P, *YREF, *ZL, *ZZL;
    // iterate (in/out)
    double *X, *U, *SL, *SU;
    // outputs
    double *PI, *LAM, *T, *STATS;
    int32_t *STATUS, *QP_ITER, *QP_STATUS;
    double* NLP_RES;             // full-SQP mode outputs
    int32_t* SQP_ITER;
    const int32_t* ORDER;        // optional launch order of the problems (warp kernel)
    int32_t* NEXT;               // work counter of the persistent warp kernel (zeroed before each launch)
    // scratch
    double* WS;                  // thread-per-problem kernel: all per-problem work arrays
    double* WPRIV;               // warp kernel: node-private rows, one block of wpriv_stride doubles per resident CTA
    size_t wpriv_stride;
};

// ---- products with S = [B A] (5 x 7), exploiting the generated sparsity pattern -------------
// out[5] = [B A] * v[7]
__device__ __forceinline__ void b2_ba_mul(const double* sv, double h, double h2, const double* v, double* out)
{
    out[0] = v[2] + sv[SV_A02] * v[4] + sv[SV_A03] * v[5] + sv[SV_A04] * v[6] + sv[SV_B00] * v[0] + sv[SV_B01] * v[1];
    out[1] = v[3] + sv[SV_A12] * v[4] + sv[SV_A13] * v[5] + sv[SV_A14] * v[6] + sv[SV_B10] * v[0] + sv[SV_B11] * v[1];
    out[2] = v[4] + h * v[6] + h2 * v[1];
    out[3] = v[5] + h * v[0];
    out[4] = v[6] + h * v[1];
}
// out[7] = [B A]^T * w[5]
__device__ __forceinline__ void b2_bat_mul(const double* sv, double h, double h2, const double* w, double* out)
{
    out[0] = sv[SV_B00] * w[0] + sv[SV_B10] * w[1] + h * w[3];
    out[1] = sv[SV_B01] * w[0] + sv[SV_B11] * w[1] + h2 * w[2] + h * w[4];
    out[2] = w[0];
    out[3] = w[1];
    out[4] = sv[SV_A02] * w[0] + sv[SV_A12] * w[1] + w[2];
    out[5] = sv[SV_A03] * w[0] + sv[SV_A13] * w[1] + w[3];
    out[6] = sv[SV_A04] * w[0] + sv[SV_A14] * w[1] + h * w[2] + w[4];
}
// symmetric 5x5 stored packed lower: idx(i,j) i>=j -> i*(i+1)/2 + j
__host__ __device__ __forceinline__ constexpr int b2_sym(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
// y[5] = P * x[5]
__device__ __forceinline__ void b2_symv(const double* P, const double* x, double* y)
{
#pragma unroll
    for (int i = 0; i < 5; i++) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 5; j++) a += P[b2_sym(i, j)] * x[j];
        y[i] = a;
    }
}
// Mfull[28] (packed lower 7x7, v order) = [B A]^T P [B A]
__device__ __forceinline__ void b2_bat_p_ba(const double* sv, double h, double h2, const double* P, double* Mp)
{
    // W[:,j] = P * col_j([B A]) for the 7 columns
    double W[7][5];
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const double p0 = P[b2_sym(r, 0)], p1 = P[b2_sym(r, 1)], p2 = P[b2_sym(r, 2)], p3 = P[b2_sym(r, 3)], p4 = P[b2_sym(r, 4)];
        W[0][r] = sv[SV_B00] * p0 + sv[SV_B10] * p1 + h * p3;
        W[1][r] = sv[SV_B01] * p0 + sv[SV_B11] * p1 + h2 * p2 + h * p4;
        W[2][r] = p0;
        W[3][r] = p1;
        W[4][r] = sv[SV_A02] * p0 + sv[SV_A12] * p1 + p2;
        W[5][r] = sv[SV_A03] * p0 + sv[SV_A13] * p1 + p3;
        W[6][r] = sv[SV_A04] * p0 + sv[SV_A14] * p1 + h * p2 + p4;
    }
#pragma unroll
    for (int j = 0; j < 7; j++) {
        double col[7];
        b2_bat_mul(sv, h, h2, W[j], col);     // col[i] = col_i . W[:,j] = M[i][j]
#pragma unroll
        for (int i = j; i < 7; i++) Mp[i * (i + 1) / 2 + j] = col[i];
    }
}

// ---- square-root helpers (thread-level) ---------------------------------------------------------
// in-place Cholesky of a packed lower-triangular n x n matrix (idx(i,j) = i(i+1)/2 + j);
// non-positive pivot -> zero column (BLASFEO dpotrf convention)
template <int NN>
__device__ __forceinline__ void b2_chol_packed(double* Mp)
{
#pragma unroll
    for (int j = 0; j < NN; j++) {
        double d = Mp[j * (j + 1) / 2 + j];
#pragma unroll
        for (int q = 0; q < j; q++) d -= Mp[j * (j + 1) / 2 + q] * Mp[j * (j + 1) / 2 + q];
        const double l = d > 0.0 ? sqrt(d) : 0.0;
        const double li = l > 0.0 ? 1.0 / l : 0.0;
        Mp[j * (j + 1) / 2 + j] = l;
#pragma unroll
        for (int i = j + 1; i < NN; i++) {
            double a = Mp[i * (i + 1) / 2 + j];
#pragma unroll
            for (int q = 0; q < j; q++) a -= Mp[i * (i + 1) / 2 + q] * Mp[j * (j + 1) / 2 + q];
            Mp[i * (i + 1) / 2 + j] = a * li;
        }
    }
}
// y[5] = L (L^T x) for a packed lower 5x5 factor
__device__ __forceinline__ void b2_llt_mul(const double* L, const double* x, double* y)
{
    double t[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double acc = 0.0;
#pragma unroll
        for (int r = j; r < 5; r++) acc += L[b2_sym(r, j)] * x[r];
        t[j] = acc;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j <= i; j++) acc += L[b2_sym(i, j)] * t[j];
        y[i] = acc;
    }
}
// Mp[28] (packed lower 7x7) = AL AL^T with AL = [B A]^T L  (L packed lower 5x5)
__device__ __forceinline__ void b2_bat_l_sq(const double* sv, double h, double h2, const double* L, double* Mp)
{
    double AL[5][7];
#pragma unroll
    for (int q = 0; q < 5; q++) {
        double col[5];
#pragma unroll
        for (int r = 0; r < 5; r++) col[r] = r >= q ? L[b2_sym(r, q)] : 0.0;
        b2_bat_mul(sv, h, h2, col, AL[q]);
    }
#pragma unroll
    for (int i = 0; i < 7; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 5; q++) acc += AL[q][i] * AL[q][j];
            Mp[i * (i + 1) / 2 + j] = acc;
        }
}
