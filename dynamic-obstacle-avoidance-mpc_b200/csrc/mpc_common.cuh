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

// dynamic shared memory of a kernel (the CPU SIMT emulator under oracle/simt/ hands out a per-block buffer instead)
#ifdef SIMT_EMU
#define B2_DYNAMIC_SMEM(type, name) SIMT_DYNAMIC_SMEM(type, name)
#else
#define B2_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

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
    const int32_t* NQ;           // optional: length of the queue ORDER (closed loop: finished environments are not queued); NULL: B
    // scratch
    double* WPRIV;               // warp kernel: node-private rows, one block of wpriv_stride doubles per resident CTA
    size_t wpriv_stride;
};

// ---- products with S = [B A] (5 x 7), exploiting the generated sparsity pattern -------------
// Floating-point contraction is OFF in every build of the kernels (device: -fmad=false, CPU mirror: -ffp-contract=off):
// each fused multiply-add below is written out, so that the GPU and the mirror execute the same operations.
// out[5] = [B A] * v[7]
__host__ __device__ __forceinline__ void b2_ba_mul(const double* sv, double h, double h2, const double* v, double* out)
{
    out[0] = fma(sv[SV_A03], v[5], fma(sv[SV_A02], v[4], v[2])) + fma(sv[SV_B01], v[1], fma(sv[SV_B00], v[0], sv[SV_A04] * v[6]));
    out[1] = fma(sv[SV_A13], v[5], fma(sv[SV_A12], v[4], v[3])) + fma(sv[SV_B11], v[1], fma(sv[SV_B10], v[0], sv[SV_A14] * v[6]));
    out[2] = fma(h2, v[1], fma(h, v[6], v[4]));
    out[3] = fma(h, v[0], v[5]);
    out[4] = fma(h, v[1], v[6]);
}
// out[7] = [B A]^T * w[5]
__host__ __device__ __forceinline__ void b2_bat_mul(const double* sv, double h, double h2, const double* w, double* out)
{
    out[0] = fma(h, w[3], fma(sv[SV_B10], w[1], sv[SV_B00] * w[0]));
    out[1] = fma(sv[SV_B11], w[1], sv[SV_B01] * w[0]) + fma(h, w[4], h2 * w[2]);
    out[2] = w[0];
    out[3] = w[1];
    out[4] = fma(sv[SV_A12], w[1], fma(sv[SV_A02], w[0], w[2]));
    out[5] = fma(sv[SV_A13], w[1], fma(sv[SV_A03], w[0], w[3]));
    out[6] = fma(sv[SV_A14], w[1], fma(sv[SV_A04], w[0], w[4])) + h * w[2];
}
// symmetric 5x5 stored packed lower: idx(i,j) i>=j -> i*(i+1)/2 + j
__host__ __device__ __forceinline__ constexpr int b2_sym(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
