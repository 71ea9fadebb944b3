/*
 * mpc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the one hot path of
 * abdelhakim96/Dynamic-Obstacle-Avoidance-MPC: the acados SQP_RTI solve issued at
 * src/simulation/robot_ocp_problem.py:195 and the per-step glue around it
 * (:145-166 parameterisation, :207-212 plant step, :253-258 shift, :286-306 reset,
 * src/utils/visualization.py:25-79 obstacle model).
 *
 * PARITY UNPINNED: the arithmetic of that call lives in acados / HPIPM / BLASFEO /
 * CasADi-generated C, none of which is vendored under /root/reference, pinned by
 * requirements.txt (empty) or installable offline.  This file restates the
 * published algorithms (SQP_RTI Gauss-Newton preparation + HPIPM's Mehrotra
 * predictor-corrector Riccati interior point, acados' default option set for
 * `PARTIAL_CONDENSING_HPIPM`) from SURVEY.md Appendices A-D.  It is validated in
 * tests/ against independent dense-KKT solves, a generic IRK-Newton integrator and
 * KKT optimality checks, not against acados binaries.  The one link to the real solver
 * is at the closed-loop level: the reference's recorded experiment outcomes
 * (src/simulation/test_data, 8 settings x 100 seeded episodes) are reproduced per seed
 * by the same algorithm on the GPU (54-63 % of the episodes
 * identical at QP_ITER = 100) -- the iterates of a single solve() remain unpinned.
 * Round 2: HPIPM's lq_fact switch is restated here too (orc_config.lq_fact) and replayed against
 * those recordings: no effect beyond their noise.
 *
 * This is the INDEPENDENT checker (different code, plain Cholesky, sequential sums, libm sincos);
 * the bit-exact checker of the CUDA kernels is the mirror under oracle/simt/ (the kernel source
 * compiled for the CPU).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The CUDA product never links it.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NX 5
#define ORC_NU 2
#define ORC_NV 7
#define ORC_NY 6
#define ORC_NBX 4
#define ORC_MAXM 32   /* max obstacles (= general constraints) per node */
#define ORC_MAXN 128  /* max horizon */

/* Everything robot_ocp_problem.py:54-132 and world_specification.py:7-48 fix at
 * solver-construction time, plus the HPIPM option set acados installs. */
typedef struct orc_config {
    int32_t N;               /* N_SOLV, world_specification.py:44 */
    int32_t M;               /* N_OBST, world_specification.py:25 */
    int32_t integrator;      /* 0 = IRK Gauss-Legendre-4 closed form (parity), 1 = RK4 */
    int32_t soft;            /* robot_ocp_problem.py:106 `slack` */
    int32_t qp_iter_max;     /* QP_ITER, world_specification.py:48 */
    int32_t lm_scaled_by_dt; /* acados-version switch, SURVEY Appendix A */
    int32_t pred_corr, cond_pred_corr, itref_corr_max;
    int32_t lq_fact;         /* HPIPM lq_fact: 0 Cholesky only, 1 switch to the LQ factorisation when the predictor is
                                inaccurate (mode BALANCE, what acados installs), 2 LQ always; -1 = analysis only: classical
                                (non-square-root) Riccati recursion */
    double tf;               /* TF, world_specification.py:43 */
    double lm;               /* levenberg_marquardt, robot_ocp_problem.py:128 */
    double W[6];             /* diag of W over y=[x,y,v,omega,u_a,u_alpha], :78-80 */
    double We[4];            /* diag of W_e over [x,y,v,omega], :83 */
    double lbx[4], ubx[4];   /* on state idx {0,1,3,4}, :91-93 */
    double lbu[2], ubu[2];   /* :95-97 */
    double r2;               /* (R_OBST+R_ROBOT+MARGIN)^2, robot_model.py:62 */
    double lh, uh;           /* :99-103 */
    double zu, Zu;           /* upper-side slack penalties (0, :118-122) */
    double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max;
    double reg_prim, lam_min, t_min, tau_min, thr0;
} orc_config;

void orc_default_config(orc_config *c, int N, int M, double tf);

/* ---- model (src/models/robot_model.py) ---------------------------------- */
void orc_ode(const double x[5], const double u[2], double f[5]);        /* :39-43 */
void orc_gl4_step(const double x[5], const double u[2], double h,
                  double xn[5], double A[25], double B[10]);            /* IRK GL4, closed form */
void orc_rk4_step(const double x[5], const double u[2], double h,
                  double xn[5], double A[25], double B[10]);
void orc_con_h(const double x[5], const double *p, int M, double r2,
               double *h, double *cx, double *cy);                      /* :60-65 */

/* ---- one SQP_RTI solve (robot_ocp_problem.py:195) ----------------------- */
/* Per-problem arrays, stage-major:
 *  x[(N+1)*5] u[N*2] sl[(N+1)*M] su[(N+1)*M]  : iterate, updated in place
 *  x0bar[5]  p[(N+1)*2M]  yref[(N+1)*6]  zl[(N+1)*M]  Zl[(N+1)*M] : inputs
 *  pi[N*5]  (multipliers of the dynamics)  -- output, may be NULL
 *  lam/t[(N+1)*(12+4M)] in HPIPM order [lb lg ub ug ls us] padded per node -- may be NULL
 *  stats[8]: res_g,res_b,res_d,res_m (inf norms at exit), mu, alpha, n_itref, n_lq_flag
 * returns acados status (0 ok, 4 QP failure); *qp_iter = HPIPM iteration count,
 * *qp_status = HPIPM status (0 ok, 1 max-iter, 2 min-step, 3 NaN) */
int orc_rti_solve(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  double *x, double *u, double *sl, double *su,
                  double *pi, double *lam, double *t,
                  int32_t *qp_iter, int32_t *qp_status, double *stats);

/* batch of independent solves, arrays problem-major (b*len + i); OpenMP over b */
void orc_rti_solve_batch(const orc_config *c, int64_t B, int nthreads,
                         const double *x0bar, const double *p, const double *yref,
                         const double *zl, const double *Zl,
                         double *x, double *u, double *sl, double *su,
                         int32_t *status, int32_t *qp_iter, double *stats);

void orc_rti_solve_batch_duals(const orc_config *c, int64_t B, int nthreads,
                               const double *x0bar, const double *p, const double *yref,
                               const double *zl, const double *Zl,
                               double *x, double *u, double *sl, double *su,
                               double *pi, double *lam, double *t,
                               int32_t *status, int32_t *qp_iter, double *stats);

/* ---- full SQP mode (north_star kernel 4; the reference itself runs SQP_RTI only, :130) ---- */
/* res[4] = inf norms of [stationarity, dynamics gap, g_i - t_i (incl. |x0bar - x_0|), lam_i t_i] */
void orc_nlp_residuals(const orc_config *c, const double *x0bar, const double *p,
                       const double *yref, const double *zl, const double *Zl,
                       const double *x, const double *u, const double *sl, const double *su,
                       const double *pi, const double *lam, const double *t, double res[4]);
/* acados `SQP`: [residuals -> test against tol[4] -> QP -> step] up to max_iter QPs; globalization 0 = FIXED_STEP
 * (full steps), 1 = MERIT_BACKTRACKING (alpha *= 0.7 while the l1 merit function does not decrease, alpha_min 0.05).
 * pi/lam/t in/out (previous multipliers or zeros).  Returns 0 converged, 2 max_iter, 4 QP failure. */
int orc_sqp_solve(const orc_config *c, int max_iter, const double tol[4], int globalization,
                  const double ls[3] /* alpha_min, alpha_reduction, eps (accept when merit(a) < merit(0) + eps max(1,|merit(0)|)); NULL = 0.05, 0.7, 0 */,
                  const double *x0bar, const double *p, const double *yref, const double *zl, const double *Zl,
                  double *x, double *u, double *sl, double *su, double *pi, double *lam, double *t,
                  int32_t *sqp_iter, int32_t *qp_iter, double *res, double *stats);
void orc_sqp_solve_batch(const orc_config *c, int64_t B, int nthreads, int max_iter, const double tol[4], int globalization,
                         const double ls[3],
                         const double *x0bar, const double *p, const double *yref, const double *zl, const double *Zl,
                         double *x, double *u, double *sl, double *su, double *pi, double *lam, double *t,
                         int32_t *status, int32_t *sqp_iter, int32_t *qp_iter, double *res,
                         int32_t *qp_maxed /* per problem: QPs stopped by qp_iter_max (may be NULL) */);

/* ---- QP export for validation (tests only) ------------------------------ */
/* Builds the RTI QP of one problem and returns it densely so that numpy can
 * assemble the KKT system independently. Layout documented in pyoracle.py. */
int orc_export_qp(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  const double *x, const double *u, const double *sl, const double *su,
                  double *BAt, double *b, double *H, double *g, double *Cxy,
                  double *lb, double *ub, double *lg, double *ug,
                  double *ls, double *us, double *Zl_qp, double *Zu_qp,
                  double *zl_qp, double *zu_qp);
/* Solves that QP and returns the QP solution (step), for KKT checks. */
int orc_qp_solve_export(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  const double *x, const double *u, const double *sl, const double *su,
                  double *v, double *dsl, double *dsu, double *pi, double *lam, double *t,
                  int32_t *qp_iter, int32_t *qp_status, double *stats);

/* ---- glue around the solve --------------------------------------------- */
void orc_slack_weights(const orc_config *c, const double x0[5], const double goal[2],
                       double *zl, double *Zl);                         /* :145-152 */
void orc_shift(const orc_config *c, double *x, double *u);              /* :253-258 */
void orc_initial_guess(const orc_config *c, double x0[5], double *x, double *u,
                       double *sl, double *su);                         /* :286-306 */
/* obstacle model, src/utils/visualization.py:25-79; obst = [x,y,vx,vy] */
void orc_obst_predict_step(double o[4], double dt, double xmin, double xmax,
                           double ymin, double ymax);                   /* :25-60, noise=False */
void orc_obst_step(double o[4], double dt, int random_move, double randomness, double vmax,
                   const double noise[2], double xmin, double xmax, double ymin, double ymax);
void orc_obst_predict_traj(const double o[4], int n, double dt, int quirk_vx_is_vy,
                           double xmin, double xmax, double ymin, double ymax,
                           double *traj /* (n+1)*2 */);                 /* :62-79 */

#ifdef __cplusplus
}
#endif
#endif
