/*
 * b200mpc.h -- C-ABI of libb200mpc.so: batched FP64 SQP-RTI nonlinear MPC on B200 (sm_100a).
 *
 * Drop-in boundary for ONE hot path of abdelhakim96/Dynamic-Obstacle-Avoidance-MPC: the
 * acados solver object driven from src/simulation/robot_ocp_problem.py.  Each entry point
 * names the reference call it replaces (file:line, relative to the reference repo); the
 * ctypes binding a maintainer would add is shown in INTEGRATION.md and implemented in
 * dynamic-obstacle-avoidance-mpc_b200/b200mpc/.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  `stream` is a cudaStream_t passed as void*
 *     (e.g. torch.cuda.current_stream().cuda_stream); 0 = legacy default stream.
 *   - device arrays are FP64, contiguous, "batch-last" structure-of-arrays: field[len][batch]
 *     (B2MPC_LAYOUT_SOA), or batch-first field[batch][len] (B2MPC_LAYOUT_AOS, transposed on
 *     device).  Host variants (`*_host`) take batch-first host memory (what numpy users hold)
 *     and copy on `stream` straight between the caller's buffer and a device staging buffer of the
 *     handle.  b2mpc_set_host does NOT synchronise: a pinned source buffer must stay alive and unchanged
 *     until the stream has passed the copy; the get_*_host variants and b2mpc_solve_host synchronise.
 *   - every entry point runs on the handle's device (cfg.device) whatever device is current in the
 *     calling thread and restores the caller's current device before returning: handles on different
 *     GPUs can be used from one thread.
 *   - the handle owns all device memory (sized by max_batch at create; no allocation later);
 *     callers own what they pass; pointers are borrowed for the stream-ordered call only.
 *   - every function returns 0 on success or a negative B2MPC_E* code, message via
 *     b2mpc_last_error().  Per-problem numerical status uses acados codes (0 ok, 4 QP failure).
 *   - a handle is not thread-safe; no call synchronises the stream unless documented.
 *   - there is NO CPU fallback: without a CUDA device b2mpc_create fails.
 *   - obstacle constraints are SOFT on both sides (Jsh = I, robot_ocp_problem.py:106-122, the `slack=True` every experiment of
 *     the reference uses).  The hard-constraint variant (`RobotOcpProblem(..., slack=False)`, :14) is not part of the
 *     accelerated path: the Python mirror raises NotImplementedError for it.
 */
#ifndef B200MPC_H
#define B200MPC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2MPC_ABI_VERSION 3
#define B2MPC_MAX_OBST 32
#define B2MPC_MAX_N 128

typedef struct b2mpc b2mpc;   /* opaque; one per (device, max_batch, dims) */

/* Everything robot_ocp_problem.py:54-132 + world_specification.py:7-48 freeze when the
 * reference constructs AcadosOcpSolver (:135), plus acados' HPIPM option block. */
typedef struct b2mpc_config {
    int32_t N;               /* ocp.dims.N = N_SOLV, robot_ocp_problem.py:57 */
    int32_t n_obst;          /* N_OBST, world_specification.py:25; 1..B2MPC_MAX_OBST */
    int32_t device;          /* CUDA ordinal */
    int32_t max_batch;       /* independent problems per solve */
    int32_t integrator;      /* 0 = IRK Gauss-Legendre-4 (reference, :129; the closed form of the collocation system, which
                                the model generator derives AND checks against the IRK equations: codegen/gen_model.py
                                verify_collocation; a model whose ODE is not triangular is rejected at generation time --
                                there is no generic on-device IRK-Newton), 1 = RK4 */
    int32_t qp_iter_max;     /* solver_options.qp_solver_iter_max, :131 */
    int32_t lm_scaled_by_dt; /* acados-version switch for levenberg_marquardt scaling */
    int32_t kernel;          /* 0 = auto (2; batches that fit the GPU in one wave: 3), 2 = warp-per-problem with the node-private
                                rows in L2-resident global memory (any N <= 128), 3 = warp-per-problem, everything in shared
                                memory (needs the whole problem to fit: N = 20, M = 5 does), 4 = CTA-per-problem (mpc_cta.cuh:
                                five warps per problem, everything in shared memory, the latency-oriented mapping; reference
                                dimensions only: 5 obstacles, N = 20 or 10; SQP_RTI only).  1 (a thread-per-problem
                                cross-check kernel) was removed in ABI 3: EINVAL */
    int32_t pred_corr, cond_pred_corr, itref_corr_max, reserved;
    double tf;               /* solver_options.tf, :132 */
    double lm;               /* solver_options.levenberg_marquardt, :128 */
    double W[6];             /* diag(cost.W) over y=[x,y,v,omega,u_a,u_alpha], :78-80 */
    double W_e[4];           /* diag(cost.W_e) over [x,y,v,omega], :83 */
    double lbx[4], ubx[4];   /* constraints.lbx/ubx on idxbx={0,1,3,4}, :91-93 */
    double lbu[2], ubu[2];   /* constraints.lbu/ubu, :95-97 */
    double r2;               /* (R_OBST+R_ROBOT+MARGIN)^2 inside con_h_expr, robot_model.py:62 */
    double lh, uh;           /* constraints.lh/uh (= lh_e/uh_e), :99-103 */
    double zu, Zu;           /* cost.zu/Zu (upper-side slack penalty), :118-122 */
    double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max;   /* HPIPM */
    double reg_prim, lam_min, t_min, tau_min, thr0;                      /* HPIPM */
    /* outer loop of b2mpc_solve: 0 = 'SQP_RTI' (the reference, solver_options.nlp_solver_type, :130),
     * 1 = acados 'SQP' (full steps until the NLP residuals pass nlp_tol or nlp_iter_max QPs were solved) */
    int32_t nlp_solver;
    int32_t nlp_iter_max;    /* solver_options.nlp_solver_max_iter (acados default 100); SQP only */
    double nlp_tol[4];       /* nlp_solver_tol_stat / _eq / _ineq / _comp (acados default 1e-6 each); SQP only */
    int32_t nlp_globalization; /* solver_options.globalization: 0 = 'FIXED_STEP' (full steps, acados default),
                                  1 = 'MERIT_BACKTRACKING' (l1 merit function, alpha <- alpha * ls_alpha_reduction while it
                                  does not decrease, primal and dual variables move by alpha); SQP only */
    int32_t reserved2;
    double ls_alpha_min, ls_alpha_reduction;   /* alpha_min (0.05), alpha_reduction (0.7) */
    double ls_eps;           /* a step is accepted when merit(alpha) < merit(0) + ls_eps max(1, |merit(0)|); default 0 = the
                                plain decrease test.  Near convergence the true decrease (~|step|^2) sinks below the rounding
                                noise of the merit sum (slack penalties of 1e5..1e7): see DESIGN.md "line search" */
} b2mpc_config;

/* fields of b2mpc_set / b2mpc_get; len = doubles per problem */
enum {
    B2MPC_X0 = 0,     /* [5]            set(0,'lbx'/'ubx',x0), :191-192                     */
    B2MPC_P = 1,      /* [(N+1)*2M]     set(i,'p',P[i]) / set_params_sparse, :164-166        */
    B2MPC_YREF = 2,   /* [(N+1)*6]      cost_set(i,'yref',..), :70,77,284 (node N: first 4)  */
    B2MPC_ZL = 3,     /* [(N+1)*M]      cost_set(i,'zl',..), :151                            */
    B2MPC_ZZL = 4,    /* [(N+1)*M]      cost_set(i,'Zl',..), :152                            */
    B2MPC_X = 5,      /* [(N+1)*5]      set/get(i,'x'), :232,254,257,306                     */
    B2MPC_U = 6,      /* [N*2]          set/get(i,'u'), :140,198,255,258,305                 */
    B2MPC_SL = 7,     /* [(N+1)*M]      get(i,'sl') (lower slack iterate; reset() zeroes)    */
    B2MPC_SU = 8,     /* [(N+1)*M]      get(i,'su')                                          */
    B2MPC_PI = 9,     /* [N*5]          get(i,'pi')  (output; SQP: in/out, reset() zeroes)   */
    B2MPC_LAM = 10,   /* [(N+1)*(12+4M)] get(i,'lam') padded [lb6 lg ub6 ug ls us] (same)    */
    B2MPC_T = 11,     /* [(N+1)*(12+4M)] get(i,'t')   (same)                                 */
    B2MPC_STATS = 12, /* [8] res_g,res_b,res_d,res_m,mu,alpha,n_itref,0 (SQP: last line-search step) */
    B2MPC_GOAL = 13,  /* [2]            sub-goal used by b2mpc_prepare_slack / closed loop   */
    B2MPC_OBST = 14,  /* [M*4]          obstacle states (x,y,vx,vy), visualization.py:10-18  */
    B2MPC_MARGIN = 15, /* [2]           closed loop: min_margin_traj, distance_to_goal (:222-228,:246) */
    B2MPC_U0 = 16,    /* [2]            control applied by the last b2mpc_loop_post (get(0,'u'), :198) */
    B2MPC_NLP_RES = 17, /* [4]          SQP: res_stat,res_eq,res_ineq,res_comp at exit (get_residuals()) */
    B2MPC_N_FIELDS = 18
};
enum { B2MPC_STATUS = 0, B2MPC_QP_ITER = 1, B2MPC_QP_STATUS = 2,
       B2MPC_STEPS = 3,   /* closed loop: control steps taken (i of :184-260) */
       B2MPC_DONE = 4,    /* closed loop: reached the goal tolerance (reached_subgoal, :247-250) */
       B2MPC_OOB = 5,     /* closed loop: left the arena at least once (:213-214) */
       B2MPC_NFAIL = 6,   /* closed loop: number of status-4 solves (:203) */
       B2MPC_SQP_ITER = 7, /* SQP: QPs solved by the last b2mpc_solve (get_stats('sqp_iter')); QP_ITER then holds their IPM total */
       B2MPC_N_IFIELDS = 8 };
enum { B2MPC_LAYOUT_SOA = 0, B2MPC_LAYOUT_AOS = 1 };

enum {
    B2MPC_OK = 0, B2MPC_EINVAL = -1, B2MPC_ENODEV = -2, B2MPC_ENOMEM = -3,
    B2MPC_ECUDA = -4, B2MPC_EBATCH = -5
};

int b2mpc_abi_version(void);
const char *b2mpc_last_error(void);                              /* thread-local */

/* reference defaults: world_specification.py + robot_ocp_problem.py:24-27,91-132 */
void b2mpc_default_config(b2mpc_config *cfg, int32_t N, int32_t n_obst, double tf,
                          int32_t max_batch, int32_t device);

/* AcadosOcpSolver(ocp, json_file=...) + AcadosSimSolver(...), robot_ocp_problem.py:135-136 */
int b2mpc_create(const b2mpc_config *cfg, b2mpc **out);
int b2mpc_destroy(b2mpc *h);                                     /* solver teardown */
int64_t b2mpc_field_len(const b2mpc *h, int field);              /* doubles per problem */
int64_t b2mpc_workspace_bytes(const b2mpc *h);                   /* device scratch of the solve kernel (node-private rows / work arrays) */

/* ocp_solver.set / cost_set / set_params_sparse (:151-152,:164-166,:191-192,:254-258,:305-306) */
int b2mpc_set(b2mpc *h, int field, const double *dptr, int64_t batch, int layout, void *stream);
/* ocp_solver.get(i, 'u'/'x') (:198,:232,:240) and multipliers/statistics */
int b2mpc_get(b2mpc *h, int field, double *dptr, int64_t batch, int layout, void *stream);
int b2mpc_get_i32(b2mpc *h, int ifield, int32_t *dptr, int64_t batch, void *stream);
/* same, host memory, batch-first; H2D/D2H on `stream`; get_* synchronise the stream */
int b2mpc_set_host(b2mpc *h, int field, const double *hptr, int64_t batch, void *stream);
int b2mpc_get_host(b2mpc *h, int field, double *hptr, int64_t batch, void *stream);
int b2mpc_get_i32_host(b2mpc *h, int ifield, int32_t *hptr, int64_t batch, void *stream);
/* raw device pointer of a handle-owned field (SoA, row stride = max_batch) for zero-copy views */
double *b2mpc_field_ptr(b2mpc *h, int field);

/* parameterize_slack(), robot_ocp_problem.py:145-152: ZL,ZZL <- f(X0, GOAL) on device */
int b2mpc_prepare_slack(b2mpc *h, int64_t batch, void *stream);
/* ocp_solver.reset() + set_initial_guess(), :286-306 (mask: uint8 per problem, NULL = all);
 * like the reference it zeroes v,omega of X0 itself (aliasing quirk :301-302) */
int b2mpc_reset(b2mpc *h, const uint8_t *dmask, int64_t batch, void *stream);
/* ocp_solver.solve(), :195 -- one SQP_RTI iteration for every problem (cfg.nlp_solver = 0); fills STATUS, QP_ITER.
 * With cfg.nlp_solver = 1 the same launch runs acados' full SQP loop per problem (status 0 converged,
 * 2 max_iter, 4 QP failure; also fills SQP_ITER, NLP_RES). */
int b2mpc_solve(b2mpc *h, int64_t batch, void *stream);
/* full SQP on a handle of either kind (solver.options_set('max_iter'/'tol_*') + solve()): max_iter <= 0 and
 * tol = NULL take the handle's configuration. */
int b2mpc_solve_sqp(b2mpc *h, int64_t batch, int32_t max_iter, const double *tol4, void *stream);
/* ocp_integrator.set/solve/get, :207-212: X0 <- F(X0, U[0]) */
int b2mpc_sim(b2mpc *h, int64_t batch, void *stream);
/* warm-start shift, :253-258 */
int b2mpc_shift(b2mpc *h, int64_t batch, void *stream);
/* ---- closed loop around the solve: everything of robot_ocp_problem.py:184-260
 * that is not the solve itself, on device ---- */
/* parameterize_model(), :154-166 + Obstacle.predict_trajectory, visualization.py:62-79:
 * P <- predicted obstacle positions over the horizon from OBST; quirk_vx_is_vy reproduces :69 */
int b2mpc_obst_predict(b2mpc *h, int64_t batch, int32_t quirk_vx_is_vy, void *stream);
/* zero the closed-loop bookkeeping (init_experiment, :35-52) */
int b2mpc_loop_reset(b2mpc *h, int64_t batch, void *stream);
/* head of RobotOcpProblem.step(), :177-183: reached_subgoal = False, out_of_bounds = False, i = 0 -- clears DONE, OOB
 * and STEPS so that set_subgoal() followed by step() moves an environment that had reached its previous sub-goal;
 * the episode's min_margin_traj (MARGIN[0]) and NFAIL are kept */
int b2mpc_loop_begin(b2mpc *h, int64_t batch, void *stream);
/* the rest of one control step after b2mpc_solve, :198-260: read u0, on status 4 re-initialise the
 * guess (init_guess_when_error, incl. the x0 aliasing quirk), plant step, out-of-bounds flag, obstacle
 * step with noise[batch][M][2] ~ N(0,1) (device, batch-first; NULL or random_move=0: no noise),
 * margins, goal test with tolerance `tol`, warm-start shift.  Finished problems are frozen. */
int b2mpc_loop_post(b2mpc *h, int64_t batch, const double *dnoise, int32_t random_move,
                    int32_t init_guess_when_error, double tol, void *stream);

/* The whole `while i < max_iter` loop of RobotOcpProblem.step() (robot_ocp_problem.py:184-260) for `steps` control steps of
 * every environment, enqueued on `stream` by ONE call (per control step: queue scatter, solve, and one glue kernel that
 * finishes the step -- b2mpc_loop_post -- and starts the next -- b2mpc_obst_predict + b2mpc_prepare_slack).  Environments
 * that have reached their goal leave the queue.  Velocity noise of the obstacles (visualization.py:28-33):
 *   noise_mode 0  none;
 *   noise_mode 1  on-device Philox4x32-10 + Box-Muller stream keyed by (seed, env_offset + environment, control step since
 *                 the last b2mpc_loop_reset, obstacle): the same for any batch size or sharding of the environments;
 *   noise_mode 2  replay of dnoise[steps][batch][M][2] ~ N(0,1) (device, batch-first), e.g. the streams
 *                 `np.random.seed(i)` gives the reference's episodes.
 * random_move = 0 switches the noise off whatever the mode.  Does not synchronise. */
int b2mpc_closed_loop(b2mpc *h, int64_t batch, int32_t steps, int32_t noise_mode, uint64_t seed, int64_t env_offset,
                      const double *dnoise, int32_t random_move, int32_t init_guess_when_error,
                      int32_t quirk_vx_is_vy, double tol, void *stream);

/* which solve kernel the handle resolved to: 2 = warp-per-problem (node-private rows in global memory),
   3 = warp-per-problem (all shared memory), 4 = CTA-per-problem */
int b2mpc_kernel(const b2mpc *h);
/* number of kernels launched by this handle so far (for bench accounting) */
int64_t b2mpc_launch_count(const b2mpc *h);

/* whole call a user of the reference makes per control step, on HOST batch-first buffers:
 * set x0/p/yref/zl/Zl + warm iterate (x,u,sl,su) -> solve -> read back iterate, status, qp_iter.
 * Synchronises `stream` before returning. */
int b2mpc_solve_host(b2mpc *h, int64_t batch, const double *x0, const double *p, const double *yref,
                     const double *zl, const double *Zl, double *x, double *u, double *sl, double *su,
                     int32_t *status, int32_t *qp_iter, void *stream);

/* FP64 FMA-pipe microbenchmark (register-resident DFMA chains) -> achieved TFLOP/s; the
 * roofline denominator bench.py quotes, since MEASURED_PEAKS.json has no FP64 entry */
int b2mpc_dfma_peak(int device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif
