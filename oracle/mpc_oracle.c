/*
 * mpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; PARITY UNPINNED,
 * see mpc_oracle.h).  Restates, in plain dense C the way acados/HPIPM/BLASFEO would
 * execute it, the solve behind src/simulation/robot_ocp_problem.py:195:
 *
 *   preparation  : IRK Gauss-Legendre-4 discretisation + forward sensitivities of
 *                  src/models/robot_model.py:39-43, LINEAR_LS Gauss-Newton cost of
 *                  robot_ocp_problem.py:59-83, BGH constraints :85-122
 *   feedback     : x0 embedding and elimination, HPIPM-style Mehrotra predictor-corrector
 *                  interior point with a square-root Riccati factorisation
 *   update       : full step (SQP_RTI, :130)
 *
 * Deliberately generic/dense inside the QP solver (per-stage dims nu,nx,nb,ng,ns, dense
 * Cholesky) -- the CUDA product exploits the model structure instead, so the two are
 * independent implementations of the same mathematics.
 */
#include "mpc_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>
#include <stdatomic.h>

#define NX ORC_NX
#define NU ORC_NU
#define NV ORC_NV
#define MM ORC_MAXM
#define MAXC (12 + 4 * ORC_MAXM)

/* ------------------------------------------------------------------ config */
void orc_default_config(orc_config *c, int N, int M, double tf)
{
    memset(c, 0, sizeof(*c));
    c->N = N; c->M = M; c->integrator = 0; c->soft = 1;
    c->qp_iter_max = 50;                      /* world_specification.py:48 */
    c->lm_scaled_by_dt = 1;   /* see DESIGN.md "Levenberg-Marquardt scaling": fixtures favour Ts*lm */
    c->pred_corr = 1; c->cond_pred_corr = 1; c->itref_corr_max = 2;
    c->tf = tf; c->lm = 2.0;                  /* robot_ocp_problem.py:128,132 */
    /* W = blkdiag(Q=2*I4, R=0.15*I2), robot_ocp_problem.py:24-26,78-80 */
    for (int i = 0; i < 4; i++) c->W[i] = 2.0;
    c->W[4] = c->W[5] = 0.15;
    for (int i = 0; i < 4; i++) c->We[i] = 5.0;   /* :27,83 */
    c->lbx[0] = c->lbx[1] = -7.0; c->lbx[2] = c->lbx[3] = -10.0;  /* :91 */
    c->ubx[0] = c->ubx[1] = 7.0;  c->ubx[2] = c->ubx[3] = 10.0;   /* :92 */
    c->lbu[0] = c->lbu[1] = -8.0; c->ubu[0] = c->ubu[1] = 8.0;    /* :95-96 */
    c->r2 = (1.0 + 0.2 + 1.2) * (1.0 + 0.2 + 1.2);               /* robot_model.py:62 */
    c->lh = 0.0; c->uh = 1e5;                                    /* :99-103 */
    c->zu = 0.0; c->Zu = 0.0;                                    /* :118-122 */
    /* acados' HPIPM option block (mode BALANCE + overrides), SURVEY App. C */
    c->mu0 = 1e0; c->alpha_min = 1e-8;
    c->res_g_max = 1e-6; c->res_b_max = 1e-8; c->res_d_max = 1e-8; c->res_m_max = 1e-8;
    c->reg_prim = 1e-15; c->lam_min = 1e-16; c->t_min = 1e-16; c->tau_min = 1e-16;
    c->thr0 = 1e-1;
}

/* ------------------------------------------------------------------- model */
/* robot_model.py:39-43 */
void orc_ode(const double x[5], const double u[2], double f[5])
{
    f[0] = x[3] * cos(x[2]);
    f[1] = x[3] * sin(x[2]);
    f[2] = x[4];
    f[3] = u[0];
    f[4] = u[1];
}

/* 4-stage Gauss-Legendre nodes/weights on [0,1] (acados IRK default) */
static void gl4_nodes(double c[4], double b[4])
{
    double s30 = sqrt(30.0), s65 = sqrt(6.0 / 5.0);
    double ci = sqrt(3.0 / 7.0 - 2.0 / 7.0 * s65), co = sqrt(3.0 / 7.0 + 2.0 / 7.0 * s65);
    c[0] = 0.5 * (1.0 - co); c[1] = 0.5 * (1.0 - ci); c[2] = 0.5 * (1.0 + ci); c[3] = 0.5 * (1.0 + co);
    b[0] = b[3] = (18.0 - s30) / 72.0;
    b[1] = b[2] = (18.0 + s30) / 72.0;
}

/* One IRK-GL4 step with forward sensitivities.  For this model the collocation system
 * is triangular (v,omega affine in t, psi quadratic, x,y pure quadratures), so the
 * Newton iteration acados runs converges to this closed form (proven
 * numerically in tests/test_oracle_model.py against oracle/irk_generic.py). */
void orc_gl4_step(const double x[5], const double u[2], double h,
                  double xn[5], double A[25], double B[10])
{
    double c[4], b[4];
    gl4_nodes(c, b);
    double sx = 0, sy = 0;                       /* sum b v cos, sum b v sin */
    double dxp = 0, dxv = 0, dxw = 0, dxa = 0, dxal = 0;
    double dyp = 0, dyv = 0, dyw = 0, dya = 0, dyal = 0;
    for (int i = 0; i < 4; i++) {
        double ti = h * c[i];
        double psi = x[2] + x[4] * ti + 0.5 * u[1] * ti * ti;
        double v = x[3] + u[0] * ti;
        double cs = cos(psi), sn = sin(psi);
        sx += b[i] * v * cs;
        sy += b[i] * v * sn;
        dxp += -b[i] * v * sn;            dyp += b[i] * v * cs;
        dxv += b[i] * cs;                 dyv += b[i] * sn;
        dxw += -b[i] * v * sn * ti;       dyw += b[i] * v * cs * ti;
        dxa += b[i] * ti * cs;            dya += b[i] * ti * sn;
        dxal += -b[i] * v * sn * 0.5 * ti * ti;  dyal += b[i] * v * cs * 0.5 * ti * ti;
    }
    xn[0] = x[0] + h * sx;
    xn[1] = x[1] + h * sy;
    xn[2] = x[2] + h * x[4] + 0.5 * h * h * u[1];
    xn[3] = x[3] + h * u[0];
    xn[4] = x[4] + h * u[1];
    if (A) {
        memset(A, 0, 25 * sizeof(double));
        for (int i = 0; i < 5; i++) A[i * 5 + i] = 1.0;
        A[0 * 5 + 2] = h * dxp; A[0 * 5 + 3] = h * dxv; A[0 * 5 + 4] = h * dxw;
        A[1 * 5 + 2] = h * dyp; A[1 * 5 + 3] = h * dyv; A[1 * 5 + 4] = h * dyw;
        A[2 * 5 + 4] = h;
    }
    if (B) {
        memset(B, 0, 10 * sizeof(double));
        B[0 * 2 + 0] = h * dxa; B[0 * 2 + 1] = h * dxal;
        B[1 * 2 + 0] = h * dya; B[1 * 2 + 1] = h * dyal;
        B[2 * 2 + 1] = 0.5 * h * h;
        B[3 * 2 + 0] = h;
        B[4 * 2 + 1] = h;
    }
}

/* df/dx (5x5) and df/du (5x2) of robot_model.py:39-43 */
static void ode_jac(const double x[5], double fx[25], double fu[10])
{
    memset(fx, 0, 25 * sizeof(double));
    memset(fu, 0, 10 * sizeof(double));
    double cs = cos(x[2]), sn = sin(x[2]);
    fx[0 * 5 + 2] = -x[3] * sn; fx[0 * 5 + 3] = cs;
    fx[1 * 5 + 2] = x[3] * cs;  fx[1 * 5 + 3] = sn;
    fx[2 * 5 + 4] = 1.0;
    fu[3 * 2 + 0] = 1.0;
    fu[4 * 2 + 1] = 1.0;
}

/* classical RK4, one step, with forward sensitivities (north_star's named alternative;
 * NOT what the reference runs -- robot_ocp_problem.py:129 sets IRK) */
void orc_rk4_step(const double x[5], const double u[2], double h,
                  double xn[5], double A[25], double B[10])
{
    /* augmented state: x (5), Sx (25), Su (10) */
    double k[4][5], Kx[4][25], Ku[4][10];
    double xs[5], Sx[25], Su[10];
    const double cc[4] = {0.0, 0.5, 0.5, 1.0};
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < 5; i++) xs[i] = x[i] + (s ? cc[s] * h * k[s - 1][i] : 0.0);
        for (int i = 0; i < 25; i++) Sx[i] = ((i / 5 == i % 5) ? 1.0 : 0.0) + (s ? cc[s] * h * Kx[s - 1][i] : 0.0);
        for (int i = 0; i < 10; i++) Su[i] = (s ? cc[s] * h * Ku[s - 1][i] : 0.0);
        double fx[25], fu[10];
        orc_ode(xs, u, k[s]);
        ode_jac(xs, fx, fu);
        for (int r = 0; r < 5; r++) {
            for (int cidx = 0; cidx < 5; cidx++) {
                double a = 0;
                for (int q = 0; q < 5; q++) a += fx[r * 5 + q] * Sx[q * 5 + cidx];
                Kx[s][r * 5 + cidx] = a;
            }
            for (int cidx = 0; cidx < 2; cidx++) {
                double a = fu[r * 2 + cidx];
                for (int q = 0; q < 5; q++) a += fx[r * 5 + q] * Su[q * 2 + cidx];
                Ku[s][r * 2 + cidx] = a;
            }
        }
    }
    for (int i = 0; i < 5; i++) xn[i] = x[i] + h / 6.0 * (k[0][i] + 2 * k[1][i] + 2 * k[2][i] + k[3][i]);
    if (A) for (int i = 0; i < 25; i++)
        A[i] = ((i / 5 == i % 5) ? 1.0 : 0.0) + h / 6.0 * (Kx[0][i] + 2 * Kx[1][i] + 2 * Kx[2][i] + Kx[3][i]);
    if (B) for (int i = 0; i < 10; i++)
        B[i] = h / 6.0 * (Ku[0][i] + 2 * Ku[1][i] + 2 * Ku[2][i] + Ku[3][i]);
}

/* robot_model.py:60-65: h_j = (x-p_2j)^2 + (y-p_2j+1)^2 - r^2 ; gradient on (x,y) */
void orc_con_h(const double x[5], const double *p, int M, double r2,
               double *h, double *cx, double *cy)
{
    for (int j = 0; j < M; j++) {
        double dx = x[0] - p[2 * j], dy = x[1] - p[2 * j + 1];
        h[j] = dx * dx + dy * dy - r2;
        if (cx) cx[j] = 2.0 * dx;
        if (cy) cy[j] = 2.0 * dy;
    }
}

/* ------------------------------------------------------------- QP container */
typedef struct {
    int nu, nx, nv, nb, ng, ns, nxn;
    int idxb[6];
    double BAt[NV][NX];      /* row i = d x+ / d v_i, v = [u; x] */
    double b[NX];
    double H[NV][NV];
    double g[NV];
    double Ct[NV][MM];       /* column j = gradient of general constraint j w.r.t. v */
    double lb[6], ub[6];
    double lg[MM], ug[MM], ls[MM], us[MM];
    double Zl[MM], Zu[MM], zl[MM], zu[MM];
} stage_qp;

typedef struct {
    double v[NV], sl[MM], su[MM], pi[NX], lam[MAXC], t[MAXC];
} stage_sol;

typedef struct {
    double rg[NV], rgsl[MM], rgsu[MM], rb[NX], rd[MAXC], rm[MAXC];
} stage_res;

typedef struct {
    double L[NV][NV];        /* Cholesky factor of the Riccati matrix of this stage */
    double p[NX];
    double lu[NU];
    double tinv[MAXC], Gam[MAXC];
    double Dl[MM], Du[MM];
    double rhol[MM], rhou[MM];
} stage_ws;

#define OFF_LB(q) 0
#define OFF_LG(q) ((q)->nb)
#define OFF_UB(q) ((q)->nb + (q)->ng)
#define OFF_UG(q) (2 * (q)->nb + (q)->ng)
#define OFF_LS(q) (2 * (q)->nb + 2 * (q)->ng)
#define OFF_US(q) (2 * (q)->nb + 2 * (q)->ng + (q)->ns)
#define NCT(q)    (2 * (q)->nb + 2 * (q)->ng + 2 * (q)->ns)

typedef struct {
    int N;
    stage_qp *qp;
    stage_sol *sol, *step, *corr;
    stage_res *res, *lin;
    stage_ws *ws;
    double *rm_bkp;          /* (N+1)*MAXC */
} ipm_mem;

static ipm_mem *ipm_alloc(int N)
{
    ipm_mem *m = (ipm_mem *)calloc(1, sizeof(ipm_mem));
    m->N = N;
    m->qp = (stage_qp *)calloc(N + 1, sizeof(stage_qp));
    m->sol = (stage_sol *)calloc(N + 1, sizeof(stage_sol));
    m->step = (stage_sol *)calloc(N + 1, sizeof(stage_sol));
    m->corr = (stage_sol *)calloc(N + 1, sizeof(stage_sol));
    m->res = (stage_res *)calloc(N + 1, sizeof(stage_res));
    m->lin = (stage_res *)calloc(N + 1, sizeof(stage_res));
    m->ws = (stage_ws *)calloc(N + 1, sizeof(stage_ws));
    m->rm_bkp = (double *)calloc((size_t)(N + 1) * MAXC, sizeof(double));
    return m;
}
static void ipm_free_real(ipm_mem *m)
{
    free(m->qp); free(m->sol); free(m->step); free(m->corr); free(m->res); free(m->lin);
    free(m->ws); free(m->rm_bkp); free(m);
}
/* one cached workspace per thread (a solver object reuses its memory between solves, as acados does) */
static _Thread_local ipm_mem *tl_mem = NULL;
static ipm_mem *ipm_get(int N)
{
    if (tl_mem && tl_mem->N != N) { ipm_free_real(tl_mem); tl_mem = NULL; }
    if (!tl_mem) tl_mem = ipm_alloc(N);
    return tl_mem;
}
static void ipm_free(ipm_mem *m) { (void)m; }

/* ------------------------------------------------ RTI preparation: build QP */
/* SURVEY App. A/B.  Node 0 is emitted with x0 already eliminated (HPIPM
 * d_ocp_qp_reduce_eq_dof as called by acados' xcond layer): nx_0 = 0, b_0 += A_0 dx0,
 * general-constraint bounds shifted by C_0 dx0. */
static void build_qp(const orc_config *c, const double *x0bar, const double *p,
                     const double *yref, const double *zl, const double *Zl,
                     const double *x, const double *u, const double *sl, const double *su,
                     stage_qp *qp, double dx0[NX])
{
    const int N = c->N, M = c->M;
    const double dt = c->tf / N;
    static const int idxbx[4] = {0, 1, 3, 4};
    for (int i = 0; i < NX; i++) dx0[i] = x0bar[i] - x[i];
    for (int k = 0; k <= N; k++) {
        stage_qp *q = &qp[k];
        memset(q, 0, sizeof(*q));
        const double *xk = x + k * NX;
        const double *uk = u + k * NU;      /* valid for k<N */
        q->nu = (k < N) ? NU : 0;
        q->nx = (k == 0) ? 0 : NX;
        q->nv = q->nu + q->nx;
        q->nxn = (k < N) ? NX : 0;
        q->ng = M;
        q->ns = c->soft ? M : 0;
        const int nu = q->nu;
        const double scal = (k < N) ? dt : 1.0;          /* acados cost time-scaling */
        const double lm = (k < N && c->lm_scaled_by_dt) ? dt * c->lm : c->lm;
        /* ---- dynamics */
        double A[25], Bm[10], xn[5];
        if (k < N) {
            if (c->integrator == 1) orc_rk4_step(xk, uk, dt, xn, A, Bm);
            else orc_gl4_step(xk, uk, dt, xn, A, Bm);
            for (int r = 0; r < NX; r++) {
                q->b[r] = xn[r] - x[(k + 1) * NX + r];
                for (int j = 0; j < NU; j++) q->BAt[j][r] = Bm[r * 2 + j];
                if (k > 0) for (int j = 0; j < NX; j++) q->BAt[nu + j][r] = A[r * 5 + j];
                else for (int j = 0; j < NX; j++) q->b[r] += A[r * 5 + j] * dx0[j];
            }
        }
        /* ---- LINEAR_LS Gauss-Newton cost, robot_ocp_problem.py:59-83 */
        const double *yr = yref + k * ORC_NY;
        if (k < N) {
            double Hx[5] = {c->W[0], c->W[1], 0.0, c->W[2], c->W[3]};
            double gx[5] = {c->W[0] * (xk[0] - yr[0]), c->W[1] * (xk[1] - yr[1]), 0.0,
                            c->W[2] * (xk[3] - yr[2]), c->W[3] * (xk[4] - yr[3])};
            for (int j = 0; j < NU; j++) {
                q->H[j][j] = scal * c->W[4 + j] + lm;
                q->g[j] = scal * c->W[4 + j] * (uk[j] - yr[4 + j]);
            }
            if (k > 0) for (int j = 0; j < NX; j++) {
                q->H[nu + j][nu + j] = scal * Hx[j] + lm;
                q->g[nu + j] = scal * gx[j];
            }
        } else {
            double Hx[5] = {c->We[0], c->We[1], 0.0, c->We[2], c->We[3]};
            double gx[5] = {c->We[0] * (xk[0] - yr[0]), c->We[1] * (xk[1] - yr[1]), 0.0,
                            c->We[2] * (xk[3] - yr[2]), c->We[3] * (xk[4] - yr[3])};
            for (int j = 0; j < NX; j++) {
                q->H[j][j] = Hx[j] + lm;
                q->g[j] = gx[j];
            }
        }
        /* ---- box constraints on the step, :87-97 */
        int nb = 0;
        if (k < N) for (int j = 0; j < NU; j++) {
            q->idxb[nb] = j; q->lb[nb] = c->lbu[j] - uk[j]; q->ub[nb] = c->ubu[j] - uk[j]; nb++;
        }
        if (k > 0 && k < N) for (int j = 0; j < ORC_NBX; j++) {
            q->idxb[nb] = nu + idxbx[j];
            q->lb[nb] = c->lbx[j] - xk[idxbx[j]]; q->ub[nb] = c->ubx[j] - xk[idxbx[j]]; nb++;
        }
        q->nb = nb;
        /* ---- obstacle constraints (soft), robot_model.py:60-65, :99-122 */
        double hh[MM], cx[MM], cy[MM];
        orc_con_h(xk, p + (size_t)k * 2 * M, M, c->r2, hh, cx, cy);
        for (int j = 0; j < M; j++) {
            double s_l = c->soft ? sl[k * M + j] : 0.0, s_u = c->soft ? su[k * M + j] : 0.0;
            q->lg[j] = c->lh - hh[j] - s_l;
            q->ug[j] = c->uh - hh[j] + s_u;
            if (k > 0) { q->Ct[nu + 0][j] = cx[j]; q->Ct[nu + 1][j] = cy[j]; }
            else {
                double cd = cx[j] * dx0[0] + cy[j] * dx0[1];
                q->lg[j] -= cd; q->ug[j] -= cd;
            }
            if (c->soft) {
                q->ls[j] = -s_l; q->us[j] = -s_u;
                q->Zl[j] = scal * Zl[k * M + j];
                q->zl[j] = scal * (zl[k * M + j] + Zl[k * M + j] * s_l);
                q->Zu[j] = scal * c->Zu;
                q->zu[j] = scal * (c->zu + c->Zu * s_u);
            }
        }
    }
}

/* ---------------------------------------------------------- IPM primitives */
static double cdot(const stage_qp *q, int j, const double *v)
{
    double a = 0;
    for (int i = 0; i < q->nv; i++) a += q->Ct[i][j] * v[i];
    return a;
}

/* HPIPM d_ocp_qp_init_var, var_init_scheme 0, cold start */
static void ipm_init(const orc_config *c, ipm_mem *m)
{
    const double thr0 = c->thr0, mu0 = c->mu0;
    for (int k = 0; k <= m->N; k++) {
        const stage_qp *q = &m->qp[k];
        stage_sol *s = &m->sol[k];
        memset(s, 0, sizeof(*s));
        double *lam = s->lam, *t = s->t;
        for (int j = 0; j < q->nb; j++) {
            int iv = q->idxb[j];
            double tl = -q->lb[j] + s->v[iv];
            double tu = q->ub[j] - s->v[iv];
            if (tl < thr0) {
                if (tu < thr0) { s->v[iv] = 0.5 * (q->lb[j] + q->ub[j]); tl = thr0; tu = thr0; }
                else { tl = thr0; s->v[iv] = q->lb[j] + thr0; }
            } else if (tu < thr0) { tu = thr0; s->v[iv] = q->ub[j] - thr0; }
            t[OFF_LB(q) + j] = tl; t[OFF_UB(q) + j] = tu;
            lam[OFF_LB(q) + j] = mu0 / tl; lam[OFF_UB(q) + j] = mu0 / tu;
        }
        for (int j = 0; j < q->ng; j++) {
            double cv = cdot(q, j, s->v);
            double tl = cv - q->lg[j], tu = q->ug[j] - cv;
            tl = thr0 > tl ? thr0 : tl;
            tu = thr0 > tu ? thr0 : tu;
            t[OFF_LG(q) + j] = tl; t[OFF_UG(q) + j] = tu;
            lam[OFF_LG(q) + j] = mu0 / tl; lam[OFF_UG(q) + j] = mu0 / tu;
        }
        for (int j = 0; j < q->ns; j++) {
            t[OFF_LS(q) + j] = 1.0; t[OFF_US(q) + j] = 1.0;
            lam[OFF_LS(q) + j] = mu0; lam[OFF_US(q) + j] = mu0;
        }
    }
}

/* residuals of the KKT conditions (HPIPM d_ocp_qp_res_compute); rd = g_i - t_i */
static void ipm_res(ipm_mem *m, stage_res *res, double *mu, double nrm[4])
{
    const int N = m->N;
    double summ = 0; int nc = 0;
    nrm[0] = nrm[1] = nrm[2] = nrm[3] = 0;
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        stage_res *r = &res[k];
        const int nu = q->nu;
        for (int i = 0; i < q->nv; i++) {
            double a = q->g[i];
            for (int j = 0; j < q->nv; j++) a += q->H[i][j] * s->v[j];
            for (int j = 0; j < q->nxn; j++) a += q->BAt[i][j] * s->pi[j];
            r->rg[i] = a;
        }
        if (k > 0) for (int i = 0; i < q->nx; i++) r->rg[nu + i] -= m->sol[k - 1].pi[i];
        for (int j = 0; j < q->nb; j++) {
            int iv = q->idxb[j];
            r->rg[iv] += -s->lam[OFF_LB(q) + j] + s->lam[OFF_UB(q) + j];
            r->rd[OFF_LB(q) + j] = s->v[iv] - q->lb[j] - s->t[OFF_LB(q) + j];
            r->rd[OFF_UB(q) + j] = q->ub[j] - s->v[iv] - s->t[OFF_UB(q) + j];
        }
        for (int j = 0; j < q->ng; j++) {
            double cv = cdot(q, j, s->v);
            double dl = s->lam[OFF_UG(q) + j] - s->lam[OFF_LG(q) + j];
            for (int i = 0; i < q->nv; i++) r->rg[i] += q->Ct[i][j] * dl;
            double s_l = q->ns ? s->sl[j] : 0.0, s_u = q->ns ? s->su[j] : 0.0;
            r->rd[OFF_LG(q) + j] = cv + s_l - q->lg[j] - s->t[OFF_LG(q) + j];
            r->rd[OFF_UG(q) + j] = q->ug[j] - cv + s_u - s->t[OFF_UG(q) + j];
        }
        for (int j = 0; j < q->ns; j++) {
            r->rgsl[j] = q->Zl[j] * s->sl[j] + q->zl[j] - s->lam[OFF_LG(q) + j] - s->lam[OFF_LS(q) + j];
            r->rgsu[j] = q->Zu[j] * s->su[j] + q->zu[j] - s->lam[OFF_UG(q) + j] - s->lam[OFF_US(q) + j];
            r->rd[OFF_LS(q) + j] = s->sl[j] - q->ls[j] - s->t[OFF_LS(q) + j];
            r->rd[OFF_US(q) + j] = s->su[j] - q->us[j] - s->t[OFF_US(q) + j];
        }
        if (k < N) {
            const stage_sol *sn = &m->sol[k + 1];
            const int nun = m->qp[k + 1].nu;
            for (int j = 0; j < q->nxn; j++) {
                double a = q->b[j] - sn->v[nun + j];
                for (int i = 0; i < q->nv; i++) a += q->BAt[i][j] * s->v[i];
                r->rb[j] = a;
            }
        }
        const int nct = NCT(q);
        for (int i = 0; i < nct; i++) { r->rm[i] = s->lam[i] * s->t[i]; summ += r->rm[i]; }
        nc += nct;
        for (int i = 0; i < q->nv; i++) nrm[0] = fmax(nrm[0], fabs(r->rg[i]));
        for (int j = 0; j < q->ns; j++) { nrm[0] = fmax(nrm[0], fabs(r->rgsl[j])); nrm[0] = fmax(nrm[0], fabs(r->rgsu[j])); }
        for (int j = 0; j < q->nxn; j++) nrm[1] = fmax(nrm[1], fabs(r->rb[j]));
        for (int i = 0; i < nct; i++) { nrm[2] = fmax(nrm[2], fabs(r->rd[i])); nrm[3] = fmax(nrm[3], fabs(r->rm[i])); }
    }
    *mu = summ / nc;
}

/* dense Cholesky, lower, in place on the leading n x n of a NVxNV array */
static void chol_ln(double Lm[NV][NV], int n)
{
    for (int j = 0; j < n; j++) {
        double d = Lm[j][j];
        for (int q = 0; q < j; q++) d -= Lm[j][q] * Lm[j][q];
        d = d > 0.0 ? sqrt(d) : 0.0;            /* BLASFEO dpotrf: non-positive pivot -> 0 */
        Lm[j][j] = d;
        double dinv = d > 0.0 ? 1.0 / d : 0.0;
        for (int i = j + 1; i < n; i++) {
            double a = Lm[i][j];
            for (int q = 0; q < j; q++) a -= Lm[i][q] * Lm[j][q];
            Lm[i][j] = a * dinv;
        }
        for (int i = 0; i < j; i++) Lm[i][j] = 0.0;
    }
}

/* Gamma = lam/t, condensed Hessians, backward square-root Riccati factorisation
 * (HPIPM d_ocp_qp_fact_solve_kkt_step, factorisation half) */
/* LQ factorisation with positive diagonal (BLASFEO dgelqf_pd): W (n x nc, row-major with leading dimension ld) = L Q.
 * Householder reflections from the right, row by row; on return the lower triangle of the leading n x n block of W is L.
 * L L^T = W W^T without ever forming the product: the condition number of the sum of squares is not squared. */
#define LQ_MAXC (NV + 6 + ORC_MAXM + NX)
static void lq_pd(double *W, int n, int nc, int ld)
{
    for (int i = 0; i < n; i++) {
        double *wi = W + (size_t)i * ld;
        double s2 = 0.0;
        for (int j = i + 1; j < nc; j++) s2 += wi[j] * wi[j];
        if (s2 > 0.0) {
            const double x0 = wi[i];
            const double nrmx = sqrt(x0 * x0 + s2);
            const double alpha = x0 > 0.0 ? -nrmx : nrmx;
            /* v = x - alpha e_i, beta = 2 / v^T v = 1 / (nrmx (nrmx + |x0|)) */
            const double v0 = x0 - alpha;
            const double beta = 1.0 / (nrmx * (nrmx + fabs(x0)));
            for (int r = i + 1; r < n; r++) {
                double *wr = W + (size_t)r * ld;
                double d = wr[i] * v0;
                for (int j = i + 1; j < nc; j++) d += wr[j] * wi[j];
                d *= beta;
                wr[i] -= d * v0;
                for (int j = i + 1; j < nc; j++) wr[j] -= d * wi[j];
            }
            wi[i] = alpha;
            for (int j = i + 1; j < nc; j++) wi[j] = 0.0;
        }
        if (wi[i] < 0.0) for (int r = i; r < n; r++) W[(size_t)r * ld + i] = -W[(size_t)r * ld + i];   /* positive diagonal */
    }
}

/* use_lq: HPIPM d_ocp_qp_fact_lq_solve_kkt_step -- the node matrix is never summed up; its Cholesky factor is the L of
 * an LQ factorisation of [chol(H + reg I) | sqrt(Gamma_b) | C^T sqrt(Gamma_g) | BAt Lxx(k+1)] */
static void ipm_fact(const orc_config *c, ipm_mem *m, int use_lq)
{
    const int N = m->N;
    for (int k = N; k >= 0; k--) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        stage_ws *w = &m->ws[k];
        const int nct = NCT(q), nv = q->nv, nu = q->nu;
        for (int i = 0; i < nct; i++) { w->tinv[i] = 1.0 / s->t[i]; w->Gam[i] = s->lam[i] * w->tinv[i]; }
        double Mx[NV][NV];
        for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) Mx[i][j] = q->H[i][j];
        for (int i = 0; i < nv; i++) Mx[i][i] += c->reg_prim;
        double Wq[NV][LQ_MAXC];
        int ncol = 0;
        if (use_lq) {
            /* first block: Cholesky factor of the Hessian itself (HPIPM keeps it across iterations: use_hess_fact) */
            double Lh[NV][NV];
            for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) Lh[i][j] = Mx[i][j];
            chol_ln(Lh, nv);
            memset(Wq, 0, sizeof(Wq));
            for (int i = 0; i < nv; i++) for (int j = 0; j <= i; j++) Wq[i][j] = Lh[i][j];
            ncol = nv;
        }
        for (int j = 0; j < q->nb; j++) {
            int iv = q->idxb[j];
            const double gam = w->Gam[OFF_LB(q) + j] + w->Gam[OFF_UB(q) + j];
            Mx[iv][iv] += gam;
            if (use_lq) { Wq[iv][ncol] = gam > 0.0 ? sqrt(gam) : 0.0; ncol++; }
        }
        for (int j = 0; j < q->ng; j++) {
            double wl = w->Gam[OFF_LG(q) + j], wu = w->Gam[OFF_UG(q) + j], weff;
            if (q->ns) {
                w->Dl[j] = q->Zl[j] + wl + w->Gam[OFF_LS(q) + j];
                w->Du[j] = q->Zu[j] + wu + w->Gam[OFF_US(q) + j];
                weff = wl * (q->Zl[j] + w->Gam[OFF_LS(q) + j]) / w->Dl[j]
                     + wu * (q->Zu[j] + w->Gam[OFF_US(q) + j]) / w->Du[j];
            } else weff = wl + wu;
            for (int a = 0; a < nv; a++) for (int b2 = 0; b2 < nv; b2++)
                Mx[a][b2] += weff * q->Ct[a][j] * q->Ct[b2][j];
            if (use_lq) {
                const double sw = weff > 0.0 ? sqrt(weff) : 0.0;
                for (int a = 0; a < nv; a++) Wq[a][ncol] = q->Ct[a][j] * sw;
                ncol++;
            }
        }
        if (k < N) {
            /* AL = BAt * Lxx(k+1)  (dtrmm), M += AL AL^T (dsyrk) */
            const stage_ws *wn = &m->ws[k + 1];
            const int nun = m->qp[k + 1].nu, nxn = q->nxn;
            double AL[NV][NX];
            for (int i = 0; i < nv; i++) for (int j = 0; j < nxn; j++) {
                double a = 0;
                for (int r = j; r < nxn; r++) a += q->BAt[i][r] * wn->L[nun + r][nun + j];
                AL[i][j] = a;
            }
            for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) {
                double a = 0;
                for (int r = 0; r < nxn; r++) a += AL[i][r] * AL[j][r];
                Mx[i][j] += a;
            }
            if (use_lq) {
                for (int i = 0; i < nv; i++) for (int r = 0; r < nxn; r++) Wq[i][ncol + r] = AL[i][r];
                ncol += nxn;
            }
        }
        if (use_lq) {
            lq_pd(&Wq[0][0], nv, ncol, LQ_MAXC);
            for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) w->L[i][j] = j <= i ? Wq[i][j] : 0.0;
        } else {
            for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) w->L[i][j] = Mx[i][j];
            chol_ln(w->L, nv);
        }
        if (c->lq_fact == -1 && nu == 2 && q->nx == NX) {
            /* DEBUG/ANALYSIS ONLY: classical Riccati -- replace Lxx by chol(Mxx - Mxu Muu^-1 Mux) with the
             * Schur complement formed explicitly through the 2x2 inverse (what a naive GPU kernel would do) */
            double det = Mx[0][0] * Mx[1][1] - Mx[1][0] * Mx[1][0], di = 1.0 / det;
            double i00 = Mx[1][1] * di, i01 = -Mx[1][0] * di, i11 = Mx[0][0] * di;
            double P[NX][NX];
            for (int i = 0; i < NX; i++) for (int j = 0; j < NX; j++) {
                double k0 = -(i00 * Mx[2 + j][0] + i01 * Mx[2 + j][1]), k1 = -(i01 * Mx[2 + j][0] + i11 * Mx[2 + j][1]);
                P[i][j] = Mx[2 + i][2 + j] + Mx[2 + i][0] * k0 + Mx[2 + i][1] * k1;
            }
#ifdef ORC_DEBUG
            {   double maxrel = 0, maxabs = 0, pmax = 0, pmin = 1e300;
                for (int i = 0; i < NX; i++) for (int j = 0; j <= i; j++) {
                    double ps = 0; for (int r = 0; r <= j; r++) ps += w->L[2 + i][2 + r] * w->L[2 + j][2 + r];
                    double d = fabs(ps - P[i][j]); if (d > maxabs) maxabs = d;
                    if (fabs(ps) > pmax) pmax = fabs(ps);
                    if (i == j && fabs(ps) < pmin) pmin = fabs(ps);
                }
                double mmax = 0; for (int i = 0; i < nv; i++) if (fabs(Mx[i][i]) > mmax) mmax = fabs(Mx[i][i]);
                if (maxabs > 1e-7 * pmin) fprintf(stderr, "  stage %d: |P_cl-P_sqrt| %.2e  max|P| %.2e min diag P %.2e  max diag M %.2e det %.3e m00 %.3e m11 %.3e\n", k, maxabs, pmax, pmin, mmax, det, Mx[0][0], Mx[1][1]);
            }
#endif
            double Lx[NV][NV];
            memset(Lx, 0, sizeof(Lx));
            for (int i = 0; i < NX; i++) for (int j = 0; j < NX; j++) Lx[i][j] = P[i][j];
            chol_ln(Lx, NX);
            for (int i = 0; i < NX; i++) for (int j = 0; j < NX; j++) w->L[2 + i][2 + j] = Lx[i][j];
        }
        (void)nu;
    }
}

/* y = Lxx (Lxx^T x) + p of stage k  (cost-to-go gradient) */
static void ctg_grad(const stage_qp *q, const stage_ws *w, const double *x, double *y)
{
    const int nu = q->nu, nx = q->nx > NX ? NX : q->nx;
    double tmp[NX];
    for (int j = 0; j < nx; j++) {
        double a = 0;
        for (int r = j; r < nx; r++) a += w->L[nu + r][nu + j] * x[r];
        tmp[j] = a;
    }
    for (int i = 0; i < nx; i++) {
        double a = w->p[i];
        for (int j = 0; j <= i; j++) a += w->L[nu + i][nu + j] * tmp[j];
        y[i] = a;
    }
}

/* solve the Newton/KKT system for right-hand side `rhs` with the current factorisation
 * (HPIPM d_ocp_qp_solve_kkt_step + d_compute_lam_t_qp + slack expansion) */
static void ipm_solve(ipm_mem *m, const stage_res *rhs, stage_sol *out)
{
    const int N = m->N;
    /* backward vector sweep */
    for (int k = N; k >= 0; k--) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        const stage_res *r = &rhs[k];
        stage_ws *w = &m->ws[k];
        const int nv = q->nv, nu = q->nu, nx = q->nx;
        double gt[NV];
        for (int i = 0; i < nv; i++) gt[i] = r->rg[i];
        for (int j = 0; j < q->nb; j++) {
            int il = OFF_LB(q) + j, iu = OFF_UB(q) + j;
            double ql = (r->rm[il] + s->lam[il] * r->rd[il]) * w->tinv[il];
            double qu = (r->rm[iu] + s->lam[iu] * r->rd[iu]) * w->tinv[iu];
            gt[q->idxb[j]] += ql - qu;
        }
        for (int j = 0; j < q->ng; j++) {
            int il = OFF_LG(q) + j, iu = OFF_UG(q) + j;
            double ql = (r->rm[il] + s->lam[il] * r->rd[il]) * w->tinv[il];
            double qu = (r->rm[iu] + s->lam[iu] * r->rd[iu]) * w->tinv[iu];
            double coef;
            if (q->ns) {
                int jl = OFF_LS(q) + j, ju = OFF_US(q) + j;
                double qls = (r->rm[jl] + s->lam[jl] * r->rd[jl]) * w->tinv[jl];
                double qus = (r->rm[ju] + s->lam[ju] * r->rd[ju]) * w->tinv[ju];
                w->rhol[j] = r->rgsl[j] + ql + qls;
                w->rhou[j] = r->rgsu[j] + qu + qus;
                coef = (ql - w->Gam[il] * w->rhol[j] / w->Dl[j]) - (qu - w->Gam[iu] * w->rhou[j] / w->Du[j]);
            } else coef = ql - qu;
            for (int i = 0; i < nv; i++) gt[i] += q->Ct[i][j] * coef;
        }
        if (k < N) {
            double y[NX];
            ctg_grad(&m->qp[k + 1], &m->ws[k + 1], r->rb, y);
            for (int i = 0; i < nv; i++) for (int j = 0; j < q->nxn; j++) gt[i] += q->BAt[i][j] * y[j];
        }
        /* lu = Luu^{-1} gt_u ; p = gt_x - Lxu lu */
        for (int i = 0; i < nu; i++) {
            double a = gt[i];
            for (int j = 0; j < i; j++) a -= w->L[i][j] * w->lu[j];
            w->lu[i] = a / w->L[i][i];
        }
        for (int i = 0; i < nx; i++) {
            double a = gt[nu + i];
            for (int j = 0; j < nu; j++) a -= w->L[nu + i][j] * w->lu[j];
            w->p[i] = a;
        }
    }
    /* forward sweep */
    double xcur[NX] = {0};
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_ws *w = &m->ws[k];
        stage_sol *o = &out[k];
        const int nv = q->nv, nu = q->nu, nx = q->nx;
        for (int i = 0; i < nx; i++) o->v[nu + i] = xcur[i];
        /* du = -Luu^{-T} (lu + Lxu^T dx) */
        {
            double rhsu[NU];
            for (int i = 0; i < nu; i++) {
                double a = w->lu[i];
                for (int j = 0; j < nx; j++) a += w->L[nu + j][i] * xcur[j];
                rhsu[i] = -a;
            }
            for (int i = nu - 1; i >= 0; i--) {
                double a = rhsu[i];
                for (int j = i + 1; j < nu; j++) a -= w->L[j][i] * o->v[j];
                o->v[i] = a / w->L[i][i];
            }
        }
        if (k < N) {
            double xn[NX];
            for (int j = 0; j < q->nxn; j++) {
                double a = rhs[k].rb[j];
                for (int i = 0; i < nv; i++) a += q->BAt[i][j] * o->v[i];
                xn[j] = a;
            }
            ctg_grad(&m->qp[k + 1], &m->ws[k + 1], xn, o->pi);
            for (int j = 0; j < q->nxn; j++) xcur[j] = xn[j];
        }
    }
    /* expand slacks, dt, dlam */
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        const stage_res *r = &rhs[k];
        const stage_ws *w = &m->ws[k];
        stage_sol *o = &out[k];
        for (int j = 0; j < q->nb; j++) {
            int il = OFF_LB(q) + j, iu = OFF_UB(q) + j;
            o->t[il] = o->v[q->idxb[j]] + r->rd[il];
            o->t[iu] = -o->v[q->idxb[j]] + r->rd[iu];
        }
        for (int j = 0; j < q->ng; j++) {
            int il = OFF_LG(q) + j, iu = OFF_UG(q) + j;
            double cdv = cdot(q, j, o->v);
            double dsl = 0, dsu = 0;
            if (q->ns) {
                dsl = -(w->Gam[il] * cdv + w->rhol[j]) / w->Dl[j];
                dsu = -(-w->Gam[iu] * cdv + w->rhou[j]) / w->Du[j];
                o->sl[j] = dsl; o->su[j] = dsu;
                o->t[OFF_LS(q) + j] = dsl + r->rd[OFF_LS(q) + j];
                o->t[OFF_US(q) + j] = dsu + r->rd[OFF_US(q) + j];
            }
            o->t[il] = cdv + dsl + r->rd[il];
            o->t[iu] = -cdv + dsu + r->rd[iu];
        }
        const int nct = NCT(q);
        for (int i = 0; i < nct; i++) o->lam[i] = -(r->rm[i] + s->lam[i] * o->t[i]) * w->tinv[i];
    }
}

/* residual of the linear system at `st` (HPIPM d_ocp_qp_res_compute_lin) */
static void ipm_res_lin(ipm_mem *m, const stage_res *rhs, const stage_sol *st, stage_res *lin, double nrm[4])
{
    const int N = m->N;
    nrm[0] = nrm[1] = nrm[2] = nrm[3] = 0;
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        const stage_sol *d = &st[k];
        const stage_res *r = &rhs[k];
        stage_res *l = &lin[k];
        const int nu = q->nu, nv = q->nv;
        for (int i = 0; i < nv; i++) {
            double a = r->rg[i];
            for (int j = 0; j < nv; j++) a += q->H[i][j] * d->v[j];
            for (int j = 0; j < q->nxn; j++) a += q->BAt[i][j] * d->pi[j];
            l->rg[i] = a;
        }
        if (k > 0) for (int i = 0; i < q->nx; i++) l->rg[nu + i] -= st[k - 1].pi[i];
        for (int j = 0; j < q->nb; j++) {
            int iv = q->idxb[j], il = OFF_LB(q) + j, iu = OFF_UB(q) + j;
            l->rg[iv] += -d->lam[il] + d->lam[iu];
            l->rd[il] = d->v[iv] - d->t[il] + r->rd[il];
            l->rd[iu] = -d->v[iv] - d->t[iu] + r->rd[iu];
        }
        for (int j = 0; j < q->ng; j++) {
            int il = OFF_LG(q) + j, iu = OFF_UG(q) + j;
            double cdv = cdot(q, j, d->v);
            double dl = d->lam[iu] - d->lam[il];
            for (int i = 0; i < nv; i++) l->rg[i] += q->Ct[i][j] * dl;
            double dsl = q->ns ? d->sl[j] : 0.0, dsu = q->ns ? d->su[j] : 0.0;
            l->rd[il] = cdv + dsl - d->t[il] + r->rd[il];
            l->rd[iu] = -cdv + dsu - d->t[iu] + r->rd[iu];
        }
        for (int j = 0; j < q->ns; j++) {
            int jl = OFF_LS(q) + j, ju = OFF_US(q) + j;
            l->rgsl[j] = q->Zl[j] * d->sl[j] + r->rgsl[j] - d->lam[OFF_LG(q) + j] - d->lam[jl];
            l->rgsu[j] = q->Zu[j] * d->su[j] + r->rgsu[j] - d->lam[OFF_UG(q) + j] - d->lam[ju];
            l->rd[jl] = d->sl[j] - d->t[jl] + r->rd[jl];
            l->rd[ju] = d->su[j] - d->t[ju] + r->rd[ju];
        }
        if (k < N) {
            const int nun = m->qp[k + 1].nu;
            for (int j = 0; j < q->nxn; j++) {
                double a = r->rb[j] - st[k + 1].v[nun + j];
                for (int i = 0; i < nv; i++) a += q->BAt[i][j] * d->v[i];
                l->rb[j] = a;
            }
        }
        const int nct = NCT(q);
        for (int i = 0; i < nct; i++) l->rm[i] = s->lam[i] * d->t[i] + s->t[i] * d->lam[i] + r->rm[i];
        for (int i = 0; i < nv; i++) nrm[0] = fmax(nrm[0], fabs(l->rg[i]));
        for (int j = 0; j < q->ns; j++) { nrm[0] = fmax(nrm[0], fabs(l->rgsl[j])); nrm[0] = fmax(nrm[0], fabs(l->rgsu[j])); }
        for (int j = 0; j < q->nxn; j++) nrm[1] = fmax(nrm[1], fabs(l->rb[j]));
        for (int i = 0; i < nct; i++) { nrm[2] = fmax(nrm[2], fabs(l->rd[i])); nrm[3] = fmax(nrm[3], fabs(l->rm[i])); }
    }
}

/* HPIPM d_compute_alpha_qp: largest step in (0,1] keeping lam,t >= 0 */
static double ipm_alpha(const ipm_mem *m, const stage_sol *st)
{
    double alpha = -1.0;   /* negative reciprocal convention of HPIPM */
    double ap = -1.0, ad = -1.0;
    for (int k = 0; k <= m->N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        const int nct = NCT(q);
        for (int i = 0; i < nct; i++) {
            if (ad * st[k].lam[i] > s->lam[i]) ad = s->lam[i] / st[k].lam[i];
            if (ap * st[k].t[i] > s->t[i]) ap = s->t[i] / st[k].t[i];
        }
    }
    alpha = ap > ad ? ap : ad;
    return -alpha;
}

static double ipm_mu_aff(const ipm_mem *m, const stage_sol *st, double alpha)
{
    double a = 0; int nc = 0;
    for (int k = 0; k <= m->N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        const int nct = NCT(q);
        for (int i = 0; i < nct; i++)
            a += (s->lam[i] + alpha * st[k].lam[i]) * (s->t[i] + alpha * st[k].t[i]);
        nc += nct;
    }
    return a / nc;
}

/* the interior-point loop (HPIPM d_ocp_qp_ipm_solve) */
static int ipm_solve_qp(const orc_config *c, ipm_mem *m, int *iter_out, double *stats)
{
    const int N = m->N;
    double mu, nrm[4], lnrm[4];
    double alpha = 1.0;
    int n_itref = 0, n_lq = 0, force_lq = 0;
    ipm_init(c, m);
    ipm_res(m, m->res, &mu, nrm);
    int kk;
    for (kk = 0; kk < c->qp_iter_max && alpha > c->alpha_min &&
                 (nrm[0] > c->res_g_max || nrm[1] > c->res_b_max ||
                  nrm[2] > c->res_d_max || nrm[3] > c->res_m_max); kk++) {
        /* backup complementarity residual */
        for (int k = 0; k <= N; k++) memcpy(m->rm_bkp + (size_t)k * MAXC, m->res[k].rm, MAXC * sizeof(double));
        /* predictor (affine scaling) direction */
        /* HPIPM lq_fact (mode BALANCE: 1): 0 = Cholesky of the summed node matrix always; 1 = Cholesky, but the accuracy of
         * the predictor is checked (residual of the linear system, absolute bar 1e-5 on every block, NaN counts as
         * inaccurate) and on failure the iteration is REDONE with the LQ-based factorisation, which then stays on for
         * the rest of the solve (force_lq); 2 = LQ always.  stats[7] counts the iterations that used LQ. */
        if (c->lq_fact <= 0) {
            ipm_fact(c, m, 0);
            ipm_solve(m, m->res, m->step);
        } else if (c->lq_fact == 1 && !force_lq) {
            ipm_fact(c, m, 0);
            ipm_solve(m, m->res, m->step);
            ipm_res_lin(m, m->res, m->step, m->lin, lnrm);
            if (!(lnrm[0] <= 1e-5 && lnrm[1] <= 1e-5 && lnrm[2] <= 1e-5 && lnrm[3] <= 1e-5)) {
                ipm_fact(c, m, 1);
                ipm_solve(m, m->res, m->step);
                force_lq = 1;
                n_lq++;
            }
        } else {
            ipm_fact(c, m, 1);
            ipm_solve(m, m->res, m->step);
            n_lq++;
        }
#ifdef ORC_DEBUG
        fprintf(stderr, "it %d mu %.3e res %.2e %.2e %.2e %.2e | lq %d\n", kk, mu, nrm[0], nrm[1], nrm[2], nrm[3], force_lq);
#endif
        alpha = ipm_alpha(m, m->step);
        if (c->pred_corr) {
            double mu_aff = ipm_mu_aff(m, m->step, alpha);
            double tmp = mu_aff / mu;
            double sigma = tmp * tmp * tmp;
            double sigma_mu = sigma * mu;
            sigma_mu = sigma_mu > c->tau_min ? sigma_mu : c->tau_min;
            /* centering + second-order correction */
            for (int k = 0; k <= N; k++) {
                const int nct = NCT(&m->qp[k]);
                for (int i = 0; i < nct; i++)
                    m->res[k].rm[i] = m->rm_bkp[(size_t)k * MAXC + i] + m->step[k].t[i] * m->step[k].lam[i] - sigma_mu;
            }
            ipm_solve(m, m->res, m->step);
            alpha = ipm_alpha(m, m->step);
            if (c->cond_pred_corr) {
                double mu_aff0 = mu_aff;
                mu_aff = ipm_mu_aff(m, m->step, alpha);
                if (mu_aff > 2.0 * mu_aff0) {
                    /* centering direction only */
                    for (int k = 0; k <= N; k++) {
                        const int nct = NCT(&m->qp[k]);
                        for (int i = 0; i < nct; i++)
                            m->res[k].rm[i] = m->rm_bkp[(size_t)k * MAXC + i] - sigma_mu;
                    }
                    ipm_solve(m, m->res, m->step);
                    alpha = ipm_alpha(m, m->step);
                }
            }
            /* iterative refinement of the corrector direction */
            int did = 0;
            for (int it = 0; it < c->itref_corr_max; it++) {
                ipm_res_lin(m, m->res, m->step, m->lin, lnrm);
                if (lnrm[0] < c->res_g_max && lnrm[1] < c->res_b_max &&
                    lnrm[2] < c->res_d_max && lnrm[3] < c->res_m_max) break;
                ipm_solve(m, m->lin, m->corr);
                for (int k = 0; k <= N; k++) {
                    const stage_qp *q = &m->qp[k];
                    for (int i = 0; i < q->nv; i++) m->step[k].v[i] += m->corr[k].v[i];
                    for (int i = 0; i < q->ns; i++) { m->step[k].sl[i] += m->corr[k].sl[i]; m->step[k].su[i] += m->corr[k].su[i]; }
                    for (int i = 0; i < q->nxn; i++) m->step[k].pi[i] += m->corr[k].pi[i];
                    const int nct = NCT(q);
                    for (int i = 0; i < nct; i++) { m->step[k].lam[i] += m->corr[k].lam[i]; m->step[k].t[i] += m->corr[k].t[i]; }
                }
                n_itref++; did = 1;
            }
            if (did) alpha = ipm_alpha(m, m->step);
        }
#ifdef ORC_DEBUG
        {   /* which constraint limits alpha */
            double best = 1e300; int bk=-1, bi=-1, isT=0;
            for (int k = 0; k <= N; k++) { const stage_qp *q=&m->qp[k]; for (int i=0;i<NCT(q);i++){
                if (m->step[k].lam[i]<0){ double r=-m->sol[k].lam[i]/m->step[k].lam[i]; if(r<best){best=r;bk=k;bi=i;isT=0;} }
                if (m->step[k].t[i]<0){ double r=-m->sol[k].t[i]/m->step[k].t[i]; if(r<best){best=r;bk=k;bi=i;isT=1;} } } }
            fprintf(stderr, "   alpha %.3e limited by stage %d idx %d (%s) lam %.3e t %.3e dlam %.3e dt %.3e (nb %d ng %d)\n", alpha, bk, bi, isT?"t":"lam",
               m->sol[bk].lam[bi], m->sol[bk].t[bi], m->step[bk].lam[bi], m->step[bk].t[bi], m->qp[bk].nb, m->qp[bk].ng);
        }
#endif
        /* update (HPIPM d_update_var_qp): step shortening, clipping */
        double a = alpha;
        if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
        for (int k = 0; k <= N; k++) {
            const stage_qp *q = &m->qp[k];
            stage_sol *s = &m->sol[k];
            for (int i = 0; i < q->nv; i++) s->v[i] += a * m->step[k].v[i];
            for (int i = 0; i < q->ns; i++) { s->sl[i] += a * m->step[k].sl[i]; s->su[i] += a * m->step[k].su[i]; }
            for (int i = 0; i < q->nxn; i++) s->pi[i] += a * m->step[k].pi[i];
            const int nct = NCT(q);
            for (int i = 0; i < nct; i++) {
                s->lam[i] += a * m->step[k].lam[i];
                s->t[i] += a * m->step[k].t[i];
                if (s->lam[i] <= c->lam_min) s->lam[i] = c->lam_min;
                if (s->t[i] <= c->t_min) s->t[i] = c->t_min;
            }
        }
        ipm_res(m, m->res, &mu, nrm);
    }
    *iter_out = kk;
    if (stats) {
        stats[0] = nrm[0]; stats[1] = nrm[1]; stats[2] = nrm[2]; stats[3] = nrm[3];
        stats[4] = mu; stats[5] = alpha; stats[6] = n_itref; stats[7] = n_lq;
    }
    int status;
    if (kk == c->qp_iter_max) status = 1;                                                     /* MAX_ITER */
    else if (alpha <= c->alpha_min) status = 2;                                               /* MIN_STEP */
    else if (mu != mu || nrm[0] != nrm[0] || nrm[1] != nrm[1] || nrm[2] != nrm[2] || nrm[3] != nrm[3]) status = 3; /* NaN */
    else status = 0;
    return status;
}

/* ------------------------------------------------------------- RTI solve */
int orc_rti_solve(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  double *x, double *u, double *sl, double *su,
                  double *pi, double *lam, double *t,
                  int32_t *qp_iter, int32_t *qp_status, double *stats)
{
    const int N = c->N, M = c->M;
    ipm_mem *m = ipm_get(N);
    double dx0[NX];
    build_qp(c, x0bar, p, yref, zl, Zl, x, u, sl, su, m->qp, dx0);
    int it = 0;
    int qs = ipm_solve_qp(c, m, &it, stats);
    if (qp_iter) *qp_iter = it;
    if (qp_status) *qp_status = qs;
    int status = 0;
    if (qs != 0 && qs != 1) status = 4;      /* ACADOS_QP_FAILURE: iterate not updated (App. B.3) */
    else {
        /* full step, SQP_RTI */
        for (int i = 0; i < NX; i++) x[i] += dx0[i];
        for (int k = 0; k <= N; k++) {
            const stage_qp *q = &m->qp[k];
            const stage_sol *s = &m->sol[k];
            for (int j = 0; j < q->nu; j++) u[k * NU + j] += s->v[j];
            if (k > 0) for (int j = 0; j < NX; j++) x[k * NX + j] += s->v[q->nu + j];
            if (c->soft) for (int j = 0; j < M; j++) { sl[k * M + j] += s->sl[j]; su[k * M + j] += s->su[j]; }
            if (pi && k < N) for (int j = 0; j < NX; j++) pi[k * NX + j] = s->pi[j];
        }
    }
    if (lam && t) {
        /* padded fixed layout per node: [lb(6) lg(M) ub(6) ug(M) ls(M) us(M)], box order [u0,u1,x,y,v,w] */
        const int stride = 12 + 4 * M;
        memset(lam, 0, sizeof(double) * (size_t)(N + 1) * stride);
        memset(t, 0, sizeof(double) * (size_t)(N + 1) * stride);
        for (int k = 0; k <= N; k++) {
            const stage_qp *q = &m->qp[k];
            const stage_sol *s = &m->sol[k];
            double *lk = lam + (size_t)k * stride, *tk = t + (size_t)k * stride;
            for (int j = 0; j < q->nb; j++) {
                int iv = q->idxb[j];
                int slot = (iv < q->nu) ? iv : 2 + (iv - q->nu == 0 ? 0 : iv - q->nu == 1 ? 1 : iv - q->nu == 3 ? 2 : 3);
                lk[slot] = s->lam[OFF_LB(q) + j]; tk[slot] = s->t[OFF_LB(q) + j];
                lk[6 + M + slot] = s->lam[OFF_UB(q) + j]; tk[6 + M + slot] = s->t[OFF_UB(q) + j];
            }
            for (int j = 0; j < q->ng; j++) {
                lk[6 + j] = s->lam[OFF_LG(q) + j]; tk[6 + j] = s->t[OFF_LG(q) + j];
                lk[12 + M + j] = s->lam[OFF_UG(q) + j]; tk[12 + M + j] = s->t[OFF_UG(q) + j];
            }
            for (int j = 0; j < q->ns; j++) {
                lk[12 + 2 * M + j] = s->lam[OFF_LS(q) + j]; tk[12 + 2 * M + j] = s->t[OFF_LS(q) + j];
                lk[12 + 3 * M + j] = s->lam[OFF_US(q) + j]; tk[12 + 3 * M + j] = s->t[OFF_US(q) + j];
            }
        }
    }
    ipm_free(m);
    return status;
}

typedef struct {
    const orc_config *c; int64_t B; atomic_llong next;
    const double *x0bar, *p, *yref, *zl, *Zl;
    double *x, *u, *sl, *su; int32_t *status, *qp_iter; double *stats;
    double *pi, *lam, *t;          /* optional multiplier outputs (NULL: not wanted) */
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const orc_config *c = j->c;
    const int N = c->N, M = c->M;
    const size_t nxs = (size_t)(N + 1) * NX, nus = (size_t)N * NU, nss = (size_t)(N + 1) * M;
    const size_t nps = (size_t)(N + 1) * 2 * M, nys = (size_t)(N + 1) * ORC_NY;
    for (;;) {
        long long b0 = atomic_fetch_add(&j->next, 8);
        if (b0 >= j->B) break;
        long long b1 = b0 + 8 < j->B ? b0 + 8 : j->B;
        for (long long b = b0; b < b1; b++) {
            int32_t it = 0, qs = 0;
            double st[8];
            int s = orc_rti_solve(c, j->x0bar + b * NX, j->p + b * nps, j->yref + b * nys,
                                  j->zl + b * nss, j->Zl + b * nss,
                                  j->x + b * nxs, j->u + b * nus, j->sl + b * nss, j->su + b * nss,
                                  j->pi ? j->pi + b * (size_t)N * NX : NULL,
                                  j->lam ? j->lam + b * (size_t)(N + 1) * (12 + 4 * M) : NULL,
                                  j->t ? j->t + b * (size_t)(N + 1) * (12 + 4 * M) : NULL, &it, &qs, st);
            if (j->status) j->status[b] = s;
            if (j->qp_iter) j->qp_iter[b] = it;
            if (j->stats) memcpy(j->stats + b * 8, st, sizeof(st));
        }
    }
    return NULL;
}

/* batch of independent solves over `nthreads` host threads (pthreads; problems are
 * independent, so this is the CPU analogue of one-process-per-core acados) */
void orc_rti_solve_batch(const orc_config *c, int64_t B, int nthreads,
                         const double *x0bar, const double *p, const double *yref,
                         const double *zl, const double *Zl,
                         double *x, double *u, double *sl, double *su,
                         int32_t *status, int32_t *qp_iter, double *stats)
{
    orc_rti_solve_batch_duals(c, B, nthreads, x0bar, p, yref, zl, Zl, x, u, sl, su, NULL, NULL, NULL, status, qp_iter, stats);
}

/* same, also returning the multipliers pi[B,N*5], lam/t[B,(N+1)*(12+4M)] (any of them may be NULL) */
void orc_rti_solve_batch_duals(const orc_config *c, int64_t B, int nthreads,
                               const double *x0bar, const double *p, const double *yref,
                               const double *zl, const double *Zl,
                               double *x, double *u, double *sl, double *su,
                               double *pi, double *lam, double *t,
                               int32_t *status, int32_t *qp_iter, double *stats)
{
    batch_job j = {c, B, 0, x0bar, p, yref, zl, Zl, x, u, sl, su, status, qp_iter, stats, pi, lam, t};
    atomic_init(&j.next, 0);
    if (nthreads <= 1) { batch_worker(&j); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, batch_worker, &j);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------ full SQP mode */
/* north_star kernel 4 ("SQP outer loop with convergence test").  The reference itself never runs it
 * (nlp_solver_type = 'SQP_RTI', robot_ocp_problem.py:130): this follows acados' `SQP` solver as published
 * (ocp_nlp_sqp: linearise -> NLP residuals -> convergence test -> QP -> full step, FIXED_STEP globalisation,
 * tolerances nlp_solver_tol_{stat,eq,ineq,comp}), restated from memory like the rest of this file. */

/* KKT residuals of the NLP at (x,u,sl,su) with multipliers (pi, lam, t) in the padded layout of
 * orc_rti_solve: res = [stationarity, dynamics gap, inequality residual g_i - t_i (incl. the x0 embedding),
 * complementarity lam_i t_i], infinity norms.  Written directly from the model functions (no QP data). */
void orc_nlp_residuals(const orc_config *c, const double *x0bar, const double *p,
                       const double *yref, const double *zl, const double *Zl,
                       const double *x, const double *u, const double *sl, const double *su,
                       const double *pi, const double *lam, const double *t, double res[4])
{
    const int N = c->N, M = c->M, nct = 12 + 4 * M;
    const double dt = c->tf / N;
    static const int idxbx[4] = {0, 1, 3, 4};
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (int i = 0; i < NX; i++) r2 = fmax(r2, fabs(x0bar[i] - x[i]));
    for (int k = 0; k <= N; k++) {
        const double *xk = x + k * NX, *uk = u + k * NU, *yr = yref + k * ORC_NY;
        const double *lk = lam + (size_t)k * nct, *tk = t + (size_t)k * nct;
        const double scal = k < N ? dt : 1.0;
        double gu[NU] = {0, 0}, gx[NX] = {0, 0, 0, 0, 0};
        if (k < N) {
            double A[25], Bm[10], xn[5];
            if (c->integrator == 1) orc_rk4_step(xk, uk, dt, xn, A, Bm);
            else orc_gl4_step(xk, uk, dt, xn, A, Bm);
            const double *pk = pi + k * NX;
            for (int r = 0; r < NX; r++) {
                r1 = fmax(r1, fabs(xn[r] - x[(k + 1) * NX + r]));
                for (int j = 0; j < NU; j++) gu[j] += Bm[r * 2 + j] * pk[r];
                for (int j = 0; j < NX; j++) gx[j] += A[r * 5 + j] * pk[r];
            }
            for (int j = 0; j < NU; j++) gu[j] += scal * c->W[4 + j] * (uk[j] - yr[4 + j]);
            for (int j = 0; j < 4; j++) gx[idxbx[j]] += scal * c->W[j] * (xk[idxbx[j]] - yr[j]);
        } else {
            for (int j = 0; j < 4; j++) gx[idxbx[j]] += c->We[j] * (xk[idxbx[j]] - yr[j]);
        }
        if (k > 0) for (int j = 0; j < NX; j++) gx[j] -= pi[(k - 1) * NX + j];
        /* boxes: slots [u0,u1,x,y,v,w]; lower at lk[slot], upper at lk[6+M+slot] */
        for (int s = 0; s < 6; s++) {
            const int on = s < 2 ? (k < N) : (k > 0 && k < N);
            if (!on) continue;
            const double val = s < 2 ? uk[s] : xk[idxbx[s - 2]];
            const double lo = s < 2 ? c->lbu[s] : c->lbx[s - 2], hi = s < 2 ? c->ubu[s] : c->ubx[s - 2];
            const double ll = lk[s], lu = lk[6 + M + s], tl = tk[s], tu = tk[6 + M + s];
            if (s < 2) gu[s] += lu - ll; else gx[idxbx[s - 2]] += lu - ll;
            r2 = fmax(r2, fmax(fabs(val - lo - tl), fabs(hi - val - tu)));
            r3 = fmax(r3, fmax(fabs(ll * tl), fabs(lu * tu)));
        }
        double hh[MM], cx[MM], cy[MM];
        orc_con_h(xk, p + (size_t)k * 2 * M, M, c->r2, hh, cx, cy);
        for (int j = 0; j < M; j++) {
            const double s_l = c->soft ? sl[k * M + j] : 0.0, s_u = c->soft ? su[k * M + j] : 0.0;
            const double l0 = lk[6 + j], l1 = lk[12 + M + j], t0 = tk[6 + j], t1 = tk[12 + M + j];
            gx[0] += cx[j] * (l1 - l0); gx[1] += cy[j] * (l1 - l0);
            r2 = fmax(r2, fmax(fabs(hh[j] + s_l - c->lh - t0), fabs(c->uh - hh[j] + s_u - t1)));
            r3 = fmax(r3, fmax(fabs(l0 * t0), fabs(l1 * t1)));
            if (c->soft) {
                const double l2 = lk[12 + 2 * M + j], l3 = lk[12 + 3 * M + j], t2 = tk[12 + 2 * M + j], t3 = tk[12 + 3 * M + j];
                const double gsl = scal * (zl[k * M + j] + Zl[k * M + j] * s_l) - l0 - l2;
                const double gsu = scal * (c->zu + c->Zu * s_u) - l1 - l3;
                r0 = fmax(r0, fmax(fabs(gsl), fabs(gsu)));
                r2 = fmax(r2, fmax(fabs(s_l - t2), fabs(s_u - t3)));
                r3 = fmax(r3, fmax(fabs(l2 * t2), fabs(l3 * t3)));
            }
        }
        if (k < N) for (int j = 0; j < NU; j++) r0 = fmax(r0, fabs(gu[j]));
        if (k > 0) for (int j = 0; j < NX; j++) r0 = fmax(r0, fabs(gx[j]));      /* x_0 is fixed by the embedding */
    }
    res[0] = r0; res[1] = r1; res[2] = r2; res[3] = r3;
}

/* merit function of acados' MERIT_BACKTRACKING line search (ocp_nlp_evaluate_merit_fun): cost + sum w_pi |dynamics gap|
 * + sum w_lam max(0, constraint violation), slack variables and their penalties included.  x[0] is the embedded x0bar. */
static double sqp_merit(const orc_config *c, const double *p, const double *yref, const double *zl, const double *Zl,
                        const double *x, const double *u, const double *sl, const double *su,
                        const double *wpi, const double *wlam)
{
    const int N = c->N, M = c->M, nct = 12 + 4 * M;
    const double dt = c->tf / N;
    static const int idxbx[4] = {0, 1, 3, 4};
    double m = 0.0;
    for (int k = 0; k <= N; k++) {
        const double *xk = x + k * NX, *uk = u + k * NU, *yr = yref + k * ORC_NY;
        const double *wk = wlam + (size_t)k * nct;
        const double scal = k < N ? dt : 1.0;
        double mk = 0.0;
        if (k < N) {
            double xn[5];
            if (c->integrator == 1) orc_rk4_step(xk, uk, dt, xn, NULL, NULL);
            else orc_gl4_step(xk, uk, dt, xn, NULL, NULL);
            for (int r = 0; r < NX; r++) mk += wpi[k * NX + r] * fabs(xn[r] - x[(k + 1) * NX + r]);
            double cst = 0.0;
            for (int j = 0; j < NU; j++) cst += c->W[4 + j] * (uk[j] - yr[4 + j]) * (uk[j] - yr[4 + j]);
            for (int j = 0; j < 4; j++) cst += c->W[j] * (xk[idxbx[j]] - yr[j]) * (xk[idxbx[j]] - yr[j]);
            mk += 0.5 * scal * cst;
        } else {
            double cst = 0.0;
            for (int j = 0; j < 4; j++) cst += c->We[j] * (xk[idxbx[j]] - yr[j]) * (xk[idxbx[j]] - yr[j]);
            mk += 0.5 * cst;
        }
        for (int s = 0; s < 6; s++) {
            const int on = s < 2 ? (k < N) : (k > 0 && k < N);
            if (!on) continue;
            const double val = s < 2 ? uk[s] : xk[idxbx[s - 2]];
            const double lo = s < 2 ? c->lbu[s] : c->lbx[s - 2], hi = s < 2 ? c->ubu[s] : c->ubx[s - 2];
            mk += wk[s] * fmax(0.0, lo - val) + wk[6 + M + s] * fmax(0.0, val - hi);
        }
        double hh[MM], cx[MM], cy[MM];
        orc_con_h(xk, p + (size_t)k * 2 * M, M, c->r2, hh, cx, cy);
        for (int j = 0; j < M; j++) {
            const double s_l = c->soft ? sl[k * M + j] : 0.0, s_u = c->soft ? su[k * M + j] : 0.0;
            mk += wk[6 + j] * fmax(0.0, c->lh - hh[j] - s_l) + wk[12 + M + j] * fmax(0.0, hh[j] - c->uh - s_u);
            if (c->soft) {
                mk += scal * (zl[k * M + j] * s_l + 0.5 * Zl[k * M + j] * s_l * s_l + c->zu * s_u + 0.5 * c->Zu * s_u * s_u);
                mk += wk[12 + 2 * M + j] * fmax(0.0, -s_l) + wk[12 + 3 * M + j] * fmax(0.0, -s_u);
            }
        }
        m += mk;
    }
    return m;
}

/* SQP: returns acados status 0 (converged), 2 (max_iter), 4 (QP failure).
 * globalization 0 = FIXED_STEP (full steps, acados' default), 1 = MERIT_BACKTRACKING (alpha <- 0.7 alpha while the
 * merit function does not decrease, alpha_min 0.05 -- ls[3] = {alpha_min, alpha_reduction, eps}, a step being accepted when
 * merit(alpha) < merit(0) + eps max(1, |merit(0)|): eps = 0 is the plain test, a few ulp of slack keep the rounding noise
 * of the merit sum from rejecting full steps near convergence; weights |multipliers| at the first iteration, then
 * max(|mult|, (w + |mult|)/2); primal AND dual variables move by alpha; the x0 embedding is always taken in full,
 * x0 being eliminated from the QP).
 * pi/lam/t are IN/OUT (multipliers of the previous call or zeros); *sqp_iter = QPs solved, *qp_iter = their
 * interior-point iterations in total, res[4] = NLP residuals at exit. */
int orc_sqp_solve(const orc_config *c, int max_iter, const double tol[4], int globalization, const double ls[3],
                  const double *x0bar, const double *p, const double *yref, const double *zl, const double *Zl,
                  double *x, double *u, double *sl, double *su, double *pi, double *lam, double *t,
                  int32_t *sqp_iter, int32_t *qp_iter, double *res, double *stats)
{
    const int N = c->N, M = c->M, nct = 12 + 4 * M;
    const size_t nxs = (size_t)(N + 1) * NX, nus = (size_t)N * NU, nss = (size_t)(N + 1) * M, nls = (size_t)(N + 1) * nct;
    int it = 0, qtot = 0, status, n_maxed = 0;
    double r[4], st8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double ls_amin = ls ? ls[0] : 0.05, ls_red = ls ? ls[1] : 0.7, ls_eps = ls ? ls[2] : 0.0;
    double *w = NULL;
    if (globalization) w = (double *)calloc((size_t)(N + 1) * NV + 2 * nss + (size_t)N * NX * 2 + 3 * nls + nxs + nus + 2 * nss, sizeof(double));
    double *v = w, *dsl = v ? v + (size_t)(N + 1) * NV : NULL, *dsu = v ? dsl + nss : NULL, *pq = v ? dsu + nss : NULL;
    double *wpi = v ? pq + (size_t)N * NX : NULL, *lq = v ? wpi + (size_t)N * NX : NULL, *tq = v ? lq + nls : NULL, *wlam = v ? tq + nls : NULL;
    double *xt = v ? wlam + nls : NULL, *ut = v ? xt + nxs : NULL, *slt = v ? ut + nus : NULL, *sut = v ? slt + nss : NULL;
    for (;; it++) {
        orc_nlp_residuals(c, x0bar, p, yref, zl, Zl, x, u, sl, su, pi, lam, t, r);
        if (r[0] < tol[0] && r[1] < tol[1] && r[2] < tol[2] && r[3] < tol[3]) { status = 0; break; }
        if (it >= max_iter) { status = 2; break; }
        int32_t qi = 0, qs = 0;
        if (!globalization) {
            const int s = orc_rti_solve(c, x0bar, p, yref, zl, Zl, x, u, sl, su, pi, lam, t, &qi, &qs, st8);
            qtot += qi; n_maxed += (qs == 1);
            if (s != 0) { status = s; it++; break; }
            continue;
        }
        orc_qp_solve_export(c, x0bar, p, yref, zl, Zl, x, u, sl, su, v, dsl, dsu, pq, lq, tq, &qi, &qs, st8);
        qtot += qi; n_maxed += (qs == 1);
        if (qs != 0 && qs != 1) { status = 4; it++; memcpy(lam, lq, nls * sizeof(double)); memcpy(t, tq, nls * sizeof(double)); break; }
        for (int i = 0; i < NX; i++) x[i] += v[2 + i];                       /* embedding: x_0 <- x0bar */
        for (size_t i = 0; i < (size_t)N * NX; i++) { const double a = fabs(pq[i]); wpi[i] = it == 0 ? a : fmax(a, 0.5 * (wpi[i] + a)); }
        for (size_t i = 0; i < nls; i++) { const double a = fabs(lq[i]); wlam[i] = it == 0 ? a : fmax(a, 0.5 * (wlam[i] + a)); }
        const double m0 = sqp_merit(c, p, yref, zl, Zl, x, u, sl, su, wpi, wlam);
        double alpha = 1.0;
        for (;;) {
            for (int i = 0; i < NX; i++) xt[i] = x[i];
            for (int k = 1; k <= N; k++) for (int i = 0; i < NX; i++) xt[k * NX + i] = x[k * NX + i] + alpha * v[k * NV + 2 + i];
            for (int k = 0; k < N; k++) for (int i = 0; i < NU; i++) ut[k * NU + i] = u[k * NU + i] + alpha * v[k * NV + i];
            for (size_t i = 0; i < nss; i++) { slt[i] = sl[i] + alpha * dsl[i]; sut[i] = su[i] + alpha * dsu[i]; }
            if (alpha <= ls_amin) break;                                     /* alpha_min: accepted without a test */
            const double m1 = sqp_merit(c, p, yref, zl, Zl, xt, ut, slt, sut, wpi, wlam);
            if (m1 < m0 + ls_eps * fmax(1.0, fabs(m0))) break;
            alpha *= ls_red;
            if (alpha < ls_amin) alpha = ls_amin;
        }
        memcpy(x, xt, nxs * sizeof(double)); memcpy(u, ut, nus * sizeof(double));
        memcpy(sl, slt, nss * sizeof(double)); memcpy(su, sut, nss * sizeof(double));
        if (alpha == 1.0) {
            memcpy(pi, pq, (size_t)N * NX * sizeof(double)); memcpy(lam, lq, nls * sizeof(double)); memcpy(t, tq, nls * sizeof(double));
        } else {
            for (size_t i = 0; i < (size_t)N * NX; i++) pi[i] += alpha * (pq[i] - pi[i]);
            for (size_t i = 0; i < nls; i++) { lam[i] += alpha * (lq[i] - lam[i]); t[i] += alpha * (tq[i] - t[i]); }
        }
        st8[7] = alpha;
    }
    free(w);
    if (sqp_iter) *sqp_iter = it;
    if (qp_iter) *qp_iter = qtot;
    if (res) memcpy(res, r, sizeof(r));
    st8[6] = n_maxed;                                /* QPs of this call that ran into qp_iter_max */
    if (stats) memcpy(stats, st8, sizeof(st8));
    return status;
}

typedef struct {
    const orc_config *c; int64_t B; atomic_llong next; int max_iter; const double *tol; int glob; const double *ls; int32_t *qp_maxed;
    const double *x0bar, *p, *yref, *zl, *Zl;
    double *x, *u, *sl, *su, *pi, *lam, *t; int32_t *status, *sqp_iter, *qp_iter; double *res;
} sqp_job;

static void *sqp_worker(void *arg)
{
    sqp_job *j = (sqp_job *)arg;
    const orc_config *c = j->c;
    const int N = c->N, M = c->M;
    const size_t nxs = (size_t)(N + 1) * NX, nus = (size_t)N * NU, nss = (size_t)(N + 1) * M;
    const size_t nps = (size_t)(N + 1) * 2 * M, nys = (size_t)(N + 1) * ORC_NY, nls = (size_t)(N + 1) * (12 + 4 * M);
    for (;;) {
        long long b = atomic_fetch_add(&j->next, 1);
        if (b >= j->B) break;
        int32_t it = 0, qi = 0;
        double st8[8];
        int s = orc_sqp_solve(c, j->max_iter, j->tol, j->glob, j->ls, j->x0bar + b * NX, j->p + b * nps, j->yref + b * nys,
                              j->zl + b * nss, j->Zl + b * nss, j->x + b * nxs, j->u + b * nus, j->sl + b * nss, j->su + b * nss,
                              j->pi + b * (size_t)N * NX, j->lam + b * nls, j->t + b * nls, &it, &qi, j->res + b * 4, st8);
        j->status[b] = s; j->sqp_iter[b] = it; j->qp_iter[b] = qi;
        if (j->qp_maxed) j->qp_maxed[b] = (int32_t)st8[6];
    }
    return NULL;
}

void orc_sqp_solve_batch(const orc_config *c, int64_t B, int nthreads, int max_iter, const double tol[4], int globalization,
                         const double ls[3],
                         const double *x0bar, const double *p, const double *yref, const double *zl, const double *Zl,
                         double *x, double *u, double *sl, double *su, double *pi, double *lam, double *t,
                         int32_t *status, int32_t *sqp_iter, int32_t *qp_iter, double *res, int32_t *qp_maxed)
{
    sqp_job j = {c, B, 0, max_iter, tol, globalization, ls, qp_maxed, x0bar, p, yref, zl, Zl, x, u, sl, su, pi, lam, t, status, sqp_iter, qp_iter, res};
    atomic_init(&j.next, 0);
    if (nthreads <= 1) { sqp_worker(&j); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, sqp_worker, &j);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------- validation exports */
int orc_export_qp(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  const double *x, const double *u, const double *sl, const double *su,
                  double *BAt, double *b, double *H, double *g, double *Cxy,
                  double *lb, double *ub, double *lg, double *ug,
                  double *ls, double *us, double *Zl_qp, double *Zu_qp,
                  double *zl_qp, double *zu_qp)
{
    /* un-reduced view: every node exported with nv=7 ([u;x]); node 0 carries dx0 in lb=ub of
     * its x entries is NOT done here -- the caller gets the reduced node 0 (A_0 dx0 folded in
     * b_0) with zero x-columns, which is exactly what the solver sees. */
    const int N = c->N, M = c->M;
    stage_qp *qp = (stage_qp *)calloc(N + 1, sizeof(stage_qp));
    double dx0[NX];
    build_qp(c, x0bar, p, yref, zl, Zl, x, u, sl, su, qp, dx0);
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &qp[k];
        /* map local variable index -> global [u0,u1,x,y,psi,v,w] slot */
        int map[NV];
        for (int i = 0; i < q->nu; i++) map[i] = i;
        for (int i = 0; i < q->nx; i++) map[q->nu + i] = 2 + i;
        for (int i = 0; i < q->nv; i++) {
            for (int j = 0; j < q->nxn; j++) BAt[((size_t)k * NV + map[i]) * NX + j] = q->BAt[i][j];
            for (int j = 0; j < q->nv; j++) H[((size_t)k * NV + map[i]) * NV + map[j]] = q->H[i][j];
            g[(size_t)k * NV + map[i]] = q->g[i];
        }
        for (int j = 0; j < q->nxn; j++) b[(size_t)k * NX + j] = q->b[j];
        for (int j = 0; j < 6; j++) { lb[k * 6 + j] = -INFINITY; ub[k * 6 + j] = INFINITY; }
        for (int j = 0; j < q->nb; j++) {
            int gi = map[q->idxb[j]];
            int slot = gi < 2 ? gi : (gi == 2 ? 2 : gi == 3 ? 3 : gi == 5 ? 4 : 5);
            lb[k * 6 + slot] = q->lb[j]; ub[k * 6 + slot] = q->ub[j];
        }
        for (int j = 0; j < M; j++) {
            Cxy[((size_t)k * M + j) * 2 + 0] = (q->nx ? q->Ct[q->nu + 0][j] : 0.0);
            Cxy[((size_t)k * M + j) * 2 + 1] = (q->nx ? q->Ct[q->nu + 1][j] : 0.0);
            lg[k * M + j] = q->lg[j]; ug[k * M + j] = q->ug[j];
            ls[k * M + j] = q->ls[j]; us[k * M + j] = q->us[j];
            Zl_qp[k * M + j] = q->Zl[j]; Zu_qp[k * M + j] = q->Zu[j];
            zl_qp[k * M + j] = q->zl[j]; zu_qp[k * M + j] = q->zu[j];
        }
    }
    free(qp);
    return 0;
}

int orc_qp_solve_export(const orc_config *c, const double *x0bar, const double *p,
                  const double *yref, const double *zl, const double *Zl,
                  const double *x, const double *u, const double *sl, const double *su,
                  double *v, double *dsl, double *dsu, double *pi, double *lam, double *t,
                  int32_t *qp_iter, int32_t *qp_status, double *stats)
{
    const int N = c->N, M = c->M;
    ipm_mem *m = ipm_get(N);
    double dx0[NX];
    build_qp(c, x0bar, p, yref, zl, Zl, x, u, sl, su, m->qp, dx0);
    int it = 0;
    int qs = ipm_solve_qp(c, m, &it, stats);
    *qp_iter = it; *qp_status = qs;
    const int stride = 12 + 4 * M;
    for (int k = 0; k <= N; k++) {
        const stage_qp *q = &m->qp[k];
        const stage_sol *s = &m->sol[k];
        for (int i = 0; i < NV; i++) v[k * NV + i] = 0.0;
        for (int i = 0; i < q->nu; i++) v[k * NV + i] = s->v[i];
        for (int i = 0; i < q->nx; i++) v[k * NV + 2 + i] = s->v[q->nu + i];
        for (int j = 0; j < M; j++) { dsl[k * M + j] = s->sl[j]; dsu[k * M + j] = s->su[j]; }
        if (k < N) for (int j = 0; j < NX; j++) pi[k * NX + j] = s->pi[j];
        double *lk = lam + (size_t)k * stride, *tk = t + (size_t)k * stride;
        for (int i = 0; i < stride; i++) { lk[i] = 0; tk[i] = 0; }
        for (int j = 0; j < q->nb; j++) {
            int iv = q->idxb[j];
            int xi = iv - q->nu;
            int slot = (iv < q->nu) ? iv : 2 + (xi == 0 ? 0 : xi == 1 ? 1 : xi == 3 ? 2 : 3);
            lk[slot] = s->lam[OFF_LB(q) + j]; tk[slot] = s->t[OFF_LB(q) + j];
            lk[6 + M + slot] = s->lam[OFF_UB(q) + j]; tk[6 + M + slot] = s->t[OFF_UB(q) + j];
        }
        for (int j = 0; j < q->ng; j++) {
            lk[6 + j] = s->lam[OFF_LG(q) + j]; tk[6 + j] = s->t[OFF_LG(q) + j];
            lk[12 + M + j] = s->lam[OFF_UG(q) + j]; tk[12 + M + j] = s->t[OFF_UG(q) + j];
        }
        for (int j = 0; j < q->ns; j++) {
            lk[12 + 2 * M + j] = s->lam[OFF_LS(q) + j]; tk[12 + 2 * M + j] = s->t[OFF_LS(q) + j];
            lk[12 + 3 * M + j] = s->lam[OFF_US(q) + j]; tk[12 + 3 * M + j] = s->t[OFF_US(q) + j];
        }
    }
    for (int i = 0; i < NX; i++) v[2 + i] = dx0[i];   /* node 0: the eliminated dx0 */
    ipm_free(m);
    return 0;
}

/* ------------------------------------------------------------------- glue */
/* robot_ocp_problem.py:145-152 */
void orc_slack_weights(const orc_config *c, const double x0[5], const double goal[2],
                       double *zl, double *Zl)
{
    const int N = c->N, M = c->M;
    double d0 = x0[0] - goal[0], d1 = x0[1] - goal[1];
    double scale = 1e4 * (d0 * d0 + d1 * d1 + x0[3] * x0[3] + x0[4] * x0[4] + 50.0);
    for (int i = 0; i <= N; i++) {
        double alpha_i = scale * (N - i) / N;
        for (int j = 0; j < M; j++) { zl[i * M + j] = alpha_i; Zl[i * M + j] = alpha_i; }
    }
}

/* robot_ocp_problem.py:253-258 */
void orc_shift(const orc_config *c, double *x, double *u)
{
    const int N = c->N;
    for (int j = 0; j < N - 1; j++) {
        memcpy(x + j * NX, x + (j + 1) * NX, NX * sizeof(double));
        memcpy(u + j * NU, u + (j + 1) * NU, NU * sizeof(double));
    }
    memcpy(x + (N - 1) * NX, x + N * NX, NX * sizeof(double));
    u[(N - 1) * NU + 0] = 0.0; u[(N - 1) * NU + 1] = 0.0;
}

/* robot_ocp_problem.py:286-306 (reset() zeroes the slacks/multipliers; x0's v,omega are
 * zeroed IN PLACE because x_guess aliases self.x0, :301-302) */
void orc_initial_guess(const orc_config *c, double x0[5], double *x, double *u,
                       double *sl, double *su)
{
    const int N = c->N, M = c->M;
    x0[3] = 0.0; x0[4] = 0.0;
    for (int i = 0; i <= N; i++) {
        memcpy(x + i * NX, x0, NX * sizeof(double));
        if (i < N) { u[i * NU] = 0.0; u[i * NU + 1] = 0.0; }
        for (int j = 0; j < M; j++) { sl[i * M + j] = 0.0; su[i * M + j] = 0.0; }
    }
}

/* src/utils/visualization.py:35-59 (deterministic part of predict_step) */
void orc_obst_predict_step(double o[4], double dt, double xmin, double xmax, double ymin, double ymax)
{
    double x = o[0], y = o[1], vx = o[2], vy = o[3], th;
    if (vx < 0) th = (x - xmin) / fabs(vx);
    else if (vx > 0) th = (xmax - x) / fabs(vx);
    else th = INFINITY;
    if (th <= dt) { x += (vx * th - vx * (dt - th)); vx = -vx; }
    else x += vx * dt;
    if (vy < 0) th = (y - ymin) / fabs(vy);
    else if (vy > 0) th = (ymax - y) / fabs(vy);
    else th = INFINITY;
    if (th <= dt) { y += (vy * th - vy * (dt - th)); vy = -vy; }
    else y += vy * dt;
    o[0] = x; o[1] = y; o[2] = vx; o[3] = vy;
}

/* visualization.py:20-33 + :35-59 ; noise[2] ~ N(0,1) supplied by the caller */
void orc_obst_step(double o[4], double dt, int random_move, double randomness, double vmax,
                   const double noise[2], double xmin, double xmax, double ymin, double ymax)
{
    if (random_move) {
        o[2] = fmin(fmax((1.0 + randomness * noise[0]) * o[2], -vmax), vmax);
        o[3] = fmin(fmax((1.0 + randomness * noise[1]) * o[3], -vmax), vmax);
    }
    orc_obst_predict_step(o, dt, xmin, xmax, ymin, ymax);
}

/* visualization.py:62-79; quirk_vx_is_vy reproduces `vx = self.vy` (:69) */
void orc_obst_predict_traj(const double o[4], int n, double dt, int quirk_vx_is_vy,
                           double xmin, double xmax, double ymin, double ymax, double *traj)
{
    double s[4] = {o[0], o[1], quirk_vx_is_vy ? o[3] : o[2], o[3]};
    traj[0] = s[0]; traj[1] = s[1];
    for (int i = 0; i < n; i++) {
        orc_obst_predict_step(s, dt, xmin, xmax, ymin, ymax);
        traj[2 * (i + 1)] = s[0]; traj[2 * (i + 1) + 1] = s[1];
    }
}
