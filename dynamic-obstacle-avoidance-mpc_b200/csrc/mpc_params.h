// mpc_params.h -- b2mpc_config (include/b200mpc.h) -> the scalar part of the kernel parameter block.
// Host code, shared by the C-ABI (b200mpc.cu) and by the CPU mirror build of the kernels (oracle/simt/), so that both run the
// kernels with identical constants.  The QP data per node follow SURVEY Appendix A / robot_ocp_problem.py:59-122.
#pragma once
#include <string.h>

#include "../../include/b200mpc.h"
#include "mpc_common.cuh"

static inline void b2_params_from_config(const b2mpc_config& c, int Bm, B2Params& p)
{
    memset(&p, 0, sizeof(p));
    p.N = c.N; p.M = c.n_obst; p.Bm = Bm; p.B = 0;
    p.integrator = c.integrator; p.qp_iter_max = c.qp_iter_max;
    p.pred_corr = c.pred_corr; p.cond_pred_corr = c.cond_pred_corr; p.itref_corr_max = c.itref_corr_max;
    p.nct = 12 + 4 * c.n_obst;
    p.nlp_iter_max = c.nlp_iter_max; p.nlp_globalization = c.nlp_globalization;
    for (int i = 0; i < 4; i++) p.nlp_tol[i] = c.nlp_tol[i];
    p.ls_alpha_min = c.ls_alpha_min; p.ls_alpha_red = c.ls_alpha_reduction; p.ls_eps = c.ls_eps;
    p.dt = c.tf / c.N; p.dt2h = 0.5 * p.dt * p.dt;
    const double lm_s = c.lm_scaled_by_dt ? p.dt * c.lm : c.lm;   // acados-version dependent
    // v order [u_a,u_alpha,x,y,psi,v,omega]; y order [x,y,v,omega,u_a,u_alpha]; stage cost scaled by dt
    const double Wv[7] = {c.W[4], c.W[5], c.W[0], c.W[1], 0.0, c.W[2], c.W[3]};
    const double Wev[7] = {0.0, 0.0, c.W_e[0], c.W_e[1], 0.0, c.W_e[2], c.W_e[3]};
    for (int i = 0; i < 7; i++) {
        p.Hd[i] = p.dt * Wv[i] + lm_s; p.Wg[i] = p.dt * Wv[i];
        p.Hde[i] = i < 2 ? 0.0 : Wev[i] + c.lm; p.Wge[i] = Wev[i];
    }
    for (int i = 0; i < 4; i++) { p.lbx[i] = c.lbx[i]; p.ubx[i] = c.ubx[i]; }
    for (int i = 0; i < 2; i++) { p.lbu[i] = c.lbu[i]; p.ubu[i] = c.ubu[i]; }
    for (int i = 0; i < 6; i++) { p.lb6[i] = i < 2 ? c.lbu[i] : c.lbx[i - 2]; p.ub6[i] = i < 2 ? c.ubu[i] : c.ubx[i - 2]; }
    p.r2 = c.r2; p.lh = c.lh; p.uh = c.uh; p.zu = c.zu; p.Zu = c.Zu;
    p.mu0 = c.mu0; p.alpha_min = c.alpha_min; p.res_g_max = c.res_g_max; p.res_b_max = c.res_b_max;
    p.res_d_max = c.res_d_max; p.res_m_max = c.res_m_max; p.reg_prim = c.reg_prim;
    p.lam_min = c.lam_min; p.t_min = c.t_min; p.tau_min = c.tau_min; p.thr0 = c.thr0;
}
