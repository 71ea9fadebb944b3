// mpc_tpp.cuh -- thread-per-problem fused SQP-RTI kernel (generic N, M).
//
// One CUDA thread advances one independent MPC problem through the whole call the reference
// issues at src/simulation/robot_ocp_problem.py:195 (acados SQP_RTI): preparation
// (IRK-GL4 rollout + sensitivities, obstacle constraints + gradients, Gauss-Newton cost),
// the HPIPM-equivalent Mehrotra predictor-corrector interior point on a structure-exploiting
// Riccati recursion, and the full-step update.  All per-problem state lives in batch-last SoA
// global memory (coalesced across the warp, L2-resident for B ~ 4k); the 7x7 stage algebra
// runs in registers.  This is the generic path (any N <= 128, M <= 32) and the on-device
// reference for the warp-cooperative kernel in mpc_wpp.cuh.
#pragma once
#include "mpc_common.cuh"

struct B2Off {
    int sv, b, g, cx, cy, lg, ug, zlq;                 // per-solve constants
    int v, sl, su;                                     // QP primal iterate (pi, lam, t live in the output arrays)
    int dv, dsl, dsu, dpi, dlam, dt;                   // step
    int rg, rgsl, rgsu, rb, rd, dd, tinv;              // residuals, predictor product, 1/t
    int lrg, lrgsl, lrgsu, lrb;                        // linear-system residual (iterative refinement)
    int lamu, lxu, pm, pv, lu, dli, dui, rhol, rhou;   // Riccati factors and slack elimination
    int total;
    __host__ __device__ explicit B2Off(int M)
    {
        const int nct = 12 + 4 * M;
        int o = 0;
        sv = o; o += 10; b = o; o += 5; g = o; o += 7; cx = o; o += M; cy = o; o += M;
        lg = o; o += M; ug = o; o += M; zlq = o; o += M;
        v = o; o += 7; sl = o; o += M; su = o; o += M;
        dv = o; o += 7; dsl = o; o += M; dsu = o; o += M; dpi = o; o += 5; dlam = o; o += nct; dt = o; o += nct;
        rg = o; o += 7; rgsl = o; o += M; rgsu = o; o += M; rb = o; o += 5; rd = o; o += nct; dd = o; o += nct; tinv = o; o += nct;
        lrg = o; o += 7; lrgsl = o; o += M; lrgsu = o; o += M; lrb = o; o += 5;
        lamu = o; o += 3; lxu = o; o += 10; pm = o; o += 15; pv = o; o += 5; lu = o; o += 2;
        dli = o; o += M; dui = o; o += M; rhol = o; o += M; rhou = o; o += M;
        total = o;
    }
};

template <int MT>
struct TppSolver {
    const B2Params& p;
    const int b;
    const int N, M, nct;
    const B2Off o;
    const size_t Bm;
    const int o_lb, o_lg, o_ub, o_ug, o_ls, o_us;
    // running reductions
    double mu, nrm[4], sum_lt;
    double alpha, S1, S2;
    int nc;

    __device__ TppSolver(const B2Params& p_, int b_)
        : p(p_), b(b_), N(p_.N), M(MT > 0 ? MT : p_.M), nct(12 + 4 * (MT > 0 ? MT : p_.M)), o(MT > 0 ? MT : p_.M),
          Bm((size_t)p_.Bm), o_lb(0), o_lg(6), o_ub(6 + (MT > 0 ? MT : p_.M)), o_ug(12 + (MT > 0 ? MT : p_.M)),
          o_ls(12 + 2 * (MT > 0 ? MT : p_.M)), o_us(12 + 3 * (MT > 0 ? MT : p_.M))
    {
        nc = 0;
        for (int k = 0; k <= N; k++) nc += 2 * nbox(k) + 4 * M;
    }

    // ---- accessors -------------------------------------------------------------------------
    __device__ __forceinline__ double& ws(int k, int off) const { return p.WS[((size_t)k * p.slots_per_stage + off) * Bm + b]; }
    __device__ __forceinline__ double& lam(int k, int i) const { return p.LAM[((size_t)k * nct + i) * Bm + b]; }
    __device__ __forceinline__ double& tt(int k, int i) const { return p.T[((size_t)k * nct + i) * Bm + b]; }
    __device__ __forceinline__ double& pi(int k, int i) const { return p.PI[((size_t)k * 5 + i) * Bm + b]; }
    __device__ __forceinline__ double X(int k, int i) const { return p.X[((size_t)k * 5 + i) * Bm + b]; }
    __device__ __forceinline__ double U(int k, int i) const { return p.U[((size_t)k * 2 + i) * Bm + b]; }
    __device__ __forceinline__ double SLi(int k, int j) const { return p.SL[((size_t)k * M + j) * Bm + b]; }
    __device__ __forceinline__ double SUi(int k, int j) const { return p.SU[((size_t)k * M + j) * Bm + b]; }
    __device__ __forceinline__ int nbox(int k) const { return (k < N ? 2 : 0) + ((k > 0 && k < N) ? 4 : 0); }
    __device__ __forceinline__ bool box_on(int s, int k) const { return s < 2 ? (k < N) : (k > 0 && k < N); }
    __device__ __forceinline__ double scal(int k) const { return k < N ? p.dt : 1.0; }
    __device__ __forceinline__ void box_bounds(int s, int k, double& lb, double& ub) const
    {
        if (s < 2) { const double uk = U(k, s); lb = p.lbu[s] - uk; ub = p.ubu[s] - uk; }
        else { const double xk = X(k, b2_slot2v(s) - 2); lb = p.lbx[s - 2] - xk; ub = p.ubx[s - 2] - xk; }
    }
    __device__ __forceinline__ void load_sv(int k, double* sv) const
    {
#pragma unroll
        for (int i = 0; i < 10; i++) sv[i] = ws(k, o.sv + i);
    }

    // ---- RTI preparation + HPIPM cold-start initialisation ------------
    __device__ void prepare_and_init()
    {
        double dx0[5];
#pragma unroll
        for (int i = 0; i < 5; i++) dx0[i] = p.X0[(size_t)i * Bm + b] - X(0, i);
        for (int k = 0; k <= N; k++) {
            double xk[5], uk[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 5; i++) xk[i] = X(k, i);
            const double sc = scal(k);
            const double* yr = p.YREF + ((size_t)k * 6) * Bm + b;
            if (k < N) {
                uk[0] = U(k, 0); uk[1] = U(k, 1);
                double xn[5], sv[10];
                if (p.integrator == 1) model_rk4_step(xk, uk, p.dt, xn, sv);
                else model_gl4_step(xk, uk, p.dt, xn, sv);
#pragma unroll
                for (int i = 0; i < 10; i++) ws(k, o.sv + i) = sv[i];
#pragma unroll
                for (int i = 0; i < 5; i++) ws(k, o.b + i) = xn[i] - X(k + 1, i);
                // Gauss-Newton gradient of the LINEAR_LS stage cost (robot_ocp_problem.py:59-81)
                ws(k, o.g + 0) = p.Wg[0] * (uk[0] - yr[4 * Bm]);
                ws(k, o.g + 1) = p.Wg[1] * (uk[1] - yr[5 * Bm]);
                ws(k, o.g + 2) = p.Wg[2] * (xk[0] - yr[0 * Bm]);
                ws(k, o.g + 3) = p.Wg[3] * (xk[1] - yr[1 * Bm]);
                ws(k, o.g + 4) = 0.0;
                ws(k, o.g + 5) = p.Wg[5] * (xk[3] - yr[2 * Bm]);
                ws(k, o.g + 6) = p.Wg[6] * (xk[4] - yr[3 * Bm]);
            } else {
                ws(k, o.g + 0) = 0.0; ws(k, o.g + 1) = 0.0;
                ws(k, o.g + 2) = p.Wge[2] * (xk[0] - yr[0 * Bm]);
                ws(k, o.g + 3) = p.Wge[3] * (xk[1] - yr[1 * Bm]);
                ws(k, o.g + 4) = 0.0;
                ws(k, o.g + 5) = p.Wge[5] * (xk[3] - yr[2 * Bm]);
                ws(k, o.g + 6) = p.Wge[6] * (xk[4] - yr[3 * Bm]);
            }
            // initial QP point: v = 0 (node 0 carries the embedded dx0 = x0bar - x_0)
            double v[7] = {0, 0, 0, 0, 0, 0, 0};
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) v[2 + i] = dx0[i];
            }
            for (int i = 0; i < nct; i++) { lam(k, i) = 0.0; tt(k, i) = 0.0; }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                double lb, ub;
                box_bounds(s, k, lb, ub);
                const int iv = b2_slot2v(s);
                double tl = -lb + v[iv], tu = ub - v[iv];
                if (tl < p.thr0) {
                    if (tu < p.thr0) { v[iv] = 0.5 * (lb + ub); tl = p.thr0; tu = p.thr0; }
                    else { tl = p.thr0; v[iv] = lb + p.thr0; }
                } else if (tu < p.thr0) { tu = p.thr0; v[iv] = ub - p.thr0; }
                tt(k, o_lb + s) = tl; tt(k, o_ub + s) = tu;
                lam(k, o_lb + s) = p.mu0 / tl; lam(k, o_ub + s) = p.mu0 / tu;
            }
            // obstacle constraints (robot_model.py:60-65) and their soft bounds
            for (int j = 0; j < M; j++) {
                double hval, cx, cy;
                model_con_h(xk, p.P[((size_t)k * 2 * M + 2 * j) * Bm + b], p.P[((size_t)k * 2 * M + 2 * j + 1) * Bm + b], p.r2, &hval, &cx, &cy);
                const double s_l = SLi(k, j), s_u = SUi(k, j);
                const double lg = p.lh - hval - s_l, ug = p.uh - hval + s_u;
                ws(k, o.cx + j) = cx; ws(k, o.cy + j) = cy;
                ws(k, o.lg + j) = lg; ws(k, o.ug + j) = ug;
                const double zl_in = p.ZL[((size_t)k * M + j) * Bm + b], Zl_in = p.ZZL[((size_t)k * M + j) * Bm + b];
                ws(k, o.zlq + j) = sc * (zl_in + Zl_in * s_l);
                const double cv = cx * v[2] + cy * v[3];
                double tl = cv - lg, tu = ug - cv;
                tl = p.thr0 > tl ? p.thr0 : tl;
                tu = p.thr0 > tu ? p.thr0 : tu;
                tt(k, o_lg + j) = tl; tt(k, o_ug + j) = tu;
                lam(k, o_lg + j) = p.mu0 / tl; lam(k, o_ug + j) = p.mu0 / tu;
                tt(k, o_ls + j) = 1.0; tt(k, o_us + j) = 1.0;
                lam(k, o_ls + j) = p.mu0; lam(k, o_us + j) = p.mu0;
                ws(k, o.sl + j) = 0.0; ws(k, o.su + j) = 0.0;
            }
#pragma unroll
            for (int i = 0; i < 7; i++) ws(k, o.v + i) = v[i];
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 5; i++) pi(k, i) = 0.0;
            }
        }
    }

    // ---- KKT residuals (HPIPM d_ocp_qp_res_compute); rd = g_i(v,s) - t_i ----------------------
    __device__ void residuals()
    {
        double sum = 0.0;
        nrm[0] = nrm[1] = nrm[2] = nrm[3] = 0.0;
        for (int k = 0; k <= N; k++) {
            double v[7], rg[7];
            const double* Hd = k < N ? p.Hd : p.Hde;
            const double sc = scal(k);
#pragma unroll
            for (int i = 0; i < 7; i++) { v[i] = ws(k, o.v + i); rg[i] = Hd[i] * v[i] + ws(k, o.g + i); }
            double sv[10];
            if (k < N) {
                load_sv(k, sv);
                double pk[5], bt[7];
#pragma unroll
                for (int i = 0; i < 5; i++) pk[i] = pi(k, i);
                b2_bat_mul(sv, p.dt, p.dt2h, pk, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) rg[i] += bt[i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) rg[2 + i] -= pi(k - 1, i);
            }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                double lb, ub;
                box_bounds(s, k, lb, ub);
                const int iv = b2_slot2v(s);
                const double ll = lam(k, o_lb + s), lu_ = lam(k, o_ub + s), tl = tt(k, o_lb + s), tu = tt(k, o_ub + s);
                rg[iv] += -ll + lu_;
                const double rdl = v[iv] - lb - tl, rdu = ub - v[iv] - tu;
                ws(k, o.rd + o_lb + s) = rdl; ws(k, o.rd + o_ub + s) = rdu;
                nrm[2] = fmax(nrm[2], fmax(fabs(rdl), fabs(rdu)));
                const double ml = ll * tl, mu_ = lu_ * tu;
                nrm[3] = fmax(nrm[3], fmax(fabs(ml), fabs(mu_)));
                sum += ml; sum += mu_;
            }
            const double Zuq = sc * p.Zu;
            for (int j = 0; j < M; j++) {
                const double cx = ws(k, o.cx + j), cy = ws(k, o.cy + j);
                const double cv = cx * v[2] + cy * v[3];
                const double llg = lam(k, o_lg + j), lug = lam(k, o_ug + j), lls = lam(k, o_ls + j), lus = lam(k, o_us + j);
                const double tlg = tt(k, o_lg + j), tug = tt(k, o_ug + j), tls = tt(k, o_ls + j), tus = tt(k, o_us + j);
                const double dl = lug - llg;
                rg[2] += cx * dl; rg[3] += cy * dl;
                const double s_l = ws(k, o.sl + j), s_u = ws(k, o.su + j);
                const double sli = SLi(k, j), sui = SUi(k, j);
                const double Zlq = sc * p.ZZL[((size_t)k * M + j) * Bm + b];
                const double rgsl = Zlq * s_l + ws(k, o.zlq + j) - llg - lls;
                const double rgsu = Zuq * s_u + sc * (p.zu + p.Zu * sui) - lug - lus;
                ws(k, o.rgsl + j) = rgsl; ws(k, o.rgsu + j) = rgsu;
                nrm[0] = fmax(nrm[0], fmax(fabs(rgsl), fabs(rgsu)));
                const double rdlg = cv + s_l - ws(k, o.lg + j) - tlg;
                const double rdug = ws(k, o.ug + j) - cv + s_u - tug;
                const double rdls = s_l + sli - tls;       // ls = -sl_iter
                const double rdus = s_u + sui - tus;
                ws(k, o.rd + o_lg + j) = rdlg; ws(k, o.rd + o_ug + j) = rdug;
                ws(k, o.rd + o_ls + j) = rdls; ws(k, o.rd + o_us + j) = rdus;
                nrm[2] = fmax(nrm[2], fmax(fmax(fabs(rdlg), fabs(rdug)), fmax(fabs(rdls), fabs(rdus))));
                const double m0 = llg * tlg, m1 = lug * tug, m2 = lls * tls, m3 = lus * tus;
                nrm[3] = fmax(nrm[3], fmax(fmax(fabs(m0), fabs(m1)), fmax(fabs(m2), fabs(m3))));
                sum += m0; sum += m1; sum += m2; sum += m3;
            }
            if (k < N) {
                double rb[5];
                b2_ba_mul(sv, p.dt, p.dt2h, v, rb);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    rb[i] += ws(k, o.b + i) - ws(k + 1, o.v + 2 + i);
                    ws(k, o.rb + i) = rb[i];
                    nrm[1] = fmax(nrm[1], fabs(rb[i]));
                }
            }
            if (k == 0) { rg[2] = rg[3] = rg[4] = rg[5] = rg[6] = 0.0; }      // x_0 is not a variable
            if (k == N) { rg[0] = rg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) { ws(k, o.rg + i) = rg[i]; nrm[0] = fmax(nrm[0], fabs(rg[i])); }
        }
        sum_lt = sum;
        mu = sum / nc;
    }

    // ---- condensed Hessians + backward Riccati factorisation -------------
    __device__ void factorize()
    {
        double Pn[15];
        for (int k = N; k >= 0; k--) {
            const double* Hd = k < N ? p.Hd : p.Hde;
            const double sc = scal(k);
            double Hx[7], hxy = 0.0;
#pragma unroll
            for (int i = 0; i < 7; i++) Hx[i] = Hd[i] + p.reg_prim;
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const double til = 1.0 / tt(k, o_lb + s), tiu = 1.0 / tt(k, o_ub + s);
                ws(k, o.tinv + o_lb + s) = til; ws(k, o.tinv + o_ub + s) = tiu;
                Hx[b2_slot2v(s)] += lam(k, o_lb + s) * til + lam(k, o_ub + s) * tiu;
            }
            const double Zuq = sc * p.Zu;
            for (int j = 0; j < M; j++) {
                const double t0 = 1.0 / tt(k, o_lg + j), t1 = 1.0 / tt(k, o_ug + j), t2 = 1.0 / tt(k, o_ls + j), t3 = 1.0 / tt(k, o_us + j);
                ws(k, o.tinv + o_lg + j) = t0; ws(k, o.tinv + o_ug + j) = t1;
                ws(k, o.tinv + o_ls + j) = t2; ws(k, o.tinv + o_us + j) = t3;
                const double Gl = lam(k, o_lg + j) * t0, Gu = lam(k, o_ug + j) * t1, Gls = lam(k, o_ls + j) * t2, Gus = lam(k, o_us + j) * t3;
                const double Zlq = sc * p.ZZL[((size_t)k * M + j) * Bm + b];
                const double Dli = 1.0 / (Zlq + Gl + Gls), Dui = 1.0 / (Zuq + Gu + Gus);
                ws(k, o.dli + j) = Dli; ws(k, o.dui + j) = Dui;
                const double weff = Gl * (Zlq + Gls) * Dli + Gu * (Zuq + Gus) * Dui;
                const double cx = ws(k, o.cx + j), cy = ws(k, o.cy + j);
                Hx[2] += weff * cx * cx; Hx[3] += weff * cy * cy; hxy += weff * cx * cy;
            }
            // square-root Riccati (HPIPM square_root_alg = 1): Pn holds Lxx, P = Lxx Lxx^T is never formed
            // by subtraction (barrier weights reach 1e15+ near convergence; DESIGN.md "square-root Riccati")
            if (k == N) {
#pragma unroll
                for (int i = 0; i < 15; i++) Pn[i] = 0.0;
#pragma unroll
                for (int i = 0; i < 5; i++) Pn[b2_sym(i, i)] = Hx[2 + i];
                Pn[b2_sym(1, 0)] = hxy;
                b2_chol_packed<5>(Pn);
            } else {
                double sv[10], Mp[28];
                load_sv(k, sv);
                b2_bat_l_sq(sv, p.dt, p.dt2h, Pn, Mp);
#pragma unroll
                for (int i = 0; i < 7; i++) Mp[i * (i + 1) / 2 + i] += Hx[i];
                Mp[3 * 4 / 2 + 2] += hxy;
                b2_chol_packed<7>(Mp);
                const double l00i = Mp[0] > 0.0 ? 1.0 / Mp[0] : 0.0, l11i = Mp[2] > 0.0 ? 1.0 / Mp[2] : 0.0;
                ws(k, o.lamu + 0) = l00i; ws(k, o.lamu + 1) = Mp[1]; ws(k, o.lamu + 2) = l11i;
                if (k > 0) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const int r = 2 + i;
                        ws(k, o.lxu + 2 * i) = Mp[r * (r + 1) / 2 + 0]; ws(k, o.lxu + 2 * i + 1) = Mp[r * (r + 1) / 2 + 1];
                    }
#pragma unroll
                    for (int i = 0; i < 5; i++)
#pragma unroll
                        for (int j = 0; j <= i; j++) {
                            const int r = 2 + i, c = 2 + j;
                            Pn[b2_sym(i, j)] = Mp[r * (r + 1) / 2 + c];
                        }
                }
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 15; i++) ws(k, o.pm + i) = Pn[i];
            }
        }
    }

    // complementarity right-hand side by solve kind: 0 affine, 1 corrector, 2 centering, 3 refinement
    __device__ __forceinline__ double rm_of(int kind, double l, double t, double dd, double smu) const
    {
        if (kind == 0) return l * t;
        if (kind == 1) return l * t + dd - smu;
        if (kind == 2) return l * t - smu;
        return 0.0;
    }

    // ---- solve the KKT system with the current factorisation, expand, and reduce alpha ---------
    // kind 3 = iterative refinement: right-hand side is the linear residual, result is ADDED.
    __device__ void solve(int kind, double smu)
    {
        const bool itr = (kind == 3);
        const int oRG = itr ? o.lrg : o.rg, oRGSL = itr ? o.lrgsl : o.rgsl, oRGSU = itr ? o.lrgsu : o.rgsu, oRB = itr ? o.lrb : o.rb;
        // backward vector sweep
        double pn[5] = {0, 0, 0, 0, 0};
        for (int k = N; k >= 0; k--) {
            const double sc = scal(k);
            double gt[7];
#pragma unroll
            for (int i = 0; i < 7; i++) gt[i] = ws(k, oRG + i);
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const int il = o_lb + s, iu = o_ub + s;
                const double ll = lam(k, il), lu_ = lam(k, iu);
                const double rml = rm_of(kind, ll, tt(k, il), itr ? 0.0 : ws(k, o.dd + il), smu);
                const double rmu = rm_of(kind, lu_, tt(k, iu), itr ? 0.0 : ws(k, o.dd + iu), smu);
                const double rdl = itr ? 0.0 : ws(k, o.rd + il), rdu = itr ? 0.0 : ws(k, o.rd + iu);
                const double ql = (rml + ll * rdl) * ws(k, o.tinv + il);
                const double qu = (rmu + lu_ * rdu) * ws(k, o.tinv + iu);
                gt[b2_slot2v(s)] += ql - qu;
            }
            (void)sc;
            for (int j = 0; j < M; j++) {
                const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
                const double l0 = lam(k, i0), l1 = lam(k, i1), l2 = lam(k, i2), l3 = lam(k, i3);
                const double ti0 = ws(k, o.tinv + i0), ti1 = ws(k, o.tinv + i1), ti2 = ws(k, o.tinv + i2), ti3 = ws(k, o.tinv + i3);
                double q0, q1, q2, q3;
                if (itr) { q0 = q1 = q2 = q3 = 0.0; }
                else {
                    q0 = (rm_of(kind, l0, tt(k, i0), ws(k, o.dd + i0), smu) + l0 * ws(k, o.rd + i0)) * ti0;
                    q1 = (rm_of(kind, l1, tt(k, i1), ws(k, o.dd + i1), smu) + l1 * ws(k, o.rd + i1)) * ti1;
                    q2 = (rm_of(kind, l2, tt(k, i2), ws(k, o.dd + i2), smu) + l2 * ws(k, o.rd + i2)) * ti2;
                    q3 = (rm_of(kind, l3, tt(k, i3), ws(k, o.dd + i3), smu) + l3 * ws(k, o.rd + i3)) * ti3;
                }
                const double rhol = ws(k, oRGSL + j) + q0 + q2;
                const double rhou = ws(k, oRGSU + j) + q1 + q3;
                ws(k, o.rhol + j) = rhol; ws(k, o.rhou + j) = rhou;
                const double coef = (q0 - l0 * ti0 * rhol * ws(k, o.dli + j)) - (q1 - l1 * ti1 * rhou * ws(k, o.dui + j));
                gt[2] += ws(k, o.cx + j) * coef;
                gt[3] += ws(k, o.cy + j) * coef;
            }
            if (k < N) {
                double sv[10], Pn[15], rb[5], y[5], bt[7];
                load_sv(k, sv);
#pragma unroll
                for (int i = 0; i < 15; i++) Pn[i] = ws(k + 1, o.pm + i);
#pragma unroll
                for (int i = 0; i < 5; i++) rb[i] = ws(k, oRB + i);
                b2_llt_mul(Pn, rb, y);
#pragma unroll
                for (int i = 0; i < 5; i++) y[i] += pn[i];
                b2_bat_mul(sv, p.dt, p.dt2h, y, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) gt[i] += bt[i];
                const double l00i = ws(k, o.lamu + 0), l10 = ws(k, o.lamu + 1), l11i = ws(k, o.lamu + 2);
                const double lu0 = gt[0] * l00i;
                const double lu1 = (gt[1] - l10 * lu0) * l11i;
                ws(k, o.lu + 0) = lu0; ws(k, o.lu + 1) = lu1;
                if (k > 0) {
#pragma unroll
                    for (int i = 0; i < 5; i++) pn[i] = gt[2 + i] - ws(k, o.lxu + 2 * i) * lu0 - ws(k, o.lxu + 2 * i + 1) * lu1;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 5; i++) pn[i] = gt[2 + i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) ws(k, o.pv + i) = pn[i];
            }
        }
        // forward sweep + expansion + step-length / mu_aff reductions
        double dx[5] = {0, 0, 0, 0, 0};
        double ap = -1.0, ad = -1.0;
        S1 = 0.0; S2 = 0.0;
        for (int k = 0; k <= N; k++) {
            double dv[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 5; i++) dv[2 + i] = dx[i];
            double sv[10];
            if (k < N) {
                load_sv(k, sv);
                const double l00i = ws(k, o.lamu + 0), l10 = ws(k, o.lamu + 1), l11i = ws(k, o.lamu + 2);
                double r0 = ws(k, o.lu + 0), r1 = ws(k, o.lu + 1);
                if (k > 0) {
#pragma unroll
                    for (int i = 0; i < 5; i++) { r0 += ws(k, o.lxu + 2 * i) * dx[i]; r1 += ws(k, o.lxu + 2 * i + 1) * dx[i]; }
                }
                r0 = -r0; r1 = -r1;
                dv[1] = r1 * l11i;
                dv[0] = (r0 - l10 * dv[1]) * l00i;
                double xn[5], Pn[15], dpi[5];
                b2_ba_mul(sv, p.dt, p.dt2h, dv, xn);
#pragma unroll
                for (int i = 0; i < 5; i++) xn[i] += ws(k, oRB + i);
#pragma unroll
                for (int i = 0; i < 15; i++) Pn[i] = ws(k + 1, o.pm + i);
                b2_llt_mul(Pn, xn, dpi);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    dpi[i] += ws(k + 1, o.pv + i);
                    if (itr) ws(k, o.dpi + i) += dpi[i]; else ws(k, o.dpi + i) = dpi[i];
                    dx[i] = xn[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 7; i++) { if (itr) ws(k, o.dv + i) += dv[i]; else ws(k, o.dv + i) = dv[i]; }
            // box constraints
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const int iv = b2_slot2v(s);
                const int idx[2] = {o_lb + s, o_ub + s};
                const double sg[2] = {1.0, -1.0};
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int i = idx[q];
                    const double l = lam(k, i), t = tt(k, i);
                    const double rd = itr ? 0.0 : ws(k, o.rd + i);
                    const double rm = rm_of(kind, l, t, itr ? 0.0 : ws(k, o.dd + i), smu);
                    double dti = sg[q] * dv[iv] + rd;
                    double dli = -(rm + l * dti) * ws(k, o.tinv + i);
                    finish_constraint(kind, k, i, l, t, dti, dli, ap, ad);
                }
            }
            for (int j = 0; j < M; j++) {
                const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
                const double cdv = ws(k, o.cx + j) * dv[2] + ws(k, o.cy + j) * dv[3];
                const double l0 = lam(k, i0), l1 = lam(k, i1), l2 = lam(k, i2), l3 = lam(k, i3);
                const double t0 = tt(k, i0), t1 = tt(k, i1), t2 = tt(k, i2), t3 = tt(k, i3);
                const double ti0 = ws(k, o.tinv + i0), ti1 = ws(k, o.tinv + i1), ti2 = ws(k, o.tinv + i2), ti3 = ws(k, o.tinv + i3);
                const double dsl = -(l0 * ti0 * cdv + ws(k, o.rhol + j)) * ws(k, o.dli + j);
                const double dsu = -(-l1 * ti1 * cdv + ws(k, o.rhou + j)) * ws(k, o.dui + j);
                if (itr) { ws(k, o.dsl + j) += dsl; ws(k, o.dsu + j) += dsu; }
                else { ws(k, o.dsl + j) = dsl; ws(k, o.dsu + j) = dsu; }
                double rd0 = 0, rd1 = 0, rd2 = 0, rd3 = 0, rm0 = 0, rm1 = 0, rm2 = 0, rm3 = 0;
                if (!itr) {
                    rd0 = ws(k, o.rd + i0); rd1 = ws(k, o.rd + i1); rd2 = ws(k, o.rd + i2); rd3 = ws(k, o.rd + i3);
                    rm0 = rm_of(kind, l0, t0, ws(k, o.dd + i0), smu); rm1 = rm_of(kind, l1, t1, ws(k, o.dd + i1), smu);
                    rm2 = rm_of(kind, l2, t2, ws(k, o.dd + i2), smu); rm3 = rm_of(kind, l3, t3, ws(k, o.dd + i3), smu);
                }
                const double dt0 = cdv + dsl + rd0, dt1 = -cdv + dsu + rd1, dt2 = dsl + rd2, dt3 = dsu + rd3;
                finish_constraint(kind, k, i0, l0, t0, dt0, -(rm0 + l0 * dt0) * ti0, ap, ad);
                finish_constraint(kind, k, i1, l1, t1, dt1, -(rm1 + l1 * dt1) * ti1, ap, ad);
                finish_constraint(kind, k, i2, l2, t2, dt2, -(rm2 + l2 * dt2) * ti2, ap, ad);
                finish_constraint(kind, k, i3, l3, t3, dt3, -(rm3 + l3 * dt3) * ti3, ap, ad);
            }
        }
        const double a = ap > ad ? ap : ad;
        alpha = -a;
    }

    // store (or accumulate) dlam/dt of one constraint and fold it into the reductions
    __device__ __forceinline__ void finish_constraint(int kind, int k, int i, double l, double t, double dti, double dli, double& ap, double& ad)
    {
        if (kind == 3) {
            dti += ws(k, o.dt + i);
            dli += ws(k, o.dlam + i);
        }
        ws(k, o.dt + i) = dti;
        ws(k, o.dlam + i) = dli;
        if (kind == 0) ws(k, o.dd + i) = dti * dli;
        // HPIPM d_compute_alpha_qp (negative-reciprocal convention)
        if (ad * dli > l) ad = l / dli;
        if (ap * dti > t) ap = t / dti;
        S1 += l * dti + t * dli;
        S2 += dli * dti;
    }

    // ---- residual of the linear system at the current step (HPIPM d_ocp_qp_res_compute_lin) -----
    __device__ void lin_residuals(double lnrm[2])
    {
        lnrm[0] = lnrm[1] = 0.0;
        for (int k = 0; k <= N; k++) {
            const double* Hd = k < N ? p.Hd : p.Hde;
            const double sc = scal(k);
            double dv[7], lrg[7];
#pragma unroll
            for (int i = 0; i < 7; i++) { dv[i] = ws(k, o.dv + i); lrg[i] = ws(k, o.rg + i) + Hd[i] * dv[i]; }
            double sv[10];
            if (k < N) {
                load_sv(k, sv);
                double dp[5], bt[7];
#pragma unroll
                for (int i = 0; i < 5; i++) dp[i] = ws(k, o.dpi + i);
                b2_bat_mul(sv, p.dt, p.dt2h, dp, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) lrg[i] += bt[i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) lrg[2 + i] -= ws(k - 1, o.dpi + i);
            }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                lrg[b2_slot2v(s)] += -ws(k, o.dlam + o_lb + s) + ws(k, o.dlam + o_ub + s);
            }
            const double Zuq = sc * p.Zu;
            for (int j = 0; j < M; j++) {
                const double dl0 = ws(k, o.dlam + o_lg + j), dl1 = ws(k, o.dlam + o_ug + j);
                const double dl2 = ws(k, o.dlam + o_ls + j), dl3 = ws(k, o.dlam + o_us + j);
                const double dl = dl1 - dl0;
                lrg[2] += ws(k, o.cx + j) * dl; lrg[3] += ws(k, o.cy + j) * dl;
                const double Zlq = sc * p.ZZL[((size_t)k * M + j) * Bm + b];
                const double a0 = Zlq * ws(k, o.dsl + j) + ws(k, o.rgsl + j) - dl0 - dl2;
                const double a1 = Zuq * ws(k, o.dsu + j) + ws(k, o.rgsu + j) - dl1 - dl3;
                ws(k, o.lrgsl + j) = a0; ws(k, o.lrgsu + j) = a1;
                lnrm[0] = fmax(lnrm[0], fmax(fabs(a0), fabs(a1)));
            }
            if (k < N) {
                double lrb[5];
                b2_ba_mul(sv, p.dt, p.dt2h, dv, lrb);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    lrb[i] += ws(k, o.rb + i) - ws(k + 1, o.dv + 2 + i);
                    ws(k, o.lrb + i) = lrb[i];
                    lnrm[1] = fmax(lnrm[1], fabs(lrb[i]));
                }
            }
            if (k == 0) { lrg[2] = lrg[3] = lrg[4] = lrg[5] = lrg[6] = 0.0; }
            if (k == N) { lrg[0] = lrg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) { ws(k, o.lrg + i) = lrg[i]; lnrm[0] = fmax(lnrm[0], fabs(lrg[i])); }
        }
    }

    // ---- HPIPM d_update_var_qp ----------------------------------------------------------------
    __device__ void update(double a)
    {
        for (int k = 0; k <= N; k++) {
#pragma unroll
            for (int i = 0; i < 7; i++) ws(k, o.v + i) += a * ws(k, o.dv + i);
            for (int j = 0; j < M; j++) { ws(k, o.sl + j) += a * ws(k, o.dsl + j); ws(k, o.su + j) += a * ws(k, o.dsu + j); }
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 5; i++) pi(k, i) += a * ws(k, o.dpi + i);
            }
            for (int i = 0; i < nct; i++) {
                if (i < 6 ? !box_on(i, k) : (i >= o_ub && i < o_ub + 6 ? !box_on(i - o_ub, k) : false)) continue;
                double l = lam(k, i) + a * ws(k, o.dlam + i);
                double t = tt(k, i) + a * ws(k, o.dt + i);
                if (l <= p.lam_min) l = p.lam_min;
                if (t <= p.t_min) t = p.t_min;
                lam(k, i) = l; tt(k, i) = t;
            }
        }
    }

    // ---- the whole RTI call ---------------------------------------------------------------------
    __device__ void run()
    {
        prepare_and_init();
        residuals();
        alpha = 1.0;
        int kk = 0, n_itref = 0;
        const double inv_nc = 1.0 / nc;
        for (; kk < p.qp_iter_max && alpha > p.alpha_min &&
               (nrm[0] > p.res_g_max || nrm[1] > p.res_b_max || nrm[2] > p.res_d_max || nrm[3] > p.res_m_max); kk++) {
            factorize();
            solve(0, 0.0);
            if (p.pred_corr) {
                double mu_aff = (sum_lt + alpha * S1 + alpha * alpha * S2) * inv_nc;
                const double tmp = mu_aff / mu;
                double smu = tmp * tmp * tmp * mu;
                smu = smu > p.tau_min ? smu : p.tau_min;
                solve(1, smu);
                if (p.cond_pred_corr) {
                    const double mu_aff0 = mu_aff;
                    mu_aff = (sum_lt + alpha * S1 + alpha * alpha * S2) * inv_nc;
                    if (mu_aff > 2.0 * mu_aff0) solve(2, smu);
                }
                for (int it = 0; it < p.itref_corr_max; it++) {
                    double lnrm[2];
                    lin_residuals(lnrm);
                    if (lnrm[0] < p.res_g_max && lnrm[1] < p.res_b_max) break;
                    solve(3, 0.0);
                    n_itref++;
                }
            }
            double a = alpha;
            if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
            update(a);
            residuals();
        }
        int qs;
        if (kk == p.qp_iter_max) qs = 1;
        else if (alpha <= p.alpha_min) qs = 2;
        else if (mu != mu || nrm[0] != nrm[0] || nrm[1] != nrm[1] || nrm[2] != nrm[2] || nrm[3] != nrm[3]) qs = 3;
        else qs = 0;
        int status = 0;
        if (qs != 0 && qs != 1) status = 4;     // ACADOS_QP_FAILURE: iterate untouched
        else {
            // full step (SQP_RTI): x += dx, u += du, s += ds
            for (int k = 0; k <= N; k++) {
#pragma unroll
                for (int i = 0; i < 5; i++) p.X[((size_t)k * 5 + i) * Bm + b] += ws(k, o.v + 2 + i);
                if (k < N) {
                    p.U[((size_t)k * 2 + 0) * Bm + b] += ws(k, o.v + 0);
                    p.U[((size_t)k * 2 + 1) * Bm + b] += ws(k, o.v + 1);
                }
                for (int j = 0; j < M; j++) {
                    p.SL[((size_t)k * M + j) * Bm + b] += ws(k, o.sl + j);
                    p.SU[((size_t)k * M + j) * Bm + b] += ws(k, o.su + j);
                }
            }
        }
        p.STATUS[b] = status; p.QP_ITER[b] = kk; p.QP_STATUS[b] = qs;
        p.STATS[0 * Bm + b] = nrm[0]; p.STATS[1 * Bm + b] = nrm[1]; p.STATS[2 * Bm + b] = nrm[2]; p.STATS[3 * Bm + b] = nrm[3];
        p.STATS[4 * Bm + b] = mu; p.STATS[5 * Bm + b] = alpha; p.STATS[6 * Bm + b] = (double)n_itref; p.STATS[7 * Bm + b] = 0.0;
    }
};

template <int MT>
__global__ void __launch_bounds__(64) b2_rti_solve_tpp(const B2Params p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    TppSolver<MT> s(p, b);
    s.run();
}
