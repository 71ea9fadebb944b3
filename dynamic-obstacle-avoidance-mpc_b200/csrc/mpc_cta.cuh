// mpc_cta.cuh -- CTA-per-problem fused SQP-RTI kernel: the LATENCY-oriented mapping of the same solve as mpc_wpp.cuh
// (acados SQP_RTI + HPIPM-style Mehrotra IPM behind src/simulation/robot_ocp_problem.py:195).
//
// mpc_wpp.cuh gives a problem ONE warp: the node-parallel passes of an interior-point iteration (KKT residuals, barrier /
// slack elimination, expansion of the steps, step length) then walk the 6 box slots and M obstacles of a node sequentially,
// and the lanes beyond N+1 idle.  Here a problem owns a CTA of ceil((N+1)/NPW) warps (5 at N = 20, M = 5) and EVERYTHING of
// the problem lives in shared memory (57 KB: four problems per SM, no scratch in global memory at all):
//   * passes: one thread per (node, chunk), chunk 0 = the node-level algebra + the six box slots, chunk 1..M = one obstacle
//     (its two general constraints and two slack bounds).  The M+1 chunks of a node sit in adjacent lanes; what they
//     contribute to the node's condensed Hessian / gradient / residual is summed by a fixed shuffle tree, problem-wide
//     reductions (mu, step length, norms) go warp butterfly -> shared memory -> every thread, in a fixed order;
//   * sequential Riccati sweeps: warp 0 runs the factorisation and the vector sweeps of mpc_wpp.cuh unchanged (lane = matrix
//     entry / vector component) while the other warps wait at the block barrier.
// Per-problem latency is what this buys (single episodes, small batches: BASELINE config 1 and the p50 single-problem
// latency); instantiated for the reference's dimensions (M = 5; N = 20 and N = 10).  Arithmetic is written with explicit fma()
// like mpc_wpp.cuh; the order of the sums differs from the warp kernel (tree over chunks instead of a sequential loop), so
// the two kernels agree to rounding, not bit for bit -- each has its own bit-exact CPU mirror (oracle/simt/).
#pragma once
#include "mpc_wpp.cuh"

template <int MT, int NT>
struct CtaSolver {
    static constexpr unsigned long long SMASK = B2_WPP_ALL_SHARED;
    static constexpr int M = MT, N = NT, NS = (NT + 1) | 1, nct = 12 + 4 * MT;
    static constexpr int CH = MT + 1;                       // chunks per node
    static constexpr int NPW = 32 / CH;                     // nodes per warp in the passes
    static constexpr int NWARPS = (NT + 1 + NPW - 1) / NPW;
    static constexpr int NTHR = 32 * NWARPS;
    static_assert(MT >= 1 && MT <= 7 && NWARPS <= 8, "CTA kernel: M <= 7, (N+1) <= 8 * (32 / (M+1))");
    static constexpr int o_lb = 0, o_lg = 6, o_ub = 6 + MT, o_ug = 12 + MT, o_ls = 12 + 2 * MT, o_us = 12 + 3 * MT;

    const B2Params& p;
    const int b, tid, lane, warp;
    double* const sm;
    const WOff o;
    const size_t Bm;
    const int c;            // chunk of this thread: 0 = node level + box slots, 1..M = obstacle c-1
    const int lead;         // lane of chunk 0 of the same node
    const bool act;         // the thread owns a (node, chunk) pair
    const int k;            // its node (clamped into range for idle threads)
    int nc;
    double mu, nrm[4], sum_lt, alpha, S1, S2;
    double rm_cdd, rm_smu;
    double ad_n, ad_d, ap_n, ap_d;

    __device__ CtaSolver(const B2Params& p_, int b_, double* sm_)
        : p(p_), b(b_), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), sm(sm_), o(MT, SMASK, false),
          Bm((size_t)p_.Bm), c((threadIdx.x & 31) % CH), lead((threadIdx.x & 31) - (threadIdx.x & 31) % CH),
          act((threadIdx.x & 31) < NPW * CH && (int)(threadIdx.x >> 5) * NPW + (int)(threadIdx.x & 31) / CH <= NT),
          k(((int)(threadIdx.x >> 5) * NPW + (int)(threadIdx.x & 31) / CH) <= NT ? ((int)(threadIdx.x >> 5) * NPW + (int)(threadIdx.x & 31) / CH) : NT)
    {
        nc = 0;
        for (int kk = 0; kk <= N; kk++) nc += 2 * ((kk < N ? 2 : 0) + ((kk > 0 && kk < N) ? 4 : 0)) + 4 * M;
    }

    __device__ __forceinline__ double& S(int row, int kk) const { return sm[row * NS + kk]; }
    template <int F>
    __device__ __forceinline__ double& Q(int i) const { return sm[(o.pf[F] + i) * NS + k]; }
    template <int F>
    __device__ __forceinline__ double& R(int i, int kk) const { return sm[(o.pf[F] + i) * NS + kk]; }
    __device__ __forceinline__ double& VV(int i, int kk) const { return sm[(o.pf[PF_V] + i) * NS + kk]; }
    __device__ __forceinline__ double& PP(int i, int kk) const { return sm[(o.pf[PF_PI] + i) * NS + kk]; }
    __device__ __forceinline__ bool box_on(int s, int kk) const { return s < 2 ? (kk < N) : (kk > 0 && kk < N); }
    __device__ __forceinline__ double scal(int kk) const { return kk < N ? p.dt : 1.0; }
    __device__ __forceinline__ void box_bounds(int s, double& lb, double& ub) const
    {
        const double xv = Q<PF_XU>(s);
        lb = p.lb6[s] - xv; ub = p.ub6[s] - xv;
    }
    __device__ __forceinline__ bool is_lead() const { return act && c == 0; }
    __device__ __forceinline__ bool is_obst() const { return act && c > 0; }

    // ---- reductions ------------------------------------------------------------------------------------------------
    // sum over the CH adjacent lanes of a node, result in its chunk-0 lane: a fixed tree over power-of-two offsets
    // (CH = 6: (v0 + v4 + v2) + (v1 + v5 + v3)); every lane of the warp takes part in the shuffles
    __device__ __forceinline__ double seg_sum(double v) const
    {
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
            if (off < CH) {
                const double t = __shfl_down_sync(0xffffffffu, v, off);
                if (c < off && c + off < CH) v += t;
            }
        }
        return v;
    }
    __device__ __forceinline__ double from_lead(double v) const { return __shfl_sync(0xffffffffu, v, lead); }

    // problem-wide reduction of up to 8 values per thread: sums (first nsum entries) by warp butterflies, maxima of
    // non-negative values (the rest) by REDUX; the per-warp results meet in shared memory (the AL scratch of the
    // factorisation, idle during the passes) and every thread combines them in warp order
    // (NNEG: maxima of NEGATIVE values, the step-length ratios in HPIPM's negative-reciprocal convention)
    template <int NSUM, int NMAX, int NNEG = 0>
    __device__ __forceinline__ void cta_reduce(double* v)
    {
        constexpr int NV = NSUM + NMAX + NNEG;
        double* const red = sm + o.total * NS;              // B2_WPP_SCRATCH = 40 doubles >= NWARPS * NV
        static_assert(NWARPS * NV <= B2_WPP_SCRATCH, "reduction scratch");
#pragma unroll
        for (int i = 0; i < NSUM; i++) v[i] = b2_warp_sum(v[i]);
#pragma unroll
        for (int i = NSUM; i < NSUM + NMAX; i++) v[i] = b2_warp_max_nonneg(v[i]);
#pragma unroll
        for (int i = NSUM + NMAX; i < NV; i++) v[i] = b2_warp_max_neg(v[i]);
        __syncthreads();                                     // previous readers of the scratch are done
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; i++) red[warp * NV + i] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double r = red[i];
            for (int w = 1; w < NWARPS; w++) {
                const double x = red[w * NV + i];
                r = i < NSUM ? r + x : (i < NSUM + NMAX ? ((x > r || x != x) ? x : r) : (x > r ? x : r));
            }
            v[i] = r;
        }
    }

    // ---- preparation + cold-start initialisation -------------------------------------------------------------------
    __device__ void prepare_and_init()
    {
        // rows that start at zero: lam, t, 1/t, dlam of every constraint slot, the step rows
        for (int e = tid; e < nct * NS; e += NTHR) {
            sm[o.pf[PF_LAM] * NS + e] = 0.0; sm[o.pf[PF_T] * NS + e] = 0.0; sm[o.pf[PF_TINV] * NS + e] = 0.0; sm[o.pf[PF_DD] * NS + e] = 0.0;
        }
        for (int e = tid; e < 7 * NS; e += NTHR) sm[o.pf[PF_DV] * NS + e] = 0.0;
        for (int e = tid; e < 5 * NS; e += NTHR) sm[o.pf[PF_DPI] * NS + e] = 0.0;
        __syncthreads();
        double xk[5] = {0, 0, 0, 0, 0}, v[7] = {0, 0, 0, 0, 0, 0, 0};
        if (is_lead()) {
            double uk[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 5; i++) xk[i] = p.X[((size_t)k * 5 + i) * Bm + b];
            const double* yr = p.YREF + ((size_t)k * 6) * Bm + b;
            double sv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (k < N) {
                uk[0] = p.U[((size_t)k * 2 + 0) * Bm + b]; uk[1] = p.U[((size_t)k * 2 + 1) * Bm + b];
                double xn[5];
                if (p.integrator == 1) model_rk4_step(xk, uk, p.dt, xn, sv);
                else model_gl4_step(xk, uk, p.dt, xn, sv);
#pragma unroll
                for (int i = 0; i < 5; i++) Q<PF_B>(i) = xn[i] - p.X[((size_t)(k + 1) * 5 + i) * Bm + b];
                Q<PF_G>(0) = p.Wg[0] * (uk[0] - yr[4 * Bm]);
                Q<PF_G>(1) = p.Wg[1] * (uk[1] - yr[5 * Bm]);
                Q<PF_G>(2) = p.Wg[2] * (xk[0] - yr[0 * Bm]);
                Q<PF_G>(3) = p.Wg[3] * (xk[1] - yr[1 * Bm]);
                Q<PF_G>(4) = 0.0;
                Q<PF_G>(5) = p.Wg[5] * (xk[3] - yr[2 * Bm]);
                Q<PF_G>(6) = p.Wg[6] * (xk[4] - yr[3 * Bm]);
            } else {
                Q<PF_G>(0) = 0.0; Q<PF_G>(1) = 0.0;
                Q<PF_G>(2) = p.Wge[2] * (xk[0] - yr[0 * Bm]);
                Q<PF_G>(3) = p.Wge[3] * (xk[1] - yr[1 * Bm]);
                Q<PF_G>(4) = 0.0;
                Q<PF_G>(5) = p.Wge[5] * (xk[3] - yr[2 * Bm]);
                Q<PF_G>(6) = p.Wge[6] * (xk[4] - yr[3 * Bm]);
            }
#pragma unroll
            for (int i = 0; i < 10; i++) S(o.sv + i, k) = sv[i];
            S(o.sv + 10, k) = 1.0; S(o.sv + 11, k) = 0.0; S(o.sv + 12, k) = p.dt; S(o.sv + 13, k) = p.dt2h;
            Q<PF_XU>(0) = uk[0]; Q<PF_XU>(1) = uk[1];
            Q<PF_XU>(2) = xk[0]; Q<PF_XU>(3) = xk[1]; Q<PF_XU>(4) = xk[3]; Q<PF_XU>(5) = xk[4];
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) v[2 + i] = p.X0[(size_t)i * Bm + b] - xk[i];
            }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                double lb, ub;
                box_bounds(s, lb, ub);
                const int iv = b2_slot2v(s);
                double tl = -lb + v[iv], tu = ub - v[iv];
                if (tl < p.thr0) {
                    if (tu < p.thr0) { v[iv] = 0.5 * (lb + ub); tl = p.thr0; tu = p.thr0; }
                    else { tl = p.thr0; v[iv] = lb + p.thr0; }
                } else if (tu < p.thr0) { tu = p.thr0; v[iv] = ub - p.thr0; }
                Q<PF_T>(o_lb + s) = tl; Q<PF_T>(o_ub + s) = tu;
                Q<PF_LAM>(o_lb + s) = p.mu0 / tl; Q<PF_LAM>(o_ub + s) = p.mu0 / tu;
            }
#pragma unroll
            for (int i = 0; i < 7; i++) VV(i, k) = v[i];
#pragma unroll
            for (int i = 0; i < 5; i++) PP(i, k) = 0.0;
        }
        // the obstacle chunks need the linearisation point and the (possibly shifted) initial v of their node
        xk[0] = from_lead(xk[0]); xk[1] = from_lead(xk[1]);
        const double v2 = from_lead(v[2]), v3 = from_lead(v[3]);
        if (is_obst()) {
            const int j = c - 1;
            const double sc = scal(k);
            double hval, cx, cy;
            model_con_h(xk, p.P[((size_t)k * 2 * M + 2 * j) * Bm + b], p.P[((size_t)k * 2 * M + 2 * j + 1) * Bm + b], p.r2, &hval, &cx, &cy);
            const double s_l = p.SL[((size_t)k * M + j) * Bm + b], s_u = p.SU[((size_t)k * M + j) * Bm + b];
            const double lg = p.lh - hval - s_l, ug = p.uh - hval + s_u;
            Q<PF_CX>(j) = cx; Q<PF_CY>(j) = cy; Q<PF_LG>(j) = lg; Q<PF_UG>(j) = ug;
            Q<PF_SLI>(j) = s_l; Q<PF_SUI>(j) = s_u;
            const double zl_in = p.ZL[((size_t)k * M + j) * Bm + b], Zl_in = p.ZZL[((size_t)k * M + j) * Bm + b];
            Q<PF_ZLQ>(j) = sc * (zl_in + Zl_in * s_l);
            Q<PF_ZZLQ>(j) = sc * Zl_in;
            const double cv = cx * v2 + cy * v3;
            double tl = cv - lg, tu = ug - cv;
            tl = p.thr0 > tl ? p.thr0 : tl;
            tu = p.thr0 > tu ? p.thr0 : tu;
            Q<PF_T>(o_lg + j) = tl; Q<PF_T>(o_ug + j) = tu;
            Q<PF_LAM>(o_lg + j) = p.mu0 / tl; Q<PF_LAM>(o_ug + j) = p.mu0 / tu;
            Q<PF_T>(o_ls + j) = 1.0; Q<PF_T>(o_us + j) = 1.0;
            Q<PF_LAM>(o_ls + j) = p.mu0; Q<PF_LAM>(o_us + j) = p.mu0;
            Q<PF_SL>(j) = 0.0; Q<PF_SU>(j) = 0.0; Q<PF_DSL>(j) = 0.0; Q<PF_DSU>(j) = 0.0;
        }
        __syncthreads();
    }

    __device__ __forceinline__ static double slot_get(const double* a, int s) { return a[b2_slot2v(s)]; }
    __device__ __forceinline__ static void slot_add(double* a, int s, double x) { a[b2_slot2v(s)] += x; }

    __device__ __forceinline__ double rm_of(double l, double t, double dd) const { return fma(l, t, rm_cdd * dd) - rm_smu; }

    // ---- head of an interior-point iteration (see mpc_wpp.cuh head()): apply the previous step, KKT residuals and mu,
    // barrier / slack elimination, condensed affine right-hand side.  Chunk 0: node algebra + box slots; chunk j+1: obstacle j
    __device__ void head(double a)
    {
        for (int e = tid; e < 7 * (N + 1); e += NTHR) {
            const int i = e / (N + 1), kk = e - i * (N + 1);
            const double vo = VV(i, kk);
            S(o.cv + i, kk) = vo;                              // the old iterate, for the recomputation of dt below
            VV(i, kk) = fma(a, R<PF_DV>(i, kk), vo);
        }
        for (int e = tid; e < 5 * N; e += NTHR) {
            const int i = e / N, kk = e - i * N;
            PP(i, kk) = fma(a, R<PF_DPI>(i, kk), PP(i, kk));
        }
        __syncthreads();
        double sum = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
        double rg[7] = {0, 0, 0, 0, 0, 0, 0}, Hx[7] = {0, 0, 0, 0, 0, 0, 0}, gt[7] = {0, 0, 0, 0, 0, 0, 0}, hxy = 0.0;
        double v[7] = {0, 0, 0, 0, 0, 0, 0}, sv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        if (is_lead()) {
            const double* Hd = k < N ? p.Hd : p.Hde;
#pragma unroll
            for (int i = 0; i < 7; i++) { v[i] = VV(i, k); rg[i] = fma(Hd[i], v[i], Q<PF_G>(i)); Hx[i] = Hd[i] + p.reg_prim; }
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 10; i++) sv[i] = S(o.sv + i, k);
                double pk[5], bt[7];
#pragma unroll
                for (int i = 0; i < 5; i++) pk[i] = PP(i, k);
                b2_bat_mul(sv, p.dt, p.dt2h, pk, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) rg[i] += bt[i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) rg[2 + i] -= PP(i, k - 1);
            }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const int il = o_lb + s, iu = o_ub + s;
                double ll = fma(a, Q<PF_DD>(il), Q<PF_LAM>(il)), lu_ = fma(a, Q<PF_DD>(iu), Q<PF_LAM>(iu));
                double lb, ub;
                box_bounds(s, lb, ub);
                const double tlo = Q<PF_T>(il), tuo = Q<PF_T>(iu);
                const double vos = S(o.cv + b2_slot2v(s), k), dvs = R<PF_DV>(b2_slot2v(s), k);
                const double dtl = dvs + ((vos - lb) - tlo), dtu = -dvs + ((ub - vos) - tuo);      // = expand()'s dtl, dtu
                double tl = fma(a, dtl, tlo), tu = fma(a, dtu, tuo);
                if (ll <= p.lam_min) ll = p.lam_min;
                if (lu_ <= p.lam_min) lu_ = p.lam_min;
                if (tl <= p.t_min) tl = p.t_min;
                if (tu <= p.t_min) tu = p.t_min;
                Q<PF_LAM>(il) = ll; Q<PF_LAM>(iu) = lu_; Q<PF_T>(il) = tl; Q<PF_T>(iu) = tu;
                slot_add(rg, s, -ll + lu_);
                const double vv = slot_get(v, s);
                const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
                n2 = fmax(n2, fmax(fabs(rdl), fabs(rdu)));
                const double ml = ll * tl, mu_ = lu_ * tu;
                n3 = fmax(n3, fmax(fabs(ml), fabs(mu_)));
                sum += ml; sum += mu_;
                const double rlu = 1.0 / (tl * tu);
                const double til = b2_rcp_polish(tl, rlu * tu), tiu = b2_rcp_polish(tu, rlu * tl);
                Q<PF_TINV>(il) = til; Q<PF_TINV>(iu) = tiu;
                slot_add(Hx, s, fma(lu_, tiu, ll * til));
                slot_add(gt, s, fma(-fma(lu_, rdu, mu_), tiu, fma(ll, rdl, ml) * til));
            }
        }
        // what the obstacle chunks need of their node
        const double v2 = VV(2, k), v3 = VV(3, k);
        double o_rg2 = 0.0, o_rg3 = 0.0, o_h2 = 0.0, o_h3 = 0.0, o_hxy = 0.0, o_gt2 = 0.0, o_gt3 = 0.0;
        if (is_obst()) {
            const int j = c - 1;
            const double vo2 = S(o.cv + 2, k), vo3 = S(o.cv + 3, k), dv2 = R<PF_DV>(2, k), dv3 = R<PF_DV>(3, k);
            const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
            double l0 = fma(a, Q<PF_DD>(i0), Q<PF_LAM>(i0)), l1 = fma(a, Q<PF_DD>(i1), Q<PF_LAM>(i1));
            double l2 = fma(a, Q<PF_DD>(i2), Q<PF_LAM>(i2)), l3 = fma(a, Q<PF_DD>(i3), Q<PF_LAM>(i3));
            const double cx = Q<PF_CX>(j), cy = Q<PF_CY>(j);
            const double cdv = fma(cy, dv3, cx * dv2);
            const double cvalo = fma(cy, vo3, cx * vo2);
            const double t0o = Q<PF_T>(i0), t1o = Q<PF_T>(i1), t2o = Q<PF_T>(i2), t3o = Q<PF_T>(i3);
            const double slo = Q<PF_SL>(j), suo = Q<PF_SU>(j), dslo = Q<PF_DSL>(j), dsuo = Q<PF_DSU>(j);
            const double lgj = Q<PF_LG>(j), ugj = Q<PF_UG>(j), sli = Q<PF_SLI>(j), sui = Q<PF_SUI>(j);
            const double rd0o = cvalo + slo - lgj - t0o;
            const double rd1o = ugj - cvalo + suo - t1o;
            const double rd2o = slo + sli - t2o;
            const double rd3o = suo + sui - t3o;
            const double dt0 = cdv + dslo + rd0o, dt1 = -cdv + dsuo + rd1o, dt2 = dslo + rd2o, dt3 = dsuo + rd3o;
            double t0 = fma(a, dt0, t0o), t1 = fma(a, dt1, t1o), t2 = fma(a, dt2, t2o), t3 = fma(a, dt3, t3o);
            if (l0 <= p.lam_min) l0 = p.lam_min;
            if (l1 <= p.lam_min) l1 = p.lam_min;
            if (l2 <= p.lam_min) l2 = p.lam_min;
            if (l3 <= p.lam_min) l3 = p.lam_min;
            if (t0 <= p.t_min) t0 = p.t_min;
            if (t1 <= p.t_min) t1 = p.t_min;
            if (t2 <= p.t_min) t2 = p.t_min;
            if (t3 <= p.t_min) t3 = p.t_min;
            Q<PF_LAM>(i0) = l0; Q<PF_LAM>(i1) = l1; Q<PF_LAM>(i2) = l2; Q<PF_LAM>(i3) = l3;
            Q<PF_T>(i0) = t0; Q<PF_T>(i1) = t1; Q<PF_T>(i2) = t2; Q<PF_T>(i3) = t3;
            const double s_l = fma(a, dslo, slo), s_u = fma(a, dsuo, suo);
            Q<PF_SL>(j) = s_l; Q<PF_SU>(j) = s_u;
            const double cv = fma(cy, v3, cx * v2);
            const double dl = l1 - l0;
            o_rg2 = cx * dl; o_rg3 = cy * dl;
            const double Zlq = Q<PF_ZZLQ>(j);
            const double rgsl = fma(Zlq, s_l, Q<PF_ZLQ>(j)) - l0 - l2;
            const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, sui, p.zu)) - l1 - l3;
            n0 = fmax(fabs(rgsl), fabs(rgsu));
            const double rd0 = cv + s_l - lgj - t0;
            const double rd1 = ugj - cv + s_u - t1;
            const double rd2 = s_l + sli - t2;
            const double rd3 = s_u + sui - t3;
            n2 = fmax(fmax(fabs(rd0), fabs(rd1)), fmax(fabs(rd2), fabs(rd3)));
            const double m0 = l0 * t0, m1 = l1 * t1, m2 = l2 * t2, m3 = l3 * t3;
            n3 = fmax(fmax(fabs(m0), fabs(m1)), fmax(fabs(m2), fabs(m3)));
            sum = (m0 + m1) + (m2 + m3);
            const double p01 = t0 * t1, p23 = t2 * t3;
            const double r4 = 1.0 / (p01 * p23);
            const double r01 = r4 * p23, r23 = r4 * p01;
            const double ti0 = b2_rcp_polish(t0, r01 * t1), ti1 = b2_rcp_polish(t1, r01 * t0);
            const double ti2 = b2_rcp_polish(t2, r23 * t3), ti3 = b2_rcp_polish(t3, r23 * t2);
            Q<PF_TINV>(i0) = ti0; Q<PF_TINV>(i1) = ti1; Q<PF_TINV>(i2) = ti2; Q<PF_TINV>(i3) = ti3;
            const double Gl = l0 * ti0, Gu = l1 * ti1, Gls = l2 * ti2, Gus = l3 * ti3;
            const double dl_ = Zlq + Gl + Gls, du_ = Zuq + Gu + Gus;
            const double rd_ = 1.0 / (dl_ * du_);
            const double Dli = b2_rcp_polish(dl_, rd_ * du_), Dui = b2_rcp_polish(du_, rd_ * dl_);
            Q<PF_DLI>(j) = Dli; Q<PF_DUI>(j) = Dui;
            const double weff = fma(Gu * (Zuq + Gus), Dui, Gl * (Zlq + Gls) * Dli);
            const double wcx = weff * cx;
            o_h2 = wcx * cx; o_h3 = (weff * cy) * cy; o_hxy = wcx * cy;
            const double q0 = fma(l0, rd0, m0) * ti0, q1 = fma(l1, rd1, m1) * ti1;
            const double q2 = fma(l2, rd2, m2) * ti2, q3 = fma(l3, rd3, m3) * ti3;
            const double rhol = rgsl + q0 + q2, rhou = rgsu + q1 + q3;
            const double coef = fma(-(Gl * rhol), Dli, q0) - fma(-(Gu * rhou), Dui, q1);
            o_gt2 = cx * coef; o_gt3 = cy * coef;
        }
        o_rg2 = seg_sum(o_rg2); o_rg3 = seg_sum(o_rg3); o_h2 = seg_sum(o_h2); o_h3 = seg_sum(o_h3);
        o_hxy = seg_sum(o_hxy); o_gt2 = seg_sum(o_gt2); o_gt3 = seg_sum(o_gt3);
        if (is_lead()) {
            rg[2] += o_rg2; rg[3] += o_rg3; Hx[2] += o_h2; Hx[3] += o_h3; hxy = o_hxy; gt[2] += o_gt2; gt[3] += o_gt3;
            if (k < N) {
                double rb[5];
                b2_ba_mul(sv, p.dt, p.dt2h, v, rb);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    rb[i] += Q<PF_B>(i) - VV(2 + i, k + 1);
                    S(o.rb + i, k) = rb[i];
                    n1 = fmax(n1, fabs(rb[i]));
                }
            }
            if (k == 0) { rg[2] = rg[3] = rg[4] = rg[5] = rg[6] = 0.0; }
            if (k == N) { rg[0] = rg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) {
                Q<PF_RG>(i) = rg[i]; n0 = fmax(n0, fabs(rg[i]));
                S(o.hg + i, k) = rg[i] + gt[i];
            }
        }
        double red[5] = {sum, n0, n1, n2, n3};
        cta_reduce<1, 4, 0>(red);                            // (its barriers also order the reads of cv above ...
        if (is_lead()) {
#pragma unroll
            for (int i = 0; i < 7; i++) S(o.cv + i, k) = Hx[i];   // ... against these stores of the node Hessian)
            S(o.hg + 7, k) = hxy;
        }
        sum_lt = red[0]; nrm[0] = red[1]; nrm[1] = red[2]; nrm[2] = red[3]; nrm[3] = red[4];
        mu = sum_lt / nc;
        __syncthreads();
    }

    // ---- backward Riccati recursion in square-root form: warp 0, as mpc_wpp.cuh riccati_factorize() ---------------------
    __device__ void riccati_factorize()
    {
        const int e = lane < 28 ? lane : 0;
        int mi = 0;
        while ((mi + 1) * (mi + 2) / 2 <= e) mi++;
        const int mj = e - mi * (mi + 1) / 2;
        const int ai0 = lane / 5, aq0 = lane % 5, ai1 = (32 + lane) / 5, aq1 = (32 + lane) % 5;
        int ba0[5], ba1[5];
#pragma unroll
        for (int r = 0; r < 5; r++) { ba0[r] = o.sv + B2_BATAB[r][ai0]; ba1[r] = o.sv + B2_BATAB[r][ai1 < 7 ? ai1 : 0]; }
        const bool xx = lane < 28 && mi >= 2 && mj >= 2;
        const bool xu = lane < 28 && mi >= 2 && mj < 2;
        const int xrow = xx ? o.lxx + b2_sym(mi - 2, mj - 2) : 0;
        const int lrow = xu ? o.lxu + 2 * (mi - 2) + mj : 0;
        double* const scr = sm + o.total * NS;
        for (int kk = N; kk >= 0; kk--) {
            double m = 0.0;
            if (mi == mj) m = S(o.cv + mi, kk);
            if (mi == 3 && mj == 2) m = S(o.hg + 7, kk);
            if (kk < N) {
                {
                    double acc = 0.0;
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        const double l = (r >= aq0) ? S(o.lxx + b2_sym(r, aq0 <= r ? aq0 : r), kk + 1) : 0.0;
                        acc = fma(S(ba0[r], kk), l, acc);
                    }
                    scr[lane] = acc;
                    if (lane < 3) {
                        double acc1 = 0.0;
#pragma unroll
                        for (int r = 0; r < 5; r++) {
                            const double l = (r >= aq1) ? S(o.lxx + b2_sym(r, aq1 <= r ? aq1 : r), kk + 1) : 0.0;
                            acc1 = fma(S(ba1[r], kk), l, acc1);
                        }
                        scr[32 + lane] = acc1;
                    }
                }
                __syncwarp();
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 5; q++) acc = fma(scr[mi * 5 + q], scr[mj * 5 + q], acc);
                m += acc;
            }
            double sigma = 1.0, my_sp = 0.0, sp0 = 0.0;
#pragma unroll 1
            for (int cc = 0; cc < 7; cc++) {
                const double pv_ = b2_bcast(m, cc * (cc + 1) / 2 + cc);
                const double la = b2_bcast(m, mi * (mi + 1) / 2 + (cc <= mi ? cc : mi));
                const double lb = b2_bcast(m, mj * (mj + 1) / 2 + (cc <= mj ? cc : mj));
                const bool ok = pv_ > 0.0;
                const double sp = ok ? sigma * pv_ : 0.0;
                if (cc == 0) sp0 = sp;
                if (mj == cc) my_sp = sp;
                if (ok) {
                    const int ex = ((__double2hiint(pv_) >> 20) & 0x7ff) - 1023;
                    const double scale = __hiloint2double((1023 - ex) << 20, 0);
                    const double pn = pv_ * scale;
                    if (mj > cc) m = fma(m, pn, -((la * lb) * scale));
                    sigma *= pn;
                }
            }
            const double myrs = my_sp > 0.0 ? 1.0 / sqrt(my_sp) : 0.0;
            m *= myrs;
            const double rs0 = b2_bcast(myrs, 0);
            const double sig1 = sp0 > 0.0 ? sp0 * __hiloint2double((1023 - (((__double2hiint(sp0) >> 20) & 0x7ff) - 1023)) << 20, 0) : 1.0;
            const double rs1 = sig1 * b2_bcast(myrs, 2);
            if (kk < N && lane == 0) { S(o.lamu + 0, kk) = rs0; S(o.lamu + 2, kk) = rs1; }
            if (kk < N && lane == 1) S(o.lamu + 1, kk) = m;
            if (kk > 0) {
                if (xx) S(xrow, kk) = m;
                if (xu) S(lrow, kk) = m;
            }
            __syncwarp();
        }
    }

    // q_lg, q_ug and rho_l, rho_u of obstacle j at node k (base kinds)
    __device__ __forceinline__ void slack_rho(int j, double v2, double v3, double Zuq, double sc,
                                              double& rhol, double& rhou, double& q0, double& q1) const
    {
        const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
        const double l0 = Q<PF_LAM>(i0), l1 = Q<PF_LAM>(i1), l2 = Q<PF_LAM>(i2), l3 = Q<PF_LAM>(i3);
        const double t0 = Q<PF_T>(i0), t1 = Q<PF_T>(i1), t2 = Q<PF_T>(i2), t3 = Q<PF_T>(i3);
        const double s_l = Q<PF_SL>(j), s_u = Q<PF_SU>(j);
        const double cv = fma(Q<PF_CY>(j), v3, Q<PF_CX>(j) * v2);
        const double rd0 = cv + s_l - Q<PF_LG>(j) - t0;
        const double rd1 = Q<PF_UG>(j) - cv + s_u - t1;
        const double rd2 = s_l + Q<PF_SLI>(j) - t2;
        const double rd3 = s_u + Q<PF_SUI>(j) - t3;
        q0 = fma(l0, rd0, rm_of(l0, t0, Q<PF_DD>(i0))) * Q<PF_TINV>(i0);
        q1 = fma(l1, rd1, rm_of(l1, t1, Q<PF_DD>(i1))) * Q<PF_TINV>(i1);
        const double q2 = fma(l2, rd2, rm_of(l2, t2, Q<PF_DD>(i2))) * Q<PF_TINV>(i2);
        const double q3 = fma(l3, rd3, rm_of(l3, t3, Q<PF_DD>(i3))) * Q<PF_TINV>(i3);
        const double rgsl = fma(Q<PF_ZZLQ>(j), s_l, Q<PF_ZLQ>(j)) - l0 - l2;
        const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, Q<PF_SUI>(j), p.zu)) - l1 - l3;
        rhol = rgsl + q0 + q2;
        rhou = rgsu + q1 + q3;
    }

    // condensed gradient of the right-hand side `kind` (1 corrector, 2 centering, 3 refinement) -> rows hg[0..6]
    __device__ void solve_prep(int kind)
    {
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        double gt[7] = {0, 0, 0, 0, 0, 0, 0};
        double o2 = 0.0, o3 = 0.0;
        if (kind == 3) {
            if (is_obst()) {
                const int j = c - 1;
                const double G0 = Q<PF_LAM>(o_lg + j) * Q<PF_TINV>(o_lg + j), G1 = Q<PF_LAM>(o_ug + j) * Q<PF_TINV>(o_ug + j);
                const double coef = fma(G1 * Q<PF_LRGSU>(j), Q<PF_DUI>(j), -((G0 * Q<PF_LRGSL>(j)) * Q<PF_DLI>(j)));
                o2 = Q<PF_CX>(j) * coef; o3 = Q<PF_CY>(j) * coef;
            }
            o2 = seg_sum(o2); o3 = seg_sum(o3);
            if (is_lead()) { S(o.hg + 2, k) += o2; S(o.hg + 3, k) += o3; }       // lin_residuals left lrg in hg
        } else {
            if (is_lead()) {
                double v[7];
#pragma unroll
                for (int i = 0; i < 7; i++) { v[i] = VV(i, k); gt[i] = Q<PF_RG>(i); }
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    if (!box_on(s, k)) continue;
                    const int il = o_lb + s, iu = o_ub + s;
                    const double ll = Q<PF_LAM>(il), lu_ = Q<PF_LAM>(iu), tl = Q<PF_T>(il), tu = Q<PF_T>(iu);
                    double lb, ub;
                    box_bounds(s, lb, ub);
                    const double vv = slot_get(v, s);
                    const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
                    const double ql = fma(ll, rdl, rm_of(ll, tl, Q<PF_DD>(il))) * Q<PF_TINV>(il);
                    const double qu = fma(lu_, rdu, rm_of(lu_, tu, Q<PF_DD>(iu))) * Q<PF_TINV>(iu);
                    slot_add(gt, s, ql - qu);
                }
            }
            if (is_obst()) {
                const int j = c - 1;
                double rhol, rhou, q0, q1;
                slack_rho(j, VV(2, k), VV(3, k), Zuq, sc, rhol, rhou, q0, q1);
                const double G0 = Q<PF_LAM>(o_lg + j) * Q<PF_TINV>(o_lg + j), G1 = Q<PF_LAM>(o_ug + j) * Q<PF_TINV>(o_ug + j);
                const double coef = fma(-(G0 * rhol), Q<PF_DLI>(j), q0) - fma(-(G1 * rhou), Q<PF_DUI>(j), q1);
                o2 = Q<PF_CX>(j) * coef; o3 = Q<PF_CY>(j) * coef;
            }
            o2 = seg_sum(o2); o3 = seg_sum(o3);
            if (is_lead()) {
                gt[2] += o2; gt[3] += o3;
#pragma unroll
                for (int i = 0; i < 7; i++) S(o.hg + i, k) = gt[i];
            }
        }
        __syncthreads();
    }

    __device__ __forceinline__ void ctg_mul(int kk, const double* x, double* y) const
    {
        double L[15], t[5];
#pragma unroll
        for (int i = 0; i < 15; i++) L[i] = S(o.lxx + i, kk);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double acc = 0.0;
#pragma unroll
            for (int r = j; r < 5; r++) acc = fma(L[b2_sym(r, j)], x[r], acc);
            t[j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 5; i++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j <= i; j++) acc = fma(L[b2_sym(i, j)], t[j], acc);
            y[i] = acc;
        }
    }

    // backward + forward vector sweeps with the current factorisation (mpc_wpp.cuh riccati_solve()): the node-parallel
    // products with the cost-to-go factor by the chunk-0 threads, the sequential cores by warp 0
    __device__ void riccati_solve(int kind)
    {
        const bool itr = (kind == 3);
        const int oRB = itr ? o.lrb : o.rb, oDV = itr ? o.cv : o.pf[PF_DV], oDPI = itr ? o.pf[PF_CPI] : o.pf[PF_DPI];
        if (is_lead() && k < N) {
            double rb[5], z[5];
#pragma unroll
            for (int i = 0; i < 5; i++) rb[i] = S(oRB + i, k);
            ctg_mul(k + 1, rb, z);
#pragma unroll
            for (int i = 0; i < 5; i++) S(oDPI + i, k) = z[i];                // z_k (overwritten by dpi_k below)
        }
        __syncthreads();
        if (warp == 0) {
            const int c5 = lane < 5 ? lane : 4;
            int acol[5], arow[7];
#pragma unroll
            for (int r = 0; r < 5; r++) acol[r] = o.sv + B2_BATAB[r][2 + c5];
#pragma unroll
            for (int cc = 0; cc < 7; cc++) arow[cc] = o.sv + B2_BATAB[c5][cc];
            const double h = p.dt, h2 = p.dt2h;
            double pf[5];
            {
                const double pc = S(o.hg + 2 + c5, N);
                if (lane < 5) R<PF_PV>(lane, N) = pc;
#pragma unroll
                for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
            }
            for (int kk = N - 1; kk >= 0; kk--) {
                double s[5];
#pragma unroll
                for (int r = 0; r < 5; r++) s[r] = pf[r] + S(oDPI + r, kk);
                const double w0 = fma(S(o.sv + SV_B00, kk), s[0], S(o.hg + 0, kk)) + fma(S(o.sv + SV_B10, kk), s[1], h * s[3]);
                const double w1 = fma(h, s[4], fma(S(o.sv + SV_B01, kk), s[0], S(o.hg + 1, kk)) + fma(S(o.sv + SV_B11, kk), s[1], h2 * s[2]));
                const double l00i = S(o.lamu + 0, kk), l10 = S(o.lamu + 1, kk), l11i = S(o.lamu + 2, kk);
                const double lu0 = w0 * l00i;
                const double lu1 = fma(-l10, lu0, w1) * l11i;
                if (lane == 0) { S(o.lu + 0, kk) = lu0; S(o.lu + 1, kk) = lu1; }
                if (kk > 0) {
                    const double pa = fma(S(acol[0], kk), s[0], S(o.hg + 2 + c5, kk)) + fma(S(acol[1], kk), s[1], S(acol[2], kk) * s[2]);
                    const double pb = fma(S(acol[3], kk), s[3], S(acol[4], kk) * s[4]);
                    const double pc = (pa + pb) - fma(S(o.lxu + 2 * c5, kk), lu0, S(o.lxu + 2 * c5 + 1, kk) * lu1);
                    if (lane < 5) R<PF_PV>(lane, kk) = pc;
#pragma unroll
                    for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
                }
            }
            __syncwarp();
            double dx[5] = {0, 0, 0, 0, 0};
            if (lane < 5) S(oDV + 2 + lane, 0) = 0.0;
            for (int kk = 0; kk < N; kk++) {
                double r0 = S(o.lu + 0, kk), r1 = S(o.lu + 1, kk);
                if (kk > 0) {
                    r0 = fma(S(o.lxu + 2, kk), dx[1], fma(S(o.lxu + 0, kk), dx[0], r0)) + fma(S(o.lxu + 8, kk), dx[4], fma(S(o.lxu + 6, kk), dx[3], S(o.lxu + 4, kk) * dx[2]));
                    r1 = fma(S(o.lxu + 3, kk), dx[1], fma(S(o.lxu + 1, kk), dx[0], r1)) + fma(S(o.lxu + 9, kk), dx[4], fma(S(o.lxu + 7, kk), dx[3], S(o.lxu + 5, kk) * dx[2]));
                }
                const double l00i = S(o.lamu + 0, kk), l10 = S(o.lamu + 1, kk), l11i = S(o.lamu + 2, kk);
                const double du1 = -r1 * l11i;
                const double du0 = fma(-l10, du1, -r0) * l00i;
                const double xa = fma(S(arow[3], kk), dx[1], fma(S(arow[2], kk), dx[0], S(oRB + c5, kk))) + fma(S(arow[6], kk), dx[4], fma(S(arow[5], kk), dx[3], S(arow[4], kk) * dx[2]));
                const double xn = xa + fma(S(arow[0], kk), du0, S(arow[1], kk) * du1);
                if (lane == 0) { S(oDV + 0, kk) = du0; S(oDV + 1, kk) = du1; }
                if (lane < 5) S(oDV + 2 + lane, kk + 1) = xn;
#pragma unroll
                for (int r = 0; r < 5; r++) dx[r] = b2_bcast(xn, r);
            }
            if (lane == 0) { S(oDV + 0, N) = 0.0; S(oDV + 1, N) = 0.0; }
        }
        __syncthreads();
        if (is_lead() && k < N) {
            double dxn[5], y[5];
#pragma unroll
            for (int i = 0; i < 5; i++) dxn[i] = S(oDV + 2 + i, k + 1);
            ctg_mul(k + 1, dxn, y);
#pragma unroll
            for (int i = 0; i < 5; i++) S(oDPI + i, k) = y[i] + R<PF_PV>(i, k + 1);
        }
        __syncthreads();
    }

    __device__ __forceinline__ void consume(double l, double t, double dti, double dli, double& s1, double& s2)
    {
        if (ad_n * dli < l * ad_d) { ad_n = l; ad_d = dli; }
        if (ap_n * dti < t * ap_d) { ap_n = t; ap_d = dti; }
        s1 += fma(l, dti, t * dli);
        s2 = fma(dli, dti, s2);
    }

    // expansion of the step of this thread's chunk (see mpc_wpp.cuh expand()): MODE XP_SOLVE (base solve `kind`),
    // XP_REFINE (add the refinement correction), XP_LIN (residual of the linear system; lrg2/lrg3: obstacle parts)
    template <int MODE>
    __device__ __forceinline__ void expand(int kind, double& s1, double& s2, double* lrg, double& o2, double& o3, double& ln0)
    {
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        const bool keep_dd = (MODE == XP_SOLVE) && kind == 0 && p.pred_corr;
        if (MODE == XP_REFINE) {
            // dv += cv, dpi += cpi: every chunk reads what it needs of the old rows first
            double dvn[7] = {0, 0, 0, 0, 0, 0, 0};
            if (is_lead()) {
#pragma unroll
                for (int i = 0; i < 7; i++) dvn[i] = R<PF_DV>(i, k) + S(o.cv + i, k);
            }
            __syncwarp();
            if (is_lead()) {
#pragma unroll
                for (int i = 0; i < 7; i++) R<PF_DV>(i, k) = dvn[i];
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < 5; i++) R<PF_DPI>(i, k) += R<PF_CPI>(i, k);
                }
            }
            __syncwarp();
        }
        if (is_lead()) {
            double v[7], dv[7], cv[7];
#pragma unroll
            for (int i = 0; i < 7; i++) { v[i] = VV(i, k); dv[i] = R<PF_DV>(i, k); cv[i] = MODE == XP_REFINE ? S(o.cv + i, k) : 0.0; }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const int il = o_lb + s, iu = o_ub + s;
                const double ll = Q<PF_LAM>(il), lu_ = Q<PF_LAM>(iu), tl = Q<PF_T>(il), tu = Q<PF_T>(iu);
                double lb, ub;
                box_bounds(s, lb, ub);
                const double vv = slot_get(v, s), dvv = slot_get(dv, s);
                const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
                const double dtl = dvv + rdl, dtu = -dvv + rdu;
                double dll, dlu;
                if (MODE == XP_SOLVE) {
                    dll = -fma(ll, dtl, rm_of(ll, tl, Q<PF_DD>(il))) * Q<PF_TINV>(il);
                    dlu = -fma(lu_, dtu, rm_of(lu_, tu, Q<PF_DD>(iu))) * Q<PF_TINV>(iu);
                    Q<PF_DD>(il) = keep_dd ? dtl * dll : dll;
                    Q<PF_DD>(iu) = keep_dd ? dtu * dlu : dlu;
                } else if (MODE == XP_REFINE) {
                    const double cvv = slot_get(cv, s);
                    dll = fma(-(ll * cvv), Q<PF_TINV>(il), Q<PF_DD>(il));
                    dlu = fma(lu_ * cvv, Q<PF_TINV>(iu), Q<PF_DD>(iu));
                    Q<PF_DD>(il) = dll; Q<PF_DD>(iu) = dlu;
                } else {
                    dll = Q<PF_DD>(il); dlu = Q<PF_DD>(iu);
                }
                if (MODE != XP_LIN) { consume(ll, tl, dtl, dll, s1, s2); consume(lu_, tu, dtu, dlu, s1, s2); }
                if (MODE == XP_LIN) slot_add(lrg, s, -dll + dlu);
            }
        }
        if (is_obst()) {
            const int j = c - 1;
            const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
            const double cx = Q<PF_CX>(j), cy = Q<PF_CY>(j);
            const double cdv = fma(cy, R<PF_DV>(3, k), cx * R<PF_DV>(2, k));
            const double l0 = Q<PF_LAM>(i0), l1 = Q<PF_LAM>(i1), l2 = Q<PF_LAM>(i2), l3 = Q<PF_LAM>(i3);
            const double t0 = Q<PF_T>(i0), t1 = Q<PF_T>(i1), t2 = Q<PF_T>(i2), t3 = Q<PF_T>(i3);
            const double s_l = Q<PF_SL>(j), s_u = Q<PF_SU>(j);
            const double v2 = VV(2, k), v3 = VV(3, k);
            const double cval = fma(cy, v3, cx * v2);
            const double rd0 = cval + s_l - Q<PF_LG>(j) - t0;
            const double rd1 = Q<PF_UG>(j) - cval + s_u - t1;
            const double rd2 = s_l + Q<PF_SLI>(j) - t2;
            const double rd3 = s_u + Q<PF_SUI>(j) - t3;
            double dsl, dsu, dl0, dl1, dl2, dl3;
            if (MODE == XP_SOLVE) {
                const double ti0 = Q<PF_TINV>(i0), ti1 = Q<PF_TINV>(i1), ti2 = Q<PF_TINV>(i2), ti3 = Q<PF_TINV>(i3);
                double rhol, rhou, q0, q1;
                slack_rho(j, v2, v3, Zuq, sc, rhol, rhou, q0, q1);
                dsl = -fma(l0 * ti0, cdv, rhol) * Q<PF_DLI>(j);
                dsu = -fma(-(l1 * ti1), cdv, rhou) * Q<PF_DUI>(j);
                if (!keep_dd) { Q<PF_DSL>(j) = dsl; Q<PF_DSU>(j) = dsu; }
                const double dt0 = cdv + dsl + rd0, dt1 = -cdv + dsu + rd1, dt2 = dsl + rd2, dt3 = dsu + rd3;
                dl0 = -fma(l0, dt0, rm_of(l0, t0, Q<PF_DD>(i0))) * ti0;
                dl1 = -fma(l1, dt1, rm_of(l1, t1, Q<PF_DD>(i1))) * ti1;
                dl2 = -fma(l2, dt2, rm_of(l2, t2, Q<PF_DD>(i2))) * ti2;
                dl3 = -fma(l3, dt3, rm_of(l3, t3, Q<PF_DD>(i3))) * ti3;
                Q<PF_DD>(i0) = keep_dd ? dt0 * dl0 : dl0; Q<PF_DD>(i1) = keep_dd ? dt1 * dl1 : dl1;
                Q<PF_DD>(i2) = keep_dd ? dt2 * dl2 : dl2; Q<PF_DD>(i3) = keep_dd ? dt3 * dl3 : dl3;
            } else if (MODE == XP_REFINE) {
                const double ti0 = Q<PF_TINV>(i0), ti1 = Q<PF_TINV>(i1), ti2 = Q<PF_TINV>(i2), ti3 = Q<PF_TINV>(i3);
                const double ccv = fma(cy, S(o.cv + 3, k), cx * S(o.cv + 2, k));
                const double csl = -fma(l0 * ti0, ccv, Q<PF_LRGSL>(j)) * Q<PF_DLI>(j);
                const double csu = -fma(-(l1 * ti1), ccv, Q<PF_LRGSU>(j)) * Q<PF_DUI>(j);
                dsl = Q<PF_DSL>(j) + csl; dsu = Q<PF_DSU>(j) + csu;
                Q<PF_DSL>(j) = dsl; Q<PF_DSU>(j) = dsu;
                dl0 = fma(-(l0 * (ccv + csl)), ti0, Q<PF_DD>(i0));
                dl1 = fma(-(l1 * (-ccv + csu)), ti1, Q<PF_DD>(i1));
                dl2 = fma(-(l2 * csl), ti2, Q<PF_DD>(i2));
                dl3 = fma(-(l3 * csu), ti3, Q<PF_DD>(i3));
                Q<PF_DD>(i0) = dl0; Q<PF_DD>(i1) = dl1; Q<PF_DD>(i2) = dl2; Q<PF_DD>(i3) = dl3;
            } else {
                dsl = Q<PF_DSL>(j); dsu = Q<PF_DSU>(j);
                dl0 = Q<PF_DD>(i0); dl1 = Q<PF_DD>(i1); dl2 = Q<PF_DD>(i2); dl3 = Q<PF_DD>(i3);
            }
            const double dt0 = cdv + dsl + rd0, dt1 = -cdv + dsu + rd1, dt2 = dsl + rd2, dt3 = dsu + rd3;
            if (MODE != XP_LIN) {
                consume(l0, t0, dt0, dl0, s1, s2); consume(l1, t1, dt1, dl1, s1, s2);
                consume(l2, t2, dt2, dl2, s1, s2); consume(l3, t3, dt3, dl3, s1, s2);
            }
            if (MODE == XP_LIN) {
                const double dl = dl1 - dl0;
                o2 = cx * dl; o3 = cy * dl;
                const double rgsl = fma(Q<PF_ZZLQ>(j), s_l, Q<PF_ZLQ>(j)) - l0 - l2;
                const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, Q<PF_SUI>(j), p.zu)) - l1 - l3;
                const double a0 = fma(Q<PF_ZZLQ>(j), dsl, rgsl) - dl0 - dl2;
                const double a1 = fma(Zuq, dsu, rgsu) - dl1 - dl3;
                Q<PF_LRGSL>(j) = a0; Q<PF_LRGSU>(j) = a1;
                ln0 = fmax(fabs(a0), fabs(a1));
            }
        }
    }

    // one linear solve with the current factorisation + step-length reductions (kind 0 affine, 1 corrector, 2 centering,
    // 3 one iterative-refinement step)
    __device__ void solve_any(int kind, double smu)
    {
        rm_cdd = kind == 1 ? 1.0 : 0.0;
        rm_smu = kind == 0 ? 0.0 : smu;
        if (kind != 0) solve_prep(kind);
        riccati_solve(kind);
        double s1 = 0.0, s2 = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0;
        ad_n = 1.0; ad_d = -1.0; ap_n = 1.0; ap_d = -1.0;
        if (kind == 3) expand<XP_REFINE>(3, s1, s2, nullptr, d0, d1, d2);
        else expand<XP_SOLVE>(kind, s1, s2, nullptr, d0, d1, d2);
        // negative-reciprocal convention (HPIPM d_compute_alpha_qp): the largest of the negative ratios limits the step
        double red[4] = {s1, s2, ap_n / ap_d, ad_n / ad_d};
        cta_reduce<2, 0, 2>(red);
        S1 = red[0]; S2 = red[1];
        alpha = -(red[2] > red[3] ? red[2] : red[3]);
        __syncthreads();
    }

    // residual of the linear system at the stored step; leaves lrg in rows hg[0..6] and lrb in rows lrb[0..4]
    __device__ void lin_residuals(double lnrm[2])
    {
        double ln0 = 0.0, ln1 = 0.0, o2 = 0.0, o3 = 0.0, s1 = 0.0, s2 = 0.0;
        double lrg[7] = {0, 0, 0, 0, 0, 0, 0}, dv[7] = {0, 0, 0, 0, 0, 0, 0}, sv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (is_lead()) {
            const double* Hd = k < N ? p.Hd : p.Hde;
#pragma unroll
            for (int i = 0; i < 7; i++) { dv[i] = R<PF_DV>(i, k); lrg[i] = fma(Hd[i], dv[i], Q<PF_RG>(i)); }
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 10; i++) sv[i] = S(o.sv + i, k);
                double dp[5], bt[7];
#pragma unroll
                for (int i = 0; i < 5; i++) dp[i] = R<PF_DPI>(i, k);
                b2_bat_mul(sv, p.dt, p.dt2h, dp, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) lrg[i] += bt[i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) lrg[2 + i] -= R<PF_DPI>(i, k - 1);
            }
        }
        expand<XP_LIN>(0, s1, s2, lrg, o2, o3, ln0);
        o2 = seg_sum(o2); o3 = seg_sum(o3);
        if (is_lead()) {
            lrg[2] += o2; lrg[3] += o3;
            if (k < N) {
                double lrb[5];
                b2_ba_mul(sv, p.dt, p.dt2h, dv, lrb);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    lrb[i] += S(o.rb + i, k) - R<PF_DV>(2 + i, k + 1);
                    S(o.lrb + i, k) = lrb[i];
                    ln1 = fmax(ln1, fabs(lrb[i]));
                }
            }
            if (k == 0) { lrg[2] = lrg[3] = lrg[4] = lrg[5] = lrg[6] = 0.0; }
            if (k == N) { lrg[0] = lrg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) { S(o.hg + i, k) = lrg[i]; ln0 = fmax(ln0, fabs(lrg[i])); }
        }
        double red[2] = {ln0, ln1};
        cta_reduce<0, 2, 0>(red);
        lnrm[0] = red[0]; lnrm[1] = red[1];
        __syncthreads();
    }

    // ---- one RTI iteration (robot_ocp_problem.py:195) ---------------------------------------------------------------
    __device__ void run()
    {
        int status = 0, qs = 0, n_itref = 0;
        prepare_and_init();
        alpha = 1.0;
        int kk = 0;
        const double inv_nc = 1.0 / nc;
        double a = 0.0;
        for (;; kk++) {
            head(a);
            if (!(kk < p.qp_iter_max && alpha > p.alpha_min &&
                  (nrm[0] > p.res_g_max || nrm[1] > p.res_b_max || nrm[2] > p.res_d_max || nrm[3] > p.res_m_max))) break;
            if (warp == 0) riccati_factorize();
            __syncthreads();
            int kind = 0, itref = 0;
            double smu = 0.0, mu_aff0 = 0.0;
            for (;;) {
                solve_any(kind, smu);
                if (!p.pred_corr) break;
                if (kind == 0) {
                    const double mu_aff = (sum_lt + alpha * S1 + alpha * alpha * S2) * inv_nc;
                    const double tmp = mu_aff / mu;
                    smu = tmp * tmp * tmp * mu;
                    smu = smu > p.tau_min ? smu : p.tau_min;
                    mu_aff0 = mu_aff;
                    kind = 1;
                    continue;
                }
                if (kind == 1 && p.cond_pred_corr) {
                    const double mu_aff = (sum_lt + alpha * S1 + alpha * alpha * S2) * inv_nc;
                    if (mu_aff > 2.0 * mu_aff0) { kind = 2; continue; }
                }
                if (itref >= p.itref_corr_max) break;
                double lnrm[2];
                lin_residuals(lnrm);
                if (lnrm[0] < p.res_g_max && lnrm[1] < p.res_b_max) break;
                kind = 3; itref++; n_itref++;
            }
            a = alpha;
            if (a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
        }
        if (kk == p.qp_iter_max) qs = 1;
        else if (alpha <= p.alpha_min) qs = 2;
        else if (mu != mu || nrm[0] != nrm[0] || nrm[1] != nrm[1] || nrm[2] != nrm[2] || nrm[3] != nrm[3]) qs = 3;
        else qs = 0;
        status = (qs != 0 && qs != 1) ? 4 : 0;     // ACADOS_QP_FAILURE leaves the iterate and the multipliers untouched
        if (status == 0) {
            if (is_lead()) {
#pragma unroll
                for (int i = 0; i < 5; i++) p.X[((size_t)k * 5 + i) * Bm + b] += VV(2 + i, k);
                if (k < N) {
                    p.U[((size_t)k * 2 + 0) * Bm + b] += VV(0, k);
                    p.U[((size_t)k * 2 + 1) * Bm + b] += VV(1, k);
#pragma unroll
                    for (int i = 0; i < 5; i++) p.PI[((size_t)k * 5 + i) * Bm + b] = PP(i, k);
                }
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    p.LAM[((size_t)k * nct + o_lb + s) * Bm + b] = Q<PF_LAM>(o_lb + s); p.LAM[((size_t)k * nct + o_ub + s) * Bm + b] = Q<PF_LAM>(o_ub + s);
                    p.T[((size_t)k * nct + o_lb + s) * Bm + b] = Q<PF_T>(o_lb + s); p.T[((size_t)k * nct + o_ub + s) * Bm + b] = Q<PF_T>(o_ub + s);
                }
            }
            if (is_obst()) {
                const int j = c - 1;
                p.SL[((size_t)k * M + j) * Bm + b] += Q<PF_SL>(j);
                p.SU[((size_t)k * M + j) * Bm + b] += Q<PF_SU>(j);
                const int idx[4] = {o_lg + j, o_ug + j, o_ls + j, o_us + j};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    p.LAM[((size_t)k * nct + idx[i]) * Bm + b] = Q<PF_LAM>(idx[i]);
                    p.T[((size_t)k * nct + idx[i]) * Bm + b] = Q<PF_T>(idx[i]);
                }
            }
        }
        if (tid == 0) {
            p.STATUS[b] = status; p.QP_ITER[b] = kk; p.QP_STATUS[b] = qs;
            p.STATS[0 * Bm + b] = nrm[0]; p.STATS[1 * Bm + b] = nrm[1]; p.STATS[2 * Bm + b] = nrm[2]; p.STATS[3 * Bm + b] = nrm[3];
            p.STATS[4 * Bm + b] = mu; p.STATS[5 * Bm + b] = alpha; p.STATS[6 * Bm + b] = (double)n_itref; p.STATS[7 * Bm + b] = 0.0;
        }
        __syncthreads();
    }
};

template <int MT, int NT>
__global__ void __launch_bounds__(CtaSolver<MT, NT>::NTHR, 4) b2_rti_solve_cta(const B2Params p)
{
    B2_DYNAMIC_SMEM(double, b2_smem);
    __shared__ int s_next;
    const int nq = p.NQ ? *p.NQ : p.B;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(p.NEXT, 1);
        __syncthreads();
        const int i = s_next;
        __syncthreads();
        if (i >= nq) break;
        const int b = p.ORDER ? p.ORDER[i] : i;
        CtaSolver<MT, NT> s(p, b, b2_smem);
        s.run();
    }
}
