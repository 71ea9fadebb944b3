// mpc_cta.cuh -- CTA-per-problem fused SQP-RTI kernel: the LATENCY-oriented mapping of the same solve as mpc_wpp.cuh
// (acados SQP_RTI + HPIPM-style Mehrotra IPM behind src/simulation/robot_ocp_problem.py:195).
//
// mpc_wpp.cuh gives a problem ONE warp: the node-parallel passes of an interior-point iteration (KKT residuals, barrier /
// slack elimination, expansion of the steps, step length) then walk the 6 box slots and M obstacles of a node sequentially,
// and the lanes beyond N+1 idle.  Here a problem owns a CTA and EVERYTHING of the problem lives in shared memory (78 KB at N = 20:
// two problems per SM, no scratch in global memory at all):
//   * obstacle warps (4 at N = 20, M = 5): one thread per (node, obstacle) -- its two general constraints and two slack
//     bounds; the M obstacles of a node sit in adjacent lanes and their contributions to the node's condensed Hessian /
//     gradient / residual are summed by a fixed shuffle tree, then handed to the node warp through shared memory;
//   * node warp (the last one): one thread per node -- the node-level algebra and the six box slots -- and the Riccati
//     recursion of mpc_wpp.cuh (lane = matrix entry) while the obstacle warps wait at the barrier; the two vector recursions of
//     a solve are parallel scans over the nodes (B2_CTA_SCAN below);
//   * problem-wide reductions (mu, step length, norms): warp butterfly -> shared memory -> every thread, in a fixed order.
// Every warp has ONE role: a first version that mixed node-level and obstacle lanes in the same warp left the warps
// permanently split into two convergence groups (measured: 5 active lanes inside the sweeps, every shuffle through the
// WARPSYNC.COLLECTIVE slow path, 3.4x slower than the one-warp kernel).
// Per-problem latency is what this buys (single episodes, small batches: BASELINE config 1 and the p50 single-problem
// latency); instantiated for the reference's dimensions (M = 5; N = 20 and N = 10).  Arithmetic is written with explicit fma()
// like mpc_wpp.cuh; the order of the sums differs from the warp kernel (tree over obstacles instead of a sequential loop), so
// the two kernels agree to rounding, not bit for bit -- each has its own bit-exact CPU mirror (oracle/simt/).
#pragma once
#include "mpc_wpp.cuh"

// B2_CTA_SCAN: the two vector recursions of a KKT solve (p_k = c'_k + Phi_k^T p_{k+1}, dx_{k+1} = f_k + Phi_k dx_k) as PARALLEL
// SCANS over the nodes (Hillis-Steele, ceil(log2 N) levels of one 5 x 5 matrix-vector product per node) instead of N sequential
// steps: the products G(d)_j = Phi_{j-1} Phi_{j-2} ... Phi_{j-2^d} are formed once per factorisation by all threads of the CTA
// (both directions use them: the backward operator of level d at node k is G(d)_{k+2^d} transposed).  More arithmetic, a
// dependent chain of 5 levels instead of 19 nodes: this is the latency kernel.  (The same idea LOST as a throughput kernel,
// profiles/experiments/r01_scan_variant.md: there the instruction count is what matters.)  0 restores the sequential sweeps.
#ifndef B2_CTA_SCAN
#define B2_CTA_SCAN 1
#endif
#define B2_CTA_SCAN_LEVELS 4     // stored product levels G(1) .. G(4): horizons up to 32 nodes
// rows appended to the all-shared layout of mpc_wpp.cuh: the entries of Phi_k (feedback form) + the scan products
#define B2_CTA_PHI_ROWS (25 + (B2_CTA_SCAN ? 25 * B2_CTA_SCAN_LEVELS : 0))

// NOW_: obstacle warps per problem (0 = one (node, obstacle) pair per thread: ceil((N+1) / (32/M)) warps, the latency
// configuration; fewer warps walk the pairs in rounds).  SMASK_: placement of the node-private fields as in mpc_wpp.cuh
// (B2_WPP_ALL_SHARED: everything in shared memory; B2_WPP_ALL_GLOBAL: the private rows in a per-CTA block of L2-resident
// global memory, which is what lets 8-10 problems share an SM).
template <int MT, int NT, int NOW_ = 0, unsigned long long SMASK_ = B2_WPP_ALL_SHARED>
struct CtaSolver {
    static constexpr unsigned long long SMASK = SMASK_;
    static constexpr int M = MT, N = NT, NS = (NT + 1) | 1, GS = NT + 1, nct = 12 + 4 * MT;
    static constexpr int NPW = 32 / MT;                     // nodes per obstacle warp and round
    static constexpr int NOW = NOW_ > 0 ? NOW_ : (NT + 1 + NPW - 1) / NPW;    // obstacle warps
    static constexpr int OROUNDS = (NT + 1 + NPW * NOW - 1) / (NPW * NOW);    // rounds in which they walk the (node, obstacle) pairs
    static constexpr int NODE_WARP = NOW;                   // the node warp (also runs the sequential sweeps)
    static constexpr int NWARPS = NOW + 1;
    static constexpr int NTHR = 32 * NWARPS;
    static_assert(MT >= 1 && MT <= 8 && NT + 1 <= 32 && NWARPS <= 8, "CTA kernel: M <= 8, N <= 31");
    static constexpr int PHI_ROWS = B2_CTA_PHI_ROWS;                     // Phi_k of the feedback form, appended to the all-shared layout
    static constexpr int o_lb = 0, o_lg = 6, o_ub = 6 + MT, o_ug = 12 + MT, o_ls = 12 + 2 * MT, o_us = 12 + 3 * MT;

    const B2Params& p;
    const int b, tid, lane, warp;
    double* const sm;
    double* const gm;       // this CTA's block of node-private rows in global memory (fields whose SMASK bit is clear)
    const WOff o;
    const int PHI;          // first of the 25 rows of Phi
    const int GSC;          // first row of the scan products G(1), G(2), ... (25 rows each)
    static constexpr int LV = NT > 16 ? 5 : NT > 8 ? 4 : NT > 4 ? 3 : NT > 2 ? 2 : 1;      // scan levels: 2^LV >= N
    const size_t Bm;
    const int c;            // obstacle warps: 1 + obstacle index of this thread; node warp: 0
    bool act;               // the thread owns a node (node warp) or, in the current round, a (node, obstacle) pair (obstacle warps)
    int k;                  // its node (clamped into range for idle threads)
    int nc;
    double mu, nrm[4], sum_lt, alpha, S1, S2;
    double rm_cdd, rm_smu;
    double ad_n, ad_d, ap_n, ap_d;
#ifdef B2_DEBUG_CLOCK
    long long clk_fact = 0, clk_sweep = 0, clk_all = 0, clk_head = 0, clk_prep = 0, clk_rs = 0, clk_exp = 0, clk_lin = 0;
    long long clk_fa = 0, clk_fb = 0, clk_fc = 0, clk_fe = 0;       // -DB2_DEBUG_CLOCK_FACT: phases of the factorisation
#endif

    __device__ CtaSolver(const B2Params& p_, int b_, double* sm_, double* gm_)
        : p(p_), b(b_), tid(threadIdx.x), lane(threadIdx.x & 31),
          // the warp index through a shuffle: tells the compiler that it is WARP-UNIFORM, so that `if (warp == 0)` around the
          // sweeps is a uniform branch and the shuffles inside are compiled for a converged warp (without it every shuffle
          // of the factorisation went through the divergent-warp path: WARPSYNC.COLLECTIVE per SHFL, 7x slower sweeps)
          warp(__shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0)), sm(sm_), gm(gm_), o(MT, SMASK, false),
          PHI(WOff(MT, SMASK, false).total), GSC(WOff(MT, SMASK, false).total + 25), Bm((size_t)p_.Bm), c((int)(threadIdx.x >> 5) == NODE_WARP ? 0 : 1 + (int)(threadIdx.x & 31) % MT)
    {
        obst_round(0);
        nc = 0;
        for (int kk = 0; kk <= N; kk++) nc += 2 * ((kk < N ? 2 : 0) + ((kk > 0 && kk < N) ? 4 : 0)) + 4 * M;
    }

    // the (node, obstacle) pair of an obstacle thread in round r (node-warp threads: their node, whatever the round)
    __device__ __forceinline__ void obst_round(int r)
    {
        const int kn = warp == NODE_WARP ? lane : (r * NOW + warp) * NPW + lane / MT;
        act = kn <= NT && (warp == NODE_WARP || lane < NPW * MT);
        k = kn <= NT ? kn : NT;
    }
    __device__ __forceinline__ double& S(int row, int kk) const { return sm[row * NS + kk]; }
    // row i of node-private field F at this thread's node: shared memory or the CTA's block of global memory
    template <int F>
    __device__ __forceinline__ double& Q(int i) const
    {
        if ((SMASK >> F) & 1u) return sm[(o.pf[F] + i) * NS + k];
        return gm[(size_t)((o.pf[F] + i) * GS + k)];
    }
    // fields that other lanes read (v, pi, the steps, the refinement correction, the cost-to-go gradient) stay in shared memory
    static_assert(((SMASK >> PF_V) & 1u) && ((SMASK >> PF_PI) & 1u) && ((SMASK >> PF_DV) & 1u) && ((SMASK >> PF_DPI) & 1u) &&
                  ((SMASK >> PF_CPI) & 1u) && ((SMASK >> PF_PV) & 1u), "CTA kernel: v, pi, dv, dpi, cpi, pv live in shared memory");
    template <int F>
    __device__ __forceinline__ double& R(int i, int kk) const { return sm[(o.pf[F] + i) * NS + kk]; }
    __device__ __forceinline__ double& VV(int i, int kk) const { return sm[(o.pf[PF_V] + i) * NS + kk]; }
    __device__ __forceinline__ double& PP(int i, int kk) const { return sm[(o.pf[PF_PI] + i) * NS + kk]; }
    // start of the rows of private field F (zero fill), its row stride, and its length
    template <int F>
    __device__ __forceinline__ double* field_base() const { return ((SMASK >> F) & 1u) ? sm + o.pf[F] * NS : gm + (size_t)o.pf[F] * GS; }
    template <int F>
    __device__ __forceinline__ static constexpr int field_stride() { return ((SMASK >> F) & 1u) ? NS : GS; }
    __device__ __forceinline__ bool box_on(int s, int kk) const { return s < 2 ? (kk < N) : (kk > 0 && kk < N); }
    __device__ __forceinline__ double scal(int kk) const { return kk < N ? p.dt : 1.0; }
    __device__ __forceinline__ void box_bounds(int s, double& lb, double& ub) const
    {
        const double xv = Q<PF_XU>(s);
        lb = p.lb6[s] - xv; ub = p.ub6[s] - xv;
    }
    __device__ __forceinline__ bool node_warp() const { return warp == NODE_WARP; }       // warp-uniform
    __device__ __forceinline__ bool is_lead() const { return act; }                        // (inside node_warp() branches)
    __device__ __forceinline__ bool is_obst() const { return act; }                        // (inside !node_warp() branches)
    // what the obstacles of a node add to its condensed quantities travels through dead rows of shared memory: the
    // refinement correction cpi (5 rows) and the cost-to-go gradient pv (5 rows) are not live during the passes that use them
    __device__ __forceinline__ double& XCH(int i) const { return i < 5 ? sm[(o.pf[PF_CPI] + i) * NS + k] : sm[(o.pf[PF_PV] + i - 5) * NS + k]; }

    // ---- reductions ------------------------------------------------------------------------------------------------
    // sum over the M adjacent lanes of a node (obstacle warps), result in the lane of obstacle 0: a fixed tree over
    // power-of-two offsets (M = 5: (v0 + v4 + v2) + (v1 + v3)); every lane of the warp takes part in the shuffles
    __device__ __forceinline__ double seg_sum(double v) const
    {
        const int j = c - 1;
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
            if (off < MT) {
                const double t = __shfl_down_sync(0xffffffffu, v, off);
                if (j < off && j + off < MT) v += t;
            }
        }
        return v;
    }

    // problem-wide reduction of up to 8 values per thread: sums (first nsum entries) by warp butterflies, maxima of
    // non-negative values (the rest) by REDUX; the per-warp results meet in shared memory (the AL scratch of the
    // factorisation, idle during the passes) and every thread combines them in warp order
    // (NNEG: maxima of NEGATIVE values, the step-length ratios in HPIPM's negative-reciprocal convention)
    template <int NSUM, int NMAX, int NNEG = 0>
    __device__ __forceinline__ void cta_reduce(double* v)
    {
        constexpr int NV = NSUM + NMAX + NNEG;
        double* const red = sm + (o.total + PHI_ROWS) * NS; // B2_WPP_SCRATCH = 40 doubles >= NWARPS * NV
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
        for (int e = tid; e < nct * field_stride<PF_LAM>(); e += NTHR) {
            field_base<PF_LAM>()[e] = 0.0; field_base<PF_T>()[e] = 0.0; field_base<PF_TINV>()[e] = 0.0; field_base<PF_DD>()[e] = 0.0;
        }
        for (int e = tid; e < 7 * NS; e += NTHR) sm[o.pf[PF_DV] * NS + e] = 0.0;
        for (int e = tid; e < 5 * NS; e += NTHR) sm[o.pf[PF_DPI] * NS + e] = 0.0;
        __syncthreads();
        if (node_warp() && is_lead()) {
            double xk[5], v[7] = {0, 0, 0, 0, 0, 0, 0};
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
        __syncthreads();
        // the obstacle threads need the linearisation point and the (possibly shifted) initial v of their node
        if (!node_warp())
        for (int r_ = 0; r_ < OROUNDS; r_++) {
            obst_round(r_);
            if (!is_obst()) continue;
            const int j = c - 1;
            const double sc = scal(k);
            double xk[5] = {Q<PF_XU>(2), Q<PF_XU>(3), 0.0, 0.0, 0.0};
            const double v2 = VV(2, k), v3 = VV(3, k);
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
        if (node_warp() && is_lead()) {
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
        if (!node_warp())
        for (int r_ = 0; r_ < OROUNDS; r_++) {
        obst_round(r_);
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        double o_rg2 = 0.0, o_rg3 = 0.0, o_h2 = 0.0, o_h3 = 0.0, o_hxy = 0.0, o_gt2 = 0.0, o_gt3 = 0.0;
        if (is_obst()) {
            const int j = c - 1;
            const double v2 = VV(2, k), v3 = VV(3, k);
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
            n0 = fmax(n0, fmax(fabs(rgsl), fabs(rgsu)));
            const double rd0 = cv + s_l - lgj - t0;
            const double rd1 = ugj - cv + s_u - t1;
            const double rd2 = s_l + sli - t2;
            const double rd3 = s_u + sui - t3;
            n2 = fmax(n2, fmax(fmax(fabs(rd0), fabs(rd1)), fmax(fabs(rd2), fabs(rd3))));
            const double m0 = l0 * t0, m1 = l1 * t1, m2 = l2 * t2, m3 = l3 * t3;
            n3 = fmax(n3, fmax(fmax(fabs(m0), fabs(m1)), fmax(fabs(m2), fabs(m3))));
            sum += (m0 + m1) + (m2 + m3);
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
        if (is_obst() && c == 1) {
            XCH(0) = o_rg2; XCH(1) = o_rg3; XCH(2) = o_h2; XCH(3) = o_h3; XCH(4) = o_hxy; XCH(5) = o_gt2; XCH(6) = o_gt3;
        }
        }
        obst_round(0);
        __syncthreads();
        if (node_warp() && is_lead()) {
            rg[2] += XCH(0); rg[3] += XCH(1); Hx[2] += XCH(2); Hx[3] += XCH(3); hxy = XCH(4); gt[2] += XCH(5); gt[3] += XCH(6);
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
        if (node_warp() && is_lead()) {                      // (the barrier above ordered the obstacle threads' reads of the old v in cv)
#pragma unroll
            for (int i = 0; i < 7; i++) S(o.cv + i, k) = Hx[i];
            S(o.hg + 7, k) = hxy;
        }
        cta_reduce<1, 4, 0>(red);
        sum_lt = red[0]; nrm[0] = red[1]; nrm[1] = red[2]; nrm[2] = red[3]; nrm[3] = red[4];
        mu = sum_lt / nc;
        __syncthreads();
    }

    // ---- backward Riccati recursion in square-root form: warp 0, as mpc_wpp.cuh riccati_factorize() ---------------------
    __device__ void riccati_factorize()
    {
        const int e = lane < 28 ? lane : 0;
        // row of the packed lower triangle, branch-free: a `while` loop with a lane-dependent trip count left the lanes of
        // the last row in a convergence group of their own for the whole factorisation (measured active mask 0xf01fffff)
        const int mi = (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15) + (e >= 21);
        const int mj = e - mi * (mi + 1) / 2;
        // (structure and arithmetic of mpc_wpp.cuh riccati_factorize(): U = [B A]^T Sxx with the unnormalised factor, M = H + U D^2 U^T,
        //  the normalisation of the factor after the recursion -- there by the warp, here by all threads: normalize_factor())
        const int ai0 = lane < 30 ? lane / 5 : 0, aq0 = lane % 5;
        int ba0[5], ba1[5];
#pragma unroll
        for (int r = 0; r < 5; r++) { ba0[r] = o.sv + B2_BATAB[r][ai0]; ba1[r] = o.sv + B2_BATAB[r][6]; }
        const bool xx = lane < 28 && mi >= 2 && mj >= 2;
        const bool xu = lane < 28 && mi >= 2 && mj < 2;
        const int xrow = xx ? o.lxx + b2_sym(mi - 2, mj - 2) : 0;
        const int lrow = xu ? o.lxu + 2 * (mi - 2) + mj : 0;
        const bool mine = xx || xu || lane == 1;
        const int myrow = xx ? xrow : xu ? lrow : o.lamu + 1;
        const bool dg = lane < 28 && mi == mj;
        const int sprow = mj == 0 ? o.lamu + 0 : mj == 1 ? o.lamu + 2 : o.lrb + (mj - 2);
        double* const scr = sm + (o.total + PHI_ROWS) * NS;
        for (int kk = N; kk >= 0; kk--) {
#ifdef B2_DEBUG_CLOCK_FACT
            const long long tf0_ = clock64();
            long long tfa_ = tf0_;
#endif
            double m = 0.0;
            if (mi == mj) m = S(o.cv + mi, kk);
            if (mi == 3 && mj == 2) m = S(o.hg + 7, kk);
            if (kk < N) {
                {
                    const double spq = S(o.lrb + aq0, kk + 1);
                    const double dq = spq > 0.0 ? 1.0 / spq : 0.0;
                    double acc = 0.0;
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        const double l = (r >= aq0) ? S(o.lxx + b2_sym(r, aq0 <= r ? aq0 : r), kk + 1) : 0.0;
                        acc = fma(S(ba0[r], kk), l, acc);
                    }
                    // (row 6 of U: computed by every lane, stored by lanes 0..4 -- no divergent branch inside the recursion)
                    double acc1 = 0.0;
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        const double l = (r >= aq0) ? S(o.lxx + b2_sym(r, aq0 <= r ? aq0 : r), kk + 1) : 0.0;
                        acc1 = fma(S(ba1[r], kk), l, acc1);
                    }
                    if (lane < 30) { scr[lane] = acc; scr[40 + lane] = acc * dq; }
                    if (lane < 5) { scr[30 + lane] = acc1; scr[70 + lane] = acc1 * dq; }
                }
                __syncwarp();
#ifdef B2_DEBUG_CLOCK_FACT
                tfa_ = clock64();
#endif
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 5; q++) acc = fma(scr[40 + mi * 5 + q], scr[mj * 5 + q], acc);
                m += acc;
            }
#ifdef B2_DEBUG_CLOCK_FACT
            const long long tfb_ = clock64() + (long long)(m == 1.2345e-300);
#endif
            double sigma = 1.0, my_sp = 0.0;
            // (unrolled: one problem per CTA is latency-bound, the compile-time shuffle lanes save the index arithmetic of the
            //  rolled loop that the throughput kernel prefers for its instruction-cache footprint)
            // (Experiment, negative: TWO columns per exchange round -- 7 broadcasts, the second step's pivot and column entries
            //  recomputed by every lane that needs them, bit-identical -- halves the dependent exchange rounds but adds 9 FP64
            //  operations per lane and pair: 0.818 against 0.811 ms.  The rounds are issue-bound, not latency-bound.)
            auto column = [&](const int cc) {
                const double pv_ = b2_bcast(m, cc * (cc + 1) / 2 + cc);
                const double la = b2_bcast(m, mi * (mi + 1) / 2 + (cc <= mi ? cc : mi));
                const double lb = b2_bcast(m, mj * (mj + 1) / 2 + (cc <= mj ? cc : mj));
                const bool ok = pv_ > 0.0;
                if (mj == cc) my_sp = ok ? sigma * pv_ : 0.0;
#ifdef B2_CHOL_BRANCH
                if (ok) {
                    const double scale = __hiloint2double(0x7fe00000 - (__double2hiint(pv_) & 0x7ff00000), 0);
                    if (mj > cc) m = fma(m, pv_, -(la * lb)) * scale;
                    sigma *= pv_ * scale;
                }
#else
                {
                    // scale = 2^-e(p): exponent field 2046 - e, i.e. 0x7fe00000 - (high word & 0x7ff00000) (p > 0, finite).  The
                    // update is written as fma(m, p, -(la lb)) * scale: bit-identical to fma(m, p scale, -(la lb) scale) (a power
                    // of two commutes with the rounding), one multiplication shorter on the dependent chain.  Branch-free: a
                    // non-positive (or NaN) pivot leaves m and sigma as they are by selects, not by a divergent branch
                    const double scale = __hiloint2double(0x7fe00000 - (__double2hiint(pv_) & 0x7ff00000), 0);
                    const double mn = fma(m, pv_, -(la * lb)) * scale, sn = sigma * (pv_ * scale);
                    m = (ok && mj > cc) ? mn : m;
                    sigma = ok ? sn : sigma;
                }
#endif
            };
            column(0); column(1);
            if (kk > 0) {                                                // (node 0: the state is eliminated, only the control block is used)
#pragma unroll
                for (int cc = 2; cc < 6; cc++) column(cc);
            }
            // (the last column has nothing below or to the right of its pivot: its lane needs sigma and its own entry, no exchange)
            if (mj == 6) my_sp = m > 0.0 ? sigma * m : 0.0;
#ifdef B2_DEBUG_CLOCK_FACT
            const long long tfc_ = clock64() + (long long)(m == 1.2345e-300);
#endif
            if (mine) S(myrow, kk) = m;                                  // (see mpc_wpp.cuh: loop-invariant predicate)
            if (dg) S(sprow, kk) = my_sp;
            __syncwarp();
#ifdef B2_DEBUG_CLOCK_FACT
            const long long tfe_ = clock64();
            clk_fa += tfa_ - tf0_; clk_fb += tfb_ - tfa_; clk_fc += tfc_ - tfb_; clk_fe += tfe_ - tfc_;
#endif
        }
    }

    // end of the factorisation, all threads of the CTA, one (node, column) pair each: L_ic = S_ic / sqrt(sigma_c p_c), 1/L00, 1/L11
    // (see mpc_wpp.cuh normalize_factor(); the thread of control column 0 handles both control columns, whose scales it overwrites)
    __device__ void normalize_factor()
    {
        for (int e = tid; e < 7 * (N + 1); e += NTHR) {
            const int kk = e / 7, cq = e - 7 * kk;
            if (cq == 1) continue;
            if (cq == 0) {
                const double sp0 = S(o.lamu + 0, kk), sp1 = S(o.lamu + 2, kk);
                const double rs0 = sp0 > 0.0 ? 1.0 / sqrt(sp0) : 0.0, rs1 = sp1 > 0.0 ? 1.0 / sqrt(sp1) : 0.0;
                if (kk < N) {
                    const double sig1 = sp0 > 0.0 ? sp0 * __hiloint2double(0x7fe00000 - (__double2hiint(sp0) & 0x7ff00000), 0) : 1.0;
                    S(o.lamu + 0, kk) = rs0; S(o.lamu + 2, kk) = sig1 * rs1;
                    S(o.lamu + 1, kk) *= rs0;
                }
                if (kk > 0) {
#pragma unroll
                    for (int i = 0; i < 5; i++) { S(o.lxu + 2 * i, kk) *= rs0; S(o.lxu + 2 * i + 1, kk) *= rs1; }
                }
            } else if (kk > 0) {
                const int j = cq - 2;
                const double spj = S(o.lrb + j, kk);
                const double rsj = spj > 0.0 ? 1.0 / sqrt(spj) : 0.0;
                for (int i = j; i < 5; i++) S(o.lxx + b2_sym(i, j), kk) *= rsj;
            }
        }
        __syncthreads();
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
        if (!node_warp()) {
            for (int r_ = 0; r_ < OROUNDS; r_++) {
            obst_round(r_);
            const double sc = scal(k);
            const double Zuq = sc * p.Zu;
            double o2 = 0.0, o3 = 0.0;
            if (is_obst()) {
                const int j = c - 1;
                const double G0 = Q<PF_LAM>(o_lg + j) * Q<PF_TINV>(o_lg + j), G1 = Q<PF_LAM>(o_ug + j) * Q<PF_TINV>(o_ug + j);
                double coef;
                if (kind == 3) {
                    coef = fma(G1 * Q<PF_LRGSU>(j), Q<PF_DUI>(j), -((G0 * Q<PF_LRGSL>(j)) * Q<PF_DLI>(j)));
                } else {
                    double rhol, rhou, q0, q1;
                    slack_rho(j, VV(2, k), VV(3, k), Zuq, sc, rhol, rhou, q0, q1);
                    coef = fma(-(G0 * rhol), Q<PF_DLI>(j), q0) - fma(-(G1 * rhou), Q<PF_DUI>(j), q1);
                }
                o2 = Q<PF_CX>(j) * coef; o3 = Q<PF_CY>(j) * coef;
            }
            o2 = seg_sum(o2); o3 = seg_sum(o3);
            if (is_obst() && c == 1) { XCH(0) = o2; XCH(1) = o3; }
            }
            obst_round(0);
        } else if (is_lead() && kind != 3) {
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
        __syncthreads();
        if (node_warp() && is_lead()) {
            if (kind == 3) { S(o.hg + 2, k) += XCH(0); S(o.hg + 3, k) += XCH(1); }       // lin_residuals left lrg in hg
            else {
                gt[2] += XCH(0); gt[3] += XCH(1);
#pragma unroll
                for (int i = 0; i < 7; i++) S(o.hg + i, k) = gt[i];
            }
        }
        // (no block barrier: the node warp, which wrote hg, is the one that runs the sweeps next)
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

    // Feedback form of the factorisation, once per interior-point iteration (node-parallel, node warp):
    //   G_k = Lxu Luu^-1 (5 x 2, stored over Lxu, which nothing reads afterwards),  Phi_k = A_k - B_k G_k^T (5 x 5).
    // With them the sequential part of every KKT solve is one 5 x 5 matrix-vector product per node,
    //   backward  p_k = c'_k + Phi_k^T p_{k+1},      c'_k = gx_k - G_k gu_k + Phi_k^T z_k,   z_k = P_{k+1} rb_k
    //   forward   dx_{k+1} = f_k + Phi_k dx_k,        f_k = rb_k + B_k du0_k,  du0_k = -Luu^-T Luu^-1 (gu_k + B_k^T (p_{k+1} + z_k))
    // and everything else (z, c', du0, f, du_k = du0_k - G_k^T dx_k, dpi_k = P_{k+1} dx_{k+1} + p_{k+1}) is node-parallel.
    // Same mathematics as the substitution with the factor that mpc_wpp.cuh (and HPIPM) runs inside the sweep; the dependent
    // chain per node shrinks from ~350 to ~100 cycles, which is what a latency-bound single problem needs.
    // Node 0 has no G / Phi: its state is eliminated (dx_0 = 0, p_0 is never formed).
    __device__ void feedback_matrices()
    {
        if (node_warp() && is_lead() && k >= 1 && k < N) {
            const double l00i = S(o.lamu + 0, k), l10 = S(o.lamu + 1, k), l11i = S(o.lamu + 2, k);
            double G0[5], G1[5];
#pragma unroll
            for (int i = 0; i < 5; i++) {
                G1[i] = S(o.lxu + 2 * i + 1, k) * l11i;
                G0[i] = fma(-G1[i], l10, S(o.lxu + 2 * i, k)) * l00i;
                S(o.lxu + 2 * i, k) = G0[i]; S(o.lxu + 2 * i + 1, k) = G1[i];
            }
            const double h = p.dt, h2 = p.dt2h;
            const double b00 = S(o.sv + SV_B00, k), b01 = S(o.sv + SV_B01, k), b10 = S(o.sv + SV_B10, k), b11 = S(o.sv + SV_B11, k);
            const double a0[5] = {1.0, 0.0, S(o.sv + SV_A02, k), S(o.sv + SV_A03, k), S(o.sv + SV_A04, k)};
            const double a1[5] = {0.0, 1.0, S(o.sv + SV_A12, k), S(o.sv + SV_A13, k), S(o.sv + SV_A14, k)};
#pragma unroll
            for (int j = 0; j < 5; j++) {
                S(PHI + 0 * 5 + j, k) = a0[j] - fma(b00, G0[j], b01 * G1[j]);
                S(PHI + 1 * 5 + j, k) = a1[j] - fma(b10, G0[j], b11 * G1[j]);
                S(PHI + 2 * 5 + j, k) = (j == 2 ? 1.0 : (j == 4 ? h : 0.0)) - h2 * G1[j];
                S(PHI + 3 * 5 + j, k) = (j == 3 ? 1.0 : 0.0) - h * G0[j];
                S(PHI + 4 * 5 + j, k) = (j == 4 ? 1.0 : 0.0) - h * G1[j];
            }
        }
    }

    // entry (r, q) of G(d)_j, d >= 0:  G(0)_j = Phi_{j-1},  G(d)_j = Phi_{j-1} ... Phi_{j-2^d}  (defined for j - 2^d >= 1)
    __device__ __forceinline__ double& GD(int d, int r, int q, int j) const
    {
        return d == 0 ? S(PHI + r * 5 + q, j - 1) : S(GSC + (d - 1) * 25 + r * 5 + q, j);
    }
    // G(d+1)_j = G(d)_j G(d)_{j-2^d} for every node j >= 2^(d+1) + 1, one thread per (node, row); all threads of the CTA, once per
    // factorisation (block barriers inside: every thread of the CTA must call it)
    __device__ void scan_products()
    {
#if B2_CTA_SCAN
#pragma unroll 1
        for (int d = 0; d + 1 < LV; d++) {
            const int s_ = 1 << d;
            for (int e = tid; e < 5 * (N + 1); e += NTHR) {
                const int j = e / 5, r = e - 5 * j;
                if (j < 2 * s_ + 1 || j > N) continue;
                double a[5], acc[5];
#pragma unroll
                for (int q = 0; q < 5; q++) a[q] = GD(d, r, q, j);
#pragma unroll
                for (int cq = 0; cq < 5; cq++) {
                    double t = a[0] * GD(d, 0, cq, j - s_);
#pragma unroll
                    for (int q = 1; q < 5; q++) t = fma(a[q], GD(d, q, cq, j - s_), t);
                    acc[cq] = t;
                }
#pragma unroll
                for (int cq = 0; cq < 5; cq++) GD(d + 1, r, cq, j) = acc[cq];
            }
            __syncthreads();
        }
#endif
    }

    // (Experiment, negative: every lane carrying the whole 5-vector with all 25 entries of Phi as broadcast loads -- no shuffle
    //  on the chain -- measured 2.1x SLOWER sweeps (38.6 k vs 18.3 k cycles per iteration): 25 FP64 operations per node and
    //  lane on a half-rate pipe cost more than the 26-cycle broadcast they avoid.)
    // one KKT solve with the current factorisation in feedback form (see feedback_matrices()); kind 3 reads the linear
    // residual (lrb) and writes the correction to cv / cpi.  Everything runs in the node warp: node-parallel phases with
    // lane = node, the two sequential recursions with lanes 0..4 = vector components
    __device__ void riccati_solve(int kind)
    {
        const bool itr = (kind == 3);
        const int oRB = itr ? o.lrb : o.rb, oDV = itr ? o.cv : o.pf[PF_DV], oDPI = itr ? o.pf[PF_CPI] : o.pf[PF_DPI];
        const double h = p.dt, h2 = p.dt2h;
        if (node_warp()) {
            // ---- pre: z_k, c'_k
            if (is_lead() && k < N) {
                double rb[5], z[5];
#pragma unroll
                for (int i = 0; i < 5; i++) rb[i] = S(oRB + i, k);
                ctg_mul(k + 1, rb, z);
#pragma unroll
                for (int i = 0; i < 5; i++) S(oDPI + i, k) = z[i];                // z_k (overwritten by dpi_k below)
                if (k >= 1) {
                    const double gu0 = S(o.hg + 0, k), gu1 = S(o.hg + 1, k);
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        double acc = S(o.hg + 2 + i, k) - fma(S(o.lxu + 2 * i, k), gu0, S(o.lxu + 2 * i + 1, k) * gu1);
#pragma unroll
                        for (int j = 0; j < 5; j++) acc = fma(S(PHI + j * 5 + i, k), z[j], acc);
                        S(o.hg + 2 + i, k) = acc;
                    }
                }
            }
            __syncwarp();
#ifdef B2_DEBUG_CLOCK
            const long long c1_ = clock64();
#endif
            // ---- backward: p_k = c'_k + Phi_k^T p_{k+1}
#if B2_CTA_SCAN
            // suffix scan, lane = node k: after level d the row holds p_k up to the influence of p_{k+2^(d+1)}; node N (p_N given,
            // no operator) ends every chain.  Branch-free: every lane computes (indices clamped), the lanes of the level store.
            if (is_lead() && k >= 1) {
#pragma unroll
                for (int i = 0; i < 5; i++) R<PF_PV>(i, k) = S(o.hg + 2 + i, k);
            }
            __syncwarp();
#pragma unroll
            for (int d = 0; d < LV; d++) {
                const int s_ = 1 << d;
                const bool on = is_lead() && k >= 1 && k + s_ <= N;
                const int kn = k + s_ <= N ? k + s_ : N, kc = k >= 1 ? k : 1;
                double cn[5], acc[5];
#pragma unroll
                for (int q = 0; q < 5; q++) cn[q] = R<PF_PV>(q, kn);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    double t = R<PF_PV>(i, kc);
#pragma unroll
                    for (int q = 0; q < 5; q++) t = fma(GD(d, q, i, kn), cn[q], t);
                    acc[i] = t;
                }
                __syncwarp();
                if (on) {
#pragma unroll
                    for (int i = 0; i < 5; i++) R<PF_PV>(i, k) = acc[i];
                }
                __syncwarp();
            }
#else
            const int c5 = lane < 5 ? lane : 4;
            double pf[5];
            {
                const double pc = S(o.hg + 2 + c5, N);
                if (lane < 5) R<PF_PV>(lane, N) = pc;
#pragma unroll
                for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
            }
            for (int kk = N - 1; kk >= 1; kk--) {
                const double t01 = fma(S(PHI + 1 * 5 + c5, kk), pf[1], S(PHI + 0 * 5 + c5, kk) * pf[0]);
                const double t23 = fma(S(PHI + 3 * 5 + c5, kk), pf[3], S(PHI + 2 * 5 + c5, kk) * pf[2]);
                const double pc = (fma(S(PHI + 4 * 5 + c5, kk), pf[4], S(o.hg + 2 + c5, kk)) + t01) + t23;
                if (lane < 5) R<PF_PV>(lane, kk) = pc;
#pragma unroll
                for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
            }
#endif
#ifdef B2_DEBUG_CLOCK
            clk_sweep += clock64() - c1_;
#endif
            __syncwarp();
            // ---- mid: du0_k, f_k
            if (is_lead() && k < N) {
                double sx[5];
#pragma unroll
                for (int i = 0; i < 5; i++) sx[i] = R<PF_PV>(i, k + 1) + S(oDPI + i, k);
                const double b00 = S(o.sv + SV_B00, k), b01 = S(o.sv + SV_B01, k), b10 = S(o.sv + SV_B10, k), b11 = S(o.sv + SV_B11, k);
                const double w0 = fma(b00, sx[0], S(o.hg + 0, k)) + fma(b10, sx[1], h * sx[3]);
                const double w1 = fma(h, sx[4], fma(b01, sx[0], S(o.hg + 1, k)) + fma(b11, sx[1], h2 * sx[2]));
                const double l00i = S(o.lamu + 0, k), l10 = S(o.lamu + 1, k), l11i = S(o.lamu + 2, k);
                const double lu0 = w0 * l00i;
                const double lu1 = fma(-l10, lu0, w1) * l11i;
                const double du1 = -lu1 * l11i;
                const double du0 = fma(-l10, du1, -lu0) * l00i;
                S(oDV + 0, k) = du0; S(oDV + 1, k) = du1;
                S(oDV + 2 + 0, k + 1) = S(oRB + 0, k) + fma(b00, du0, b01 * du1);
                S(oDV + 2 + 1, k + 1) = S(oRB + 1, k) + fma(b10, du0, b11 * du1);
                S(oDV + 2 + 2, k + 1) = fma(h2, du1, S(oRB + 2, k));
                S(oDV + 2 + 3, k + 1) = fma(h, du0, S(oRB + 3, k));
                S(oDV + 2 + 4, k + 1) = fma(h, du1, S(oRB + 4, k));
            }
            if (lane < 5) S(oDV + 2 + lane, 0) = 0.0;
            if (lane == 0) { S(oDV + 0, N) = 0.0; S(oDV + 1, N) = 0.0; }
            __syncwarp();
#ifdef B2_DEBUG_CLOCK
            const long long c2_ = clock64();
#endif
            // ---- forward: dx_{k+1} = f_k + Phi_k dx_k  (dx_1 = f_0)
#if B2_CTA_SCAN
            // prefix scan, lane = node j: the row dv[2..6](j) holds f_{j-1} on entry (written by the mid phase) and dx_j on exit
#pragma unroll
            for (int d = 0; d < LV; d++) {
                const int s_ = 1 << d;
                const bool on = is_lead() && k - s_ >= 1;
                const int kc = k >= s_ + 1 ? k : s_ + 1;        // (clamped: G(d)_j exists for j >= 2^d + 1)
                double yn[5], acc[5];
#pragma unroll
                for (int q = 0; q < 5; q++) yn[q] = S(oDV + 2 + q, kc - s_);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    double t = S(oDV + 2 + i, kc);
#pragma unroll
                    for (int q = 0; q < 5; q++) t = fma(GD(d, i, q, kc), yn[q], t);
                    acc[i] = t;
                }
                __syncwarp();
                if (on) {
#pragma unroll
                    for (int i = 0; i < 5; i++) S(oDV + 2 + i, k) = acc[i];
                }
                __syncwarp();
            }
#else
            double dx[5];
            {
                const double xn = S(oDV + 2 + c5, 1);
#pragma unroll
                for (int r = 0; r < 5; r++) dx[r] = b2_bcast(xn, r);
            }
            for (int kk = 1; kk < N; kk++) {
                const double t01 = fma(S(PHI + c5 * 5 + 1, kk), dx[1], S(PHI + c5 * 5 + 0, kk) * dx[0]);
                const double t23 = fma(S(PHI + c5 * 5 + 3, kk), dx[3], S(PHI + c5 * 5 + 2, kk) * dx[2]);
                const double xn = (fma(S(PHI + c5 * 5 + 4, kk), dx[4], S(oDV + 2 + c5, kk + 1)) + t01) + t23;
                if (lane < 5) S(oDV + 2 + lane, kk + 1) = xn;
#pragma unroll
                for (int r = 0; r < 5; r++) dx[r] = b2_bcast(xn, r);
            }
#endif
#ifdef B2_DEBUG_CLOCK
            clk_sweep += clock64() - c2_;
#endif
            __syncwarp();
            // ---- post: du_k = du0_k - G_k^T dx_k, dpi_k = P_{k+1} dx_{k+1} + p_{k+1}
            if (is_lead() && k < N) {
                if (k >= 1) {
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const double dxi = S(oDV + 2 + i, k);
                        d0 = fma(S(o.lxu + 2 * i, k), dxi, d0); d1 = fma(S(o.lxu + 2 * i + 1, k), dxi, d1);
                    }
                    S(oDV + 0, k) -= d0; S(oDV + 1, k) -= d1;
                }
                double dxn[5], y[5];
#pragma unroll
                for (int i = 0; i < 5; i++) dxn[i] = S(oDV + 2 + i, k + 1);
                ctg_mul(k + 1, dxn, y);
#pragma unroll
                for (int i = 0; i < 5; i++) S(oDPI + i, k) = y[i] + R<PF_PV>(i, k + 1);
            }
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
    __device__ __forceinline__ void expand_box(int kind, double& s1, double& s2, double* lrg)
    {
        const bool keep_dd = (MODE == XP_SOLVE) && kind == 0 && p.pred_corr;
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
    }
    // ... and of the (node, obstacle) pair of an obstacle thread in the current round
    template <int MODE>
    __device__ __forceinline__ void expand_obst(int kind, double& s1, double& s2, double& o2, double& o3, double& ln0)
    {
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        const bool keep_dd = (MODE == XP_SOLVE) && kind == 0 && p.pred_corr;
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
                ln0 = fmax(ln0, fmax(fabs(a0), fabs(a1)));
            }
        }
    }

    // one linear solve with the current factorisation + step-length reductions (kind 0 affine, 1 corrector, 2 centering,
    // 3 one iterative-refinement step)
    __device__ void solve_any(int kind, double smu)
    {
        rm_cdd = kind == 1 ? 1.0 : 0.0;
        rm_smu = kind == 0 ? 0.0 : smu;
#ifdef B2_DEBUG_CLOCK
        const long long cp_ = clock64();
#endif
        if (kind != 0) solve_prep(kind);
#ifdef B2_DEBUG_CLOCK
        const long long cr_ = clock64(); clk_prep += cr_ - cp_;
#endif
        riccati_solve(kind);
#ifdef B2_DEBUG_CLOCK
        const long long ce_ = clock64(); clk_rs += ce_ - cr_;
#endif
        if (kind == 3) {
            // add the refinement correction to the stored step before anybody expands it: dv += cv, dpi += cpi
            if (node_warp() && is_lead()) {
#pragma unroll
                for (int i = 0; i < 7; i++) R<PF_DV>(i, k) += S(o.cv + i, k);
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < 5; i++) R<PF_DPI>(i, k) += R<PF_CPI>(i, k);
                }
            }
            __syncthreads();
        }
        double s1 = 0.0, s2 = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0;
        ad_n = 1.0; ad_d = -1.0; ap_n = 1.0; ap_d = -1.0;
        if (node_warp()) {
            if (kind == 3) expand_box<XP_REFINE>(3, s1, s2, nullptr);
            else expand_box<XP_SOLVE>(kind, s1, s2, nullptr);
        } else {
            for (int r_ = 0; r_ < OROUNDS; r_++) {
                obst_round(r_);
                if (kind == 3) expand_obst<XP_REFINE>(3, s1, s2, d0, d1, d2);
                else expand_obst<XP_SOLVE>(kind, s1, s2, d0, d1, d2);
            }
            obst_round(0);
        }
        // negative-reciprocal convention (HPIPM d_compute_alpha_qp): the largest of the negative ratios limits the step
        double red[4] = {s1, s2, ap_n / ap_d, ad_n / ad_d};
        cta_reduce<2, 0, 2>(red);
        S1 = red[0]; S2 = red[1];
        alpha = -(red[2] > red[3] ? red[2] : red[3]);
        __syncthreads();
#ifdef B2_DEBUG_CLOCK
        clk_exp += clock64() - ce_;
#endif
    }

    // residual of the linear system at the stored step; leaves lrg in rows hg[0..6] and lrb in rows lrb[0..4]
    __device__ void lin_residuals(double lnrm[2])
    {
        double ln0 = 0.0, ln1 = 0.0, o2 = 0.0, o3 = 0.0, s1 = 0.0, s2 = 0.0;
        double lrg[7] = {0, 0, 0, 0, 0, 0, 0}, dv[7] = {0, 0, 0, 0, 0, 0, 0}, sv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (node_warp() && is_lead()) {
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
        if (node_warp()) expand_box<XP_LIN>(0, s1, s2, lrg);
        else {
            for (int r_ = 0; r_ < OROUNDS; r_++) {
                obst_round(r_);
                o2 = 0.0; o3 = 0.0;
                expand_obst<XP_LIN>(0, s1, s2, o2, o3, ln0);
                o2 = seg_sum(o2); o3 = seg_sum(o3);
                if (is_obst() && c == 1) { XCH(0) = o2; XCH(1) = o3; }
            }
            obst_round(0);
        }
        __syncthreads();
        if (node_warp() && is_lead()) {
            lrg[2] += XCH(0); lrg[3] += XCH(1);
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
#ifdef B2_DEBUG_CLOCK
        const long long call0_ = clock64();
#endif
        prepare_and_init();
        alpha = 1.0;
        int kk = 0;
        const double inv_nc = 1.0 / nc;
        double a = 0.0;
        for (;; kk++) {
#ifdef B2_DEBUG_CLOCK
            const long long ch_ = clock64();
#endif
            head(a);
#ifdef B2_DEBUG_CLOCK
            clk_head += clock64() - ch_;
#endif
            if (!(kk < p.qp_iter_max && alpha > p.alpha_min &&
                  (nrm[0] > p.res_g_max || nrm[1] > p.res_b_max || nrm[2] > p.res_d_max || nrm[3] > p.res_m_max))) break;
#ifdef B2_DEBUG_CLOCK
            const long long c0_ = clock64();
#endif
            if (node_warp()) riccati_factorize();
            __syncthreads();
            normalize_factor();
            if (node_warp()) feedback_matrices();
            __syncthreads();
            scan_products();
#ifdef B2_DEBUG_CLOCK
            clk_fact += clock64() - c0_;
#endif
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
#ifdef B2_DEBUG_CLOCK
                const long long cl_ = clock64();
#endif
                lin_residuals(lnrm);
#ifdef B2_DEBUG_CLOCK
                clk_lin += clock64() - cl_;
#endif
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
            if (node_warp() && is_lead()) {
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
            if (!node_warp())
            for (int r_ = 0; r_ < OROUNDS; r_++) {
                obst_round(r_);
                if (!is_obst()) continue;
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
#ifdef B2_DEBUG_CLOCK
        clk_all = clock64() - call0_;
        if (tid == NODE_WARP * 32) {
            p.STATS[5 * Bm + b] = (double)clk_fact; p.STATS[6 * Bm + b] = (double)clk_sweep; p.STATS[7 * Bm + b] = (double)clk_all;
            p.STATS[0 * Bm + b] = (double)clk_head; p.STATS[1 * Bm + b] = (double)clk_prep; p.STATS[2 * Bm + b] = (double)clk_rs;
            p.STATS[3 * Bm + b] = (double)clk_exp; p.STATS[4 * Bm + b] = (double)clk_lin;
#ifdef B2_DEBUG_CLOCK_FACT
            p.STATS[0 * Bm + b] = (double)clk_fa; p.STATS[1 * Bm + b] = (double)clk_fb; p.STATS[2 * Bm + b] = (double)clk_fc; p.STATS[3 * Bm + b] = (double)clk_fe;
#endif
        }
        if (tid != 0 || true) {}
#endif
        if (tid == 0) {
            p.STATUS[b] = status; p.QP_ITER[b] = kk; p.QP_STATUS[b] = qs;
#ifndef B2_DEBUG_CLOCK
            p.STATS[0 * Bm + b] = nrm[0]; p.STATS[1 * Bm + b] = nrm[1]; p.STATS[2 * Bm + b] = nrm[2]; p.STATS[3 * Bm + b] = nrm[3];
            p.STATS[4 * Bm + b] = mu;
#endif
#ifndef B2_DEBUG_CLOCK
            p.STATS[5 * Bm + b] = alpha; p.STATS[6 * Bm + b] = (double)n_itref; p.STATS[7 * Bm + b] = 0.0;
#endif
        }
        __syncthreads();
    }
};

#ifndef B2_CTA_MINB
#define B2_CTA_MINB 2           // resident CTAs per SM the kernel is compiled for: 168 registers.  (78 KB of shared memory per problem
                                // would allow 3 at 128 registers: measured 6.5 % slower per problem, 1.426 vs 1.333 ms)
#endif
// dynamic shared memory of the CTA kernel: the shared rows of the chosen placement (mpc_wpp.cuh WOff) + the Phi rows + scratch
__host__ __device__ inline size_t b2_cta_smem_bytes(int N, int M, unsigned long long smask = B2_WPP_ALL_SHARED)
{
    return sizeof(double) * ((size_t)(WOff(M, smask, false).total + B2_CTA_PHI_ROWS) * b2_wpp_stride(N) + B2_WPP_SCRATCH);
}

// MINB: resident CTAs per SM the instantiation is compiled for (register budget 65536 / (MINB * threads)).
//   latency kernel  <5, N, 0, ALL_SHARED, 2>: 168 registers (78 KB of shared memory per problem; at 61 KB, before the scan products, 3 CTAs at 128
//                                             registers: measured 6.5 % slower per problem, 1.426 vs 1.333 ms)
//   throughput candidates <5, 20, 1 or 2, ALL_GLOBAL, 8 or 6>: node-private rows in L2 like the warp kernel, 22 KB of shared memory
template <int MT, int NT, int NOW_ = 0, unsigned long long SMASK_ = B2_WPP_ALL_SHARED, int MINB = 2>
__global__ void __launch_bounds__(CtaSolver<MT, NT, NOW_, SMASK_>::NTHR, MINB) b2_rti_solve_cta(const B2Params p)
{
    B2_DYNAMIC_SMEM(double, b2_smem);
    __shared__ int s_next;
    double* const gm = p.WPRIV + (size_t)blockIdx.x * p.wpriv_stride;      // this CTA's block of node-private rows (if any)
    const int nq = p.NQ ? *p.NQ : p.B;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(p.NEXT, 1);
        __syncthreads();
        const int i = s_next;
        __syncthreads();
        if (i >= nq) break;
        const int b = p.ORDER ? p.ORDER[i] : i;
        CtaSolver<MT, NT, NOW_, SMASK_> s(p, b, b2_smem, gm);
        s.run();
    }
}
