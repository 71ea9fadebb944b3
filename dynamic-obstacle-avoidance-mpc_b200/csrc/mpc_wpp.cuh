// mpc_wpp.cuh -- warp-per-problem fused SQP-RTI kernel: the whole solve of one MPC problem stays
// on chip (shared memory + registers) from linearisation to the full-step update.
//
// Same mathematics as mpc_tpp.cuh / the oracle (acados SQP_RTI + HPIPM-style Mehrotra IPM behind
// src/simulation/robot_ocp_problem.py:195), different mapping:
//   * node-parallel phases (linearise, constraint evaluation, residuals, multiplier/slack
//     elimination, expansion, step length, update): lane k owns shooting node k; reductions
//     (step length, mu, residual norms) are warp shuffles;
//   * the sequential backward/forward Riccati sweeps walk the nodes with the 7x7 node algebra
//     spread over the lanes: lane e < 28 owns entry (i,j) of the symmetric Riccati matrix in the
//     factorisation, lanes 0..6 own rows/columns in the vector sweeps; exchange by shuffles;
//   * per-node data that other lanes read (model sensitivities, Riccati factors, steps, v, pi) live in
//     shared memory as rows [field][node]: lane==node accesses are conflict-free, the odd row stride keeps
//     lane==field accesses conflict-free too;
//   * NODE-PRIVATE fields (multipliers, slacks of the inequality constraints and their steps, the QP's
//     constraint data: only ever touched by the lane that owns the node) are addressed through Q<field>():
//     the template mask SMASK places each field either in shared memory or in a per-CTA block of global
//     memory (rows [field][node], coalesced, L2-resident: resident CTAs x ~40 KB).  Fewer shared-memory
//     rows = more resident problems per SM, the lever that matters for this latency-bound kernel.
//   * horizons with more than 32 nodes: a lane owns nodes lane, lane+32, ... (node-parallel phases loop).
// One CTA = one warp = one problem at a time; CTAs are persistent and pull problems from a counter.
//
// Numerical note (why dlam is STORED and never re-derived): dlam_i = -(rm_i + lam_i dt_i)/t_i carries
// lam/t up to ~1e18 on slack bounds, so iterative refinement only works if the correction is added
// to the stored dlam; re-deriving dlam from the refined dv would draw a fresh rounding sample.
#pragma once
#include "mpc_common.cuh"

// table of [B A] (5 x 7, rows x+, cols [u_a,u_alpha,x,y,psi,v,omega]) as indices into the extended
// per-node vector ext = [sv0..sv9, 1, 0, h, h^2/2]  (pattern checked in mpc_common.cuh)
__constant__ unsigned char B2_BATAB[5][7] = {
    {SV_B00, SV_B01, 10, 11, SV_A02, SV_A03, SV_A04},
    {SV_B10, SV_B11, 11, 10, SV_A12, SV_A13, SV_A14},
    {11, 13, 11, 11, 10, 11, 12},
    {12, 11, 11, 11, 11, 10, 11},
    {11, 12, 11, 11, 11, 11, 10},
};

// node-private fields (see header): rows per field are nct for the first four, M for the obstacle fields
enum { PF_LAM = 0, PF_T, PF_TINV, PF_DD, PF_DT,              // per constraint; dd = dlam (or dt*dlam of the predictor), dt = dt
       PF_B, PF_G, PF_RG, PF_XU,                            // 5, 7, 7, 6 rows
       PF_CX, PF_CY, PF_LG, PF_UG, PF_ZLQ, PF_ZZLQ, PF_SLI, PF_SUI, PF_SL, PF_SU,
       PF_LRGSL, PF_LRGSU, PF_DSL, PF_DSU, PF_DLI, PF_DUI,
       PF_WPI, PF_WLAM,                                      // merit-function weights of the SQP line search (SQP kernels only)
       PF_V, PF_PI,                                          // QP iterate v (7 rows) and pi (5): touched by the owner lane and its
                                                             // neighbours k-1 / k+1 only, never inside a sequential sweep
       PF_DV, PF_DPI, PF_CPI, PF_PV,                         // step dv (7), dpi (5), refinement correction cpi (5), cost-to-go gradient
                                                             // pv (5): the sweeps only STORE them (z_k, kept in the dpi / cpi rows, is
                                                             // read by the backward sweep: prefetched one node ahead when in global memory)
       PF_COUNT };
#define B2_WPP_ALL_SHARED 0xffffffffffffffffull      // (every placement bit; 33 fields -> 64-bit mask)
// Placement used by default: EVERY node-private field in global memory.  Measured on B200 (N=20, M=5, 4096 problems,
// profiles/experiments/r01_hybrid_layout.md): all global, 12 CTAs/SM 11.6 ms; lam,t shared, 7 CTAs/SM 13.1 ms;
// lam,t,1/t,dlam shared, 5 CTAs/SM 14.3 ms; all shared, 4 CTAs/SM 12.7 ms.
#define B2_WPP_ALL_GLOBAL ((1ull << PF_V) | (1ull << PF_PI) | (1ull << PF_DV) | (1ull << PF_DPI) | (1ull << PF_CPI) | (1ull << PF_PV))
                                           // ... except v, pi, dv, dpi, cpi, pv (34 rows), which stay in shared memory
// Long horizons (N = 100: 103 shared rows x 101 nodes = 85 KB, 2 CTAs/SM): those 34 rows in global memory as well ->
// 69 rows, 56 KB, 4 CTAs/SM.
#define B2_WPP_LONG 0ull
// The six box slots of a node: unrolled by default.  -DB2_WPP_ROLL_BOX walks them with a rolled loop (select chains /
// predicated adds instead of compile-time register indices; same arithmetic, bit-identical results): 20 KB less code
// (8104 -> 6936 instructions).  Measured on B200: N=20 +2.0 % at 32 768 problems, -0.2 % at 4096; N=10 -2.5 %; N=100 -10 %
// (more instructions executed; only the 12-CTA/SM N=20 kernel is instruction-fetch bound enough to profit).
#ifdef B2_WPP_ROLL_BOX
#define B2_BOX_LOOP _Pragma("unroll 1")
#else
#define B2_BOX_LOOP _Pragma("unroll")
#endif
// B2_WPP_NO_DT: the slack steps dt are not stored (one row per constraint) but recomputed by head() from dv, ds and the
// residual of the OLD iterate -- the very expression expand() evaluated: 12 % fewer node-private rows, 4 stores per obstacle
// and solve less.  Measured on B200: N=20 9.89 -> 9.58 ms per 4096 problems, 67.0 -> 64.7 ms per 32 768 (507 k solves/s);
// N=100 / M=20 284 -> 271 ms; identical iteration counts.  -DB2_WPP_NO_DT=0 restores the stored rows.
#ifndef B2_WPP_NO_DT
#define B2_WPP_NO_DT 1
#endif
#ifndef B2_WPP_MINB
#define B2_WPP_MINB 12          // resident CTAs per SM the N=20 layout is compiled for (register cap 65536/32/MINB = 170)
#endif
// Software prefetch of the node-private rows that live in global memory.  With one warp per problem a node-parallel pass is a
// chain of round trips: the loads of one box slot / obstacle are issued back to back, the warp sleeps until they return, the
// next obstacle follows.  At N = 100, M = 20 the private rows of the resident problems (592 x 0.68 MB) are 3x the L2, so every
// one of the ~600 round trips of an interior-point iteration goes to DRAM (measured: 1.6 M cycles per iteration = 600 x 2.7 k).
// `prefetch.global.L2` of the rows of obstacle j + PFD (and of the next node round) costs no registers and turns those
// round trips into L2 hits.
//   B2_WPP_PF_LONG : cache level for horizons > 32 nodes (0 off, 2 = L2, 1 = L1), B2_WPP_PF_SHORT: the same for N <= 32
//   B2_WPP_PFD     : prefetch distance in obstacles
#ifndef B2_WPP_PF_LONG
#define B2_WPP_PF_LONG 2
#endif
#ifndef B2_WPP_PF_SHORT
#define B2_WPP_PF_SHORT 0
#endif
#ifndef B2_WPP_PFD
#define B2_WPP_PFD 1
#endif
#ifndef B2_WPP_FEEDBACK_GAIN
#define B2_WPP_FEEDBACK_GAIN 0   // experiment: G = Lxu Luu^-1 formed once per factorisation, no triangular solve inside the sweeps
#endif
#ifndef B2_WPP_BOX_PIPE            // head(): next box slot's loads in flight while this one is processed
#ifdef B2_WPP_ROLL_BOX
#define B2_WPP_BOX_PIPE 0
#else
#define B2_WPP_BOX_PIPE B2_WPP_NO_DT
#endif
#endif
#ifndef B2_WPP_FUSE_LIN
#define B2_WPP_FUSE_LIN 1        // 0 never, 1 horizons > 32 nodes, 2 always: the linear-residual pass fused into the expansion pass
#endif
#ifndef B2_WPP_PF_NODE
#define B2_WPP_PF_NODE 0          // also prefetch the node-level rows / box slots of the next node round
#endif
template <int LEVEL>
__device__ __forceinline__ void b2_prefetch(const void* ptr)
{
#ifndef SIMT_EMU
    if (LEVEL == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    if (LEVEL == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

struct WOff {
    int sv, hg, rb, lrb, cv, lamu, lxu, lxx, lu;                               // shared-memory rows read across lanes
    int pf[PF_COUNT];                                                           // row of each private field in ITS space
    int total;                                                                  // shared-memory rows (incl. private fields placed there)
    int gtotal;                                                                 // global-memory rows
    __host__ __device__ static int pf_rows(int f, int M, bool sqp)
    {
        if (f == PF_DT && B2_WPP_NO_DT) return 0;
        if (f == PF_V || f == PF_DV) return 7;
        if (f == PF_PI || f == PF_DPI || f == PF_CPI || f == PF_PV) return 5;
        if (f >= PF_WPI) return !sqp ? 0 : f == PF_WPI ? 5 : 12 + 4 * M;
        return f <= PF_DT ? 12 + 4 * M : f == PF_B ? 5 : (f == PF_G || f == PF_RG) ? 7 : f == PF_XU ? 6 : M;
    }
    __host__ __device__ WOff(int M, unsigned long long smask, bool sqp)
    {
        int o = 0, g = 0;
        // rows that may live in either space keep their historical place in shared memory
#define B2_PLACE(f) do { if ((smask >> (f)) & 1u) { pf[f] = o; o += pf_rows(f, M, sqp); } else { pf[f] = g; g += pf_rows(f, M, sqp); } } while (0)
        sv = o; o += 14; hg = o; o += 8; rb = o; o += 5; lrb = o; o += 5;
        B2_PLACE(PF_DV); B2_PLACE(PF_DPI);
        cv = o; o += 7;
        B2_PLACE(PF_CPI);
        lamu = o; o += 3; lxu = o; o += 10; lxx = o; o += 15;
        B2_PLACE(PF_PV);
        lu = o; o += 2;
        B2_PLACE(PF_V); B2_PLACE(PF_PI);
#undef B2_PLACE
        for (int f = 0; f < PF_V; f++) {
            if ((smask >> f) & 1u) { pf[f] = o; o += pf_rows(f, M, sqp); }
            else { pf[f] = g; g += pf_rows(f, M, sqp); }
        }
        total = o;
        gtotal = g;
    }
};

__host__ __device__ inline int b2_wpp_stride(int N) { return (N + 1) | 1; }          // odd shared-memory row stride
__host__ __device__ inline int b2_wpp_gstride(int N) { return N + 1; }               // global rows packed (footprint over sector alignment)
#define B2_WPP_SCRATCH 80                                                             // flat doubles after the rows (U = [B A]^T S 7x5 at 0, U D at 40)
__host__ __device__ inline size_t b2_wpp_smem_bytes(int N, int M, unsigned long long smask, bool sqp = false)
{
    return sizeof(double) * ((size_t)WOff(M, smask, sqp).total * b2_wpp_stride(N) + B2_WPP_SCRATCH);
}
__host__ __device__ inline size_t b2_wpp_priv_doubles(int N, int M, unsigned long long smask, bool sqp = false)    // per resident CTA
{
    return (size_t)WOff(M, smask, sqp).gtotal * b2_wpp_gstride(N);
}

// y ~ 1/x (a few ulp) -> one Newton step: the result is within an ulp of the correctly rounded reciprocal
__device__ __forceinline__ double b2_rcp_polish(double x, double y) { return fma(y, fma(-x, y, 1.0), y); }

// Warp-wide maximum of NON-NEGATIVE doubles (norms; a NaN wins): the IEEE bit pattern of such values is monotone as an
// unsigned integer, so two REDUX.MAX.U32 (high word, then low word among the lanes holding the maximal high word)
// replace five rounds of double shuffles + compares.  Exact.
__device__ __forceinline__ double b2_warp_max_nonneg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
// Warp-wide maximum of NEGATIVE doubles (the step-length ratios in negative-reciprocal form): largest value = smallest
// magnitude = smallest bit pattern.
__device__ __forceinline__ double b2_warp_max_neg(double v)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double b2_warp_sum(double v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}
__device__ __forceinline__ double b2_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// expansion modes
enum { XP_SOLVE = 0, XP_REFINE = 1, XP_UPDATE = 2, XP_LIN = 3 };

template <int MT, int NT, unsigned long long SMASK, bool SQP = false>
struct WppSolver {
    const B2Params& p;
    const int b, lane, N, NS, GS, M, nct;
    double* const sm;
    double* const gm;
    const WOff o;
    const size_t Bm;
    // obstacle loops stay rolled: unrolling them (more loads in flight per L2 round trip) was measured 40-70 %
    // SLOWER -- the kernel is instruction-fetch sensitive (profiles/experiments/r01_hybrid_layout.md)
    static constexpr int JU = 1;
    int k;                  // node this lane is working on (node-parallel phases: lane, lane+32, ...)
    const int o_lb, o_lg, o_ub, o_ug, o_ls, o_us;
    int nc;
    double mu, nrm[4], sum_lt, alpha, S1, S2;

    __device__ WppSolver(const B2Params& p_, int b_, double* sm_, double* gm_)
        : p(p_), b(b_), lane(threadIdx.x & 31), N(NT > 0 ? NT : p_.N), NS(b2_wpp_stride(NT > 0 ? NT : p_.N)),
          GS(b2_wpp_gstride(NT > 0 ? NT : p_.N)), M(MT > 0 ? MT : p_.M),
          nct(12 + 4 * (MT > 0 ? MT : p_.M)), sm(sm_), gm(gm_), o(MT > 0 ? MT : p_.M, SMASK, SQP), Bm((size_t)p_.Bm),
          k(threadIdx.x & 31),
          o_lb(0), o_lg(6), o_ub(6 + (MT > 0 ? MT : p_.M)), o_ug(12 + (MT > 0 ? MT : p_.M)),
          o_ls(12 + 2 * (MT > 0 ? MT : p_.M)), o_us(12 + 3 * (MT > 0 ? MT : p_.M))
    {
        nc = 0;
        for (int kk = 0; kk <= N; kk++) nc += 2 * ((kk < N ? 2 : 0) + ((kk > 0 && kk < N) ? 4 : 0)) + 4 * M;
    }

    __device__ __forceinline__ double& S(int row, int kk) const { return sm[row * NS + kk]; }
    // row i of private field F at the node this lane is working on
    template <int F>
    __device__ __forceinline__ double& Q(int i) const
    {
        if ((SMASK >> F) & 1u) return sm[(o.pf[F] + i) * NS + k];
        return gm[(size_t)((o.pf[F] + i) * GS + k)];
    }
    // rows of v / pi at node kk (the owner's or a neighbour's)
    __device__ __forceinline__ double& VV(int i, int kk) const
    {
        if ((SMASK >> PF_V) & 1u) return sm[(o.pf[PF_V] + i) * NS + kk];
        return gm[(size_t)((o.pf[PF_V] + i) * GS + kk)];
    }
    __device__ __forceinline__ double& PP(int i, int kk) const
    {
        if ((SMASK >> PF_PI) & 1u) return sm[(o.pf[PF_PI] + i) * NS + kk];
        return gm[(size_t)((o.pf[PF_PI] + i) * GS + kk)];
    }
    // row i of placeable field F at node kk
    template <int F>
    __device__ __forceinline__ double& R(int i, int kk) const
    {
        if ((SMASK >> F) & 1u) return sm[(o.pf[F] + i) * NS + kk];
        return gm[(size_t)((o.pf[F] + i) * GS + kk)];
    }
    __device__ __forceinline__ bool box_on(int s, int kk) const { return s < 2 ? (kk < N) : (kk > 0 && kk < N); }
    __device__ __forceinline__ double scal(int kk) const { return kk < N ? p.dt : 1.0; }
    __device__ __forceinline__ void box_bounds(int s, double& lb, double& ub) const
    {
        const double xv = Q<PF_XU>(s);
        lb = p.lb6[s] - xv; ub = p.ub6[s] - xv;          // [lbu, lbx] / [ubu, ubx] in slot order
    }

    // the nine private values of box slot s at this lane's node (head() has its own seven): requested one slot ahead of their
    // use (B2_WPP_BOX_PIPE), in front of the `box_on` branch of the slot being processed
    struct Box9 { double ll, lu, tl, tu, xu, dl, du, til, tiu; };
    template <bool TINV_>
    __device__ __forceinline__ Box9 box_load9(const int s_) const
    {
        Box9 r;
        r.ll = Q<PF_LAM>(o_lb + s_); r.lu = Q<PF_LAM>(o_ub + s_); r.tl = Q<PF_T>(o_lb + s_); r.tu = Q<PF_T>(o_ub + s_);
        r.xu = Q<PF_XU>(s_); r.dl = Q<PF_DD>(o_lb + s_); r.du = Q<PF_DD>(o_ub + s_);
        r.til = TINV_ ? Q<PF_TINV>(o_lb + s_) : 0.0; r.tiu = TINV_ ? Q<PF_TINV>(o_ub + s_) : 0.0;
        return r;
    }

    // ---- software prefetch of private rows (see B2_WPP_PF_LONG above) ---------------------------------------------
    static constexpr int PFL = (NT > 32) ? B2_WPP_PF_LONG : (NT > 0 ? B2_WPP_PF_SHORT : 0);      // cache level, 0 = off
    static constexpr int PFD = B2_WPP_PFD;
    static constexpr bool FUSE_LIN = B2_WPP_FUSE_LIN == 2 || (B2_WPP_FUSE_LIN == 1 && NT > 32);       // see solve_any()
#define B2_FM(f) (1ull << (f))
    // fields READ per obstacle / per node by each pass (write-only rows need no fetch: full sectors are written)
    static constexpr unsigned long long FM_OBST_IN = B2_FM(PF_CX) | B2_FM(PF_CY) | B2_FM(PF_LG) | B2_FM(PF_UG) | B2_FM(PF_SLI) |
                                                     B2_FM(PF_SUI) | B2_FM(PF_SL) | B2_FM(PF_SU);
    static constexpr unsigned long long FM_HEAD = FM_OBST_IN | B2_FM(PF_LAM) | B2_FM(PF_T) | B2_FM(PF_DD) | B2_FM(PF_DSL) | B2_FM(PF_DSU) |
                                                  B2_FM(PF_ZLQ) | B2_FM(PF_ZZLQ) | B2_FM(PF_G) | B2_FM(PF_XU) | B2_FM(PF_V) | B2_FM(PF_PI) |
                                                  B2_FM(PF_DV) | B2_FM(PF_B);
    static constexpr unsigned long long FM_SOLVE = FM_OBST_IN | B2_FM(PF_LAM) | B2_FM(PF_T) | B2_FM(PF_TINV) | B2_FM(PF_DD) | B2_FM(PF_ZLQ) |
                                                   B2_FM(PF_ZZLQ) | B2_FM(PF_DLI) | B2_FM(PF_DUI) | B2_FM(PF_XU) | B2_FM(PF_V) | B2_FM(PF_DV);
    static constexpr unsigned long long FM_PREP = FM_SOLVE | B2_FM(PF_RG);
    static constexpr unsigned long long FM_PREP3 = B2_FM(PF_LAM) | B2_FM(PF_TINV) | B2_FM(PF_LRGSL) | B2_FM(PF_LRGSU) | B2_FM(PF_DLI) |
                                                   B2_FM(PF_DUI) | B2_FM(PF_CX) | B2_FM(PF_CY);
    static constexpr unsigned long long FM_REFINE = FM_OBST_IN | B2_FM(PF_LAM) | B2_FM(PF_T) | B2_FM(PF_TINV) | B2_FM(PF_DD) | B2_FM(PF_LRGSL) |
                                                    B2_FM(PF_LRGSU) | B2_FM(PF_DLI) | B2_FM(PF_DUI) | B2_FM(PF_DSL) | B2_FM(PF_DSU) | B2_FM(PF_XU) |
                                                    B2_FM(PF_V) | B2_FM(PF_DV) | B2_FM(PF_DPI) | B2_FM(PF_CPI);
    static constexpr unsigned long long FM_LIN = FM_OBST_IN | B2_FM(PF_LAM) | B2_FM(PF_T) | B2_FM(PF_DD) | B2_FM(PF_DSL) | B2_FM(PF_DSU) |
                                                 B2_FM(PF_ZLQ) | B2_FM(PF_ZZLQ) | B2_FM(PF_XU) | B2_FM(PF_DV) | B2_FM(PF_RG) | B2_FM(PF_DPI);
    template <int F>
    __device__ __forceinline__ void pfq(int i, int kk) const
    {
        if (!((SMASK >> F) & 1u)) b2_prefetch<PFL>(gm + (size_t)((o.pf[F] + i) * GS + kk));
    }
    template <int F, unsigned long long FM>
    __device__ __forceinline__ void pf_con4(int j, int kk) const       // the four constraint rows of obstacle j
    {
        if ((FM >> F) & 1u) { pfq<F>(o_lg + j, kk); pfq<F>(o_ug + j, kk); pfq<F>(o_ls + j, kk); pfq<F>(o_us + j, kk); }
    }
    template <int F, unsigned long long FM>
    __device__ __forceinline__ void pf_one(int j, int kk) const
    {
        if ((FM >> F) & 1u) pfq<F>(j, kk);
    }
    template <int F, unsigned long long FM>
    __device__ __forceinline__ void pf_rows(int i0, int n, int kk) const
    {
        if ((FM >> F) & 1u) {
#pragma unroll
            for (int i = 0; i < n; i++) pfq<F>(i0 + i, kk);
        }
    }
    // rows of obstacle j at node kk that pass FM reads
    template <unsigned long long FM>
    __device__ __forceinline__ void prefetch_obst(int j, int kk) const
    {
        if (PFL == 0 || kk > N) return;
        pf_con4<PF_LAM, FM>(j, kk); pf_con4<PF_T, FM>(j, kk); pf_con4<PF_TINV, FM>(j, kk); pf_con4<PF_DD, FM>(j, kk);
        pf_one<PF_CX, FM>(j, kk); pf_one<PF_CY, FM>(j, kk); pf_one<PF_LG, FM>(j, kk); pf_one<PF_UG, FM>(j, kk);
        pf_one<PF_ZLQ, FM>(j, kk); pf_one<PF_ZZLQ, FM>(j, kk); pf_one<PF_SLI, FM>(j, kk); pf_one<PF_SUI, FM>(j, kk);
        pf_one<PF_SL, FM>(j, kk); pf_one<PF_SU, FM>(j, kk); pf_one<PF_LRGSL, FM>(j, kk); pf_one<PF_LRGSU, FM>(j, kk);
        pf_one<PF_DSL, FM>(j, kk); pf_one<PF_DSU, FM>(j, kk); pf_one<PF_DLI, FM>(j, kk); pf_one<PF_DUI, FM>(j, kk);
    }
    // node-level rows and the six box slots of node kk that pass FM reads, plus its first PFD obstacles
    template <unsigned long long FM>
    __device__ __forceinline__ void prefetch_node(int kk) const
    {
        if (PFL == 0 || kk > N) return;
        pf_rows<PF_LAM, FM>(o_lb, 6, kk); pf_rows<PF_LAM, FM>(o_ub, 6, kk);
        pf_rows<PF_T, FM>(o_lb, 6, kk); pf_rows<PF_T, FM>(o_ub, 6, kk);
        pf_rows<PF_TINV, FM>(o_lb, 6, kk); pf_rows<PF_TINV, FM>(o_ub, 6, kk);
        pf_rows<PF_DD, FM>(o_lb, 6, kk); pf_rows<PF_DD, FM>(o_ub, 6, kk);
        pf_rows<PF_B, FM>(0, 5, kk); pf_rows<PF_G, FM>(0, 7, kk); pf_rows<PF_RG, FM>(0, 7, kk); pf_rows<PF_XU, FM>(0, 6, kk);
        pf_rows<PF_V, FM>(0, 7, kk); pf_rows<PF_PI, FM>(0, 5, kk); pf_rows<PF_DV, FM>(0, 7, kk); pf_rows<PF_DPI, FM>(0, 5, kk);
        pf_rows<PF_CPI, FM>(0, 5, kk);
        for (int j = 0; j < PFD && j < M; j++) prefetch_obst<FM>(j, kk);
    }
    // inside the obstacle loop of node k, before obstacle j is processed: obstacle j + PFD of this node, or -- past the last
    // obstacle -- nothing (the first PFD obstacles of the next node round were prefetched with its node-level rows at j = 0)
    template <unsigned long long FM>
    __device__ __forceinline__ void prefetch_ahead(int j) const
    {
        if (PFL == 0) return;
        if (B2_WPP_PF_NODE && j == 0) prefetch_node<FM>(k + 32);
        if (j + PFD < M) prefetch_obst<FM>(j + PFD, k);
    }

    // ---- preparation + cold-start initialisation: lane k handles node k ----------------------------
    __device__ void prepare_and_init()
    {
        for (k = lane; k <= N; k += 32) {
            double xk[5], uk[2] = {0.0, 0.0};
#pragma unroll
            for (int i = 0; i < 5; i++) xk[i] = p.X[((size_t)k * 5 + i) * Bm + b];
            const double sc = scal(k);
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
            double v[7] = {0, 0, 0, 0, 0, 0, 0};
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) v[2 + i] = p.X0[(size_t)i * Bm + b] - xk[i];
            }
            for (int i = 0; i < nct; i++) { Q<PF_LAM>(i) = 0.0; Q<PF_T>(i) = 0.0; Q<PF_TINV>(i) = 0.0; Q<PF_DD>(i) = 0.0; if (!B2_WPP_NO_DT) Q<PF_DT>(i) = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) R<PF_DV>(i, k) = 0.0;
#pragma unroll
            for (int i = 0; i < 5; i++) R<PF_DPI>(i, k) = 0.0;
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
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                double hval, cx, cy;
                model_con_h(xk, p.P[((size_t)k * 2 * M + 2 * j) * Bm + b], p.P[((size_t)k * 2 * M + 2 * j + 1) * Bm + b], p.r2, &hval, &cx, &cy);
                const double s_l = p.SL[((size_t)k * M + j) * Bm + b], s_u = p.SU[((size_t)k * M + j) * Bm + b];
                const double lg = p.lh - hval - s_l, ug = p.uh - hval + s_u;
                Q<PF_CX>(j) = cx; Q<PF_CY>(j) = cy; Q<PF_LG>(j) = lg; Q<PF_UG>(j) = ug;
                Q<PF_SLI>(j) = s_l; Q<PF_SUI>(j) = s_u;
                const double zl_in = p.ZL[((size_t)k * M + j) * Bm + b], Zl_in = p.ZZL[((size_t)k * M + j) * Bm + b];
                Q<PF_ZLQ>(j) = sc * (zl_in + Zl_in * s_l);
                Q<PF_ZZLQ>(j) = sc * Zl_in;
                const double cv = cx * v[2] + cy * v[3];
                double tl = cv - lg, tu = ug - cv;
                tl = p.thr0 > tl ? p.thr0 : tl;
                tu = p.thr0 > tu ? p.thr0 : tu;
                Q<PF_T>(o_lg + j) = tl; Q<PF_T>(o_ug + j) = tu;
                Q<PF_LAM>(o_lg + j) = p.mu0 / tl; Q<PF_LAM>(o_ug + j) = p.mu0 / tu;
                Q<PF_T>(o_ls + j) = 1.0; Q<PF_T>(o_us + j) = 1.0;
                Q<PF_LAM>(o_ls + j) = p.mu0; Q<PF_LAM>(o_us + j) = p.mu0;
                Q<PF_SL>(j) = 0.0; Q<PF_SU>(j) = 0.0; Q<PF_DSL>(j) = 0.0; Q<PF_DSU>(j) = 0.0;
            }
#pragma unroll
            for (int i = 0; i < 7; i++) VV(i, k) = v[i];
#pragma unroll
            for (int i = 0; i < 5; i++) PP(i, k) = 0.0;
        }
        __syncwarp();
    }

    // Element of a per-node register array (v order [u_a,u_alpha,x,y,psi,v,omega]) that belongs to box slot s (the slots
    // skip psi).  In the ROLLED box loops s is a run-time value: a select chain / predicated adds instead of dynamic
    // indexing (which would push the arrays to local memory); with a compile-time s both fold to a[iv].
    __device__ __forceinline__ static double slot_get(const double* a, int s)
    {
        double r = a[0];
        if (s == 1) r = a[1];
        if (s == 2) r = a[2];
        if (s == 3) r = a[3];
        if (s == 4) r = a[5];
        if (s == 5) r = a[6];
        return r;
    }
    __device__ __forceinline__ static void slot_add(double* a, int s, double x)
    {
        if (s == 0) a[0] += x;
        if (s == 1) a[1] += x;
        if (s == 2) a[2] += x;
        if (s == 3) a[3] += x;
        if (s == 4) a[5] += x;
        if (s == 5) a[6] += x;
    }

    // g_i(v,s) of the box constraints (rd_i = g_i - t_i)
    __device__ __forceinline__ double g_box_l(int s, const double* v) const { double lb, ub; box_bounds(s, lb, ub); return v[b2_slot2v(s)] - lb; }
    __device__ __forceinline__ double g_box_u(int s, const double* v) const { double lb, ub; box_bounds(s, lb, ub); return ub - v[b2_slot2v(s)]; }

    // ---- head of an interior-point iteration: ONE pass over the node-private rows ---------------------
    //  (1) apply the previous step with length a (a = 0 on entry: the step rows are zeroed by prepare_and_init),
    //  (2) KKT residuals, their norms and mu at the new iterate,
    //  (3) barrier/slack elimination: 1/t, D^-1, condensed node Hessian (-> rows cv[0..6], cpi[0] for the recursion),
    //  (4) condensed gradient of the AFFINE right-hand side (-> rows hg[0..6], what solve_prep(0) would produce).
    // Fused because every one of these walks the same multipliers/slacks: with the private rows in L2 a pass costs a
    // round trip per obstacle, and each pass is straight-line code that has to be fetched.  (3)+(4) are wasted on the
    // last (converged) iteration only.
    __device__ void head(double a)
    {
        for (k = lane; k <= N; k += 32) {
#pragma unroll
            for (int i = 0; i < 7; i++) {
                const double vo = VV(i, k);
                if (B2_WPP_NO_DT) S(o.cv + i, k) = vo;          // the old iterate, for the recomputation of dt below (cv is free here)
                VV(i, k) = fma(a, R<PF_DV>(i, k), vo);
            }
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 5; i++) PP(i, k) = fma(a, R<PF_DPI>(i, k), PP(i, k));
            }
        }
        __syncwarp();
        double sum = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
        for (k = lane; k <= N; k += 32) {
            double v[7], rg[7], Hx[7], gt[7], hxy = 0.0;
            const double* Hd = k < N ? p.Hd : p.Hde;
            const double sc = scal(k);
#if B2_WPP_BOX_PIPE
            // the private rows of box slot s+1 are requested before slot s is processed (its seven loads sit in front of the
            // `box_on` branch of slot s): two slots share an L2 round trip instead of paying one each
            struct BoxIn { double dl, ll, du, lu, xu, tl, tu; };
            auto box_load = [&](const int s_) {
                BoxIn r;
                r.dl = Q<PF_DD>(o_lb + s_); r.ll = Q<PF_LAM>(o_lb + s_); r.du = Q<PF_DD>(o_ub + s_); r.lu = Q<PF_LAM>(o_ub + s_);
                r.xu = Q<PF_XU>(s_); r.tl = Q<PF_T>(o_lb + s_); r.tu = Q<PF_T>(o_ub + s_);
                return r;
            };
            BoxIn bnext = box_load(0);                 // (with the node-level rows below: one round trip)
#endif

#pragma unroll
            for (int i = 0; i < 7; i++) { v[i] = VV(i, k); rg[i] = fma(Hd[i], v[i], Q<PF_G>(i)); Hx[i] = Hd[i] + p.reg_prim; gt[i] = 0.0; }
            double sv[10];
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
            B2_BOX_LOOP
            for (int s = 0; s < 6; s++) {
#if B2_WPP_BOX_PIPE
                const BoxIn bin = bnext;
                if (s + 1 < 6) bnext = box_load(s + 1);
#endif
                if (!box_on(s, k)) continue;
                const int il = o_lb + s, iu = o_ub + s;
#if B2_WPP_BOX_PIPE
                double ll = fma(a, bin.dl, bin.ll), lu_ = fma(a, bin.du, bin.lu);
                const double lb = p.lb6[s] - bin.xu, ub = p.ub6[s] - bin.xu;
#else
                double ll = fma(a, Q<PF_DD>(il), Q<PF_LAM>(il)), lu_ = fma(a, Q<PF_DD>(iu), Q<PF_LAM>(iu));
                double lb, ub;
                box_bounds(s, lb, ub);
#endif
                double tl, tu;
                if (B2_WPP_NO_DT) {
#if B2_WPP_BOX_PIPE
                    const double tlo = bin.tl, tuo = bin.tu;
#else
                    const double tlo = Q<PF_T>(il), tuo = Q<PF_T>(iu);
#endif
                    const double vos = S(o.cv + b2_slot2v(s), k), dvs = R<PF_DV>(b2_slot2v(s), k);
                    const double dtl = dvs + ((vos - lb) - tlo), dtu = -dvs + ((ub - vos) - tuo);      // = expand()'s dtl, dtu
                    tl = fma(a, dtl, tlo); tu = fma(a, dtu, tuo);
                } else {
                    tl = fma(a, Q<PF_DT>(il), Q<PF_T>(il)); tu = fma(a, Q<PF_DT>(iu), Q<PF_T>(iu));
                }
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
                const double rlu = 1.0 / (tl * tu);                    // one division per pair: 1/tl = tu/(tl tu)
                const double til = b2_rcp_polish(tl, rlu * tu), tiu = b2_rcp_polish(tu, rlu * tl);
                Q<PF_TINV>(il) = til; Q<PF_TINV>(iu) = tiu;
                slot_add(Hx, s, fma(lu_, tiu, ll * til));
                slot_add(gt, s, fma(-fma(lu_, rdu, mu_), tiu, fma(ll, rdl, ml) * til));
            }
            const double Zuq = sc * p.Zu;
            double vo2 = 0.0, vo3 = 0.0, dv2 = 0.0, dv3 = 0.0;
            if (B2_WPP_NO_DT) { vo2 = S(o.cv + 2, k); vo3 = S(o.cv + 3, k); dv2 = R<PF_DV>(2, k); dv3 = R<PF_DV>(3, k); }
            // everything that the obstacles do not touch is finished here and leaves the registers before their loop: the
            // dynamics residual, and every component of the stationarity residual / condensed Hessian / affine right-hand side
            // but the two position components
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
            if (k == 0) { rg[4] = rg[5] = rg[6] = 0.0; }
            if (k == N) { rg[0] = rg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) {
                if (i == 2 || i == 3) continue;
                Q<PF_RG>(i) = rg[i]; n0 = fmax(n0, fabs(rg[i]));
                S(o.hg + i, k) = rg[i] + gt[i];
                S(o.cv + i, k) = Hx[i];
            }
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                prefetch_ahead<FM_HEAD>(j);
                const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
                double l0 = fma(a, Q<PF_DD>(i0), Q<PF_LAM>(i0)), l1 = fma(a, Q<PF_DD>(i1), Q<PF_LAM>(i1));
                double l2 = fma(a, Q<PF_DD>(i2), Q<PF_LAM>(i2)), l3 = fma(a, Q<PF_DD>(i3), Q<PF_LAM>(i3));
                double t0, t1, t2, t3;
                if (B2_WPP_NO_DT) {
                    // dt of the stored step, recomputed the way expand() computed it (old v, old slacks, old t)
                    const double cxo = Q<PF_CX>(j), cyo = Q<PF_CY>(j);
                    const double cdv = fma(cyo, dv3, cxo * dv2);
                    const double cvalo = fma(cyo, vo3, cxo * vo2);
                    const double t0o = Q<PF_T>(i0), t1o = Q<PF_T>(i1), t2o = Q<PF_T>(i2), t3o = Q<PF_T>(i3);
                    const double slo = Q<PF_SL>(j), suo = Q<PF_SU>(j), dslo = Q<PF_DSL>(j), dsuo = Q<PF_DSU>(j);
                    const double rd0o = cvalo + slo - Q<PF_LG>(j) - t0o;
                    const double rd1o = Q<PF_UG>(j) - cvalo + suo - t1o;
                    const double rd2o = slo + Q<PF_SLI>(j) - t2o;
                    const double rd3o = suo + Q<PF_SUI>(j) - t3o;
                    const double dt0 = cdv + dslo + rd0o, dt1 = -cdv + dsuo + rd1o, dt2 = dslo + rd2o, dt3 = dsuo + rd3o;
                    t0 = fma(a, dt0, t0o); t1 = fma(a, dt1, t1o); t2 = fma(a, dt2, t2o); t3 = fma(a, dt3, t3o);
                } else {
                    t0 = fma(a, Q<PF_DT>(i0), Q<PF_T>(i0)); t1 = fma(a, Q<PF_DT>(i1), Q<PF_T>(i1));
                    t2 = fma(a, Q<PF_DT>(i2), Q<PF_T>(i2)); t3 = fma(a, Q<PF_DT>(i3), Q<PF_T>(i3));
                }
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
                const double s_l = fma(a, Q<PF_DSL>(j), Q<PF_SL>(j)), s_u = fma(a, Q<PF_DSU>(j), Q<PF_SU>(j));
                Q<PF_SL>(j) = s_l; Q<PF_SU>(j) = s_u;
                const double cx = Q<PF_CX>(j), cy = Q<PF_CY>(j);
                const double cv = fma(cy, v[3], cx * v[2]);
                const double dl = l1 - l0;
                rg[2] = fma(cx, dl, rg[2]); rg[3] = fma(cy, dl, rg[3]);
                const double sli = Q<PF_SLI>(j), sui = Q<PF_SUI>(j);
                const double Zlq = Q<PF_ZZLQ>(j);
                const double rgsl = fma(Zlq, s_l, Q<PF_ZLQ>(j)) - l0 - l2;
                const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, sui, p.zu)) - l1 - l3;
                n0 = fmax(n0, fmax(fabs(rgsl), fabs(rgsu)));
                const double rd0 = cv + s_l - Q<PF_LG>(j) - t0;
                const double rd1 = Q<PF_UG>(j) - cv + s_u - t1;
                const double rd2 = s_l + sli - t2;
                const double rd3 = s_u + sui - t3;
                n2 = fmax(n2, fmax(fmax(fabs(rd0), fabs(rd1)), fmax(fabs(rd2), fabs(rd3))));
                const double m0 = l0 * t0, m1 = l1 * t1, m2 = l2 * t2, m3 = l3 * t3;
                n3 = fmax(n3, fmax(fmax(fabs(m0), fabs(m1)), fmax(fabs(m2), fabs(m3))));
                sum += m0; sum += m1; sum += m2; sum += m3;
                // one division for the four slacks of the obstacle (t in [1e-16, ~1e5]: the product stays in range)
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
                Hx[2] = fma(wcx, cx, Hx[2]); Hx[3] = fma(weff * cy, cy, Hx[3]); hxy = fma(wcx, cy, hxy);
                // affine right-hand side (kind 0: rm = lam t)
                const double q0 = fma(l0, rd0, m0) * ti0, q1 = fma(l1, rd1, m1) * ti1;
                const double q2 = fma(l2, rd2, m2) * ti2, q3 = fma(l3, rd3, m3) * ti3;
                const double rhol = rgsl + q0 + q2, rhou = rgsu + q1 + q3;
                const double coef = fma(-(Gl * rhol), Dli, q0) - fma(-(Gu * rhou), Dui, q1);
                gt[2] = fma(cx, coef, gt[2]); gt[3] = fma(cy, coef, gt[3]);
            }
            if (k == 0) { rg[2] = rg[3] = 0.0; }
#pragma unroll
            for (int i = 2; i < 4; i++) {
                Q<PF_RG>(i) = rg[i]; n0 = fmax(n0, fabs(rg[i]));
                S(o.hg + i, k) = rg[i] + gt[i];
                S(o.cv + i, k) = Hx[i];
            }
            S(o.hg + 7, k) = hxy;
        }
        sum = b2_warp_sum(sum);
        nrm[0] = b2_warp_max_nonneg(n0); nrm[1] = b2_warp_max_nonneg(n1); nrm[2] = b2_warp_max_nonneg(n2); nrm[3] = b2_warp_max_nonneg(n3);
        sum_lt = sum;
        mu = sum / nc;
        __syncwarp();
    }

    // Backward Riccati recursion in SQUARE-ROOT form (what HPIPM does with square_root_alg = 1):
    //   AL = [B A]^T Lxx+ ,  M = Htilde + AL AL^T ,  L = chol(M) ,  pass Lxx on.
    // The cost-to-go Hessian P = Lxx Lxx^T is never formed and never obtained by subtraction: near
    // convergence the barrier weights on active state bounds reach 1e15+, where the classical recursion
    // P = Mxx - Mxu Muu^-1 Mux loses all digits (and definiteness) -- see DESIGN.md "square-root Riccati".
    // Lane e<28 owns entry (mi,mj) of the lower triangle; the 7 Cholesky columns are sequential.
    __device__ void riccati_factorize()
    {
        const int e = lane < 28 ? lane : 0;
        int mi = 0;
        while ((mi + 1) * (mi + 2) / 2 <= e) mi++;
        const int mj = e - mi * (mi + 1) / 2;
        // phase A roles: lane < 30 owns entry (ai, aq) = (lane / 5, lane % 5) of U, lanes 0..4 additionally entry (6, lane)
        const int ai0 = lane < 30 ? lane / 5 : 0, aq0 = lane % 5;
        int ba0[5], ba1[5];
#pragma unroll
        for (int r = 0; r < 5; r++) { ba0[r] = o.sv + B2_BATAB[r][ai0]; ba1[r] = o.sv + B2_BATAB[r][6]; }
        const bool xx = lane < 28 && mi >= 2 && mj >= 2;
        const bool xu = lane < 28 && mi >= 2 && mj < 2;
        const int xrow = xx ? o.lxx + b2_sym(mi - 2, mj - 2) : 0;
        const int lrow = xu ? o.lxu + 2 * (mi - 2) + mj : 0;
        // the diagonal lane of column c keeps sigma_c p_c (the square of the column's normalisation) in a row of its own: the
        // two control columns in the rows that will hold 1/L00 and 1/L11, the state columns in the rows of the refinement
        // right-hand side (lrb), which are dead until the linear residual is assembled
        const bool mine = xx || xu || lane == 1;                        // S10 (lane 1), Sxu, Sxx
        const int myrow = xx ? xrow : xu ? lrow : o.lamu + 1;
        const bool dg = lane < 28 && mi == mj;
        const int sprow = mj == 0 ? o.lamu + 0 : mj == 1 ? o.lamu + 2 : o.lrb + (mj - 2);
        double* const scr = sm + o.total * NS;
        for (int kk = N; kk >= 0; kk--) {
            double m = 0.0;
            if (mi == mj) m = S(o.cv + mi, kk);                 // condensed node Hessian, left there by head()
            if (mi == 3 && mj == 2) m = S(o.hg + 7, kk);
            if (kk < N) {
                // A: U = [B A]^T Sxx(kk+1) with the UNNORMALISED factor Sxx = Lxx D^-1 of the next node, and U D^2 with
                // D_q^2 = 1 / (sigma_q p_q): AL AL^T = U D^2 U^T needs one division per column and no square root, so the
                // normalisation of the factor (square roots, 28 products, stores: normalize_factor()) leaves the sequential chain
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
                // B: M = Htilde + (U D^2) U^T
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 5; q++) acc = fma(scr[40 + mi * 5 + q], scr[mj * 5 + q], acc);
                m += acc;
            }
            // C: right-looking Cholesky with DELAYED NORMALISATION.  The sequential chain per column is one
            // shuffle round + two FMAs: instead of T_ik -= T_ic T_kc / T_cc (rsqrt/division on the critical
            // path) the scaled Schur complement S = sigma * T is updated fraction-free,
            //     S_ik <- (S_ik p - S_ic S_kc) * 2^-e(p),   sigma <- sigma * p 2^-e(p),   p = S_cc,
            // (the power of two keeps sigma in [1, 2^7)); column c of the factor is L_ic = S_ic / sqrt(sigma_c p_c), formed
            // by normalize_factor() after the recursion.  Non-positive pivot -> zero column, as BLASFEO.
            double sigma = 1.0, my_sp = 0.0;
            // rolled: the unrolled 7 columns made the loop body 5.9 KB, the size of an L0 instruction cache that three
            // warps share (measured +0.6 % at 4096 problems, +2.2 % at 32 768; same arithmetic).  End of round 2, six columns,
            // -DB2_WPP_UNROLL_CHOL: 8150 -> 8520 instructions, 9.10 -> 9.29 ms per 4096 problems, 61.3 -> 63.3 ms per 32 768)
#ifdef B2_WPP_UNROLL_CHOL
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int c = 0; c < 6; c++) {
                const double pv_ = b2_bcast(m, c * (c + 1) / 2 + c);
                const double la = b2_bcast(m, mi * (mi + 1) / 2 + (c <= mi ? c : mi));
                const double lb = b2_bcast(m, mj * (mj + 1) / 2 + (c <= mj ? c : mj));
                const bool ok = pv_ > 0.0;
                if (mj == c) my_sp = ok ? sigma * pv_ : 0.0;
#ifdef B2_CHOL_BRANCH
                if (ok) {
                    const double scale = __hiloint2double(0x7fe00000 - (__double2hiint(pv_) & 0x7ff00000), 0);
                    if (mj > c) m = fma(m, pv_, -(la * lb)) * scale;
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
                    m = (ok && mj > c) ? mn : m;
                    sigma = ok ? sn : sigma;
                }
#endif
            }
            // (the last column has nothing below or to the right of its pivot: its lane needs sigma and its own entry, no exchange)
            if (mj == 6) my_sp = m > 0.0 ? sigma * m : 0.0;
            // unnormalised column entries and the column scales; normalize_factor() turns them into L, 1/L00, 1/L11
            // (one store per lane under a loop-invariant predicate; the rows of node 0 / the terminal node that nobody reads are
            //  written too, which keeps lane- and node-dependent branches out of the recursion)
            if (mine) S(myrow, kk) = m;
            if (dg) S(sprow, kk) = my_sp;
            __syncwarp();
        }
        normalize_factor();
    }

    // node-parallel end of the factorisation: L_ic = S_ic / sqrt(sigma_c p_c) for the stored columns (Lxu, Lxx; l10),
    //   1/L00 = 1 / sqrt(p_0),  1/L11 = sqrt(sigma_1 / p_1) = sigma_1 / sqrt(sigma_1 p_1),  sigma_1 = p_0 2^-e(p_0).
    // 1/sqrt through the two correctly rounded operations (rsqrt() is not: the CPU mirror could not reproduce it).
    __device__ void normalize_factor()
    {
        for (k = lane; k <= N; k += 32) {
            double sig1 = 1.0;
            // (rolled: one copy of the square root / division sequences in the instruction stream)
#pragma unroll 1
            for (int c = 0; c < 7; c++) {
                const int row = c == 0 ? o.lamu : c == 1 ? o.lamu + 2 : o.lrb + c - 2;
                const double sp = S(row, k);
                const double rs = sp > 0.0 ? 1.0 / sqrt(sp) : 0.0;
                if (c == 0) {
                    sig1 = sp > 0.0 ? sp * __hiloint2double(0x7fe00000 - (__double2hiint(sp) & 0x7ff00000), 0) : 1.0;
                    if (k < N) { S(o.lamu + 0, k) = rs; S(o.lamu + 1, k) *= rs; }
                }
                if (c == 1 && k < N) S(o.lamu + 2, k) = sig1 * rs;
                if (k > 0) {
                    if (c < 2) {
#pragma unroll
                        for (int i = 0; i < 5; i++) S(o.lxu + 2 * i + c, k) *= rs;
                    } else {
#pragma unroll 1
                        for (int i = c - 2; i < 5; i++) S(o.lxx + i * (i + 1) / 2 + (c - 2), k) *= rs;
                    }
                }
            }
#if B2_WPP_FEEDBACK_GAIN
            // feedback gain G_k = Lxu Luu^-1 (5 x 2), stored over Lxu (nothing else reads it): the KKT solves then need no
            // triangular solve with Luu inside their sequential sweeps -- p_k = gx + A^T s - G (gu + B^T s) going backward,
            // du_k = du0_k - G^T dx_k going forward, du0_k = -(Luu Luu^T)^-1 w_k formed node-parallel in between
            if (k > 0 && k < N) {
                const double l00i = S(o.lamu + 0, k), l10 = S(o.lamu + 1, k), l11i = S(o.lamu + 2, k);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const double g1 = S(o.lxu + 2 * i + 1, k) * l11i;
                    const double g0 = fma(-g1, l10, S(o.lxu + 2 * i, k)) * l00i;
                    S(o.lxu + 2 * i, k) = g0; S(o.lxu + 2 * i + 1, k) = g1;
                }
            }
#endif
        }
        __syncwarp();
    }

    // complementarity right-hand side of a base solve: kind 0 (affine) lam t; 1 (corrector) lam t + dt_aff dlam_aff - sigma mu;
    // 2 (centering) lam t - sigma mu.  Branch-free in the run-time kind: rm_cdd = [kind == 1], rm_smu = kind ? sigma mu : 0
    // (set by solve_any), evaluated as fma(lam, t, rm_cdd * dd) - rm_smu.
    double rm_cdd, rm_smu;
    __device__ __forceinline__ double rm_of(int, double l, double t, double dd, double) const
    {
        return fma(l, t, rm_cdd * dd) - rm_smu;
    }

    // q_lg, q_ug and the slack-elimination right-hand sides rho_l, rho_u of obstacle j at node k, base kinds
    __device__ __forceinline__ void slack_rho(int kind, double smu, int j, const double* v, double Zuq, double sc,
                                              double& rhol, double& rhou, double& q0, double& q1) const
    {
        const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
        const double l0 = Q<PF_LAM>(i0), l1 = Q<PF_LAM>(i1), l2 = Q<PF_LAM>(i2), l3 = Q<PF_LAM>(i3);
        const double t0 = Q<PF_T>(i0), t1 = Q<PF_T>(i1), t2 = Q<PF_T>(i2), t3 = Q<PF_T>(i3);
        const double s_l = Q<PF_SL>(j), s_u = Q<PF_SU>(j);
        const double cv = fma(Q<PF_CY>(j), v[3], Q<PF_CX>(j) * v[2]);
        const double rd0 = cv + s_l - Q<PF_LG>(j) - t0;
        const double rd1 = Q<PF_UG>(j) - cv + s_u - t1;
        const double rd2 = s_l + Q<PF_SLI>(j) - t2;
        const double rd3 = s_u + Q<PF_SUI>(j) - t3;
        q0 = fma(l0, rd0, rm_of(kind, l0, t0, Q<PF_DD>(i0), smu)) * Q<PF_TINV>(i0);
        q1 = fma(l1, rd1, rm_of(kind, l1, t1, Q<PF_DD>(i1), smu)) * Q<PF_TINV>(i1);
        const double q2 = fma(l2, rd2, rm_of(kind, l2, t2, Q<PF_DD>(i2), smu)) * Q<PF_TINV>(i2);
        const double q3 = fma(l3, rd3, rm_of(kind, l3, t3, Q<PF_DD>(i3), smu)) * Q<PF_TINV>(i3);
        const double rgsl = fma(Q<PF_ZZLQ>(j), s_l, Q<PF_ZLQ>(j)) - l0 - l2;
        const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, Q<PF_SUI>(j), p.zu)) - l1 - l3;
        rhol = rgsl + q0 + q2;
        rhou = rgsu + q1 + q3;
    }

    // node-local part of a solve: condensed gradient of node k for the right-hand side `kind`
    // (0 affine, 1 corrector, 2 centering, 3 refinement = linear residual) -> rows hg[0..6]
    __device__ void solve_prep(int kind, double smu)
    {
        for (k = lane; k <= N; k += 32) {
            const double sc = scal(k);
            double gt[7];
            if (kind == 3) {
                gt[2] = S(o.hg + 2, k); gt[3] = S(o.hg + 3, k);                // the linear residual lrg is in hg: the obstacles add to two rows
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                    prefetch_ahead<FM_PREP3>(j);
                    const double G0 = Q<PF_LAM>(o_lg + j) * Q<PF_TINV>(o_lg + j), G1 = Q<PF_LAM>(o_ug + j) * Q<PF_TINV>(o_ug + j);
                    const double coef = fma(G1 * Q<PF_LRGSU>(j), Q<PF_DUI>(j), -((G0 * Q<PF_LRGSL>(j)) * Q<PF_DLI>(j)));
                    gt[2] = fma(Q<PF_CX>(j), coef, gt[2]);
                    gt[3] = fma(Q<PF_CY>(j), coef, gt[3]);
                }
            } else {
                double v[7];
#if B2_WPP_BOX_PIPE
                Box9 bnext = box_load9<true>(0);
#endif
#pragma unroll
                for (int i = 0; i < 7; i++) { v[i] = VV(i, k); gt[i] = Q<PF_RG>(i); }
                B2_BOX_LOOP
                for (int s = 0; s < 6; s++) {
#if B2_WPP_BOX_PIPE
                    const Box9 bin = bnext;
                    if (s + 1 < 6) bnext = box_load9<true>(s + 1);
                    if (!box_on(s, k)) continue;
                    const double ll = bin.ll, lu_ = bin.lu, tl = bin.tl, tu = bin.tu;
                    const double lb = p.lb6[s] - bin.xu, ub = p.ub6[s] - bin.xu;
                    const double vv = slot_get(v, s);
                    const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
                    const double ql = fma(ll, rdl, rm_of(kind, ll, tl, bin.dl, smu)) * bin.til;
                    const double qu = fma(lu_, rdu, rm_of(kind, lu_, tu, bin.du, smu)) * bin.tiu;
#else
                    if (!box_on(s, k)) continue;
                    const int il = o_lb + s, iu = o_ub + s;
                    const double ll = Q<PF_LAM>(il), lu_ = Q<PF_LAM>(iu), tl = Q<PF_T>(il), tu = Q<PF_T>(iu);
                    double lb, ub;
                    box_bounds(s, lb, ub);
                    const double vv = slot_get(v, s);
                    const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
                    const double ql = fma(ll, rdl, rm_of(kind, ll, tl, Q<PF_DD>(il), smu)) * Q<PF_TINV>(il);
                    const double qu = fma(lu_, rdu, rm_of(kind, lu_, tu, Q<PF_DD>(iu), smu)) * Q<PF_TINV>(iu);
#endif
                    slot_add(gt, s, ql - qu);
                }
                // (the components that the obstacles do not touch leave the registers before their loop)
#pragma unroll
                for (int i = 0; i < 7; i++) {
                    if (i != 2 && i != 3) S(o.hg + i, k) = gt[i];
                }
                const double Zuq = sc * p.Zu;
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                    prefetch_ahead<FM_PREP>(j);
                    double rhol, rhou, q0, q1;
                    slack_rho(kind, smu, j, v, Zuq, sc, rhol, rhou, q0, q1);
                    const double G0 = Q<PF_LAM>(o_lg + j) * Q<PF_TINV>(o_lg + j), G1 = Q<PF_LAM>(o_ug + j) * Q<PF_TINV>(o_ug + j);
                    const double coef = fma(-(G0 * rhol), Q<PF_DLI>(j), q0) - fma(-(G1 * rhou), Q<PF_DUI>(j), q1);
                    gt[2] = fma(Q<PF_CX>(j), coef, gt[2]);
                    gt[3] = fma(Q<PF_CY>(j), coef, gt[3]);
                }
            }
            S(o.hg + 2, k) = gt[2]; S(o.hg + 3, k) = gt[3];
        }
        __syncwarp();
    }

    // y = Lxx(kk) (Lxx(kk)^T x): product with the cost-to-go Hessian through its factor
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

    // backward + forward vector sweeps with the current factorisation.  Everything that does not
    // depend on the recursion (z_k = P+ rb_k before, dpi_k = P+ dx+ + p+ after, both through the factor)
    // is node-parallel; the sequential cores are
    //   backward:  s = p+ + z,  w = gu + B^T s,  lu = Luu^-1 w,  p = gx + A^T s - Lxu lu
    //   forward :  du = -Luu^-T (lu + Lxu^T dx),  dx+ = A dx + B du + rb
    // (lanes 0..4 own the components, 2-vectors are computed redundantly by every lane).
    // kind 3 reads the linear residual (lrb) and writes the correction to cv/cpi.
    __device__ void riccati_solve(int kind)
    {
        const bool itr = (kind == 3);
        // right-hand side rb / lrb and the outputs dv / cv, dpi / cpi.  cv always lives in shared memory (it doubles as the
        // node Hessian of the factorisation); dv, dpi, cpi, pv may live in global memory (long horizons): the sweeps only
        // store them, except z_k (kept in the dpi / cpi rows), which the backward sweep then prefetches one node ahead.
        constexpr bool DVS = (SMASK >> PF_DV) & 1u, DPS = (SMASK >> PF_DPI) & 1u;
        static_assert(((SMASK >> PF_DPI) & 1u) == ((SMASK >> PF_CPI) & 1u), "dpi and cpi share a placement");
        const int oRB = itr ? o.lrb : o.rb, oDV = itr ? o.cv : o.pf[PF_DV], oDPI = itr ? o.pf[PF_CPI] : o.pf[PF_DPI];
        auto rDV = [&](int i, int kk) -> double& {
            if (DVS || itr) return sm[(oDV + i) * NS + kk];
            return gm[(size_t)((oDV + i) * GS + kk)];
        };
        auto rDPI = [&](int i, int kk) -> double& {
            if (DPS) return sm[(oDPI + i) * NS + kk];
            return gm[(size_t)((oDPI + i) * GS + kk)];
        };
        for (k = lane; k < N; k += 32) {
            double rb[5], z[5];
#pragma unroll
            for (int i = 0; i < 5; i++) rb[i] = S(oRB + i, k);
            ctg_mul(k + 1, rb, z);
#pragma unroll
            for (int i = 0; i < 5; i++) rDPI(i, k) = z[i];                // z_k (overwritten by dpi_k below)
        }
        __syncwarp();
#ifdef B2_DEBUG_CLOCK
        const long long c1_ = clock64();
#endif
        const int c5 = lane < 5 ? lane : 4;
        int acol[5], arow[7];
#pragma unroll
        for (int r = 0; r < 5; r++) acol[r] = o.sv + B2_BATAB[r][2 + c5];
#pragma unroll
        for (int c = 0; c < 7; c++) arow[c] = o.sv + B2_BATAB[c5][c];
        const double h = p.dt, h2 = p.dt2h;
        // ---- backward
        double pf[5];
        {
            const double pc = S(o.hg + 2 + c5, N);
            if (lane < 5) R<PF_PV>(lane, N) = pc;
#pragma unroll
            for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
        }
        double zc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (!DPS) {
#pragma unroll
            for (int r = 0; r < 5; r++) zc[r] = rDPI(r, N - 1);
        }
        for (int kk = N - 1; kk >= 0; kk--) {
            double s[5], zn[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            if (!DPS && kk > 0) {                                           // next node's z: an L2 round trip off the critical chain
#pragma unroll
                for (int r = 0; r < 5; r++) zn[r] = rDPI(r, kk - 1);
            }
#pragma unroll
            for (int r = 0; r < 5; r++) s[r] = pf[r] + (DPS ? S(oDPI + r, kk) : zc[r]);
            // short dependency trees: the chain through this loop is the critical path of a solve
            const double w0 = fma(S(o.sv + SV_B00, kk), s[0], S(o.hg + 0, kk)) + fma(S(o.sv + SV_B10, kk), s[1], h * s[3]);
            const double w1 = fma(h, s[4], fma(S(o.sv + SV_B01, kk), s[0], S(o.hg + 1, kk)) + fma(S(o.sv + SV_B11, kk), s[1], h2 * s[2]));
#if B2_WPP_FEEDBACK_GAIN
            const double lu0 = w0, lu1 = w1;                        // (G w instead of Lxu (Luu^-1 w))
#else
            const double l00i = S(o.lamu + 0, kk), l10 = S(o.lamu + 1, kk), l11i = S(o.lamu + 2, kk);
            const double lu0 = w0 * l00i;
            const double lu1 = fma(-l10, lu0, w1) * l11i;
#endif
            S(o.lu + 0, kk) = lu0; S(o.lu + 1, kk) = lu1;           // (every lane holds the 2-vector: same value, same address)
            if (kk > 0) {
                const double pa = fma(S(acol[0], kk), s[0], S(o.hg + 2 + c5, kk)) + fma(S(acol[1], kk), s[1], S(acol[2], kk) * s[2]);
                const double pb = fma(S(acol[3], kk), s[3], S(acol[4], kk) * s[4]);
                const double pc = (pa + pb) - fma(S(o.lxu + 2 * c5, kk), lu0, S(o.lxu + 2 * c5 + 1, kk) * lu1);
                R<PF_PV>(c5, kk) = pc;                              // (lanes >= 5 recompute component 4: no divergent store)
#pragma unroll
                for (int r = 0; r < 5; r++) pf[r] = b2_bcast(pc, r);
            }
            if (!DPS) {
#pragma unroll
                for (int r = 0; r < 5; r++) zc[r] = zn[r];
            }
        }
        __syncwarp();
#if B2_WPP_FEEDBACK_GAIN
        // ---- between the sweeps, node-parallel: du0_k = -Luu^-T Luu^-1 w_k (over w_k)
        for (k = lane; k < N; k += 32) {
            const double l00i = S(o.lamu + 0, k), l10 = S(o.lamu + 1, k), l11i = S(o.lamu + 2, k);
            const double lu0 = S(o.lu + 0, k) * l00i;
            const double lu1 = fma(-l10, lu0, S(o.lu + 1, k)) * l11i;
            const double d1 = -lu1 * l11i;
            S(o.lu + 1, k) = d1;
            S(o.lu + 0, k) = fma(-l10, d1, -lu0) * l00i;
        }
        __syncwarp();
#endif
        // ---- forward
        double dx[5] = {0, 0, 0, 0, 0};
        if (lane < 5) rDV(2 + lane, 0) = 0.0;
        for (int kk = 0; kk < N; kk++) {
#if B2_WPP_FEEDBACK_GAIN
            double du0 = S(o.lu + 0, kk), du1 = S(o.lu + 1, kk);
            if (kk > 0) {                                           // du = du0 - G^T dx
                du0 -= fma(S(o.lxu + 2, kk), dx[1], S(o.lxu + 0, kk) * dx[0]) + fma(S(o.lxu + 8, kk), dx[4], fma(S(o.lxu + 6, kk), dx[3], S(o.lxu + 4, kk) * dx[2]));
                du1 -= fma(S(o.lxu + 3, kk), dx[1], S(o.lxu + 1, kk) * dx[0]) + fma(S(o.lxu + 9, kk), dx[4], fma(S(o.lxu + 7, kk), dx[3], S(o.lxu + 5, kk) * dx[2]));
            }
#else
            double r0 = S(o.lu + 0, kk), r1 = S(o.lu + 1, kk);
            if (kk > 0) {
                r0 = fma(S(o.lxu + 2, kk), dx[1], fma(S(o.lxu + 0, kk), dx[0], r0)) + fma(S(o.lxu + 8, kk), dx[4], fma(S(o.lxu + 6, kk), dx[3], S(o.lxu + 4, kk) * dx[2]));
                r1 = fma(S(o.lxu + 3, kk), dx[1], fma(S(o.lxu + 1, kk), dx[0], r1)) + fma(S(o.lxu + 9, kk), dx[4], fma(S(o.lxu + 7, kk), dx[3], S(o.lxu + 5, kk) * dx[2]));
            }
            const double l00i = S(o.lamu + 0, kk), l10 = S(o.lamu + 1, kk), l11i = S(o.lamu + 2, kk);
            const double du1 = -r1 * l11i;
            const double du0 = fma(-l10, du1, -r0) * l00i;
#endif
            const double xa = fma(S(arow[3], kk), dx[1], fma(S(arow[2], kk), dx[0], S(oRB + c5, kk))) + fma(S(arow[6], kk), dx[4], fma(S(arow[5], kk), dx[3], S(arow[4], kk) * dx[2]));
            const double xn = xa + fma(S(arow[0], kk), du0, S(arow[1], kk) * du1);
            rDV(0, kk) = du0; rDV(1, kk) = du1;
            rDV(2 + c5, kk + 1) = xn;
#pragma unroll
            for (int r = 0; r < 5; r++) dx[r] = b2_bcast(xn, r);
        }
        if (lane == 0) { rDV(0, N) = 0.0; rDV(1, N) = 0.0; }
        __syncwarp();
#ifdef B2_DEBUG_CLOCK
        clk_sweep += clock64() - c1_;
#endif
        for (k = lane; k < N; k += 32) {
            double dxn[5], y[5];
#pragma unroll
            for (int i = 0; i < 5; i++) dxn[i] = rDV(2 + i, k + 1);
            ctg_mul(k + 1, dxn, y);
#pragma unroll
            for (int i = 0; i < 5; i++) rDPI(i, k) = y[i] + R<PF_PV>(i, k + 1);
        }
        __syncwarp();
    }

    // node-local expansion of node k: slack steps, dt, dlam for every constraint, consumed on the fly.
    //  XP_SOLVE : base solve `kind` from dv: store dlam (kind 0 with predictor-corrector: store dt*dlam
    //             in the same slot instead, the corrector consumes it), store dsl/dsu, reduce alpha/S1/S2
    //  XP_REFINE: add the correction (cv,cpi) of one refinement step: dlam += -lam*dt_c/t, dsl/dsu += ...,
    //             dv += cv, dpi += cpi; reduce alpha/S1/S2 on the totals
    //  XP_UPDATE: apply the stored step with length a
    //  `lin`    : also assemble the residual of the linear system at the new step (see solve_any())
    template <int MODE>
    __device__ __forceinline__ void expand(int kind, double smu, double a, double& ap, double& ad, double& s1, double& s2,
                                           double* lrg, double& ln0, const bool lin)
    {
        const double sc = scal(k);
        const double Zuq = sc * p.Zu;
        const bool keep_dd = (MODE == XP_SOLVE) && kind == 0 && p.pred_corr;
        double v[7], dv[7], cv[7];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            v[i] = VV(i, k);
            dv[i] = R<PF_DV>(i, k);
            cv[i] = 0.0;
            if (MODE == XP_REFINE) cv[i] = S(o.cv + i, k);       // (dv, dpi already hold the corrected step: apply_correction())
        }
#if B2_WPP_BOX_PIPE
        Box9 bnext = box_load9<MODE != XP_LIN>(0);
#endif
        B2_BOX_LOOP
        for (int s = 0; s < 6; s++) {
#if B2_WPP_BOX_PIPE
            const Box9 bin = bnext;
            if (s + 1 < 6) bnext = box_load9<MODE != XP_LIN>(s + 1);
#endif
            if (!box_on(s, k)) continue;
            const int il = o_lb + s, iu = o_ub + s;
#if B2_WPP_BOX_PIPE
            const double ll = bin.ll, lu_ = bin.lu, tl = bin.tl, tu = bin.tu;
            const double lb = p.lb6[s] - bin.xu, ub = p.ub6[s] - bin.xu;
            const double ddl_ = bin.dl, ddu_ = bin.du, til_ = bin.til, tiu_ = bin.tiu;
#else
            const double ll = Q<PF_LAM>(il), lu_ = Q<PF_LAM>(iu), tl = Q<PF_T>(il), tu = Q<PF_T>(iu);
            double lb, ub;
            box_bounds(s, lb, ub);
            const double ddl_ = Q<PF_DD>(il), ddu_ = Q<PF_DD>(iu);
            const double til_ = MODE != XP_LIN ? Q<PF_TINV>(il) : 0.0, tiu_ = MODE != XP_LIN ? Q<PF_TINV>(iu) : 0.0;
#endif
            const double vv = slot_get(v, s), dvv = slot_get(dv, s);
            const double rdl = (vv - lb) - tl, rdu = (ub - vv) - tu;
            const double dtl = dvv + rdl, dtu = -dvv + rdu;
            double dll, dlu;
            if (MODE == XP_SOLVE) {
                dll = -fma(ll, dtl, rm_of(kind, ll, tl, ddl_, smu)) * til_;
                dlu = -fma(lu_, dtu, rm_of(kind, lu_, tu, ddu_, smu)) * tiu_;
                Q<PF_DD>(il) = keep_dd ? dtl * dll : dll;
                Q<PF_DD>(iu) = keep_dd ? dtu * dlu : dlu;
                if (!B2_WPP_NO_DT) { Q<PF_DT>(il) = dtl; Q<PF_DT>(iu) = dtu; }
            } else if (MODE == XP_REFINE) {
                const double cvv = slot_get(cv, s);
                dll = fma(-(ll * cvv), til_, ddl_);
                dlu = fma(lu_ * cvv, tiu_, ddu_);
                Q<PF_DD>(il) = dll; Q<PF_DD>(iu) = dlu;
                if (!B2_WPP_NO_DT) { Q<PF_DT>(il) = dtl; Q<PF_DT>(iu) = dtu; }
            } else {
                dll = ddl_; dlu = ddu_;
            }
            consume<MODE>(il, ll, tl, dtl, dll, a, ap, ad, s1, s2);
            consume<MODE>(iu, lu_, tu, dtu, dlu, a, ap, ad, s1, s2);
            if (lin) slot_add(lrg, s, -dll + dlu);
        }
        // residual of the linear system: every component but the two that the obstacles touch is complete here -- stored now,
        // so that only lrg[2], lrg[3] stay live across the obstacle loop
        if (lin) {
            if (k == 0) { lrg[4] = lrg[5] = lrg[6] = 0.0; }
            if (k == N) { lrg[0] = lrg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) {
                if (i == 2 || i == 3) continue;
                S(o.hg + i, k) = lrg[i]; ln0 = fmax(ln0, fabs(lrg[i]));
            }
        }
#pragma unroll JU
        for (int j = 0; j < M; j++) {
            prefetch_ahead<MODE == XP_SOLVE ? (FM_SOLVE | FM_LIN) : MODE == XP_REFINE ? (FM_REFINE | FM_LIN) : FM_LIN>(j);
            const int i0 = o_lg + j, i1 = o_ug + j, i2 = o_ls + j, i3 = o_us + j;
            const double cx = Q<PF_CX>(j), cy = Q<PF_CY>(j);
            const double cdv = fma(cy, dv[3], cx * dv[2]);
            const double l0 = Q<PF_LAM>(i0), l1 = Q<PF_LAM>(i1), l2 = Q<PF_LAM>(i2), l3 = Q<PF_LAM>(i3);
            const double t0 = Q<PF_T>(i0), t1 = Q<PF_T>(i1), t2 = Q<PF_T>(i2), t3 = Q<PF_T>(i3);
            const double s_l = Q<PF_SL>(j), s_u = Q<PF_SU>(j);
            const double cval = fma(cy, v[3], cx * v[2]);
            const double rd0 = cval + s_l - Q<PF_LG>(j) - t0;
            const double rd1 = Q<PF_UG>(j) - cval + s_u - t1;
            const double rd2 = s_l + Q<PF_SLI>(j) - t2;
            const double rd3 = s_u + Q<PF_SUI>(j) - t3;
            double dsl, dsu, dl0, dl1, dl2, dl3;
            if (MODE == XP_SOLVE) {
                const double ti0 = Q<PF_TINV>(i0), ti1 = Q<PF_TINV>(i1), ti2 = Q<PF_TINV>(i2), ti3 = Q<PF_TINV>(i3);
                double rhol, rhou, q0, q1;
                slack_rho(kind, smu, j, v, Zuq, sc, rhol, rhou, q0, q1);
                dsl = -fma(l0 * ti0, cdv, rhol) * Q<PF_DLI>(j);
                dsu = -fma(-(l1 * ti1), cdv, rhou) * Q<PF_DUI>(j);
                if (!keep_dd) { Q<PF_DSL>(j) = dsl; Q<PF_DSU>(j) = dsu; }      // (the predictor's slack steps are never read)
                const double dt0 = cdv + dsl + rd0, dt1 = -cdv + dsu + rd1, dt2 = dsl + rd2, dt3 = dsu + rd3;
                dl0 = -fma(l0, dt0, rm_of(kind, l0, t0, Q<PF_DD>(i0), smu)) * ti0;
                dl1 = -fma(l1, dt1, rm_of(kind, l1, t1, Q<PF_DD>(i1), smu)) * ti1;
                dl2 = -fma(l2, dt2, rm_of(kind, l2, t2, Q<PF_DD>(i2), smu)) * ti2;
                dl3 = -fma(l3, dt3, rm_of(kind, l3, t3, Q<PF_DD>(i3), smu)) * ti3;
                Q<PF_DD>(i0) = keep_dd ? dt0 * dl0 : dl0; Q<PF_DD>(i1) = keep_dd ? dt1 * dl1 : dl1;
                Q<PF_DD>(i2) = keep_dd ? dt2 * dl2 : dl2; Q<PF_DD>(i3) = keep_dd ? dt3 * dl3 : dl3;
            } else if (MODE == XP_REFINE) {
                const double ti0 = Q<PF_TINV>(i0), ti1 = Q<PF_TINV>(i1), ti2 = Q<PF_TINV>(i2), ti3 = Q<PF_TINV>(i3);
                const double ccv = fma(cy, cv[3], cx * cv[2]);
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
            if (!B2_WPP_NO_DT && (MODE == XP_SOLVE || MODE == XP_REFINE)) { Q<PF_DT>(i0) = dt0; Q<PF_DT>(i1) = dt1; Q<PF_DT>(i2) = dt2; Q<PF_DT>(i3) = dt3; }
            consume<MODE>(i0, l0, t0, dt0, dl0, a, ap, ad, s1, s2);
            consume<MODE>(i1, l1, t1, dt1, dl1, a, ap, ad, s1, s2);
            consume<MODE>(i2, l2, t2, dt2, dl2, a, ap, ad, s1, s2);
            consume<MODE>(i3, l3, t3, dt3, dl3, a, ap, ad, s1, s2);
            if (lin) {
                const double dl = dl1 - dl0;
                lrg[2] = fma(cx, dl, lrg[2]); lrg[3] = fma(cy, dl, lrg[3]);
                const double rgsl = fma(Q<PF_ZZLQ>(j), s_l, Q<PF_ZLQ>(j)) - l0 - l2;
                const double rgsu = fma(Zuq, s_u, sc * fma(p.Zu, Q<PF_SUI>(j), p.zu)) - l1 - l3;
                const double a0 = fma(Q<PF_ZZLQ>(j), dsl, rgsl) - dl0 - dl2;
                const double a1 = fma(Zuq, dsu, rgsu) - dl1 - dl3;
                Q<PF_LRGSL>(j) = a0; Q<PF_LRGSU>(j) = a1;
                ln0 = fmax(ln0, fmax(fabs(a0), fabs(a1)));
            }
        }
    }

    // running "most limiting" ratios lam/dlam and t/dt (HPIPM d_compute_alpha_qp, negative-reciprocal
    // convention) kept as numerator/denominator pairs: one division per lane at the end
    double ad_n, ad_d, ap_n, ap_d;
#ifdef B2_DEBUG_CLOCK
    long long clk_fact = 0, clk_sweep = 0, clk_all = 0;
#endif

    template <int MODE>
    __device__ __forceinline__ void consume(int i, double l, double t, double dti, double dli, double a,
                                            double& ap, double& ad, double& s1, double& s2)
    {
        if (MODE == XP_SOLVE || MODE == XP_REFINE) {
            // (ad_n/ad_d) * dli > l  with ad_d < 0   <=>   ad_n * dli < l * ad_d
            if (ad_n * dli < l * ad_d) { ad_n = l; ad_d = dli; }
            if (ap_n * dti < t * ap_d) { ap_n = t; ap_d = dti; }
            s1 += fma(l, dti, t * dli);
            s2 = fma(dli, dti, s2);
        }
    }

    __device__ void reduce_step_length(double ap, double ad, double s1, double s2)
    {
        // max of the negative reciprocals == min of the admissible step lengths
        ap = b2_warp_max_neg(ap); ad = b2_warp_max_neg(ad);
        S1 = b2_warp_sum(s1); S2 = b2_warp_sum(s2);
        alpha = -(ap > ad ? ap : ad);
    }

    // One linear solve with the current factorisation + step-length reductions.  kind 0 affine, 1 corrector,
    // 2 centering: base right-hand sides; kind 3: one iterative-refinement step (solve for the linear residual
    // and add the correction).  `kind` is a RUN-TIME value and run() has a single call site: every phase of the
    // interior-point iteration exists once in the instruction stream (the kernel is instruction-fetch
    // sensitive: many warps per SM, each at its own place in a long straight-line program).
    __device__ __forceinline__ void solve_any(int kind, double smu, double lnrm[2])
    {
        rm_cdd = kind == 1 ? 1.0 : 0.0;
        rm_smu = kind == 0 ? 0.0 : smu;
        if (kind != 0) solve_prep(kind, smu);          // the affine right-hand side was assembled by head()
        riccati_solve(kind);
        double ap = -1.0, ad = -1.0, s1 = 0.0, s2 = 0.0, ln0 = 0.0, ln1 = 0.0;
        ad_n = 1.0; ad_d = -1.0; ap_n = 1.0; ap_d = -1.0;
        // FUSE_LIN: the residual of the linear system at the new step (what iterative refinement tests and corrects) is assembled by
        // the SAME pass that expands the step -- it needs exactly the multipliers, slacks and steps that pass holds in registers --
        // instead of a pass of its own (lin_residuals()).  Bit-identical results.  Measured on B200: N = 100, M = 20 (the private
        // rows stream from DRAM, a pass less is 12 % less traffic) 17.8 -> 18.8 k solves/s; N = 20 (rows L2-resident, the linear
        // residual was the cheapest pass) 9.50 -> 9.56 ms per 4096 problems: on for long horizons only.
        // Every base solve but the predictor may be followed by the refinement test.
        const bool lin = FUSE_LIN && kind != 0 && p.itref_corr_max > 0;
        if (kind == 3) {
            // the correction enters the stored step first: the residual of node k reads dv, dpi of its neighbours
            for (k = lane; k <= N; k += 32) {
#pragma unroll
                for (int i = 0; i < 7; i++) R<PF_DV>(i, k) += S(o.cv + i, k);
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < 5; i++) R<PF_DPI>(i, k) += R<PF_CPI>(i, k);
                }
            }
            __syncwarp();
        }
        for (k = lane; k <= N; k += 32) {
            double lrg[7] = {0, 0, 0, 0, 0, 0, 0};
            if (lin) lin_pre(lrg);
            if (kind == 3) expand<XP_REFINE>(3, 0.0, 0.0, ap, ad, s1, s2, lrg, ln0, lin);
            else expand<XP_SOLVE>(kind, smu, 0.0, ap, ad, s1, s2, lrg, ln0, lin);
            if (lin) lin_post(lrg, ln0, ln1);
        }
        ap = ap_n / ap_d; ad = ad_n / ad_d;
        reduce_step_length(ap, ad, s1, s2);
        if (lin) { lnrm[0] = b2_warp_max_nonneg(ln0); lnrm[1] = b2_warp_max_nonneg(ln1); }
        __syncwarp();
    }

    // residual of the linear system at the stored step as a pass of its own (!FUSE_LIN); leaves lrg in rows hg[0..6] for solve_prep(3)
    __device__ void lin_residuals(double lnrm[2])
    {
        double ln0 = 0.0, ln1 = 0.0;
        for (k = lane; k <= N; k += 32) {
            double lrg[7];
            lin_pre(lrg);
            double ap = 0, ad = 0, s1 = 0, s2 = 0;
            expand<XP_LIN>(0, 0.0, 0.0, ap, ad, s1, s2, lrg, ln0, true);
            lin_post(lrg, ln0, ln1);
        }
        lnrm[0] = b2_warp_max_nonneg(ln0); lnrm[1] = b2_warp_max_nonneg(ln1);
        __syncwarp();
    }

    // residual of the linear system at the stored step, node-level parts around expand(): the stationarity residual lrg
    // (-> rows hg[0..6], the right-hand side of a refinement solve) and the dynamics residual lrb (-> rows lrb[0..4])
    __device__ __forceinline__ void lin_pre(double* lrg)
    {
        const double* Hd = k < N ? p.Hd : p.Hde;
#pragma unroll
        for (int i = 0; i < 7; i++) lrg[i] = fma(Hd[i], R<PF_DV>(i, k), Q<PF_RG>(i));
        if (k < N) {
            double sv[10], dp[5], bt[7];
#pragma unroll
            for (int i = 0; i < 10; i++) sv[i] = S(o.sv + i, k);
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
    __device__ __forceinline__ void lin_post(double* lrg, double& ln0, double& ln1)
    {
        if (k < N) {
            double sv[10], dv[7], lrb[5];
#pragma unroll
            for (int i = 0; i < 10; i++) sv[i] = S(o.sv + i, k);
#pragma unroll
            for (int i = 0; i < 7; i++) dv[i] = R<PF_DV>(i, k);
            b2_ba_mul(sv, p.dt, p.dt2h, dv, lrb);
#pragma unroll
            for (int i = 0; i < 5; i++) {
                lrb[i] += S(o.rb + i, k) - R<PF_DV>(2 + i, k + 1);
                S(o.lrb + i, k) = lrb[i];
                ln1 = fmax(ln1, fabs(lrb[i]));
            }
        }
        if (k == 0) { lrg[2] = lrg[3] = 0.0; }
#pragma unroll
        for (int i = 2; i < 4; i++) { S(o.hg + i, k) = lrg[i]; ln0 = fmax(ln0, fabs(lrg[i])); }
    }

    // ---- full-SQP mode only: KKT residuals of the NLP at the current iterate with the multipliers of the previous
    // QP (PI, LAM, T in global memory), i.e. acados' ocp_nlp_res_compute: [stationarity, dynamics gap,
    // g_i - t_i incl. the x0 embedding, lam_i t_i] as infinity norms.  Runs right after prepare_and_init(), whose
    // linearisation (sv, b, g, cx, cy, lg, ug, ...) is everything it needs: node-parallel, one pass.
    __device__ void nlp_residuals(double res[4])
    {
        double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
        for (k = lane; k <= N; k += 32) {
            const double* lk = p.LAM + ((size_t)k * nct) * Bm + b;
            const double* tk = p.T + ((size_t)k * nct) * Bm + b;
            double rg[7];
#pragma unroll
            for (int i = 0; i < 7; i++) rg[i] = Q<PF_G>(i);
            if (k < N) {
                double sv[10], pk[5], bt[7];
#pragma unroll
                for (int i = 0; i < 10; i++) sv[i] = S(o.sv + i, k);
#pragma unroll
                for (int i = 0; i < 5; i++) { pk[i] = p.PI[((size_t)k * 5 + i) * Bm + b]; n1 = fmax(n1, fabs(Q<PF_B>(i))); }
                b2_bat_mul(sv, p.dt, p.dt2h, pk, bt);
#pragma unroll
                for (int i = 0; i < 7; i++) rg[i] += bt[i];
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) rg[2 + i] -= p.PI[((size_t)(k - 1) * 5 + i) * Bm + b];
            } else {
#pragma unroll
                for (int i = 0; i < 5; i++) n2 = fmax(n2, fabs(p.X0[(size_t)i * Bm + b] - p.X[(size_t)i * Bm + b]));
            }
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                double lb, ub;
                box_bounds(s, lb, ub);
                const double ll = lk[(size_t)(o_lb + s) * Bm], lu_ = lk[(size_t)(o_ub + s) * Bm];
                const double tl = tk[(size_t)(o_lb + s) * Bm], tu = tk[(size_t)(o_ub + s) * Bm];
                rg[b2_slot2v(s)] += lu_ - ll;
                n2 = fmax(n2, fmax(fabs(-lb - tl), fabs(ub - tu)));
                n3 = fmax(n3, fmax(fabs(ll * tl), fabs(lu_ * tu)));
            }
            const double sc = scal(k);
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                const double l0 = lk[(size_t)(o_lg + j) * Bm], l1 = lk[(size_t)(o_ug + j) * Bm];
                const double l2 = lk[(size_t)(o_ls + j) * Bm], l3 = lk[(size_t)(o_us + j) * Bm];
                const double t0 = tk[(size_t)(o_lg + j) * Bm], t1 = tk[(size_t)(o_ug + j) * Bm];
                const double t2 = tk[(size_t)(o_ls + j) * Bm], t3 = tk[(size_t)(o_us + j) * Bm];
                const double dl = l1 - l0;
                rg[2] += Q<PF_CX>(j) * dl; rg[3] += Q<PF_CY>(j) * dl;
                const double sli = Q<PF_SLI>(j), sui = Q<PF_SUI>(j);
                const double gsl = Q<PF_ZLQ>(j) - l0 - l2;
                const double gsu = sc * (p.zu + p.Zu * sui) - l1 - l3;
                n0 = fmax(n0, fmax(fabs(gsl), fabs(gsu)));
                // lg = lh - h - sl, ug = uh - h + su  =>  h + sl - lh = -lg,  uh - h + su = ug
                n2 = fmax(n2, fmax(fmax(fabs(-Q<PF_LG>(j) - t0), fabs(Q<PF_UG>(j) - t1)), fmax(fabs(sli - t2), fabs(sui - t3))));
                n3 = fmax(n3, fmax(fmax(fabs(l0 * t0), fabs(l1 * t1)), fmax(fabs(l2 * t2), fabs(l3 * t3))));
            }
            if (k == 0) { rg[2] = rg[3] = rg[4] = rg[5] = rg[6] = 0.0; }       // x_0 is fixed by the embedding
            if (k == N) { rg[0] = rg[1] = 0.0; }
#pragma unroll
            for (int i = 0; i < 7; i++) n0 = fmax(n0, fabs(rg[i]));
        }
        res[0] = b2_warp_max_nonneg(n0); res[1] = b2_warp_max_nonneg(n1); res[2] = b2_warp_max_nonneg(n2); res[3] = b2_warp_max_nonneg(n3);
    }

    // ---- full-SQP mode, globalization MERIT_BACKTRACKING (acados ocp_nlp_line_search): l1 merit function
    //   cost + sum w_pi |dynamics gap| + sum w_lam max(0, violation)   (slacks and their penalties included)
    // at the trial point iterate + al * QP step; node-parallel, one model evaluation per node.  The x0 embedding
    // (node 0, eliminated from the QP) is always taken in full.
    __device__ double merit(double al)
    {
        double m = 0.0;
        for (k = lane; k <= N; k += 32) {
            double xk[5], uk[2] = {0.0, 0.0};
            const double ax = k == 0 ? 1.0 : al;
#pragma unroll
            for (int i = 0; i < 5; i++) xk[i] = p.X[((size_t)k * 5 + i) * Bm + b] + ax * VV(2 + i, k);
            const double* yr = p.YREF + ((size_t)k * 6) * Bm + b;
            const double sc = scal(k);
            const double* Wc = k < N ? p.Wg : p.Wge;
            double cst = Wc[2] * (xk[0] - yr[0 * Bm]) * (xk[0] - yr[0 * Bm]) + Wc[3] * (xk[1] - yr[1 * Bm]) * (xk[1] - yr[1 * Bm])
                       + Wc[5] * (xk[3] - yr[2 * Bm]) * (xk[3] - yr[2 * Bm]) + Wc[6] * (xk[4] - yr[3 * Bm]) * (xk[4] - yr[3 * Bm]);
            double mk = 0.0;
            if (k < N) {
                uk[0] = p.U[((size_t)k * 2 + 0) * Bm + b] + al * VV(0, k);
                uk[1] = p.U[((size_t)k * 2 + 1) * Bm + b] + al * VV(1, k);
                cst += Wc[0] * (uk[0] - yr[4 * Bm]) * (uk[0] - yr[4 * Bm]) + Wc[1] * (uk[1] - yr[5 * Bm]) * (uk[1] - yr[5 * Bm]);
                double xn[5], sv[10];
                if (p.integrator == 1) model_rk4_step(xk, uk, p.dt, xn, sv);
                else model_gl4_step(xk, uk, p.dt, xn, sv);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const double xn1 = p.X[((size_t)(k + 1) * 5 + i) * Bm + b] + al * VV(2 + i, k + 1);
                    mk += Q<PF_WPI>(i) * fabs(xn[i] - xn1);
                }
            }
            mk += 0.5 * cst;
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (!box_on(s, k)) continue;
                const double val = s < 2 ? uk[s] : xk[b2_slot2v(s) - 2];
                const double lo = s < 2 ? p.lbu[s] : p.lbx[s - 2], hi = s < 2 ? p.ubu[s] : p.ubx[s - 2];
                mk += Q<PF_WLAM>(o_lb + s) * fmax(0.0, lo - val) + Q<PF_WLAM>(o_ub + s) * fmax(0.0, val - hi);
            }
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                double hval, cx, cy;
                model_con_h(xk, p.P[((size_t)k * 2 * M + 2 * j) * Bm + b], p.P[((size_t)k * 2 * M + 2 * j + 1) * Bm + b], p.r2, &hval, &cx, &cy);
                const double s_l = p.SL[((size_t)k * M + j) * Bm + b] + al * Q<PF_SL>(j);
                const double s_u = p.SU[((size_t)k * M + j) * Bm + b] + al * Q<PF_SU>(j);
                const double zl_in = p.ZL[((size_t)k * M + j) * Bm + b], Zl_in = p.ZZL[((size_t)k * M + j) * Bm + b];
                mk += Q<PF_WLAM>(o_lg + j) * fmax(0.0, p.lh - hval - s_l) + Q<PF_WLAM>(o_ug + j) * fmax(0.0, hval - p.uh - s_u);
                mk += sc * (zl_in * s_l + 0.5 * Zl_in * s_l * s_l + p.zu * s_u + 0.5 * p.Zu * s_u * s_u);
                mk += Q<PF_WLAM>(o_ls + j) * fmax(0.0, -s_l) + Q<PF_WLAM>(o_us + j) * fmax(0.0, -s_u);
            }
            m += mk;
        }
        return b2_warp_sum(m);
    }

    // weights |multipliers| of the QP at the first SQP iteration, max(|mult|, (w + |mult|)/2) afterwards; then
    // alpha = 1, 0.7, 0.49, ... until the merit function decreases (alpha_min is accepted untested; ls_eps: a few ulp of
    // slack so that the rounding noise of the merit sum does not reject full steps near convergence)
    __device__ double line_search(bool first)
    {
        for (k = lane; k <= N; k += 32) {
            if (k < N) {
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const double a = fabs(PP(i, k));
                    Q<PF_WPI>(i) = first ? a : fmax(a, 0.5 * (Q<PF_WPI>(i) + a));
                }
            }
            for (int i = 0; i < nct; i++) {
                const double a = fabs(Q<PF_LAM>(i));
                Q<PF_WLAM>(i) = first ? a : fmax(a, 0.5 * (Q<PF_WLAM>(i) + a));
            }
        }
        __syncwarp();
        const double m0 = merit(0.0);
        double al = 1.0;
        for (;;) {
            if (al <= p.ls_alpha_min) break;
            const double m1 = merit(al);
            if (m1 < m0 + p.ls_eps * fmax(1.0, fabs(m0))) break;
            al *= p.ls_alpha_red;
            if (al < p.ls_alpha_min) al = p.ls_alpha_min;
        }
        return al;
    }

    // ---- the whole call: one RTI iteration (SQP = false, the reference: nlp_solver_type 'SQP_RTI',
    // robot_ocp_problem.py:130), or acados' full `SQP` loop with FIXED_STEP globalisation (SQP = true:
    // [linearise -> NLP residuals -> convergence test -> QP -> full step] up to nlp_iter_max QPs, the iterate
    // and the multipliers going through global memory between the iterations; status 0 converged, 2 max_iter,
    // 4 QP failure).  Every problem leaves the loop at its own iteration: the CTA pulls the next one.
    __device__ void run()
    {
        int sqp_it = 0, qp_total = 0, status = 0, qs = 0, n_itref = 0;
        double nres[4] = {0.0, 0.0, 0.0, 0.0}, last_al = 1.0;
#ifdef B2_DEBUG_CLOCK
        const long long call0_ = clock64();
#endif
        for (;;) {
        prepare_and_init();
        if (SQP) {
            nlp_residuals(nres);
            if (nres[0] < p.nlp_tol[0] && nres[1] < p.nlp_tol[1] && nres[2] < p.nlp_tol[2] && nres[3] < p.nlp_tol[3]) { status = 0; break; }
            if (sqp_it >= p.nlp_iter_max) { status = 2; break; }
        }
        alpha = 1.0;
        int kk = 0;
        const double inv_nc = 1.0 / nc;
        double a = 0.0;
        for (;; kk++) {
            head(a);
            if (!(kk < p.qp_iter_max && alpha > p.alpha_min &&
                  (nrm[0] > p.res_g_max || nrm[1] > p.res_b_max || nrm[2] > p.res_d_max || nrm[3] > p.res_m_max))) break;
#ifdef B2_DEBUG_CLOCK
            const long long c0_ = clock64();
#endif
            riccati_factorize();
#ifdef B2_DEBUG_CLOCK
            clk_fact += clock64() - c0_;
#endif
            // predictor, corrector, (conditional) centering, <= itref_corr_max refinement steps
            int kind = 0, itref = 0;
            double smu = 0.0, mu_aff0 = 0.0;
            double lnrm[2] = {0.0, 0.0};
            for (;;) {
                solve_any(kind, smu, lnrm);
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
                if (!FUSE_LIN) lin_residuals(lnrm);                              // (FUSE_LIN: assembled by solve_any)
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
        status = (qs != 0 && qs != 1) ? 4 : 0;     // ACADOS_QP_FAILURE leaves the iterate untouched
        qp_total += kk;
        double al = 1.0;
        if (SQP && p.nlp_globalization && status == 0) al = line_search(sqp_it == 0);
        last_al = al;
        if (SQP && al < 1.0) {
            // damped step: primal and dual variables move by al (the x0 embedding in full)
            for (k = lane; k <= N; k += 32) {
                const double ax = k == 0 ? 1.0 : al;
#pragma unroll
                for (int i = 0; i < 5; i++) p.X[((size_t)k * 5 + i) * Bm + b] += ax * VV(2 + i, k);
                if (k < N) {
                    p.U[((size_t)k * 2 + 0) * Bm + b] += al * VV(0, k);
                    p.U[((size_t)k * 2 + 1) * Bm + b] += al * VV(1, k);
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        double* d = p.PI + ((size_t)k * 5 + i) * Bm + b;
                        *d += al * (PP(i, k) - *d);
                    }
                }
#pragma unroll JU
                for (int j = 0; j < M; j++) {
                    p.SL[((size_t)k * M + j) * Bm + b] += al * Q<PF_SL>(j);
                    p.SU[((size_t)k * M + j) * Bm + b] += al * Q<PF_SU>(j);
                }
                for (int i = 0; i < nct; i++) {
                    double* dl = p.LAM + ((size_t)k * nct + i) * Bm + b;
                    double* dt_ = p.T + ((size_t)k * nct + i) * Bm + b;
                    *dl += al * (Q<PF_LAM>(i) - *dl);
                    *dt_ += al * (Q<PF_T>(i) - *dt_);
                }
            }
        } else
        for (k = lane; k <= N; k += 32) {
            if (status == 0) {
#pragma unroll
                for (int i = 0; i < 5; i++) p.X[((size_t)k * 5 + i) * Bm + b] += VV(2 + i, k);
                if (k < N) {
                    p.U[((size_t)k * 2 + 0) * Bm + b] += VV(0, k);
                    p.U[((size_t)k * 2 + 1) * Bm + b] += VV(1, k);
                }
#pragma unroll JU
            for (int j = 0; j < M; j++) {
                    p.SL[((size_t)k * M + j) * Bm + b] += Q<PF_SL>(j);
                    p.SU[((size_t)k * M + j) * Bm + b] += Q<PF_SU>(j);
                }
                // multipliers <- QP multipliers; like the primal iterate they stay untouched on ACADOS_QP_FAILURE
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < 5; i++) p.PI[((size_t)k * 5 + i) * Bm + b] = PP(i, k);
                }
                for (int i = 0; i < nct; i++) {
                    p.LAM[((size_t)k * nct + i) * Bm + b] = Q<PF_LAM>(i);
                    p.T[((size_t)k * nct + i) * Bm + b] = Q<PF_T>(i);
                }
            }
        }
        sqp_it++;
        __syncwarp();                                          // the next linearisation reads what other lanes wrote
        if (!SQP || status != 0) break;
        }
        if (lane == 0) {
            p.STATUS[b] = status; p.QP_ITER[b] = qp_total; p.QP_STATUS[b] = qs;
            p.STATS[0 * Bm + b] = nrm[0]; p.STATS[1 * Bm + b] = nrm[1]; p.STATS[2 * Bm + b] = nrm[2]; p.STATS[3 * Bm + b] = nrm[3];
            p.STATS[4 * Bm + b] = mu; p.STATS[5 * Bm + b] = alpha; p.STATS[6 * Bm + b] = (double)n_itref; p.STATS[7 * Bm + b] = 0.0;
#ifdef B2_DEBUG_CLOCK
            p.STATS[5 * Bm + b] = (double)clk_fact; p.STATS[6 * Bm + b] = (double)clk_sweep; p.STATS[7 * Bm + b] = (double)(clock64() - call0_);
#endif
            if (SQP) {
                p.STATS[7 * Bm + b] = last_al;
                p.SQP_ITER[b] = sqp_it;
                p.NLP_RES[0 * Bm + b] = nres[0]; p.NLP_RES[1 * Bm + b] = nres[1]; p.NLP_RES[2 * Bm + b] = nres[2]; p.NLP_RES[3 * Bm + b] = nres[3];
            }
        }
        __syncwarp();
    }
};

#ifdef B2_EXP_MAXNREG          // experiment: register budget instead of a minimum CTA count
#define B2_WPP_BOUNDS(NT, SMASK) __maxnreg__(B2_EXP_MAXNREG)
#else
#define B2_WPP_BOUNDS(NT, SMASK) __launch_bounds__(32, (NT > 0 && NT < 32 && SMASK != B2_WPP_ALL_SHARED) ? B2_WPP_MINB : 1)
#endif
template <int MT, int NT, unsigned long long SMASK, bool SQP = false>
__global__ void B2_WPP_BOUNDS(NT, SMASK) b2_rti_solve_wpp(const B2Params p)
{
    B2_DYNAMIC_SMEM(double, b2_smem);
    double* const gm = p.WPRIV + (size_t)blockIdx.x * p.wpriv_stride;      // this CTA's block of node-private rows
    const int nq = p.NQ ? *p.NQ : p.B;
#if defined(B2_WPP_STAGGER) && !defined(SIMT_EMU)
    // experiment: the resident CTAs of an SM start B2_WPP_STAGGER cycles apart, so that the first wave of a launch does not walk
    // through the phases of the interior-point iteration in lock-step (timing only, results unchanged)
    {
        const long long t0 = clock64(), d = (long long)((blockIdx.x / 148u) % 12u) * B2_WPP_STAGGER;
        while (clock64() - t0 < d) __nanosleep(500);
    }
#endif
    for (;;) {
        int i = 0;
        if ((threadIdx.x & 31) == 0) i = atomicAdd(p.NEXT, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nq) break;
        const int b = p.ORDER ? p.ORDER[i] : i;      // expensive problems first (shorter tail), see b2_order_problems
        WppSolver<MT, NT, SMASK, SQP> s(p, b, b2_smem, gm);
        s.run();
    }
}
