/*
 * wpp_mirror.cpp -- CPU MIRROR of the CUDA solve kernels.  TEST INFRASTRUCTURE, not product code.
 *
 * The solver kernels of dynamic-obstacle-avoidance-mpc_b200/csrc/ (mpc_wpp.cuh and the headers it includes) are compiled
 * here, UNCHANGED, with g++ against the SIMT emulator of simt_emu.h and executed warp by warp on the host.  The kernels are
 * written with explicit fma() and a software sincos (csrc/b2_math.cuh); the device build uses -fmad=false, this build
 * -ffp-contract=off, so both execute the same IEEE-754 operations in the same order (divisions and square roots are
 * correctly rounded on both sides; reductions are the same shuffle butterflies).  tests/ therefore demand BIT-IDENTICAL
 * iterates, multipliers and iteration counts between this library and the GPU -- the answer to "is the +-1 iteration noise
 * between GPU and oracle inherent?": against this mirror there is none.
 *
 * What this proves: the GPU executes the kernel source faithfully (no races, no uninitialised reads, no compiler
 * re-association).  What it does NOT prove: that the source implements the right algorithm -- that is the job of the
 * INDEPENDENT oracle (oracle/mpc_oracle.c: plain Cholesky, sequential sums, different code), compared with tolerances.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's checker legs load the library built from this file.
 */
#define SIMT_EMU_IMPL
#include "simt_emu.h"

#include "../../dynamic-obstacle-avoidance-mpc_b200/csrc/mpc_params.h"
#include "../../dynamic-obstacle-avoidance-mpc_b200/csrc/mpc_wpp.cuh"
#include "../../dynamic-obstacle-avoidance-mpc_b200/csrc/mpc_cta.cuh"

namespace {

// batch-first [B][len] <-> batch-last [len][B]
void to_soa(const double* a, double* s, int64_t len, int64_t B)
{
    for (int64_t b = 0; b < B; b++) for (int64_t l = 0; l < len; l++) s[l * B + b] = a[b * len + l];
}
void to_aos(const double* s, double* a, int64_t len, int64_t B)
{
    for (int64_t b = 0; b < B; b++) for (int64_t l = 0; l < len; l++) a[b * len + l] = s[l * B + b];
}

template <int MT, int NT, unsigned long long SMASK, bool SQP>
void run_kernel(const B2Params& p, unsigned grid, size_t smem, int host_threads)
{
    simt::run_grid(grid, 32, smem, host_threads, [&]() { b2_rti_solve_wpp<MT, NT, SMASK, SQP>(p); });
}

}   // namespace

extern "C" {

/* One b2mpc_solve() of `B` problems on the CPU, through the kernel source.  Arrays batch-first (what numpy holds):
 *   x0[B,5] p[B,(N+1)*2M] yref[B,(N+1)*6] zl,Zl[B,(N+1)*M]   inputs
 *   x[B,(N+1)*5] u[B,N*2] sl,su[B,(N+1)*M]                   iterate, in/out
 *   pi[B,N*5] lam,t[B,(N+1)*(12+4M)] stats[B,8]              outputs (pi/lam/t are in/out in SQP mode; may be NULL otherwise)
 *   status, qp_iter, qp_status, sqp_iter [B]                 outputs (sqp_iter may be NULL)
 * layout: 2 = node-private rows in "global" memory (the persistent throughput kernel), 3 = everything in shared memory,
 *         4 = the CTA-per-problem latency kernel (mpc_cta.cuh; M = 5, N = 20 or 10, RTI only).
 * sqp: 0 = one RTI iteration (cfg->nlp_solver ignored), 1 = the full SQP loop.
 * Returns 0, or -1 for an unsupported configuration. */
int b2mirror_solve(const b2mpc_config* cfg, int layout, int sqp, int64_t B, int host_threads,
                   const double* x0, const double* p_in, const double* yref, const double* zl, const double* Zl,
                   double* x, double* u, double* sl, double* su, double* pi, double* lam, double* t, double* stats,
                   double* nlp_res, int32_t* status, int32_t* qp_iter, int32_t* qp_status, int32_t* sqp_iter)
{
    const int N = cfg->N, M = cfg->n_obst, nct = 12 + 4 * M;
    if (N < 1 || N > B2MPC_MAX_N || M < 1 || M > B2MPC_MAX_OBST || B < 1) return -1;
    if (layout != 2 && layout != 3 && layout != 4) return -1;
    if (layout == 4 && (sqp || M != 5 || (N != 20 && N != 10))) return -1;      // CTA-per-problem kernel: reference dimensions, RTI
    B2Params prm;
    b2_params_from_config(*cfg, (int)B, prm);
    prm.B = (int)B;
    const int64_t len[] = {5, (int64_t)(N + 1) * 2 * M, (int64_t)(N + 1) * 6, (int64_t)(N + 1) * M, (int64_t)(N + 1) * M,
                           (int64_t)(N + 1) * 5, (int64_t)N * 2, (int64_t)(N + 1) * M, (int64_t)(N + 1) * M,
                           (int64_t)N * 5, (int64_t)(N + 1) * nct, (int64_t)(N + 1) * nct, 8, 4};
    std::vector<std::vector<double>> f(14);
    for (int i = 0; i < 14; i++) f[i].assign((size_t)len[i] * B, 0.0);
    const double* in[] = {x0, p_in, yref, zl, Zl, x, u, sl, su, sqp ? pi : nullptr, sqp ? lam : nullptr, sqp ? t : nullptr};
    for (int i = 0; i < 12; i++) if (in[i]) to_soa(in[i], f[i].data(), len[i], B);
    std::vector<int32_t> ist(B, 0), iit(B, 0), iqs(B, 0), isq(B, 0);
    int32_t next = 0;
    prm.X0 = f[0].data(); prm.P = f[1].data(); prm.YREF = f[2].data(); prm.ZL = f[3].data(); prm.ZZL = f[4].data();
    prm.X = f[5].data(); prm.U = f[6].data(); prm.SL = f[7].data(); prm.SU = f[8].data();
    prm.PI = f[9].data(); prm.LAM = f[10].data(); prm.T = f[11].data(); prm.STATS = f[12].data(); prm.NLP_RES = f[13].data();
    prm.STATUS = ist.data(); prm.QP_ITER = iit.data(); prm.QP_STATUS = iqs.data(); prm.SQP_ITER = isq.data();
    prm.ORDER = nullptr; prm.NEXT = &next;
    if (sqp) prm.nlp_iter_max = cfg->nlp_iter_max;

    const bool ref_dims = (M == 5 && N == 20), ref_short = (M == 5 && N == 10), long_dims = (M == 20 && N == 100);
    // the same instantiation the C-ABI picks (csrc/b200mpc.cu: b2mpc_create), so that compile-time dimensions are mirrored too
    unsigned long long smask = layout == 3 ? B2_WPP_ALL_SHARED : ((long_dims && !sqp) ? B2_WPP_LONG : B2_WPP_ALL_GLOBAL);
    const unsigned grid = (unsigned)(B < 64 ? B : 64);               // resident "CTAs" of the emulated launch: any number works
    const size_t smem = b2_wpp_smem_bytes(N, M, smask, sqp != 0);
    prm.wpriv_stride = b2_wpp_priv_doubles(N, M, smask, sqp != 0);
    std::vector<double> wpriv((size_t)prm.wpriv_stride * grid + 1, 0.0);
    prm.WPRIV = wpriv.data();
    if (host_threads < 1) host_threads = 1;
#define RUN(MT, NT, SM) do { if (sqp) run_kernel<MT, NT, SM, true>(prm, grid, smem, host_threads); \
                             else run_kernel<MT, NT, SM, false>(prm, grid, smem, host_threads); } while (0)
    if (layout == 4) {
        const size_t smem4 = b2_cta_smem_bytes(N, M);
        if (N == 20) simt::run_grid(grid, CtaSolver<5, 20>::NTHR, smem4, host_threads, [&]() { b2_rti_solve_cta<5, 20>(prm); });
        else simt::run_grid(grid, CtaSolver<5, 10>::NTHR, smem4, host_threads, [&]() { b2_rti_solve_cta<5, 10>(prm); });
    }
    else if (layout == 3) { if (ref_dims) RUN(5, 20, B2_WPP_ALL_SHARED); else RUN(0, 0, B2_WPP_ALL_SHARED); }
    else if (ref_dims) RUN(5, 20, B2_WPP_ALL_GLOBAL);
    else if (ref_short && !sqp) RUN(5, 10, B2_WPP_ALL_GLOBAL);
    else if (long_dims && !sqp) RUN(20, 100, B2_WPP_LONG);
    else RUN(0, 0, B2_WPP_ALL_GLOBAL);
#undef RUN
    double* out[] = {nullptr, nullptr, nullptr, nullptr, nullptr, x, u, sl, su, pi, lam, t, stats, nlp_res};
    for (int i = 5; i < 14; i++) if (out[i]) to_aos(f[i].data(), out[i], len[i], B);
    for (int64_t b = 0; b < B; b++) {
        if (status) status[b] = ist[b];
        if (qp_iter) qp_iter[b] = iit[b];
        if (qp_status) qp_status[b] = iqs[b];
        if (sqp_iter) sqp_iter[b] = isq[b];
    }
    return 0;
}

/* the model functions as the kernels evaluate them (b2_math.cuh sincos), for unit tests */
void b2mirror_gl4_step(const double* x, const double* u, double h, double* xn, double* sv) { model_gl4_step(x, u, h, xn, sv); }
void b2mirror_rk4_step(const double* x, const double* u, double h, double* xn, double* sv) { model_rk4_step(x, u, h, xn, sv); }

}   // extern "C"
