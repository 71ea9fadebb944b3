// b200mpc.cu -- C-ABI of libb200mpc.so (include/b200mpc.h) and the glue kernels around the solve.
//
// Host side mirrors the object the reference builds at src/simulation/robot_ocp_problem.py:135-136
// (AcadosOcpSolver + AcadosSimSolver) for a BATCH of independent problems.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/b200mpc.h"
#include "mpc_common.cuh"
#include "mpc_params.h"
#include "mpc_wpp.cuh"
#include "mpc_cta.cuh"

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) return fail(B2MPC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// Every entry point runs on the handle's device whatever device is current in the calling thread, and leaves the
// caller's current device as it found it (two handles on two GPUs in one process must not disturb each other).
struct B2DeviceGuard {
    int prev = -1, dev;
    explicit B2DeviceGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
    ~B2DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

#define B2_ORDER_MAXKEYS 72         // >= number of keys of the launch-order counting sort (B2_ORDER_NKEYS)

struct b2mpc {
    b2mpc_config cfg;
    int N, M, Bm, nct;
    int64_t len[B2MPC_N_FIELDS];
    double* f[B2MPC_N_FIELDS];       // SoA [len][Bm]
    int32_t* fi[B2MPC_N_IFIELDS];    // [Bm]
    double* stage_d;                 // device staging for layout conversion (max field len * Bm)
    size_t stage_elems;
    int64_t launches;
    int64_t loop_step;               // control steps taken since the last b2mpc_loop_reset (Philox counter of the noise stream)
    int32_t* order;                  // launch order for the warp kernel + 2 counters
    int kernel;                      // resolved: 1 thread-per-problem, 2 warp-per-problem (hybrid layout), 3 warp-per-problem (all shared)
    size_t wpp_smem;
    void (*wpp_fn)(const B2Params);  // selected instantiation of the warp kernel
    int wpp_grid;                    // resident CTAs (SMs x occupancy): the persistent grid
    int l2_persist;                  // experiment (B2MPC_L2_PERSIST=1): launch with an L2 access-policy window over wpriv
    float l2_hit_ratio;
    double* wpriv;                   // node-private rows of the resident CTAs
    size_t wpriv_bytes;
    void (*wpp_fn_small)(const B2Params);   // auto mode: all-shared instantiation for batches that fit the GPU at once
    void (*wpp_fn_sqp)(const B2Params);     // full-SQP instantiation: same grid; its layout adds the merit-weight rows
    size_t wpp_smem_sqp, wpriv_stride_sqp;
    size_t wpp_smem_small;
    int small_batch_max;
    void (*cta_fn)(const B2Params);         // CTA-per-problem latency kernel (mpc_cta.cuh), reference dimensions only
    size_t cta_smem;
    int cta_threads, cta_grid;
    B2Params prm;
};

extern "C" int b2mpc_abi_version(void) { return B2MPC_ABI_VERSION; }
extern "C" const char* b2mpc_last_error(void) { return g_err.c_str(); }

extern "C" void b2mpc_default_config(b2mpc_config* c, int32_t N, int32_t n_obst, double tf, int32_t max_batch, int32_t device)
{
    memset(c, 0, sizeof(*c));
    c->N = N; c->n_obst = n_obst; c->device = device; c->max_batch = max_batch;
    c->integrator = 0; c->qp_iter_max = 50; c->lm_scaled_by_dt = 1; c->kernel = 0;   // world_specification.py:48
    c->pred_corr = 1; c->cond_pred_corr = 1; c->itref_corr_max = 2;
    c->tf = tf; c->lm = 2.0;                                       // robot_ocp_problem.py:128,132
    for (int i = 0; i < 4; i++) c->W[i] = 2.0;                     // Q = 2*I4, :26
    c->W[4] = c->W[5] = 0.15;                                      // R = 0.15*I2, :24
    for (int i = 0; i < 4; i++) c->W_e[i] = 5.0;                   // Q_e = 5*I4, :27
    c->lbx[0] = c->lbx[1] = -7.0; c->lbx[2] = c->lbx[3] = -MODEL_WORLD_V_MAX_ROBOT;   // :91
    c->ubx[0] = c->ubx[1] = 7.0; c->ubx[2] = c->ubx[3] = MODEL_WORLD_V_MAX_ROBOT;     // :92
    c->lbu[0] = c->lbu[1] = -MODEL_WORLD_C_MAX; c->ubu[0] = c->ubu[1] = MODEL_WORLD_C_MAX;   // :95-96
    c->r2 = MODEL_R2;                                              // robot_model.py:62
    c->lh = 0.0; c->uh = 1e5;                                      // :99-103
    c->zu = 0.0; c->Zu = 0.0;                                      // :118-122
    c->mu0 = 1e0; c->alpha_min = 1e-8;                             // acados' HPIPM defaults
    c->res_g_max = 1e-6; c->res_b_max = 1e-8; c->res_d_max = 1e-8; c->res_m_max = 1e-8;
    c->reg_prim = 1e-15; c->lam_min = 1e-16; c->t_min = 1e-16; c->tau_min = 1e-16; c->thr0 = 1e-1;
    c->nlp_solver = 0;                                             // 'SQP_RTI', robot_ocp_problem.py:130
    c->nlp_iter_max = 100;                                         // acados defaults for the full-SQP mode
    for (int i = 0; i < 4; i++) c->nlp_tol[i] = 1e-6;
    c->nlp_globalization = 0;                                      // 'FIXED_STEP'
    c->ls_alpha_min = 0.05; c->ls_alpha_reduction = 0.7;           // acados line-search defaults
    c->ls_eps = 0.0;                                               // relative slack of the merit decrease test (0 = the plain test)
}

static void fill_params(b2mpc* h)
{
    B2Params& p = h->prm;
    b2_params_from_config(h->cfg, h->Bm, p);          // everything that is not a pointer (shared with the CPU mirror build)
    p.X0 = h->f[B2MPC_X0]; p.P = h->f[B2MPC_P]; p.YREF = h->f[B2MPC_YREF]; p.ZL = h->f[B2MPC_ZL]; p.ZZL = h->f[B2MPC_ZZL];
    p.X = h->f[B2MPC_X]; p.U = h->f[B2MPC_U]; p.SL = h->f[B2MPC_SL]; p.SU = h->f[B2MPC_SU];
    p.PI = h->f[B2MPC_PI]; p.LAM = h->f[B2MPC_LAM]; p.T = h->f[B2MPC_T]; p.STATS = h->f[B2MPC_STATS];
    p.STATUS = h->fi[B2MPC_STATUS]; p.QP_ITER = h->fi[B2MPC_QP_ITER]; p.QP_STATUS = h->fi[B2MPC_QP_STATUS];
    p.NLP_RES = h->f[B2MPC_NLP_RES]; p.SQP_ITER = h->fi[B2MPC_SQP_ITER];
}

extern "C" int b2mpc_create(const b2mpc_config* cfg, b2mpc** out)
{
    if (!cfg || !out) return fail(B2MPC_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->N < 1 || cfg->N > B2MPC_MAX_N) return fail(B2MPC_EINVAL, "N=%d out of range [1,%d]", cfg->N, B2MPC_MAX_N);
    if (cfg->n_obst < 1 || cfg->n_obst > B2MPC_MAX_OBST) return fail(B2MPC_EINVAL, "n_obst=%d out of range [1,%d]", cfg->n_obst, B2MPC_MAX_OBST);
    if (cfg->max_batch < 1) return fail(B2MPC_EINVAL, "max_batch must be >= 1");
    if (!(cfg->tf > 0.0)) return fail(B2MPC_EINVAL, "tf must be > 0");
    if (cfg->integrator != 0 && cfg->integrator != 1) return fail(B2MPC_EINVAL, "integrator must be 0 (IRK-GL4) or 1 (RK4)");
    if (cfg->qp_iter_max < 1) return fail(B2MPC_EINVAL, "qp_iter_max must be >= 1");
    if (cfg->nlp_solver != 0 && cfg->nlp_solver != 1) return fail(B2MPC_EINVAL, "nlp_solver must be 0 (SQP_RTI) or 1 (SQP)");
    if (cfg->nlp_solver == 1 && cfg->nlp_iter_max < 1) return fail(B2MPC_EINVAL, "nlp_iter_max must be >= 1");
    if (cfg->kernel == 1) return fail(B2MPC_EINVAL, "kernel 1 (the thread-per-problem cross-check) was removed in ABI 3: the cross-check is the CPU mirror build under oracle/simt/");
    if (cfg->nlp_globalization != 0 && cfg->nlp_globalization != 1) return fail(B2MPC_EINVAL, "nlp_globalization must be 0 (FIXED_STEP) or 1 (MERIT_BACKTRACKING)");
    if (cfg->nlp_globalization == 1 && !(cfg->ls_alpha_min > 0.0 && cfg->ls_alpha_min <= 1.0 && cfg->ls_alpha_reduction > 0.0 && cfg->ls_alpha_reduction < 1.0))
        return fail(B2MPC_EINVAL, "line search needs 0 < ls_alpha_min <= 1 and 0 < ls_alpha_reduction < 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(B2MPC_ENODEV, "no CUDA device: libb200mpc has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(B2MPC_ENODEV, "device %d not present (%d devices)", cfg->device, ndev);
    B2DeviceGuard guard(cfg->device);
    b2mpc* h = new b2mpc();
    memset((void*)h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->N = cfg->N; h->M = cfg->n_obst; h->Bm = cfg->max_batch; h->nct = 12 + 4 * cfg->n_obst;
    const int N = h->N, M = h->M;
    h->len[B2MPC_X0] = 5; h->len[B2MPC_P] = (int64_t)(N + 1) * 2 * M; h->len[B2MPC_YREF] = (int64_t)(N + 1) * 6;
    h->len[B2MPC_ZL] = h->len[B2MPC_ZZL] = h->len[B2MPC_SL] = h->len[B2MPC_SU] = (int64_t)(N + 1) * M;
    h->len[B2MPC_X] = (int64_t)(N + 1) * 5; h->len[B2MPC_U] = (int64_t)N * 2; h->len[B2MPC_PI] = (int64_t)N * 5;
    h->len[B2MPC_LAM] = h->len[B2MPC_T] = (int64_t)(N + 1) * h->nct;
    h->len[B2MPC_STATS] = 8; h->len[B2MPC_GOAL] = 2; h->len[B2MPC_OBST] = (int64_t)M * 4;
    h->len[B2MPC_MARGIN] = 2; h->len[B2MPC_U0] = 2; h->len[B2MPC_NLP_RES] = 4;
    int64_t maxlen = 0;
    for (int i = 0; i < B2MPC_N_FIELDS; i++) {
        if (h->len[i] > maxlen) maxlen = h->len[i];
        size_t bytes = sizeof(double) * (size_t)h->len[i] * h->Bm;
        if (cudaMalloc(&h->f[i], bytes) != cudaSuccess) { b2mpc_destroy(h); return fail(B2MPC_ENOMEM, "cudaMalloc(field %d, %zu B) failed", i, bytes); }
        cudaMemset(h->f[i], 0, bytes);
    }
    for (int i = 0; i < B2MPC_N_IFIELDS; i++) {
        if (cudaMalloc(&h->fi[i], sizeof(int32_t) * h->Bm) != cudaSuccess) { b2mpc_destroy(h); return fail(B2MPC_ENOMEM, "cudaMalloc(int field) failed"); }
        cudaMemset(h->fi[i], 0, sizeof(int32_t) * h->Bm);
    }
    h->stage_elems = (size_t)maxlen * h->Bm;
    if (cudaMalloc(&h->stage_d, sizeof(double) * h->stage_elems) != cudaSuccess) {
        b2mpc_destroy(h);
        return fail(B2MPC_ENOMEM, "staging allocation failed");
    }
    fill_params(h);
    // kernel selection.  The warp-per-problem kernel keeps the cross-lane part of a problem in shared
    // memory; the node-private rows go to shared memory too (kernel 3, needs the whole problem to fit) or,
    // per field, to an L2-resident block of global memory per resident CTA (kernel 2, any N <= 128).
    unsigned long long smask = B2_WPP_ALL_GLOBAL, smask_sqp = B2_WPP_ALL_GLOBAL;      // layouts of wpp_fn / wpp_fn_sqp
    h->wpp_fn = nullptr;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const bool ref_dims = (M == 5 && N == 20);          // the reference's configuration (world_specification.py:44-48)
    const bool ref_dims_short = (M == 5 && N == 10);    // the reference's other recorded setting: TF = 1 s => N_SOLV = 10
    const bool long_dims = (M == 20 && N == 100);       // BASELINE.json config 4
    if (cfg->kernel < 0 || cfg->kernel > 4) { b2mpc_destroy(h); return fail(B2MPC_EINVAL, "kernel must be 0 (auto), 2, 3 or 4"); }
    if (ref_dims) { h->cta_fn = b2_rti_solve_cta<5, 20>; h->cta_threads = CtaSolver<5, 20>::NTHR; }
    else if (ref_dims_short) { h->cta_fn = b2_rti_solve_cta<5, 10>; h->cta_threads = CtaSolver<5, 10>::NTHR; }
    if (cfg->kernel == 4 && !h->cta_fn) {
        b2mpc_destroy(h);
        return fail(B2MPC_EINVAL, "kernel 4 (CTA-per-problem) is instantiated for the reference's dimensions: 5 obstacles, N = 20 or N = 10");
    }
    if (cfg->kernel == 4 && cfg->nlp_solver == 1) { b2mpc_destroy(h); return fail(B2MPC_EINVAL, "kernel 4 runs SQP_RTI only (nlp_solver = 0)"); }
    if (cfg->kernel == 3) {
        smask = smask_sqp = B2_WPP_ALL_SHARED;
        if (b2_wpp_smem_bytes(N, M, smask) > (size_t)max_smem) {
            b2mpc_destroy(h);
            return fail(B2MPC_EINVAL, "all-shared warp kernel needs %zu B > %d B of shared memory", b2_wpp_smem_bytes(N, M, smask), max_smem);
        }
        h->wpp_fn = ref_dims ? b2_rti_solve_wpp<5, 20, B2_WPP_ALL_SHARED> : b2_rti_solve_wpp<0, 0, B2_WPP_ALL_SHARED>;
        h->wpp_fn_sqp = ref_dims ? b2_rti_solve_wpp<5, 20, B2_WPP_ALL_SHARED, true> : b2_rti_solve_wpp<0, 0, B2_WPP_ALL_SHARED, true>;
        h->kernel = 3;
    } else {
        h->wpp_fn = ref_dims ? b2_rti_solve_wpp<5, 20, B2_WPP_ALL_GLOBAL>
                  : ref_dims_short ? b2_rti_solve_wpp<5, 10, B2_WPP_ALL_GLOBAL>
                  : long_dims ? b2_rti_solve_wpp<20, 100, B2_WPP_LONG> : b2_rti_solve_wpp<0, 0, B2_WPP_ALL_GLOBAL>;
        if (long_dims) smask = B2_WPP_LONG;                  // v, pi, dv, dpi, cpi, pv in global memory too: 4 CTAs/SM instead of 2
        h->wpp_fn_sqp = ref_dims ? b2_rti_solve_wpp<5, 20, B2_WPP_ALL_GLOBAL, true> : b2_rti_solve_wpp<0, 0, B2_WPP_ALL_GLOBAL, true>;
        h->kernel = cfg->kernel == 4 ? 4 : 2;
    }
    if (h->cta_fn) {
        h->cta_smem = b2_cta_smem_bytes(N, M);
        int occ_c = 0, sms_c = 0;
        cudaDeviceGetAttribute(&sms_c, cudaDevAttrMultiProcessorCount, cfg->device);
        if (h->cta_smem > (size_t)max_smem ||
            cudaFuncSetAttribute(h->cta_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->cta_smem) != cudaSuccess ||
            cudaFuncSetAttribute(h->cta_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, h->cta_fn, h->cta_threads, h->cta_smem) != cudaSuccess || occ_c < 1) {
            if (cfg->kernel == 4) { b2mpc_destroy(h); return fail(B2MPC_ECUDA, "CTA-per-problem kernel does not fit an SM"); }
            h->cta_fn = nullptr;
        }
        h->cta_grid = occ_c * sms_c;
        if (const char* e = getenv("B2MPC_CTA_GRID")) { if (atoi(e) > 0) h->cta_grid = atoi(e); }       // experiment: resident CTAs
    }
    {
        h->wpp_smem = b2_wpp_smem_bytes(N, M, smask);
        h->wpp_smem_sqp = b2_wpp_smem_bytes(N, M, smask_sqp, true);
        if (h->wpp_smem_sqp > (size_t)max_smem) { b2mpc_destroy(h); return fail(B2MPC_EINVAL, "warp kernel needs %zu B > %d B of shared memory", h->wpp_smem, max_smem); }
        if (cudaMalloc(&h->order, sizeof(int32_t) * (2 * (size_t)h->Bm + 4 * B2_ORDER_MAXKEYS + 8)) != cudaSuccess) { b2mpc_destroy(h); return fail(B2MPC_ENOMEM, "cudaMalloc(order) failed"); }
        cudaError_t e = cudaFuncSetAttribute(h->wpp_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->wpp_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(h->wpp_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(h->wpp_fn_sqp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->wpp_smem_sqp);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(h->wpp_fn_sqp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { b2mpc_destroy(h); return fail(B2MPC_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
        int occ = 0, sms = 0, occ_q = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_q, h->wpp_fn_sqp, 32, h->wpp_smem_sqp);
        if (occ_q < 1) { b2mpc_destroy(h); return fail(B2MPC_ECUDA, "SQP warp kernel does not fit an SM"); }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->wpp_fn, 32, h->wpp_smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        if (occ < 1) { b2mpc_destroy(h); return fail(B2MPC_ECUDA, "warp kernel does not fit an SM"); }
        if (const char* e = getenv("B2MPC_WPP_OCC")) { if (atoi(e) > 0 && atoi(e) < occ) occ = atoi(e); }   // experiment: fewer resident CTAs per SM
        h->wpp_grid = occ * sms;
        if (h->wpp_grid > h->Bm) h->wpp_grid = h->Bm;
        h->prm.wpriv_stride = b2_wpp_priv_doubles(N, M, smask);              // RTI launches keep their compact blocks (L2 footprint)
        h->wpriv_stride_sqp = b2_wpp_priv_doubles(N, M, smask_sqp, true);
        h->wpriv_bytes = sizeof(double) * (h->wpriv_stride_sqp > h->prm.wpriv_stride ? h->wpriv_stride_sqp : h->prm.wpriv_stride) * (size_t)h->wpp_grid;
        if (h->wpriv_bytes) {
            if (cudaMalloc(&h->wpriv, h->wpriv_bytes) != cudaSuccess) { b2mpc_destroy(h); return fail(B2MPC_ENOMEM, "cudaMalloc(node-private rows, %zu B) failed", h->wpriv_bytes); }
            cudaMemset(h->wpriv, 0, h->wpriv_bytes);
        }
        h->prm.WPRIV = h->wpriv;
        if (const char* e = getenv("B2MPC_L2_PERSIST")) {
            if (atoi(e) > 0) {
                cudaDeviceProp prop;
                cudaGetDeviceProperties(&prop, cfg->device);
                size_t want = sizeof(double) * h->prm.wpriv_stride * (size_t)h->wpp_grid;
                size_t lim = want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize;
                if (atoi(e) > 1) lim = lim * atoi(e) / 100;                    // B2MPC_L2_PERSIST=<percent of the maximum>
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim);
                h->l2_persist = 1;
                h->l2_hit_ratio = want > lim ? (float)((double)lim / (double)want) : 1.0f;
                if ((size_t)prop.accessPolicyMaxWindowSize < want) h->l2_persist = 0;
                fprintf(stderr, "[b200mpc] L2 persist: want %zu B, set-aside %zu B (max %d), window max %d, hitRatio %.2f, on=%d\n",
                        want, lim, prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize, h->l2_hit_ratio, h->l2_persist);
            }
        }
        // auto mode: a batch that fits the GPU in one wave of all-shared CTAs (<= 4 per SM at N=20, M=5) has no use
        // for more resident CTAs, and the all-shared layout has the shorter per-problem latency (no L2 round trips):
        // single closed-loop episodes (BASELINE config 1) and small batches take that instantiation.
        if (cfg->kernel == 0 && b2_wpp_smem_bytes(N, M, B2_WPP_ALL_SHARED) <= (size_t)max_smem) {
            h->wpp_fn_small = ref_dims ? b2_rti_solve_wpp<5, 20, B2_WPP_ALL_SHARED> : b2_rti_solve_wpp<0, 0, B2_WPP_ALL_SHARED>;
            h->wpp_smem_small = b2_wpp_smem_bytes(N, M, B2_WPP_ALL_SHARED);
            int occ_s = 0;
            if (cudaFuncSetAttribute(h->wpp_fn_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->wpp_smem_small) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, h->wpp_fn_small, 32, h->wpp_smem_small) != cudaSuccess || occ_s < 1)
                h->wpp_fn_small = nullptr;
            h->small_batch_max = occ_s * sms;
        }
    }
    CUDA_TRY(cudaDeviceSynchronize());
    *out = h;
    return B2MPC_OK;
}

extern "C" int b2mpc_destroy(b2mpc* h)
{
    if (!h) return B2MPC_OK;
    B2DeviceGuard guard(h->cfg.device);
    for (int i = 0; i < B2MPC_N_FIELDS; i++) if (h->f[i]) cudaFree(h->f[i]);
    for (int i = 0; i < B2MPC_N_IFIELDS; i++) if (h->fi[i]) cudaFree(h->fi[i]);
    if (h->wpriv) cudaFree(h->wpriv);
    if (h->order) cudaFree(h->order);
    if (h->stage_d) cudaFree(h->stage_d);
    delete h;
    return B2MPC_OK;
}

extern "C" int64_t b2mpc_field_len(const b2mpc* h, int field)
{
    if (!h || field < 0 || field >= B2MPC_N_FIELDS) return -1;
    return h->len[field];
}
extern "C" int64_t b2mpc_workspace_bytes(const b2mpc* h) { return h ? (int64_t)h->wpriv_bytes : -1; }
extern "C" int64_t b2mpc_launch_count(const b2mpc* h) { return h ? h->launches : -1; }
extern "C" int b2mpc_kernel(const b2mpc* h) { return h ? h->kernel : -1; }
extern "C" double* b2mpc_field_ptr(b2mpc* h, int field)
{
    if (!h || field < 0 || field >= B2MPC_N_FIELDS) return nullptr;
    return h->f[field];
}

// ------------------------------------------------------------------ layout conversion kernels
// dst[len][Bm] (SoA) <- src[batch][len] (AoS), tiled through shared memory so both sides coalesce
__global__ void b2_aos_to_soa(const double* __restrict__ src, double* __restrict__ dst, int len, int batch, int Bm)
{
    __shared__ double tile[32][33];
    const int b0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r, l = l0 + threadIdx.x;
        if (b < batch && l < len) tile[r][threadIdx.x] = src[(size_t)b * len + l];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int l = l0 + r, b = b0 + threadIdx.x;
        if (b < batch && l < len) dst[(size_t)l * Bm + b] = tile[threadIdx.x][r];
    }
}
__global__ void b2_soa_to_aos(const double* __restrict__ src, double* __restrict__ dst, int len, int batch, int Bm)
{
    __shared__ double tile[32][33];
    const int b0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int l = l0 + r, b = b0 + threadIdx.x;
        if (b < batch && l < len) tile[r][threadIdx.x] = src[(size_t)l * Bm + b];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r, l = l0 + threadIdx.x;
        if (b < batch && l < len) dst[(size_t)b * len + l] = tile[threadIdx.x][r];
    }
}

static int check_batch(const b2mpc* h, int64_t batch)
{
    if (!h) return fail(B2MPC_EINVAL, "null handle");
    if (batch < 1 || batch > h->Bm) return fail(B2MPC_EBATCH, "batch %lld out of range [1,%d]", (long long)batch, h->Bm);
    return B2MPC_OK;
}

extern "C" int b2mpc_set(b2mpc* h, int field, const double* dptr, int64_t batch, int layout, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if (field < 0 || field >= B2MPC_N_FIELDS || !dptr) return fail(B2MPC_EINVAL, "bad field %d or null pointer", field);
    cudaStream_t st = (cudaStream_t)stream;
    const int len = (int)h->len[field];
    if (layout == B2MPC_LAYOUT_SOA) {
        CUDA_TRY(cudaMemcpy2DAsync(h->f[field], sizeof(double) * h->Bm, dptr, sizeof(double) * batch,
                                   sizeof(double) * batch, len, cudaMemcpyDeviceToDevice, st));
    } else if (layout == B2MPC_LAYOUT_AOS) {
        dim3 grid((unsigned)((batch + 31) / 32), (unsigned)((len + 31) / 32)), block(32, 8);
        b2_aos_to_soa<<<grid, block, 0, st>>>(dptr, h->f[field], len, (int)batch, h->Bm);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    } else return fail(B2MPC_EINVAL, "bad layout %d", layout);
    return B2MPC_OK;
}

extern "C" int b2mpc_get(b2mpc* h, int field, double* dptr, int64_t batch, int layout, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if (field < 0 || field >= B2MPC_N_FIELDS || !dptr) return fail(B2MPC_EINVAL, "bad field %d or null pointer", field);
    cudaStream_t st = (cudaStream_t)stream;
    const int len = (int)h->len[field];
    if (layout == B2MPC_LAYOUT_SOA) {
        CUDA_TRY(cudaMemcpy2DAsync(dptr, sizeof(double) * batch, h->f[field], sizeof(double) * h->Bm,
                                   sizeof(double) * batch, len, cudaMemcpyDeviceToDevice, st));
    } else if (layout == B2MPC_LAYOUT_AOS) {
        dim3 grid((unsigned)((batch + 31) / 32), (unsigned)((len + 31) / 32)), block(32, 8);
        b2_soa_to_aos<<<grid, block, 0, st>>>(h->f[field], dptr, len, (int)batch, h->Bm);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    } else return fail(B2MPC_EINVAL, "bad layout %d", layout);
    return B2MPC_OK;
}

extern "C" int b2mpc_get_i32(b2mpc* h, int ifield, int32_t* dptr, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if (ifield < 0 || ifield >= B2MPC_N_IFIELDS || !dptr) return fail(B2MPC_EINVAL, "bad int field %d or null pointer", ifield);
    CUDA_TRY(cudaMemcpyAsync(dptr, h->fi[ifield], sizeof(int32_t) * batch, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B2MPC_OK;
}

// host variants: batch-first host memory <-> SoA device fields.  The copy goes straight between the caller's buffer and
// the handle's DEVICE staging buffer (cudaMemcpyAsync on `stream`), then a layout kernel.  b2mpc_set_host does not
// synchronise: with pinned host memory the copy is truly asynchronous and the caller must keep the buffer alive and
// unchanged until the stream has passed it (pageable memory is staged by the driver before the call returns).  The
// get_* variants synchronise the stream before returning.
extern "C" int b2mpc_set_host(b2mpc* h, int field, const double* hptr, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if (field < 0 || field >= B2MPC_N_FIELDS || !hptr) return fail(B2MPC_EINVAL, "bad field %d or null pointer", field);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->len[field] * batch;
    CUDA_TRY(cudaMemcpyAsync(h->stage_d, hptr, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    return b2mpc_set(h, field, h->stage_d, batch, B2MPC_LAYOUT_AOS, stream);
}
extern "C" int b2mpc_get_host(b2mpc* h, int field, double* hptr, int64_t batch, void* stream)
{
    int rc = b2mpc_get(h, field, h ? h->stage_d : nullptr, batch, B2MPC_LAYOUT_AOS, stream);
    if (rc) return rc;
    if (!hptr) return fail(B2MPC_EINVAL, "null pointer");
    B2DeviceGuard guard(h->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->len[field] * batch;
    CUDA_TRY(cudaMemcpyAsync(hptr, h->stage_d, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2MPC_OK;
}
extern "C" int b2mpc_get_i32_host(b2mpc* h, int ifield, int32_t* hptr, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if (ifield < 0 || ifield >= B2MPC_N_IFIELDS || !hptr) return fail(B2MPC_EINVAL, "bad int field %d or null pointer", ifield);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(hptr, h->fi[ifield], sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2MPC_OK;
}

// ------------------------------------------------------------------ glue kernels
// parameterize_slack(), robot_ocp_problem.py:145-152
__device__ __forceinline__ void b2_slack_one(const double* __restrict__ X0, const double* __restrict__ GOAL, double* __restrict__ ZL,
                                             double* __restrict__ ZZL, int N, int M, int b, int Bm)
{
    const double d0 = X0[0 * (size_t)Bm + b] - GOAL[0 * (size_t)Bm + b], d1 = X0[1 * (size_t)Bm + b] - GOAL[1 * (size_t)Bm + b];
    const double v = X0[3 * (size_t)Bm + b], w = X0[4 * (size_t)Bm + b];
    const double scale = 1e4 * (d0 * d0 + d1 * d1 + v * v + w * w + 50.0);
    for (int i = 0; i <= N; i++) {
        const double a = scale * (N - i) / N;
        for (int j = 0; j < M; j++) {
            ZL[((size_t)i * M + j) * Bm + b] = a;
            ZZL[((size_t)i * M + j) * Bm + b] = a;
        }
    }
}
__global__ void b2_prepare_slack(const double* __restrict__ X0, const double* __restrict__ GOAL, double* __restrict__ ZL,
                                 double* __restrict__ ZZL, int N, int M, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    b2_slack_one(X0, GOAL, ZL, ZZL, N, M, b, Bm);
}
// ocp_solver.reset() + set_initial_guess(), robot_ocp_problem.py:286-306 (incl. aliasing quirk :301-302)
__global__ void b2_reset(double* __restrict__ X0, double* __restrict__ X, double* __restrict__ U, double* __restrict__ SL,
                         double* __restrict__ SU, double* __restrict__ PI, double* __restrict__ LAM, double* __restrict__ T,
                         const uint8_t* __restrict__ mask, int N, int M, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (mask && !mask[b]) return;
    X0[3 * (size_t)Bm + b] = 0.0; X0[4 * (size_t)Bm + b] = 0.0;
    double x0[5];
    for (int i = 0; i < 5; i++) x0[i] = X0[i * (size_t)Bm + b];
    for (int k = 0; k <= N; k++) {
        for (int i = 0; i < 5; i++) X[((size_t)k * 5 + i) * Bm + b] = x0[i];
        if (k < N) { U[((size_t)k * 2) * Bm + b] = 0.0; U[((size_t)k * 2 + 1) * Bm + b] = 0.0; }
        for (int j = 0; j < M; j++) { SL[((size_t)k * M + j) * Bm + b] = 0.0; SU[((size_t)k * M + j) * Bm + b] = 0.0; }
        // reset() also forgets the multipliers (they only matter to the full-SQP mode's first convergence test)
        if (k < N) for (int i = 0; i < 5; i++) PI[((size_t)k * 5 + i) * Bm + b] = 0.0;
        const int nct = 12 + 4 * M;
        for (int i = 0; i < nct; i++) { LAM[((size_t)k * nct + i) * Bm + b] = 0.0; T[((size_t)k * nct + i) * Bm + b] = 0.0; }
    }
}
// ocp_integrator: X0 <- F(X0, U[0]), robot_ocp_problem.py:207-212 (same IRK as the OCP, :136)
__global__ void b2_sim(double* __restrict__ X0, const double* __restrict__ U, double dt, int integrator, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double x[5], u[2], xn[5], sv[10];
    for (int i = 0; i < 5; i++) x[i] = X0[i * (size_t)Bm + b];
    u[0] = U[b]; u[1] = U[(size_t)Bm + b];
    if (integrator == 1) model_rk4_step(x, u, dt, xn, sv); else model_gl4_step(x, u, dt, xn, sv);
    for (int i = 0; i < 5; i++) X0[i * (size_t)Bm + b] = xn[i];
}
// warm-start shift, robot_ocp_problem.py:253-258
__global__ void b2_shift(double* __restrict__ X, double* __restrict__ U, int N, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int j = 0; j < N - 1; j++) {
        for (int i = 0; i < 5; i++) X[((size_t)j * 5 + i) * Bm + b] = X[((size_t)(j + 1) * 5 + i) * Bm + b];
        for (int i = 0; i < 2; i++) U[((size_t)j * 2 + i) * Bm + b] = U[((size_t)(j + 1) * 2 + i) * Bm + b];
    }
    for (int i = 0; i < 5; i++) X[((size_t)(N - 1) * 5 + i) * Bm + b] = X[((size_t)N * 5 + i) * Bm + b];
    U[((size_t)(N - 1) * 2) * Bm + b] = 0.0; U[((size_t)(N - 1) * 2 + 1) * Bm + b] = 0.0;
}

extern "C" int b2mpc_prepare_slack(b2mpc* h, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_prepare_slack<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->f[B2MPC_X0], h->f[B2MPC_GOAL], h->f[B2MPC_ZL], h->f[B2MPC_ZZL], h->N, h->M, (int)batch, h->Bm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_reset(b2mpc* h, const uint8_t* dmask, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_reset<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->f[B2MPC_X0], h->f[B2MPC_X], h->f[B2MPC_U], h->f[B2MPC_SL], h->f[B2MPC_SU], h->f[B2MPC_PI], h->f[B2MPC_LAM], h->f[B2MPC_T],
        dmask, h->N, h->M, (int)batch, h->Bm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_sim(b2mpc* h, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_sim<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->f[B2MPC_X0], h->f[B2MPC_U], h->prm.dt, h->cfg.integrator, (int)batch, h->Bm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_shift(b2mpc* h, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_shift<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->f[B2MPC_X], h->f[B2MPC_U], h->N, (int)batch, h->Bm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}

// ------------------------------------------------------------------ closed loop around the solve
// world_specification.py:7-10,28-31 (arena, obstacle radius/noise/speed limit)
#define B2_XMIN (-8.0)
#define B2_XMAX (8.0)
#define B2_YMIN (-8.0)
#define B2_YMAX (8.0)
#define B2_R_OBST (1.0)
#define B2_R_ROBOT (0.2)
#define B2_RANDOMNESS (0.1)
#define B2_V_MAX_OBST (2.0)

// Obstacle.predict_step without noise, visualization.py:35-59
__device__ __forceinline__ void b2_obst_move(double& x, double& vx, double& y, double& vy, double dt)
{
    double th;
    if (vx < 0) th = (x - B2_XMIN) / fabs(vx);
    else if (vx > 0) th = (B2_XMAX - x) / fabs(vx);
    else th = INFINITY;
    if (th <= dt) { x += (vx * th - vx * (dt - th)); vx = -vx; }
    else x += vx * dt;
    if (vy < 0) th = (y - B2_YMIN) / fabs(vy);
    else if (vy > 0) th = (B2_YMAX - y) / fabs(vy);
    else th = INFINITY;
    if (th <= dt) { y += (vy * th - vy * (dt - th)); vy = -vy; }
    else y += vy * dt;
}

// parameterize_model(): one thread per (problem, obstacle), robot_ocp_problem.py:154-166
__device__ __forceinline__ void b2_predict_one(const double* __restrict__ OBST, double* __restrict__ P, int N, int M, double dt,
                                               int quirk, int b, int j, int Bm)
{
    double x = OBST[((size_t)j * 4 + 0) * Bm + b], y = OBST[((size_t)j * 4 + 1) * Bm + b];
    double vy = OBST[((size_t)j * 4 + 3) * Bm + b];
    double vx = quirk ? vy : OBST[((size_t)j * 4 + 2) * Bm + b];        // `vx = self.vy`, visualization.py:69
    P[((size_t)0 * 2 * M + 2 * j) * Bm + b] = x; P[((size_t)0 * 2 * M + 2 * j + 1) * Bm + b] = y;
    for (int i = 0; i < N; i++) {
        b2_obst_move(x, vx, y, vy, dt);
        P[((size_t)(i + 1) * 2 * M + 2 * j) * Bm + b] = x; P[((size_t)(i + 1) * 2 * M + 2 * j + 1) * Bm + b] = y;
    }
}
__global__ void b2_obst_predict(const double* __restrict__ OBST, double* __restrict__ P, const int32_t* __restrict__ DONE,
                                int N, int M, double dt, int quirk, int B, int Bm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t % B, j = t / B;
    if (j >= M) return;
    if (DONE[b]) return;
    b2_predict_one(OBST, P, N, M, dt, quirk, b, j, Bm);
}

__global__ void b2_loop_reset(double* __restrict__ MARGIN, int32_t* STEPS, int32_t* DONE, int32_t* OOB, int32_t* NFAIL, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    MARGIN[b] = INFINITY; MARGIN[(size_t)Bm + b] = INFINITY;
    STEPS[b] = 0; DONE[b] = 0; OOB[b] = 0; NFAIL[b] = 0;
}

// head of RobotOcpProblem.step() (robot_ocp_problem.py:177-183): reached_subgoal = False, out_of_bounds = False, i = 0;
// min_margin_traj belongs to the experiment (init_experiment, :51) and survives
__global__ void b2_loop_begin(int32_t* STEPS, int32_t* DONE, int32_t* OOB, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    STEPS[b] = 0; DONE[b] = 0; OOB[b] = 0;
}

// ---- Philox4x32-10 (Salmon et al., SC'11: the counter-based generator behind cuRAND's and numpy's Philox), written out so
// that the velocity-noise stream of an environment (visualization.py:28-33: np.random.normal per obstacle and control step)
// depends only on (seed, GLOBAL environment index, control step, obstacle) -- never on the batch or the shard it runs in.
__host__ __device__ __forceinline__ void b2_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two independent N(0,1) draws for (environment, control step, obstacle): Box-Muller on two 53-bit uniforms in (0,1)
__device__ __forceinline__ void b2_philox_normal2(uint64_t seed, uint64_t env, uint32_t step, uint32_t obstacle, double* n0, double* n1)
{
    uint32_t r[4];
    b2_philox4x32_10(step, obstacle, (uint32_t)env, (uint32_t)(env >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    *n0 = rad * cs; *n1 = rad * sn;
}

// where the velocity noise of the closed loop comes from
struct B2Noise {
    int mode;                 // 0 none, 1 device Philox stream, 2 replay of a caller-provided stream
    const double* replay;     // mode 2: [batch][M][2] for THIS control step (batch-first)
    uint64_t seed;            // mode 1
    int64_t env_offset;       //         global index of environment 0 of this batch (shards draw the unsharded run's streams)
    uint32_t step;            //         control-step counter of the experiment
};

// robot_ocp_problem.py:198-260 for one problem
__device__ __forceinline__ void b2_post_one(const B2Params& p, int b, double* __restrict__ X0w, const double* __restrict__ GOAL,
                                            double* __restrict__ OBST, double* __restrict__ MARGIN, double* __restrict__ U0,
                                            int32_t* STEPS, int32_t* DONE, int32_t* OOB, int32_t* NFAIL, const B2Noise& nz,
                                            int random_move, int init_guess_when_error, double tol)
{
    const size_t Bm = (size_t)p.Bm;
    const int N = p.N, M = p.M;
    // u = ocp_solver.get(0,'u') (:198): on status 4 acados left the iterate untouched -> stale u0
    double u[2] = {p.U[b], p.U[Bm + b]};
    U0[b] = u[0]; U0[Bm + b] = u[1];
    double x[5];
    for (int i = 0; i < 5; i++) x[i] = X0w[i * Bm + b];
    if (p.STATUS[b] == 4) {
        NFAIL[b] += 1;
        if (init_guess_when_error) {
            // set_initial_guess(): reset + x_i <- x0 with v,omega zeroed IN x0 ITSELF (:301-302)
            x[3] = 0.0; x[4] = 0.0;
            for (int k = 0; k <= N; k++) {
                for (int i = 0; i < 5; i++) p.X[((size_t)k * 5 + i) * Bm + b] = x[i];
                if (k < N) { p.U[((size_t)k * 2) * Bm + b] = 0.0; p.U[((size_t)k * 2 + 1) * Bm + b] = 0.0; }
                for (int j = 0; j < M; j++) { p.SL[((size_t)k * M + j) * Bm + b] = 0.0; p.SU[((size_t)k * M + j) * Bm + b] = 0.0; }
            }
        }
    }
    // plant step (:207-212)
    double xn[5], sv[10];
    if (p.integrator == 1) model_rk4_step(x, u, p.dt, xn, sv); else model_gl4_step(x, u, p.dt, xn, sv);
    for (int i = 0; i < 5; i++) X0w[i * Bm + b] = xn[i];
    if (xn[0] < B2_XMIN || xn[0] > B2_XMAX || xn[1] < B2_YMIN || xn[1] > B2_YMAX) OOB[b] = 1;   // :213-214
    // obstacles move (:217-218) and margin bookkeeping (:222-228)
    double min_margin = INFINITY;
    for (int j = 0; j < M; j++) {
        double ox = OBST[((size_t)j * 4 + 0) * Bm + b], oy = OBST[((size_t)j * 4 + 1) * Bm + b];
        double vx = OBST[((size_t)j * 4 + 2) * Bm + b], vy = OBST[((size_t)j * 4 + 3) * Bm + b];
        if (random_move && nz.mode) {
            double n0, n1;
            if (nz.mode == 2) { n0 = nz.replay[((size_t)b * M + j) * 2 + 0]; n1 = nz.replay[((size_t)b * M + j) * 2 + 1]; }
            else b2_philox_normal2(nz.seed, (uint64_t)(nz.env_offset + b), nz.step, (uint32_t)j, &n0, &n1);
            vx = fmin(fmax((1.0 + B2_RANDOMNESS * n0) * vx, -B2_V_MAX_OBST), B2_V_MAX_OBST);   // visualization.py:28-33
            vy = fmin(fmax((1.0 + B2_RANDOMNESS * n1) * vy, -B2_V_MAX_OBST), B2_V_MAX_OBST);
        }
        b2_obst_move(ox, vx, oy, vy, p.dt);
        OBST[((size_t)j * 4 + 0) * Bm + b] = ox; OBST[((size_t)j * 4 + 1) * Bm + b] = oy;
        OBST[((size_t)j * 4 + 2) * Bm + b] = vx; OBST[((size_t)j * 4 + 3) * Bm + b] = vy;
        const double margin = sqrt((xn[0] - ox) * (xn[0] - ox) + (xn[1] - oy) * (xn[1] - oy)) - (B2_R_OBST + B2_R_ROBOT);
        if (margin < min_margin) min_margin = margin;
    }
    if (min_margin < MARGIN[b]) MARGIN[b] = min_margin;
    // goal test on the position only (:246-250); the loop `break`s BEFORE i += 1 and before the shift
    const double gx = GOAL[b], gy = GOAL[Bm + b];
    const double dist = sqrt((xn[0] - gx) * (xn[0] - gx) + (xn[1] - gy) * (xn[1] - gy));
    MARGIN[Bm + b] = dist;
    if (dist <= tol) { DONE[b] = 1; return; }
    // warm-start shift (:253-258)
    for (int j = 0; j < N - 1; j++) {
        for (int i = 0; i < 5; i++) p.X[((size_t)j * 5 + i) * Bm + b] = p.X[((size_t)(j + 1) * 5 + i) * Bm + b];
        for (int i = 0; i < 2; i++) p.U[((size_t)j * 2 + i) * Bm + b] = p.U[((size_t)(j + 1) * 2 + i) * Bm + b];
    }
    for (int i = 0; i < 5; i++) p.X[((size_t)(N - 1) * 5 + i) * Bm + b] = p.X[((size_t)N * 5 + i) * Bm + b];
    p.U[((size_t)(N - 1) * 2) * Bm + b] = 0.0; p.U[((size_t)(N - 1) * 2 + 1) * Bm + b] = 0.0;
    STEPS[b] += 1;
}
__global__ void b2_loop_post(B2Params p, double* __restrict__ X0w, const double* __restrict__ GOAL, double* __restrict__ OBST,
                             double* __restrict__ MARGIN, double* __restrict__ U0, int32_t* STEPS, int32_t* DONE, int32_t* OOB,
                             int32_t* NFAIL, B2Noise nz, int random_move, int init_guess_when_error, double tol)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    if (DONE[b]) return;
    b2_post_one(p, b, X0w, GOAL, OBST, MARGIN, U0, STEPS, DONE, OOB, NFAIL, nz, random_move, init_guess_when_error, tol);
}


extern "C" int b2mpc_obst_predict(b2mpc* h, int64_t batch, int32_t quirk, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    const int64_t n = batch * h->M;
    b2_obst_predict<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->f[B2MPC_OBST], h->f[B2MPC_P], h->fi[B2MPC_DONE], h->N, h->M, h->prm.dt, quirk, (int)batch, h->Bm);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_loop_reset(b2mpc* h, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_loop_reset<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->f[B2MPC_MARGIN], h->fi[B2MPC_STEPS], h->fi[B2MPC_DONE], h->fi[B2MPC_OOB], h->fi[B2MPC_NFAIL], (int)batch, h->Bm);
    h->loop_step = 0;
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_loop_begin(b2mpc* h, int64_t batch, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    b2_loop_begin<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        h->fi[B2MPC_STEPS], h->fi[B2MPC_DONE], h->fi[B2MPC_OOB], (int)batch);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_loop_post(b2mpc* h, int64_t batch, const double* dnoise, int32_t random_move,
                               int32_t init_guess_when_error, double tol, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    B2Params p = h->prm;
    p.B = (int)batch;
    b2_loop_post<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        p, h->f[B2MPC_X0], h->f[B2MPC_GOAL], h->f[B2MPC_OBST], h->f[B2MPC_MARGIN], h->f[B2MPC_U0], h->fi[B2MPC_STEPS],
        h->fi[B2MPC_DONE], h->fi[B2MPC_OOB], h->fi[B2MPC_NFAIL], B2Noise{dnoise ? 2 : 0, dnoise, 0, 0, 0}, random_move,
        init_guess_when_error, tol);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}

// Launch-order heuristic for the warp kernel.  Interior-point iteration counts range from 13 to the limit (50);
// the persistent CTAs pull problems from a queue, so the order of the queue decides the tail of the launch.
// Predictor (4096 random scenarios, profiles/experiments/r01_launch_order.md): the FIRST shooting node whose
// linearisation point lies inside an obstacle's safety disc -- node 0/1: 45 iterations on average (62 % hit the limit),
// node 10: 30, node 20: 15; none: 15-20 depending on the smallest clearance.  Spearman correlation with the iteration
// count 0.84.  Counting sort by that key, expensive first.  Pure scheduling: per-problem results are unaffected.
#define B2_ORDER_NODEKEYS 64
#define B2_ORDER_NKEYS (B2_ORDER_NODEKEYS + 3)
__device__ __forceinline__ int b2_order_key_one(const double* __restrict__ X, const double* __restrict__ P, int N, int M, double r2, int b, int Bm)
{
    int first = -1;
    double hmin = 1e300;
    for (int k = 0; k <= N && first < 0; k++) {
        const double x = X[((size_t)k * 5 + 0) * Bm + b], y = X[((size_t)k * 5 + 1) * Bm + b];
        for (int j = 0; j < M; j++) {
            const double dx = x - P[((size_t)k * 2 * M + 2 * j) * Bm + b], dy = y - P[((size_t)k * 2 * M + 2 * j + 1) * Bm + b];
            const double h = dx * dx + dy * dy - r2;
            hmin = h < hmin ? h : hmin;
            if (h < 0.0) first = k;
        }
    }
    return first >= 0 ? (first < B2_ORDER_NODEKEYS ? first : B2_ORDER_NODEKEYS - 1)
                      : B2_ORDER_NODEKEYS + (hmin < 1.0 ? 0 : (hmin < 3.0 ? 1 : 2));
}
__global__ void b2_order_keys(const double* __restrict__ X, const double* __restrict__ P, int32_t* __restrict__ key,
                              int32_t* __restrict__ hist, int N, int M, double r2, int B, int Bm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kk = b2_order_key_one(X, P, N, M, r2, b, Bm);
    key[b] = kk;
    atomicAdd(&hist[kk], 1);
}
__global__ void b2_order_keys_prev(const int32_t* __restrict__ it, int32_t* __restrict__ key, int32_t* __restrict__ hist, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int i = it[b];
    const int kk = B2_ORDER_NODEKEYS - 1 - (i < B2_ORDER_NODEKEYS ? (i > 0 ? i : 0) : B2_ORDER_NODEKEYS - 1);
    key[b] = kk;
    atomicAdd(&hist[kk], 1);
}
// key < 0: the problem is not queued (closed loop: an environment that has reached its goal).  The first thread also
// prepares what the persistent solve kernel and the NEXT scatter need: work counter, queue length, and -- the key
// histogram being double-buffered between control steps -- zeroes of the other buffer.
__global__ void b2_order_scatter(const int32_t* __restrict__ key, const int32_t* __restrict__ hist, int32_t* __restrict__ fill,
                                 int32_t* __restrict__ order, int nkeys, int B, int32_t* next, int32_t* nqueue,
                                 int32_t* other_hist, int32_t* other_fill)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) {
        if (next) *next = 0;
        if (nqueue) { int tot = 0; for (int c = 0; c < nkeys; c++) tot += hist[c]; *nqueue = tot; }
    }
    if (other_hist && b < nkeys) { other_hist[b] = 0; other_fill[b] = 0; }
    if (b >= B) return;
    const int kk = key[b];
    if (kk < 0) return;
    int start = 0;
    for (int c = 0; c < kk; c++) start += hist[c];
    order[start + atomicAdd(&fill[kk], 1)] = b;
}

// ---- the fused closed loop (b2mpc_closed_loop): everything of robot_ocp_problem.py:184-260 that is not the solve, for one
// environment per thread, in ONE kernel between two solves: the rest of control step i (b2_post_one: u0, failure handling,
// plant step, obstacle step with noise, margins, goal test, shift) and the head of step i+1 (parameterize_model,
// parameterize_slack, queue key).  Environments that reached their goal leave the queue: the solve kernel only sees the
// active ones.  Queue key of a warm-started problem: the interior-point iteration count of its previous solve (already on
// the device; expensive first), of a fresh one: the geometric predictor of b2_order_key_one.
struct B2LoopArgs {
    double *X0w, *GOAL, *OBST, *MARGIN, *U0;
    int32_t *STEPS, *DONE, *OOB, *NFAIL;
    B2Noise nz;
    int random_move, init_guess_when_error, quirk, do_post, do_pre;
    int key_mode;             // 0: previous iteration count where available, 1: always the geometric predictor (experiment: B2MPC_LOOP_KEY=geo)
    double tol;
    int32_t *key, *hist;
};
__global__ void b2_loop_glue(B2Params p, B2LoopArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    bool fresh = !a.do_post;
    if (a.do_post && !a.DONE[b]) {
        fresh = p.STATUS[b] == 4 && a.init_guess_when_error;       // re-initialised iterate: the previous count says nothing
        b2_post_one(p, b, a.X0w, a.GOAL, a.OBST, a.MARGIN, a.U0, a.STEPS, a.DONE, a.OOB, a.NFAIL, a.nz, a.random_move,
                    a.init_guess_when_error, a.tol);
    }
    if (!a.do_pre) return;
    if (a.DONE[b]) { a.key[b] = -1; return; }
    for (int j = 0; j < p.M; j++) b2_predict_one(a.OBST, const_cast<double*>(p.P), p.N, p.M, p.dt, a.quirk, b, j, p.Bm);
    b2_slack_one(a.X0w, a.GOAL, const_cast<double*>(p.ZL), const_cast<double*>(p.ZZL), p.N, p.M, b, p.Bm);
    const int it = p.QP_ITER[b];
    int kk;
    if (!fresh && it > 0 && a.key_mode == 0) kk = B2_ORDER_NODEKEYS - 1 - (it < B2_ORDER_NODEKEYS ? it : B2_ORDER_NODEKEYS - 1);
    else kk = b2_order_key_one(p.X, p.P, p.N, p.M, p.r2, b, p.Bm);
    a.key[b] = kk;
    atomicAdd(&a.hist[kk], 1);
}

// ------------------------------------------------------------------ the solve
// order buffer of a handle: [order (Bm) | key (Bm) | hist0 | fill0 | hist1 | fill1 (MAXKEYS each) | work counter | queue length]
struct B2Order {
    int32_t *order, *key, *hist[2], *fill[2], *next, *nq;
    explicit B2Order(b2mpc* h)
    {
        order = h->order; key = order + h->Bm;
        hist[0] = key + h->Bm; fill[0] = hist[0] + B2_ORDER_MAXKEYS;
        hist[1] = fill[0] + B2_ORDER_MAXKEYS; fill[1] = hist[1] + B2_ORDER_MAXKEYS;
        next = fill[1] + B2_ORDER_MAXKEYS; nq = next + 1;
    }
};

// launch the solve kernel on a prepared queue (p.ORDER / p.NEXT / p.NQ set by the caller)
static int launch_solve(b2mpc* h, B2Params& p, int64_t batch, bool sqp, cudaStream_t st)
{
    if (sqp) {
        const unsigned blocks = (unsigned)(batch < h->wpp_grid ? batch : h->wpp_grid);
        p.wpriv_stride = h->wpriv_stride_sqp;
        h->wpp_fn_sqp<<<blocks, 32, h->wpp_smem_sqp, st>>>(p);
    } else if (h->kernel == 4) {
        const unsigned blocks = (unsigned)(batch < h->cta_grid ? batch : h->cta_grid);      // persistent: one CTA per resident slot
        h->cta_fn<<<blocks, h->cta_threads, h->cta_smem, st>>>(p);
    } else if (h->wpp_fn_small && batch <= h->small_batch_max) {
        h->wpp_fn_small<<<(unsigned)batch, 32, h->wpp_smem_small, st>>>(p);            // one wave, everything on chip
    } else {
        const unsigned blocks = (unsigned)(batch < h->wpp_grid ? batch : h->wpp_grid);   // persistent: one warp-CTA per resident slot
        if (h->l2_persist && h->wpriv_bytes) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(blocks); lc.blockDim = dim3(32); lc.dynamicSmemBytes = h->wpp_smem; lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeAccessPolicyWindow;
            at[0].val.accessPolicyWindow.base_ptr = h->wpriv;
            at[0].val.accessPolicyWindow.num_bytes = sizeof(double) * h->prm.wpriv_stride * (size_t)blocks;
            at[0].val.accessPolicyWindow.hitRatio = h->l2_hit_ratio;
            at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            lc.attrs = at; lc.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&lc, h->wpp_fn, (const B2Params)p));
        } else
        h->wpp_fn<<<blocks, 32, h->wpp_smem, st>>>(p);
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}

static int solve_impl(b2mpc* h, int64_t batch, bool sqp, int32_t max_iter, const double* tol4, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    B2DeviceGuard guard(h->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    B2Params p = h->prm;
    p.B = (int)batch;
    if (sqp) {
        if (max_iter > 0) p.nlp_iter_max = max_iter;
        if (tol4) for (int i = 0; i < 4; i++) p.nlp_tol[i] = tol4[i];
    }
    B2Order o(h);
    if (batch == 1) {
        // one problem: nothing to order -- the work counter is all the solve kernel needs (two launches less on the
        // single-episode latency path)
        CUDA_TRY(cudaMemsetAsync(o.next, 0, 2 * sizeof(int32_t), st));
        p.ORDER = nullptr; p.NEXT = o.next; p.NQ = nullptr;
        return launch_solve(h, p, batch, sqp, st);
    }
    CUDA_TRY(cudaMemsetAsync(o.hist[0], 0, (4 * B2_ORDER_MAXKEYS + 8) * sizeof(int32_t), st));
    const unsigned ob = (unsigned)((batch + 127) / 128);
    static const int key_prev = []() { const char* e = getenv("B2MPC_SOLVE_KEY"); return (e && e[0] == 'p') ? 1 : 0; }();
    if (key_prev)      // experiment (B2MPC_SOLVE_KEY=prev): the iteration counts of the PREVIOUS solve as queue keys -- with the same
                       // problems solved twice from the same iterate this is the perfect cost predictor (upper bound of what a
                       // better launch order can recover of the drain tail)
        b2_order_keys_prev<<<ob, 128, 0, st>>>(h->fi[B2MPC_QP_ITER], o.key, o.hist[0], (int)batch);
    else
    b2_order_keys<<<ob, 128, 0, st>>>(h->f[B2MPC_X], h->f[B2MPC_P], o.key, o.hist[0], h->N, h->M, h->cfg.r2, (int)batch, h->Bm);
    b2_order_scatter<<<ob, 128, 0, st>>>(o.key, o.hist[0], o.fill[0], o.order, B2_ORDER_NKEYS, (int)batch, nullptr, nullptr, nullptr, nullptr);
    h->launches += 2;
    p.ORDER = o.order;
    p.NEXT = o.next;
    p.NQ = nullptr;
    return launch_solve(h, p, batch, sqp, st);
}

// RobotOcpProblem.step()'s `while i < max_iter` loop (robot_ocp_problem.py:184-260) for `steps` control steps, enqueued by ONE
// call: per control step three launches -- queue scatter, solve, glue (rest of this step + head of the next) -- instead of
// six launches and a trip through the host language.
extern "C" int b2mpc_closed_loop(b2mpc* h, int64_t batch, int32_t steps, int32_t noise_mode, uint64_t seed, int64_t env_offset,
                                 const double* dnoise, int32_t random_move, int32_t init_guess_when_error, int32_t quirk_vx_is_vy,
                                 double tol, void* stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    if (steps < 0) return fail(B2MPC_EINVAL, "steps must be >= 0");
    if (noise_mode < 0 || noise_mode > 2) return fail(B2MPC_EINVAL, "noise_mode must be 0 (none), 1 (device Philox stream) or 2 (replay)");
    if (noise_mode == 2 && !dnoise) return fail(B2MPC_EINVAL, "noise_mode 2 needs a device noise stream [steps][batch][M][2]");
    if (steps == 0) return B2MPC_OK;
    B2DeviceGuard guard(h->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    B2Params p = h->prm;
    p.B = (int)batch;
    B2Order o(h);
    CUDA_TRY(cudaMemsetAsync(o.hist[0], 0, (4 * B2_ORDER_MAXKEYS + 8) * sizeof(int32_t), st));
    B2LoopArgs a;
    a.X0w = h->f[B2MPC_X0]; a.GOAL = h->f[B2MPC_GOAL]; a.OBST = h->f[B2MPC_OBST]; a.MARGIN = h->f[B2MPC_MARGIN]; a.U0 = h->f[B2MPC_U0];
    a.STEPS = h->fi[B2MPC_STEPS]; a.DONE = h->fi[B2MPC_DONE]; a.OOB = h->fi[B2MPC_OOB]; a.NFAIL = h->fi[B2MPC_NFAIL];
    a.random_move = random_move; a.init_guess_when_error = init_guess_when_error; a.quirk = quirk_vx_is_vy; a.tol = tol;
    a.key = o.key;
    { const char* e = getenv("B2MPC_LOOP_KEY"); a.key_mode = (e && e[0] == 'g') ? 1 : 0; }
    const unsigned ob = (unsigned)((batch + 127) / 128);
    const size_t noise_stride = (size_t)batch * h->M * 2;
    int cur = 0;
    // head of the first control step
    a.do_post = 0; a.do_pre = 1; a.hist = o.hist[cur];
    a.nz = B2Noise{0, nullptr, 0, 0, 0};
    b2_loop_glue<<<ob, 128, 0, st>>>(p, a);
    h->launches++;
    for (int i = 0; i < steps; i++) {
        b2_order_scatter<<<ob, 128, 0, st>>>(o.key, o.hist[cur], o.fill[cur], o.order, B2_ORDER_NKEYS, (int)batch, o.next, o.nq,
                                             o.hist[cur ^ 1], o.fill[cur ^ 1]);
        h->launches++;
        p.ORDER = o.order; p.NEXT = o.next; p.NQ = o.nq;
        if ((rc = launch_solve(h, p, batch, h->cfg.nlp_solver == 1, st))) return rc;      // (nlp_solver = 1: a full SQP solve per control step)
        cur ^= 1;
        a.do_post = 1; a.do_pre = (i + 1 < steps) ? 1 : 0; a.hist = o.hist[cur];
        a.nz = B2Noise{random_move ? noise_mode : 0, noise_mode == 2 ? dnoise + (size_t)i * noise_stride : nullptr, seed, env_offset,
                       (uint32_t)(h->loop_step + i)};
        b2_loop_glue<<<ob, 128, 0, st>>>(p, a);
        h->launches++;
    }
    h->loop_step += steps;
    CUDA_TRY(cudaGetLastError());
    return B2MPC_OK;
}
extern "C" int b2mpc_solve(b2mpc* h, int64_t batch, void* stream)
{
    if (!h) return fail(B2MPC_EINVAL, "null handle");
    return solve_impl(h, batch, h->cfg.nlp_solver == 1, 0, nullptr, stream);
}
extern "C" int b2mpc_solve_sqp(b2mpc* h, int64_t batch, int32_t max_iter, const double* tol4, void* stream)
{
    if (!h) return fail(B2MPC_EINVAL, "null handle");
    return solve_impl(h, batch, true, max_iter, tol4, stream);
}

extern "C" int b2mpc_solve_host(b2mpc* h, int64_t batch, const double* x0, const double* p, const double* yref,
                                const double* zl, const double* Zl, double* x, double* u, double* sl, double* su,
                                int32_t* status, int32_t* qp_iter, void* stream)
{
    int rc;
    if (!x0 || !p || !yref || !zl || !Zl || !x || !u || !sl || !su) return fail(B2MPC_EINVAL, "null pointer");
    if ((rc = check_batch(h, batch))) return rc;
    B2DeviceGuard guard(h->cfg.device);
    if ((rc = b2mpc_set_host(h, B2MPC_X0, x0, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_P, p, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_YREF, yref, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_ZL, zl, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_ZZL, Zl, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_X, x, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_U, u, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_SL, sl, batch, stream))) return rc;
    if ((rc = b2mpc_set_host(h, B2MPC_SU, su, batch, stream))) return rc;
    if ((rc = b2mpc_solve(h, batch, stream))) return rc;
    // results: the four iterate fields are transposed into disjoint parts of the staging buffer (it holds the largest
    // field, (N+1)(12+4M) doubles per problem, more than x+u+sl+su together) and copied back with ONE synchronisation
    cudaStream_t st = (cudaStream_t)stream;
    const int fields[4] = {B2MPC_X, B2MPC_U, B2MPC_SL, B2MPC_SU};
    double* const dst[4] = {x, u, sl, su};
    size_t off = 0;
    for (int i = 0; i < 4; i++) {
        const size_t n = (size_t)h->len[fields[i]] * batch;
        if (off + n > h->stage_elems) return fail(B2MPC_EINVAL, "staging buffer too small");
        if ((rc = b2mpc_get(h, fields[i], h->stage_d + off, batch, B2MPC_LAYOUT_AOS, stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(dst[i], h->stage_d + off, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        off += n;
    }
    if (status) CUDA_TRY(cudaMemcpyAsync(status, h->fi[B2MPC_STATUS], sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, st));
    if (qp_iter) CUDA_TRY(cudaMemcpyAsync(qp_iter, h->fi[B2MPC_QP_ITER], sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2MPC_OK;
}

// ------------------------------------------------------------------ FP64 pipe microbenchmark
__global__ void b2_dfma_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int b2mpc_dfma_peak(int device, double* tflops_out)
{
    if (!tflops_out) return fail(B2MPC_EINVAL, "null pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(B2MPC_ENODEV, "no such CUDA device %d", device);
    B2DeviceGuard guard(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        b2_dfma_kernel<<<blocks, threads>>>(out, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *tflops_out = best;
    return B2MPC_OK;
}
