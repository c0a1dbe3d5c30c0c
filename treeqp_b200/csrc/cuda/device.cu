/*
 * device.cu -- device arena, stream, and the host-side driver of the dual Newton loop
 * (treeqp_tdunes_solve, dual_Newton_tree.c:1166-1228) over the kernels in tdunes_kernels.cuh.
 *
 * The Newton loop is level-synchronous: per iteration one launch per phase and one launch per
 * tree depth for the factorisation / substitution sweeps.  All decisions (termination, descent
 * test, Armijo) are taken on the device per tree; the host only reads two counters ("trees still
 * running", "trees with a pending line-search trial") through pinned memory to know when to stop
 * launching.
 */
#include <cuda_runtime.h>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tdunes_kernels.cuh"
#include "tdunes_fused.cuh"
#include "tdunes_stage.cuh"
#include "tdunes_small.cuh"
#include "comm.cuh"

static thread_local char g_err[512] = "";
#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            snprintf(g_err, sizeof g_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            fprintf(stderr, "[treeqp_b200] CUDA error %s\n", g_err);                             \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

enum { K_INIT, K_STAGE, K_GRAD, K_REDUCE, K_DECIDE, K_HESS, K_FACTOR, K_BACK, K_DIR, K_TRIAL, K_STEP, K_FINISH, K_COMM, K_SMALL, K_NCLASS };
/* event classes carry the names of the kernels that run: uniform trees take the compile-time-dimension
 * kernels (tdunes_fused.cuh, tdunes_stage.cuh), everything else the runtime-dimension ones; the table in
 * use is that of the most recent solve of this thread */
static const char *const k_names_generic[K_NCLASS] = {"k_init", "k_stage<0>", "k_gradient", "k_chunk_reduce", "k_decide_*",
                                                      "k_hessian", "k_factor_level", "k_backsolve_level", "k_stage<2>",
                                                      "k_trial_dense+k_ls_eval", "k_lambda_step", "k_finish", "nccl+pack", "k_tdunes_small"};
static const char *const k_names_fused[K_NCLASS] = {"k_init", "k_stage_fused<0>", "k_gradient", "k_chunk_reduce", "k_decide_*",
                                                    "k_hessian", "k_build_factor", "k_backsolve_fused", "k_stage_fused<2>",
                                                    "k_ls_eval", "k_lambda_step", "k_finish", "nccl+pack", "k_tdunes_small"};
static thread_local const char *const *k_names = k_names_generic;
static int g_profile = -1;
#define TQB_MAX_EVENTS 65536

/* 0 off | 1 an event pair around EVERY launch (serialises the launches: no PDL overlap between the level
 * kernels) | 2 as 1, but one pair around each leaves-to-root / root-to-leaves sweep, whose launches stay
 * chained with programmatic dependent launch -- the sweep is timed the way it runs */
extern "C" void tqb_set_profiling(int on) { g_profile = on < 0 ? 0 : on; }
extern "C" int tqb_num_kernel_classes(void) { return K_NCLASS; }
extern "C" const char *tqb_kernel_class_name(int k) { return (k >= 0 && k < K_NCLASS) ? k_names[k] : ""; }
static int profiling_on(void)
{
    if (g_profile < 0)
    {
        const char *e = getenv("TREEQP_B200_PROFILE");
        g_profile = (e && e[0] >= '1' && e[0] <= '2') ? e[0] - '0' : 0;
    }
    return g_profile;
}

struct tqb_device
{
    cudaEvent_t *pev; /* profiling event pool (pairs) */
    unsigned char *pev_class;
    unsigned short *pev_cnt; /* launches an event pair brackets */
    unsigned short *pev_iter; /* Newton iteration an event pair belongs to (65535: before / after the loop) */
    int cur_it;
    int pev_n, pev_on;
    int in_sweep, sweep_launches, sweep_pe; /* profiling level 2: ONE event pair around a PDL-chained level sweep */
    tqb_dims dims;
    int *d_int;
    double *d_in, *d_work;
    unsigned long long *d_kkt; /* [nT] per-tree KKT residual bits (tqb_device_kkt) */
    int *d_tint;      /* per-tree ints + counters + ls history */
    double *d_tdbl;   /* per-tree doubles */
    int *h_counters;  /* pinned [4] */
    int *lvl_ptr;     /* host copy [Nh+2] */
    View v;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int wpb_stage, wpb_hess, wpb_fact, wpb_back, wpb_dense;
    size_t sm_stage, sm_hess, sm_fact, sm_back, sm_dense;
    int pw_fact, pw_back;
    long long launches;
    size_t out_doubles;
    /* fast path (uniform dimensions, see tdunes_fused.cuh) */
    int fast, f_nx, f_nu, f_pnx;
    int *lvl_md, *lvl_kp; /* [Nh] children per parent / kids are parents, per depth */
    int *chain_link;      /* [Nh] depth l continues the chains of depth l-1 (one kid per parent, same order) */
    int *lvl_udim;        /* [Nh] d when every block of the depth has dimension d and a parent state of d entries, else 0 */
    int sw_chain_dc, sw_chain_pf, sw_dense_first; /* switches read at create: TREEQP_B200_NO_CHAIN_DC / NO_CHAIN_PF (0 = off), DENSE_FIRST_FULL (-1 = by size) */
    int has_ghost;        /* sharded local forest: reduction chunks cover the owned nodes only */
    int pair_begin;       /* level descriptors from here on are non-root parents with one kid (k_stage_fused pairs them) */
    double *d_schur;
    int *d_lvl_desc;
    int *d_leaf_desc; /* LeafDesc records (tdunes_stage.cuh) */
    int nleaf, fast_stage, stage_grid, n_sm;
    int n_sm_all; /* SMs of the device (n_sm is only set on the fused stage path) */
    /* subtree sharding (tqb_device_enable_sharding) */
    tqb_comm *comm;
    int cut_depth, ncut, my_cut0, my_ncut;
    int *d_cut_nodes;
    double *d_send, *d_recv, *d_tred;
    const tqb_p2p_dev *p2p; /* peer mailboxes over NVLink; NULL: NCCL collectives */
    int norm_on_cut;        /* sharded: the termination norm rides on the cut-node exchange */
    /* small trees: the whole Newton loop in one persistent kernel, one CTA per tree (tdunes_small.cuh) */
    int affine_refresh;       /* > 0: stage QPs from the affine model, refreshed from lambda every so many iterations (k_step_affine) */
    int *d_flags;             /* chg | bdirty | breg per node (re-use of unchanged blocks) */
    int *d_ghost;
    int skip_on;
    char *d_abstage;          /* staging of the strided [A | B] DMA copy (tqb_device_upload_ab_direct) */
    size_t abstage_bytes;
    int small, small_warps, small_mdmax;
    size_t small_smem;
    SmallPlan small_plan;
};

extern "C" const char *tqb_cuda_last_error(void) { return g_err; }

/* event pair `pe` brackets a launch of class `cls` of the Newton iteration being enqueued */
static inline void pev_tag(tqb_device *D, int pe, int cls)
{
    D->pev_class[pe] = (unsigned char) cls;
    D->pev_iter[pe] = (unsigned short) (D->cur_it < 65535 ? D->cur_it : 65535);
}

#ifdef TQB_BF_TIMING
/* debug builds only: the phase stamps of k_build_factor (tdunes_fused.cuh), 32 slots x 16 values */
extern "C" int tqb_debug_bf_timing(long long *out)
{
    return cudaMemcpyFromSymbol(out, g_bf_t, sizeof(long long) * 32 * 16) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int tqb_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int tqb_cuda_set_device(int dev) { return cudaSetDevice(dev) == cudaSuccess ? 0 : 1; }

extern "C" void *tqb_pinned_alloc(size_t bytes)
{
    void *p = NULL;
    if (bytes == 0) bytes = 8;
    if (cudaMallocHost(&p, bytes) != cudaSuccess)
    {
        snprintf(g_err, sizeof g_err, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        fprintf(stderr, "[treeqp_b200] %s\n", g_err);
        return NULL;
    }
    return p;
}

extern "C" void tqb_pinned_free(void *p) { if (p) cudaFreeHost(p); }

static int pick_wpb(size_t per_warp_bytes, size_t limit, int *wpb, size_t *smem)
{
    int w = 8;
    while (w > 1 && (size_t) w * per_warp_bytes > limit) w >>= 1;
    if ((size_t) w * per_warp_bytes > limit) return 1;
    *wpb = w;
    *smem = (size_t) w * per_warp_bytes;
    return 0;
}

/* ---- fast-path dispatch: the (NX,NU,MD) combinations compiled in ---------------------------- */
#define TQB_FUSED_LIST(X) \
    X(4, 1, 1) X(4, 1, 2) X(4, 1, 3) X(4, 1, 4) \
    X(6, 4, 1) X(6, 4, 2) X(6, 4, 3) X(6, 4, 4) \
    X(8, 3, 1) X(8, 3, 2) X(8, 3, 3) X(8, 3, 4) \
    X(12, 4, 1) X(12, 4, 2)

static bool fused_supported(int nx, int nu, int md)
{
#define X(a, b, c) if (nx == a && nu == b && md == c) return true;
    TQB_FUSED_LIST(X)
#undef X
    return false;
}

/* launch with programmatic stream serialisation (see pdl_wait in tdunes_fused.cuh); plain launch
 * while profiling (the event records between launches would serialise them anyway) or with
 * TREEQP_B200_NO_PDL=1 */
static int g_pdl = -1;
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args)
{
    if (g_pdl < 0)
    {
        const char *e = getenv("TREEQP_B200_NO_PDL");
        g_pdl = (e && e[0] == '1') ? 0 : 1;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifndef TQB_AFFINE_REFRESH
#define TQB_AFFINE_REFRESH 4 /* every so many Newton iterations the stage QPs are recomputed from lambda (k_stage_fused<0>) */
#endif
#ifndef TQB_AFFINE_MIN_NODES
#define TQB_AFFINE_MIN_NODES 4096
#endif
#ifndef TQB_SKIP_MIN_BLOCKS
#define TQB_SKIP_MIN_BLOCKS 1024 /* blocks in a level from which the re-use of unchanged blocks is tried */
#endif
template <int NX, int NU, int MD, int PNX, bool ROOT>
static int launch_bf(tqb_device *D, int begin, int cnt, int kp, int pre_ok)
{
    using Cfg = BfSel<NX, NU, MD, PNX, ROOT>;
    static bool attr_done = false;
    const int wpb = 4;
    const size_t smem = (size_t) wpb * Cfg::SMEM * sizeof(double);
    if (!attr_done)
    {
        CK(cudaFuncSetAttribute(k_build_factor<NX, NU, MD, PNX, ROOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_done = true;
    }
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_FACTOR); }
    /* wide levels try to re-use the factorisation of blocks whose subtree kept its active set; on narrow levels almost
     * every block has a changed descendant, and the build of a full block should start before the kids' level ends.
     * Sharded solves: a wide level lies below the cut, where every block and all its kids are this rank's own (ghost
     * cut nodes are kids of the replicated top levels only, which are narrow) */
    static int min_blocks = 0; /* TREEQP_B200_SKIP_MIN_BLOCKS overrides the compiled-in threshold (A/B runs) */
    if (min_blocks == 0) { const char *e = getenv("TREEQP_B200_SKIP_MIN_BLOCKS"); min_blocks = (e && atoi(e) > 0) ? atoi(e) : TQB_SKIP_MIN_BLOCKS; }
    const int try_skip = (D->skip_on && !ROOT && Cfg::BLOCKED && cnt >= min_blocks) ? 1 : 0;
    CK(launch_pdl(k_build_factor<NX, NU, MD, PNX, ROOT>, dim3((cnt + wpb - 1) / wpb), dim3(32 * wpb), smem, D->stream, pe_ < 0,
                  D->v, begin, cnt, kp, pre_ok, try_skip));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

template <int NX, int MD>
static int launch_bs(tqb_device *D, int begin, int cnt)
{
    const int wpb = 4;
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_BACK); }
    CK(launch_pdl(k_backsolve_fused<NX, MD>, dim3((cnt + wpb - 1) / wpb), dim3(32 * wpb), 0, D->stream, pe_ < 0, D->v, begin, cnt));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

/* fat levels of the root-to-leaves sweep: persistent warps + bulk copies (k_backsolve_tma); TREEQP_B200_NO_TMA=1
 * keeps the one-block-per-warp kernel everywhere */
#ifndef TQB_BS_TMA_MIN_PER_SM
#define TQB_BS_TMA_MIN_PER_SM 64 /* blocks per SM from which a level takes the persistent bulk-copy kernel */
#endif
static int g_bs_tma = -1;
template <int NX, int MD>
static int launch_bs_tma(tqb_device *D, int begin, int cnt)
{
    using Cfg = BsTmaCfg<NX, MD>;
    static int per_sm = 0;
    const size_t smem = (size_t) 4 * 2 * Cfg::BUF;
    if (per_sm == 0)
    {
        CK(cudaFuncSetAttribute(k_backsolve_tma<NX, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_backsolve_tma<NX, MD>, 128, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int full = (cnt + 3) / 4, res = D->n_sm_all * per_sm, grid = full < res ? full : res;
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_BACK); }
    CK(launch_pdl(k_backsolve_tma<NX, MD>, dim3(grid), dim3(128), smem, D->stream, pe_ < 0, D->v, begin, cnt));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

template <int NX, int NU>
static int launch_bf_chain(tqb_device *D, int top_begin, int cnt, int nlev, int kp_bottom, int pre_ok)
{
    using Cfg = FusedCfg<NX, NU, 1, NX, false>;
    static bool attr_done = false;
    const int wpb = 4;
    const size_t smem = (size_t) wpb * Cfg::SMEM * sizeof(double);
    if (!attr_done)
    {
        CK(cudaFuncSetAttribute(k_build_factor_chain<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_done = true;
    }
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_FACTOR); }
    CK(launch_pdl(k_build_factor_chain<NX, NU>, dim3((cnt + wpb - 1) / wpb), dim3(32 * wpb), smem, D->stream, pe_ < 0,
                  D->v, top_begin, cnt, nlev, kp_bottom, pre_ok));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

template <int NX>
static int launch_bs_chain(tqb_device *D, int top_begin, int cnt, int nlev)
{
    const int wpb = 4;
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_BACK); }
    CK(launch_pdl(k_backsolve_chain<NX>, dim3((cnt + wpb - 1) / wpb), dim3(32 * wpb), 0, D->stream, pe_ < 0, D->v, top_begin, cnt, nlev));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

template <int NX, int NU>
static int launch_bf_chain2(tqb_device *D, int top_begin, int cnt, int nlev, int kp_bottom, int pre_ok)
{
    using Cfg = HalfCfg<NX, NU>;
    const int wpb = 4, warps = (cnt + 1) / 2;
    const size_t smem = (size_t) wpb * 2 * Cfg::SMEM * sizeof(double);
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_FACTOR); }
    CK(launch_pdl(k_build_factor_chain2<NX, NU>, dim3((warps + wpb - 1) / wpb), dim3(32 * wpb), smem, D->stream, pe_ < 0,
                  D->v, top_begin, cnt, nlev, kp_bottom, pre_ok));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

template <int NX>
static int launch_bs_chain2(tqb_device *D, int top_begin, int cnt, int nlev)
{
    const int wpb = 4, warps = (cnt + 1) / 2;
    const int pe_ = (D->pev_on && !D->in_sweep && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (D->in_sweep) D->sweep_launches++;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_BACK); }
    CK(launch_pdl(k_backsolve_chain2<NX>, dim3((warps + wpb - 1) / wpb), dim3(32 * wpb), 0, D->stream, pe_ < 0, D->v, top_begin, cnt, nlev));
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

/* two chains per warp for blocks of dimension <= 8 (TREEQP_B200_NO_CHAIN2=1: one chain per warp) */
static int g_chain2 = -1;
static bool chain2_on()
{
    if (g_chain2 < 0)
    {
        const char *e = getenv("TREEQP_B200_NO_CHAIN2");
        g_chain2 = (e && e[0] == '1') ? 0 : 1;
    }
    return g_chain2 == 1;
}

/* chain segment that has depth l as its deepest level: returns its top depth (== l: no segment) */
static int chain_top(const tqb_device *D, int l)
{
    if (!D->chain_link || D->comm) return l;
    int lt = l;
    while (lt >= 2 && D->chain_link[lt]) lt--;
    return lt;
}

/* chain segment that has depth l as its top level: returns its deepest depth */
static int chain_bottom(const tqb_device *D, int l)
{
    if (!D->chain_link || D->comm) return l;
    int lb = l;
    while (lb + 1 < D->dims.Nh && D->chain_link[lb + 1]) lb++;
    return lb;
}

static int fused_factor_chain(tqb_device *D, int lt, int lb, int pre_ok)
{
    const int begin = D->lvl_ptr[lt], cnt = D->lvl_ptr[lt + 1] - begin;
    const int nx = D->f_nx, nu = D->f_nu;
    if (chain2_on())
    {
#define X(a, b) if (nx == a && nu == b) return launch_bf_chain2<a, b>(D, begin, cnt, lb - lt + 1, D->lvl_kp[lb], pre_ok);
        X(4, 1) X(6, 4) X(8, 3)
#undef X
    }
#define X(a, b) if (nx == a && nu == b) return launch_bf_chain<a, b>(D, begin, cnt, lb - lt + 1, D->lvl_kp[lb], pre_ok);
    X(4, 1) X(6, 4) X(8, 3) X(12, 4)
#undef X
    return 1;
}

static int fused_backsolve_chain(tqb_device *D, int lt, int lb)
{
    const int begin = D->lvl_ptr[lt], cnt = D->lvl_ptr[lt + 1] - begin;
    const int nx = D->f_nx;
    if (chain2_on())
    {
#define X(a) if (nx == a) return launch_bs_chain2<a>(D, begin, cnt, lb - lt + 1);
        X(4) X(6) X(8)
#undef X
    }
#define X(a) if (nx == a) return launch_bs_chain<a>(D, begin, cnt, lb - lt + 1);
    X(4) X(6) X(8) X(12)
#undef X
    return 1;
}

static int fused_factor_level(tqb_device *D, int l, int pre_ok)
{
    const int begin = D->lvl_ptr[l], cnt = D->lvl_ptr[l + 1] - begin;
    if (cnt <= 0) return 0;
    const int nx = D->f_nx, nu = D->f_nu, md = D->lvl_md[l], kp = D->lvl_kp[l];
    const bool root = (l == 0), pnx0 = root && D->f_pnx == 0;
#define X(a, b, c)                                                                   \
    if (nx == a && nu == b && md == c)                                               \
    {                                                                                \
        if (!root) return launch_bf<a, b, c, a, false>(D, begin, cnt, kp, pre_ok);   \
        if (pnx0) return launch_bf<a, b, c, 0, true>(D, begin, cnt, kp, pre_ok);     \
        return launch_bf<a, b, c, a, true>(D, begin, cnt, kp, pre_ok);               \
    }
    TQB_FUSED_LIST(X)
#undef X
    return 1;
}

static int fused_backsolve_level(tqb_device *D, int l)
{
    const int begin = D->lvl_ptr[l], cnt = D->lvl_ptr[l + 1] - begin;
    if (cnt <= 0) return 0;
    const int nx = D->f_nx, nu = D->f_nu, md = D->lvl_md[l];
    if (g_bs_tma < 0)
    {
        const char *e = getenv("TREEQP_B200_NO_TMA");
        g_bs_tma = (e && e[0] == '1') ? 0 : 1;
    }
    /* at least four blocks per resident warp: below that the one-block-per-warp kernel overlaps better */
    if (g_bs_tma && cnt >= TQB_BS_TMA_MIN_PER_SM * D->n_sm_all)
    {
#define X(a, b, c) if (nx == a && nu == b && md == c) { if constexpr (BsTmaCfg<a, c>::OK) return launch_bs_tma<a, c>(D, begin, cnt); }
        TQB_FUSED_LIST(X)
#undef X
    }
#define X(a, b, c) if (nx == a && nu == b && md == c) return launch_bs<a, c>(D, begin, cnt);
    TQB_FUSED_LIST(X)
#undef X
    return 1;
}

/* chain links: depth l continues depth l-1 when every parent of both depths has one kid and the w-th parent of
 * depth l is the kid of the w-th parent of depth l-1 (the scenario chains below the robust horizon of a
 * multistage tree): such runs of depths are walked by one warp per chain in one launch */
static int compute_chain_links(tqb_device *D, const tqb_tables *T, const int *ghost)
{
    const tqb_dims *d = &D->dims;
    if (D->chain_link) return 0;
    const char *e0 = getenv("TREEQP_B200_NO_CHAIN");
    D->chain_link = (int *) calloc((size_t) d->Nh + 2, sizeof(int));
    if ((e0 && e0[0] == '1') || ghost) return 0;
    for (int l = 2; l < d->Nh; l++)
    {
        const int c0 = D->lvl_ptr[l] - D->lvl_ptr[l - 1], c1 = D->lvl_ptr[l + 1] - D->lvl_ptr[l];
        if (c0 != c1 || c1 <= 0) continue;
        int ok = 1;
        for (int w = 0; w < c1 && ok; w++)
        {
            const int gp = T->lvl_nodes[D->lvl_ptr[l - 1] + w], gk = T->lvl_nodes[D->lvl_ptr[l] + w];
            ok = T->nkids[gp] == 1 && T->nkids[gk] == 1 && T->kid0[gp] == gk;
        }
        D->chain_link[l] = ok;
    }
    /* depths whose blocks all have d == nx == one value: their chain segments run the compile-time-dimension
     * instances of the level kernels (k_factor_level<DC>, k_backsolve_level<DC>) */
    D->lvl_udim = (int *) calloc((size_t) d->Nh + 2, sizeof(int));
    for (int l = 0; l < d->Nh; l++)
    {
        const int c = D->lvl_ptr[l + 1] - D->lvl_ptr[l];
        if (c <= 0) continue;
        const int u = T->dim[T->lvl_nodes[D->lvl_ptr[l]]];
        int ok = 1;
        for (int w = 0; w < c && ok; w++)
        {
            const int g = T->lvl_nodes[D->lvl_ptr[l] + w];
            ok = T->dim[g] == u && T->nx[g] == u;
        }
        D->lvl_udim[l] = ok ? u : 0;
    }
    return 0;
}

/* compile-time dimension of a chain segment (depths lt..lb) on the runtime-dimension path, 0 = none */
static int chain_udim(const tqb_device *D, int lt, int lb)
{
    if (!D->lvl_udim || lb <= lt) return 0;
    if (!D->sw_chain_dc) return 0; /* TREEQP_B200_NO_CHAIN_DC=1 */
    const int u = D->lvl_udim[lt];
    for (int l = lt; l <= lb; l++)
        if (D->lvl_udim[l] != u) return 0;
    return u;
}

/* dense stage QPs: does the first line-search pass solve all TQB_LS_NT step lengths?  Yes when that is at most
 * four warps per resident-CTA slot of the device (TREEQP_B200_DENSE_FIRST_FULL = 0 / 1 forces it, A/B runs) */
static bool dense_first_pass_full(const tqb_device *D)
{
    if (D->sw_dense_first >= 0) return D->sw_dense_first == 1;
    return (long long) D->dims.Nn * TQB_LS_NT <= 4LL * D->n_sm_all * D->wpb_dense;
}

/* chain segments of 2..32 depths take the kernels that move the next block's inputs while a block is processed
 * (k_factor_chain_pf, k_backsolve_chain_pf); TREEQP_B200_NO_CHAIN_PF=1 keeps the level kernels (A/B runs) */
static bool chain_pf_on(const tqb_device *D, int nlev)
{
    return D->sw_chain_pf && nlev >= 2 && nlev <= 32;
}

/* dimensions with a compile-time instance of the level kernels */
#define TQB_CHAIN_DC_LIST(X) X(2) X(4) X(6) X(8) X(10) X(12) X(16)

/* The fast path needs: every non-root node nx == NX, every root nx in {0, NX} (the same for all
 * trees), every parent nu == NU, all parents of one depth the same number of kids, kids of one
 * depth either all parents or all leaves, and the combination compiled in. */
static int detect_fast_path(tqb_device *D, const int *h_int, const int *ghost)
{
    const tqb_dims *d = &D->dims;
    tqb_tables T;
    tqb_layout_tables(d, const_cast<int *>(h_int), &T);
    D->fast = 0;
    const char *env = getenv("TREEQP_B200_GENERIC");
    if (env && env[0] == '1') return 0;
    if (d->Np == 0 || d->Nn < 2) return 0;
    if (d->has_dense) return 0; /* dense stage QPs run on the runtime-dimension kernels */
    int NX = -1, NU = -1, PNX = -1;
    D->lvl_md = (int *) calloc((size_t) d->Nh + 1, sizeof(int));
    D->lvl_kp = (int *) calloc((size_t) d->Nh + 1, sizeof(int));
    for (int l = 0; l <= d->Nh; l++) D->lvl_kp[l] = -1;
    for (int g = 0; g < d->Nn; g++)
    {
        const bool root = T.dad[g] < 0;
        if (root) { if (PNX < 0) PNX = T.nx[g]; else if (PNX != T.nx[g]) return 0; }
        else { if (NX < 0) NX = T.nx[g]; else if (NX != T.nx[g]) return 0; }
        if (T.nkids[g] > 0)
        {
            if (NU < 0) NU = T.nu[g]; else if (NU != T.nu[g]) return 0;
            const int l = T.depth[g];
            if (D->lvl_md[l] == 0) D->lvl_md[l] = T.nkids[g]; else if (D->lvl_md[l] != T.nkids[g]) return 0;
            for (int j = 0; j < T.nkids[g]; j++)
            {
                const int kid = T.kid0[g] + j;
                const int kp = T.nkids[kid] > 0 || (ghost && ghost[kid]);
                if (D->lvl_kp[l] < 0) D->lvl_kp[l] = kp; else if (D->lvl_kp[l] != kp) return 0;
            }
        }
    }
    if (NX <= 0 || NU < 0 || (PNX != 0 && PNX != NX)) return 0;
    for (int l = 0; l < d->Nh; l++)
        if (D->lvl_md[l] > 0 && !fused_supported(NX, NU, D->lvl_md[l])) return 0;
    D->f_nx = NX; D->f_nu = NU; D->f_pnx = PNX;
    CK(cudaMalloc(&D->d_schur, sizeof(double) * ((size_t) d->Nn * NX * NX + (size_t) d->X + 4)));
    D->v.schur = D->d_schur;
    D->v.rsub = D->d_schur + (size_t) d->Nn * NX * NX;
    {   /* re-use of unchanged blocks (tdunes_fused.cuh): per-node flags; TREEQP_B200_NO_SKIP=1 re-factors every block
         * in every iteration (A/B switch; the iterates are bit-identical either way) */
        const char *es = getenv("TREEQP_B200_NO_SKIP");
        CK(cudaMalloc(&D->d_flags, sizeof(int) * 3 * (size_t) d->Nn));
        D->v.chg = D->d_flags; D->v.bdirty = D->d_flags + d->Nn; D->v.breg = D->d_flags + 2 * (size_t) d->Nn;
        D->skip_on = !(es && es[0] == '1');
    }
    {   /* stage QPs from the affine model between refreshes (k_step_affine, tdunes_stage.cuh): TREEQP_B200_AFFINE=0 recomputes
         * them from lambda in every iteration, TREEQP_B200_AFFINE_REFRESH sets the period */
        const char *ea = getenv("TREEQP_B200_AFFINE"), *er = getenv("TREEQP_B200_AFFINE_REFRESH");
        D->affine_refresh = (ea && ea[0] == '0') ? 0 : ((er && atoi(er) > 0) ? atoi(er) : TQB_AFFINE_REFRESH);
        /* Only where the pass over A and B costs something (large trees), unless forced (=1).  Converged solutions, iteration
         * and trial counts are those of the exact path on every golden case; what moves is the path of a solve that is cut off
         * by maxIter in the middle of an ill-conditioned stretch (regularised singular blocks amplify any rounding: the
         * maxIter = 7 iterate of spring_mass_tight leaves the 1e-5 band around the oracle's), so small trees keep recomputing
         * from lambda as the reference does. */
        if (!(ea && ea[0] == '1') && d->Nn / (d->nT > 0 ? d->nT : 1) < TQB_AFFINE_MIN_NODES) D->affine_refresh = 0;
        if (NX + NU > 32) D->affine_refresh = 0;
        if (ghost)
        {   /* ghost cut nodes take their stage solution from the exchange: k_step_affine moves their multipliers only */
            CK(cudaMalloc(&D->d_ghost, sizeof(int) * (size_t) d->Nn));
            CK(cudaMemcpy(D->d_ghost, ghost, sizeof(int) * (size_t) d->Nn, cudaMemcpyHostToDevice));
            D->v.ghost = D->d_ghost;
        }
    }
    {   /* level descriptors (LvlDesc): one 48-byte record per parent in level order */
        const size_t n = (size_t) d->Np;
        int *h = (int *) calloc(12 * n + 4, sizeof(int));
        for (size_t e = 0; e < n; e++)
        {
            const int g = T.lvl_nodes[e], k0 = T.kid0[g];
            int *r = h + 12 * e;
            r[0] = g; r[1] = k0; r[2] = T.xoff[g]; r[3] = T.uoff[g];
            r[4] = T.xoff[k0]; r[5] = T.aoff[k0]; r[6] = T.boff[k0]; r[7] = T.woff[g];
            r[8] = T.utoff[g]; r[9] = T.tree[g]; r[10] = T.nkids[g];
        }
        CK(cudaMalloc(&D->d_lvl_desc, sizeof(int) * (12 * n + 4)));
        CK(cudaMemcpy(D->d_lvl_desc, h, sizeof(int) * (12 * n + 4), cudaMemcpyHostToDevice));
        free(h);
        D->v.lvl_desc = D->d_lvl_desc;
    }
    D->fast = 1;
    if (compute_chain_links(D, &T, ghost)) return 1;
    {   /* leaf records for k_stage_fused; it needs leaves without inputs and few (kid, column) pairs */
        int nleaf = 0, ok = 1;
        /* ghost nodes of a sharded solve (cut nodes owned by another rank) are leaves of the local tree
         * whose stage solution arrives with the exchange: no stage QP here */
        for (int g = 0; g < d->Nn; g++)
            if (T.nkids[g] == 0 && !(ghost && ghost[g])) { nleaf++; if (T.nu[g] != 0 || T.nx[g] != (T.dad[g] < 0 ? PNX : NX)) ok = 0; }
        for (int l = 0; l < d->Nh; l++)
            if (D->lvl_md[l] * (NX + NU) > TQB_STAGE_MAXPAIR || D->lvl_md[l] * NX > 32) ok = 0;
        const char *e2 = getenv("TREEQP_B200_GENERIC_STAGE");
        if (e2 && e2[0] == '1') ok = 0;
        if (ok)
        {
            int *h = (int *) calloc(4 * (size_t) nleaf + 4, sizeof(int));
            size_t e = 0;
            for (int g = 0; g < d->Nn; g++)
                if (T.nkids[g] == 0 && !(ghost && ghost[g]))
                {
                    int *r = h + 4 * e++;
                    r[0] = g; r[1] = T.xoff[g]; r[2] = T.tree[g]; r[3] = T.dad[g] < 0;
                }
            CK(cudaMalloc(&D->d_leaf_desc, sizeof(int) * (4 * (size_t) nleaf + 4)));
            CK(cudaMemcpy(D->d_leaf_desc, h, sizeof(int) * (4 * (size_t) nleaf + 4), cudaMemcpyHostToDevice));
            free(h);
            D->v.leaf_desc = D->d_leaf_desc;
            D->nleaf = nleaf;
            {   /* deepest run of depths (>= 1) whose parents have one kid each */
                const char *e4 = getenv("TREEQP_B200_NO_PAIR");
                int l = d->Nh;
                while (l - 1 >= 1 && D->lvl_md[l - 1] == 1) l--;
                D->pair_begin = (ghost || (e4 && e4[0] == '1')) ? d->Np : D->lvl_ptr[l];
            }
            D->fast_stage = 1;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const char *e3 = getenv("TREEQP_B200_STAGE_CTAS_PER_SM");
            D->n_sm = sms;
            D->stage_grid = e3 ? sms * atoi(e3) : 0; /* persistent warps: as many CTAs as are resident at once */
        }
    }
    return 0;
}

/* ---- small trees: one persistent CTA per tree (tdunes_small.cuh) ------------------------------ */
#ifndef TQB_SMALL_MAX_NODES
/* Measured (tools/small_probe.py, profiles/r02_small_probe.txt): the persistent kernel wins on a single tiny tree
 * (85 nodes, nx = 4: 58 iterations in 5.6 instead of 9.7 ms), is level with the level-synchronous driver around
 * 175 nodes and loses beyond (its per-level cost is one block latency, while the level kernels of a batch or of a
 * wide tree keep every SM busy): it is the default for ONE tree of at most this many nodes only. */
#define TQB_SMALL_MAX_NODES 100
#endif
template <int NX, int NU, int MD>
struct SmallSm
{
    static constexpr int A = BfSel<NX, NU, MD, NX, false>::SMEM, B = BfSel<NX, NU, MD, 0, true>::SMEM, C = BfSel<NX, NU, MD, NX, true>::SMEM;
    static constexpr int V = A > B ? (A > C ? A : C) : (B > C ? B : C);
};
template <int NX, int NU, int MDMAX>
static constexpr int small_smem_warp()
{
    int m = SmallSm<NX, NU, 1>::V;
    if constexpr (MDMAX >= 2) m = SmallSm<NX, NU, 2>::V > m ? SmallSm<NX, NU, 2>::V : m;
    if constexpr (MDMAX >= 3) m = SmallSm<NX, NU, 3>::V > m ? SmallSm<NX, NU, 3>::V : m;
    if constexpr (MDMAX >= 4) m = SmallSm<NX, NU, 4>::V > m ? SmallSm<NX, NU, 4>::V : m;
    return m;
}

/* (NX, NU, largest MD compiled in) of the persistent kernel: one instance per (NX, NU) of TQB_FUSED_LIST */
#define TQB_SMALL_LIST(X) X(4, 1, 4) X(6, 4, 4) X(8, 3, 4) X(12, 4, 2)

/* The persistent kernel needs the fused stage path, no sharding, every tree of the forest the same shape (so that
 * tree t owns the descriptors [t cnt_l, (t+1) cnt_l) of depth l and node / x- / u-space ranges of equal length)
 * and small trees.  TREEQP_B200_SMALL=0 switches it off, =1 takes it whatever the size of the trees. */
static int detect_small_path(tqb_device *D, const int *h_int, const int *ghost)
{
    const tqb_dims *d = &D->dims;
    D->small = 0;
    if (!D->fast || !D->fast_stage || ghost) return 0;
    const char *env = getenv("TREEQP_B200_SMALL");
    if (env && env[0] == '0') return 0;
    const bool force = env && env[0] == '1';
    const int nT = d->nT, Nh = d->Nh;
    if (nT <= 0 || d->Nn % nT || d->Np % nT || Nh > TQB_SMALL_MAXH) return 0;
    const int nodes = d->Nn / nT;
    if (!force && (nodes > TQB_SMALL_MAX_NODES || nT > 1)) return 0;
    tqb_tables T;
    tqb_layout_tables(d, const_cast<int *>(h_int), &T);
    SmallPlan &pl = D->small_plan;
    memset(&pl, 0, sizeof pl);
    int maxcnt = 1;
    for (int l = 0; l < Nh; l++)
    {
        const int begin = D->lvl_ptr[l], tot = D->lvl_ptr[l + 1] - begin;
        if (tot <= 0 || tot % nT || tot / nT > 32767) return 0;
        const int cnt = tot / nT;
        for (int e = 0; e < tot; e++)
            if (T.tree[T.lvl_nodes[begin + e]] != e / cnt) return 0;
        pl.lvl_begin[l] = begin; pl.lvl_cnt[l] = (short) cnt;
        pl.lvl_md[l] = (signed char) D->lvl_md[l]; pl.lvl_kp[l] = (signed char) D->lvl_kp[l];
        if (cnt > maxcnt) maxcnt = cnt;
    }
    for (int t = 0; t < nT; t++)
        if (T.tnode0[t] != t * nodes) return 0;
    if (D->nleaf % nT) return 0;
    pl.Nh = Nh; pl.nroots_total = nT; pl.nleaf = D->nleaf / nT; pl.nodes = nodes;
    pl.pnx = D->f_pnx;
    pl.xlen = D->f_pnx + (nodes - 1) * D->f_nx;
    pl.ulen = (d->Np / nT) * D->f_nu;
    if ((long long) pl.xlen * nT != d->X || (long long) pl.ulen * nT != d->U) return 0;
    {   /* leaves of tree t at [t nleaf, (t+1) nleaf) of the leaf records (node order) */
        int seen = 0;
        for (int g = 0; g < d->Nn; g++)
            if (T.nkids[g] == 0) { if (T.tree[g] != seen / pl.nleaf) return 0; seen++; }
    }
    int mdmax = 0, smw = 0;
#define X(a, b, c) if (D->f_nx == a && D->f_nu == b) { mdmax = c; smw = small_smem_warp<a, b, c>(); }
    TQB_SMALL_LIST(X)
#undef X
    if (!mdmax) return 0;
    for (int l = 0; l < Nh; l++) if (D->lvl_md[l] > mdmax) return 0;
    pl.smem_warp = smw;
    /* warps per CTA: four for a batch (four CTAs per SM), up to 16 for a few trees (a level of a single tree is
     * spread over the warps of its CTA) */
    int nw = 4;
    if (nT < 2 * D->n_sm_all)
    {
        const int lpw = 32 / tqb_pow2_ge(D->f_nx), lgrp = (pl.nleaf + lpw - 1) / lpw;
        const int want = maxcnt > lgrp ? maxcnt : lgrp;
        nw = want <= 4 ? 4 : want <= 8 ? 8 : 16; /* the reductions need a divisor of TQB_SMALL_VL / 32 */
    }
    const char *ew = getenv("TREEQP_B200_SMALL_WARPS");
    if (ew && (atoi(ew) == 4 || atoi(ew) == 8 || atoi(ew) == 16)) nw = atoi(ew);
    D->small_warps = nw;
    D->small_smem = (size_t) nw * smw * sizeof(double);
    D->small_mdmax = mdmax;
#define X(a, b, c) if (D->f_nx == a && D->f_nu == b) CK(cudaFuncSetAttribute(k_tdunes_small<a, b, c>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->small_smem));
    TQB_SMALL_LIST(X)
#undef X
    D->small = 1;
    return 0;
}

static int launch_small(tqb_device *D)
{
    const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_SMALL); }
    int done = 0;
#define X(a, b, c) if (D->f_nx == a && D->f_nu == b) { k_tdunes_small<a, b, c><<<D->dims.nT, 32 * D->small_warps, D->small_smem, D->stream>>>(D->v, D->small_plan); done = 1; }
    TQB_SMALL_LIST(X)
#undef X
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    CK(cudaGetLastError());
    return done ? 0 : 1;
}

/* stage QPs of every node: compile-time-dimension kernel on uniform trees, generic otherwise */
template <int MODE>
static int launch_stage(tqb_device *D, int cls)
{
    View &v = D->v;
    if (D->fast_stage)
    {
        constexpr int wpb = TQB_STAGE_WPB;
        const int nx = D->f_nx, nu = D->f_nu, Np = D->dims.Np, nroots = D->lvl_ptr[1] - D->lvl_ptr[0];
#define X(a, b)                                                                                        \
        if (nx == a && nu == b)                                                                        \
        {                                                                                              \
            constexpr int lpw = 32 / tqb_pow2_ge(a);                                                   \
            const int PB = (a + b <= 16 && a <= 8) ? D->pair_begin : Np;                                        \
            const int warps = PB + (Np - PB + 1) / 2 + (D->nleaf + lpw - 1) / lpw;                     \
            const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;                \
            if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, cls); }    \
            static int per_sm = 0;                                                                     \
            if (per_sm == 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stage_fused<a, b, MODE>, 32 * wpb, 0) != cudaSuccess) per_sm = 2; \
            const int res = D->stage_grid > 0 ? D->stage_grid : D->n_sm * per_sm;                      \
            const int full = (warps + wpb - 1) / wpb, grid = full < res ? full : res;                  \
            k_stage_fused<a, b, MODE><<<grid, 32 * wpb, 0, D->stream>>>(v, Np, nroots, D->f_pnx, D->nleaf, PB); \
            if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);                             \
            D->launches++;                                                                             \
            return 0;                                                                                  \
        }
        X(4, 1) X(6, 4) X(8, 3) X(12, 4)
#undef X
        return 1;
    }
    const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, cls); }
    k_stage<MODE><<<(v.Nn + D->wpb_stage - 1) / D->wpb_stage, 32 * D->wpb_stage, D->sm_stage, D->stream>>>(v);
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches++;
    return 0;
}

static int device_build(tqb_device *D, const int *h_int, const int *lvl_ptr, const int *ghost)
{
    const tqb_dims *d = &D->dims;
    D->has_ghost = ghost != NULL;
    CK(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&D->ev0));
    CK(cudaEventCreate(&D->ev1));
    CK(cudaMalloc(&D->d_int, sizeof(int) * (size_t) d->n_int));
    CK(cudaMemcpy(D->d_int, h_int, sizeof(int) * (size_t) d->n_int, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&D->d_in, sizeof(double) * (size_t) d->n_in));
    D->lvl_ptr = (int *) malloc(sizeof(int) * (size_t) (d->Nh + 2));
    memcpy(D->lvl_ptr, lvl_ptr, sizeof(int) * (size_t) (d->Nh + 2));

    View &v = D->v;
    memset(&v, 0, sizeof v);
    tqb_layout_tables(d, D->d_int, &v.T);
    tqb_layout_inputs(d, D->d_in, &v.in);

    const size_t X = d->X, U = d->U;
    D->out_doubles = 3 * X + 2 * U + (size_t) d->NC;
    size_t nwork = D->out_doubles + 8 /* x u lam mux muu mud */ + (size_t) d->NC + 2
                   + 8 * X + 4 * U + 16      /* dlam res resMod xUnc P Linv qmod0 dq | uUnc Rt rmod0 dr */
                   + (size_t) d->szW + (size_t) d->szUt + 4
                   + 4 * (size_t) d->Nn + (1 + TQB_LS_NT) * (size_t) d->nchunk + 4
                   + (d->has_dense ? (size_t) d->szH + (size_t) TQB_LS_NT * d->Nn + 4 : 0);
    CK(cudaMalloc(&D->d_work, sizeof(double) * nwork));
    double *p = D->d_work;
    v.x = p; p += X; v.u = p; p += U; v.lam = p; p += X; v.mux = p; p += X; v.muu = p; p += U;
    v.mud = p; p += d->NC; v.wsd = p; p += d->NC;
    v.ncmax = d->ncmax; v.NC = d->NC;
    p += ((uintptr_t) p & 8) ? 1 : 0;
    v.dlam = p; p += X; v.res = p; p += X; v.resMod = p; p += X; v.xUnc = p; p += X; v.P = p; p += X;
    v.Linv = p; p += X; v.uUnc = p; p += U; v.Rt = p; p += U;
    v.qmod0 = p; p += X; v.dq = p; p += X; v.rmod0 = p; p += U; v.dr = p; p += U;
    p += ((uintptr_t) p & 8) ? 1 : 0;
    v.W = p; p += d->szW; v.Ut = p; p += d->szUt;
    v.nva = p; p += d->Nn; v.nvb = p; p += d->Nn; v.cmod0 = p; p += d->Nn; v.dc = p; p += d->Nn; v.cva = p; p += d->nchunk; v.cvb = p; p += (size_t) TQB_LS_NT * d->nchunk;
    if (d->has_dense) { v.Pd = p; p += d->szH; v.fdense = p; p += (size_t) TQB_LS_NT * d->Nn; }
    v.has_dense = d->has_dense; v.nvmax = d->nvmax;
    if ((size_t) (p - D->d_work) > nwork) { snprintf(g_err, sizeof g_err, "work arena overflow"); return 1; }

    const size_t nT = d->nT;
    CK(cudaMalloc(&D->d_tint, sizeof(int) * (5 * nT + 4 + 256)));
    CK(cudaMalloc(&D->d_tdbl, sizeof(double) * (6 * nT)));
    CK(cudaMemset(D->d_tint, 0, sizeof(int) * (5 * nT + 4 + 256)));
    v.status = D->d_tint; v.iter = v.status + nT; v.ls_cnt = v.iter + nT; v.ls_active = v.ls_cnt + nT;
    v.restart_cnt = v.ls_active + nT; v.counters = v.restart_cnt + nT; v.ls_hist = v.counters + 4;
    v.t_f0 = D->d_tdbl; v.t_dot = v.t_f0 + nT; v.t_tau = v.t_dot + nT; v.t_tauprev = v.t_tau + nT;
    v.t_err = v.t_tauprev + nT; v.t_taufin = v.t_err + nT;
    CK(cudaMallocHost(&D->h_counters, sizeof(int) * 8));

    v.Nn = d->Nn; v.nT = d->nT; v.X = d->X; v.U = d->U; v.nchunk = d->nchunk;
    v.abmax = d->abmax; v.dmax = d->dmax; v.nxmax = d->nxmax; v.numax = d->numax;

    {
        int dev0 = 0, sms = 148;
        cudaGetDevice(&dev0);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev0);
        D->n_sm_all = sms;
    }
    if (detect_fast_path(D, h_int, ghost)) return 1;
    if (detect_small_path(D, h_int, ghost)) return 1;
    if (!D->chain_link)
    {   /* trees on the runtime-dimension kernels walk their chain segments in one launch as well */
        tqb_tables Th;
        tqb_layout_tables(d, const_cast<int *>(h_int), &Th);
        if (compute_chain_links(D, &Th, ghost)) return 1;
    }

    /* shared memory per warp (bytes) and warps per CTA for the block kernels */
    int dev = 0, smem_max = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t lim = (size_t) smem_max > 8192 ? (size_t) smem_max - 2048 : 40000;
    const int ldw = d->dmax | 1, ldu = d->nxmax | 1;
    v.pw_stage = d->abmax + d->dmax + (d->has_dense ? tqb_dense_smem(d->nvmax, d->ncmax) : 0);
    v.pw_hess = d->abmax + d->nxmax + d->numax + d->dmax +
                (d->has_dense ? d->dmax / 2 + 1 + d->nvmax * d->nvmax + d->dmax * d->nvmax : 0);
    const size_t pw_stage = 8ull * v.pw_stage;
    const size_t pw_hess = 8ull * v.pw_hess;
    if (d->has_dense)
    {
        if (pick_wpb(8ull * tqb_dense_smem(d->nvmax, d->ncmax), lim, &D->wpb_dense, &D->sm_dense)) return 1;
        CK(cudaFuncSetAttribute(k_trial_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_dense));
    }
    D->pw_fact = ldw * d->dmax + 2 * d->dmax + ldu * d->dmax;
    D->pw_back = ldw * d->dmax + 2 * d->dmax + d->nxmax;
    if (pick_wpb(pw_stage > 8 ? pw_stage : 8, lim, &D->wpb_stage, &D->sm_stage) ||
        pick_wpb(pw_hess > 8 ? pw_hess : 8, lim, &D->wpb_hess, &D->sm_hess) ||
        pick_wpb(8ull * (D->pw_fact > 1 ? D->pw_fact : 1), lim, &D->wpb_fact, &D->sm_fact) ||
        pick_wpb(8ull * (D->pw_back > 1 ? D->pw_back : 1), lim, &D->wpb_back, &D->sm_back))
    {
        snprintf(g_err, sizeof g_err, "block dimension %d needs more shared memory than one SM has", d->dmax);
        fprintf(stderr, "[treeqp_b200] %s\n", g_err);
        return 1;
    }
    CK(cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_stage));
    CK(cudaFuncSetAttribute(k_stage<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_stage));
    CK(cudaFuncSetAttribute(k_stage<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_stage));
    CK(cudaFuncSetAttribute(k_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_hess));
    CK(cudaFuncSetAttribute(k_factor_level<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_fact));
    CK(cudaFuncSetAttribute(k_backsolve_level<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_back));
#define X(a)                                                                                                       \
    CK(cudaFuncSetAttribute(k_factor_level<a>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_fact));    \
    CK(cudaFuncSetAttribute(k_backsolve_level<a>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_back)); \
    CK(cudaFuncSetAttribute(k_factor_chain_pf<a>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_fact)); \
    CK(cudaFuncSetAttribute(k_backsolve_chain_pf<a>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->sm_back));
    TQB_CHAIN_DC_LIST(X)
#undef X
    return 0;
}

extern "C" tqb_device *tqb_device_create(const tqb_dims *dims, const int *h_int, const int *lvl_ptr, const int *ghost)
{
    if (tqb_cuda_device_count() <= 0)
    {
        snprintf(g_err, sizeof g_err, "no CUDA device visible: libtreeqp_b200 has no CPU fallback");
        fprintf(stderr, "[treeqp_b200] %s\n", g_err);
        return NULL;
    }
    tqb_device *D = (tqb_device *) calloc(1, sizeof(tqb_device));
    D->dims = *dims;
    {   /* diagnostics switches of the runtime-dimension path */
        const char *e1 = getenv("TREEQP_B200_NO_CHAIN_DC"), *e2 = getenv("TREEQP_B200_NO_CHAIN_PF"), *e3 = getenv("TREEQP_B200_DENSE_FIRST_FULL");
        D->sw_chain_dc = (e1 && e1[0] == '1') ? 0 : 1;
        D->sw_chain_pf = (e2 && e2[0] == '1') ? 0 : 1;
        D->sw_dense_first = e3 ? (e3[0] == '1' ? 1 : 0) : -1;
    }
    if (device_build(D, h_int, lvl_ptr, ghost)) { tqb_device_destroy(D); return NULL; }
    return D;
}

extern "C" void tqb_device_destroy(tqb_device *D)
{
    if (!D) return;
    if (D->stream) cudaStreamSynchronize(D->stream);
    cudaFree(D->d_int); cudaFree(D->d_in); cudaFree(D->d_work); cudaFree(D->d_tint); cudaFree(D->d_tdbl);
    cudaFree(D->d_schur); cudaFree(D->d_lvl_desc); cudaFree(D->d_leaf_desc); free(D->lvl_md); free(D->lvl_kp); free(D->chain_link); free(D->lvl_udim);
    cudaFree(D->d_abstage); cudaFree(D->d_flags); cudaFree(D->d_ghost);
    cudaFree(D->d_cut_nodes); cudaFree(D->d_send); cudaFree(D->d_recv); cudaFree(D->d_tred);
    if (D->h_counters) cudaFreeHost(D->h_counters);
    if (D->ev0) cudaEventDestroy(D->ev0);
    if (D->ev1) cudaEventDestroy(D->ev1);
    if (D->pev) { for (int i = 0; i < 2 * TQB_MAX_EVENTS; i++) cudaEventDestroy(D->pev[i]); free(D->pev); free(D->pev_class); free(D->pev_cnt); free(D->pev_iter); }
    if (D->d_kkt) cudaFree(D->d_kkt);
    if (D->stream) cudaStreamDestroy(D->stream);
    free(D->lvl_ptr);
    free(D);
}

/* pipelined upload: the host packs node ranges while earlier ranges are already on the wire;
 * begin/end bracket the copies with events (ms_total = first byte to last byte on the stream,
 * ms_tail = what the host still had to wait for after its last range was packed) */
extern "C" int tqb_device_upload_begin(tqb_device *D)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    return 0;
}

extern "C" int tqb_device_upload_range(tqb_device *D, const double *h_in, long long off, long long n)
{
    if (n <= 0) return 0;
    CK(cudaMemcpyAsync(D->d_in + off, h_in + off, sizeof(double) * (size_t) n, cudaMemcpyHostToDevice, D->stream));
    return 0;
}

extern "C" int tqb_device_upload_2d(tqb_device *D, const double *h_in, long long off, long long pitch, long long width, int rows)
{
    if (width <= 0 || rows <= 0) return 0;
    CK(cudaMemcpy2DAsync(D->d_in + off, sizeof(double) * (size_t) pitch, h_in + off, sizeof(double) * (size_t) pitch,
                         sizeof(double) * (size_t) width, (size_t) rows, cudaMemcpyHostToDevice, D->stream));
    return 0;
}

extern "C" int tqb_device_upload_end(tqb_device *D, double *ms_total)
{
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms_total) *ms_total = t;
    return 0;
}

/* ---- caller-owned host memory as a DMA source / target -------------------------------------------------------
 * A buffer is page-locked by its OWNER: treeqp_b200_host_register(ptr, bytes) (include/treeqp_b200.h) before the
 * solve loop, treeqp_b200_host_unregister(ptr) before the buffer is freed.  The library never page-locks caller
 * memory on its own: a registration that outlives the buffer makes later copies from whatever the allocator puts at
 * that address fail.  Only the pages that lie ENTIRELY inside [ptr, ptr + bytes) are registered -- a page shared with a
 * neighbouring allocation must stay pageable, or a plain cudaMemcpy from that neighbour straddles registered and
 * unregistered memory and is refused -- so a buffer qualifies as a whole when it is page-aligned and a multiple of
 * the page size (posix_memalign(4096); the Python wrapper allocates its buffers that way). */
#define TQB_PIN_SLOTS 16
static struct { const char *lo, *hi; } g_pin[TQB_PIN_SLOTS];
static std::mutex g_pin_mu; /* solvers of several host threads share the table */

extern "C" int tqb_host_register(const void *base, size_t bytes)
{
    if (!base || bytes == 0) return 1;
    std::lock_guard<std::mutex> lock(g_pin_mu);
    const uintptr_t pg = 4096;
    const char *lo = (const char *) (((uintptr_t) base + pg - 1) & ~(pg - 1)), *hi = (const char *) (((uintptr_t) base + bytes) & ~(pg - 1));
    if (hi <= lo) return 1;
    int slot = -1;
    for (int i = 0; i < TQB_PIN_SLOTS; i++)
    {
        if (g_pin[i].lo && lo >= g_pin[i].lo && hi <= g_pin[i].hi) return 0; /* already inside a registered range */
        if (g_pin[i].lo && lo < g_pin[i].hi && hi > g_pin[i].lo)
        {   /* overlaps an older entry: that buffer is gone */
            if (cudaHostUnregister((void *) g_pin[i].lo) != cudaSuccess) cudaGetLastError();
            g_pin[i].lo = g_pin[i].hi = NULL;
        }
        if (!g_pin[i].lo && slot < 0) slot = i;
    }
    if (slot < 0) return 1;
    cudaError_t e = cudaHostRegister((void *) lo, (size_t) (hi - lo), cudaHostRegisterDefault);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        if (e != cudaErrorHostMemoryAlreadyRegistered) return 1;
    }
    g_pin[slot].lo = lo; g_pin[slot].hi = hi;
    return 0;
}

extern "C" int tqb_host_unregister(const void *base)
{   /* the range that holds `base` or starts within one page behind it (an unaligned buffer's first whole page) */
    std::lock_guard<std::mutex> lock(g_pin_mu);
    int rc = 1;
    for (int i = 0; i < TQB_PIN_SLOTS; i++)
        if (g_pin[i].lo && (const char *) base + 4096 > g_pin[i].lo && (const char *) base < g_pin[i].hi)
        {
            if (cudaHostUnregister((void *) g_pin[i].lo) != cudaSuccess) cudaGetLastError();
            g_pin[i].lo = g_pin[i].hi = NULL;
            rc = 0;
        }
    return rc;
}

/* is [base, base + bytes) page-locked? */
extern "C" int tqb_host_pin(const void *base, size_t bytes)
{
    if (bytes == 0) return 0;
    std::lock_guard<std::mutex> lock(g_pin_mu);
    const char *lo = (const char *) base, *hi = lo + bytes;
    for (int i = 0; i < TQB_PIN_SLOTS; i++)
        if (g_pin[i].lo && lo >= g_pin[i].lo && hi <= g_pin[i].hi) return 1;
    return 0;
}

extern "C" void tqb_host_unpin_all(void)
{
    std::lock_guard<std::mutex> lock(g_pin_mu);
    for (int i = 0; i < TQB_PIN_SLOTS; i++)
    {
        if (g_pin[i].lo && cudaHostUnregister((void *) g_pin[i].lo) != cudaSuccess) cudaGetLastError();
        g_pin[i].lo = g_pin[i].hi = NULL;
    }
}

/* One run of nodes in the caller's layout -> the device arrays, one warp per node: [A | B] BLASFEO panel-major (4-row
 * panels, cn columns per panel row) -> column-major, and the node's vectors b q r xmin xmax umin umax -> x- / u-space */
struct RunDev
{
    int rows, g_first, ma, na, cna, mb, nb, cnb, nx, nu;
    long long mwidth, b_off, vwidth;
    int voff[7];
};

__global__ void __launch_bounds__(256) k_unpack_run(View v, const char *mst, const char *vst, RunDev r)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= r.rows) return;
    const int g = r.g_first + w;
    const double *sa = reinterpret_cast<const double *>(mst + (size_t) w * r.mwidth);
    const double *sb = reinterpret_cast<const double *>(mst + (size_t) w * r.mwidth + r.b_off);
    const double *sv = reinterpret_cast<const double *>(vst + (size_t) w * r.vwidth);
    double *A = v.in.A + v.T.aoff[g], *B = v.in.B + v.T.boff[g];
    for (int e = lane; e < r.ma * r.na; e += 32)
    {
        const int j = e / r.ma, i = e - j * r.ma;
        A[e] = sa[(size_t) (i >> 2) * 4 * r.cna + 4 * j + (i & 3)];
    }
    for (int e = lane; e < r.mb * r.nb; e += 32)
    {
        const int j = e / r.mb, i = e - j * r.mb;
        B[e] = sb[(size_t) (i >> 2) * 4 * r.cnb + 4 * j + (i & 3)];
    }
    const int xo = v.T.xoff[g], uo = v.T.uoff[g];
    for (int i = lane; i < r.nx; i += 32)
    {
        v.in.b[xo + i] = sv[r.voff[0] + i];
        v.in.q[xo + i] = sv[r.voff[1] + i];
        v.in.xmin[xo + i] = sv[r.voff[3] + i];
        v.in.xmax[xo + i] = sv[r.voff[4] + i];
    }
    for (int i = lane; i < r.nu; i += 32)
    {
        v.in.r[uo + i] = sv[r.voff[2] + i];
        v.in.umin[uo + i] = sv[r.voff[5] + i];
        v.in.umax[uo + i] = sv[r.voff[6] + i];
    }
}

/* results of one run of nodes in the caller's tree_qp_out layout (rows of `width` bytes, padding zeroed) */
struct OutRunDev
{
    int rows, g_first, nx, nu;
    long long width;
    int off[5];
};

__global__ void __launch_bounds__(256) k_gather_run(View v, char *st, OutRunDev r)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= r.rows) return;
    const int g = r.g_first + w, xo = v.T.xoff[g], uo = v.T.uoff[g];
    double *d = reinterpret_cast<double *>(st + (size_t) w * r.width);
    for (int e = lane; e < (int) (r.width >> 3); e += 32) d[e] = 0.0;
    __syncwarp();
    for (int i = lane; i < r.nx; i += 32)
    {
        d[r.off[0] + i] = v.x[xo + i];
        d[r.off[2] + i] = v.mux[xo + i];
        d[r.off[4] + i] = v.lam[xo + i];
    }
    for (int i = lane; i < r.nu; i += 32)
    {
        d[r.off[1] + i] = v.u[uo + i];
        d[r.off[3] + i] = v.muu[uo + i];
    }
}

extern "C" int tqb_device_stage_reserve(tqb_device *D, size_t bytes)
{
    if (bytes <= D->abstage_bytes) return 0;
    CK(cudaStreamSynchronize(D->stream));
    cudaFree(D->d_abstage);
    D->d_abstage = NULL; D->abstage_bytes = 0;
    CK(cudaMalloc(&D->d_abstage, bytes));
    D->abstage_bytes = bytes;
    return 0;
}

extern "C" int tqb_device_upload_run(tqb_device *D, const tqb_in_run *r, int g_first, size_t stage_off)
{
    const int rows = r->k1 - r->k0;
    char *mst = D->d_abstage + stage_off, *vst = mst + (size_t) rows * (size_t) r->mwidth;
    if (stage_off + (size_t) rows * (size_t) (r->mwidth + r->vwidth) > D->abstage_bytes) return 1;
    CK(cudaMemcpy2DAsync(mst, (size_t) r->mwidth, r->mbase, (size_t) r->mpitch, (size_t) r->mwidth, (size_t) rows, cudaMemcpyHostToDevice, D->stream));
    /* the vector chunks of consecutive nodes usually touch (nothing else is carved between them): one linear copy --
     * as a 2D copy the 131 k rows of 480 bytes, none of them aligned to more than 32 bytes, ran at 6 GB/s */
    if (r->vpitch == r->vwidth) CK(cudaMemcpyAsync(vst, r->vbase, (size_t) r->vwidth * (size_t) rows, cudaMemcpyHostToDevice, D->stream));
    else CK(cudaMemcpy2DAsync(vst, (size_t) r->vwidth, r->vbase, (size_t) r->vpitch, (size_t) r->vwidth, (size_t) rows, cudaMemcpyHostToDevice, D->stream));
    RunDev rd;
    rd.rows = rows; rd.g_first = g_first; rd.ma = r->ma; rd.na = r->na; rd.cna = r->cna; rd.mb = r->mb; rd.nb = r->nb; rd.cnb = r->cnb;
    rd.nx = r->nx; rd.nu = r->nu; rd.mwidth = r->mwidth; rd.b_off = r->b_off; rd.vwidth = r->vwidth;
    for (int i = 0; i < 7; i++) rd.voff[i] = r->voff[i];
    k_unpack_run<<<(rows + 7) / 8, 256, 0, D->stream>>>(D->v, mst, vst, rd);
    CK(cudaGetLastError());
    return 0;
}

__global__ void k_recip(View v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.X) v.in.Qinv[i] = 1.0 / v.in.Qd[i];
    if (i < v.U) v.in.Rinv[i] = 1.0 / v.in.Rd[i];
}

extern "C" int tqb_device_finish_direct_upload(tqb_device *D, long long lam0_off, int lam0_zero)
{
    const int n = D->v.X > D->v.U ? D->v.X : D->v.U;
    k_recip<<<(n + 255) / 256, 256, 0, D->stream>>>(D->v);
    CK(cudaGetLastError());
    if (lam0_zero) CK(cudaMemsetAsync(D->d_in + lam0_off, 0, sizeof(double) * (size_t) D->v.X, D->stream));
    return 0;
}

extern "C" int tqb_device_download_begin(tqb_device *D)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    return 0;
}

extern "C" int tqb_device_download_run(tqb_device *D, const tqb_out_run *r, int g_first, size_t stage_off)
{
    const int rows = r->k1 - r->k0;
    if (stage_off + (size_t) rows * (size_t) r->width > D->abstage_bytes) return 1;
    char *st = D->d_abstage + stage_off;
    OutRunDev rd;
    rd.rows = rows; rd.g_first = g_first; rd.nx = r->nx; rd.nu = r->nu; rd.width = r->width;
    for (int i = 0; i < 5; i++) rd.off[i] = r->off[i];
    k_gather_run<<<(rows + 7) / 8, 256, 0, D->stream>>>(D->v, st, rd);
    CK(cudaGetLastError());
    if (r->pitch == r->width) CK(cudaMemcpyAsync(r->base, st, (size_t) r->width * (size_t) rows, cudaMemcpyDeviceToHost, D->stream));
    else CK(cudaMemcpy2DAsync(r->base, (size_t) r->pitch, st, (size_t) r->width, (size_t) r->width, (size_t) rows, cudaMemcpyDeviceToHost, D->stream));
    return 0;
}

extern "C" int tqb_device_download_nodes(tqb_device *D, double *h_out, long long x0, long long x1, long long u0, long long u1)
{
    const long long X = D->v.X, U = D->v.U;
    double *hx = h_out, *hu = hx + X, *hl = hu + U, *hmx = hl + X, *hmu = hmx + X;
    if (x1 > x0)
    {
        const size_t nb = sizeof(double) * (size_t) (x1 - x0);
        CK(cudaMemcpyAsync(hx + x0, D->v.x + x0, nb, cudaMemcpyDeviceToHost, D->stream));
        CK(cudaMemcpyAsync(hl + x0, D->v.lam + x0, nb, cudaMemcpyDeviceToHost, D->stream));
        CK(cudaMemcpyAsync(hmx + x0, D->v.mux + x0, nb, cudaMemcpyDeviceToHost, D->stream));
    }
    if (u1 > u0)
    {
        const size_t nb = sizeof(double) * (size_t) (u1 - u0);
        CK(cudaMemcpyAsync(hu + u0, D->v.u + u0, nb, cudaMemcpyDeviceToHost, D->stream));
        CK(cudaMemcpyAsync(hmu + u0, D->v.muu + u0, nb, cudaMemcpyDeviceToHost, D->stream));
    }
    return 0;
}

extern "C" int tqb_device_download_end(tqb_device *D, double *ms)
{
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms) *ms = t;
    return 0;
}

/* the multipliers of the last solve (still in HBM) become the dual start of the next one */
extern "C" int tqb_device_warm_start(tqb_device *D, long long lam0_off, long long n)
{
    CK(cudaMemcpyAsync(D->d_in + lam0_off, D->v.lam, sizeof(double) * (size_t) n, cudaMemcpyDeviceToDevice, D->stream));
    return 0;
}

extern "C" int tqb_device_upload(tqb_device *D, const double *h_in, long long n_in, double *ms)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    CK(cudaMemcpyAsync(D->d_in, h_in, sizeof(double) * (size_t) n_in, cudaMemcpyHostToDevice, D->stream));
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms) *ms = t;
    return 0;
}

extern "C" int tqb_device_download(tqb_device *D, double *h_out, double *ms)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    CK(cudaMemcpyAsync(h_out, D->v.x, sizeof(double) * D->out_doubles, cudaMemcpyDeviceToHost, D->stream));
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms) *ms = t;
    return 0;
}

/* KKT residual of every tree from the resident iterates (k_kkt): h_kkt[nT] */
/* Per Newton iteration, the event time of the last profiled solve (tqb_set_profiling(1)) under the reference's four
 * phase names (utils/profiling.h:58-68), in seconds: ph[0..n) solve_stage_problems (k_stage*<0>), ph[n..2n) build_dual_problem
 * (k_gradient, k_hessian, reductions, termination test; on the fused path gradient and Hessian live inside k_build_factor and
 * are counted with the Newton direction), ph[2n..3n) calculate_delta_lambda (factorisation + substitutions, exchanges),
 * ph[3n..4n) line_search (direction model, trials, Armijo, step).  Returns the number of iterations that have entries. */
extern "C" int tqb_device_iter_phase_times(tqb_device *D, double *ph, int n)
{
    static const signed char phase_of[K_NCLASS] = {/*K_INIT*/ -1, /*K_STAGE*/ 0, /*K_GRAD*/ 1, /*K_REDUCE*/ 1, /*K_DECIDE*/ 3, /*K_HESS*/ 1,
                                                   /*K_FACTOR*/ 2, /*K_BACK*/ 2, /*K_DIR*/ 3, /*K_TRIAL*/ 3, /*K_STEP*/ 3, /*K_FINISH*/ -1,
                                                   /*K_COMM*/ 2, /*K_SMALL*/ 2};
    for (int i = 0; i < 4 * n; i++) ph[i] = 0.0;
    int last = -1;
    for (int i = 0; i < D->pev_n; i++)
    {
        const int it = D->pev_iter[i], p = phase_of[D->pev_class[i]];
        if (it >= n || p < 0) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, D->pev[2 * i], D->pev[2 * i + 1]) != cudaSuccess) continue;
        ph[p * n + it] += 1e-3 * t;
        if (it > last) last = it;
    }
    return last + 1;
}

extern "C" int tqb_device_kkt(tqb_device *D, double *h_kkt)
{
    const int nT = D->dims.nT;
    if (!D->d_kkt) CK(cudaMalloc(&D->d_kkt, sizeof(unsigned long long) * (size_t) nT));
    CK(cudaMemsetAsync(D->d_kkt, 0, sizeof(unsigned long long) * (size_t) nT, D->stream));
    k_kkt<<<(D->v.Nn + 7) / 8, 256, 0, D->stream>>>(D->v, D->d_kkt);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_kkt, D->d_kkt, sizeof(double) * (size_t) nT, cudaMemcpyDeviceToHost, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    return 0;
}

#define LAUNCH(cls, kern, grid, block, smem, ...)                                          \
    do {                                                                                   \
        const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;        \
        if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, (cls)); } \
        kern<<<(grid), (block), (smem), D->stream>>>(__VA_ARGS__);                         \
        if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);                     \
        D->launches++;                                                                     \
    } while (0)

static inline int cdiv(long long a, int b) { return (int) ((a + b - 1) / b); }

static void sweep_begin(tqb_device *D, int cls)
{
    if (D->pev_on != 2 || D->pev_n >= TQB_MAX_EVENTS) return;
    D->sweep_pe = D->pev_n++;
    pev_tag(D, D->sweep_pe, cls);
    D->in_sweep = 1; D->sweep_launches = 0;
    cudaEventRecord(D->pev[2 * D->sweep_pe], D->stream);
}

static void sweep_end(tqb_device *D)
{
    if (!D->in_sweep) return;
    cudaEventRecord(D->pev[2 * D->sweep_pe + 1], D->stream);
    D->pev_cnt[D->sweep_pe] = (unsigned short) D->sweep_launches;
    D->in_sweep = 0;
}

extern "C" int tqb_device_enable_sharding(tqb_device *D, tqb_comm *comm, int cut_depth, int ncut, int my_cut0,
                                          int my_ncut, const int *cut_nodes_local)
{
    if (!D->fast)
    {
        snprintf(g_err, sizeof g_err, "subtree sharding needs a uniform tree (compile-time-dimension kernels)");
        fprintf(stderr, "[treeqp_b200] %s\n", g_err);
        return 1;
    }
    D->comm = comm; D->cut_depth = cut_depth; D->ncut = ncut; D->my_cut0 = my_cut0; D->my_ncut = my_ncut;
    const size_t per = (size_t) D->f_nx * (D->f_nx + 3);
    CK(cudaMalloc(&D->d_cut_nodes, sizeof(int) * (size_t) ncut));
    CK(cudaMemcpy(D->d_cut_nodes, cut_nodes_local, sizeof(int) * (size_t) ncut, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&D->d_send, sizeof(double) * per * (size_t) my_ncut));
    CK(cudaMalloc(&D->d_recv, sizeof(double) * per * (size_t) ncut));
    CK(cudaMalloc(&D->d_tred, sizeof(double) * (1 + TQB_LS_NT) * (size_t) D->dims.nT));
    D->v.t_red = D->d_tred;
    D->v.sharded = 1;
    {   /* peer mailboxes for the per-iteration exchanges (comm.cuh); NCCL stays as the fall-back */
        size_t cap = per * (size_t) my_ncut + (size_t) D->dims.nT, cap2 = (size_t) (1 + TQB_LS_NT) * D->dims.nT;
        if (cap2 > cap) cap = cap2;
        D->p2p = (ncut * per <= 65536 && tqb_comm_p2p_setup(comm, cap, D->stream) == 0) ? tqb_comm_p2p(comm) : NULL;
        /* the termination norm rides on the cut exchange (one exchange less per Newton iteration): the ranks' partial
         * norms over their subtrees travel behind the cut nodes, the replicated top is added by every rank itself.
         * Needs the top to be the first local nodes (it keeps its numbering) and at least one replicated level. */
        const char *en = getenv("TREEQP_B200_NORM_EXCHANGE");
        D->norm_on_cut = (D->p2p && cut_depth >= 1 && cut_nodes_local[0] > 0 && !(en && en[0] == '1')) ? 1 : 0;
        D->v.top_nodes = D->norm_on_cut ? cut_nodes_local[0] : 0;
    }
    return 0;
}

/* cut nodes -> every rank: one all-gather per Newton iteration, between the sweeps of the
 * subtrees (depth >= cut) and of the replicated top (depth < cut) */
static int exchange_cut_nodes(tqb_device *D, int is_max)
{
    const int nx = D->f_nx;
    const size_t per = (size_t) nx * (nx + 3);
    const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_COMM); }
    if (D->p2p)
    {
        if (D->norm_on_cut)
        {   /* this rank's part of the termination norm over the subtrees below the cut rides along */
            k_chunk_reduce<<<D->v.nchunk, 256, 0, D->stream>>>(D->v, D->v.nva, D->v.cva, is_max, D->v.top_nodes);
            k_tree_partial<<<(D->dims.nT + 3) / 4, 128, 0, D->stream>>>(D->v, D->v.cva, D->d_tred, is_max);
            D->launches += 2;
        }
        k_cut_exchange<<<1, 512, 0, D->stream>>>(D->v, *D->p2p, tqb_comm_p2p_next_seq(D->comm), D->d_cut_nodes, D->my_cut0,
                                                  D->my_ncut, D->ncut, nx, D->norm_on_cut ? D->d_tred : NULL, D->dims.nT, is_max);
        if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
        D->launches++;
        return 0;
    }
    k_cut_pack<<<(D->my_ncut + 3) / 4, 128, 0, D->stream>>>(D->v, D->d_cut_nodes, D->my_cut0, D->my_ncut, nx, D->d_send);
    if (tqb_comm_allgather(D->comm, D->d_send, D->d_recv, per * (size_t) D->my_ncut, D->stream)) return 1;
    k_cut_unpack<<<(D->ncut + 3) / 4, 128, 0, D->stream>>>(D->v, D->d_cut_nodes, D->ncut, nx, D->d_recv);
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    D->launches += 2;
    return 0;
}

/* this rank's per-tree partial sums (cva -> t_red[0..nT), cvb -> t_red[nT..2nT)), all-reduced */
static int allreduce_tree_sums(tqb_device *D, int use_a, int use_b, int a_is_max)
{
    View &v = D->v;
    const int nT = D->dims.nT, grid = (nT + 3) / 4;
    const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_COMM); }
    if (use_a) { k_tree_partial<<<grid, 128, 0, D->stream>>>(v, v.cva, D->d_tred, a_is_max); D->launches++; }
    if (use_b) { k_tree_partial<<<grid, 128, 0, D->stream>>>(v, v.cvb, D->d_tred + nT, 0); D->launches++; }
    int rc = 0;
    if (D->p2p)
    {   /* one exchange: [0, nT) from cva (max or sum), [nT, 2 nT) from cvb (sum) */
        double *buf = use_a ? D->d_tred : D->d_tred + nT;
        const int cnt = (use_a ? nT : 0) + (use_b ? nT : 0);
        k_red_exchange<<<1, 256, 0, D->stream>>>(*D->p2p, tqb_comm_p2p_next_seq(D->comm), buf, cnt, (use_a && a_is_max) ? nT : 0);
        D->launches++;
    }
    else if (use_a && use_b && !a_is_max) rc = tqb_comm_allreduce(D->comm, D->d_tred, 2 * (size_t) nT, 0, D->stream);
    else
    {
        if (use_a) rc |= tqb_comm_allreduce(D->comm, D->d_tred, (size_t) nT, a_is_max, D->stream);
        if (use_b) rc |= tqb_comm_allreduce(D->comm, D->d_tred + nT, (size_t) nT, 0, D->stream);
    }
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    return rc;
}

/* batched line search: res'*dlam (first batch only) and the NT dual function values per tree */
static int allreduce_ls_batch(tqb_device *D, int first)
{
    View &v = D->v;
    const int nT = D->dims.nT;
    const int pe_ = (D->pev_on && D->pev_n < TQB_MAX_EVENTS) ? D->pev_n++ : -1;
    if (pe_ >= 0) { cudaEventRecord(D->pev[2 * pe_], D->stream); pev_tag(D, pe_, K_COMM); }
    k_tree_partial_ls<<<dim3((nT + 3) / 4, 1 + TQB_LS_NT), 128, 0, D->stream>>>(v, D->d_tred);
    D->launches++;
    const size_t off = first ? 0 : (size_t) nT;
    int rc = 0;
    if (D->p2p)
    {
        k_red_exchange<<<1, 256, 0, D->stream>>>(*D->p2p, tqb_comm_p2p_next_seq(D->comm), D->d_tred + off,
                                                  (int) ((size_t) (1 + TQB_LS_NT) * nT - off), 0);
        D->launches++;
    }
    else rc = tqb_comm_allreduce(D->comm, D->d_tred + off, (size_t) (1 + TQB_LS_NT) * nT - off, 0, D->stream);
    if (pe_ >= 0) cudaEventRecord(D->pev[2 * pe_ + 1], D->stream);
    return rc;
}

/* counters[idx] -> pinned slot, stream-ordered; read after the next synchronisation */
static int peek_counter(tqb_device *D, int idx, int slot)
{
    CK(cudaMemcpyAsync(D->h_counters + slot, D->v.counters + idx, sizeof(int), cudaMemcpyDeviceToHost, D->stream));
    return 0;
}

extern "C" int tqb_device_solve(tqb_device *D, const tqb_opts *o, int *h_status, int *h_iter,
                                int *ls_iters256, int *ls_total, int *n_reg, double *gpu_ms,
                                long long *launches, double *kernel_ms16, int *kernel_count16)
{
    D->pev_on = profiling_on();
    D->pev_n = 0;
    D->in_sweep = 0;
    if (D->pev_cnt) memset(D->pev_cnt, 0, sizeof(unsigned short) * TQB_MAX_EVENTS);
    k_names = D->fast ? k_names_fused : k_names_generic;
    if (D->pev_on && !D->pev)
    {
        D->pev = (cudaEvent_t *) calloc(2 * TQB_MAX_EVENTS, sizeof(cudaEvent_t));
        D->pev_class = (unsigned char *) calloc(TQB_MAX_EVENTS, 1);
        D->pev_cnt = (unsigned short *) calloc(TQB_MAX_EVENTS, sizeof(unsigned short));
        D->pev_iter = (unsigned short *) calloc(TQB_MAX_EVENTS, sizeof(unsigned short));
        for (int i = 0; i < 2 * TQB_MAX_EVENTS; i++) CK(cudaEventCreate(&D->pev[i]));
    }
    View &v = D->v;
    v.o = *o;
    const tqb_dims *d = &D->dims;
    const int nT = d->nT, Nh = d->Nh;
    const int tgrid = cdiv(nT, 4);
    D->launches = 0;

    D->cur_it = 65535;
    CK(cudaEventRecord(D->ev0, D->stream));
    {
        int n = v.X > nT ? v.X : nT;
        if (n < 256) n = 256;
        if (n < v.Nn) n = v.Nn;
        if (n < v.U) n = v.U;
        if (n < d->NC) n = d->NC;
        LAUNCH(K_INIT, k_init, cdiv(n, 256), 256, 0, v);
    }
    /* Host round trips: ONE per Newton iteration (after the first line-search batch).  The
     * termination counter is fetched asynchronously and everything behind the termination test is
     * enqueued at once: kernels of a finished tree return at their status check. */
    if (D->small && !D->comm && D->pev_on != 1)
    {   /* small trees: the Newton loop of every tree runs inside one persistent CTA (not under per-launch profiling, whose
         * per-iteration phase times -- the PROFILE > 0 build's timers -- need the launches) */
        if (launch_small(D)) { snprintf(g_err, sizeof g_err, "small-tree dispatch failed"); return 1; }
    }
    else
    for (int it = 0; it < o->maxIter; it++)
    {
        CK(cudaMemsetAsync(v.counters, 0, sizeof(int) * 2, D->stream));
        D->cur_it = it;
        /* the step kernel of the last iteration has already formed this iteration's stage solution from the affine model,
         * except at iteration 0 and at every refresh */
        const bool affine = D->fast && D->fast_stage && D->affine_refresh > 0;
        if (!(affine && it % D->affine_refresh != 0))
        {
            if (launch_stage<0>(D, K_STAGE)) return 1;
        }
        if (D->fast)
        {   /* gradient + Hessian + factorisation fused, leaves to root; the termination test comes
             * after the sweep (the sweep of the last iteration is the price of the fusion) */
            sweep_begin(D, K_FACTOR);
            for (int l = Nh - 1; l >= 0; l--)
            {
                const bool exch = D->comm && l == D->cut_depth - 1;
                if (exch)
                {   /* the exchange is its own class: close the bracket around it */
                    sweep_end(D);
                    if (exchange_cut_nodes(D, o->termCondition == 2)) return 1;
                    sweep_begin(D, K_FACTOR);
                }
                const int pre_ok = (l < Nh - 1 && !exch) ? 1 : 0;
                const int lt = chain_top(D, l);
                if (lt < l)
                {   /* depths lt..l are chains: one launch walks them all */
                    if (fused_factor_chain(D, lt, l, pre_ok)) { snprintf(g_err, sizeof g_err, "fused dispatch failed"); return 1; }
                    l = lt;
                    continue;
                }
                if (fused_factor_level(D, l, pre_ok)) { snprintf(g_err, sizeof g_err, "fused dispatch failed"); return 1; }
            }
            sweep_end(D);
            if (!(D->comm && D->norm_on_cut))
            {
                LAUNCH(K_REDUCE, k_chunk_reduce, v.nchunk, 256, 0, v, v.nva, v.cva, o->termCondition == 2, 0);
                if (D->comm && allreduce_tree_sums(D, 1, 0, o->termCondition == 2)) return 1;
            }
            LAUNCH(K_DECIDE, k_decide_term, tgrid, 128, 0, v);
            if (peek_counter(D, 0, 4)) return 1;
            sweep_begin(D, K_BACK);
            for (int l = 1; l < Nh; l++)
            {
                const int lb = chain_bottom(D, l);
                if (lb > l && D->chain_link[l + 1])
                {
                    if (fused_backsolve_chain(D, l, lb)) { snprintf(g_err, sizeof g_err, "fused dispatch failed"); return 1; }
                    l = lb;
                    continue;
                }
                if (fused_backsolve_level(D, l)) { snprintf(g_err, sizeof g_err, "fused dispatch failed"); return 1; }
            }
            sweep_end(D);
        }
        else
        {
            LAUNCH(K_GRAD, k_gradient, cdiv(v.Nn, 8), 256, 0, v);
            LAUNCH(K_REDUCE, k_chunk_reduce, v.nchunk, 256, 0, v, v.nva, v.cva, o->termCondition == 2, 0);
            LAUNCH(K_DECIDE, k_decide_term, tgrid, 128, 0, v);
            if (peek_counter(D, 0, 4)) return 1;

            LAUNCH(K_HESS, k_hessian, cdiv(v.Nn, D->wpb_hess), 32 * D->wpb_hess, D->sm_hess, v);
            for (int l = Nh - 1; l >= 0; l--)
            {   /* depths lt..l that form chain segments are walked by one launch */
                const int lt = chain_top(D, l);
                const int cnt = D->lvl_ptr[lt + 1] - D->lvl_ptr[lt];
                if (cnt > 0)
                {
                    const int dc = chain_udim(D, lt, l);
                    bool done = false;
#define X(a) if (dc == a) { if (chain_pf_on(D, l - lt + 1)) LAUNCH(K_FACTOR, k_factor_chain_pf<a>, cdiv(cnt, D->wpb_fact), 32 * D->wpb_fact, D->sm_fact, v, D->lvl_ptr[lt], cnt, D->pw_fact, l - lt + 1); \
                             else LAUNCH(K_FACTOR, k_factor_level<a>, cdiv(cnt, D->wpb_fact), 32 * D->wpb_fact, D->sm_fact, v, D->lvl_ptr[lt], cnt, D->pw_fact, l - lt + 1); done = true; }
                    TQB_CHAIN_DC_LIST(X)
#undef X
                    if (!done) LAUNCH(K_FACTOR, k_factor_level<0>, cdiv(cnt, D->wpb_fact), 32 * D->wpb_fact, D->sm_fact, v, D->lvl_ptr[lt], cnt, D->pw_fact, l - lt + 1);
                }
                l = lt;
            }
            for (int l = 1; l < Nh; l++)
            {
                int lb = chain_bottom(D, l);
                if (!(lb > l && D->chain_link[l + 1])) lb = l;
                const int cnt = D->lvl_ptr[l + 1] - D->lvl_ptr[l];
                if (cnt > 0)
                {
                    const int dc = chain_udim(D, l, lb);
                    bool done = false;
#define X(a) if (dc == a) { if (chain_pf_on(D, lb - l + 1)) LAUNCH(K_BACK, k_backsolve_chain_pf<a>, cdiv(cnt, D->wpb_back), 32 * D->wpb_back, D->sm_back, v, D->lvl_ptr[l], cnt, D->pw_back, lb - l + 1); \
                             else LAUNCH(K_BACK, k_backsolve_level<a>, cdiv(cnt, D->wpb_back), 32 * D->wpb_back, D->sm_back, v, D->lvl_ptr[l], cnt, D->pw_back, lb - l + 1); done = true; }
                    TQB_CHAIN_DC_LIST(X)
#undef X
                    if (!done) LAUNCH(K_BACK, k_backsolve_level<0>, cdiv(cnt, D->wpb_back), 32 * D->wpb_back, D->sm_back, v, D->lvl_ptr[l], cnt, D->pw_back, lb - l + 1);
                }
                l = lb;
            }
        }
        /* affine model of the stage gradients along dlam + res'*dlam, then the line search */
        if (launch_stage<2>(D, K_DIR)) return 1;
        {   /* f(0), res'*dlam and the first TQB_LS_NT-1 step lengths in one pass, Armijo on the device;
             * dense stage QPs: the same step lengths solved by k_trial_dense first */
            for (int batch = 0, trials = 0; trials <= o->lsMaxIter + 1; batch++)
            {
                /* slots looked at in this pass: all TQB_LS_NT for clipping trees (a pass only streams vectors);
                 * with dense stage QPs the first pass solves tau = 0 and the full step only -- unless the whole
                 * pass is less than one wave of warps anyway (a single small tree: the extra step lengths cost
                 * no time and save the passes of a backtracking iteration; same Armijo sequence either way) */
                const int nt = (v.has_dense && batch == 0 && !dense_first_pass_full(D)) ? 2 : TQB_LS_NT;
                if (v.has_dense) LAUNCH(K_TRIAL, k_trial_dense, cdiv((long long) v.Nn * nt, D->wpb_dense), 32 * D->wpb_dense, D->sm_dense, v, batch == 0, nt);
                LAUNCH(K_TRIAL, k_ls_eval, v.nchunk, 256, 0, v, batch == 0);
                if (D->comm && allreduce_ls_batch(D, batch == 0)) return 1;
                LAUNCH(K_DECIDE, k_ls_decide, nT, 32 * (TQB_LS_NT + 1), 0, v, batch == 0, nt);
                if (peek_counter(D, 1, 5)) return 1;
                CK(cudaStreamSynchronize(D->stream));
                if (D->h_counters[4] == 0 || D->h_counters[5] == 0) break;
                trials += nt - (batch == 0 ? 1 : 0);
            }
        }
        if (D->h_counters[4] == 0) break; /* every tree has terminated */
        if (affine)
        {   /* lambda step, fused with the next iteration's stage solution unless that iteration refreshes it */
            const int ws = ((it + 1) % D->affine_refresh != 0) ? 1 : 0;
            const int nv = D->f_nx + D->f_nu;
            if (nv <= 8) LAUNCH(K_STEP, k_step_affine<8>, cdiv(v.Nn, 32), 256, 0, v, ws);
            else if (nv <= 16) LAUNCH(K_STEP, k_step_affine<16>, cdiv(v.Nn, 16), 256, 0, v, ws);
            else LAUNCH(K_STEP, k_step_affine<32>, cdiv(v.Nn, 8), 256, 0, v, ws);
        }
        else if (!D->comm && !D->has_ghost) LAUNCH(K_STEP, k_lambda_step_chunks, v.nchunk, 256, 0, v);
        else if (v.nxmax <= 4) LAUNCH(K_STEP, k_lambda_step_w<4>, cdiv(v.Nn, 64), 256, 0, v);
        else if (v.nxmax <= 8) LAUNCH(K_STEP, k_lambda_step_w<8>, cdiv(v.Nn, 32), 256, 0, v);
        else if (v.nxmax <= 16) LAUNCH(K_STEP, k_lambda_step_w<16>, cdiv(v.Nn, 16), 256, 0, v);
        else if (v.nxmax <= 32) LAUNCH(K_STEP, k_lambda_step_w<32>, cdiv(v.Nn, 8), 256, 0, v);
        else LAUNCH(K_STEP, k_lambda_step, cdiv(v.Nn, 256), 256, 0, v);
    }
    D->cur_it = 65535;
    /* trees that ran out of iterations: x,u of the last trial point (what the reference holds) */
    if (launch_stage<3>(D, K_STAGE)) return 1;
    LAUNCH(K_DECIDE, k_finalize_status, cdiv(nT, 256), 256, 0, v);
    {
        int n = v.X > v.U ? v.X : v.U;
        LAUNCH(K_FINISH, k_finish, cdiv(n > 0 ? n : 1, 256), 256, 0, v);
        if (v.has_dense) LAUNCH(K_FINISH, k_finish_dense, cdiv(v.Nn, 256), 256, 0, v);
    }
    CK(cudaEventRecord(D->ev1, D->stream));
    if (D->comm) tqb_comm_p2p_error(D->comm, D->stream);
    CK(cudaMemcpyAsync(h_status, v.status, sizeof(int) * (size_t) (2 * nT), cudaMemcpyDeviceToHost, D->stream));
    CK(cudaMemcpyAsync(D->h_counters, v.counters, sizeof(int) * 4, cudaMemcpyDeviceToHost, D->stream));
    if (ls_iters256)
        CK(cudaMemcpyAsync(ls_iters256, v.ls_hist, sizeof(int) * 256, cudaMemcpyDeviceToHost, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    CK(cudaGetLastError());
    if (D->comm && tqb_comm_p2p_error(D->comm, D->stream))
    {
        snprintf(g_err, sizeof g_err, "sharded solve: a peer GPU did not answer an exchange within 3 s");
        fprintf(stderr, "[treeqp_b200] %s\n", g_err);
        return 1;
    }
    (void) h_iter; /* h_iter == h_status + nT: filled by the copy above (status and iter are adjacent) */
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, D->ev0, D->ev1));
    if (gpu_ms) *gpu_ms = ms;
    if (ls_total) *ls_total = D->h_counters[3];
    if (n_reg) *n_reg = D->h_counters[2];
    if (launches) *launches = D->launches;
    if (kernel_ms16 && kernel_count16)
    {
        for (int k = 0; k < 16; k++) { kernel_ms16[k] = 0.0; kernel_count16[k] = 0; }
        for (int i = 0; i < D->pev_n; i++)
        {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, D->pev[2 * i], D->pev[2 * i + 1]));
            kernel_ms16[D->pev_class[i]] += t;
            kernel_count16[D->pev_class[i]] += D->pev_cnt[i] ? D->pev_cnt[i] : 1;
        }
    }
    return 0;
}
