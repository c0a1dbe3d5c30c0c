/*
 * tdunes.c -- the reference's TDUNES C API (treeqp/src/dual_Newton_tree.h:154-168) on top of the
 * CUDA engine, plus the additive device-resident and batch entry points.
 *
 * Ownership follows the reference: the caller owns the host buffer handed to
 * treeqp_tdunes_create (sized by treeqp_tdunes_calculate_size).  Device memory cannot live in
 * that buffer, so `create` also allocates a device arena that treeqp_tdunes_destroy releases
 * (an API extension: callers that never call it leak the arena at exit, nothing else changes).
 */
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "treeqp_b200.h"
#include "host_internal.h"

#define TQB_MAGIC 0x7D0B200u

typedef struct tqb_ctx
{
    unsigned magic;
    tqb_forest *forest;
    int lam_dirty;  /* lambda0 changed on the host since the last upload */
    int uploaded;   /* problem data present on the device */
    int solved;     /* a Newton loop has run: the device holds its final multipliers */
    return_t last_status;
} tqb_ctx;

struct treeqp_tdunes_batch_
{
    int n;
    tqb_forest *forest;
    tqb_opts o;
    int opts_ok;
    int uploaded;
};

/* TREEQP_B200_VERBOSE=1: one line per drop-in solve on stderr (which library ran, how many launches) */
int tqb_verbose(void)
{
    static int v = -1;
    if (v < 0)
    {
        const char *e = getenv("TREEQP_B200_VERBOSE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

const char *treeqp_b200_version(void) { return "treeqp_b200 0.1 (sm_100a)"; }
int treeqp_b200_live_contexts(void) { return tqb_registry_size(); }
int treeqp_b200_host_register(void *ptr, size_t bytes) { return tqb_host_register(ptr, bytes); }
int treeqp_b200_host_unregister(void *ptr)
{
    if (ptr == NULL) { tqb_host_unpin_all(); return 0; } /* everything, including ranges registered under TREEQP_B200_PIN */
    return tqb_host_unregister(ptr);
}
int treeqp_b200_device_count(void) { return tqb_cuda_device_count(); }
int treeqp_b200_set_device(int device) { return tqb_cuda_set_device(device); }
void treeqp_b200_set_profiling(int on) { tqb_set_profiling(on); }
int treeqp_b200_num_kernel_classes(void) { return tqb_num_kernel_classes(); }
const char *treeqp_b200_kernel_class_name(int k) { return tqb_kernel_class_name(k); }

/* ----------------------------------------------------------------------------- options ---- */

int treeqp_tdunes_opts_calculate_size(int Nn) { return Nn * (int) sizeof(stage_qp_t); }

void treeqp_tdunes_opts_create(int Nn, treeqp_tdunes_opts_t *opts, void *ptr)
{
    (void) Nn;
    opts->qp_solver = (stage_qp_t *) ptr;
}

/* defaults of dual_Newton_tree.c:92-120 */
void treeqp_tdunes_opts_set_default(int Nn, treeqp_tdunes_opts_t *opts)
{
    opts->maxIter = 100;
    opts->termCondition = TREEQP_INFNORM;
    opts->stationarityTolerance = 1.0e-8;
    opts->checkLastActiveSet = 1;
    for (int i = 0; i < Nn; i++) opts->qp_solver[i] = TREEQP_CLIPPING_SOLVER;
    opts->lineSearchMaxIter = 50;
    opts->lineSearchGamma = 0.1;
    opts->lineSearchBeta = 0.6;
    opts->lineSearchRestartTrigger = -1;
    opts->regType = TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT;
    opts->regTol = 1.0e-6;
    opts->regValue = 1.0e-6;
}

/* dual_Newton_tree.c:1078-1100 */
static return_t validate_opts(const treeqp_tdunes_opts_t *o)
{
    if (o->termCondition != TREEQP_SUMSQUAREDERRORS && o->termCondition != TREEQP_TWONORM &&
        o->termCondition != TREEQP_INFNORM)
        return TREEQP_INVALID_OPTION;
    if ((int) o->regType != TREEQP_NO_REGULARIZATION &&
        (int) o->regType != TREEQP_ALWAYS_LEVENBERG_MARQUARDT &&
        (int) o->regType != TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT)
        return TREEQP_INVALID_OPTION;
    if (o->regValue < 0) return TREEQP_INVALID_OPTION;
    return TREEQP_OK;
}

static void forward_opts(const treeqp_tdunes_opts_t *o, tqb_opts *d)
{
    d->maxIter = o->maxIter;
    d->termCondition = (int) o->termCondition;
    d->regType = (int) o->regType;
    d->lsMaxIter = o->lineSearchMaxIter;
    d->lsRestartTrigger = o->lineSearchRestartTrigger;
    d->tol = o->stationarityTolerance;
    d->regTol = o->regTol;
    d->regValue = o->regValue;
    d->gamma = o->lineSearchGamma;
    d->beta = o->lineSearchBeta;
}

/* --------------------------------------------------------------------------- workspace ---- */

static int host_bytes(int Nn, int Nh)
{
    long long b = 0;
    b += (long long) (Nh + 1) * sizeof(int);   /* npar */
    b += (long long) Nn * sizeof(int);         /* idxpos */
    b += 2LL * Nn * sizeof(double);            /* fval, cmod */
    b += (long long) Nn * sizeof(void *);      /* stage_qp_data */
    b = (b + 63) / 64 * 64 + 3 * 64;
    return b > 2147483647LL ? -1 : (int) b; /* the reference's int return type cannot hold more (callers check <= 0) */
}

int treeqp_tdunes_calculate_size(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts)
{
    int Nn = qp_in->N;
    int bytes = host_bytes(Nn, qp_in->tree[Nn - 1].stage);
#if defined(PROFILE) && PROFILE > 0
    if (bytes > 0) bytes += timers_calculate_size(opts->maxIter) + 64; /* dual_Newton_tree.c:1399-1401 */
#else
    (void) opts;
#endif
    return bytes;
}

static void check_stage_solvers(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts)
{
    for (int k = 0; k < qp_in->N; k++)
    {
        if (opts->qp_solver[k] == TREEQP_CLIPPING_SOLVER)
        {
            if (tqb_clipping_applicable(qp_in, k) == NO)
            {   /* same reaction as clipping.c:70-74 */
                printf("[TREEQP]: Error! Specified stage QP solver (clipping) not applicable.\n");
                exit(1);
            }
        }
        else if (opts->qp_solver[k] == TREEQP_QPOASES_SOLVER)
        {   /* dense stage QP: one warp per node, one lane per variable and per general constraint row */
            if (qp_in->nx[k] + qp_in->nu[k] + qp_in->nc[k] > 32 || qp_in->nc[k] > TQB_NC_MAX)
            {
                printf("[TREEQP_B200]: Error! Dense stage QPs need nx + nu + nc <= 32 and nc <= %d (node %d: nx=%d nu=%d nc=%d).\n",
                       TQB_NC_MAX, k, qp_in->nx[k], qp_in->nu[k], qp_in->nc[k]);
                exit(1);
            }
        }
        else
        {   /* dual_Newton_tree.c:158-160 */
            printf("[TREEQP] Error! Unknown stage QP solver specified.\n");
            exit(1);
        }
    }
}

static void ctx_free(void *p);

void treeqp_tdunes_create(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                          treeqp_tdunes_workspace *work, void *ptr)
{
    const int Nn = qp_in->N, Nh = qp_in->tree[Nn - 1].stage;
    memset(work, 0, sizeof *work);
    check_stage_solvers(qp_in, opts);

    char *cur = (char *) ptr;
    uintptr_t a = ((uintptr_t) cur + 63u) / 64u * 64u;
    cur = (char *) a;
    work->stage_qp_data = (void **) cur; cur += (size_t) Nn * sizeof(void *);
    work->fval = (double *) cur;         cur += (size_t) Nn * sizeof(double);
    work->cmod = (double *) cur;         cur += (size_t) Nn * sizeof(double);
    work->npar = (int *) cur;            cur += (size_t) (Nh + 1) * sizeof(int);
    work->idxpos = (int *) cur;          cur += (size_t) Nn * sizeof(int);
    assert(cur <= (char *) ptr + host_bytes(Nn, Nh));
#if defined(PROFILE) && PROFILE > 0
    {   /* dual_Newton_tree.c:1638-1642: the timers live behind the workspace's own tables */
        char *tp = (char *) ptr + host_bytes(Nn, Nh);
        tp = (char *) (((uintptr_t) tp + 63u) / 64u * 64u);
        timers_create(opts->maxIter, &work->timings, tp);
        timers_initialize(&work->timings);
        tqb_set_profiling(1); /* an event pair around every launch feeds them */
    }
#endif

    int Np = 0;
    for (int l = 0; l <= Nh; l++) work->npar[l] = 0;
    for (int k = 0; k < Nn; k++)
    {   /* setup_npar :166, setup_idxpos :177 */
        const struct node *nd = &qp_in->tree[k];
        work->npar[nd->stage]++;
        if (nd->nkids > 0) Np++;
        int acc = 0;
        for (int j = 0; j < nd->nkids; j++)
        {
            work->idxpos[nd->kids[j]] = acc;
            acc += qp_in->nx[nd->kids[j]];
        }
        work->stage_qp_data[k] = NULL;
        work->fval[k] = work->cmod[k] = 0.0;
    }
    work->idxpos[0] = 0;
    work->Nn = Nn;
    work->Np = Np;

    tqb_ctx *ctx = (tqb_ctx *) calloc(1, sizeof(tqb_ctx));
    ctx->magic = TQB_MAGIC;
    const tree_qp_in *one[1] = {qp_in};
    ctx->forest = tqb_forest_create(1, one, (const int *) opts->qp_solver);
    if (!ctx->forest)
    {
        fprintf(stderr, "[treeqp_b200] treeqp_tdunes_create: no device arena (%s); solve will return "
                        "TREEQP_UNKNOWN_ERROR\n", tqb_cuda_last_error());
    }
    else
    {   /* dual start defaults to zero until set_dual_initialization is called */
        memset(ctx->forest->hin.lam0, 0, sizeof(double) * (size_t) ctx->forest->dims.X);
    }
    ctx->lam_dirty = 1;
    work->stage_qp_data[0] = ctx;
    /* the reference API has no destroy: a re-create over the same caller memory releases the arena of the
     * workspace that lived there (registry.c) */
    tqb_registry_replace(ptr, ctx, ctx_free);
}

static tqb_ctx *ctx_of(const treeqp_tdunes_workspace *work)
{
    if (!work || !work->stage_qp_data) return NULL;
    tqb_ctx *ctx = (tqb_ctx *) work->stage_qp_data[0];
    if (!ctx || ctx->magic != TQB_MAGIC) return NULL;
    return ctx;
}

static void ctx_free(void *p)
{
    tqb_ctx *ctx = (tqb_ctx *) p;
    if (!ctx || ctx->magic != TQB_MAGIC) return;
    tqb_forest_destroy(ctx->forest);
    ctx->magic = 0;
    free(ctx);
}

void treeqp_tdunes_destroy(treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx) return;
    tqb_registry_remove(ctx);
    ctx_free(ctx);
    work->stage_qp_data[0] = NULL;
}

void treeqp_tdunes_set_dual_initialization(const double *lambda, treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest) return;
    tqb_forest_set_lambda0(ctx->forest, 0, lambda);
    ctx->lam_dirty = 1;
}

void treeqp_tdunes_get_stats(const treeqp_tdunes_workspace *work, treeqp_b200_stats *s)
{
    memset(s, 0, sizeof *s);
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest) return;
    const tqb_forest *f = ctx->forest;
    s->iter = f->iter_max;
    s->ls_total = f->ls_total;
    memcpy(s->ls_iters, f->ls_iters, sizeof s->ls_iters);
    s->n_reg = f->n_reg;
    s->gpu_ms = f->gpu_ms; s->h2d_ms = f->h2d_ms; s->d2h_ms = f->d2h_ms; s->pack_ms = f->pack_ms;
    s->h2d_bytes = f->h2d_bytes; s->d2h_bytes = f->d2h_bytes;
    s->kernel_launches = f->launches;
    memcpy(s->kernel_ms, f->kernel_ms, sizeof s->kernel_ms);
    memcpy(s->kernel_count, f->kernel_count, sizeof s->kernel_count);
}

/* ------------------------------------------------------------------------------- solve ---- */

static return_t forest_solve(tqb_forest *f, const tqb_opts *o)
{
    if (tqb_device_solve(f->dev, o, f->h_status, f->h_iter, f->ls_iters, &f->ls_total, &f->n_reg,
                         &f->gpu_ms, &f->launches, f->kernel_ms, f->kernel_count))
        return TREEQP_UNKNOWN_ERROR;
    int it = 0;
    for (int t = 0; t < f->dims.nT; t++)
        if (f->h_iter[t] > it) it = f->h_iter[t];
    f->iter_max = it;
    return TREEQP_OK;
}

static return_t forest_download(tqb_forest *f)
{
    if (tqb_device_download(f->dev, f->h_out, &f->d2h_ms)) return TREEQP_UNKNOWN_ERROR;
    f->d2h_bytes = (long long) sizeof(double) * (3LL * f->dims.X + 2LL * f->dims.U + f->dims.NC);
    return TREEQP_OK;
}

/* bytes of the last pack_upload: the packed input block, or -- DMA runs -- their rows in the caller's layout (padding
 * included) plus what still goes through the staging block */
static long long upload_bytes(const tqb_forest *f)
{
    if (!f->in_direct) return (long long) sizeof(double) * f->dims.n_in;
    long long by = (long long) sizeof(double) * ((f->lam0_zero ? 1LL : 2LL) * f->dims.X + f->dims.U); /* Qd (lam0) | Rd */
    int covered = 0;
    for (int i = 0; i < f->n_in_runs; i++)
    {
        const tqb_in_run *r = &f->in_runs[i];
        by += (long long) (r->k1 - r->k0) * (r->mwidth + r->vwidth);
        covered += r->k1 - r->k0;
    }
    by += (long long) sizeof(double) * (f->dims.n_in / (f->dims.Nn > 0 ? f->dims.Nn : 1)) * (f->dims.Nn - covered); /* staged nodes */
    return by;
}

return_t treeqp_tdunes_upload(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                              treeqp_tdunes_workspace *work)
{
    (void) opts;
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest) return TREEQP_UNKNOWN_ERROR;
    /* pack and H2D pipelined by node range: pack_ms = wall time of the whole phase minus the
     * wait for the last ranges (h2d_tail), h2d_ms = first to last byte on the stream */
    double t0 = tqb_wall_ms(), tail = 0.0;
    if (tqb_forest_pack_upload(ctx->forest, (const tree_qp_in *const *) &qp_in, &tail)) return TREEQP_UNKNOWN_ERROR;
    ctx->forest->pack_ms = tqb_wall_ms() - t0 - tail;
    ctx->forest->h2d_bytes = (long long) sizeof(double) * ctx->forest->dims.n_in;
    ctx->forest->h2d_bytes = upload_bytes(ctx->forest);
    ctx->uploaded = 1;
    ctx->lam_dirty = 0;
    return TREEQP_OK;
}

return_t treeqp_tdunes_solve_resident(const treeqp_tdunes_opts_t *opts, treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest || !ctx->uploaded) return TREEQP_UNKNOWN_ERROR;
    return_t st = validate_opts(opts);
    if (st != TREEQP_OK) return st;
    tqb_forest *f = ctx->forest;
    if (ctx->lam_dirty)
    {   /* a new dual start was set on the host: send that vector only */
        if (tqb_forest_upload_lambda0(f)) return TREEQP_UNKNOWN_ERROR;
        f->h2d_bytes = (long long) sizeof(double) * f->dims.X;
        ctx->lam_dirty = 0;
    }
    tqb_opts o;
    forward_opts(opts, &o);
    work->lineSearchRestartCounter = 0;
    if (forest_solve(f, &o) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    ctx->solved = 1;
    ctx->last_status = (return_t) f->h_status[0];
    int it = f->h_iter[0];
    work->lsIter = (it > 0 && it <= 256) ? f->ls_iters[it - 1] : 0;
    return ctx->last_status;
}

/* only b of the root's children, r[0] and the root's bounds depend on x0 (tree_qp_common.c:2154-2227):
 * re-pack and re-send the root and its children */
return_t treeqp_tdunes_update_x0(const tree_qp_in *qp_in, treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest || !ctx->uploaded) return TREEQP_UNKNOWN_ERROR;
    tqb_forest *f = ctx->forest;
    const int n = 1 + qp_in->tree[0].nkids;
    double t0 = tqb_wall_ms();
    if (tqb_forest_update_nodes(f, 0, qp_in, 0, n)) return TREEQP_UNKNOWN_ERROR;
    f->pack_ms = tqb_wall_ms() - t0;
    const int last = n >= f->dims.Nn;
    const long long x1 = last ? f->dims.X : f->ht.xoff[n], u1 = last ? f->dims.U : f->ht.uoff[n];
    const long long a1 = last ? f->dims.szA : f->ht.aoff[n], b1 = last ? f->dims.szB : f->ht.boff[n];
    f->h2d_bytes = (long long) sizeof(double) * (a1 + b1 + 6 * x1 + 5 * u1);
    return TREEQP_OK;
}

return_t treeqp_tdunes_warm_start_resident(treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest || !ctx->uploaded) return TREEQP_UNKNOWN_ERROR;
    if (tqb_forest_warm_start(ctx->forest)) return TREEQP_UNKNOWN_ERROR;
    ctx->lam_dirty = 0;
    return TREEQP_OK;
}

/* tree_qp_out_max_KKT_res (tree_qp_common.c:769) of the last solve, computed where the iterates are */
return_t treeqp_tdunes_kkt_resident(treeqp_tdunes_workspace *work, double *kkt)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest || !ctx->uploaded || !kkt) return TREEQP_UNKNOWN_ERROR;
    return tqb_device_kkt(ctx->forest->dev, kkt) ? TREEQP_UNKNOWN_ERROR : TREEQP_OK;
}

return_t treeqp_tdunes_download(tree_qp_out *qp_out, treeqp_tdunes_workspace *work)
{
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest) return TREEQP_UNKNOWN_ERROR;
    tqb_forest *f = ctx->forest;
    if (!tqb_forest_download_direct(f, qp_out, NULL, NULL))
    {   /* one D2H copy into the pinned block, then the host loop over the BLASFEO views */
        if (forest_download(f) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
        double t0 = tqb_wall_ms();
        tqb_forest_unpack(f, 0, qp_out);
        f->pack_ms += tqb_wall_ms() - t0;
    }
    qp_out->info.iter = f->h_iter[0];
    return TREEQP_OK;
}

return_t treeqp_tdunes_solve(const tree_qp_in *qp_in, tree_qp_out *qp_out,
                             const treeqp_tdunes_opts_t *opts, treeqp_tdunes_workspace *work)
{
    const double t_total = tqb_wall_ms();
    return_t st = validate_opts(opts);
    if (st != TREEQP_OK) return st;
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest)
    {
        fprintf(stderr, "[treeqp_b200] treeqp_tdunes_solve: workspace has no device arena (no GPU?)\n");
        return TREEQP_UNKNOWN_ERROR;
    }
    /* The reference iterates work->slambda in place (dual_Newton_tree.c:966): a solve that is not
     * preceded by set_dual_initialization continues from the multipliers of the previous one. */
    const int continue_duals = ctx->solved && !ctx->lam_dirty;
    /* "interface" part 1: BLASFEO views -> pinned SoA -> HBM */
    st = treeqp_tdunes_upload(qp_in, opts, work);
    if (st != TREEQP_OK) return st;
    if (continue_duals && tqb_forest_warm_start(ctx->forest)) return TREEQP_UNKNOWN_ERROR;
    const double t_solver = tqb_wall_ms();
    qp_out->info.interface_time = 1e-3 * (t_solver - t_total);

    st = treeqp_tdunes_solve_resident(opts, work);
    const double t_after = tqb_wall_ms();
    if (st == TREEQP_UNKNOWN_ERROR) return st;
    /* the reference returns early, leaving qp_out untouched, when the line search reports a
     * non-descent direction (dual_Newton_tree.c:1214-1215) */
    if (st == TREEQP_DN_NOT_DESCENT_DIRECTION) return st;
    qp_out->info.solver_time = 1e-3 * (t_after - t_solver);

    if (treeqp_tdunes_download(qp_out, work) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    const double t_end = tqb_wall_ms();
    qp_out->info.interface_time += 1e-3 * (t_end - t_after);
    qp_out->info.total_time = 1e-3 * (t_end - t_total);
#if defined(PROFILE) && PROFILE > 0
    {   /* dual_Newton_tree.c:1168-1226, :1257-1260 with CUDA-event times per Newton iteration */
        treeqp_profiling_t *tm = &work->timings;
        if (tm->num_iter != opts->maxIter)
        {
            fprintf(stderr, "[treeqp_b200] Number of iterations cannot be changed after initializing solver (when PROFILE > 0)\n");
            exit(1);
        }
        const int n = tm->num_iter > 0 ? tm->num_iter : 1;
        double *ph = (double *) calloc(4 * (size_t) n, sizeof(double));
        const int nit = tqb_device_iter_phase_times(ctx->forest->dev, ph, n);
        tqb_timers_fill(tm, ph, n, nit < qp_out->info.iter ? nit : qp_out->info.iter, ctx->forest->ls_iters, qp_out->info.total_time);
        free(ph);
    }
#endif
    if (tqb_verbose())
        fprintf(stderr, "[treeqp_b200] treeqp_tdunes_solve: status %d, %d Newton iterations, %lld kernel launches, gpu %.3f ms, total %.3f ms "
                        "(upload %.3f, solve %.3f, download %.3f)\n",
                (int) st, qp_out->info.iter, ctx->forest->launches, ctx->forest->gpu_ms, t_end - t_total, t_solver - t_total,
                t_after - t_solver, t_end - t_after);
    return st;
}

/* ------------------------------------------------------------------- write_solution_to_txt ---- */

/* same format as write_double_vector_to_txt / write_int_vector_to_txt (utils/utils.c:98-143) */
int tqb_write_doubles_txt(const double *v, long long n, const char *filename)
{
    FILE *f = fopen(filename, "w");
    if (f == NULL)
    {
        printf("Error opening file (%s)\n", filename);
        return 1;
    }
    for (long long i = 0; i < n; i++) fprintf(f, "%.16e\n", v[i]);
    fclose(f);
    return 0;
}

int tqb_write_ints_txt(const int *v, long long n, const char *filename)
{
    FILE *f = fopen(filename, "w");
    if (f == NULL)
    {
        printf("Error opening file (%s)\n", filename);
        return 1;
    }
    for (long long i = 0; i < n; i++) fprintf(f, "%d\n", v[i]);
    fclose(f);
    return 0;
}

void write_solution_to_txt(const tree_qp_in *qp_in, int Np, int iter, struct node *tree,
                           treeqp_tdunes_workspace *work)
{
    (void) Np; (void) tree;
    tqb_ctx *ctx = ctx_of(work);
    if (!ctx || !ctx->forest || !ctx->solved)
    {
        fprintf(stderr, "[treeqp_b200] write_solution_to_txt: no solve on this workspace yet\n");
        return;
    }
    tqb_forest *f = ctx->forest;
    /* the iterates live in HBM: bring x, u, lambda back (one copy) */
    if (forest_download(f) != TREEQP_OK) return;
    const tqb_dims *d = &f->dims;
    const double *x = f->h_out, *u = x + d->X, *lam = u + d->U;
    const int nx0 = qp_in->nx[0];
    tqb_write_doubles_txt(x, d->X, "examples/spring_mass_utils/x_opt.txt");
    tqb_write_doubles_txt(u, d->U, "examples/spring_mass_utils/u_opt.txt");
    tqb_write_doubles_txt(lam + nx0, d->X - nx0, "examples/spring_mass_utils/deltalambda_opt.txt");
    tqb_write_doubles_txt(lam + nx0, d->X - nx0, "examples/spring_mass_utils/lambda_opt.txt");
    tqb_write_ints_txt(&iter, 1, "examples/spring_mass_utils/iter.txt");
#if defined(PROFILE) && PROFILE > 0
    timers_write_to_txt(&work->timings); /* dual_Newton_tree.c:1065-1067 */
#endif
}

/* ------------------------------------------------------------------ PROFILE > 0 guard ---- */

/* The reference's workspace structs grow a `timings` member when the CALLER is compiled with PROFILE > 0
 * (dual_Newton_tree.h:146-148, dual_Newton_scenarios.h:135-137) and its drivers then read arrays that
 * only the reference's solve fills (utils/profiling.c).  This ABI is the PROFILE = 0 one; a PROFILE > 0
 * program reaches for these two symbols right after its solve (examples/...:145, :152, random_qp.c:191),
 * binds them here (the library precedes libtreeqp.a on the link line) and is stopped with a message
 * instead of printing uninitialised timers.  Per-kernel-class GPU times: treeqp_tdunes_get_stats. */
#if !(defined(PROFILE) && PROFILE > 0)
static void profile_unsupported(const char *fn)
{
    fprintf(stderr, "[treeqp_b200] %s: this program was compiled with PROFILE > 0; libtreeqp_b200 implements the "
                    "PROFILE = 0 workspace ABI only (rebuild the caller with -DPROFILE=0; GPU timings: "
                    "treeqp_tdunes_get_stats / treeqp_sdunes_get_stats)\n", fn);
    exit(1);
}
void timers_print(void *timings) { (void) timings; profile_unsupported("timers_print"); }
void timers_write_to_txt(void *timings) { (void) timings; profile_unsupported("timers_write_to_txt"); }
#endif

/* ------------------------------------------------------------------------------- batch ---- */

treeqp_tdunes_batch *treeqp_tdunes_batch_create(int n, const tree_qp_in *const *qp_in,
                                                const treeqp_tdunes_opts_t *opts)
{
    if (n <= 0) return NULL;
    for (int i = 0; i < n; i++) check_stage_solvers(qp_in[i], opts);
    treeqp_tdunes_batch *b = (treeqp_tdunes_batch *) calloc(1, sizeof *b);
    b->n = n;
    b->forest = tqb_forest_create(n, qp_in, (const int *) opts->qp_solver);
    if (!b->forest) { free(b); return NULL; }
    memset(b->forest->hin.lam0, 0, sizeof(double) * (size_t) b->forest->dims.X);
    b->opts_ok = validate_opts(opts) == TREEQP_OK;
    forward_opts(opts, &b->o);
    return b;
}

void treeqp_tdunes_batch_set_dual_initialization(treeqp_tdunes_batch *b, int i, const double *lambda)
{
    if (b && i >= 0 && i < b->n) tqb_forest_set_lambda0(b->forest, i, lambda);
}

return_t treeqp_tdunes_batch_upload(treeqp_tdunes_batch *b, const tree_qp_in *const *qp_in)
{
    double t0 = tqb_wall_ms(), tail = 0.0;
    if (tqb_forest_pack_upload(b->forest, qp_in, &tail)) return TREEQP_UNKNOWN_ERROR;
    b->forest->pack_ms = tqb_wall_ms() - t0 - tail;
    b->forest->h2d_bytes = (long long) sizeof(double) * b->forest->dims.n_in;
    b->uploaded = 1;
    return TREEQP_OK;
}

/* closed-loop step of a batch: only x0 of every QP changed (tree_qp_in_set_x0_* was called on each qp_in) */
return_t treeqp_tdunes_batch_update_x0(treeqp_tdunes_batch *b, const tree_qp_in *const *qp_in)
{
    if (!b || !b->uploaded) return TREEQP_UNKNOWN_ERROR;
    double t0 = tqb_wall_ms();
    long long bytes = 0;
    if (tqb_forest_update_x0_batch(b->forest, qp_in, &bytes)) return TREEQP_UNKNOWN_ERROR;
    b->forest->pack_ms = tqb_wall_ms() - t0;
    b->forest->h2d_bytes = bytes;
    return TREEQP_OK;
}

return_t treeqp_tdunes_batch_solve_resident(treeqp_tdunes_batch *b, return_t *status)
{
    if (!b->opts_ok) return TREEQP_INVALID_OPTION;
    if (!b->uploaded) return TREEQP_UNKNOWN_ERROR;
    if (forest_solve(b->forest, &b->o) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    if (status)
        for (int i = 0; i < b->n; i++) status[i] = (return_t) b->forest->h_status[i];
    return TREEQP_OK;
}

return_t treeqp_tdunes_batch_download(treeqp_tdunes_batch *b, tree_qp_out *const *qp_out)
{
    if (forest_download(b->forest) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    double t0 = tqb_wall_ms();
#pragma omp parallel for schedule(dynamic, 16) if (b->n > 32)
    for (int i = 0; i < b->n; i++)
    {
        if (b->forest->h_status[i] == TREEQP_DN_NOT_DESCENT_DIRECTION) continue; /* untouched */
        tqb_forest_unpack(b->forest, i, qp_out[i]);
        qp_out[i]->info.iter = b->forest->h_iter[i];
    }
    b->forest->pack_ms += tqb_wall_ms() - t0;
    return TREEQP_OK;
}

return_t treeqp_tdunes_solve_batch(treeqp_tdunes_batch *b, const tree_qp_in *const *qp_in,
                                   tree_qp_out *const *qp_out, return_t *status)
{
    return_t st = treeqp_tdunes_batch_upload(b, qp_in);
    if (st != TREEQP_OK) return st;
    st = treeqp_tdunes_batch_solve_resident(b, status);
    if (st != TREEQP_OK) return st;
    return treeqp_tdunes_batch_download(b, qp_out);
}

return_t treeqp_tdunes_batch_kkt_resident(treeqp_tdunes_batch *b, double *kkt)
{
    if (!b || !b->uploaded || !kkt) return TREEQP_UNKNOWN_ERROR;
    return tqb_device_kkt(b->forest->dev, kkt) ? TREEQP_UNKNOWN_ERROR : TREEQP_OK;
}

void treeqp_tdunes_batch_get_stats(const treeqp_tdunes_batch *b, treeqp_b200_stats *s)
{
    memset(s, 0, sizeof *s);
    const tqb_forest *f = b->forest;
    s->iter = f->iter_max;
    s->ls_total = f->ls_total;
    memcpy(s->ls_iters, f->ls_iters, sizeof s->ls_iters);
    s->n_reg = f->n_reg;
    s->gpu_ms = f->gpu_ms; s->h2d_ms = f->h2d_ms; s->d2h_ms = f->d2h_ms; s->pack_ms = f->pack_ms;
    s->h2d_bytes = f->h2d_bytes; s->d2h_bytes = f->d2h_bytes;
    s->kernel_launches = f->launches;
    memcpy(s->kernel_ms, f->kernel_ms, sizeof s->kernel_ms);
    memcpy(s->kernel_count, f->kernel_count, sizeof s->kernel_count);
}

void treeqp_tdunes_batch_destroy(treeqp_tdunes_batch *b)
{
    if (!b) return;
    tqb_forest_destroy(b->forest);
    free(b);
}

/* ------------------------------------------------------------- one tree over several GPUs ---- */

struct treeqp_b200_comm_ { tqb_comm *c; };

int treeqp_b200_comm_unique_id(char *id128) { return tqb_comm_unique_id(id128); }

treeqp_b200_comm *treeqp_b200_comm_create(int rank, int nranks, const char *id128, int device)
{
    tqb_comm *c = tqb_comm_create(rank, nranks, id128, device);
    if (!c) return NULL;
    treeqp_b200_comm *w = (treeqp_b200_comm *) calloc(1, sizeof *w);
    w->c = c;
    return w;
}

void treeqp_b200_comm_destroy(treeqp_b200_comm *comm)
{
    if (!comm) return;
    tqb_comm_destroy(comm->c);
    free(comm);
}

struct treeqp_tdunes_sharded_
{
    tqb_shard_plan *plan;
    tqb_forest *forest;
    tqb_opts o;
    int opts_ok, uploaded, lam_dirty;
    return_t last_status;
};

int treeqp_b200_shard_plan(const tree_qp_in *qp_in, int rank, int nranks, int *cut_depth, int *ncut, int *n_local,
                           int *owned_global, int *ghost_global)
{
    tqb_shard_plan *p = tqb_shard_plan_create(qp_in, rank, nranks);
    if (!p) return 1;
    if (cut_depth) *cut_depth = p->cut_depth;
    if (ncut) *ncut = p->ncut;
    if (n_local) *n_local = p->Nl;
    if (owned_global) memset(owned_global, 0, sizeof(int) * (size_t) qp_in->N);
    if (ghost_global) memset(ghost_global, 0, sizeof(int) * (size_t) qp_in->N);
    for (int j = 0; j < p->Nl; j++)
    {
        if (owned_global && p->owned[j]) owned_global[p->gmap[j]] = 1;
        if (ghost_global && p->ghost[j]) ghost_global[p->gmap[j]] = 1;
    }
    tqb_shard_plan_destroy(p);
    return 0;
}

treeqp_tdunes_sharded *treeqp_tdunes_sharded_create(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                                                    treeqp_b200_comm *comm)
{
    if (!comm) return NULL;
    check_stage_solvers(qp_in, opts);
    for (int k = 0; k < qp_in->N; k++)
        if (opts->qp_solver[k] != TREEQP_CLIPPING_SOLVER)
        {
            fprintf(stderr, "[treeqp_b200] subtree sharding supports the clipping stage solver only\n");
            return NULL;
        }
    tqb_shard_plan *plan = tqb_shard_plan_create(qp_in, tqb_comm_rank(comm->c), tqb_comm_size(comm->c));
    if (!plan) return NULL;
    treeqp_tdunes_sharded *sh = (treeqp_tdunes_sharded *) calloc(1, sizeof *sh);
    sh->plan = plan;
    sh->forest = tqb_forest_create_sharded(qp_in, plan, comm->c);
    if (!sh->forest) { tqb_shard_plan_destroy(plan); free(sh); return NULL; }
    memset(sh->forest->hin.lam0, 0, sizeof(double) * (size_t) sh->forest->dims.X);
    sh->opts_ok = validate_opts(opts) == TREEQP_OK;
    forward_opts(opts, &sh->o);
    sh->lam_dirty = 1;
    return sh;
}

void treeqp_tdunes_sharded_set_dual_initialization(treeqp_tdunes_sharded *sh, const double *lambda,
                                                   const tree_qp_in *qp_in)
{
    tqb_forest_set_lambda0_sharded(sh->forest, lambda, qp_in, sh->plan);
    sh->lam_dirty = 1;
}

return_t treeqp_tdunes_sharded_upload(treeqp_tdunes_sharded *sh, const tree_qp_in *qp_in)
{
    /* pack and H2D pipelined by node range, as the single-GPU upload */
    double t0 = tqb_wall_ms(), tail = 0.0;
    if (tqb_forest_pack_upload_mapped(sh->forest, qp_in, sh->plan->gmap, &tail)) return TREEQP_UNKNOWN_ERROR;
    sh->forest->pack_ms = tqb_wall_ms() - t0 - tail;
    sh->forest->h2d_bytes = upload_bytes(sh->forest);
    sh->uploaded = 1;
    sh->lam_dirty = 0;
    return TREEQP_OK;
}

return_t treeqp_tdunes_sharded_solve_resident(treeqp_tdunes_sharded *sh)
{
    if (!sh->opts_ok) return TREEQP_INVALID_OPTION;
    if (!sh->uploaded) return TREEQP_UNKNOWN_ERROR;
    if (sh->lam_dirty)
    {   /* a new dual start was set on the host: send that vector only (after a DMA-mode upload the pinned block does
         * not even hold the other inputs of the nodes that travelled straight from the caller's buffer) */
        if (tqb_forest_upload_lambda0(sh->forest)) return TREEQP_UNKNOWN_ERROR;
        sh->forest->h2d_bytes = (long long) sizeof(double) * sh->forest->dims.X;
        sh->lam_dirty = 0;
    }
    if (forest_solve(sh->forest, &sh->o) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    sh->last_status = (return_t) sh->forest->h_status[0];
    return sh->last_status;
}

return_t treeqp_tdunes_sharded_download(treeqp_tdunes_sharded *sh, tree_qp_out *qp_out)
{
    if (!tqb_forest_download_direct(sh->forest, qp_out, sh->plan->gmap, sh->plan->owned))
    {
        if (forest_download(sh->forest) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
        double t0 = tqb_wall_ms();
        tqb_forest_unpack_sharded(sh->forest, sh->plan, qp_out);
        sh->forest->pack_ms += tqb_wall_ms() - t0;
    }
    qp_out->info.iter = sh->forest->h_iter[0];
    return TREEQP_OK;
}

return_t treeqp_tdunes_sharded_solve(treeqp_tdunes_sharded *sh, const tree_qp_in *qp_in, tree_qp_out *qp_out)
{
    return_t st = treeqp_tdunes_sharded_upload(sh, qp_in);
    if (st != TREEQP_OK) return st;
    st = treeqp_tdunes_sharded_solve_resident(sh);
    if (st == TREEQP_UNKNOWN_ERROR || st == TREEQP_INVALID_OPTION || st == TREEQP_DN_NOT_DESCENT_DIRECTION) return st;
    if (treeqp_tdunes_sharded_download(sh, qp_out) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    return st;
}

void treeqp_tdunes_sharded_get_stats(const treeqp_tdunes_sharded *sh, treeqp_b200_stats *s)
{
    treeqp_tdunes_batch tmp;
    memset(&tmp, 0, sizeof tmp);
    tmp.forest = sh->forest;
    treeqp_tdunes_batch_get_stats(&tmp, s);
}

void treeqp_tdunes_sharded_get_plan(const treeqp_tdunes_sharded *sh, int *cut_depth, int *ncut, int *n_local,
                                    int *n_owned)
{
    const tqb_shard_plan *p = sh->plan;
    if (cut_depth) *cut_depth = p->cut_depth;
    if (ncut) *ncut = p->ncut;
    if (n_local) *n_local = p->Nl;
    if (n_owned)
    {
        int n = 0;
        for (int j = 0; j < p->Nl; j++) n += p->owned[j];
        *n_owned = n;
    }
}

void treeqp_tdunes_sharded_destroy(treeqp_tdunes_sharded *sh)
{
    if (!sh) return;
    tqb_forest_destroy(sh->forest);
    tqb_shard_plan_destroy(sh->plan);
    free(sh);
}
