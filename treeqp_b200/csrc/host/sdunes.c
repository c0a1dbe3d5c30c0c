/*
 * sdunes.c -- the reference's SDUNES C API (treeqp/src/dual_Newton_scenarios.h:143-163) on top
 * of the persistent CUDA kernel in cuda/sdunes_device.cu, plus the additive device-resident entry points.
 *
 * Host work: scenario bookkeeping (nodeIdx, boundsRemoved, commonNodes: scenarios.c:262-291,
 * 1607-1668), un-scaling of the weights (scenarios.c:1885-1898), BLASFEO views <-> flat blocks.
 * Ownership follows the reference: the caller owns the buffer handed to treeqp_sdunes_create; the
 * host-visible members the reference's callers read (Ns, Nh, Nr, md, sx/su/smu/slambda views,
 * examples/solve_qp_json.cpp:165-188) live there and are valid after every solve.  The device
 * arena is an addition released by treeqp_sdunes_destroy.
 */
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "treeqp_b200.h"
#include "host_internal.h"

#define TQB_SD_MAGIC 0x5D0B200u

typedef struct tqb_sd_ctx
{
    unsigned magic;
    tqb_sd_dims dims;
    int *h_int;
    tqb_sd_tables ht;
    double *h_in;   /* pinned */
    tqb_sd_inputs hin;
    double *h_out;  /* pinned */
    tqb_sd_outputs hout;
    int *h_info;    /* pinned */
    int *node_nx, *node_nu; /* actual per-node dimensions (root nx = 0, leaves nu = 0) */
    tqb_sd_device *dev;
    int uploaded, dual_dirty;
    int solved; /* a Newton loop has run: hout.lam_s / mu_s hold its final multipliers */
    /* statistics of the last solve */
    double gpu_ms, h2d_ms, d2h_ms, pack_ms;
    long long launches;
} tqb_sd_ctx;

static long long pad2(long long n) { return (n + 1) / 2 * 2; }

void tqb_sd_layout_tables(const tqb_sd_dims *d, int *base, tqb_sd_tables *t)
{
    int *p = base;
#define TK(field, n) t->field = p; p += (n)
    TK(nodeIdx, d->Ns * (d->Nh + 1)); TK(removed, d->Ns * (d->Nh + 1));
    TK(cn, d->Ns); TK(lo, d->Ns + 1); TK(jdo, d->Ns + 1); TK(jlo, d->Ns + 1);
    TK(pair_of, d->nl > 0 ? d->nl : 1);
    TK(nfirst, d->Nn); TK(ncount, d->Nn); TK(nstage, d->Nn);
#undef TK
    assert(p - base == d->n_int);
}

static int sd_tables_size(const tqb_sd_dims *d)
{
    return 2 * d->Ns * (d->Nh + 1) + d->Ns + 3 * (d->Ns + 1) + (d->nl > 0 ? d->nl : 1) + 3 * d->Nn;
}

void tqb_sd_layout_inputs(const tqb_sd_dims *d, double *base, tqb_sd_inputs *in)
{
    double *p = base;
    const long long Nn = d->Nn, nx = d->nx, nu = d->nu;
#define TK(field, n) in->field = p; p += pad2(n)
    TK(A, Nn * nx * nx); TK(B, Nn * nx * nu); TK(b, Nn * nx);
    TK(Q, Nn * nx); TK(q, Nn * nx); TK(Qinv, Nn * nx); TK(xmin, Nn * nx); TK(xmax, Nn * nx);
    TK(R, Nn * nu); TK(r, Nn * nu); TK(Rinv, Nn * nu); TK(umin, Nn * nu); TK(umax, Nn * nu);
    TK(lam0, d->nl); TK(mu0, (long long) d->SK * nx);
#undef TK
    assert(p - base == d->n_in);
}

static long long sd_inputs_size(const tqb_sd_dims *d)
{
    const long long Nn = d->Nn, nx = d->nx, nu = d->nu;
    return pad2(Nn * nx * nx) + pad2(Nn * nx * nu) + 6 * pad2(Nn * nx) + 5 * pad2(Nn * nu) + pad2(d->nl) +
           pad2((long long) d->SK * nx);
}

void tqb_sd_layout_outputs(const tqb_sd_dims *d, double *base, tqb_sd_outputs *out)
{
    double *p = base;
    const long long Nn = d->Nn, nx = d->nx, nu = d->nu;
#define TK(field, n) out->field = p; p += pad2(n)
    TK(x, Nn * nx); TK(u, Nn * nu); TK(lam, Nn * nx); TK(mux, Nn * nx); TK(muu, Nn * nu);
    TK(lam_s, d->nl); TK(mu_s, (long long) d->SK * nx); TK(xs, (long long) d->SK * nx); TK(us, (long long) d->SK * nu);
#undef TK
    assert(p - base == d->n_out);
}

static long long sd_outputs_size(const tqb_sd_dims *d)
{
    const long long Nn = d->Nn, nx = d->nx, nu = d->nu;
    return 3 * pad2(Nn * nx) + 2 * pad2(Nn * nu) + pad2(d->nl) + 2 * pad2((long long) d->SK * nx) +
           pad2((long long) d->SK * nu);
}

/* ----------------------------------------------------------------------------- options ---- */

int treeqp_sdunes_opts_calculate_size(int Nn) { (void) Nn; return 0; }

void treeqp_sdunes_opts_create(int Nn, treeqp_sdunes_opts_t *opts, void *ptr)
{
    (void) Nn; (void) opts; (void) ptr; /* no pointers in the options (scenarios.c:73-76) */
}

/* defaults of scenarios.c:80-95 */
void treeqp_sdunes_opts_set_default(int Nn, treeqp_sdunes_opts_t *opts)
{
    (void) Nn;
    opts->maxIter = 100;
    opts->termCondition = TREEQP_INFNORM;
    opts->stationarityTolerance = 1.0e-8;
    opts->checkLastActiveSet = 1;
    opts->lineSearchMaxIter = 50;
    opts->lineSearchGamma = 0.1;
    opts->lineSearchBeta = 0.6;
    opts->regType = TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT;
    opts->regTol = 1.0e-6;
    opts->regValue = 1.0e-6;
}

static int ipow_i(int b, int e)
{
    int r = 1;
    for (int t = 0; t < e; t++) r *= b;
    return r;
}

/* scenarios.c:99-107 */
int treeqp_sdunes_calculate_dual_dimension(int Nr, int md, int nu)
{
    int Ns = ipow_i(md, Nr);
    if (Ns == 1) return -1;
    return (Nr * Ns - (Ns - 1) / (md - 1)) * nu;
}

/* --------------------------------------------------------------------------- workspace ---- */

typedef struct sd_shape
{
    int Nn, Np, Ns, Nh, Nr, md, nx, nu;
} sd_shape;

static void shape_of(const tree_qp_in *qp, sd_shape *s)
{
    const struct node *tree = qp->tree;
    s->Nn = qp->N;
    s->nx = qp->N > 1 ? qp->nx[1] : 0;
    s->nu = qp->nu[0];
    s->Nh = tree[s->Nn - 1].stage;
    s->Np = 0;
    for (int k = 0; k < s->Nn; k++)
        if (tree[k].nkids > 0) s->Np++;
    s->Ns = s->Nn - s->Np;
    s->Nr = 0; /* utils/tree.c:65-81 */
    for (int k = 0; k < s->Nn; k++)
    {
        if (tree[k].nkids > 1) s->Nr = tree[k].stage + 1;
        else break;
    }
    s->md = tree[0].nkids;
}

static long long sd_host_bytes(const sd_shape *s)
{
    long long b = 0;
    const long long SK = (long long) s->Ns * s->Nh;
    b += (long long) (s->Ns > 1 ? s->Ns - 1 : 1) * sizeof(int);                 /* commonNodes */
    b += (long long) s->Ns * sizeof(double);                                     /* fvals */
    b += 2LL * (s->Ns * sizeof(int *) + (long long) s->Ns * (s->Nh + 1) * sizeof(int)); /* nodeIdx, boundsRemoved */
    b += 3LL * (s->Ns * sizeof(struct blasfeo_dvec *) + SK * sizeof(struct blasfeo_dvec)); /* sx, su, smu */
    b += (long long) s->Ns * sizeof(struct blasfeo_dvec);                        /* slambda */
    b += 8 * 64;
    return (b + 63) / 64 * 64;
}

int treeqp_sdunes_calculate_size(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts)
{
    sd_shape s;
    shape_of(qp_in, &s);
    long long b = sd_host_bytes(&s);
#if defined(PROFILE) && PROFILE > 0
    b += timers_calculate_size(opts->maxIter) + 64; /* dual_Newton_scenarios.c (PROFILE > 0 branch of calculate_size) */
#else
    (void) opts;
#endif
    return (int) b;
}

static char *align64(char *p) { return (char *) (((uintptr_t) p + 63u) / 64u * 64u); }

static tqb_sd_ctx *sd_ctx_of(const treeqp_sdunes_workspace *work)
{
    if (!work) return NULL;
    tqb_sd_ctx *ctx = (tqb_sd_ctx *) (void *) work->sTmpMats;
    if (!ctx || ctx->magic != TQB_SD_MAGIC) return NULL;
    return ctx;
}

static void sd_ctx_free(tqb_sd_ctx *ctx)
{
    if (!ctx || ctx->magic != TQB_SD_MAGIC) return;
    if (ctx->dev) tqb_sd_device_destroy(ctx->dev);
    if (ctx->h_in) tqb_pinned_free(ctx->h_in);
    if (ctx->h_out) tqb_pinned_free(ctx->h_out);
    if (ctx->h_info) tqb_pinned_free(ctx->h_info);
    free(ctx->h_int); free(ctx->node_nx); free(ctx->node_nu);
    ctx->magic = 0;
    free(ctx);
}

void treeqp_sdunes_create(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work, void *ptr)
{
    (void) opts;
    sd_shape s;
    shape_of(qp_in, &s);
    const struct node *tree = qp_in->tree;
    const int Nn = s.Nn, Ns = s.Ns, Nh = s.Nh, nx = s.nx, nu = s.nu;
    memset(work, 0, sizeof *work);
    work->Ns = Ns; work->Nh = Nh; work->Nr = s.Nr; work->md = s.md;

    /* preconditions of the reference (scenarios.c:1498-1504, 1630): x0 eliminated, uniform
     * dimensions, diagonal weights, bounds only */
    int ok = qp_in->nx[0] == 0;
    for (int k = 1; k < Nn && ok; k++) ok = qp_in->nx[k] == nx;
    for (int k = 0; k < Nn && ok; k++) ok = (tree[k].nkids > 0 ? qp_in->nu[k] == nu : qp_in->nu[k] == 0);
    for (int k = 0; k < Nn && ok; k++) ok = tqb_clipping_applicable(qp_in, k) == YES;
    if (!ok)
    {
        printf("[TREEQP_B200]: Error! SDUNES needs x0 eliminated, uniform nx/nu, diagonal weights and bounds only.\n");
        exit(1);
    }

    /* ---- host-visible members carved from the caller's buffer ---- */
    char *cur = align64((char *) ptr);
    work->commonNodes = (int *) cur; cur += (size_t) (Ns > 1 ? Ns - 1 : 1) * sizeof(int);
    cur = align64(cur);
    work->fvals = (double *) cur; cur += (size_t) Ns * sizeof(double);
    work->nodeIdx = (int **) cur; cur += (size_t) Ns * sizeof(int *);
    work->boundsRemoved = (int **) cur; cur += (size_t) Ns * sizeof(int *);
    for (int i = 0; i < Ns; i++) { work->nodeIdx[i] = (int *) cur; cur += (size_t) (Nh + 1) * sizeof(int); }
    for (int i = 0; i < Ns; i++) { work->boundsRemoved[i] = (int *) cur; cur += (size_t) (Nh + 1) * sizeof(int); }
    cur = align64(cur);
    struct blasfeo_dvec ***views[3] = {&work->sx, &work->su, &work->smu};
    for (int f = 0; f < 3; f++)
    {
        *views[f] = (struct blasfeo_dvec **) cur; cur += (size_t) Ns * sizeof(struct blasfeo_dvec *);
        for (int i = 0; i < Ns; i++) { (*views[f])[i] = (struct blasfeo_dvec *) cur; cur += (size_t) Nh * sizeof(struct blasfeo_dvec); }
    }
    work->slambda = (struct blasfeo_dvec *) cur; cur += (size_t) Ns * sizeof(struct blasfeo_dvec);
    assert(cur <= (char *) ptr + sd_host_bytes(&s));

    /* scenario bookkeeping (scenarios.c:1650-1668, 262-291) */
    char *seen = (char *) calloc((size_t) Nn, 1);
    for (int i = 0; i < Ns; i++)
    {
        int node = Nn - Ns + i;
        work->nodeIdx[i][Nh] = node;
        work->boundsRemoved[i][Nh] = 0;
        for (int k = Nh; k > 0; k--)
        {
            node = tree[node].dad;
            work->nodeIdx[i][k - 1] = node;
            work->boundsRemoved[i][k - 1] = seen[node] ? 1 : 0;
            seen[node] = 1;
        }
        work->fvals[i] = 0.0;
    }
    free(seen);
    for (int i = 0; i < Ns - 1; i++)
    {
        int n1 = Nn - Ns + i, n2 = n1 + 1, c = -1;
        for (int k = Nh; k >= 0; k--)
        {
            if (n1 == n2) { c = k + 1; break; }
            n1 = tree[n1].dad; n2 = tree[n2].dad;
        }
        work->commonNodes[i] = c;
    }

    /* ---- context: dimensions, tables, pinned staging, device arena ---- */
    tqb_sd_ctx *ctx = (tqb_sd_ctx *) calloc(1, sizeof(tqb_sd_ctx));
    ctx->magic = TQB_SD_MAGIC;
    work->sTmpMats = (struct blasfeo_dmat *) (void *) ctx;
    tqb_sd_dims *d = &ctx->dims;
    d->Nn = Nn; d->Np = s.Np; d->Ns = Ns; d->Nh = Nh; d->Nr = s.Nr; d->md = s.md; d->nx = nx; d->nu = nu;
    d->nK = nu * s.Nr; d->SK = Ns * Nh;
    int nl = 0, njd = 0, njl = 0, mJ = 0;
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * work->commonNodes[i];
        nl += dim; njd += dim * dim;
        if (i < Ns - 2) njl += nu * work->commonNodes[i + 1] * dim;
        if (dim > mJ) mJ = dim;
    }
    d->nl = nl; d->njd = njd; d->njl = njl; d->mJ = mJ;
    d->n_int = sd_tables_size(d);
    d->n_in = sd_inputs_size(d);
    d->n_out = sd_outputs_size(d);

    ctx->h_int = (int *) calloc((size_t) d->n_int, sizeof(int));
    tqb_sd_layout_tables(d, ctx->h_int, &ctx->ht);
    tqb_sd_tables *T = &ctx->ht;
    ctx->node_nx = (int *) malloc(sizeof(int) * (size_t) Nn);
    ctx->node_nu = (int *) malloc(sizeof(int) * (size_t) Nn);
    for (int k = 0; k < Nn; k++)
    {
        ctx->node_nx[k] = qp_in->nx[k]; ctx->node_nu[k] = qp_in->nu[k];
        T->nstage[k] = tree[k].stage; T->nfirst[k] = -1; T->ncount[k] = 0;
    }
    int lo = 0, jdo = 0, jlo = 0;
    for (int i = 0; i < Ns; i++)
    {
        for (int k = 0; k <= Nh; k++)
        {
            const int node = work->nodeIdx[i][k];
            T->nodeIdx[i * (Nh + 1) + k] = node;
            T->removed[i * (Nh + 1) + k] = work->boundsRemoved[i][k];
            if (T->nfirst[node] < 0) T->nfirst[node] = i;
            T->ncount[node]++;
        }
        T->lo[i] = lo; T->jdo[i] = jdo; T->jlo[i] = jlo;
        T->cn[i] = i < Ns - 1 ? work->commonNodes[i] : 0;
        if (i < Ns - 1)
        {
            const int dim = nu * work->commonNodes[i];
            for (int e = 0; e < dim; e++) T->pair_of[lo + e] = i;
            lo += dim; jdo += dim * dim;
            if (i < Ns - 2) jlo += nu * work->commonNodes[i + 1] * dim;
        }
    }
    T->lo[Ns] = lo; T->jdo[Ns] = jdo; T->jlo[Ns] = jlo;

    ctx->h_in = (double *) tqb_pinned_alloc(sizeof(double) * (size_t) d->n_in);
    ctx->h_out = (double *) tqb_pinned_alloc(sizeof(double) * (size_t) d->n_out);
    ctx->h_info = (int *) tqb_pinned_alloc(sizeof(int) * TQB_SD_INFO_INTS);
    if (ctx->h_in && ctx->h_out && ctx->h_info)
    {
        memset(ctx->h_in, 0, sizeof(double) * (size_t) d->n_in);
        memset(ctx->h_out, 0, sizeof(double) * (size_t) d->n_out);
        tqb_sd_layout_inputs(d, ctx->h_in, &ctx->hin);
        tqb_sd_layout_outputs(d, ctx->h_out, &ctx->hout);
        ctx->dev = tqb_sd_device_create(d, ctx->h_int);
    }
    if (!ctx->dev)
        fprintf(stderr, "[treeqp_b200] treeqp_sdunes_create: no device arena (%s); solve will return "
                        "TREEQP_UNKNOWN_ERROR\n", tqb_cuda_last_error());
    ctx->dual_dirty = 1;
#if defined(PROFILE) && PROFILE > 0
    {   /* the timers live behind the workspace's own tables */
        char *tp = align64((char *) ptr + sd_host_bytes(&s));
        timers_create(opts->maxIter, &work->timings, tp);
        timers_initialize(&work->timings);
    }
#endif
    /* no destroy in the reference API: a re-create over the same caller memory releases the previous arena */
    tqb_registry_replace(ptr, ctx, (void (*)(void *)) sd_ctx_free);

    /* views over the pinned output block: valid after every solve (what solve_qp_json.cpp reads) */
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {
            struct blasfeo_dvec *vx = &work->sx[i][k], *vu = &work->su[i][k], *vm = &work->smu[i][k];
            vx->m = nx; vx->pm = nx; vx->memsize = 0;
            vu->m = nu; vu->pm = nu; vu->memsize = 0;
            vm->m = nx; vm->pm = nx; vm->memsize = 0;
            vx->pa = ctx->h_out ? ctx->hout.xs + ((size_t) i * Nh + k) * nx : NULL;
            vu->pa = ctx->h_out ? ctx->hout.us + ((size_t) i * Nh + k) * nu : NULL;
            vm->pa = ctx->h_out ? ctx->hout.mu_s + ((size_t) i * Nh + k) * nx : NULL;
        }
    for (int i = 0; i < Ns - 1; i++)
    {
        struct blasfeo_dvec *vl = &work->slambda[i];
        vl->m = nu * work->commonNodes[i]; vl->pm = vl->m; vl->memsize = 0;
        vl->pa = ctx->h_out ? ctx->hout.lam_s + T->lo[i] : NULL;
    }
}

void treeqp_sdunes_destroy(treeqp_sdunes_workspace *work)
{
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx) return;
    tqb_registry_remove(ctx);
    sd_ctx_free(ctx);
    work->sTmpMats = NULL;
}

/* scenarios.c:2094-2117: lam = concatenation over the scenario pairs, mu over (scenario, stage) */
void treeqp_sdunes_set_dual_initialization(double *lam, double *mu, treeqp_sdunes_workspace *work)
{
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->h_in) return;
    const tqb_sd_dims *d = &ctx->dims;
    if (d->nl > 0) memcpy(ctx->hin.lam0, lam, sizeof(double) * (size_t) d->nl);
    memcpy(ctx->hin.mu0, mu, sizeof(double) * (size_t) d->SK * d->nx);
    /* the reference writes the start into work->slambda / work->smu: keep those views current */
    if (d->nl > 0) memcpy(ctx->hout.lam_s, lam, sizeof(double) * (size_t) d->nl);
    memcpy(ctx->hout.mu_s, mu, sizeof(double) * (size_t) d->SK * d->nx);
    ctx->dual_dirty = 1;
}

void treeqp_sdunes_get_duals(double *lam, double *mu, const treeqp_sdunes_workspace *work)
{
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->h_out) return;
    const tqb_sd_dims *d = &ctx->dims;
    if (d->nl > 0) memcpy(lam, ctx->hout.lam_s, sizeof(double) * (size_t) d->nl);
    memcpy(mu, ctx->hout.mu_s, sizeof(double) * (size_t) d->SK * d->nx);
}

void treeqp_sdunes_get_stats(const treeqp_sdunes_workspace *work, treeqp_b200_stats *s)
{
    memset(s, 0, sizeof *s);
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->h_info) return;
    s->iter = ctx->h_info[1];
    s->ls_total = ctx->h_info[2];
    s->n_reg = ctx->h_info[3];
    memcpy(s->ls_iters, ctx->h_info + 8, sizeof s->ls_iters);
    s->gpu_ms = ctx->gpu_ms; s->h2d_ms = ctx->h2d_ms; s->d2h_ms = ctx->d2h_ms; s->pack_ms = ctx->pack_ms;
    s->h2d_bytes = (long long) sizeof(double) * ctx->dims.n_in;
    s->d2h_bytes = (long long) sizeof(double) * ctx->dims.n_out;
    s->kernel_launches = ctx->launches;
    /* one kernel: the per-"class" slots carry the phase split measured inside it (globaltimer) */
    for (int k = 0; k < TQB_SD_NPHASE; k++)
    {
        s->kernel_ms[k] = 1e-6 * ctx->h_info[8 + 256 + k];
        s->kernel_count[k] = ctx->h_info[8 + 256 + k] > 0 ? 1 : 0;
    }
}

static const char *const sd_phase_names[TQB_SD_NPHASE] = {
    "sd_stage_solve", "sd_build+norm", "sd_factor_scenario", "sd_jay_chain", "sd_delta_mu+dot", "sd_fval(0)",
    "sd_line_search", "sd_step_barrier", "sd_export", ""};
int treeqp_sdunes_num_phases(void) { return TQB_SD_NPHASE - 1; }
const char *treeqp_sdunes_phase_name(int k) { return (k >= 0 && k < TQB_SD_NPHASE) ? sd_phase_names[k] : ""; }

/* ------------------------------------------------------------------------------- solve ---- */

/* BLASFEO views -> node slots; weights un-scaled exactly as scenarios.c:1885-1898 */
static void sd_pack(tqb_sd_ctx *ctx, const tree_qp_in *qp)
{
    const tqb_sd_dims *d = &ctx->dims;
    tqb_sd_inputs *in = &ctx->hin;
    const int nx = d->nx, nu = d->nu, Nr = d->Nr, md = d->md;
    const int mdNr = ipow_i(md, Nr);
    for (int k = 0; k < d->Nn; k++)
    {
        const int nxk = qp->nx[k], nuk = qp->nu[k];
        const int st = qp->tree[k].stage < Nr ? qp->tree[k].stage : Nr;
        const double sc = (double) ipow_i(md, st) / mdNr;
        if (k > 0)
        {
            const struct blasfeo_dmat *A = &qp->A[k - 1], *B = &qp->B[k - 1];
            double *Ad = in->A + (size_t) k * nx * nx, *Bd = in->B + (size_t) k * nx * nu;
            for (int j = 0; j < A->n; j++)
                for (int i = 0; i < A->m; i++) Ad[i + (size_t) j * nx] = BLASFEO_DMATEL(A, i, j);
            for (int j = 0; j < B->n; j++)
                for (int i = 0; i < B->m; i++) Bd[i + (size_t) j * nx] = BLASFEO_DMATEL(B, i, j);
            memcpy(in->b + (size_t) k * nx, qp->b[k - 1].pa, sizeof(double) * (size_t) nxk);
        }
        for (int i = 0; i < nxk; i++)
        {
            const size_t g = (size_t) k * nx + i;
            in->Q[g] = sc * BLASFEO_DMATEL(&qp->Q[k], i, i);
            in->q[g] = sc * qp->q[k].pa[i];
            in->Qinv[g] = 1.0 / in->Q[g];
            in->xmin[g] = qp->xmin[k].pa[i];
            in->xmax[g] = qp->xmax[k].pa[i];
        }
        for (int i = 0; i < nuk; i++)
        {
            const size_t g = (size_t) k * nu + i;
            in->R[g] = sc * BLASFEO_DMATEL(&qp->R[k], i, i);
            in->r[g] = sc * qp->r[k].pa[i];
            in->Rinv[g] = 1.0 / in->R[g];
            in->umin[g] = qp->umin[k].pa[i];
            in->umax[g] = qp->umax[k].pa[i];
        }
    }
}

static void forward_sd_opts(const treeqp_sdunes_opts_t *o, tqb_opts *d)
{
    memset(d, 0, sizeof *d);
    d->maxIter = o->maxIter;
    d->termCondition = (int) o->termCondition;
    d->regType = (int) o->regType;
    d->lsMaxIter = o->lineSearchMaxIter;
    d->lsRestartTrigger = -1;
    d->tol = o->stationarityTolerance;
    d->regTol = o->regTol;
    d->regValue = o->regValue;
    d->gamma = o->lineSearchGamma;
    d->beta = o->lineSearchBeta;
}

return_t treeqp_sdunes_upload(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work)
{
    (void) opts;
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->dev) return TREEQP_UNKNOWN_ERROR;
    const double t0 = tqb_wall_ms();
    sd_pack(ctx, qp_in);
    ctx->pack_ms = tqb_wall_ms() - t0;
    if (tqb_sd_device_upload(ctx->dev, ctx->h_in, &ctx->h2d_ms)) return TREEQP_UNKNOWN_ERROR;
    ctx->uploaded = 1;
    ctx->dual_dirty = 0;
    return TREEQP_OK;
}

return_t treeqp_sdunes_solve_resident(treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work)
{
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->dev || !ctx->uploaded) return TREEQP_UNKNOWN_ERROR;
    if (ctx->dual_dirty)
    {   /* the dual start is the tail of the input block: re-send the block */
        if (tqb_sd_device_upload(ctx->dev, ctx->h_in, &ctx->h2d_ms)) return TREEQP_UNKNOWN_ERROR;
        ctx->dual_dirty = 0;
    }
    tqb_opts o;
    forward_sd_opts(opts, &o);
    work->reverseCholesky = opts->checkLastActiveSet;
    if (tqb_sd_device_solve(ctx->dev, &o, opts->checkLastActiveSet, ctx->h_info, &ctx->gpu_ms, &ctx->launches))
        return TREEQP_UNKNOWN_ERROR;
    /* scenarios.c:1858, 2080-2081: UNKNOWN_ERROR unless converged (0) or out of iterations (1) */
    return (return_t) ctx->h_info[0];
}

return_t treeqp_sdunes_download(tree_qp_out *qp_out, treeqp_sdunes_workspace *work)
{
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->dev) return TREEQP_UNKNOWN_ERROR;
    if (tqb_sd_device_download(ctx->dev, ctx->h_out, &ctx->d2h_ms)) return TREEQP_UNKNOWN_ERROR;
    const double t0 = tqb_wall_ms();
    const tqb_sd_dims *d = &ctx->dims;
    const tqb_sd_outputs *o = &ctx->hout;
    const int nx = d->nx, nu = d->nu;
    for (int k = 0; k < d->Nn; k++)
    {
        const int nxk = ctx->node_nx[k], nuk = ctx->node_nu[k];
        memcpy(qp_out->x[k].pa, o->x + (size_t) k * nx, sizeof(double) * (size_t) nxk);
        memcpy(qp_out->mu_x[k].pa, o->mux + (size_t) k * nx, sizeof(double) * (size_t) nxk);
        memcpy(qp_out->u[k].pa, o->u + (size_t) k * nu, sizeof(double) * (size_t) nuk);
        memcpy(qp_out->mu_u[k].pa, o->muu + (size_t) k * nu, sizeof(double) * (size_t) nuk);
        if (k > 0) memcpy(qp_out->lam[k - 1].pa, o->lam + (size_t) k * nx, sizeof(double) * (size_t) nxk);
    }
    qp_out->info.iter = ctx->h_info[1];
    ctx->pack_ms += tqb_wall_ms() - t0;
    return TREEQP_OK;
}

void write_scenarios_solution_to_txt(int Ns, int Nh, int Nr, int md, int nx, int nu, int NewtonIter,
                                     treeqp_sdunes_workspace *work)
{
    (void) Nr; (void) md;
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->h_out || !ctx->solved)
    {
        fprintf(stderr, "[treeqp_b200] write_scenarios_solution_to_txt: no solve on this workspace yet\n");
        return;
    }
    /* scenario iterates and duals of the last solve: the output block treeqp_sdunes_solve downloaded */
    const tqb_sd_dims *d = &ctx->dims;
    const tqb_sd_outputs *o = &ctx->hout;
    if (Ns != d->Ns || Nh != d->Nh || nx != d->nx || nu != d->nu)
        fprintf(stderr, "[treeqp_b200] write_scenarios_solution_to_txt: dimensions differ from the workspace\n");
    tqb_write_doubles_txt(o->lam_s, d->nl, "examples/spring_mass_utils/lambdaIter.txt");
    tqb_write_doubles_txt(o->mu_s, (long long) d->SK * d->nx, "examples/spring_mass_utils/muIter.txt");
    tqb_write_doubles_txt(o->xs, (long long) d->SK * d->nx, "examples/spring_mass_utils/xIter.txt");
    tqb_write_doubles_txt(o->us, (long long) d->SK * d->nu, "examples/spring_mass_utils/uIter.txt");
    tqb_write_ints_txt(&NewtonIter, 1, "examples/spring_mass_utils/iter.txt");
#if defined(PROFILE) && PROFILE > 0
    timers_write_to_txt(&work->timings); /* dual_Newton_scenarios.c:241-243 */
#endif
}

return_t treeqp_sdunes_solve(tree_qp_in *qp_in, tree_qp_out *qp_out, treeqp_sdunes_opts_t *opts,
                             treeqp_sdunes_workspace *work)
{
    const double t_total = tqb_wall_ms();
    tqb_sd_ctx *ctx = sd_ctx_of(work);
    if (!ctx || !ctx->dev)
    {
        fprintf(stderr, "[treeqp_b200] treeqp_sdunes_solve: workspace has no device arena (no GPU?)\n");
        return TREEQP_UNKNOWN_ERROR;
    }
    if (ctx->solved && !ctx->dual_dirty)
    {   /* the reference iterates work->slambda / work->smu in place (scenarios.c:1172-1186): without a
         * new set_dual_initialization the next solve continues from the last multipliers */
        const tqb_sd_dims *d = &ctx->dims;
        if (d->nl > 0) memcpy(ctx->hin.lam0, ctx->hout.lam_s, sizeof(double) * (size_t) d->nl);
        memcpy(ctx->hin.mu0, ctx->hout.mu_s, sizeof(double) * (size_t) d->SK * d->nx);
    }
    return_t st = treeqp_sdunes_upload(qp_in, opts, work);
    if (st != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    const double t_solver = tqb_wall_ms();
    qp_out->info.interface_time = 1e-3 * (t_solver - t_total);
    st = treeqp_sdunes_solve_resident(opts, work);
    const double t_after = tqb_wall_ms();
    if (st == TREEQP_UNKNOWN_ERROR) return st;
    qp_out->info.solver_time = 1e-3 * (t_after - t_solver);
    if (treeqp_sdunes_download(qp_out, work) != TREEQP_OK) return TREEQP_UNKNOWN_ERROR;
    ctx->solved = 1;
    const double t_end = tqb_wall_ms();
    qp_out->info.interface_time += 1e-3 * (t_end - t_after);
    qp_out->info.total_time = 1e-3 * (t_end - t_total);
#if defined(PROFILE) && PROFILE > 0
    {   /* The whole Newton loop is ONE kernel: its phase split is measured inside it per SOLVE (globaltimer stamps,
         * treeqp_sdunes_get_stats), so the per-iteration entries are that split divided evenly over the iterations.
         * Phases as dual_Newton_scenarios.c names them: stage QPs (incl. the Hessian blocks, as there), build, Newton
         * direction (scenario factorisations, J chain, delta_mu), line search. */
        treeqp_profiling_t *tm = &work->timings;
        const int n = tm->num_iter > 0 ? tm->num_iter : 1, nit = qp_out->info.iter < n ? qp_out->info.iter : n;
        double *ph = (double *) calloc(4 * (size_t) n, sizeof(double));
        const int *ns = ctx->h_info + 8 + 256; /* nanoseconds per phase (sd_phase_names) */
        const double tot[4] = {1e-9 * ns[0], 1e-9 * ns[1], 1e-9 * (ns[2] + ns[3] + ns[4]), 1e-9 * (ns[5] + ns[6] + ns[7])};
        for (int p4 = 0; p4 < 4; p4++)
            for (int i = 0; i < nit; i++) ph[p4 * n + i] = tot[p4] / (nit > 0 ? nit : 1);
        tqb_timers_fill(tm, ph, n, nit, ctx->h_info + 8, qp_out->info.total_time);
        free(ph);
    }
#endif
    if (tqb_verbose())
        fprintf(stderr, "[treeqp_b200] treeqp_sdunes_solve: status %d, %d Newton iterations, %lld kernel launches, gpu %.3f ms, total %.3f ms\n",
                (int) st, qp_out->info.iter, ctx->launches, ctx->gpu_ms, t_end - t_total);
    return st;
}
