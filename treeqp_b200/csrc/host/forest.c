/*
 * forest.c -- host side of the boundary: turns one or many tree_qp_in (BLASFEO panel-major
 * views, arrays of structs per node) into the flat structure-of-arrays the kernels stream, and
 * writes the device results back into tree_qp_out.
 *
 * Replaces the per-node bookkeeping of treeqp_tdunes_create (dual_Newton_tree.c:1411-1648:
 * setup_npar :166, setup_idxpos :177) and the data staging of stage_qp_clipping_init
 * (dual_Newton_tree_clipping.c:149-170: diagonal extraction + reciprocals).
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <limits.h>
#include <string.h>
#include <time.h>

#include "treeqp_b200.h"
#include "host_internal.h"

#define TQB_CHUNK 128 /* nodes per fixed-shape reduction chunk */

double tqb_wall_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double) ts.tv_sec + 1e-6 * (double) ts.tv_nsec;
}

/* carve the integer block; identical code computes host and device pointers */
void tqb_layout_tables(const tqb_dims *d, int *base, tqb_tables *t)
{
    int *p = base;
#define TK(field, n) t->field = p; p += (n)
    TK(nx, d->Nn); TK(nu, d->Nn); TK(dad, d->Nn); TK(kid0, d->Nn); TK(nkids, d->Nn);
    TK(tree, d->Nn); TK(depth, d->Nn); TK(pos, d->Nn); TK(dim, d->Nn);
    TK(xoff, d->Nn); TK(uoff, d->Nn); TK(aoff, d->Nn); TK(boff, d->Nn);
    TK(woff, d->Nn); TK(utoff, d->Nn); TK(hoff, d->Nn);
    TK(ncn, d->Nn); TK(doff, d->Nn); TK(cgoff, d->Nn);
    TK(lvl_nodes, d->Np > 0 ? d->Np : 1);
    TK(tnode0, d->nT + 1);
    TK(chunk_beg, d->nchunk); TK(chunk_end, d->nchunk); TK(chunk_tree, d->nchunk);
    TK(chunk_x0, d->nchunk); TK(chunk_x1, d->nchunk); TK(chunk_u0, d->nchunk); TK(chunk_u1, d->nchunk);
    TK(tchunk0, d->nT + 1);
#undef TK
    assert(p - base == d->n_int);
}

static long long pad2(long long n) { return (n + 1) / 2 * 2; } /* keep 16-byte alignment */

void tqb_layout_inputs(const tqb_dims *d, double *base, tqb_inputs *in)
{
    double *p = base;
#define TK(field, n) in->field = p; p += pad2(n)
    TK(A, d->szA); TK(B, d->szB);
    TK(b, d->X); TK(q, d->X); TK(Qd, d->X); TK(Qinv, d->X); TK(xmin, d->X); TK(xmax, d->X);
    TK(r, d->U); TK(Rd, d->U); TK(Rinv, d->U); TK(umin, d->U); TK(umax, d->U);
    TK(lam0, d->X);
    if (d->has_dense)
    {
        TK(H, d->szH); TK(dxmin, d->X); TK(dxmax, d->X); TK(dumin, d->U); TK(dumax, d->U);
    }
    else
        in->H = in->dxmin = in->dxmax = in->dumin = in->dumax = NULL;
    if (d->NC > 0) { TK(Cg, d->szCg); TK(dmin, d->NC); TK(dmax, d->NC); }
    else in->Cg = in->dmin = in->dmax = NULL;
#undef TK
    assert(p - base == d->n_in);
}

static long long inputs_size(const tqb_dims *d)
{
    long long n = pad2(d->szA) + pad2(d->szB) + 7 * pad2(d->X) + 5 * pad2(d->U);
    if (d->has_dense) n += pad2(d->szH) + 2 * pad2(d->X) + 2 * pad2(d->U);
    if (d->NC > 0) n += pad2(d->szCg) + 2 * pad2(d->NC);
    return n;
}

int tqb_clipping_applicable(const tree_qp_in *qp, int k)
{
    int nx = qp->nx[k], nu = qp->nu[k];
    if (qp->nc[k] > 0) return NO;
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nx; i++)
            if (i != j && BLASFEO_DMATEL(&qp->Q[k], i, j) != 0.0) return NO;
    for (int j = 0; j < nu; j++)
        for (int i = 0; i < nu; i++)
            if (i != j && BLASFEO_DMATEL(&qp->R[k], i, j) != 0.0) return NO;
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nu; i++)
            if (BLASFEO_DMATEL(&qp->S[k], i, j) != 0.0) return NO;
    return YES;
}

tqb_forest *tqb_forest_create(int nT, const tree_qp_in *const *qps, const int *solver)
{
    return tqb_forest_create_ex(nT, qps, NULL, NULL, solver);
}

/* number of reduction chunks of one tree: runs of owned nodes cut into pieces of TQB_CHUNK */
static int count_chunks(int N, const int *owned)
{
    if (!owned) return (N + TQB_CHUNK - 1) / TQB_CHUNK;
    int n = 0, k = 0;
    while (k < N)
    {
        if (!owned[k]) { k++; continue; }
        int e = k;
        while (e < N && owned[e] && e - k < TQB_CHUNK) e++;
        n++;
        k = e;
    }
    return n;
}

tqb_forest *tqb_forest_create_ex(int nT, const tree_qp_in *const *qps, const int *ghost, const int *owned,
                                 const int *solver)
{
    tqb_forest *f = (tqb_forest *) calloc(1, sizeof(tqb_forest));
    tqb_dims *d = &f->dims;
    d->nT = nT;
    long long Nn = 0;
    for (int t = 0; t < nT; t++) Nn += qps[t]->N;
    if (Nn > 2000000000LL) { free(f); return NULL; }
    d->Nn = (int) Nn;

    /* pass 1: counts */
    long long X = 0, U = 0, szA = 0, szB = 0, szW = 0, szUt = 0, szH = 0, NC = 0, szCg = 0;
    int Np = 0, Nh = 0, nchunk = 0;
    for (int t = 0; t < nT; t++)
    {
        const tree_qp_in *qp = qps[t];
        nchunk += count_chunks(qp->N, owned);
        for (int k = 0; k < qp->N; k++)
        {
            const struct node *nd = &qp->tree[k];
            X += qp->nx[k];
            U += qp->nu[k];
            if (solver && solver[k] == TREEQP_QPOASES_SOLVER)
            {
                const int nv = qp->nx[k] + qp->nu[k];
                szH += (long long) nv * nv;
                d->has_dense = 1;
                if (nv > d->nvmax) d->nvmax = nv;
                if (qp->nc[k] > 0)
                {   /* general stage constraints ride on the dense stage QP (dual_Newton_tree_qpoases.c:288-298) */
                    NC += qp->nc[k];
                    szCg += (long long) qp->nc[k] * nv;
                    if (qp->nc[k] > d->ncmax) d->ncmax = qp->nc[k];
                }
            }
            if (nd->stage > Nh) Nh = nd->stage;
            if (k > 0)
            {
                szA += (long long) qp->nx[k] * qp->nx[nd->dad];
                szB += (long long) qp->nx[k] * qp->nu[nd->dad];
            }
            if (nd->nkids > 0)
            {
                long long dim = 0;
                for (int j = 0; j < nd->nkids; j++) dim += qp->nx[nd->kids[j]];
                Np++;
                szW += dim * dim;
                if (k > 0) szUt += dim * qp->nx[k];
            }
        }
    }
    /* every offset table holds 32-bit element offsets, and so does the table block's own size */
    const long long n_int_ll = 19LL * d->Nn + (Np > 0 ? Np : 1) + 2LL * (nT + 1) + 7LL * nchunk;
    if (X > 2000000000LL || U > 2000000000LL || szA > 2000000000LL || szW > 2000000000LL || szB > 2000000000LL ||
        szH > 2000000000LL || szUt > 2000000000LL || szCg > 2000000000LL || n_int_ll > 2000000000LL)
    {
        fprintf(stderr, "[treeqp_b200] problem too large for 32-bit element offsets\n");
        free(f);
        return NULL;
    }
    d->X = (int) X; d->U = (int) U; d->szA = szA; d->szB = szB; d->szW = szW; d->szUt = szUt; d->szH = szH;
    d->Np = Np; d->Nh = Nh; d->nchunk = nchunk;
    d->NC = (int) NC; d->szCg = szCg;
    d->n_int = 19 * d->Nn + (Np > 0 ? Np : 1) + 2 * (nT + 1) + 7 * nchunk;
    d->n_in = inputs_size(d);

    f->h_int = (int *) malloc(sizeof(int) * (size_t) d->n_int);
    tqb_layout_tables(d, f->h_int, &f->ht);
    tqb_tables *T = &f->ht;

    /* pass 2: tables */
    int g0 = 0, xo = 0, uo = 0, ao = 0, bo = 0, wo = 0, uto = 0, c = 0, ho = 0, dofs = 0, cgo = 0;
    int *lvl_count = (int *) calloc((size_t) Nh + 2, sizeof(int));
    for (int t = 0; t < nT; t++)
    {
        const tree_qp_in *qp = qps[t];
        T->tnode0[t] = g0;
        T->tchunk0[t] = c;
        for (int s = 0; s < qp->N;)
        {
            if (owned && !owned[s]) { s++; continue; }
            int e = s;
            while (e < qp->N && (!owned || owned[e]) && e - s < TQB_CHUNK) e++;
            T->chunk_beg[c] = g0 + s;
            T->chunk_end[c] = g0 + e;
            T->chunk_tree[c] = t;
            c++;
            s = e;
        }
        for (int k = 0; k < qp->N; k++)
        {
            const struct node *nd = &qp->tree[k];
            int g = g0 + k;
            T->nx[g] = qp->nx[k];
            T->nu[g] = qp->nu[k];
            T->dad[g] = nd->dad < 0 ? -1 : g0 + nd->dad;
            T->nkids[g] = nd->nkids;
            T->kid0[g] = nd->nkids > 0 ? g0 + nd->kids[0] : -1;
            T->tree[g] = t;
            T->depth[g] = nd->stage;
            T->xoff[g] = xo; xo += qp->nx[k];
            T->uoff[g] = uo; uo += qp->nu[k];
            T->hoff[g] = -1;
            T->ncn[g] = 0; T->doff[g] = dofs; T->cgoff[g] = cgo;
            if (solver && solver[k] == TREEQP_QPOASES_SOLVER)
            {
                T->hoff[g] = ho;
                ho += (qp->nx[k] + qp->nu[k]) * (qp->nx[k] + qp->nu[k]);
                T->ncn[g] = qp->nc[k];
                dofs += qp->nc[k];
                cgo += qp->nc[k] * (qp->nx[k] + qp->nu[k]);
            }
            T->aoff[g] = ao; T->boff[g] = bo;
            if (k > 0)
            {
                ao += qp->nx[k] * qp->nx[nd->dad];
                bo += qp->nx[k] * qp->nu[nd->dad];
            }
            int dim = 0;
            for (int j = 0; j < nd->nkids; j++)
            {
                /* children must be numbered consecutively (breadth-first numbering) */
                assert(nd->kids[j] == nd->kids[0] + j);
                dim += qp->nx[nd->kids[j]];
            }
            T->dim[g] = dim;
            T->woff[g] = wo; T->utoff[g] = uto;
            if (nd->nkids > 0)
            {
                wo += dim * dim;
                if (k > 0) uto += dim * qp->nx[k];
                lvl_count[nd->stage]++;
                if (dim > d->dmax) d->dmax = dim;
                if (nd->nkids > d->nkmax) d->nkmax = nd->nkids;
                int ab = dim * (qp->nx[k] + qp->nu[k]);
                if (ab > d->abmax) d->abmax = ab;
            }
            if (qp->nx[k] > d->nxmax) d->nxmax = qp->nx[k];
            if (qp->nu[k] > d->numax) d->numax = qp->nu[k];
        }
        /* idxpos (dual_Newton_tree.c:177-194) */
        for (int k = 0; k < qp->N; k++)
        {
            const struct node *nd = &qp->tree[k];
            int acc = 0;
            for (int j = 0; j < nd->nkids; j++)
            {
                T->pos[g0 + nd->kids[j]] = acc;
                acc += qp->nx[nd->kids[j]];
            }
            if (k == 0) T->pos[g0] = 0;
        }
        g0 += qp->N;
    }
    T->tnode0[nT] = g0;
    T->tchunk0[nT] = c;
    for (int i = 0; i < nchunk; i++)
    {
        const int e = T->chunk_end[i];
        T->chunk_x0[i] = T->xoff[T->chunk_beg[i]];
        T->chunk_u0[i] = T->uoff[T->chunk_beg[i]];
        T->chunk_x1[i] = e < d->Nn ? T->xoff[e] : d->X;
        T->chunk_u1[i] = e < d->Nn ? T->uoff[e] : d->U;
    }

    /* level lists of parents: counting sort by depth, node order inside a level */
    f->lvl_ptr = (int *) malloc(sizeof(int) * (size_t) (Nh + 2));
    f->lvl_ptr[0] = 0;
    for (int l = 0; l <= Nh; l++) f->lvl_ptr[l + 1] = f->lvl_ptr[l] + lvl_count[l];
    memset(lvl_count, 0, sizeof(int) * (size_t) (Nh + 2));
    for (int g = 0; g < d->Nn; g++)
        if (T->nkids[g] > 0)
        {
            int l = T->depth[g];
            T->lvl_nodes[f->lvl_ptr[l] + lvl_count[l]++] = g;
        }
    free(lvl_count);

    f->h_in = (double *) tqb_pinned_alloc(sizeof(double) * (size_t) d->n_in);
    size_t nout = (size_t) 3 * d->X + 2 * d->U + (size_t) d->NC + 8;
    f->h_out = (double *) tqb_pinned_alloc(sizeof(double) * nout);
    f->h_status = (int *) tqb_pinned_alloc(sizeof(int) * (size_t) (2 * nT + 2));
    if (!f->h_in || !f->h_out || !f->h_status) { tqb_forest_destroy(f); return NULL; }
    f->h_iter = f->h_status + nT;
    memset(f->h_in, 0, sizeof(double) * (size_t) d->n_in);
    tqb_layout_inputs(d, f->h_in, &f->hin);

    f->dev = tqb_device_create(d, f->h_int, f->lvl_ptr, ghost);
    if (!f->dev) { tqb_forest_destroy(f); return NULL; }
    return f;
}

void tqb_forest_destroy(tqb_forest *f)
{
    if (!f) return;
    if (f->dev) tqb_device_destroy(f->dev);
    if (f->h_in) tqb_pinned_free(f->h_in);
    if (f->h_out) tqb_pinned_free(f->h_out);
    if (f->h_status) tqb_pinned_free(f->h_status);
    free(f->lvl_ptr);
    free(f->h_int);
    free(f);
}

/* panel-major -> column-major, 4 rows of a panel at a time */
static void unpack_dmat(const struct blasfeo_dmat *M, double *dst)
{
    const int m = M->m, n = M->n, sda = M->cn;
    const double *pA = M->pA;
    int ip = 0;
    for (; ip + TQB_PS <= m; ip += TQB_PS)
    {   /* a full panel: the 4 rows of column j are contiguous on both sides */
        const double *src = pA + (size_t) ip * sda;
        double *d = dst + ip;
        for (int j = 0; j < n; j++) memcpy(d + (size_t) j * m, src + TQB_PS * j, TQB_PS * sizeof(double));
    }
    if (ip < m)
    {
        const int r = m - ip;
        const double *src = pA + (size_t) ip * sda;
        double *d = dst + ip;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < r; i++) d[i + (size_t) j * m] = src[TQB_PS * j + i];
    }
}

static inline int node_in_run(const tqb_forest *f, int k)
{
    for (int r = 0; r < f->n_in_runs; r++)
        if (k >= f->in_runs[r].k0 && k < f->in_runs[r].k1) return 1;
    return 0;
}

/* gmap == NULL: tree t of the forest is qp itself; otherwise local node j is node gmap[j] of qp */
static void pack_range(tqb_forest *f, int t, const tree_qp_in *qp, const int *gmap, int j0, int j1)
{
    const tqb_tables *T = &f->ht;
    tqb_inputs *in = &f->hin;
    int g0 = T->tnode0[t];
    const int direct = f->in_direct; /* nodes inside a run (local index): only the weights are packed here, the rest travels by strided DMA */
    assert(gmap || T->tnode0[t + 1] - g0 == qp->N);
#pragma omp parallel for schedule(static) if (j1 - j0 > 2048)
    for (int j = j0; j < j1; j++)
    {
        const int k = gmap ? gmap[j] : j;
        int g = g0 + j, nx = qp->nx[k], nu = qp->nu[k], xo = T->xoff[g], uo = T->uoff[g];
        assert(nx == T->nx[g] && nu == T->nu[g]);
        if (direct && node_in_run(f, j))
        {   /* clipping node of a DMA run: the diagonal weights only (the device forms the reciprocals, k_recip) */
            for (int i = 0; i < nx; i++) in->Qd[xo + i] = BLASFEO_DMATEL(&qp->Q[k], i, i);
            for (int i = 0; i < nu; i++) in->Rd[uo + i] = BLASFEO_DMATEL(&qp->R[k], i, i);
            continue;
        }
        if (k > 0)
        {
            unpack_dmat(&qp->A[k - 1], in->A + T->aoff[g]);
            unpack_dmat(&qp->B[k - 1], in->B + T->boff[g]);
            memcpy(in->b + xo, qp->b[k - 1].pa, sizeof(double) * (size_t) nx);
        }
        else
            for (int i = 0; i < nx; i++) in->b[xo + i] = 0.0;
        memcpy(in->q + xo, qp->q[k].pa, sizeof(double) * (size_t) nx);
        memcpy(in->r + uo, qp->r[k].pa, sizeof(double) * (size_t) nu);
        if (T->hoff[g] >= 0)
        {   /* dense stage QP (stage_qp_qpoases_init, dual_Newton_tree_qpoases.c:241-249):
             * H = [Q S'; S R] column-major, real bounds in the d* arrays */
            const int nv = nx + nu;
            double *H = in->H + T->hoff[g];
            for (int j = 0; j < nx; j++)
            {
                for (int i = 0; i < nx; i++) H[i + j * nv] = BLASFEO_DMATEL(&qp->Q[k], i, j);
                for (int i = 0; i < nu; i++) H[nx + i + j * nv] = H[j + (nx + i) * nv] = BLASFEO_DMATEL(&qp->S[k], i, j);
            }
            for (int j = 0; j < nu; j++)
                for (int i = 0; i < nu; i++) H[nx + i + (nx + j) * nv] = BLASFEO_DMATEL(&qp->R[k], i, j);
            memcpy(in->dxmin + xo, qp->xmin[k].pa, sizeof(double) * (size_t) nx);
            memcpy(in->dxmax + xo, qp->xmax[k].pa, sizeof(double) * (size_t) nx);
            memcpy(in->dumin + uo, qp->umin[k].pa, sizeof(double) * (size_t) nu);
            memcpy(in->dumax + uo, qp->umax[k].pa, sizeof(double) * (size_t) nu);
            if (T->ncn[g] > 0)
            {   /* [C D] column-major nc x nv, dmin, dmax (stage_qp_qpoases_init :288-298) */
                const int nc = T->ncn[g];
                double *Cg = in->Cg + T->cgoff[g];
                for (int jc = 0; jc < nx; jc++)
                    for (int a = 0; a < nc; a++) Cg[a + jc * nc] = BLASFEO_DMATEL(&qp->C[k], a, jc);
                for (int jc = 0; jc < nu; jc++)
                    for (int a = 0; a < nc; a++) Cg[a + (nx + jc) * nc] = BLASFEO_DMATEL(&qp->D[k], a, jc);
                memcpy(in->dmin + T->doff[g], qp->dmin[k].pa, sizeof(double) * (size_t) nc);
                memcpy(in->dmax + T->doff[g], qp->dmax[k].pa, sizeof(double) * (size_t) nc);
            }
            for (int i = 0; i < nx; i++)
            {
                in->Qd[xo + i] = 1.0; in->Qinv[xo + i] = 0.0;
                in->xmin[xo + i] = -1e300; in->xmax[xo + i] = 1e300;
            }
            for (int i = 0; i < nu; i++)
            {
                in->Rd[uo + i] = 1.0; in->Rinv[uo + i] = 0.0;
                in->umin[uo + i] = -1e300; in->umax[uo + i] = 1e300;
            }
            continue;
        }
        for (int i = 0; i < nx; i++)
        {
            double w = BLASFEO_DMATEL(&qp->Q[k], i, i);
            in->Qd[xo + i] = w;
            in->Qinv[xo + i] = 1.0 / w;
        }
        for (int i = 0; i < nu; i++)
        {
            double w = BLASFEO_DMATEL(&qp->R[k], i, i);
            in->Rd[uo + i] = w;
            in->Rinv[uo + i] = 1.0 / w;
        }
        memcpy(in->xmin + xo, qp->xmin[k].pa, sizeof(double) * (size_t) nx);
        memcpy(in->xmax + xo, qp->xmax[k].pa, sizeof(double) * (size_t) nx);
        memcpy(in->umin + uo, qp->umin[k].pa, sizeof(double) * (size_t) nu);
        memcpy(in->umax + uo, qp->umax[k].pa, sizeof(double) * (size_t) nu);
    }
}

void tqb_forest_pack_mapped_(tqb_forest *f, int t, const tree_qp_in *qp, const int *gmap)
{
    pack_range(f, t, qp, gmap, 0, f->ht.tnode0[t + 1] - f->ht.tnode0[t]);
}

void tqb_forest_pack(tqb_forest *f, int t, const tree_qp_in *qp) { tqb_forest_pack_mapped_(f, t, qp, NULL); }

/* H2D of the slices of every input array that belong to forest nodes [ga, gb): the offset tables
 * are prefix sums over the node order, so each slice is one contiguous range */
/* with_lam0 = 0: the dual start is left alone (an x0 update must not overwrite a device-resident warm
 * start with the stale host copy) */
static int upload_nodes(tqb_forest *f, int ga, int gb, int with_lam0)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    const tqb_inputs *in = &f->hin;
    const int last = gb >= d->Nn;
    const long long x0 = T->xoff[ga], x1 = last ? d->X : T->xoff[gb];
    const long long u0 = T->uoff[ga], u1 = last ? d->U : T->uoff[gb];
    int rc = 0;
    /* weights, their reciprocals, the dual start (and the dense nodes' bounds) always come from the staging block */
    /* DMA mode: the reciprocals are formed on the device from Qd / Rd (IEEE division: the same bits as the host's
     * 1.0 / w), and a dual start of zeros is a device memset -- 34 MB less on the wire for config 3 */
    const int dm = f->in_direct;
    const double *xs_all[] = {in->Qd, dm ? NULL : in->Qinv, (with_lam0 && !(dm && f->lam0_zero)) ? in->lam0 : NULL, in->dxmin, in->dxmax};
    const double *us_all[] = {in->Rd, dm ? NULL : in->Rinv, in->dumin, in->dumax};
    for (unsigned i = 0; i < sizeof xs_all / sizeof xs_all[0]; i++)
        if (xs_all[i]) rc |= tqb_device_upload_range(f->dev, f->h_in, (xs_all[i] - f->h_in) + x0, x1 - x0);
    for (unsigned i = 0; i < sizeof us_all / sizeof us_all[0]; i++)
        if (us_all[i]) rc |= tqb_device_upload_range(f->dev, f->h_in, (us_all[i] - f->h_in) + u0, u1 - u0);
    /* A, B and the vectors: only for the nodes of [ga, gb) outside the DMA runs */
    const double *xs[] = {in->b, in->q, in->xmin, in->xmax};
    const double *us[] = {in->r, in->umin, in->umax};
    int s0 = ga;
    for (int r = 0; r <= (f->in_direct ? f->n_in_runs : 0) && s0 < gb; r++)
    {
        /* host segment [s0, s1): up to the next run that starts at or after s0 (runs are sorted; single tree: g = k) */
        int s1 = gb, next = gb;
        if (f->in_direct && r < f->n_in_runs)
        {
            const int r0 = f->in_runs[r].k0 + T->tnode0[0], r1 = f->in_runs[r].k1 + T->tnode0[0];
            if (r1 <= s0) continue;
            if (r0 < gb) { s1 = r0 > s0 ? r0 : s0; next = r1; }
        }
        if (s1 > s0)
        {
            const int l1 = s1 >= d->Nn;
            const long long ha0 = T->aoff[s0], ha1 = l1 ? d->szA : T->aoff[s1], hb0 = T->boff[s0], hb1 = l1 ? d->szB : T->boff[s1];
            const long long hx0 = T->xoff[s0], hx1 = l1 ? d->X : T->xoff[s1], hu0 = T->uoff[s0], hu1 = l1 ? d->U : T->uoff[s1];
            rc |= tqb_device_upload_range(f->dev, f->h_in, (in->A - f->h_in) + ha0, ha1 - ha0);
            rc |= tqb_device_upload_range(f->dev, f->h_in, (in->B - f->h_in) + hb0, hb1 - hb0);
            for (unsigned i = 0; i < sizeof xs / sizeof xs[0]; i++) rc |= tqb_device_upload_range(f->dev, f->h_in, (xs[i] - f->h_in) + hx0, hx1 - hx0);
            for (unsigned i = 0; i < sizeof us / sizeof us[0]; i++) rc |= tqb_device_upload_range(f->dev, f->h_in, (us[i] - f->h_in) + hu0, hu1 - hu0);
        }
        s0 = next;
    }
    if (d->NC > 0)
    {
        const long long c0 = T->cgoff[ga], c1 = last ? d->szCg : T->cgoff[gb];
        const long long e0 = T->doff[ga], e1 = last ? d->NC : T->doff[gb];
        rc |= tqb_device_upload_range(f->dev, f->h_in, (in->Cg - f->h_in) + c0, c1 - c0);
        rc |= tqb_device_upload_range(f->dev, f->h_in, (in->dmin - f->h_in) + e0, e1 - e0);
        rc |= tqb_device_upload_range(f->dev, f->h_in, (in->dmax - f->h_in) + e0, e1 - e0);
    }
    return rc;
}

int tqb_forest_update_nodes(tqb_forest *f, int t, const tree_qp_in *qp, int j0, int j1)
{
    const int g0 = f->ht.tnode0[t];
    f->in_direct = 0; /* a partial update re-sends everything of its nodes through the staging buffer */
    pack_range(f, t, qp, NULL, j0, j1);
    int rc = tqb_device_upload_begin(f->dev);
    rc |= upload_nodes(f, g0 + j0, g0 + j1, 0);
    rc |= tqb_device_upload_end(f->dev, &f->h2d_ms);
    return rc;
}

/* Batch of closed-loop MPC QPs: only x0 changed since the last upload.  Re-pack the root and its children of every
 * tree (b of the root's children, r[0] and the root's bounds / dmin / dmax depend on x0, tree_qp_common.c:2154-2227)
 * and send those slices: trees of one shape sit at a constant stride in every array, so each array travels as ONE
 * strided copy for the whole batch. */
int tqb_forest_update_x0_batch(tqb_forest *f, const tree_qp_in *const *qp, long long *bytes)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    const tqb_inputs *in = &f->hin;
    const int nT = d->nT;
    const int n0 = 1 + qp[0]->tree[0].nkids, N0 = T->tnode0[1] - T->tnode0[0];
    int uniform = d->NC == 0;
    for (int t = 0; t < nT && uniform; t++)
        uniform = qp[t]->N == N0 && 1 + qp[t]->tree[0].nkids == n0;
    /* strides of one tree in every space; equal shapes give equal strides */
    const long long sX = nT > 1 ? T->xoff[T->tnode0[1]] - T->xoff[0] : d->X, sU = nT > 1 ? T->uoff[T->tnode0[1]] - T->uoff[0] : d->U;
    const long long sA = nT > 1 ? T->aoff[T->tnode0[1]] - T->aoff[0] : d->szA, sB = nT > 1 ? T->boff[T->tnode0[1]] - T->boff[0] : d->szB;
    for (int t = 1; t < nT && uniform; t++)
    {
        const int g = T->tnode0[t];
        uniform = T->xoff[g] == t * sX && T->uoff[g] == t * sU && T->aoff[g] == t * sA && T->boff[g] == t * sB;
        for (int j = 0; j <= n0 && uniform && j < N0; j++)
            uniform = T->xoff[g + j] - T->xoff[g] == T->xoff[j] && T->uoff[g + j] - T->uoff[g] == T->uoff[j];
    }
    f->in_direct = 0;
#pragma omp parallel for schedule(static) if (nT > 32)
    for (int t = 0; t < nT; t++) pack_range(f, t, qp[t], NULL, 0, 1 + qp[t]->tree[0].nkids);
    int rc = tqb_device_upload_begin(f->dev);
    long long sent = 0;
    if (!uniform)
    {
        for (int t = 0; t < nT; t++)
        {
            const int g0 = T->tnode0[t];
            rc |= upload_nodes(f, g0, g0 + 1 + qp[t]->tree[0].nkids, 0);
        }
        sent = -1;
    }
    else
    {
        const int last = n0 >= N0;
        const long long wx = last ? sX : T->xoff[n0], wu = last ? sU : T->uoff[n0];
        const long long wa = last ? sA : T->aoff[n0], wb = last ? sB : T->boff[n0];
        const double *xs[] = {in->b, in->q, in->Qd, in->Qinv, in->xmin, in->xmax, in->dxmin, in->dxmax};
        const double *us[] = {in->r, in->Rd, in->Rinv, in->umin, in->umax, in->dumin, in->dumax};
        rc |= tqb_device_upload_2d(f->dev, f->h_in, in->A - f->h_in, sA, wa, nT);
        rc |= tqb_device_upload_2d(f->dev, f->h_in, in->B - f->h_in, sB, wb, nT);
        sent += (wa + wb) * nT;
        for (unsigned i = 0; i < sizeof xs / sizeof xs[0]; i++)
            if (xs[i]) { rc |= tqb_device_upload_2d(f->dev, f->h_in, xs[i] - f->h_in, sX, wx, nT); sent += wx * nT; }
        for (unsigned i = 0; i < sizeof us / sizeof us[0]; i++)
            if (us[i]) { rc |= tqb_device_upload_2d(f->dev, f->h_in, us[i] - f->h_in, sU, wu, nT); sent += wu * nT; }
        if (d->has_dense)
        {   /* dense Hessian blocks do not depend on x0 */
        }
    }
    rc |= tqb_device_upload_end(f->dev, &f->h2d_ms);
    if (bytes) *bytes = sent < 0 ? -1 : (long long) sizeof(double) * sent;
    return rc;
}

int tqb_forest_upload_lambda0(tqb_forest *f)
{
    int rc = tqb_device_upload_begin(f->dev);
    rc |= tqb_device_upload_range(f->dev, f->h_in, f->hin.lam0 - f->h_in, f->dims.X);
    rc |= tqb_device_upload_end(f->dev, &f->h2d_ms);
    return rc;
}

int tqb_forest_warm_start(tqb_forest *f)
{
    return tqb_device_warm_start(f->dev, f->hin.lam0 - f->h_in, f->dims.X);
}

#ifndef TQB_UPLOAD_CHUNK_DOUBLES
#define TQB_UPLOAD_CHUNK_DOUBLES (3LL << 20) /* 24 MB of input per pipeline stage */
#endif

static int pack_upload_impl(tqb_forest *f, const tree_qp_in *const *qp, const int *gmap, double *tail_ms);

int tqb_forest_pack_upload(tqb_forest *f, const tree_qp_in *const *qp, double *tail_ms)
{
    return pack_upload_impl(f, qp, NULL, tail_ms);
}

/* sharded solve: the local tree's nodes are nodes gmap[j] of the caller's FULL tree_qp_in */
int tqb_forest_pack_upload_mapped(tqb_forest *f, const tree_qp_in *full_qp, const int *gmap, double *tail_ms)
{
    const tree_qp_in *one[1] = {full_qp};
    return pack_upload_impl(f, one, gmap, tail_ms);
}

/* Large single trees: the views of the caller's tree_qp_in sit at constant strides over long runs of nodes (all
 * parents of one shape, all leaves), so once the caller's buffer is page-locked (tqb_host_pin) A, B and the vectors
 * travel as strided DMA copies in the caller's layout and are converted on the device: the host packer only forms
 * the diagonal weights and their reciprocals, and the pinned staging block is by-passed for 90 % of the bytes. */
#ifndef TQB_DIRECT_MIN_NODES
#define TQB_DIRECT_MIN_NODES 4096 /* smallest (local) tree that tries the DMA path */
#endif
#ifndef TQB_DIRECT_MIN_RUN
#define TQB_DIRECT_MIN_RUN 1024   /* shortest run of nodes worth a strided copy of its own */
#endif
static int in_layout_equal(const tree_qp_in *qp, int k, int r)
{   /* nodes k and r (both >= 1): same dimensions and same offsets inside their matrix / vector chunks */
    const struct blasfeo_dmat *A = qp->A, *B = qp->B;
    if (qp->nx[k] != qp->nx[r] || qp->nu[k] != qp->nu[r] || qp->nc[k] != 0) return 0;
    if (A[k - 1].m != A[r - 1].m || A[k - 1].n != A[r - 1].n || A[k - 1].cn != A[r - 1].cn || A[k - 1].pm != A[r - 1].pm) return 0;
    if (B[k - 1].m != B[r - 1].m || B[k - 1].n != B[r - 1].n || B[k - 1].cn != B[r - 1].cn || B[k - 1].pm != B[r - 1].pm) return 0;
    if ((const char *) B[k - 1].pA - (const char *) A[k - 1].pA != (const char *) B[r - 1].pA - (const char *) A[r - 1].pA) return 0;
    const double *vb = qp->b[k - 1].pa, *vr = qp->b[r - 1].pa;
    return qp->q[k].pa - vb == qp->q[r].pa - vr && qp->r[k].pa - vb == qp->r[r].pa - vr &&
           qp->xmin[k].pa - vb == qp->xmin[r].pa - vr && qp->xmax[k].pa - vb == qp->xmax[r].pa - vr &&
           qp->umin[k].pa - vb == qp->umin[r].pa - vr && qp->umax[k].pa - vb == qp->umax[r].pa - vr;
}

static void build_in_runs(tqb_forest *f, const tree_qp_in *qp, const int *gmap)
{   /* runs are kept in LOCAL node indices j (the forest's own numbering); the views are those of node k = gmap[j] of the
     * caller's tree (gmap == NULL: k = j).  A sharded rank's local tree maps level by level onto consecutive nodes
     * of the full tree, so every deep level is one run. */
    const void *key[4] = {qp->A, qp->N > 1 ? (const void *) qp->A[0].pA : NULL, qp->b, gmap};
    if (!memcmp(f->in_key, key, sizeof key)) return; /* same views as last time */
    memcpy(f->in_key, key, sizeof key);
    f->n_in_runs = 0;
    const int Nl = f->ht.tnode0[1] - f->ht.tnode0[0];
#define GK(j) (gmap ? gmap[j] : (j))
    for (int j = 1; j + 1 < Nl && f->n_in_runs < TQB_MAX_RUNS;)
    {
        const int k = GK(j);
        if (k < 1 || GK(j + 1) != k + 1) { j++; continue; }
        const long long mp = (const char *) qp->A[k].pA - (const char *) qp->A[k - 1].pA;
        const long long vp = (const char *) qp->b[k].pa - (const char *) qp->b[k - 1].pa;
        int e = j + 1;
        while (e < Nl && GK(e) == k + (e - j) && in_layout_equal(qp, GK(e), k) &&
               (const char *) qp->A[GK(e) - 1].pA - (const char *) qp->A[GK(e) - 2].pA == mp &&
               (const char *) qp->b[GK(e) - 1].pa - (const char *) qp->b[GK(e) - 2].pa == vp)
            e++;
        if (e - j >= TQB_DIRECT_MIN_RUN)
        {
            tqb_in_run *r = &f->in_runs[f->n_in_runs];
            const struct blasfeo_dmat *A = &qp->A[k - 1], *B = &qp->B[k - 1];
            r->k0 = j; r->k1 = e;
            r->mbase = (const char *) A->pA; r->mpitch = mp;
            r->b_off = (const char *) B->pA - (const char *) A->pA;
            r->mwidth = r->b_off + (long long) sizeof(double) * B->pm * B->cn;
            r->ma = A->m; r->na = A->n; r->cna = A->cn; r->mb = B->m; r->nb = B->n; r->cnb = B->cn;
            r->nx = qp->nx[k]; r->nu = qp->nu[k];
            const double *vb = qp->b[k - 1].pa;
            const double *vs[7] = {vb, qp->q[k].pa, qp->r[k].pa, qp->xmin[k].pa, qp->xmax[k].pa, qp->umin[k].pa, qp->umax[k].pa};
            const int len[7] = {r->nx, r->nx, r->nu, r->nx, r->nx, r->nu, r->nu};
            long long lo = 0, hi = 0;
            for (int i = 0; i < 7; i++)
            {
                const long long o = vs[i] - vb;
                r->voff[i] = (int) o;
                if (o < lo) lo = o;
                if (o + len[i] > hi) hi = o + len[i];
            }
            r->vbase = (const char *) vb; r->vpitch = vp; r->vwidth = (long long) sizeof(double) * hi;
            const int ok = lo == 0 && mp > 0 && vp > 0 && r->b_off >= (long long) sizeof(double) * A->pm * A->cn && r->mwidth <= mp &&
                           r->vwidth <= vp && !(mp & 7) && !(vp & 7) && !(r->b_off & 7);
            if (ok) f->n_in_runs++;
        }
        j = e;
    }
#undef GK
}

/* returns 1 when the runs were sent by DMA (f->in_direct), 0 when the whole tree takes the staged path */
static int upload_in_runs(tqb_forest *f, const tree_qp_in *qp, const int *gmap)
{
    f->in_direct = 0;
    if (f->dims.Nn < TQB_DIRECT_MIN_NODES || f->dims.has_dense || f->dims.NC > 0) return 0;
    build_in_runs(f, qp, gmap);
    if (f->n_in_runs == 0) return 0;
    /* one page-locked range from the first matrix to the last vector of the runs */
    const char *lo = NULL, *hi = NULL;
    size_t stage = 0;
    for (int i = 0; i < f->n_in_runs; i++)
    {
        const tqb_in_run *r = &f->in_runs[i];
        const long long rows = r->k1 - r->k0;
        const char *p[4] = {r->mbase, r->mbase + r->mpitch * (rows - 1) + r->mwidth, r->vbase, r->vbase + r->vpitch * (rows - 1) + r->vwidth};
        for (int j = 0; j < 4; j++)
        {
            if (!lo || p[j] < lo) lo = p[j];
            if (!hi || p[j] > hi) hi = p[j];
        }
        stage += (size_t) rows * (size_t) (r->mwidth + r->vwidth);
    }
    if (!tqb_host_pin(lo, (size_t) (hi - lo))) return 0;
    if (tqb_device_stage_reserve(f->dev, stage)) return 0;
    size_t off = 0;
    for (int i = 0; i < f->n_in_runs; i++)
    {
        const tqb_in_run *r = &f->in_runs[i];
        if (tqb_device_upload_run(f->dev, r, f->ht.tnode0[0] + r->k0, off)) return 0;
        off += (size_t) (r->k1 - r->k0) * (size_t) (r->mwidth + r->vwidth);
    }
    f->in_direct = 1;
    return 1;
}

static int pack_upload_impl(tqb_forest *f, const tree_qp_in *const *qp, const int *gmap, double *tail_ms)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    const int nT = d->nT;
    long long nstage = d->n_in / TQB_UPLOAD_CHUNK_DOUBLES;
    if (nstage < 1) nstage = 1;
    if (nstage > 32) nstage = 32;
    int rc = tqb_device_upload_begin(f->dev);
    f->in_direct = 0;
    const double tv0 = tqb_wall_ms();
    if (nT == 1) upload_in_runs(f, qp[0], gmap);
    const double tv1 = tqb_wall_ms();
    if (nT == 1)
    {
        const int N = T->tnode0[1];
        for (int c = 0; c < nstage; c++)
        {
            const int j0 = (int) ((long long) N * c / nstage), j1 = (int) ((long long) N * (c + 1) / nstage);
            pack_range(f, 0, qp[0], gmap, j0, j1);
            rc |= upload_nodes(f, j0, j1, 1);
        }
    }
    else
    {
        for (int c = 0; c < nstage; c++)
        {
            const int t0 = (int) ((long long) nT * c / nstage), t1 = (int) ((long long) nT * (c + 1) / nstage);
            if (t1 == t0) continue;
#pragma omp parallel for schedule(dynamic, 8) if (t1 - t0 > 32)
            for (int t = t0; t < t1; t++) tqb_forest_pack(f, t, qp[t]);
            rc |= upload_nodes(f, T->tnode0[t0], T->tnode0[t1], 1);
        }
    }
    /* dense Hessian blocks: offsets exist only for the dense nodes, sent as one range at the end */
    if (d->has_dense) rc |= tqb_device_upload_range(f->dev, f->h_in, f->hin.H - f->h_in, d->szH);
    if (f->in_direct) rc |= tqb_device_finish_direct_upload(f->dev, f->hin.lam0 - f->h_in, f->lam0_zero);
    const double t0 = tqb_wall_ms();
    double total = 0.0;
    rc |= tqb_device_upload_end(f->dev, &total);
    if (tail_ms) *tail_ms = tqb_wall_ms() - t0;
    f->h2d_ms = total;
    if (tqb_verbose())
        fprintf(stderr, "[treeqp_b200] upload: runs %d (direct %d) enqueue %.3f ms, pack loop %.3f ms, tail %.3f ms, stream %.3f ms\n",
                f->n_in_runs, f->in_direct, tv1 - tv0, t0 - tv1, tqb_wall_ms() - t0, total);
    return rc;
}

/* lambda: multipliers of nodes 1..N-1 of tree t, concatenated (dual_Newton_tree.c:1654-1663) */
void tqb_forest_set_lambda0(tqb_forest *f, int t, const double *lambda)
{
    const tqb_tables *T = &f->ht;
    int g0 = T->tnode0[t], g1 = T->tnode0[t + 1];
    int xroot = T->xoff[g0], nx0 = T->nx[g0];
    int xend = (g1 < f->dims.Nn) ? T->xoff[g1] : f->dims.X;
    for (int i = 0; i < nx0; i++) f->hin.lam0[xroot + i] = 0.0;
    const long long n = xend - xroot - nx0;
    double *dst = f->hin.lam0 + xroot + nx0;
    if (n < (1 << 16))
    {
        memcpy(dst, lambda, sizeof(double) * (size_t) n);
        f->lam0_zero = 0;
    }
    else
    {   /* large trees: copy in parallel and notice a start of zeros (the cold start of every benchmark and example),
         * which the DMA upload path turns into a device memset */
        unsigned long long nz = 0;
#pragma omp parallel for schedule(static) reduction(| : nz)
        for (long long c = 0; c < (n + 4095) / 4096; c++)
        {
            const long long a = c * 4096, b = a + 4096 < n ? a + 4096 : n;
            unsigned long long acc = 0;
            for (long long i = a; i < b; i++)
            {
                unsigned long long bits;
                memcpy(&bits, &lambda[i], sizeof bits);
                acc |= bits;
                dst[i] = lambda[i];
            }
            nz |= acc;
        }
        f->lam0_zero = (nz == 0) && f->dims.nT == 1; /* +0.0 everywhere */
    }
}

/* device results (pinned staging) -> BLASFEO views of tree_qp_out
 * (copy-out loop of dual_Newton_tree.c:1235-1247) */
void tqb_forest_unpack(const tqb_forest *f, int t, tree_qp_out *out)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    const double *x = f->h_out, *u = x + d->X, *lam = u + d->U, *mux = lam + d->X, *muu = mux + d->X, *mud = muu + d->U;
    int g0 = T->tnode0[t], N = T->tnode0[t + 1] - g0;
#pragma omp parallel for schedule(static) if (N > 2048)
    for (int k = 0; k < N; k++)
    {
        int g = g0 + k, nx = T->nx[g], nu = T->nu[g], xo = T->xoff[g], uo = T->uoff[g];
        memcpy(out->x[k].pa, x + xo, sizeof(double) * (size_t) nx);
        memcpy(out->u[k].pa, u + uo, sizeof(double) * (size_t) nu);
        memcpy(out->mu_x[k].pa, mux + xo, sizeof(double) * (size_t) nx);
        memcpy(out->mu_u[k].pa, muu + uo, sizeof(double) * (size_t) nu);
        if (k > 0) memcpy(out->lam[k - 1].pa, lam + xo, sizeof(double) * (size_t) nx);
        if (T->ncn[g] > 0) memcpy(out->mu_d[k].pa, mud + T->doff[g], sizeof(double) * (size_t) T->ncn[g]);
    }
}

/* ---- results straight into the caller's tree_qp_out (large single trees) -------------------------------------
 * tree_qp_out_create carves [x u mu_x mu_u mu_d lam] node by node: runs of nodes of one shape sit at a constant
 * stride.  With the caller's buffer page-locked the device assembles each run in the caller's layout (k_gather_run)
 * and ONE strided D2H copy per run writes it in place; only the nodes outside the runs (the root, its children) go
 * through the pinned block and the host loop. */
static void build_out_runs(tqb_forest *f, const tree_qp_out *out, const int *gmap, const int *owned)
{   /* local indices j, views of node k = gmap[j]; nodes a sharded rank does not own are never written */
    const void *key[4] = {out->x, out->x[0].pa, out->lam, gmap};
    if (!memcmp(f->out_key, key, sizeof key)) return;
    memcpy(f->out_key, key, sizeof key);
    f->n_out_runs = 0;
    const tqb_tables *T = &f->ht;
    const int g0 = T->tnode0[0], Nl = T->tnode0[1] - g0;
#define GK(j) (gmap ? gmap[j] : (j))
#define OFFS(k, o) do { const double *b_ = out->x[k].pa; (o)[0] = 0; (o)[1] = out->u[k].pa - b_; (o)[2] = out->mu_x[k].pa - b_; \
                        (o)[3] = out->mu_u[k].pa - b_; (o)[4] = out->lam[(k) - 1].pa - b_; } while (0)
    for (int j = 1; j + 1 < Nl && f->n_out_runs < TQB_MAX_RUNS;)
    {
        const int k = GK(j);
        if (k < 1 || (owned && !owned[j]) || GK(j + 1) != k + 1) { j++; continue; }
        long long o0[5];
        OFFS(k, o0);
        const long long pitch = (const char *) out->x[k + 1].pa - (const char *) out->x[k].pa;
        int e = j + 1;
        for (; e < Nl; e++)
        {
            const int ke = GK(e);
            long long o[5];
            if (ke != k + (e - j) || (owned && !owned[e])) break;
            OFFS(ke, o);
            if (T->nx[g0 + e] != T->nx[g0 + j] || T->nu[g0 + e] != T->nu[g0 + j] || T->ncn[g0 + e] != 0 || memcmp(o, o0, sizeof o) ||
                (const char *) out->x[ke].pa - (const char *) out->x[ke - 1].pa != pitch)
                break;
        }
        if (e - j >= TQB_DIRECT_MIN_RUN && T->ncn[g0 + j] == 0)
        {
            tqb_out_run *r = &f->out_runs[f->n_out_runs];
            r->k0 = j; r->k1 = e; r->base = (char *) out->x[k].pa; r->pitch = pitch;
            r->nx = T->nx[g0 + j]; r->nu = T->nu[g0 + j];
            const int len[5] = {r->nx, r->nu, r->nx, r->nu, r->nx};
            long long lo = 0, hi = 0;
            for (int i = 0; i < 5; i++)
            {
                r->off[i] = (int) o0[i];
                if (o0[i] < lo) lo = o0[i];
                if (o0[i] + len[i] > hi) hi = o0[i] + len[i];
            }
            r->width = (long long) sizeof(double) * hi;
            if (lo == 0 && pitch > 0 && r->width <= pitch && !(pitch & 7)) f->n_out_runs++;
        }
        j = e;
    }
#undef OFFS
#undef GK
}

/* returns 1 when qp_out has been filled (runs by DMA, the other nodes through the pinned block), 0: not applicable.
 * gmap / owned: the local tree of a sharded rank (NULL: the tree itself, every node) */
int tqb_forest_download_direct(tqb_forest *f, tree_qp_out *out, const int *gmap, const int *owned)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    if (d->nT != 1 || d->Nn < TQB_DIRECT_MIN_NODES || d->NC > 0) return 0;
    const int N = d->Nn, g0 = T->tnode0[0];
    build_out_runs(f, out, gmap, owned);
    if (f->n_out_runs == 0) return 0;
    char *lo = NULL, *hi = NULL;
    size_t stage = 0;
    for (int i = 0; i < f->n_out_runs; i++)
    {
        const tqb_out_run *r = &f->out_runs[i];
        const long long rows = r->k1 - r->k0;
        char *e = r->base + r->pitch * (rows - 1) + r->width;
        if (!lo || r->base < lo) lo = r->base;
        if (!hi || e > hi) hi = e;
        stage += (size_t) rows * (size_t) r->width;
    }
    if (!tqb_host_pin(lo, (size_t) (hi - lo))) return 0;
    if (tqb_device_stage_reserve(f->dev, stage)) return 0;
    int rc = tqb_device_download_begin(f->dev);
    size_t off = 0;
    long long bytes = 0;
    for (int i = 0; i < f->n_out_runs; i++)
    {
        const tqb_out_run *r = &f->out_runs[i];
        rc |= tqb_device_download_run(f->dev, r, g0 + r->k0, off);
        off += (size_t) (r->k1 - r->k0) * (size_t) r->width;
        bytes += (long long) (r->k1 - r->k0) * r->width;
    }
    /* nodes outside the runs: their slices of x, u, lam, mu_x, mu_u into the pinned block */
    int s0 = 0;
    for (int i = 0; i <= f->n_out_runs; i++)
    {
        const int s1 = i < f->n_out_runs ? f->out_runs[i].k0 : N;
        if (s1 > s0)
        {
            const int l1 = g0 + s1 >= d->Nn;
            const long long x0 = T->xoff[g0 + s0], x1 = l1 ? d->X : T->xoff[g0 + s1], u0 = T->uoff[g0 + s0], u1 = l1 ? d->U : T->uoff[g0 + s1];
            rc |= tqb_device_download_nodes(f->dev, f->h_out, x0, x1, u0, u1);
            bytes += (long long) sizeof(double) * (3 * (x1 - x0) + 2 * (u1 - u0));
        }
        if (i < f->n_out_runs) s0 = f->out_runs[i].k1;
    }
    const double td0 = tqb_wall_ms();
    rc |= tqb_device_download_end(f->dev, &f->d2h_ms);
    if (tqb_verbose()) fprintf(stderr, "[treeqp_b200] download: %d runs, wait %.3f ms, stream %.3f ms\n", f->n_out_runs, tqb_wall_ms() - td0, f->d2h_ms);
    if (rc) return 0;
    f->d2h_bytes = bytes;
    const double *x = f->h_out, *u = x + d->X, *lam = u + d->U, *mux = lam + d->X, *muu = mux + d->X;
    s0 = 0;
    for (int i = 0; i <= f->n_out_runs; i++)
    {
        const int s1 = i < f->n_out_runs ? f->out_runs[i].k0 : N;
        for (int j = s0; j < s1; j++)
        {
            if (owned && !owned[j]) continue;
            const int k = gmap ? gmap[j] : j;
            const int g = g0 + j, nx = T->nx[g], nu = T->nu[g], xo = T->xoff[g], uo = T->uoff[g];
            memcpy(out->x[k].pa, x + xo, sizeof(double) * (size_t) nx);
            memcpy(out->u[k].pa, u + uo, sizeof(double) * (size_t) nu);
            memcpy(out->mu_x[k].pa, mux + xo, sizeof(double) * (size_t) nx);
            memcpy(out->mu_u[k].pa, muu + uo, sizeof(double) * (size_t) nu);
            if (k > 0) memcpy(out->lam[k - 1].pa, lam + xo, sizeof(double) * (size_t) nx);
        }
        if (i < f->n_out_runs) s0 = f->out_runs[i].k1;
    }
    return 1;
}
