/*
 * containers.c -- host-side QP containers for the B200 dual-Newton engine.
 *
 * The hot path (tdunes.c + the CUDA kernels) only READS tree_qp_in and WRITES tree_qp_out, both
 * laid out exactly like the reference's (BLASFEO panel-major views carved out of one caller
 * buffer).  A drop-in user keeps the reference's own tree_qp_common.o for building them; this
 * file is the small stand-alone equivalent so that libtreeqp_b200.so is usable (and benchable)
 * without the reference: same names, same argument meaning, same memory protocol
 * (calculate_size -> caller malloc -> create carves) as treeqp/src/tree_qp_common.c:60-400,
 * treeqp/utils/tree.c:36-280 and treeqp/utils/memory.c:50-229.
 *
 * Written from the interface description in SURVEY.md; no BLASFEO code is linked: the
 * panel-major accessors are three-line helpers below.
 */
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "treeqp_b200.h"
#include "host_internal.h"

/* ------------------------------------------------------------------ panel-major helpers -- */

static int round_up(int v, int to) { return (v + to - 1) / to * to; }

/* bytes of a panel-major m x n matrix incl. its inverse-diagonal scratch
 * (blasfeo_memsize_dmat, external/blasfeo/auxiliary/d_aux_lib4.c:1567) */
int tqb_dmat_bytes(int m, int n)
{
    int pm = round_up(m, TQB_PS), cn = round_up(n, 4);
    int dmin = m < n ? m : n;
    return (pm * cn + round_up(dmin, 16)) * (int) sizeof(double);
}

int tqb_dvec_bytes(int m) { return round_up(m, TQB_PS) * (int) sizeof(double); }

static void carve_dmat(int m, int n, struct blasfeo_dmat *M, char **cur)
{
    M->m = m;
    M->n = n;
    M->pm = round_up(m, TQB_PS);
    M->cn = round_up(n, 4);
    M->pA = (double *) *cur;
    M->dA = M->pA + M->pm * M->cn;
    M->use_dA = 0;
    M->memsize = tqb_dmat_bytes(m, n);
    memset(*cur, 0, (size_t) M->memsize);
    *cur += M->memsize;
}

static void carve_dvec(int m, struct blasfeo_dvec *v, char **cur)
{
    v->m = m;
    v->pm = round_up(m, TQB_PS);
    v->pa = (double *) *cur;
    v->memsize = tqb_dvec_bytes(m);
    memset(*cur, 0, (size_t) v->memsize);
    *cur += v->memsize;
}

static void align64(char **cur)
{
    uintptr_t p = (uintptr_t) *cur;
    p = (p + 63u) / 64u * 64u;
    *cur = (char *) p;
}

void tqb_dmat_from_colmajor(int m, int n, const double *src, int lds, struct blasfeo_dmat *M)
{
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) BLASFEO_DMATEL(M, i, j) = src[i + (size_t) j * lds];
    M->use_dA = 0;
}

void tqb_dmat_to_colmajor(int m, int n, const struct blasfeo_dmat *M, double *dst, int ldd)
{
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) dst[i + (size_t) j * ldd] = BLASFEO_DMATEL(M, i, j);
}

static void dmat_zero(struct blasfeo_dmat *M)
{
    memset(M->pA, 0, sizeof(double) * (size_t) M->pm * M->cn);
    M->use_dA = 0;
}

static void dmat_copy(const struct blasfeo_dmat *S, struct blasfeo_dmat *D)
{
    for (int j = 0; j < S->n; j++)
        for (int i = 0; i < S->m; i++) BLASFEO_DMATEL(D, i, j) = BLASFEO_DMATEL(S, i, j);
    D->use_dA = 0;
}

/* ------------------------------------------------------------------------------ trees ---- */

static int int_pow(int base, int e)
{
    int r = 1;
    while (e-- > 0) r *= base;
    return r;
}

/* utils/tree.c:36: nodes of a tree that branches md-fold for Nr stages and then runs Nh-Nr
 * more stages without branching */
int calculate_number_of_nodes(int md, int Nr, int Nh)
{
    if (md == 1) return Nh + 1;
    return (Nh - Nr) * int_pow(md, Nr) + (int_pow(md, Nr + 1) - 1) / (md - 1);
}

/* utils/tree.c:247: children counts in breadth-first order for the tree above */
void setup_multistage_tree(int md, int Nr, int Nh, int *nk)
{
    int at = 0, width = 1;
    for (int stage = 0; stage < Nh; stage++)
    {
        int fan = stage < Nr ? md : 1;
        for (int i = 0; i < width; i++) nk[at + i] = fan;
        at += width;
        width *= fan;
    }
    for (int i = 0; i < width; i++) nk[at + i] = 0;
}

/* utils/tree.c:105: walks nk[] level by level until a level has no children; -1 if the
 * counts are inconsistent */
int number_of_nodes_from_nkids(const int *nk)
{
    int at = 0, width = 1;
    for (;;)
    {
        int next = 0;
        for (int i = 0; i < width; i++)
        {
            if (nk[at + i] < 0) return -1;
            if (nk[at + i] == 0) break; /* first leaf ends the scan of this level */
            next += nk[at + i];
        }
        at += width;
        if (next == 0) return at;
        if (next < width) return -1;
        width = next;
    }
}

/* Breadth-first tree from children counts (utils/tree.c:171).  Linear time: the children of
 * node i start right after the children of node i-1.  `real` follows the reference: siblings
 * of a branching node are numbered 0..md-1, single children inherit. */
static void build_tree(int Nn, const int *nk, struct node *tree, int *kids_store)
{
    int next_kid = 1;
    tree[0].idx = 0;
    tree[0].dad = -1;
    tree[0].stage = 0;
    tree[0].idxkid = 0;
    tree[0].real = -1;
    for (int i = 0; i < Nn; i++)
    {
        tree[i].nkids = nk[i];
        tree[i].kids = nk[i] > 0 ? kids_store : NULL;
        for (int j = 0; j < nk[i]; j++)
        {
            int k = next_kid + j;
            assert(k < Nn);
            kids_store[j] = k;
            tree[k].idx = k;
            tree[k].dad = i;
            tree[k].stage = tree[i].stage + 1;
            tree[k].idxkid = j;
            if (nk[i] > 1)
                tree[k].real = j;
            else
                tree[k].real = i > 0 ? tree[i].real : 0;
        }
        kids_store += nk[i];
        next_kid += nk[i];
    }
}

static int dad_of(int idx, const int *nk)
{
    if (idx == 0) return -1;
    int acc = 0;
    for (int i = 0; i < idx; i++)
    {
        acc += nk[i];
        if (acc >= idx) return i;
    }
    return -1;
}

/* ------------------------------------------------------------------------- tree_qp_in ---- */

int tree_qp_in_calculate_size(int Nn, const int *nx, const int *nu, const int *nc, const int *nk)
{
    long long bytes = 0;
    int kids_total = 0;
    for (int i = 0; i < Nn; i++) kids_total += nk[i];

    bytes += (long long) Nn * sizeof(struct node) + (long long) kids_total * sizeof(int);
    bytes += 3LL * Nn * sizeof(int);
    bytes += (long long) (Nn - 1) * (2 * sizeof(struct blasfeo_dmat) + sizeof(struct blasfeo_dvec));
    bytes += (long long) Nn * (3 * sizeof(struct blasfeo_dmat) + 2 * sizeof(struct blasfeo_dvec));
    bytes += (long long) Nn * 4 * sizeof(struct blasfeo_dvec);
    bytes += (long long) Nn * (2 * sizeof(struct blasfeo_dmat) + 2 * sizeof(struct blasfeo_dvec));
    bytes += (long long) nk[0] * (2 * sizeof(int) + sizeof(struct blasfeo_dmat) + sizeof(struct blasfeo_dvec));

    /* parents in O(Nn): children are contiguous */
    int dad = 0, left = nk[0];
    for (int i = 0; i < Nn; i++)
    {
        int ncs = nc ? nc[i] : 0;
        if (i > 0)
        {
            while (left == 0) left = nk[++dad];
            left--;
            bytes += tqb_dmat_bytes(nx[i], nx[dad]) + tqb_dmat_bytes(nx[i], nu[dad]) + tqb_dvec_bytes(nx[i]);
            if (i <= nk[0]) bytes += tqb_dmat_bytes(nx[i], nx[dad]) + tqb_dvec_bytes(nx[i]);
        }
        bytes += tqb_dmat_bytes(nx[i], nx[i]) + tqb_dmat_bytes(nu[i], nu[i]) + tqb_dmat_bytes(nu[i], nx[i]);
        bytes += 3 * tqb_dvec_bytes(nx[i]) + 3 * tqb_dvec_bytes(nu[i]);
        bytes += tqb_dmat_bytes(ncs, nx[i]) + tqb_dmat_bytes(ncs, nu[i]) + 2 * tqb_dvec_bytes(ncs);
        if (i == 0)
        {
            bytes += tqb_dvec_bytes(nx[0]) + tqb_dmat_bytes(ncs, nx[0]) + 2 * tqb_dvec_bytes(ncs);
            bytes += tqb_dmat_bytes(nu[0], nx[0]) + tqb_dvec_bytes(nu[0]);
        }
    }
    bytes = (bytes + 63) / 64 * 64 + 64;
    if (bytes > 2147483647LL)
    {
        fprintf(stderr, "[treeqp_b200] tree_qp_in needs %lld bytes: exceeds the int return type "
                        "of the reference API\n", bytes);
        exit(1);
    }
    (void) dad_of;
    return (int) bytes;
}

void tree_qp_in_create(int Nn, const int *nx, const int *nu, const int *nc, const int *nk,
                       tree_qp_in *qp, void *ptr)
{
    char *cur = (char *) ptr;
    int kids_total = 0;
    for (int i = 0; i < Nn; i++) kids_total += nk[i];
    assert(number_of_nodes_from_nkids(nk) == Nn);

    qp->N = Nn;
    qp->tree = (struct node *) cur;
    cur += (size_t) Nn * sizeof(struct node);
    build_tree(Nn, nk, qp->tree, (int *) cur);
    cur += (size_t) kids_total * sizeof(int);

    qp->nx = (int *) cur; cur += (size_t) Nn * sizeof(int);
    qp->nu = (int *) cur; cur += (size_t) Nn * sizeof(int);
    qp->nc = (int *) cur; cur += (size_t) Nn * sizeof(int);
    qp_internal_t *im = &qp->internal_memory;
    im->is_A_initialized = (int *) cur; cur += (size_t) nk[0] * sizeof(int);
    im->is_b_initialized = (int *) cur; cur += (size_t) nk[0] * sizeof(int);
    for (int i = 0; i < Nn; i++)
    {
        qp->nx[i] = nx[i];
        qp->nu[i] = nu[i];
        qp->nc[i] = nc ? nc[i] : 0;
    }
    for (int i = 0; i < nk[0]; i++) im->is_A_initialized[i] = im->is_b_initialized[i] = 0;
    im->is_C_initialized = im->is_dmin_initialized = im->is_dmax_initialized = 0;
    im->is_S_initialized = im->is_r_initialized = 0;

#define TAKE(field, type, count) \
    field = (type *) cur;        \
    cur += (size_t) (count) * sizeof(type)
    TAKE(qp->A, struct blasfeo_dmat, Nn - 1);
    TAKE(qp->B, struct blasfeo_dmat, Nn - 1);
    TAKE(qp->b, struct blasfeo_dvec, Nn - 1);
    TAKE(qp->Q, struct blasfeo_dmat, Nn);
    TAKE(qp->R, struct blasfeo_dmat, Nn);
    TAKE(qp->S, struct blasfeo_dmat, Nn);
    TAKE(qp->q, struct blasfeo_dvec, Nn);
    TAKE(qp->r, struct blasfeo_dvec, Nn);
    TAKE(qp->xmin, struct blasfeo_dvec, Nn);
    TAKE(qp->xmax, struct blasfeo_dvec, Nn);
    TAKE(qp->umin, struct blasfeo_dvec, Nn);
    TAKE(qp->umax, struct blasfeo_dvec, Nn);
    TAKE(qp->C, struct blasfeo_dmat, Nn);
    TAKE(qp->D, struct blasfeo_dmat, Nn);
    TAKE(qp->dmin, struct blasfeo_dvec, Nn);
    TAKE(qp->dmax, struct blasfeo_dvec, Nn);
    TAKE(im->A0, struct blasfeo_dmat, nk[0]);
    TAKE(im->b0, struct blasfeo_dvec, nk[0]);
#undef TAKE
    align64(&cur);

    for (int i = 0; i < Nn; i++)
    {
        int p = qp->tree[i].dad;
        if (i > 0)
        {
            carve_dmat(nx[i], nx[p], &qp->A[i - 1], &cur);
            carve_dmat(nx[i], nu[p], &qp->B[i - 1], &cur);
            if (i <= nk[0]) carve_dmat(nx[i], nx[p], &im->A0[i - 1], &cur);
        }
        carve_dmat(nx[i], nx[i], &qp->Q[i], &cur);
        carve_dmat(nu[i], nu[i], &qp->R[i], &cur);
        carve_dmat(nu[i], nx[i], &qp->S[i], &cur);
        carve_dmat(qp->nc[i], nx[i], &qp->C[i], &cur);
        carve_dmat(qp->nc[i], nu[i], &qp->D[i], &cur);
    }
    carve_dmat(qp->nc[0], nx[0], &im->C0, &cur);
    carve_dmat(nu[0], nx[0], &im->S0, &cur);

    for (int i = 0; i < Nn; i++)
    {
        if (i > 0)
        {
            carve_dvec(nx[i], &qp->b[i - 1], &cur);
            if (i <= nk[0]) carve_dvec(nx[i], &im->b0[i - 1], &cur);
        }
        carve_dvec(nx[i], &qp->q[i], &cur);
        carve_dvec(nu[i], &qp->r[i], &cur);
        carve_dvec(nx[i], &qp->xmin[i], &cur);
        carve_dvec(nx[i], &qp->xmax[i], &cur);
        carve_dvec(nu[i], &qp->umin[i], &cur);
        carve_dvec(nu[i], &qp->umax[i], &cur);
        carve_dvec(qp->nc[i], &qp->dmin[i], &cur);
        carve_dvec(qp->nc[i], &qp->dmax[i], &cur);
    }
    carve_dvec(nx[0], &im->x0, &cur);
    carve_dvec(qp->nc[0], &im->dmin0, &cur);
    carve_dvec(qp->nc[0], &im->dmax0, &cur);
    carve_dvec(nu[0], &im->r0, &cur);

    tree_qp_in_set_inf_bounds(qp);
    assert(cur <= (char *) ptr + tree_qp_in_calculate_size(Nn, nx, nu, nc, nk));
}

void tree_qp_in_set_inf_bounds(tree_qp_in *qp)
{
    for (int i = 0; i < qp->N; i++)
    {
        for (int j = 0; j < qp->nx[i]; j++)
        {
            qp->xmin[i].pa[j] = -TREEQP_INF;
            qp->xmax[i].pa[j] = TREEQP_INF;
        }
        for (int j = 0; j < qp->nu[i]; j++)
        {
            qp->umin[i].pa[j] = -TREEQP_INF;
            qp->umax[i].pa[j] = TREEQP_INF;
        }
    }
}

/* edge `indx` joins node indx+1 to its parent (tree_qp_common.c:874-1105).  A copy of the
 * root's outgoing dynamics is kept so that x0 can be eliminated later. */
void tree_qp_in_set_edge_dynamics_colmajor(const double *const A, const double *const B,
                                           const double *const b, tree_qp_in *const qp,
                                           const int indx)
{
    assert(indx >= 0 && indx < qp->N - 1);
    int k = indx + 1, p = qp->tree[k].dad;
    int nxk = qp->nx[k], nxp = qp->nx[p], nup = qp->nu[p];
    qp_internal_t *im = &qp->internal_memory;
    if (A)
    {
        tqb_dmat_from_colmajor(nxk, nxp, A, nxk, &qp->A[indx]);
        if (p == 0 && nxp > 0)
        {
            dmat_copy(&qp->A[indx], &im->A0[indx]);
            im->is_A_initialized[indx] = 1;
        }
    }
    if (B) tqb_dmat_from_colmajor(nxk, nup, B, nxk, &qp->B[indx]);
    if (b)
    {
        memcpy(qp->b[indx].pa, b, sizeof(double) * (size_t) nxk);
        if (p == 0 && nxp > 0)
        {
            memcpy(im->b0[indx].pa, b, sizeof(double) * (size_t) nxk);
            im->is_b_initialized[indx] = 1;
        }
    }
}

void tree_qp_in_set_node_objective_colmajor(const double *const Q, const double *const R,
                                            const double *const S, const double *const q,
                                            const double *const r, tree_qp_in *const qp,
                                            const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    int nx = qp->nx[indx], nu = qp->nu[indx];
    qp_internal_t *im = &qp->internal_memory;
    if (Q) tqb_dmat_from_colmajor(nx, nx, Q, nx, &qp->Q[indx]);
    if (R) tqb_dmat_from_colmajor(nu, nu, R, nu, &qp->R[indx]);
    if (S)
    {
        tqb_dmat_from_colmajor(nu, nx, S, nu, &qp->S[indx]);
        if (indx == 0 && nx > 0)
        {
            dmat_copy(&qp->S[0], &im->S0);
            im->is_S_initialized = 1;
        }
    }
    if (q) memcpy(qp->q[indx].pa, q, sizeof(double) * (size_t) nx);
    if (r)
    {
        memcpy(qp->r[indx].pa, r, sizeof(double) * (size_t) nu);
        if (indx == 0 && nx > 0)
        {
            memcpy(im->r0.pa, r, sizeof(double) * (size_t) nu);
            im->is_r_initialized = 1;
        }
    }
}

/* ---- initial state: pinned on the root or eliminated (tree_qp_common.c:404-536, 2154-2235) ---- */

/* y = y0 + sign * M x  (blasfeo_dgemv_n of the reference, plain loops) */
static void gemv_n_add(const struct blasfeo_dmat *M, int n, double sign, const double *x, const double *y0, double *y)
{
    for (int i = 0; i < M->m; i++)
    {
        double acc = y0[i];
        for (int j = 0; j < n; j++) acc += sign * BLASFEO_DMATEL(M, i, j) * x[j];
        y[i] = acc;
    }
}

/* nx[0] > 0: x0 is pinned through the root's bounds.  nx[0] == 0 (x0 eliminated): only b of the
 * root's children, r[0] and dmin/dmax[0] depend on x0 and are rebuilt from the kept copies
 * (tree_qp_common.c:2154-2227) */
void tree_qp_in_set_x0_strvec(tree_qp_in *qp, struct blasfeo_dvec *sx0)
{
    qp_internal_t *im = &qp->internal_memory;
    if (qp->nx[0] > 0)
    {
        assert(qp->xmin[0].m == sx0->m && qp->xmax[0].m == sx0->m);
        memcpy(qp->xmin[0].pa, sx0->pa, sizeof(double) * (size_t) sx0->m);
        memcpy(qp->xmax[0].pa, sx0->pa, sizeof(double) * (size_t) sx0->m);
        return;
    }
    const int nx0 = sx0->m, nc0 = im->C0.m, nu0 = im->S0.m;
    assert(im->x0.m == nx0 && qp->nu[0] == nu0 && qp->nc[0] == nc0);
    if (im->x0.pa != sx0->pa) memcpy(im->x0.pa, sx0->pa, sizeof(double) * (size_t) nx0);
    for (int ii = 1; ii <= qp->tree[0].nkids; ii++)
    {
        assert(im->is_A_initialized[ii - 1] == 1 && im->is_b_initialized[ii - 1] == 1);
        gemv_n_add(&im->A0[ii - 1], nx0, 1.0, sx0->pa, im->b0[ii - 1].pa, qp->b[ii - 1].pa);
    }
    if (nc0 > 0)
    {
        assert(im->is_C_initialized == 1);
        gemv_n_add(&im->C0, nx0, -1.0, sx0->pa, im->dmin0.pa, qp->dmin[0].pa);
        gemv_n_add(&im->C0, nx0, -1.0, sx0->pa, im->dmax0.pa, qp->dmax[0].pa);
    }
    assert(im->is_S_initialized == 1 && im->is_r_initialized == 1);
    gemv_n_add(&im->S0, nx0, 1.0, sx0->pa, im->r0.pa, qp->r[0].pa);
}

void tree_qp_in_set_x0_colmaj(tree_qp_in *qp, double *x0)
{
    struct blasfeo_dvec *sx0 = &qp->internal_memory.x0;
    memcpy(sx0->pa, x0, sizeof(double) * (size_t) sx0->m);
    tree_qp_in_set_x0_strvec(qp, sx0);
}

/* removes the root's state from the QP: the pinned value xmin[0] (= xmax[0]) goes into b of the
 * root's children, r[0] and dmin/dmax[0]; copies of A, b, S, r, C, dmin, dmax of the root are kept
 * so that x0 can be changed afterwards (tree_qp_common.c:404-525) */
void tree_qp_in_eliminate_x0(tree_qp_in *const qp)
{
    const int nx0 = qp->nx[0], nc0 = qp->nc[0];
    if (nx0 == 0) return;
    qp_internal_t *im = &qp->internal_memory;
    if (im->is_C_initialized == 0 && nc0 > 0)
    {
        dmat_copy(&qp->C[0], &im->C0);
        im->is_C_initialized = 1;
    }
    if (im->is_dmin_initialized == 0 && nc0 > 0) memcpy(im->dmin0.pa, qp->dmin[0].pa, sizeof(double) * (size_t) nc0);
    if (im->is_dmax_initialized == 0 && nc0 > 0) memcpy(im->dmax0.pa, qp->dmax[0].pa, sizeof(double) * (size_t) nc0);
    qp->C[0].pA = NULL;
    qp->C[0].n = 0;
    if (im->is_S_initialized == 0)
    {
        dmat_copy(&qp->S[0], &im->S0);
        im->is_S_initialized = 1;
    }
    if (im->is_r_initialized == 0)
    {
        memcpy(im->r0.pa, qp->r[0].pa, sizeof(double) * (size_t) qp->nu[0]);
        im->is_r_initialized = 1;
    }
    qp->S[0].pA = NULL;
    qp->S[0].n = 0;
    for (int ii = 0; ii < qp->tree[0].nkids; ii++)
    {
        if (im->is_A_initialized[ii] == 0)
        {
            dmat_copy(&qp->A[ii], &im->A0[ii]);
            im->is_A_initialized[ii] = 1;
        }
        if (im->is_b_initialized[ii] == 0)
        {
            memcpy(im->b0[ii].pa, qp->b[ii].pa, sizeof(double) * (size_t) qp->b[ii].m);
            im->is_b_initialized[ii] = 1;
        }
        qp->A[ii].n = 0;
        qp->A[ii].pA = NULL;
    }
    for (int i = 0; i < nx0; i++) assert(fabs(qp->xmin[0].pa[i] - qp->xmax[0].pa[i]) < 1e-10);
    qp->nx[0] = 0;
    tree_qp_in_set_x0_strvec(qp, &qp->xmin[0]);
    qp->Q[0].pA = NULL; qp->Q[0].m = 0; qp->Q[0].n = 0;
    qp->q[0].pa = NULL; qp->q[0].m = 0;
    qp->xmin[0].pa = NULL; qp->xmin[0].m = 0;
    qp->xmax[0].pa = NULL; qp->xmax[0].m = 0;
}

void tree_qp_out_eliminate_x0(tree_qp_out *const qp_out)
{
    qp_out->x[0].pa = NULL; qp_out->x[0].m = 0;
    qp_out->mu_x[0].pa = NULL; qp_out->mu_x[0].m = 0;
}

/* diagonal weights; the cross term is zeroed (tree_qp_common.c:1377-1430) */
void tree_qp_in_set_node_objective_diag(const double *const Qd, const double *const Rd,
                                        const double *const q, const double *const r,
                                        tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    int nx = qp->nx[indx], nu = qp->nu[indx];
    if (nx > 0 && nu > 0) dmat_zero(&qp->S[indx]);
    if (nx > 0)
    {
        dmat_zero(&qp->Q[indx]);
        for (int j = 0; j < nx; j++) BLASFEO_DMATEL(&qp->Q[indx], j, j) = Qd[j];
        memcpy(qp->q[indx].pa, q, sizeof(double) * (size_t) nx);
    }
    if (nu > 0)
    {
        dmat_zero(&qp->R[indx]);
        for (int j = 0; j < nu; j++) BLASFEO_DMATEL(&qp->R[indx], j, j) = Rd[j];
        memcpy(qp->r[indx].pa, r, sizeof(double) * (size_t) nu);
    }
}

void tree_qp_in_set_node_bounds(const double *const xmin, const double *const xmax,
                                const double *const umin, const double *const umax,
                                tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    size_t bx = sizeof(double) * (size_t) qp->nx[indx], bu = sizeof(double) * (size_t) qp->nu[indx];
    if (xmin && bx) memcpy(qp->xmin[indx].pa, xmin, bx);
    if (xmax && bx) memcpy(qp->xmax[indx].pa, xmax, bx);
    if (umin && bu) memcpy(qp->umin[indx].pa, umin, bu);
    if (umax && bu) memcpy(qp->umax[indx].pa, umax, bu);
}

/* general stage constraints dmin <= C x + D u <= dmax (tree_qp_common.c:1700-1821); lda < 0 = tight */
void tree_qp_in_set_node_C_colmajor(const double *const C, const int lda, tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    const int nc = qp->nc[indx], nx = qp->nx[indx];
    if (C && nc > 0 && nx > 0) tqb_dmat_from_colmajor(nc, nx, C, lda < 0 ? nc : lda, &qp->C[indx]);
}

void tree_qp_in_set_node_D_colmajor(const double *const D, const int lda, tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    const int nc = qp->nc[indx], nu = qp->nu[indx];
    if (D && nc > 0 && nu > 0) tqb_dmat_from_colmajor(nc, nu, D, lda < 0 ? nc : lda, &qp->D[indx]);
}

void tree_qp_in_set_node_dmin(const double *const dmin, tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    if (dmin && qp->nc[indx] > 0) memcpy(qp->dmin[indx].pa, dmin, sizeof(double) * (size_t) qp->nc[indx]);
}

void tree_qp_in_set_node_dmax(const double *const dmax, tree_qp_in *const qp, const int indx)
{
    assert(indx >= 0 && indx < qp->N);
    if (dmax && qp->nc[indx] > 0) memcpy(qp->dmax[indx].pa, dmax, sizeof(double) * (size_t) qp->nc[indx]);
}

void tree_qp_in_set_node_general_constraints(const double *const C, const double *const D, const double *const dmin,
                                             const double *const dmax, tree_qp_in *const qp, const int indx)
{
    tree_qp_in_set_node_C_colmajor(C, -1, qp, indx);
    tree_qp_in_set_node_D_colmajor(D, -1, qp, indx);
    tree_qp_in_set_node_dmin(dmin, qp, indx);
    tree_qp_in_set_node_dmax(dmax, qp, indx);
}

void tree_qp_out_get_node_mu_d(double *const mu, const tree_qp_out *const out, const int i)
{
    memcpy(mu, out->mu_d[i].pa, sizeof(double) * (size_t) out->mu_d[i].m);
}

void tree_qp_in_set_ltv_dynamics_colmajor(double *A, double *B, double *b, tree_qp_in *qp)
{
    size_t ia = 0, ib = 0, iv = 0;
    for (int e = 0; e < qp->N - 1; e++)
    {
        tree_qp_in_set_edge_dynamics_colmajor(A + ia, B + ib, b + iv, qp, e);
        ia += (size_t) qp->A[e].m * qp->A[e].n;
        ib += (size_t) qp->B[e].m * qp->B[e].n;
        iv += (size_t) qp->b[e].m;
    }
}

void tree_qp_in_set_ltv_objective_diag(double *Qd, double *Rd, double *q, double *r, tree_qp_in *qp)
{
    size_t ix = 0, iu = 0;
    for (int i = 0; i < qp->N; i++)
    {
        tree_qp_in_set_node_objective_diag(Qd + ix, Rd + iu, q + ix, r + iu, qp, i);
        ix += (size_t) qp->nx[i];
        iu += (size_t) qp->nu[i];
    }
}

void tree_qp_in_set_ltv_bounds(double *xmin, double *xmax, double *umin, double *umax, tree_qp_in *qp)
{
    size_t ix = 0, iu = 0;
    for (int i = 0; i < qp->N; i++)
    {
        tree_qp_in_set_node_bounds(xmin + ix, xmax + ix, umin + iu, umax + iu, qp, i);
        ix += (size_t) qp->nx[i];
        iu += (size_t) qp->nu[i];
    }
}

/* ------------------------------------------------------------------------ tree_qp_out ---- */

int tree_qp_out_calculate_size(const int Nn, const int *const nx, const int *const nu,
                               const int *const nc)
{
    long long bytes = (long long) (6 * Nn - 1) * sizeof(struct blasfeo_dvec);
    for (int i = 0; i < Nn; i++)
    {
        bytes += 2 * tqb_dvec_bytes(nx[i]) + 2 * tqb_dvec_bytes(nu[i]) + tqb_dvec_bytes(nc ? nc[i] : 0);
        if (i > 0) bytes += tqb_dvec_bytes(nx[i]);
    }
    bytes = (bytes + 63) / 64 * 64 + 64;
    return (int) bytes;
}

void tree_qp_out_create(const int Nn, const int *const nx, const int *const nu,
                        const int *const nc, tree_qp_out *const out, void *ptr)
{
    char *cur = (char *) ptr;
    out->x = (struct blasfeo_dvec *) cur;    cur += (size_t) Nn * sizeof(struct blasfeo_dvec);
    out->u = (struct blasfeo_dvec *) cur;    cur += (size_t) Nn * sizeof(struct blasfeo_dvec);
    out->mu_x = (struct blasfeo_dvec *) cur; cur += (size_t) Nn * sizeof(struct blasfeo_dvec);
    out->mu_u = (struct blasfeo_dvec *) cur; cur += (size_t) Nn * sizeof(struct blasfeo_dvec);
    out->mu_d = (struct blasfeo_dvec *) cur; cur += (size_t) Nn * sizeof(struct blasfeo_dvec);
    out->lam = (struct blasfeo_dvec *) cur;  cur += (size_t) (Nn - 1) * sizeof(struct blasfeo_dvec);
    align64(&cur);
    for (int i = 0; i < Nn; i++)
    {
        carve_dvec(nx[i], &out->x[i], &cur);
        carve_dvec(nu[i], &out->u[i], &cur);
        carve_dvec(nx[i], &out->mu_x[i], &cur);
        carve_dvec(nu[i], &out->mu_u[i], &cur);
        carve_dvec(nc ? nc[i] : 0, &out->mu_d[i], &cur);
        if (i > 0) carve_dvec(nx[i], &out->lam[i - 1], &cur);
    }
    out->info.Nn = Nn;
    out->info.iter = 0;
    out->info.total_time = out->info.solver_time = out->info.interface_time = 0.0;
    assert(cur <= (char *) ptr + tree_qp_out_calculate_size(Nn, nx, nu, nc));
}

void tree_qp_out_get_node_x(double *const x, const tree_qp_out *const out, const int i)
{
    memcpy(x, out->x[i].pa, sizeof(double) * (size_t) out->x[i].m);
}
void tree_qp_out_get_node_u(double *const u, const tree_qp_out *const out, const int i)
{
    memcpy(u, out->u[i].pa, sizeof(double) * (size_t) out->u[i].m);
}
void tree_qp_out_get_edge_lam(double *const lam, const tree_qp_out *const out, const int e)
{
    memcpy(lam, out->lam[e].pa, sizeof(double) * (size_t) out->lam[e].m);
}
void tree_qp_out_get_node_mu_x(double *const mu, const tree_qp_out *const out, const int i)
{
    memcpy(mu, out->mu_x[i].pa, sizeof(double) * (size_t) out->mu_x[i].m);
}
void tree_qp_out_get_node_mu_u(double *const mu, const tree_qp_out *const out, const int i)
{
    memcpy(mu, out->mu_u[i].pa, sizeof(double) * (size_t) out->mu_u[i].m);
}

/* ------------------------------------------------------------------------------- KKT ----- */

static double absmax(double acc, double v)
{
    v = fabs(v);
    return (v > acc || v != v) ? v : acc; /* NaN propagates */
}

/* Infinity norm of stationarity / dynamics / bound feasibility / complementarity residuals
 * with the sign conventions of tree_qp_common.c:540-765:
 *   Q x + S'u + q + mu_x - lam_n + sum_kids A_k' lam_k = 0
 *   R u + S x + r + mu_u        + sum_kids B_k' lam_k = 0
 *   A x_dad + B u_dad + b - x = 0,  mu > 0 <=> upper bound active. */
double tree_qp_out_max_KKT_res(const tree_qp_in *const qp, const tree_qp_out *const out)
{
    double err = 0.0;
    int mx = 0, mu_ = 0;
    for (int i = 0; i < qp->N; i++)
    {
        if (qp->nx[i] > mx) mx = qp->nx[i];
        if (qp->nu[i] > mu_) mu_ = qp->nu[i];
    }
    double *tx = (double *) malloc(sizeof(double) * (size_t) (mx + 1));
    double *tu = (double *) malloc(sizeof(double) * (size_t) (mu_ + 1));
    for (int i = 0; i < qp->N; i++)
    {
        int nx = qp->nx[i], nu = qp->nu[i];
        const double *x = out->x[i].pa, *u = out->u[i].pa;
        for (int r = 0; r < nx; r++)
        {
            double s = qp->q[i].pa[r] + out->mu_x[i].pa[r];
            for (int a = 0; a < qp->nc[i]; a++) s += BLASFEO_DMATEL(&qp->C[i], a, r) * out->mu_d[i].pa[a];
            for (int c = 0; c < nx; c++) s += BLASFEO_DMATEL(&qp->Q[i], r, c) * x[c];
            for (int c = 0; c < nu; c++) s += BLASFEO_DMATEL(&qp->S[i], c, r) * u[c];
            if (i > 0) s -= out->lam[i - 1].pa[r];
            tx[r] = s;
        }
        for (int r = 0; r < nu; r++)
        {
            double s = qp->r[i].pa[r] + out->mu_u[i].pa[r];
            for (int a = 0; a < qp->nc[i]; a++) s += BLASFEO_DMATEL(&qp->D[i], a, r) * out->mu_d[i].pa[a];
            for (int c = 0; c < nu; c++) s += BLASFEO_DMATEL(&qp->R[i], r, c) * u[c];
            for (int c = 0; c < nx; c++) s += BLASFEO_DMATEL(&qp->S[i], r, c) * x[c];
            tu[r] = s;
        }
        for (int j = 0; j < qp->tree[i].nkids; j++)
        {
            int k = qp->tree[i].kids[j];
            const double *lk = out->lam[k - 1].pa;
            for (int r = 0; r < nx; r++)
                for (int c = 0; c < qp->nx[k]; c++) tx[r] += BLASFEO_DMATEL(&qp->A[k - 1], c, r) * lk[c];
            for (int r = 0; r < nu; r++)
                for (int c = 0; c < qp->nx[k]; c++) tu[r] += BLASFEO_DMATEL(&qp->B[k - 1], c, r) * lk[c];
        }
        for (int r = 0; r < nx; r++) err = absmax(err, tx[r]);
        for (int r = 0; r < nu; r++) err = absmax(err, tu[r]);

        if (i > 0)
        {
            int p = qp->tree[i].dad;
            const double *xp = out->x[p].pa, *up = out->u[p].pa;
            for (int r = 0; r < nx; r++)
            {
                double s = qp->b[i - 1].pa[r] - x[r];
                for (int c = 0; c < qp->nx[p]; c++) s += BLASFEO_DMATEL(&qp->A[i - 1], r, c) * xp[c];
                for (int c = 0; c < qp->nu[p]; c++) s += BLASFEO_DMATEL(&qp->B[i - 1], r, c) * up[c];
                err = absmax(err, s);
            }
        }
        for (int r = 0; r < nx; r++)
        {
            double lo = qp->xmin[i].pa[r], hi = qp->xmax[i].pa[r], m = out->mu_x[i].pa[r];
            if (x[r] > hi) err = absmax(err, x[r] - hi);
            else if (x[r] < lo) err = absmax(err, lo - x[r]);
            err = absmax(err, m > 0 ? m * (x[r] - hi) : m * (lo - x[r]));
        }
        for (int r = 0; r < nu; r++)
        {
            double lo = qp->umin[i].pa[r], hi = qp->umax[i].pa[r], m = out->mu_u[i].pa[r];
            if (u[r] > hi) err = absmax(err, u[r] - hi);
            else if (u[r] < lo) err = absmax(err, lo - u[r]);
            err = absmax(err, m > 0 ? m * (u[r] - hi) : m * (lo - u[r]));
        }
        for (int a = 0; a < qp->nc[i]; a++)
        {   /* general constraints: feasibility and complementarity (tree_qp_common.c:723-755) */
            double gv = 0.0;
            for (int c = 0; c < nx; c++) gv += BLASFEO_DMATEL(&qp->C[i], a, c) * x[c];
            for (int c = 0; c < nu; c++) gv += BLASFEO_DMATEL(&qp->D[i], a, c) * u[c];
            const double lo = qp->dmin[i].pa[a], hi = qp->dmax[i].pa[a], m = out->mu_d[i].pa[a];
            if (gv > hi) err = absmax(err, gv - hi);
            else if (gv < lo) err = absmax(err, lo - gv);
            err = absmax(err, m > 0 ? m * (gv - hi) : m * (lo - gv));
        }
    }
    free(tx);
    free(tu);
    return err;
}
