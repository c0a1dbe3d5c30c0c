/*
 * shard.c -- subtree sharding of one tree QP over the GPUs of a box (third shape):
 * the plan (which rank owns which subtree, what is replicated, what is a ghost), the local forest
 * built from it, and the mapped pack / unpack between the caller's FULL tree_qp_in / tree_qp_out
 * and a rank's local arrays.  The plan is pure host logic (no GPU needed): tests/test_shard_cpu.py
 * checks it across two gloo processes.
 *
 * Why a cut at one depth: the dual Newton system is block tridiagonal on the tree
 * (dual_Newton_tree.c:641-775), so everything below a cut node -- stage QPs, gradient, Hessian
 * blocks, bottom-up factorisation, top-down substitution -- touches only that subtree; the single
 * coupling is the cut node's Schur complement and right-hand side in its parent's block.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "treeqp_b200.h"
#include "host_internal.h"

tqb_shard_plan *tqb_shard_plan_create(const tree_qp_in *qp, int rank, int nranks)
{
    const int N = qp->N;
    const struct node *tree = qp->tree;
    const int Nh = tree[N - 1].stage;
    if (nranks < 1 || rank < 0 || rank >= nranks) return NULL;
    /* nodes per depth (breadth-first numbering: a depth is a contiguous index range) */
    int *first = (int *) calloc((size_t) Nh + 2, sizeof(int)), *count = (int *) calloc((size_t) Nh + 2, sizeof(int));
    for (int k = N - 1; k >= 0; k--) { first[tree[k].stage] = k; count[tree[k].stage]++; }
    int c = -1;
    for (int d = 1; d < Nh; d++)
        if (count[d] >= nranks && count[d] % nranks == 0)
        {
            int all_parents = 1;
            for (int k = first[d]; k < first[d] + count[d]; k++) all_parents &= tree[k].nkids > 0;
            if (all_parents) { c = d; break; }
        }
    if (c < 0)
    {
        fprintf(stderr, "[treeqp_b200] cannot cut this tree into %d equal groups of subtrees\n", nranks);
        free(first); free(count);
        return NULL;
    }
    tqb_shard_plan *p = (tqb_shard_plan *) calloc(1, sizeof *p);
    p->rank = rank; p->nranks = nranks;
    p->cut_depth = c; p->ncut = count[c]; p->cut_first = first[c];
    p->my_ncut = p->ncut / nranks; p->my_cut0 = rank * p->my_ncut;
    const int n_top = first[c] + count[c]; /* nodes of depth <= c */
    free(first); free(count);

    /* local node list: the top, then breadth first below this rank's cut nodes */
    int cap = n_top + 16, Nl = n_top;
    int *gmap = (int *) malloc(sizeof(int) * (size_t) cap);
    for (int k = 0; k < n_top; k++) gmap[k] = k;
    for (int j = p->cut_first; j < Nl; j++)
    {   /* j runs over local nodes of depth >= c; the list grows while we walk it */
        const int g = gmap[j];
        const int mine = j >= n_top || (j - p->cut_first >= p->my_cut0 && j - p->cut_first < p->my_cut0 + p->my_ncut);
        if (!mine) continue;
        for (int q = 0; q < tree[g].nkids; q++)
        {
            if (Nl == cap) { cap *= 2; gmap = (int *) realloc(gmap, sizeof(int) * (size_t) cap); }
            gmap[Nl++] = tree[g].kids[q];
        }
    }
    p->Nl = Nl;
    p->gmap = gmap;
    p->ghost = (int *) calloc((size_t) Nl, sizeof(int));
    p->owned = (int *) calloc((size_t) Nl, sizeof(int));
    p->ltree = (struct node *) calloc((size_t) Nl, sizeof(struct node));
    p->lkids = (int *) malloc(sizeof(int) * (size_t) (Nl > 1 ? Nl : 1));
    p->lnx = (int *) malloc(sizeof(int) * (size_t) Nl);
    p->lnu = (int *) malloc(sizeof(int) * (size_t) Nl);
    p->lnc = (int *) calloc((size_t) Nl, sizeof(int));
    int next = 1, nkid_tot = 0;
    p->ltree[0].dad = -1;
    for (int j = 0; j < Nl; j++)
    {
        const int g = gmap[j];
        const int is_cut = j >= p->cut_first && j < n_top;
        const int mine_cut = is_cut && (j - p->cut_first >= p->my_cut0 && j - p->cut_first < p->my_cut0 + p->my_ncut);
        p->ghost[j] = is_cut && !mine_cut;
        p->owned[j] = j < p->cut_first ? rank == 0 : !p->ghost[j];
        p->lnx[j] = qp->nx[g];
        p->lnu[j] = qp->nu[g];
        struct node *nd = &p->ltree[j];
        nd->idx = j;
        nd->stage = tree[g].stage;
        nd->real = tree[g].real;
        nd->idxkid = tree[g].idxkid;
        nd->nkids = p->ghost[j] ? 0 : tree[g].nkids;
        nd->kids = p->lkids + nkid_tot;
        for (int q = 0; q < nd->nkids; q++)
        {
            assert(gmap[next] == tree[g].kids[q]);
            nd->kids[q] = next;
            p->ltree[next].dad = j;
            next++;
        }
        nkid_tot += nd->nkids;
    }
    assert(next == Nl);
    return p;
}

void tqb_shard_plan_destroy(tqb_shard_plan *p)
{
    if (!p) return;
    free(p->gmap); free(p->ghost); free(p->owned); free(p->ltree); free(p->lkids);
    free(p->lnx); free(p->lnu); free(p->lnc); free(p->lam_off);
    free(p);
}

tqb_forest *tqb_forest_create_sharded(const tree_qp_in *full_qp, const tqb_shard_plan *plan, void *comm)
{
    (void) full_qp;
    /* a shell tree_qp_in carrying only what tqb_forest_create_ex reads: N, nx, nu, nc, tree */
    tree_qp_in shell;
    memset(&shell, 0, sizeof shell);
    shell.N = plan->Nl;
    shell.nx = plan->lnx; shell.nu = plan->lnu; shell.nc = plan->lnc;
    shell.tree = plan->ltree;
    const tree_qp_in *one[1] = {&shell};
    tqb_forest *f = tqb_forest_create_ex(1, one, plan->ghost, plan->owned, NULL);
    if (!f) return NULL;
    int *cut_local = (int *) malloc(sizeof(int) * (size_t) plan->ncut);
    for (int q = 0; q < plan->ncut; q++) cut_local[q] = plan->cut_first + q; /* the top keeps its numbering */
    int rc = tqb_device_enable_sharding(f->dev, (tqb_comm *) comm, plan->cut_depth, plan->ncut, plan->my_cut0,
                                        plan->my_ncut, cut_local);
    free(cut_local);
    if (rc) { tqb_forest_destroy(f); return NULL; }
    return f;
}

/* mapped variants of forest.c's pack / set_lambda0 / unpack */
void tqb_forest_pack_mapped_(tqb_forest *f, int t, const tree_qp_in *qp, const int *gmap);

void tqb_forest_pack_sharded(tqb_forest *f, const tree_qp_in *full_qp, const tqb_shard_plan *plan)
{
    tqb_forest_pack_mapped_(f, 0, full_qp, plan->gmap);
}

void tqb_forest_set_lambda0_sharded(tqb_forest *f, const double *lambda_full, const tree_qp_in *full_qp,
                                    const tqb_shard_plan *plan)
{
    const tqb_tables *T = &f->ht;
    tqb_shard_plan *pl = (tqb_shard_plan *) plan; /* the offsets are a cache inside the plan */
    if (!pl->lam_off)
    {   /* offset of node k's multiplier in the caller's vector: sum of nx over nodes 1..k-1; kept per local node */
        long long *off = (long long *) malloc(sizeof(long long) * (size_t) (full_qp->N + 1));
        off[0] = 0; off[1] = 0;
        for (int k = 1; k < full_qp->N; k++) off[k + 1] = off[k] + full_qp->nx[k];
        pl->lam_off = (long long *) malloc(sizeof(long long) * (size_t) (plan->Nl > 0 ? plan->Nl : 1));
        for (int j = 0; j < plan->Nl; j++) pl->lam_off[j] = off[plan->gmap[j]];
        free(off);
    }
    for (int i = 0; i < plan->lnx[0]; i++) f->hin.lam0[T->xoff[0] + i] = 0.0;
#pragma omp parallel for schedule(static) if (plan->Nl > 4096)
    for (int j = 1; j < plan->Nl; j++)
        memcpy(f->hin.lam0 + T->xoff[j], lambda_full + pl->lam_off[j], sizeof(double) * (size_t) plan->lnx[j]);
}

void tqb_forest_unpack_sharded(const tqb_forest *f, const tqb_shard_plan *plan, tree_qp_out *out)
{
    const tqb_tables *T = &f->ht;
    const tqb_dims *d = &f->dims;
    const double *x = f->h_out, *u = x + d->X, *lam = u + d->U, *mux = lam + d->X, *muu = mux + d->X;
#pragma omp parallel for schedule(static) if (plan->Nl > 2048)
    for (int j = 0; j < plan->Nl; j++)
    {
        if (!plan->owned[j]) continue;
        const int k = plan->gmap[j], nx = T->nx[j], nu = T->nu[j], xo = T->xoff[j], uo = T->uoff[j];
        memcpy(out->x[k].pa, x + xo, sizeof(double) * (size_t) nx);
        memcpy(out->u[k].pa, u + uo, sizeof(double) * (size_t) nu);
        memcpy(out->mu_x[k].pa, mux + xo, sizeof(double) * (size_t) nx);
        memcpy(out->mu_u[k].pa, muu + uo, sizeof(double) * (size_t) nu);
        if (k > 0) memcpy(out->lam[k - 1].pa, lam + xo, sizeof(double) * (size_t) nx);
    }
}
