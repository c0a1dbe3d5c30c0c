/*
 * oracle/tdunes_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (column-major, no BLASFEO, no SIMD, one thread) restatement of the reference's
 * TDUNES algorithm with the clipping stage solver and the dense stage-QP path
 * (dual_Newton_tree_qpoases.c; the box QP itself is restated in port_common.h), used only as the checker in tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under treeqp_b200/ may
 * include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the unmodified reference compiled from /root/reference (oracle/_ref/libtreeqp_ref.so)
 *       on spring-mass (examples/spring_mass_utils/data.c) and seeded random trees: same
 *       status, same Newton iteration count, same line-search trial counts, x/u/lam/mu to 1e-10;
 *   (b) the committed golden vectors tests/golden/ (.npz files) that were generated from the reference
 *       by tests/golden/make_golden.py.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "port_common.h"

typedef struct
{
    int Nn, Np, Nh;
    int *dad, *kid0, *idxpos, *stage, *xo, *uo, *ao, *bo, *dim, *wo, *uto;
    double *Qinv, *Rinv, *P, *Rt;          /* reciprocals and "elimination matrices" */
    double *x, *u, *xUnc, *uUnc, *qmod, *rmod;
    double *lam, *dlam, *res, *resMod;     /* indexed like x (segment of node k = lambda_k) */
    double *W, *L, *Ld, *Ut, *LUt;         /* per parent block */
    double *fval;
    int *xas, *uas;
    /* dense stage-QP path: per node offset of H / P (nv*nv), projector, multipliers, scratch */
    int *ho, nvmax;
    double *Pd, *mudx, *mudu, *scr;
    int fail;
} port_work;

static int is_dense(const port_qp *qp, int k) { return qp->solver != NULL && qp->solver[k] == 1; }

/* utils/tree.c:171-243 (breadth-first numbering: kids of node i follow kids of node i-1)
 * and dual_Newton_tree.c:177-194 (idxpos) */
static void port_setup(const port_qp *qp, port_work *w)
{
    int Nn = qp->Nn;
    w->Nn = Nn;
    w->dad = malloc(sizeof(int) * Nn);   w->kid0 = malloc(sizeof(int) * Nn);
    w->idxpos = malloc(sizeof(int) * Nn); w->stage = malloc(sizeof(int) * Nn);
    w->xo = malloc(sizeof(int) * (Nn + 1)); w->uo = malloc(sizeof(int) * (Nn + 1));
    w->ao = malloc(sizeof(int) * Nn);     w->bo = malloc(sizeof(int) * Nn);
    w->dim = malloc(sizeof(int) * Nn);    w->wo = malloc(sizeof(int) * (Nn + 1));
    w->uto = malloc(sizeof(int) * (Nn + 1));
    int next = 1;
    w->dad[0] = -1; w->stage[0] = 0; w->idxpos[0] = 0;
    w->xo[0] = 0; w->uo[0] = 0;
    for (int i = 0; i < Nn; i++)
    {
        w->kid0[i] = next;
        int pos = 0;
        for (int j = 0; j < qp->nk[i]; j++)
        {
            int k = next + j;
            w->dad[k] = i; w->stage[k] = w->stage[i] + 1; w->idxpos[k] = pos;
            pos += qp->nx[k];
        }
        w->dim[i] = pos;
        next += qp->nk[i];
        w->xo[i + 1] = w->xo[i] + qp->nx[i];
        w->uo[i + 1] = w->uo[i] + qp->nu[i];
    }
    int ia = 0, ib = 0, iw = 0, iu = 0, Np = 0;
    for (int k = 0; k < Nn; k++)
    {
        w->ao[k] = ia; w->bo[k] = ib;
        if (k > 0) { ia += qp->nx[k] * qp->nx[w->dad[k]]; ib += qp->nx[k] * qp->nu[w->dad[k]]; }
        w->wo[k] = iw; w->uto[k] = iu;
        if (qp->nk[k] > 0) { Np++; iw += w->dim[k] * w->dim[k]; iu += qp->nx[k] * w->dim[k]; }
    }
    w->wo[Nn] = iw; w->uto[Nn] = iu;
    w->Np = Np; w->Nh = w->stage[Nn - 1];
    int X = w->xo[Nn], U = w->uo[Nn];
#define D(n) calloc((size_t) (n) + 1, sizeof(double))
    w->Qinv = D(X); w->Rinv = D(U); w->P = D(X); w->Rt = D(U);
    w->x = D(X); w->u = D(U); w->xUnc = D(X); w->uUnc = D(U); w->qmod = D(X); w->rmod = D(U);
    w->lam = D(X); w->dlam = D(X); w->res = D(X); w->resMod = D(X);
    w->W = D(iw); w->L = D(iw); w->Ld = D(X); w->Ut = D(iu); w->LUt = D(iu);
    w->fval = D(Nn);
    w->ho = malloc(sizeof(int) * (Nn + 1));
    w->ho[0] = 0; w->nvmax = 0;
    for (int k = 0; k < Nn; k++)
    {
        int nv = qp->nx[k] + qp->nu[k];
        w->ho[k + 1] = w->ho[k] + nv * nv;
        if (nv > w->nvmax) w->nvmax = nv;
    }
    w->Pd = D(qp->solver ? w->ho[Nn] : 0);
    w->mudx = D(X); w->mudu = D(U);
    w->scr = D(4 * w->nvmax * w->nvmax + 16 * w->nvmax + 16);
    w->fail = 0;
#undef D
    w->xas = calloc((size_t) X + 1, sizeof(int)); w->uas = calloc((size_t) U + 1, sizeof(int));
}

static void port_free(port_work *w)
{
    free(w->dad); free(w->kid0); free(w->idxpos); free(w->stage); free(w->xo); free(w->uo);
    free(w->ao); free(w->bo); free(w->dim); free(w->wo); free(w->uto);
    free(w->Qinv); free(w->Rinv); free(w->P); free(w->Rt); free(w->x); free(w->u);
    free(w->xUnc); free(w->uUnc); free(w->qmod); free(w->rmod); free(w->lam); free(w->dlam);
    free(w->res); free(w->resMod); free(w->W); free(w->L); free(w->Ld); free(w->Ut); free(w->LUt);
    free(w->fval); free(w->xas); free(w->uas);
    free(w->ho); free(w->Pd); free(w->mudx); free(w->mudu); free(w->scr);
}

/* dual_Newton_tree_qpoases.c:311-357 (QProblem_solve: hotstart at g = -[qmod; rmod]) and, when
 * `extended`, :153-214 (projector P) -- the stage solution, multipliers and working set of node k */
static void port_dense_stage(const port_qp *qp, port_work *w, int k, int extended)
{
    int nx = qp->nx[k], nu = qp->nu[k], nv = nx + nu, xo = w->xo[k], uo = w->uo[k];
    const double *H = qp->H + w->ho[k];
    double *g = w->scr, *lb = g + nv, *ub = lb + nv, *z = ub + nv, *mu = z + nv, *t = mu + nv;
    double *L = t + 3 * nv + nv * nv;
    int ws[64];
    for (int i = 0; i < nx; i++) { g[i] = w->qmod[xo + i]; lb[i] = qp->xmin[xo + i]; ub[i] = qp->xmax[xo + i]; ws[i] = w->xas[xo + i]; }
    for (int i = 0; i < nu; i++) { g[nx + i] = w->rmod[uo + i]; lb[nx + i] = qp->umin[uo + i]; ub[nx + i] = qp->umax[uo + i]; ws[nx + i] = w->uas[uo + i]; }
    if (port_boxqp(nv, H, g, lb, ub, ws, z, mu, L, t)) w->fail = 1;
    for (int i = 0; i < nx; i++) w->x[xo + i] = z[i];
    for (int i = 0; i < nu; i++) w->u[uo + i] = z[nx + i];
    if (!extended) return;
    for (int i = 0; i < nx; i++) { w->xas[xo + i] = ws[i]; w->mudx[xo + i] = mu[i]; }
    for (int i = 0; i < nu; i++) { w->uas[uo + i] = ws[nx + i]; w->mudu[uo + i] = mu[nx + i]; }
    port_boxqp_projector(nv, H, ws, w->Pd + w->ho[k], w->scr + 8 * w->nvmax);
}

/* dual_Newton_tree.c:259-292 / :862-903: modified stage gradients at the dual point lam.
 * Returns cmod_k = sum_kids b_j' lam_j (dual_Newton_tree.c:892). */
static double port_stage_gradient(const port_qp *qp, port_work *w, int k)
{
    int nx = qp->nx[k], nu = qp->nu[k];
    double *qm = w->qmod + w->xo[k], *rm = w->rmod + w->uo[k];
    for (int i = 0; i < nx; i++)
        qm[i] = -qp->q[w->xo[k] + i] + (k > 0 ? w->lam[w->xo[k] + i] : 0.0);
    for (int i = 0; i < nu; i++) rm[i] = -qp->r[w->uo[k] + i];
    double cmod = 0.0;
    for (int j = 0; j < qp->nk[k]; j++)
    {
        int c = w->kid0[k] + j, nxc = qp->nx[c];
        const double *Ac = qp->A + w->ao[c], *Bc = qp->B + w->bo[c], *lc = w->lam + w->xo[c];
        double dot = 0.0;
        for (int t = 0; t < nxc; t++) dot += qp->b[w->xo[c] - qp->nx[0] + t] * lc[t];
        cmod += dot;
        for (int i = 0; i < nx; i++)
        {
            double s = 0.0;
            for (int t = 0; t < nxc; t++) s += Ac[t + i * nxc] * lc[t];
            qm[i] -= s;
        }
        for (int i = 0; i < nu; i++)
        {
            double s = 0.0;
            for (int t = 0; t < nxc; t++) s += Bc[t + i * nxc] * lc[t];
            rm[i] -= s;
        }
    }
    return cmod;
}

/* dual_Newton_tree.c:218-330 + dual_Newton_tree_clipping.c:188-227 */
static void port_solve_stage_problems(const port_qp *qp, port_work *w)
{
    for (int k = 0; k < qp->Nn; k++)
    {
        port_stage_gradient(qp, w, k);
        if (is_dense(qp, k)) { port_dense_stage(qp, w, k, 1); continue; }
        for (int i = 0; i < qp->nx[k]; i++)
        {
            int g = w->xo[k] + i;
            w->xUnc[g] = w->Qinv[g] * w->qmod[g];
            w->x[g] = clip(w->xUnc[g], qp->xmin[g], qp->xmax[g], &w->xas[g]);
            w->P[g] = w->xas[g] == 0 ? w->Qinv[g] : 0.0;
        }
        for (int i = 0; i < qp->nu[k]; i++)
        {
            int g = w->uo[k] + i;
            w->uUnc[g] = w->Rinv[g] * w->rmod[g];
            w->u[g] = clip(w->uUnc[g], qp->umin[g], qp->umax[g], &w->uas[g]);
            w->Rt[g] = w->uas[g] == 0 ? w->Rinv[g] : 0.0;
        }
    }
}

/* dual_Newton_tree.c:823-918 + clipping.c:231-260 (solve) + clipping.c:359-382 (dual term);
 * the sum over nodes is serial in node order (dual_Newton_tree.c:915) */
static double port_evaluate_dual_function(const port_qp *qp, port_work *w)
{
    double total = 0.0;
    for (int k = 0; k < qp->Nn; k++)
    {
        double cmod = port_stage_gradient(qp, w, k);
        double xQx = 0.0, xq = 0.0, uRu = 0.0, ur = 0.0;
        int dummy;
        if (is_dense(qp, k))
        {   /* dual_Newton_tree_qpoases.c:478-523: x'Qx and u'Ru only -- the reference leaves the
             * cross term u'Sx out of the dual function value, and so does this restatement */
            int nx = qp->nx[k], nu = qp->nu[k], nv = nx + nu;
            const double *H = qp->H + w->ho[k];
            const double *xk = w->x + w->xo[k], *uk = w->u + w->uo[k];
            port_dense_stage(qp, w, k, 0);
            for (int j = 0; j < nx; j++)
                for (int i = 0; i < nx; i++) xQx += xk[i] * H[i + j * nv] * xk[j];
            for (int j = 0; j < nu; j++)
                for (int i = 0; i < nu; i++) uRu += uk[i] * H[(nx + i) + (nx + j) * nv] * uk[j];
            for (int i = 0; i < nx; i++) xq += w->qmod[w->xo[k] + i] * xk[i];
            for (int i = 0; i < nu; i++) ur += w->rmod[w->uo[k] + i] * uk[i];
        }
        else
        {
        for (int i = 0; i < qp->nx[k]; i++)
        {
            int g = w->xo[k] + i;
            w->x[g] = clip(w->Qinv[g] * w->qmod[g], qp->xmin[g], qp->xmax[g], &dummy);
            xQx += (qp->Qd[g] * w->x[g]) * w->x[g];
            xq += w->qmod[g] * w->x[g];
        }
        for (int i = 0; i < qp->nu[k]; i++)
        {
            int g = w->uo[k] + i;
            w->u[g] = clip(w->Rinv[g] * w->rmod[g], qp->umin[g], qp->umax[g], &dummy);
            uRu += (qp->Rd[g] * w->u[g]) * w->u[g];
            ur += w->rmod[g] * w->u[g];
        }
        }
        double f = -0.5 * xQx - cmod;
        f += xq;
        f -= 0.5 * uRu;
        f += ur;
        w->fval[k] = f;
    }
    for (int k = 0; k < qp->Nn; k++) total += w->fval[k];
    return total;
}

/* dual_Newton_tree.c:519-539 (residual) and :412-442 (norm) */
static double port_dual_gradient(const port_qp *qp, port_work *w, int cond)
{
    for (int k = qp->Nn - 1; k > 0; k--)
    {
        int p = w->dad[k], nx = qp->nx[k], nxp = qp->nx[p], nup = qp->nu[p];
        const double *Ak = qp->A + w->ao[k], *Bk = qp->B + w->bo[k];
        for (int i = 0; i < nx; i++)
        {
            int g = w->xo[k] + i;
            double s = qp->b[g - qp->nx[0]] - w->x[g];
            for (int c = 0; c < nxp; c++) s += Ak[i + c * nx] * w->x[w->xo[p] + c];
            for (int c = 0; c < nup; c++) s += Bk[i + c * nx] * w->u[w->uo[p] + c];
            w->res[g] = s;
            w->resMod[g] = s;
        }
    }
    double err = 0.0;
    int X = w->xo[qp->Nn];
    for (int g = qp->nx[0]; g < X; g++)
    {
        if (cond == 2) { double a = fabs(w->res[g]); if (a > err) err = a; }
        else err += w->res[g] * w->res[g];
    }
    if (cond == 1) err = sqrt(err);
    return err;
}

/* dual_Newton_tree.c:551-615 with checkLastActiveSet == 0 semantics (every block rebuilt; the
 * cached path of the reference restores bit-identical blocks) + clipping.c:264-355.
 * W_p[k,s] = A_k P_p A_s' + B_k Rt_p B_s' (+ P_k on the diagonal block);
 * Ut_p[:,k] = -P_p A_k'. */
static void port_dual_hessian(const port_qp *qp, port_work *w)
{
    for (int p = 0; p < qp->Nn; p++)
    {
        if (qp->nk[p] == 0) continue;
        int d = w->dim[p], nxp = qp->nx[p], nup = qp->nu[p];
        double *Wp = w->W + w->wo[p], *Utp = w->Ut + w->uto[p];
        const double *Pp = w->P + w->xo[p], *Rp = w->Rt + w->uo[p];
        if (is_dense(qp, p))
        {   /* dual_Newton_tree_qpoases.c:401-475: M_s = C_s P_p (C = [A B]), W[k,s] = C_k M_s',
             * Ut[:,k] = -(first nx_p columns of M_k)'; the kid's own term is the xx block of ITS
             * elimination matrix (dense kid, :430-445) or diag(P_k) (clipping kid, clipping.c:301) */
            int nvp = nxp + nup;
            const double *PP = w->Pd + w->ho[p];
            double *M = malloc(sizeof(double) * (size_t) d * nvp + 8);
            for (int jk = 0; jk < qp->nk[p]; jk++)
            {
                int k = w->kid0[p] + jk, nxk = qp->nx[k], rk = w->idxpos[k];
                const double *Ak = qp->A + w->ao[k], *Bk = qp->B + w->bo[k];
                for (int i = 0; i < nxk; i++)
                    for (int a = 0; a < nvp; a++)
                    {
                        double acc = 0.0;
                        for (int c = 0; c < nvp; c++)
                            acc += (c < nxp ? Ak[i + c * nxk] : Bk[i + (c - nxp) * nxk]) * PP[c + a * nvp];
                        M[(rk + i) + a * d] = acc;
                    }
            }
            for (int jk = 0; jk < qp->nk[p]; jk++)
            {
                int k = w->kid0[p] + jk, nxk = qp->nx[k], rk = w->idxpos[k], nvk = nxk + qp->nu[k];
                const double *Ak = qp->A + w->ao[k], *Bk = qp->B + w->bo[k];
                for (int i = 0; i < nxk; i++)
                    for (int J = 0; J <= rk + i; J++)
                    {
                        double acc = 0.0;
                        for (int c = 0; c < nvp; c++)
                            acc += (c < nxp ? Ak[i + c * nxk] : Bk[i + (c - nxp) * nxk]) * M[J + c * d];
                        if (J >= rk)
                        {
                            if (is_dense(qp, k)) acc += w->Pd[w->ho[k] + i + (J - rk) * nvk];
                            else if (J == rk + i) acc += w->P[w->xo[k] + i];
                        }
                        Wp[(rk + i) + J * d] = acc;
                    }
                if (w->dad[p] >= 0)
                    for (int i = 0; i < nxk; i++)
                        for (int c = 0; c < nxp; c++) Utp[c + (rk + i) * nxp] = -M[(rk + i) + c * d];
            }
            free(M);
            continue;
        }
        for (int jk = 0; jk < qp->nk[p]; jk++)
        {
            int k = w->kid0[p] + jk, nxk = qp->nx[k], rk = w->idxpos[k];
            const double *Ak = qp->A + w->ao[k], *Bk = qp->B + w->bo[k];
            for (int js = 0; js <= jk; js++)
            {
                int s = w->kid0[p] + js, nxs = qp->nx[s], cs = w->idxpos[s];
                const double *As = qp->A + w->ao[s], *Bs = qp->B + w->bo[s];
                for (int i = 0; i < nxk; i++)
                    for (int j = 0; j < nxs; j++)
                    {
                        if (s == k && j > i) continue; /* lower triangle only (dsyrk_ln) */
                        double acc = 0.0;
                        for (int c = 0; c < nxp; c++) acc += Ak[i + c * nxk] * (As[j + c * nxs] * Pp[c]);
                        double acc2 = 0.0;
                        for (int c = 0; c < nup; c++) acc2 += Bk[i + c * nxk] * (Bs[j + c * nxs] * Rp[c]);
                        acc += acc2;
                        if (s == k && is_dense(qp, k)) acc += w->Pd[w->ho[k] + i + j * (nxk + qp->nu[k])];
                        else if (s == k && i == j) acc += w->P[w->xo[k] + i];
                        Wp[(rk + i) + (cs + j) * d] = acc;
                    }
            }
            if (w->dad[p] >= 0)
                for (int i = 0; i < nxk; i++)
                    for (int c = 0; c < nxp; c++) Utp[c + (rk + i) * nxp] = -(Ak[i + c * nxk] * Pp[c]);
        }
    }
}

/* dual_Newton_tree.c:641-775: leaves-to-root block Cholesky merged with the backward
 * substitution, then root-to-leaves forward substitution.  Blocks are visited by decreasing
 * node index, which visits depths Nh-1..0 (breadth-first numbering). */
static int port_newton_direction(const port_qp *qp, port_work *w, const port_opts *o)
{
    int nreg = 0;
    for (int p = qp->Nn - 1; p >= 0; p--)
    {
        if (qp->nk[p] == 0) continue;
        int d = w->dim[p], nxp = qp->nx[p];
        double *Wp = w->W + w->wo[p], *Lp = w->L + w->wo[p];
        double *inv = w->Ld + w->xo[w->kid0[p]];
        double *y = w->dlam + w->xo[w->kid0[p]];
        const double *rm = w->resMod + w->xo[w->kid0[p]];
        nreg += port_potrf_reg(d, Wp, Lp, inv, o);
        /* dtrsv_lnn: y = L^-1 resMod (uses the stored inverse diagonal) */
        for (int i = 0; i < d; i++)
        {
            double s = rm[i];
            for (int t = 0; t < i; t++) s -= Lp[i + t * d] * y[t];
            y[i] = s * inv[i];
        }
        if (p == 0) break;
        /* dtrsm_rltn: LUt = Ut L^-T */
        const double *Utp = w->Ut + w->uto[p];
        double *Xp = w->LUt + w->uto[p];
        for (int j = 0; j < d; j++)
            for (int c = 0; c < nxp; c++)
            {
                double s = Utp[c + j * nxp];
                for (int t = 0; t < j; t++) s -= Xp[c + t * nxp] * Lp[j + t * d];
                Xp[c + j * nxp] = s * inv[j];
            }
        /* dsyrk_ln / dgemv_n into the parent's block and right-hand side */
        int pp = w->dad[p], dp = w->dim[pp], off = w->idxpos[p];
        double *Wpp = w->W + w->wo[pp];
        double *rmp = w->resMod + w->xo[w->kid0[pp]] + off;
        for (int a = 0; a < nxp; a++)
        {
            for (int bb = 0; bb <= a; bb++)
            {
                double s = 0.0;
                for (int t = 0; t < d; t++) s += Xp[a + t * nxp] * Xp[bb + t * nxp];
                Wpp[(off + a) + (off + bb) * dp] -= s;
            }
            double s = 0.0;
            for (int t = 0; t < d; t++) s += Xp[a + t * nxp] * y[t];
            rmp[a] -= s;
        }
    }
    for (int p = 0; p < qp->Nn; p++)
    {
        if (qp->nk[p] == 0) continue;
        int d = w->dim[p], nxp = qp->nx[p];
        const double *Lp = w->L + w->wo[p], *inv = w->Ld + w->xo[w->kid0[p]];
        double *y = w->dlam + w->xo[w->kid0[p]];
        if (p > 0)
        {   /* dgemv_t: y -= LUt' * dlam_dad[segment of p] */
            const double *Xp = w->LUt + w->uto[p];
            const double *dp = w->dlam + w->xo[p];
            for (int j = 0; j < d; j++)
            {
                double s = 0.0;
                for (int c = 0; c < nxp; c++) s += Xp[c + j * nxp] * dp[c];
                y[j] -= s;
            }
        }
        /* dtrsv_ltn */
        for (int i = d - 1; i >= 0; i--)
        {
            double s = y[i];
            for (int t = i + 1; t < d; t++) s -= Lp[t + i * d] * y[t];
            y[i] = s * inv[i];
        }
    }
    return nreg;
}

/* Entry point.  Restates treeqp_tdunes_solve (dual_Newton_tree.c:1104-1263) including the
 * line search (:922-1019) and the multiplier export (clipping.c:386-399).
 * Outputs: x,u (concatenated per node), lam (per edge, node order), mu_x, mu_u, iteration
 * count, per-iteration line-search trial counts (first maxIter entries), final active sets.
 * Returns the reference's status code (0 optimal, 1 max iter, 2 not a descent direction). */
int oracle_tdunes_solve(const port_qp *qp, const port_opts *o, const double *lambda0,
                        double *x, double *u, double *lam, double *mu_x, double *mu_u,
                        int *iter_out, int *ls_iters, int *xas, int *uas, int *nreg_out)
{
    port_work W, *w = &W;
    port_setup(qp, w);
    int Nn = qp->Nn, X = w->xo[Nn], U = w->uo[Nn], nx0 = qp->nx[0];
    int status = 7, it, nreg = 0;
    for (int g = 0; g < X; g++) w->Qinv[g] = 1.0 / qp->Qd[g];   /* clipping.c:163-170 */
    for (int g = 0; g < U; g++) w->Rinv[g] = 1.0 / qp->Rd[g];
    /* dense nodes start from an empty working set (QProblemB_init is a cold start, qpoases.c:286) */
    for (int g = nx0; g < X; g++) w->lam[g] = lambda0[g - nx0];

    for (it = 0; it < o->maxIter; it++)
    {
        port_solve_stage_problems(qp, w);
        double err = port_dual_gradient(qp, w, o->termCondition);
        if (err < o->tol) { status = 0; break; }
        port_dual_hessian(qp, w);
        nreg += port_newton_direction(qp, w, o);

        /* line search, dual_Newton_tree.c:922-1019 */
        double dot = 0.0;
        for (int p = 0; p < Nn; p++)
        {   /* per-block ddot accumulated over blocks (dual_Newton_tree.c:814-817) */
            if (qp->nk[p] == 0) continue;
            double s = 0.0;
            int o0 = w->xo[w->kid0[p]];
            for (int i = 0; i < w->dim[p]; i++) s += w->res[o0 + i] * w->dlam[o0 + i];
            dot += s;
        }
        dot = -dot;
        double f0 = port_evaluate_dual_function(qp, w);
        if (dot > 1e-10 || !((dot > 1e-10) || (dot < 1e-10))) { status = 2; break; }
        double tau = 1.0, tauPrev = 0.0;
        int ls;
        for (ls = 1; ls <= o->lsMaxIter; ls++)
        {
            for (int g = nx0; g < X; g++) w->lam[g] += (tau - tauPrev) * w->dlam[g];
            double f = port_evaluate_dual_function(qp, w);
            if (f <= f0 + o->gamma * tau * dot) break;
            tauPrev = tau;
            tau = o->beta * tauPrev;
        }
        if (ls_iters) ls_iters[it] = ls;
    }
    if (status == 2) { port_free(w); *iter_out = it; return status; } /* qp_out untouched */
    if (it == o->maxIter) status = 1;

    memcpy(x, w->x, sizeof(double) * X);
    memcpy(u, w->u, sizeof(double) * U);
    memcpy(lam, w->lam + nx0, sizeof(double) * (X - nx0));
    for (int g = 0; g < X; g++) mu_x[g] = qp->Qd[g] * (w->xUnc[g] - x[g]);
    for (int g = 0; g < U; g++) mu_u[g] = qp->Rd[g] * (w->uUnc[g] - u[g]);
    for (int k = 0; k < Nn; k++)
    {   /* dual_Newton_tree_qpoases.c:527-565: multipliers of the last extended solve, sign flipped */
        if (!is_dense(qp, k)) continue;
        for (int i = 0; i < qp->nx[k]; i++) mu_x[w->xo[k] + i] = w->mudx[w->xo[k] + i];
        for (int i = 0; i < qp->nu[k]; i++) mu_u[w->uo[k] + i] = w->mudu[w->uo[k] + i];
    }
    if (w->fail) status = 4; /* TREEQP_DN_STAGE_QP_SOLVE_FAILED */
    if (xas) memcpy(xas, w->xas, sizeof(int) * X);
    if (uas) memcpy(uas, w->uas, sizeof(int) * U);
    *iter_out = it;
    if (nreg_out) *nreg_out = nreg;
    port_free(w);
    return status;
}
