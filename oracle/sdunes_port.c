/*
 * oracle/sdunes_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (column-major, no BLASFEO, one thread) restatement of the reference's SDUNES
 * algorithm (dual Newton strategy on the scenario decomposition,
 * treeqp/src/dual_Newton_scenarios.c), used only as the checker in tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs.  Nothing under treeqp_b200/ may include,
 * link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the unmodified reference compiled from /root/reference (oracle/_ref/libtreeqp_ref.so)
 *       on the spring-mass data of examples/spring_mass_dual_newton_scenarios.c and on seeded
 *       linear_chain trees: same status, same Newton iteration count, x/u/lam/mu to 1e-10;
 *   (b) the committed golden vectors tests/golden/sdunes_*.npz generated from the reference by
 *       tests/golden/make_golden.py.
 *
 * The restatement always rebuilds and refactorises every block; the reference's
 * checkLastActiveSet caching (scenarios.c:369-482, 567-647) restores bit-identical copies of
 * unchanged blocks, so only the ORDER of the banded factorisation depends on that option
 * (reverseCholesky = checkLastActiveSet, scenarios.c:1922) and is kept here as `reverse`.
 *
 * Preconditions of the reference kept as is (scenarios.c:1498-1504, 1630): x0 eliminated
 * (nx[0] == 0), uniform nx = nx[1] and nu = nu[0], diagonal weights, bounds only, tree
 * branching md-fold for Nr stages followed by chains; all leaves at depth Nh.
 */
#include <stdlib.h>
#include <string.h>

#include "port_common.h"

typedef struct
{
    int Nn, Ns, Nh, Nr, md, nx, nu, Np;
    int *dad, *stage, *ao, *bo, *xo, *uo;
    int *nodeIdx, *removed, *cn, *lo;      /* [Ns*(Nh+1)] x2, [Ns-1] common nodes, [Ns] lambda offsets */
    double *Q, *R, *q, *r, *Qinv, *Rinv;   /* per node, un-scaled weights (scenarios.c:1885-1898) */
    /* per (scenario, stage) */
    double *x, *xUnc, *u, *uUnc, *Qc, *Rc, *res, *resMod, *mu, *dmu;
    double *Zbar, *LD, *CLD, *CLDi, *LL, *CLL;
    /* per scenario */
    double *Ut, *K, *fvals;
    /* per scenario pair */
    double *JD, *CJD, *CJDi, *JL, *CJL, *resNA, *rhsNA, *lam, *dlam;
    int *jdo, *jlo;
} sd_work;

#define NODE(w, i, k) ((w)->nodeIdx[(i) * ((w)->Nh + 1) + (k)])
#define REMOVED(w, i, k) ((w)->removed[(i) * ((w)->Nh + 1) + (k)])
#define VX(w, a, i, k) ((w)->a + ((size_t) (i) * (w)->Nh + (k)) * (w)->nx)
#define VU(w, a, i, k) ((w)->a + ((size_t) (i) * (w)->Nh + (k)) * (w)->nu)
#define MXX(w, a, i, k) ((w)->a + ((size_t) (i) * (w)->Nh + (k)) * (w)->nx * (w)->nx)
#define MXU(w, a, i, k) ((w)->a + ((size_t) (i) * (w)->Nh + (k)) * (w)->nx * (w)->nu)

static double *dz(size_t n) { return (double *) calloc(n + 1, sizeof(double)); }

/* scenarios.c:1607-1668 (dimensions, nodeIdx, boundsRemoved, commonNodes :262-291) */
static void sd_setup(const port_qp *qp, sd_work *w)
{
    const int Nn = qp->Nn;
    memset(w, 0, sizeof *w);
    w->Nn = Nn;
    w->nx = qp->nx[1];
    w->nu = qp->nu[0];
    w->md = qp->nk[0];
    w->dad = malloc(sizeof(int) * Nn); w->stage = malloc(sizeof(int) * Nn);
    w->ao = malloc(sizeof(int) * Nn); w->bo = malloc(sizeof(int) * Nn);
    w->xo = malloc(sizeof(int) * (Nn + 1)); w->uo = malloc(sizeof(int) * (Nn + 1));
    int next = 1, Np = 0;
    w->dad[0] = -1; w->stage[0] = 0; w->xo[0] = 0; w->uo[0] = 0;
    for (int i = 0; i < Nn; i++)
    {
        for (int j = 0; j < qp->nk[i]; j++) { w->dad[next + j] = i; w->stage[next + j] = w->stage[i] + 1; }
        next += qp->nk[i];
        if (qp->nk[i] > 0) Np++;
        w->xo[i + 1] = w->xo[i] + qp->nx[i];
        w->uo[i + 1] = w->uo[i] + qp->nu[i];
    }
    int ia = 0, ib = 0;
    for (int k = 0; k < Nn; k++)
    {
        w->ao[k] = ia; w->bo[k] = ib;
        if (k > 0) { ia += qp->nx[k] * qp->nx[w->dad[k]]; ib += qp->nx[k] * qp->nu[w->dad[k]]; }
    }
    w->Np = Np;
    w->Ns = Nn - Np;
    w->Nh = w->stage[Nn - 1];
    w->Nr = 0; /* utils/tree.c:65-81 */
    for (int k = 0; k < Nn; k++) { if (qp->nk[k] > 1) w->Nr = w->stage[k] + 1; else break; }
    const int Ns = w->Ns, Nh = w->Nh;

    w->nodeIdx = malloc(sizeof(int) * Ns * (Nh + 1));
    w->removed = malloc(sizeof(int) * Ns * (Nh + 1));
    char *seen = calloc(Nn, 1);
    for (int i = 0; i < Ns; i++)
    {
        int node = Nn - Ns + i;
        NODE(w, i, Nh) = node;
        REMOVED(w, i, Nh) = 0;
        for (int k = Nh; k > 0; k--)
        {
            node = w->dad[node];
            NODE(w, i, k - 1) = node;
            REMOVED(w, i, k - 1) = seen[node] ? 1 : 0;
            seen[node] = 1;
        }
    }
    free(seen);
    w->cn = malloc(sizeof(int) * (Ns > 1 ? Ns - 1 : 1));
    w->lo = malloc(sizeof(int) * (Ns + 1));
    w->jdo = malloc(sizeof(int) * (Ns + 1));
    w->jlo = malloc(sizeof(int) * (Ns + 1));
    int nl = 0, njd = 0, njl = 0;
    for (int i = 0; i < Ns - 1; i++)
    {
        int n1 = Nn - Ns + i, n2 = n1 + 1, c = -1;
        for (int k = Nh; k >= 0; k--)
        {
            if (n1 == n2) { c = k + 1; break; }
            n1 = w->dad[n1]; n2 = w->dad[n2];
        }
        w->cn[i] = c;
    }
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = w->nu * w->cn[i];
        w->lo[i] = nl; nl += dim;
        w->jdo[i] = njd; njd += dim * dim;
        w->jlo[i] = njl;
        if (i < Ns - 2) njl += w->nu * w->cn[i + 1] * dim;
    }
    w->lo[Ns - 1 > 0 ? Ns - 1 : 0] = nl;

    const int nx = w->nx, nu = w->nu, X = w->xo[Nn], U = w->uo[Nn];
    const size_t SK = (size_t) Ns * Nh;
    w->Q = dz(X); w->q = dz(X); w->Qinv = dz(X); w->R = dz(U); w->r = dz(U); w->Rinv = dz(U);
    w->x = dz(SK * nx); w->xUnc = dz(SK * nx); w->Qc = dz(SK * nx); w->res = dz(SK * nx);
    w->resMod = dz(SK * nx); w->mu = dz(SK * nx); w->dmu = dz(SK * nx);
    w->u = dz(SK * nu); w->uUnc = dz(SK * nu); w->Rc = dz(SK * nu);
    w->Zbar = dz(SK * nx * nu);
    w->LD = dz(SK * nx * nx); w->CLD = dz(SK * nx * nx); w->CLDi = dz(SK * nx);
    w->LL = dz(SK * nx * nx); w->CLL = dz(SK * nx * nx);
    const int nK = nu * w->Nr;
    w->Ut = dz((size_t) Ns * nK * Nh * nx); w->K = dz((size_t) Ns * nK * nK); w->fvals = dz(Ns);
    w->JD = dz(njd); w->CJD = dz(njd); w->CJDi = dz(nl); w->JL = dz(njl); w->CJL = dz(njl);
    w->resNA = dz(nl); w->rhsNA = dz(nl); w->lam = dz(nl); w->dlam = dz(nl);
}

static void sd_free(sd_work *w)
{
    free(w->dad); free(w->stage); free(w->ao); free(w->bo); free(w->xo); free(w->uo);
    free(w->nodeIdx); free(w->removed); free(w->cn); free(w->lo); free(w->jdo); free(w->jlo);
    free(w->Q); free(w->q); free(w->Qinv); free(w->R); free(w->r); free(w->Rinv);
    free(w->x); free(w->xUnc); free(w->Qc); free(w->res); free(w->resMod); free(w->mu); free(w->dmu);
    free(w->u); free(w->uUnc); free(w->Rc); free(w->Zbar); free(w->LD); free(w->CLD); free(w->CLDi);
    free(w->LL); free(w->CLL); free(w->Ut); free(w->K); free(w->fvals);
    free(w->JD); free(w->CJD); free(w->CJDi); free(w->JL); free(w->CJL);
    free(w->resNA); free(w->rhsNA); free(w->lam); free(w->dlam);
}

/* unconstrained stage solutions at the dual point (mu, lambda): the common part of
 * solve_stage_problems (scenarios.c:346-362, 384-400) and evaluate_dual_function (:1275-1307) */
static void sd_unconstrained(const port_qp *qp, const sd_work *w, int i, int k, double *xv, double *uv)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns;
    const int idxm1 = NODE(w, i, k), idx = NODE(w, i, k + 1);
    const double *mu = VX(w, mu, i, k);
    for (int a = 0; a < nx; a++)
    {
        double s = mu[a];
        if (k < Nh - 1)
        {   /* x[k+1] = mu[k+1] - A[k+1]' mu[k+2]   (dgemv_t) */
            const double *An = qp->A + w->ao[NODE(w, i, k + 2)], *mun = VX(w, mu, i, k + 1);
            double t = 0.0;
            for (int c = 0; c < nx; c++) t += An[c + a * nx] * mun[c];
            s = mu[a] - t;
        }
        s = s - w->q[w->xo[idx] + a];
        xv[a] = w->Qinv[w->xo[idx] + a] * s;
    }
    const double *Bk = qp->B + w->bo[idx];
    for (int a = 0; a < nu; a++)
    {   /* u[k] = -B[k]' mu[k] - r[k] -+ lambda   (dgemv_t with alpha=-1, beta=-1) */
        double t = 0.0;
        for (int c = 0; c < nx; c++) t += Bk[c + a * nx] * mu[c];
        double s = -w->r[w->uo[idxm1] + a] - t;
        if (i < Ns - 1 && k < w->cn[i]) s = s - w->lam[w->lo[i] + k * nu + a];
        if (i > 0 && k < w->cn[i - 1]) s = s + w->lam[w->lo[i - 1] + k * nu + a];
        uv[a] = w->Rinv[w->uo[idxm1] + a] * s;
    }
}

/* scenarios.c:311-486 */
static void sd_solve_stage_problems(const port_qp *qp, sd_work *w, int reverse)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns;
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {
            const int idxm1 = NODE(w, i, k), idx = NODE(w, i, k + 1);
            double *x = VX(w, x, i, k), *u = VU(w, u, i, k);
            sd_unconstrained(qp, w, i, k, x, u);
            if (REMOVED(w, i, k + 1) != 1)
                for (int a = 0; a < nx; a++)
                {
                    int m;
                    const int g = w->xo[idx] + a;
                    x[a] = clip(x[a], qp->xmin[g], qp->xmax[g], &m);
                    VX(w, Qc, i, k)[a] = m == 0 ? w->Qinv[g] : 0.0; /* dvecze */
                }
            if (REMOVED(w, i, k) != 1)
                for (int a = 0; a < nu; a++)
                {
                    int m;
                    const int g = w->uo[idxm1] + a;
                    u[a] = clip(u[a], qp->umin[g], qp->umax[g], &m);
                    VU(w, Rc, i, k)[a] = m == 0 ? w->Rinv[g] : 0.0;
                }
        }
    /* Zbar, LambdaD, LambdaL (:417-471); needs QinvCal of stage k-1, hence the second pass */
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {
            const int idx = NODE(w, i, k + 1);
            const double *Bk = qp->B + w->bo[idx], *Ak = qp->A + w->ao[idx];
            const double *Rc = VU(w, Rc, i, k), *Qc = VX(w, Qc, i, k);
            double *Z = MXU(w, Zbar, i, k), *LD = MXX(w, LD, i, k);
            for (int c = 0; c < nu; c++)
                for (int a = 0; a < nx; a++) Z[a + c * nx] = Bk[a + c * nx] * Rc[c];
            for (int b = 0; b < nx; b++)
                for (int a = 0; a < nx; a++)
                {   /* LambdaD = Zbar B' (dgemm_nt), + QinvCal on the diagonal (ddiaad) */
                    double s = 0.0;
                    for (int c = 0; c < nu; c++) s += Z[a + c * nx] * Bk[b + c * nx];
                    if (a == b) s += Qc[a];
                    LD[a + b * nx] = s;
                }
            if (k > 0)
            {
                const double *Qp = VX(w, Qc, i, k - 1);
                double *LL = MXX(w, LL, i, k - 1);
                if (reverse)
                {   /* LL = -QinvCal[k-1] A' (stored transposed); LD -= A LL   (:447-457) */
                    for (int b = 0; b < nx; b++)
                        for (int a = 0; a < nx; a++) LL[a + b * nx] = -(Qp[a] * Ak[b + a * nx]);
                    for (int b = 0; b < nx; b++)
                        for (int a = 0; a < nx; a++)
                        {
                            double s = 0.0;
                            for (int c = 0; c < nx; c++) s += Ak[a + c * nx] * LL[c + b * nx];
                            LD[a + b * nx] -= s;
                        }
                }
                else
                {   /* LL = -A QinvCal[k-1]; LD -= LL A'   (:462-469) */
                    for (int b = 0; b < nx; b++)
                        for (int a = 0; a < nx; a++) LL[a + b * nx] = -(Ak[a + b * nx] * Qp[b]);
                    for (int b = 0; b < nx; b++)
                        for (int a = 0; a < nx; a++)
                        {
                            double s = 0.0;
                            for (int c = 0; c < nx; c++) s += LL[a + c * nx] * Ak[b + c * nx];
                            LD[a + b * nx] -= s;
                        }
                }
            }
        }
}

/* scenarios.c:489-563 and :1452-1492 */
static double sd_residuals(const port_qp *qp, sd_work *w, int cond)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns;
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {
            const int idx = NODE(w, i, k + 1);
            const double *Bk = qp->B + w->bo[idx], *Ak = qp->A + w->ao[idx];
            const double *bk = qp->b + (w->xo[idx] - qp->nx[0]);
            const double *x = VX(w, x, i, k), *u = VU(w, u, i, k);
            double *res = VX(w, res, i, k);
            for (int a = 0; a < nx; a++)
            {
                double s;
                if (k == 0)
                {   /* res[1] = -b[0] - B[0] u[0] + x[1] */
                    double t = 0.0;
                    for (int c = 0; c < nu; c++) t += Bk[a + c * nx] * u[c];
                    s = -bk[a] - t;
                    s = s + x[a];
                }
                else
                {   /* res[k] = x[k] - A x[k-1] - B u[k-1] - b */
                    const double *xp = VX(w, x, i, k - 1);
                    double t = 0.0;
                    for (int c = 0; c < nx; c++) t += Ak[a + c * nx] * xp[c];
                    s = x[a] - t;
                    t = 0.0;
                    for (int c = 0; c < nu; c++) t += Bk[a + c * nx] * u[c];
                    s = s - t;
                    s = s - bk[a];
                }
                res[a] = s;
            }
        }
    /* non-anticipativity residual (:528-563): resNA[i] = u of scenario i+1 minus u of scenario i */
    for (int i = 0; i < Ns - 1; i++)
        for (int e = 0; e < nu * w->cn[i]; e++)
        {
            const int k = e / nu, a = e % nu;
            double s = 0.0;
            s += -1.0 * VU(w, u, i, k)[a];
            s += VU(w, u, i + 1, k)[a];
            w->resNA[w->lo[i] + e] = s;
        }
    double err = 0.0;
    const int nl = Ns > 1 ? w->lo[Ns - 1] : 0;
    if (cond == 2)
    {
        for (size_t g = 0; g < (size_t) Ns * Nh * nx; g++) { double a = fabs(w->res[g]); if (a > err) err = a; }
        for (int g = 0; g < nl; g++) { double a = fabs(w->resNA[g]); if (a > err) err = a; }
    }
    else
    {
        for (int i = 0; i < Ns; i++)
        {
            for (int k = 0; k < Nh; k++)
            {
                double s = 0.0;
                for (int a = 0; a < nx; a++) s += VX(w, res, i, k)[a] * VX(w, res, i, k)[a];
                err += s;
            }
            if (i < Ns - 1)
            {
                double s = 0.0;
                for (int e = 0; e < nu * w->cn[i]; e++) s += w->resNA[w->lo[i] + e] * w->resNA[w->lo[i] + e];
                err += s;
            }
        }
        if (cond == 1) err = sqrt(err);
    }
    return err;
}

/* D = B A^-T for lower-triangular A (n x n) with stored inverse diagonal; B, D are m x n,
 * leading dimensions ldb/ldd (blasfeo_dtrsm_rltn) */
static void trsm_rltn(int m, int n, const double *A, const double *Ai, const double *B, int ldb, double *D, int ldd)
{
    for (int j = 0; j < n; j++)
        for (int r = 0; r < m; r++)
        {
            double s = B[r + j * ldb];
            for (int t = 0; t < j; t++) s -= D[r + t * ldd] * A[j + t * n];
            D[r + j * ldd] = s * Ai[j];
        }
}
/* C(lower) -= A A' for A m x k (blasfeo_dsyrk_ln with alpha=-1, beta=1) */
static void syrk_ln_sub(int m, int k, const double *A, int lda, double *C, int ldc)
{
    for (int b = 0; b < m; b++)
        for (int a = b; a < m; a++)
        {
            double s = 0.0;
            for (int t = 0; t < k; t++) s += A[a + t * lda] * A[b + t * lda];
            C[a + b * ldc] -= s;
        }
}
/* z = L^-1 y (dtrsv_lnn) and z = L^-T y (dtrsv_ltn); z may alias y */
static void trsv_lnn(int n, const double *L, const double *Li, const double *y, double *z)
{
    for (int i = 0; i < n; i++)
    {
        double s = y[i];
        for (int t = 0; t < i; t++) s -= L[i + t * n] * z[t];
        z[i] = s * Li[i];
    }
}
static void trsv_ltn(int n, const double *L, const double *Li, const double *y, double *z)
{
    for (int i = n - 1; i >= 0; i--)
    {
        double s = y[i];
        for (int t = i + 1; t < n; t++) s -= L[t + i * n] * z[t];
        z[i] = s * Li[i];
    }
}

/* scenarios.c:590-689 */
static int sd_factorize_Lambda(sd_work *w, const port_opts *o, int reverse)
{
    const int nx = w->nx, Nh = w->Nh, Ns = w->Ns;
    int nreg = 0;
    for (int i = 0; i < Ns; i++)
    {
        if (reverse)
        {
            for (int k = Nh - 1; k > 0; k--)
            {
                nreg += port_potrf_reg(nx, MXX(w, LD, i, k), MXX(w, CLD, i, k), VX(w, CLDi, i, k), o);
                trsm_rltn(nx, nx, MXX(w, CLD, i, k), VX(w, CLDi, i, k), MXX(w, LL, i, k - 1), nx, MXX(w, CLL, i, k - 1), nx);
                syrk_ln_sub(nx, nx, MXX(w, CLL, i, k - 1), nx, MXX(w, LD, i, k - 1), nx);
            }
            nreg += port_potrf_reg(nx, MXX(w, LD, i, 0), MXX(w, CLD, i, 0), VX(w, CLDi, i, 0), o);
        }
        else
        {
            for (int k = 0; k < Nh - 1; k++)
            {
                nreg += port_potrf_reg(nx, MXX(w, LD, i, k), MXX(w, CLD, i, k), VX(w, CLDi, i, k), o);
                trsm_rltn(nx, nx, MXX(w, CLD, i, k), VX(w, CLDi, i, k), MXX(w, LL, i, k), nx, MXX(w, CLL, i, k), nx);
                syrk_ln_sub(nx, nx, MXX(w, CLL, i, k), nx, MXX(w, LD, i, k + 1), nx);
            }
            nreg += port_potrf_reg(nx, MXX(w, LD, i, Nh - 1), MXX(w, CLD, i, Nh - 1), VX(w, CLDi, i, Nh - 1), o);
        }
    }
    return nreg;
}

/* scenarios.c:692-792: Ut[i] (nu*Nr x Nh*nx) by block substitution, K[i] = -Ut Ut' + diag(RinvCal) */
static void sd_form_K(sd_work *w, int reverse)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns, Nr = w->Nr, nK = nu * Nr, ldu = nK;
    double *tmp = malloc(sizeof(double) * nu * nx), *tmp2 = malloc(sizeof(double) * nu * nx);
    for (int i = 0; i < Ns; i++)
    {
        double *Ut = w->Ut + (size_t) i * nK * Nh * nx, *K = w->K + (size_t) i * nK * nK;
        for (int j = 0; j < Nr; j++)
        {
            const double *Z = MXU(w, Zbar, i, j);
            for (int a = 0; a < nx; a++)
                for (int c = 0; c < nu; c++) tmp[c + a * nu] = Z[a + c * nx];
            int k = j;
            while (reverse ? k >= 0 : k < Nh)
            {
                double *blk = Ut + j * nu + (size_t) k * nx * ldu;
                trsm_rltn(nu, nx, MXX(w, CLD, i, k), VX(w, CLDi, i, k), tmp, nu, tmp2, nu);
                for (int a = 0; a < nx; a++)
                    for (int c = 0; c < nu; c++) blk[c + a * ldu] = tmp2[c + a * nu];
                const int more = reverse ? k > 0 : k < Nh - 1;
                if (more)
                {   /* tmp = -blk * CLL' (dgemm_nt) */
                    const double *CLL = MXX(w, CLL, i, reverse ? k - 1 : k);
                    for (int a = 0; a < nx; a++)
                        for (int c = 0; c < nu; c++)
                        {
                            double s = 0.0;
                            for (int t = 0; t < nx; t++) s += tmp2[c + t * nu] * CLL[a + t * nx];
                            tmp[c + a * nu] = -s;
                        }
                }
                k += reverse ? -1 : 1;
            }
        }
        for (int b = 0; b < nK; b++)
            for (int a = b; a < nK; a++)
            {
                double s = 0.0;
                for (int t = 0; t < Nh * nx; t++) s += Ut[a + (size_t) t * ldu] * Ut[b + (size_t) t * ldu];
                K[a + b * nK] = -s;
                K[b + a * nK] = -s; /* dtrtr_l mirror */
            }
        for (int k = 0; k < Nr; k++)
            for (int c = 0; c < nu; c++) K[(k * nu + c) * (nK + 1)] += VU(w, Rc, i, k)[c];
    }
    free(tmp); free(tmp2);
}

/* scenarios.c:795-868.  JD/JL are zero on entry except for the syrk update of the previous block */
static int sd_factorize_Jay(sd_work *w, const port_opts *o)
{
    const int nu = w->nu, Ns = w->Ns, nK = nu * w->Nr;
    int nreg = 0;
    if (Ns > 1) memset(w->JD, 0, sizeof(double) * (size_t) (w->jdo[Ns - 2] + nu * w->cn[Ns - 2] * nu * w->cn[Ns - 2]));
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * w->cn[i];
        double *JD = w->JD + w->jdo[i], *CJD = w->CJD + w->jdo[i], *CJDi = w->CJDi + w->lo[i];
        const double *K0 = w->K + (size_t) i * nK * nK, *K1 = w->K + (size_t) (i + 1) * nK * nK;
        for (int b = 0; b < dim; b++)
            for (int a = 0; a < dim; a++)
            {
                JD[a + b * dim] += K0[a + b * nK];
                JD[a + b * dim] += K1[a + b * nK];
            }
        nreg += port_potrf_reg(dim, JD, CJD, CJDi, o);
        if (i < Ns - 2)
        {
            const int dimN = nu * w->cn[i + 1];
            double *JL = w->JL + w->jlo[i], *CJL = w->CJL + w->jlo[i];
            for (int b = 0; b < dim; b++)
                for (int a = 0; a < dimN; a++) JL[a + b * dimN] = 0.0 + -1.0 * K1[a + b * nK];
            trsm_rltn(dimN, dim, CJD, CJDi, JL, dimN, CJL, dimN);
            syrk_ln_sub(dimN, dim, CJL, dimN, w->JD + w->jdo[i + 1], dimN);
        }
    }
    return nreg;
}

/* banded forward/backward substitution with the factor of Lambda_i: out = Lambda_i^-1 rhs,
 * rhs (Nh x nx) is overwritten (scenarios.c:893-958 and :1129-1188 share this shape) */
static void sd_banded_solve(sd_work *w, int i, int reverse, double *rhs, double *out)
{
    const int nx = w->nx, Nh = w->Nh;
#define BLK(v, k) ((v) + (size_t) (k) * nx)
    if (reverse)
    {
        for (int k = Nh; k > 1; k--)
        {
            trsv_lnn(nx, MXX(w, CLD, i, k - 1), VX(w, CLDi, i, k - 1), BLK(rhs, k - 1), BLK(out, k - 1));
            const double *CLL = MXX(w, CLL, i, k - 2);
            for (int a = 0; a < nx; a++)
            {   /* rhs[k-2] -= CLL[k-2] out[k-1]   (dgemv_n) */
                double s = 0.0;
                for (int c = 0; c < nx; c++) s += CLL[a + c * nx] * BLK(out, k - 1)[c];
                BLK(rhs, k - 2)[a] = BLK(rhs, k - 2)[a] - s;
            }
        }
        trsv_lnn(nx, MXX(w, CLD, i, 0), VX(w, CLDi, i, 0), BLK(rhs, 0), BLK(out, 0));
        for (int k = 0; k < Nh - 1; k++)
        {
            trsv_ltn(nx, MXX(w, CLD, i, k), VX(w, CLDi, i, k), BLK(out, k), BLK(out, k));
            const double *CLL = MXX(w, CLL, i, k);
            for (int a = 0; a < nx; a++)
            {   /* out[k+1] -= CLL[k]' out[k]   (dgemv_t) */
                double s = 0.0;
                for (int c = 0; c < nx; c++) s += CLL[c + a * nx] * BLK(out, k)[c];
                BLK(out, k + 1)[a] = BLK(out, k + 1)[a] - s;
            }
        }
        trsv_ltn(nx, MXX(w, CLD, i, Nh - 1), VX(w, CLDi, i, Nh - 1), BLK(out, Nh - 1), BLK(out, Nh - 1));
    }
    else
    {
        for (int k = 0; k < Nh - 1; k++)
        {
            trsv_lnn(nx, MXX(w, CLD, i, k), VX(w, CLDi, i, k), BLK(rhs, k), BLK(out, k));
            const double *CLL = MXX(w, CLL, i, k);
            for (int a = 0; a < nx; a++)
            {
                double s = 0.0;
                for (int c = 0; c < nx; c++) s += CLL[a + c * nx] * BLK(out, k)[c];
                BLK(rhs, k + 1)[a] = BLK(rhs, k + 1)[a] - s;
            }
        }
        trsv_lnn(nx, MXX(w, CLD, i, Nh - 1), VX(w, CLDi, i, Nh - 1), BLK(rhs, Nh - 1), BLK(out, Nh - 1));
        for (int k = Nh; k > 1; k--)
        {
            trsv_ltn(nx, MXX(w, CLD, i, k - 1), VX(w, CLDi, i, k - 1), BLK(out, k - 1), BLK(out, k - 1));
            const double *CLL = MXX(w, CLL, i, k - 2);
            for (int a = 0; a < nx; a++)
            {
                double s = 0.0;
                for (int c = 0; c < nx; c++) s += CLL[c + a * nx] * BLK(out, k - 1)[c];
                BLK(out, k - 2)[a] = BLK(out, k - 2)[a] - s;
            }
        }
        trsv_ltn(nx, MXX(w, CLD, i, 0), VX(w, CLDi, i, 0), BLK(out, 0), BLK(out, 0));
    }
#undef BLK
}

/* scenarios.c:871-1000 */
static void sd_form_rhs(sd_work *w, int reverse, double *scratch)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns;
    for (int i = 0; i < Ns; i++)
    {   /* resMod[i] = Lambda_i^-1 res[i]; the reference keeps res intact through a temporary */
        memcpy(scratch, VX(w, res, i, 0), sizeof(double) * Nh * nx);
        sd_banded_solve(w, i, reverse, scratch, VX(w, resMod, i, 0));
    }
    for (int i = 0; i < Ns - 1; i++)
        for (int e = 0; e < nu * w->cn[i]; e++)
        {
            const int k = e / nu, c = e % nu;
            double s = w->resNA[w->lo[i] + e], t = 0.0;
            const double *Z = MXU(w, Zbar, i, k), *rm = VX(w, resMod, i, k);
            for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
            s = s - t;
            Z = MXU(w, Zbar, i + 1, k); rm = VX(w, resMod, i + 1, k);
            t = 0.0;
            for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
            s = s + t;
            w->rhsNA[w->lo[i] + e] = s;
        }
}

/* scenarios.c:1004-1074 */
static void sd_delta_lambda(sd_work *w)
{
    const int nu = w->nu, Ns = w->Ns;
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * w->cn[i];
        double *rhs = w->rhsNA + w->lo[i], *dl = w->dlam + w->lo[i];
        for (int e = 0; e < dim; e++) dl[e] = -1.0 * rhs[e];
        trsv_lnn(dim, w->CJD + w->jdo[i], w->CJDi + w->lo[i], dl, rhs);
        if (i < Ns - 2)
        {
            const int dimN = nu * w->cn[i + 1];
            const double *CJL = w->CJL + w->jlo[i];
            double *rn = w->rhsNA + w->lo[i + 1];
            for (int a = 0; a < dimN; a++)
            {
                double s = 0.0;
                for (int c = 0; c < dim; c++) s += CJL[a + c * dimN] * rhs[c];
                rn[a] = rn[a] + s;
            }
        }
    }
    for (int i = Ns - 2; i >= 0; i--)
    {
        const int dim = nu * w->cn[i];
        double *rhs = w->rhsNA + w->lo[i], *dl = w->dlam + w->lo[i];
        trsv_ltn(dim, w->CJD + w->jdo[i], w->CJDi + w->lo[i], rhs, dl);
        if (i > 0)
        {
            const int dimP = nu * w->cn[i - 1];
            const double *CJL = w->CJL + w->jlo[i - 1]; /* dim x dimP */
            double *rp = w->rhsNA + w->lo[i - 1];
            for (int c = 0; c < dimP; c++)
            {
                double s = 0.0;
                for (int a = 0; a < dim; a++) s += CJL[a + c * dim] * dl[a];
                rp[c] = rp[c] - s;
            }
        }
    }
}

/* scenarios.c:1077-1205 */
static void sd_delta_mu(sd_work *w, int reverse)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns, Nr = w->Nr;
    double v[64];
    for (int i = 0; i < Ns; i++)
    {
        for (int k = 0; k < Nh; k++)
        {
            double *rm = VX(w, resMod, i, k);
            const double *res = VX(w, res, i, k);
            for (int c = 0; c < nu; c++)
            {
                double s = 0.0;
                if (i > 0 && k < w->cn[i - 1]) s = s + w->dlam[w->lo[i - 1] + k * nu + c];
                if (i < Ns - 1 && k < w->cn[i]) s = s - w->dlam[w->lo[i] + k * nu + c];
                v[c] = s;
            }
            if (k < Nr)
            {
                const double *Z = MXU(w, Zbar, i, k);
                for (int a = 0; a < nx; a++)
                {
                    double s = 0.0;
                    for (int c = 0; c < nu; c++) s += Z[a + c * nx] * v[c];
                    rm[a] = -res[a] + s;
                }
            }
            else
                for (int a = 0; a < nx; a++) rm[a] = 0.0 - res[a];
        }
        sd_banded_solve(w, i, reverse, VX(w, resMod, i, 0), VX(w, dmu, i, 0));
    }
}

/* scenarios.c:1229-1382 (NEW_FVAL branch) */
static double sd_evaluate_dual_function(const port_qp *qp, sd_work *w)
{
    const int nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns;
    double fval = 0.0;
    for (int i = 0; i < Ns; i++)
    {
        double f = 0.0;
        for (int k = 0; k < Nh; k++)
        {
            const int idxm1 = NODE(w, i, k), idx = NODE(w, i, k + 1);
            double *x = VX(w, x, i, k), *u = VU(w, u, i, k), *xU = VX(w, xUnc, i, k), *uU = VU(w, uUnc, i, k);
            sd_unconstrained(qp, w, i, k, xU, uU);
            int m;
            double s = 0.0;
            for (int a = 0; a < nx; a++)
            {
                const int g = w->xo[idx] + a;
                x[a] = REMOVED(w, i, k + 1) != 1 ? clip(xU[a], qp->xmin[g], qp->xmax[g], &m) : xU[a];
                xU[a] = w->Q[g] * (x[a] + -2.0 * xU[a]);
                s += x[a] * xU[a];
            }
            f -= 0.5 * s;
            s = 0.0;
            for (int a = 0; a < nu; a++)
            {
                const int g = w->uo[idxm1] + a;
                u[a] = REMOVED(w, i, k) != 1 ? clip(uU[a], qp->umin[g], qp->umax[g], &m) : uU[a];
                uU[a] = w->R[g] * (u[a] + -2.0 * uU[a]);
                s += u[a] * uU[a];
            }
            f -= 0.5 * s;
            s = 0.0;
            const double *bk = qp->b + (w->xo[idx] - qp->nx[0]), *mu = VX(w, mu, i, k);
            for (int a = 0; a < nx; a++) s += bk[a] * mu[a];
            f -= s;
        }
        w->fvals[i] = f;
    }
    for (int i = 0; i < Ns; i++) fval += w->fvals[i];
    return fval;
}

/* Entry point: treeqp_sdunes_solve (scenarios.c:1850-2091).
 *   lam0 [nl], mu0 [Ns*Nh*nx]: dual start (treeqp_sdunes_set_dual_initialization :2094-2117)
 *   x,u,mu_x,mu_u per node, lam_edge per edge: what the reference writes into tree_qp_out
 *   lam_out [nl], mu_out [Ns*Nh*nx]: the scenario duals at exit (work->slambda / work->smu)
 * Returns 0 (optimal) or 1 (maximum iterations). */
int oracle_sdunes_solve(const port_qp *qp, const port_opts *o, int checkLastActiveSet,
                        const double *lam0, const double *mu0,
                        double *x, double *u, double *lam_edge, double *mu_x, double *mu_u,
                        double *lam_out, double *mu_out, int *iter_out, int *ls_iters, int *nreg_out)
{
    sd_work W, *w = &W;
    sd_setup(qp, w);
    const int Nn = qp->Nn, nx = w->nx, nu = w->nu, Nh = w->Nh, Ns = w->Ns, Nr = w->Nr, md = w->md;
    const int nl = Ns > 1 ? w->lo[Ns - 1] : 0, reverse = checkLastActiveSet;
    int status = 11, it, nreg = 0;
    double *scratch = malloc(sizeof(double) * (size_t) Nh * nx + 8);

    /* un-scale the weights (:1885-1898): inverse of the factor of tree_qp_in_fill_lti_data */
    int mdNr = 1;
    for (int t = 0; t < Nr; t++) mdNr *= md;
    for (int k = 0; k < Nn; k++)
    {
        int st = w->stage[k] < Nr ? w->stage[k] : Nr, p = 1;
        for (int t = 0; t < st; t++) p *= md;
        const double sc = (double) p / mdNr;
        for (int a = 0; a < qp->nx[k]; a++)
        {
            const int g = w->xo[k] + a;
            w->Q[g] = sc * qp->Qd[g]; w->q[g] = sc * qp->q[g]; w->Qinv[g] = 1.0 / w->Q[g];
        }
        for (int a = 0; a < qp->nu[k]; a++)
        {
            const int g = w->uo[k] + a;
            w->R[g] = sc * qp->Rd[g]; w->r[g] = sc * qp->r[g]; w->Rinv[g] = 1.0 / w->R[g];
        }
    }
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {   /* :1901-1917: defaults for nodes whose bounds were removed */
            memcpy(VX(w, Qc, i, k), w->Qinv + w->xo[NODE(w, i, k + 1)], sizeof(double) * nx);
            memcpy(VU(w, Rc, i, k), w->Rinv + w->uo[NODE(w, i, k)], sizeof(double) * nu);
        }
    memcpy(w->lam, lam0, sizeof(double) * nl);
    memcpy(w->mu, mu0, sizeof(double) * (size_t) Ns * Nh * nx);

    for (it = 0; it < o->maxIter; it++)
    {
        sd_solve_stage_problems(qp, w, reverse);
        const double err = sd_residuals(qp, w, o->termCondition);
        if (err < o->tol) { status = 0; break; }
        nreg += sd_factorize_Lambda(w, o, reverse);
        sd_form_K(w, reverse);
        nreg += sd_factorize_Jay(w, o);
        sd_form_rhs(w, reverse, scratch);
        sd_delta_lambda(w);
        sd_delta_mu(w, reverse);

        /* line search (:1385-1448); no descent test in SDUNES */
        double dot = 0.0;
        for (int i = 0; i < Ns - 1; i++)
        {
            double s = 0.0;
            for (int e = 0; e < nu * w->cn[i]; e++) s += w->resNA[w->lo[i] + e] * w->dlam[w->lo[i] + e];
            dot += s;
        }
        for (int i = 0; i < Ns; i++)
            for (int k = 0; k < Nh; k++)
            {
                double s = 0.0;
                for (int a = 0; a < nx; a++) s += VX(w, res, i, k)[a] * VX(w, dmu, i, k)[a];
                dot += s;
            }
        const double f0 = sd_evaluate_dual_function(qp, w);
        double tau = 1.0, tauPrev = 0.0;
        int ls;
        for (ls = 1; ls <= o->lsMaxIter; ls++)
        {
            for (size_t g = 0; g < (size_t) Ns * Nh * nx; g++) w->mu[g] += (tau - tauPrev) * w->dmu[g];
            for (int g = 0; g < nl; g++) w->lam[g] += (tau - tauPrev) * w->dlam[g];
            const double f = sd_evaluate_dual_function(qp, w);
            if (f <= f0 + o->gamma * tau * dot) break;
            tauPrev = tau;
            tau = o->beta * tauPrev;
        }
        if (ls_iters) ls_iters[it] = ls;
    }
    if (it == o->maxIter) status = 1;

    /* copy-out (:2026-2075) */
    const int X = w->xo[Nn], U = w->uo[Nn];
    memset(lam_edge, 0, sizeof(double) * (X - qp->nx[0]));
    for (int i = 0; i < Ns; i++)
        for (int k = 0; k < Nh; k++)
        {
            const int idx = NODE(w, i, k + 1), idxm1 = NODE(w, i, k);
            for (int a = 0; a < nx; a++) lam_edge[w->xo[idx] - qp->nx[0] + a] += VX(w, mu, i, k)[a];
            double xU[64], uU[64];
            sd_unconstrained(qp, w, i, k, xU, uU);
            if (REMOVED(w, i, k + 1) == 0)
                for (int a = 0; a < nx; a++)
                {
                    const int g = w->xo[idx] + a;
                    x[g] = VX(w, x, i, k)[a];
                    mu_x[g] = w->Q[g] * (xU[a] - x[g]);
                }
            if (REMOVED(w, i, k) == 0)
                for (int a = 0; a < nu; a++)
                {
                    const int g = w->uo[idxm1] + a;
                    u[g] = VU(w, u, i, k)[a];
                    mu_u[g] = w->R[g] * (uU[a] - u[g]);
                }
        }
    if (lam_out) memcpy(lam_out, w->lam, sizeof(double) * nl);
    if (mu_out) memcpy(mu_out, w->mu, sizeof(double) * (size_t) Ns * Nh * nx);
    *iter_out = it;
    if (nreg_out) *nreg_out = nreg;
    free(scratch);
    sd_free(w);
    return status;
}

/* dimensions a caller needs to size the dual vectors (scenarios.c:99-107 and work->Ns/Nh) */
void oracle_sdunes_dims(const port_qp *qp, int *Ns, int *Nh, int *Nr, int *nl)
{
    sd_work W;
    sd_setup(qp, &W);
    *Ns = W.Ns; *Nh = W.Nh; *Nr = W.Nr;
    *nl = W.Ns > 1 ? W.lo[W.Ns - 1] : 0;
    sd_free(&W);
}
