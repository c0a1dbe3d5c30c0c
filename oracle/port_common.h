/*
 * oracle/port_common.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Types and small dense helpers shared by the plain-C restatements oracle/tdunes_port.c and
 * oracle/sdunes_port.c (see the headers of those files for the parity status).
 */
#ifndef ORACLE_PORT_COMMON_H_
#define ORACLE_PORT_COMMON_H_
#include <math.h>

typedef struct
{
    int Nn;
    const int *nk;      /* children per node, breadth-first */
    const int *nx;      /* states per node */
    const int *nu;      /* inputs per node */
    const double *A;    /* per edge e=k-1: nx[k] x nx[dad] column-major, concatenated */
    const double *B;    /* per edge: nx[k] x nu[dad] */
    const double *b;    /* per edge: nx[k] */
    const double *Qd;   /* per node diag(Q), concatenated */
    const double *Rd;   /* per node diag(R) */
    const double *q, *r;
    const double *xmin, *xmax, *umin, *umax;
    /* dense stage-QP path (dual_Newton_tree_qpoases.c); both NULL = clipping everywhere */
    const int *solver;  /* per node: 0 clipping, 1 dense (TREEQP_QPOASES_SOLVER) */
    const double *H;    /* per node [Q S'; S R], (nx+nu)^2 column-major, concatenated over ALL nodes */
} port_qp;

typedef struct
{
    int maxIter;
    int termCondition; /* 0 sum of squares, 1 two-norm, 2 inf-norm (utils/types.h:46-50) */
    double tol;
    int regType; /* 0 none, 1 always, 2 on the fly (dual_Newton_common.h:41-46) */
    double regTol, regValue;
    int lsMaxIter;
    double gamma, beta;
} port_opts;

/* blasfeo_dveccl_mask, external/blasfeo/auxiliary/d_aux_lib4.c:4385-4432 (AVX branch: the
 * comparisons are inclusive and the lower bound wins when both fire) */
static inline double clip(double v, double lo, double hi, int *mask)
{
    int up = hi <= v, dn = v <= lo;
    *mask = dn ? -1 : (up ? 1 : 0);
    return dn ? lo : (up ? hi : v);
}

/* blasfeo_dpotrf_l with the pivot rule of external/blasfeo/kernel/c99/kernel_dgemm_4x4_lib4.c:
 * 4633-4646: pivot>0 ? (l=sqrt, inv=1/l) : (l=0, inv=0); sub-diagonal entries are scaled by inv */
static inline void port_potrf(int d, const double *Wm, double *L, double *Linv)
{
    for (int j = 0; j < d; j++)
    {
        for (int i = j; i < d; i++)
        {
            double s = Wm[i + j * d];
            for (int t = 0; t < j; t++) s -= L[i + t * d] * L[j + t * d];
            L[i + j * d] = s;
        }
        double piv = L[j + j * d], inv;
        if (piv > 0.0) { piv = sqrt(piv); inv = 1.0 / piv; }
        else { piv = 0.0; inv = 0.0; }
        L[j + j * d] = piv;
        Linv[j] = inv;
        for (int i = j + 1; i < d; i++) L[i + j * d] *= inv;
    }
}

/* dual_Newton_common.c:35-78 */
static inline int port_potrf_reg(int d, double *Wm, double *L, double *Linv, const port_opts *o)
{
    if (o->regType == 1)
        for (int j = 0; j < d; j++) Wm[j + j * d] += o->regValue;
    port_potrf(d, Wm, L, Linv);
    if (o->regType == 2)
        for (int j = 0; j < d; j++)
            if (L[j + j * d] <= o->regTol)
            {
                for (int t = 0; t < d; t++) Wm[t + t * d] += o->regValue;
                port_potrf(d, Wm, L, Linv);
                return 1;
            }
    return o->regType == 1;
}


/* ---------------------------------------------------------------------------------------------
 * Dense stage QP  min 1/2 z'Hz - g'z  s.t. lb <= z <= ub  (H symmetric positive definite).
 * The reference delegates this to qpOASES (QProblemB_init / QProblemB_hotstart,
 * dual_Newton_tree_qpoases.c:218-357; external/qpoases/src/QProblemB.c) and then reads three
 * things back: the minimiser x, the multipliers y and the set of free variables with the
 * Cholesky factor R of the reduced Hessian (QProblem_build_elimination_matrix, :153-214).  For a
 * strictly convex QP all three are properties of the (unique) solution, not of qpOASES' homotopy
 * path, so the restatement is a textbook primal active-set method (Nocedal & Wright alg. 16.3
 * specialised to bounds) that returns exactly those quantities.  Pinned against the reference's
 * outputs in tests/test_oracle.py (dense cases).
 *
 * ws[i] in {-1 at lower, 0 free, +1 at upper} is the warm start on entry and the final working
 * set on exit.  Work: L (n*n), t (3n).  Returns 0, or 1 if the iteration cap is hit.
 * ------------------------------------------------------------------------------------------- */
static inline int port_boxqp(int n, const double *H, const double *g, const double *lb, const double *ub,
                             int *ws, double *z, double *mu, double *L, double *t)
{
    double *Ht = t + 3 * n; /* caller provides 3n + n*n doubles in t */
    double *rhs = t, *inv = t + n, *zt = t + 2 * n;
    for (int i = 0; i < n; i++)
    {
        if (lb[i] >= ub[i]) ws[i] = -1; /* pinned variable: always fixed (lower wins) */
        if (ws[i] < 0) z[i] = lb[i];
        else if (ws[i] > 0) z[i] = ub[i];
        else z[i] = 0.0 < lb[i] ? lb[i] : (0.0 > ub[i] ? ub[i] : 0.0);
    }
    for (int iter = 0; iter < 10 * n + 20; iter++)
    {
        /* equality-constrained step on the free variables: fixed rows/columns -> identity */
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++)
                Ht[i + j * n] = (ws[i] == 0 && ws[j] == 0) ? H[i + j * n] : (i == j ? 1.0 : 0.0);
        for (int i = 0; i < n; i++)
        {
            if (ws[i] != 0) { rhs[i] = z[i]; continue; }
            double s = g[i];
            for (int a = 0; a < n; a++)
                if (ws[a] != 0) s -= H[i + a * n] * z[a];
            rhs[i] = s;
        }
        port_potrf(n, Ht, L, inv);
        for (int i = 0; i < n; i++)
        {
            double s = rhs[i];
            for (int k = 0; k < i; k++) s -= L[i + k * n] * zt[k];
            zt[i] = s * inv[i];
        }
        for (int i = n - 1; i >= 0; i--)
        {
            double s = zt[i];
            for (int k = i + 1; k < n; k++) s -= L[k + i * n] * zt[k];
            zt[i] = s * inv[i];
        }
        /* ratio test: longest feasible step towards the subspace minimiser */
        double alpha = 1.0;
        int block = -1, side = 0;
        for (int i = 0; i < n; i++)
        {
            if (ws[i] != 0) continue;
            double p = zt[i] - z[i];
            if (p < 0.0 && zt[i] < lb[i])
            {
                double a = (lb[i] - z[i]) / p;
                if (a < alpha) { alpha = a; block = i; side = -1; }
            }
            else if (p > 0.0 && zt[i] > ub[i])
            {
                double a = (ub[i] - z[i]) / p;
                if (a < alpha) { alpha = a; block = i; side = 1; }
            }
        }
        if (block >= 0)
        {
            for (int i = 0; i < n; i++)
                if (ws[i] == 0) z[i] += alpha * (zt[i] - z[i]);
            z[block] = side < 0 ? lb[block] : ub[block];
            ws[block] = side;
            continue;
        }
        for (int i = 0; i < n; i++)
            if (ws[i] == 0) z[i] = zt[i];
        /* multipliers of the fixed variables: mu = g - Hz (> 0 at an upper, < 0 at a lower bound,
         * the sign convention of tree_qp_out, tree_qp_common.h:129-131) */
        int worst = -1;
        double worst_v = 0.0;
        for (int i = 0; i < n; i++)
        {
            mu[i] = 0.0;
            if (ws[i] == 0) continue;
            double s = g[i];
            for (int a = 0; a < n; a++) s -= H[i + a * n] * z[a];
            mu[i] = s;
            if (lb[i] >= ub[i]) continue;
            double viol = ws[i] < 0 ? s : -s; /* wrong sign = the bound should be released */
            if (viol > 1e-12 * (1.0 + fabs(g[i])) && viol > worst_v) { worst_v = viol; worst = i; }
        }
        if (worst < 0) return 0;
        ws[worst] = 0;
    }
    return 1;
}

/* P = Z (Z'HZ)^-1 Z' with Z the columns of the identity of the free variables
 * (QProblem_build_elimination_matrix, dual_Newton_tree_qpoases.c:153-214): the inverse of the free
 * block of H embedded in zeros.  Work t: 2*n*n + n doubles. */
static inline void port_boxqp_projector(int n, const double *H, const int *ws, double *P, double *t)
{
    double *Ht = t, *L = t + n * n, *inv = t + 2 * n * n;
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
            Ht[i + j * n] = (ws[i] == 0 && ws[j] == 0) ? H[i + j * n] : (i == j ? 1.0 : 0.0);
    port_potrf(n, Ht, L, inv);
    /* Y = L^-1 (overwrites Ht), column by column */
    for (int c = 0; c < n; c++)
        for (int i = 0; i < n; i++)
        {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= L[i + k * n] * Ht[k + c * n];
            Ht[i + c * n] = s * inv[i];
        }
    for (int b = 0; b < n; b++)
        for (int a = 0; a < n; a++)
        {
            double s = 0.0;
            for (int i = 0; i < n; i++) s += Ht[i + a * n] * Ht[i + b * n];
            P[a + b * n] = (ws[a] == 0 && ws[b] == 0) ? s : 0.0;
        }
}

#endif
