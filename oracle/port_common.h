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

#endif
