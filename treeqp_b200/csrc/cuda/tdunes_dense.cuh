/*
 * tdunes_dense.cuh -- dense stage QPs (the reference's qpOASES plug-in, dual_Newton_tree_qpoases.c)
 * on the GPU: one warp per node, one lane per variable of z = [x; u] (nx + nu <= 32).
 *
 *   min 1/2 z'Hz - g'z   s.t.  lb <= z <= ub,      H = [Q S'; S R] > 0,  g = [qmod; rmod]
 *
 * The reference hands this to qpOASES_e (QProblemB_init / QProblemB_hotstart, qpoases.c:218-357) and
 * reads back the minimiser, the bound multipliers and the free-variable set with the Cholesky
 * factor of the reduced Hessian (QProblem_build_elimination_matrix, :153-214).  All three are
 * properties of the unique solution of a strictly convex QP, so the kernel uses a primal
 * active-set method on the warp (working set warm-started from the previous Newton iteration,
 * every step one in-shared-memory Cholesky of the free block embedded in the identity) and
 * returns the same quantities:
 *   warp_boxqp            z, mu = g - Hz on the fixed variables (sign convention of tree_qp_out),
 *                         working set ws in {-1, 0, +1}
 *   warp_boxqp_projector  P = Z (Z'HZ)^-1 Z'  (inverse of the free block of H embedded in zeros)
 *   k_trial_dense         line-search trial of the dense nodes: solve at g0 + tau*dg, dual term
 *                         -1/2 x'Qx + x'qmod - 1/2 u'Ru + u'rmod (qpoases.c:478-523: no u'Sx term,
 *                         reproduced as is)
 *   k_finish_dense        mu_x, mu_u of the dense nodes (qpoases.c:527-565)
 * The Hessian blocks with a dense P are formed in k_hessian (tdunes_kernels.cuh).
 */
#pragma once
#include "warp_helpers.cuh"

#define TQB_DENSE_OK 0
#define TQB_DENSE_FAIL 1

/* shared memory per warp (doubles) for the dense solver: Hs | sL | sY (nv x ld each) | sinv (nv) */
__host__ __device__ inline int tqb_dense_ld(int nvmax) { return nvmax | 1; }
__host__ __device__ inline int tqb_dense_smem(int nvmax) { return 3 * nvmax * tqb_dense_ld(nvmax) + nvmax + 1; }

/* lane i (< n) owns variable i.  On entry ws is the warm start, on exit the final working set;
 * sL/sinv hold the Cholesky factor of the free block (embedded in the identity) of that set. */
__device__ __forceinline__ int warp_boxqp(int n, const double *Hs, int ld, double g, double lb, double ub,
                                          int &ws, double &z, double &mu, double *sL, double *sinv, int lane)
{
    const bool act = lane < n;
    const bool pinned = act && lb >= ub;
    if (!act) { ws = 1; }
    if (pinned) ws = -1;
    if (ws < 0) z = lb;
    else if (ws > 0) z = ub;
    else z = 0.0 < lb ? lb : (0.0 > ub ? ub : 0.0);
    if (!act) z = 0.0;
    const int itmax = 10 * n + 20;
    for (int iter = 0; iter < itmax; iter++)
    {
        const unsigned fm = __ballot_sync(FULL, ws != 0); /* fixed variables (lanes >= n count as fixed) */
        __syncwarp();
        for (int e = lane; e < n * n; e += 32)
        {
            const int i = e % n, j = e / n;
            const bool fr = !((fm >> i) & 1u) && !((fm >> j) & 1u);
            sL[i + j * ld] = fr ? Hs[i + j * ld] : (i == j ? 1.0 : 0.0);
        }
        double rhs = act ? (ws != 0 ? z : g) : 0.0;
        for (int a = 0; a < n; a++)
        {
            const double za = __shfl_sync(FULL, z, a);
            if (act && ws == 0 && ((fm >> a) & 1u)) rhs -= Hs[lane + a * ld] * za;
        }
        __syncwarp();
        warp_potrf32(sL, ld, n, sinv, lane);
        const double inv = act ? sinv[lane] : 0.0;
        double zt = warp_trsv_lnn32(sL, ld, n, inv, rhs, lane);
        zt = warp_trsv_ltn32(sL, ld, n, inv, zt, lane);
        /* ratio test towards the subspace minimiser */
        const double p = zt - z;
        double a = 2.0;
        int side = 0;
        if (act && ws == 0)
        {
            if (p < 0.0 && zt < lb) { a = (lb - z) / p; side = -1; }
            else if (p > 0.0 && zt > ub) { a = (ub - z) / p; side = 1; }
        }
        double amin = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amin = fmin(amin, __shfl_xor_sync(FULL, amin, o));
        if (amin < 1.0)
        {
            const unsigned hit = __ballot_sync(FULL, a == amin);
            const int block = __ffs(hit) - 1;
            if (act && ws == 0) z += amin * p;
            if (lane == block) { z = side < 0 ? lb : ub; ws = side; }
            continue;
        }
        if (act && ws == 0) z = zt;
        /* multipliers of the fixed variables, most violated sign is released */
        double s = g;
        for (int c = 0; c < n; c++)
        {
            const double zc = __shfl_sync(FULL, z, c);
            if (act) s -= Hs[lane + c * ld] * zc;
        }
        mu = (act && ws != 0) ? s : 0.0;
        double viol = 0.0;
        if (act && ws != 0 && !pinned)
        {
            const double v = ws < 0 ? s : -s;
            if (v > 1e-12 * (1.0 + fabs(g))) viol = v;
        }
        const double vmax = warp_max(viol);
        if (!(vmax > 0.0)) return TQB_DENSE_OK;
        const unsigned hit = __ballot_sync(FULL, viol == vmax);
        if (lane == __ffs(hit) - 1) ws = 0;
    }
    return TQB_DENSE_FAIL;
}

/* P (n x n column-major, global) from the factor left in sL/sinv by warp_boxqp: Y = L^-1 by
 * columns (lane c owns column c), P = Y'Y with the rows/columns of fixed variables zeroed. */
__device__ __forceinline__ void warp_boxqp_projector(int n, int ld, int ws, const double *sL, const double *sinv,
                                                     double *sY, double *Pg, int lane)
{
    const unsigned fm = __ballot_sync(FULL, ws != 0);
    __syncwarp();
    if (lane < n)
    {
        const int c = lane;
        for (int i = 0; i < c; i++) sY[i + c * ld] = 0.0;
        for (int i = c; i < n; i++)
        {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= sL[i + k * ld] * sY[k + c * ld];
            sY[i + c * ld] = s * sinv[i];
        }
    }
    __syncwarp();
    if (lane < n)
    {
        const int b = lane;
        const bool fb = (fm >> b) & 1u;
        for (int a = 0; a < n; a++)
        {
            double s = 0.0;
            const int i0 = a > b ? a : b;
            for (int i = i0; i < n; i++) s += sY[i + a * ld] * sY[i + b * ld];
            Pg[a + b * n] = (fb || ((fm >> a) & 1u)) ? 0.0 : s;
        }
    }
    __syncwarp();
}

/* stage H_k into shared memory (n x n, leading dimension ld) */
__device__ __forceinline__ void warp_load_H(int n, int ld, const double *Hg, double *Hs, int lane)
{
    for (int e = lane; e < n * n; e += 32) Hs[(e % n) + (e / n) * ld] = Hg[e];
    __syncwarp();
}

/* -1/2 x'Qx + x'qmod - 1/2 u'Ru + u'rmod for the node of this warp (lane i holds z_i and g_i) */
__device__ __forceinline__ double warp_dense_dual_term(int nx, int n, const double *Hs, int ld, double z, double g, int lane)
{
    double s = 0.0;
    const bool isx = lane < nx;
    for (int c = 0; c < n; c++)
    {
        const double zc = __shfl_sync(FULL, z, c);
        if (lane < n && ((c < nx) == isx)) s += Hs[lane + c * ld] * zc;
    }
    double t = (lane < n) ? (-0.5 * s * z + g * z) : 0.0;
    return warp_sum(t);
}
