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

/* shared memory per warp (doubles) for the dense solver: Hs | sL | sY (nv x ld each) | sinv (nv), and with
 * general constraint rows: Cs | Ws (ncmax x ld each) | Ss (ncmax x ncmax) | sv (2 ld + 4 ncmax) */
__host__ __device__ inline int tqb_dense_ld(int nvmax) { return nvmax | 1; }
__host__ __device__ inline int tqb_dense_gen(int nvmax, int ncmax)
{
    return ncmax > 0 ? 2 * ncmax * tqb_dense_ld(nvmax) + ncmax * ncmax + 2 * tqb_dense_ld(nvmax) + 4 * ncmax + 2 : 0;
}
__host__ __device__ inline int tqb_dense_smem(int nvmax, int ncmax = 0)
{
    return 3 * nvmax * tqb_dense_ld(nvmax) + nvmax + 1 + tqb_dense_gen(nvmax, ncmax);
}

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

/* ------------------------------------------------------------------------------------------------
 * Stage QP with general constraints (the reference's QProblem path, dual_Newton_tree_qpoases.c:181-193,
 * 288-298, 318):
 *   min 1/2 z'Hz - g'z   s.t.  lb <= z <= ub,   lc <= C z <= uc       (m <= 8 rows, n + m <= 32)
 * Lane i < n owns variable i, lane n + a owns row a (its lb/ub arguments are lc/uc).  Primal active-set
 * method: the constraints never change between the solves of a node (only g does), so the solution of the
 * previous solve is a feasible start and its working set a consistent warm start; a cold start projects
 * clip(0) onto the violated rows first.  Each iteration solves the equality-constrained QP of the working
 * set through the Cholesky factor K = chol(H_free embedded in I) the box solver uses and the Schur
 * complement of the active rows,
 *   z0 = K^-1 rhs,   w_a = K^-1 c_a|free,   S = C_f W,   S nu = C z0 - d,   z = z0 - W nu,
 * followed by the ratio test over free variables and inactive rows, or, at a subspace minimiser, by the
 * sign test of the multipliers (mu = g - Hz - C'nu on fixed variables, nu on active rows: nu > 0 <=> upper
 * side, the convention of tree_qp_out.mu_d).  On exit Ws / Ss hold W and the Cholesky factor of S of the
 * final working set (read by warp_genqp_projector), `ma` the number of active rows, `arow[k]` their indices
 * (in sv). ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ double warp_bcast_dot(const double *row, const double *vec, int n)
{   /* every lane the same dot product (shared-memory operands) */
    double s = 0.0;
    for (int i = 0; i < n; i++) s += row[i] * vec[i];
    return s;
}

__device__ __forceinline__ int warp_genqp(int n, int m, const double *Hs, int ld, const double *Cs, double g, double lb,
                                          double ub, int &ws, double &z, double &mu, double *sL, double *sinv,
                                          double *Ws, double *Ss, double *sv, int &ma_out, int lane)
{
    const bool isv = lane < n, isr = lane >= n && lane < n + m;
    const bool pinned = isv && lb >= ub;
    double *sz = sv, *sp = sv + ld, *snu = sp + ld, *srr = snu + m, *sd = srr + m; /* z | p / z0 | nu | rhs | side value */
    int *arow = reinterpret_cast<int *>(sd + m);                                   /* indices of the active rows */
    ma_out = 0;
    if (!isv && !isr) ws = 1;
    if (pinned) ws = -1;
    /* ---- starting point: the caller's z if it is feasible, else clip(0) pushed onto the violated rows ---- */
    double cz = 0.0;
    for (int attempt = 0; attempt < 60; attempt++)
    {
        if (isv) { if (!(z >= lb)) z = lb; if (!(z <= ub)) z = ub; }   /* also catches NaN */
        else z = 0.0;
        __syncwarp();
        if (isv) sz[lane] = z;
        __syncwarp();
        double viol = 0.0;
        if (isr)
        {
            cz = warp_bcast_dot(Cs + (lane - n) * ld, sz, n);
            const double tolr = 1e-9 * (1.0 + fabs(lb < -1e100 ? ub : lb));
            if (cz > ub + tolr) viol = cz - ub; else if (cz < lb - tolr) viol = cz - lb; /* signed */
        }
        const double vmax = warp_max(fabs(viol));
        if (!(vmax > 0.0)) break;
        if (attempt == 59) return TQB_DENSE_FAIL;
        /* Kaczmarz step onto the most violated row, then clip to the box again */
        const unsigned hit = __ballot_sync(FULL, fabs(viol) == vmax);
        const int src = __ffs(hit) - 1, a = src - n;
        const double vv = __shfl_sync(FULL, viol, src);
        double c = isv ? Cs[a * ld + lane] : 0.0, cc = warp_sum(c * c);
        if (isv && cc > 0.0) z -= vv / cc * c;
        if (attempt == 0 && isv) ws = 0; /* the stored working set belongs to another point */
        if (attempt == 0 && isr) ws = 0;
    }
    if (pinned) ws = -1;
    /* working set consistent with the point: a fixed variable sits on its bound, an active row on its side */
    if (isv && ws != 0 && !pinned && z != (ws < 0 ? lb : ub)) ws = 0;
    if (isr && ws != 0 && fabs(cz - (ws < 0 ? lb : ub)) > 1e-9 * (1.0 + fabs(ws < 0 ? lb : ub))) ws = 0;
    if (isv && ws == 0 && !pinned) { if (z == lb && lb > -1e100) ws = -1; else if (z == ub && ub < 1e100) ws = 1; }

    const int itmax = 10 * (n + m) + 20;
    for (int iter = 0; iter < itmax; iter++)
    {
        const unsigned fm = __ballot_sync(FULL, isv ? ws != 0 : true);      /* fixed variables (other lanes count as fixed) */
        const unsigned am = __ballot_sync(FULL, isr && ws != 0) >> n;       /* active rows */
        const int ma = __popc(am);
        __syncwarp();
        for (int e = lane; e < n * n; e += 32)
        {
            const int i = e % n, j = e / n;
            const bool fr = !((fm >> i) & 1u) && !((fm >> j) & 1u);
            sL[i + j * ld] = fr ? Hs[i + j * ld] : (i == j ? 1.0 : 0.0);
        }
        double rhs = isv ? (ws != 0 ? z : g) : 0.0;
        for (int a = 0; a < n; a++)
        {
            const double za = __shfl_sync(FULL, z, a);
            if (isv && ws == 0 && ((fm >> a) & 1u)) rhs -= Hs[lane + a * ld] * za;
        }
        __syncwarp();
        warp_potrf32(sL, ld, n, sinv, lane);
        const double inv = isv ? sinv[lane] : 0.0;
        double zt = warp_trsv_lnn32(sL, ld, n, inv, rhs, lane);
        zt = warp_trsv_ltn32(sL, ld, n, inv, zt, lane);              /* z0: minimiser without the rows */
        double nu_mine = 0.0;                                        /* multiplier of this lane's row */
        if (ma > 0)
        {
            /* W = K^-1 C_f', S = C_f W (lower), r = C z0 - d */
            if (isv) sp[lane] = zt;
            int k = 0;
            for (int a = 0; a < m; a++)
            {
                if (!((am >> a) & 1u)) continue;
                double w = (isv && ws == 0) ? Cs[a * ld + lane] : 0.0;
                w = warp_trsv_lnn32(sL, ld, n, inv, w, lane);
                w = warp_trsv_ltn32(sL, ld, n, inv, w, lane);
                if (isv) Ws[k * ld + lane] = (ws == 0) ? w : 0.0;
                if (lane == 0) arow[k] = a;
                k++;
            }
            __syncwarp();
            if (lane < ma)
            {
                const int a = arow[lane];
                for (int b = 0; b <= lane; b++) Ss[lane * m + b] = warp_bcast_dot(Cs + a * ld, Ws + b * ld, n);
            }
            /* side value d_a of every active row (held by the row's lane) */
            if (isr && ws != 0) sd[lane - n] = ws < 0 ? lb : ub;
            __syncwarp();
            if (lane < ma) srr[lane] = warp_bcast_dot(Cs + arow[lane] * ld, sp, n) - sd[arow[lane]];
            __syncwarp();
            if (lane == 0)
            {   /* Cholesky of S (ma x ma, in place, lower) and nu = S^-1 r; tiny pivots flag a dependent row */
                for (int j = 0; j < ma; j++)
                {
                    double dsum = Ss[j * m + j];
                    for (int t = 0; t < j; t++) dsum -= Ss[j * m + t] * Ss[j * m + t];
                    const double piv = dsum > 1e-14 * (1.0 + fabs(Ss[j * m + j])) ? sqrt(dsum) : 0.0;
                    Ss[j * m + j] = piv;
                    for (int i2 = j + 1; i2 < ma; i2++)
                    {
                        double s2 = Ss[i2 * m + j];
                        for (int t = 0; t < j; t++) s2 -= Ss[i2 * m + t] * Ss[j * m + t];
                        Ss[i2 * m + j] = piv > 0.0 ? s2 / piv : 0.0;
                    }
                }
                for (int j = 0; j < ma; j++)
                {
                    double s2 = srr[j];
                    for (int t = 0; t < j; t++) s2 -= Ss[j * m + t] * snu[t];
                    snu[j] = Ss[j * m + j] > 0.0 ? s2 / Ss[j * m + j] : 0.0;
                }
                for (int j = ma - 1; j >= 0; j--)
                {
                    double s2 = snu[j];
                    for (int t = j + 1; t < ma; t++) s2 -= Ss[t * m + j] * snu[t];
                    snu[j] = Ss[j * m + j] > 0.0 ? s2 / Ss[j * m + j] : 0.0;
                }
            }
            __syncwarp();
            if (isv && ws == 0)
                for (int k2 = 0; k2 < ma; k2++) zt -= Ws[k2 * ld + lane] * snu[k2];
            if (isr && ws != 0)
                for (int k2 = 0; k2 < ma; k2++) if (arow[k2] == lane - n) nu_mine = snu[k2];
        }
        /* ---- ratio test towards the subspace minimiser: free variables and inactive rows ---- */
        const double p = isv ? zt - z : 0.0;
        __syncwarp();
        if (isv) { sz[lane] = z; sp[lane] = p; }
        __syncwarp();
        double a = 2.0;
        int side = 0;
        if (isv && ws == 0)
        {
            if (p < 0.0 && zt < lb) { a = (lb - z) / p; side = -1; }
            else if (p > 0.0 && zt > ub) { a = (ub - z) / p; side = 1; }
        }
        if (isr && ws == 0)
        {
            const double c0 = warp_bcast_dot(Cs + (lane - n) * ld, sz, n), cp = warp_bcast_dot(Cs + (lane - n) * ld, sp, n);
            if (cp > 0.0 && c0 + cp > ub) { a = (ub - c0) / cp; side = 1; }
            else if (cp < 0.0 && c0 + cp < lb) { a = (lb - c0) / cp; side = -1; }
            if (a < 0.0) a = 0.0;
        }
        double amin = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amin = fmin(amin, __shfl_xor_sync(FULL, amin, o));
        if (amin < 1.0)
        {
            const unsigned hit = __ballot_sync(FULL, a == amin);
            const int block = __ffs(hit) - 1;
            if (isv && ws == 0) z += amin * p;
            if (lane == block) { if (isv) z = side < 0 ? lb : ub; ws = side; }
            continue;
        }
        if (isv && ws == 0) z = zt;
        /* ---- multipliers: fixed variables mu = g - Hz - C'nu, active rows nu; release the worst sign ---- */
        double sres = g;
        for (int c = 0; c < n; c++)
        {
            const double zc = __shfl_sync(FULL, z, c);
            if (isv) sres -= Hs[lane + c * ld] * zc;
        }
        if (isv && ma > 0)
            for (int k2 = 0; k2 < ma; k2++) sres -= Cs[arow[k2] * ld + lane] * snu[k2];
        mu = (isv && ws != 0) ? sres : (isr ? nu_mine : 0.0);
        double viol = 0.0;
        if (isv && ws != 0 && !pinned)
        {
            const double v = ws < 0 ? sres : -sres;
            if (v > 1e-12 * (1.0 + fabs(g))) viol = v;
        }
        if (isr && ws != 0)
        {
            const double v = ws > 0 ? -nu_mine : nu_mine; /* upper side needs nu >= 0, lower side nu <= 0 */
            if (v > 1e-12 * (1.0 + fabs(nu_mine))) viol = v;
        }
        const double vmax = warp_max(viol);
        if (!(vmax > 0.0)) { ma_out = ma; return TQB_DENSE_OK; }
        const unsigned hit = __ballot_sync(FULL, viol == vmax);
        if (lane == __ffs(hit) - 1) ws = 0;
    }
    return TQB_DENSE_FAIL;
}

/* P -= W S^-1 W' for the active rows of the last warp_genqp (Ws, Ss = chol(S), ma): the projector on the
 * null space of ALL active constraints, P = Z (Z'HZ)^-1 Z' (QProblem_build_elimination_matrix, qpoases.c:181-214) */
__device__ __forceinline__ void warp_genqp_projector(int n, int m, int ld, int ma, const double *Ws, const double *Ss,
                                                     double *sY, double *Pg, int lane)
{
    if (ma <= 0) return;
    __syncwarp();
    /* V = W L_s^-T column by column into sY (n x ma): v_k = (w_k - sum_{t<k} L_s[k][t] v_t) / L_s[k][k] */
    if (lane < n)
        for (int k = 0; k < ma; k++)
        {
            double s2 = Ws[k * ld + lane];
            for (int t = 0; t < k; t++) s2 -= Ss[k * m + t] * sY[t * ld + lane];
            sY[k * ld + lane] = Ss[k * m + k] > 0.0 ? s2 / Ss[k * m + k] : 0.0;
        }
    __syncwarp();
    if (lane < n)
        for (int a = 0; a < n; a++)
        {
            double s2 = 0.0;
            for (int k = 0; k < ma; k++) s2 += sY[k * ld + a] * sY[k * ld + lane];
            Pg[a + lane * n] -= s2;
        }
    __syncwarp();
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
