/*
 * tdunes_kernels.cuh -- sm_100a kernels of the dual Newton strategy on a forest of tree QPs.
 *
 * One warp owns one unit of work (a node for the stage QPs / gradient, a parent block for the
 * Hessian, factorisation and substitutions).  Every unit's data is contiguous in HBM
 * (children of a node are numbered consecutively, so [A_kids], [B_kids], lambda_kids, W_p, Ut_p
 * are single ranges) and is read with unit-stride warp accesses; small dense blocks are staged
 * in shared memory and factorised there by the warp (no library call per block).
 *
 * Kernel                replaces (reference file:line)
 *   k_stage<0>          solve_stage_problems dual_Newton_tree.c:218 + stage_qp_clipping_solve_extended clipping.c:188
 *   k_stage<2>, k_ls_eval evaluate_dual_function :823 + stage_qp_clipping_solve clipping.c:231 + eval_dual_term clipping.c:359
 *   k_gradient          build_dual_problem :519-539 + calculate_error_in_residuals :412
 *   k_hessian           build_dual_problem :551-615 + clipping.c:264-355 (set_CmPnCmT/add_EPmE/add_CmPnCkT)
 *   k_factor_level      calculate_delta_lambda :668-758 + treeqp_dpotrf_l_with_reg_opts dual_Newton_common.c:35
 *   k_backsolve_level   calculate_delta_lambda :760-775
 *   k_lambda_step / k_decide_*           line_search :922-1019, gradient_trans_times_direction :808
 *   k_finish            stage_qp_clipping_export_mu clipping.c:386
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../host/host_internal.h"

#include "warp_helpers.cuh"
#include "comm.cuh"
#include "tdunes_dense.cuh"

#define TQB_RUNNING (-1)
#define TQB_STAGE_QP_SOLVE_FAILED 4 /* TREEQP_DN_STAGE_QP_SOLVE_FAILED, utils/types.h:55-78 */

struct View
{
    tqb_tables T;
    tqb_inputs in;
    /* outputs / iterates (x,u,lam,mux,muu contiguous: one D2H copy) */
    double *x, *u, *lam, *mux, *muu;
    double *dlam, *res, *resMod, *xUnc, *uUnc, *P, *Rt, *Linv;
    double *qmod0, *rmod0, *dq, *dr, *cmod0, *dc; /* affine model of the stage gradients along dlam */
    double *W, *Ut;
    const int *leaf_desc; /* fast path: LeafDesc records (tdunes_stage.cuh), one per leaf in node order */
    const int *lvl_desc;  /* fast path: LvlDesc records (tdunes_fused.cuh), one per parent in level order */
    double *schur, *rsub; /* fast path: per-node Schur complement (NX x NX) and its rhs part (x-space) */
    /* fast path, re-use of unchanged blocks (the GPU form of checkLastActiveSet, dual_Newton_tree.c:334-371): per node,
     * chg = the node's own active set differs from the last Newton iteration (k_stage_fused<0>), bdirty = its block
     * was re-factored in this iteration (k_build_factor), breg = its last factorisation was regularised on the fly */
    int *chg, *bdirty, *breg;
    const int *ghost; /* sharded local tree: cut nodes owned by another rank (their stage solution arrives with the exchange); else NULL */
    double *nva, *nvb, *cva, *cvb; /* per-node and per-chunk reduction scratch */
    /* per tree */
    int *status, *iter, *ls_cnt, *ls_active, *restart_cnt;
    double *t_f0, *t_dot, *t_tau, *t_tauprev, *t_err, *t_taufin;
    int *counters; /* [0] trees still running, [1] trees with a pending line-search trial, [2] regularised blocks, [3] total trials */
    int *ls_hist;  /* [256] trials per Newton iteration of tree 0 */
    int Nn, nT, X, U, nchunk;
    /* sharded solve: per-tree sums come all-reduced in t_red ([0,nT): norm / descent product,
     * [nT,2nT): dual function) instead of from this rank's chunk values */
    double *t_red;
    int sharded;
    int top_nodes; /* sharded solve with the termination norm riding on the cut exchange: local nodes [0, top_nodes) are the replicated top */
    /* dense stage QPs (tdunes_dense.cuh): projectors P_k at hoff[k], dual terms of the dense nodes
     * of the current line-search trial; for a dense node the x-/u-space slots of P/Rt hold its
     * working set (-1, 0, +1) and those of xUnc/uUnc its bound multipliers */
    double *Pd, *fdense;
    int has_dense, nvmax;
    /* general stage constraints of the dense nodes: multipliers (output, d-space) and working set */
    double *mud, *wsd;
    int ncmax, NC;
    /* per-warp shared memory layout, in doubles */
    int abmax, dmax, nxmax, numax, pw_stage, pw_hess;
    tqb_opts o;
};

#ifndef TQB_LS_NT
#define TQB_LS_NT 8 /* step lengths evaluated per line-search pass (k_ls_eval, k_trial_dense) */
#endif

__global__ void k_init(View v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.X) v.lam[i] = v.in.lam0[i];
    if (i < v.Nn) v.nva[i] = 0.0; /* leaves never write their slot of the termination norm */
    if (v.chg)
    {   /* no active set yet: the first comparison of k_stage_fused<0> reports a change everywhere (P, Rt >= 0 later on),
         * and kernels that never clear a flag leave their nodes marked for ever */
        if (i < v.X) v.P[i] = -1.0;
        if (i < v.U) v.Rt[i] = -1.0;
        if (i < v.Nn) { v.chg[i] = 1; v.bdirty[i] = 1; v.breg[i] = 0; }
    }
    if (v.has_dense)
    {   /* empty working sets (QProblemB_init is a cold start, qpoases.c:286) */
        if (i < v.X) v.P[i] = 0.0;
        if (i < v.U) v.Rt[i] = 0.0;
        if (i < v.Nn)
            for (int m = 0; m < TQB_LS_NT; m++) v.fdense[(size_t) m * v.Nn + i] = 0.0; /* clipping nodes keep zeros */
        if (v.ncmax > 0 && i < v.NC) { v.wsd[i] = 0.0; v.mud[i] = 0.0; }
    }
    if (i < v.nT)
    {
        v.status[i] = TQB_RUNNING;
        v.iter[i] = 0;
        v.ls_cnt[i] = 0;
        v.ls_active[i] = 0;
        v.restart_cnt[i] = 0;
    }
    if (i < 256) v.ls_hist[i] = 0;
    if (i < 4) v.counters[i] = 0;
}

/* ------------------------------------------------------------------------------------------
 * Stage gradients / stage QPs of all nodes, one warp per node.
 *   MODE 0  "extended" solve at lambda (solve_stage_problems + stage_qp_clipping_solve_extended):
 *           qmod = -q + lambda_k - sum_kids A_j' lambda_j, rmod = -r - sum_kids B_j' lambda_j,
 *           x = clip(Qinv.*qmod), P = Qinv.*[inactive] (same for u); keeps qmod/rmod/cmod as the
 *           constant part of the affine model used by the line search.
 *   MODE 2  direction: the same operator applied to dlam (without q, r): dq, dr, dc and the
 *           node's part of res'*dlam.  Along lambda + tau*dlam the stage gradients are
 *           qmod0 + tau*dq, so a line-search trial never touches A or B again (k_ls_eval).
 *   MODE 3  clip-only solve at lambda writing x,u (state of the reference's work->sx after its
 *           last line-search trial; only needed when maxIter is hit).
 * Algorithmic bytes per node (MODE 0): (5nx+4nu) + sum_kids (nx_j(nx+nu) + 2nx_j) read,
 * 5nx+5nu+1 written.
 * ---------------------------------------------------------------------------------------- */
template <int MODE>
__global__ void __launch_bounds__(256) k_stage(View v)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int g = blockIdx.x * wpb + wib;
    if (g >= v.Nn) return;
    const int t = v.T.tree[g];
    if (v.status[t] != TQB_RUNNING) return;

    double *sAB = smem + (size_t) wib * v.pw_stage;
    double *sL = sAB + v.abmax;
    const int nx = v.T.nx[g], nu = v.T.nu[g], xo = v.T.xoff[g], uo = v.T.uoff[g];
    const int hoff = v.has_dense ? v.T.hoff[g] : -1;
    double gz = 0.0; /* dense node: stage gradient component of this lane's variable */
    const int nk = v.T.nkids[g], k0 = v.T.kid0[g], d = v.T.dim[g];
    const bool root = v.T.dad[g] < 0;
    const double *src = (MODE == 2) ? v.dlam : v.lam;

    /* stage [A_kids] (d*nx), [B_kids] (d*nu) and the kids' multipliers (d): contiguous ranges */
    double cpart = 0.0;
    if (nk > 0)
    {
        const double *Ag = v.in.A + v.T.aoff[k0], *Bg = v.in.B + v.T.boff[k0];
        const double *lg = src + v.T.xoff[k0];
        const int na = d * nx, nb = d * nu;
        for (int e = lane; e < na; e += 32) sAB[e] = Ag[e];
        for (int e = lane; e < nb; e += 32) sAB[na + e] = Bg[e];
        for (int e = lane; e < d; e += 32)
        {
            const double l = lg[e];
            sL[e] = l;
            if (MODE != 3) cpart += v.in.b[v.T.xoff[k0] + e] * l;
        }
        __syncwarp();
    }
    if (MODE != 3)
    {   /* cmod = sum_kids b_j' lambda_j (dual_Newton_tree.c:892) */
        const double c = warp_sum(cpart);
        if (lane == 0) { if (MODE == 0) v.cmod0[g] = c; else v.dc[g] = c; }
    }

    double dotp = 0.0;
    for (int w = lane; w < nx + nu; w += 32)
    {
        const bool isx = w < nx;
        const int i = isx ? w : w - nx;
        double acc;
        if (MODE == 2) acc = (isx && !root) ? src[xo + i] : 0.0;
        else if (isx) acc = -v.in.q[xo + i] + (root ? 0.0 : src[xo + i]);
        else acc = -v.in.r[uo + i];
        if (MODE == 2 && isx && !root) dotp += v.res[xo + i] * src[xo + i];
        /* acc -= A_j' l_j (resp. B_j' l_j) child by child (dgemv_t per child) */
        int oa = isx ? 0 : d * nx, pos = 0;
        const int ncol = isx ? nx : nu;
        for (int j = 0; j < nk; j++)
        {
            const int nxj = v.T.nx[k0 + j];
            const double *col = sAB + oa + i * nxj;
            double s = 0.0;
            for (int r = 0; r < nxj; r++) s += col[r] * sL[pos + r];
            acc -= s;
            oa += nxj * ncol;
            pos += nxj;
        }
        if (MODE == 2)
        {
            if (isx) v.dq[xo + i] = acc; else v.dr[uo + i] = acc;
            continue;
        }
        if (hoff >= 0) { gz = acc; continue; } /* nx + nu <= 32: one variable per lane */
        int mask;
        if (isx)
        {
            const double unc = v.in.Qinv[xo + i] * acc;
            v.x[xo + i] = clip_mask(unc, v.in.xmin[xo + i], v.in.xmax[xo + i], mask);
            if (MODE == 0)
            {
                v.qmod0[xo + i] = acc;
                v.xUnc[xo + i] = unc;
                v.P[xo + i] = mask == 0 ? v.in.Qinv[xo + i] : 0.0;
            }
        }
        else
        {
            const double unc = v.in.Rinv[uo + i] * acc;
            v.u[uo + i] = clip_mask(unc, v.in.umin[uo + i], v.in.umax[uo + i], mask);
            if (MODE == 0)
            {
                v.rmod0[uo + i] = acc;
                v.uUnc[uo + i] = unc;
                v.Rt[uo + i] = mask == 0 ? v.in.Rinv[uo + i] : 0.0;
            }
        }
    }
    if (MODE == 2)
    {
        dotp = warp_sum(dotp);
        if (lane == 0) v.nvb[g] = dotp;
    }
    if (MODE != 2 && hoff >= 0)
    {   /* dense stage QP (stage_qp_qpoases_solve_extended / _solve, qpoases.c:360-398) */
        const int n = nx + nu, ld = tqb_dense_ld(v.nvmax);
        double *Hs = sL + v.dmax, *sLd = Hs + v.nvmax * ld, *sY = sLd + v.nvmax * ld, *sinv = sY + v.nvmax * ld;
        __syncwarp();
        warp_load_H(n, ld, v.in.H + hoff, Hs, lane);
        const int m = v.ncmax > 0 ? v.T.ncn[g] : 0, dof = m > 0 ? v.T.doff[g] : 0;
        const bool isx = lane < nx, act = lane < n, isr = lane >= n && lane < n + m;
        const int i = isx ? lane : lane - nx;
        double lb = 0.0, ub = 0.0, z = 0.0, mu = 0.0;
        int ws = 0;
        if (act)
        {
            lb = isx ? v.in.dxmin[xo + i] : v.in.dumin[uo + i];
            ub = isx ? v.in.dxmax[xo + i] : v.in.dumax[uo + i];
            ws = (int) (isx ? v.P[xo + i] : v.Rt[uo + i]);
        }
        int fail, ma = 0;
        double *Cs = sinv + v.nvmax + 1, *Ws = Cs + v.ncmax * ld, *Ss = Ws + v.ncmax * ld, *sv = Ss + v.ncmax * v.ncmax;
        if (m > 0)
        {   /* general constraints: rows of [C D] into shared memory, lane n + a owns row a */
            const double *Cg = v.in.Cg + v.T.cgoff[g];
            for (int e = lane; e < m * n; e += 32) Cs[(e % m) * ld + e / m] = Cg[e];
            if (isr) { lb = v.in.dmin[dof + lane - n]; ub = v.in.dmax[dof + lane - n]; ws = (int) v.wsd[dof + lane - n]; }
            /* start from the previous solution of this node (feasible: only the gradient changes); first
             * iteration of a solve: from the origin */
            if (act) z = v.iter[t] == 0 ? 0.0 : (isx ? v.x[xo + i] : v.u[uo + i]);
            __syncwarp();
            fail = warp_genqp(n, m, Hs, ld, Cs, gz, lb, ub, ws, z, mu, sLd, sinv, Ws, Ss, sv, ma, lane);
        }
        else
            fail = warp_boxqp(n, Hs, ld, gz, lb, ub, ws, z, mu, sLd, sinv, lane);
        if (fail && lane == 0) v.status[t] = TQB_STAGE_QP_SOLVE_FAILED;
        if (act)
        {
            if (isx) v.x[xo + i] = z; else v.u[uo + i] = z;
            if (MODE == 0)
            {
                if (isx) { v.qmod0[xo + i] = gz; v.P[xo + i] = (double) ws; v.xUnc[xo + i] = mu; }
                else { v.rmod0[uo + i] = gz; v.Rt[uo + i] = (double) ws; v.uUnc[uo + i] = mu; }
            }
        }
        if (isr && MODE == 0) { v.wsd[dof + lane - n] = (double) ws; v.mud[dof + lane - n] = mu; }
        if (MODE == 0)
        {
            const int wsv = act ? ws : 1; /* the projector of the box part reads the variables' working set only */
            warp_boxqp_projector(n, ld, wsv, sLd, sinv, sY, v.Pd + hoff, lane);
            if (m > 0) warp_genqp_projector(n, m, ld, ma, Ws, Ss, sY, v.Pd + hoff, lane);
        }
    }
}

/* Line-search trial of the dense nodes (stage_qp_qpoases_solve + eval_dual_term, qpoases.c:384,478):
 * box QP at the stage gradient g0 + tau*dg, warm-started from the working set of the last extended
 * solve; the node's dual term goes to fdense[g] and is picked up by k_ls_eval's chunk reduction. */
__global__ void __launch_bounds__(256) k_trial_dense(View v, int first, int nt)
{
    /* Batched like k_ls_eval: the step lengths of one line-search pass are solved in ONE launch (slot m of
     * fdense; first = 1: slot 0 is tau = 0, then 1, beta, beta^2, ...).  Every solve starts from the working
     * set (and, with general constraints, the point) of the last extended solve, so the trials are
     * independent of each other. */
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    /* one warp per (node, step length): the solves of a pass run side by side */
    const int wid = blockIdx.x * wpb + wib, g = wid / nt, mslot = wid - g * nt;
    if (g >= v.Nn) return;
    const int hoff = v.T.hoff[g];
    if (hoff < 0) return;
    const int t = v.T.tree[g];
    if (first ? (v.status[t] != TQB_RUNNING) : !v.ls_active[t]) return;
    const int nx = v.T.nx[g], nu = v.T.nu[g], xo = v.T.xoff[g], uo = v.T.uoff[g];
    const int n = nx + nu, ld = tqb_dense_ld(v.nvmax);
    double *Hs = smem + (size_t) wib * tqb_dense_smem(v.nvmax, v.ncmax), *sLd = Hs + v.nvmax * ld, *sinv = sLd + 2 * v.nvmax * ld;
    double *Cs = sinv + v.nvmax + 1, *Ws = Cs + v.ncmax * ld, *Ss = Ws + v.ncmax * ld, *sv = Ss + v.ncmax * v.ncmax;
    warp_load_H(n, ld, v.in.H + hoff, Hs, lane);
    const int mrows = v.ncmax > 0 ? v.T.ncn[g] : 0, dof = mrows > 0 ? v.T.doff[g] : 0;
    const bool isx = lane < nx, act = lane < n, isr = lane >= n && lane < n + mrows;
    const int i = isx ? lane : lane - nx;
    double lb = 0.0, ub = 0.0, g0 = 0.0, dg = 0.0, z0 = 0.0;
    int ws0 = 0;
    if (act)
    {
        lb = isx ? v.in.dxmin[xo + i] : v.in.dumin[uo + i];
        ub = isx ? v.in.dxmax[xo + i] : v.in.dumax[uo + i];
        ws0 = (int) (isx ? v.P[xo + i] : v.Rt[uo + i]);
        g0 = isx ? v.qmod0[xo + i] : v.rmod0[uo + i];
        dg = isx ? v.dq[xo + i] : v.dr[uo + i];
        z0 = isx ? v.x[xo + i] : v.u[uo + i];
    }
    if (mrows > 0)
    {
        const double *Cg = v.in.Cg + v.T.cgoff[g];
        for (int e = lane; e < mrows * n; e += 32) Cs[(e % mrows) * ld + e / mrows] = Cg[e];
        if (isr) { lb = v.in.dmin[dof + lane - n]; ub = v.in.dmax[dof + lane - n]; ws0 = (int) v.wsd[dof + lane - n]; }
        __syncwarp();
    }
    double tm = first ? 1.0 : v.t_tau[t];
    for (int m = 0; m < nt; m++)
    {   /* nt <= TQB_LS_NT slots: a dense solve is not a stream over vectors, the first pass of an iteration
         * only looks at tau = 0 and the full step */
        double tau = tm;
        if (first && m == 0) tau = 0.0; else tm = v.o.beta * tm;
        if (m != mslot) continue; /* the step lengths are generated by repeated multiplication, as k_ls_decide does */
        const double gz = g0 + tau * dg;
        double z = z0, mu = 0.0;
        int ws = ws0, fail, ma = 0;
        if (mrows > 0) fail = warp_genqp(n, mrows, Hs, ld, Cs, gz, lb, ub, ws, z, mu, sLd, sinv, Ws, Ss, sv, ma, lane);
        else fail = warp_boxqp(n, Hs, ld, gz, lb, ub, ws, z, mu, sLd, sinv, lane);
        if (fail && lane == 0) v.status[t] = TQB_STAGE_QP_SOLVE_FAILED;
        const double f = warp_dense_dual_term(nx, n, Hs, ld, act ? z : 0.0, gz, lane);
        if (lane == 0) v.fdense[(size_t) m * v.Nn + g] = f;
        __syncwarp();
    }
}

/* mu_x, mu_u of the dense nodes: the multipliers of the last extended solve (qpoases.c:527-565) */
__global__ void k_finish_dense(View v)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= v.Nn || v.T.hoff[g] < 0) return;
    const int nx = v.T.nx[g], nu = v.T.nu[g], xo = v.T.xoff[g], uo = v.T.uoff[g];
    for (int i = 0; i < nx; i++) v.mux[xo + i] = v.xUnc[xo + i];
    for (int i = 0; i < nu; i++) v.muu[uo + i] = v.uUnc[uo + i];
}

/* ------------------------------------------------------------------------------------------
 * Dual gradient: res_k = b_k - x_k + A_k x_dad + B_k u_dad for every non-root node, plus the
 * node's contribution to the termination norm in nva[g].
 * Bytes per edge: nx_k(nx_p+nu_p) + 2nx_k + nx_p + nu_p read, 2nx_k written.
 * ---------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) k_gradient(View v)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= v.Nn) return;
    if (v.status[v.T.tree[g]] != TQB_RUNNING) return;
    const int p = v.T.dad[g];
    double part = 0.0;
    if (p >= 0)
    {
        const int nx = v.T.nx[g], nxp = v.T.nx[p], nup = v.T.nu[p], xo = v.T.xoff[g];
        const double *A = v.in.A + v.T.aoff[g], *B = v.in.B + v.T.boff[g];
        const double *xp = v.x + v.T.xoff[p], *up = v.u + v.T.uoff[p];
        for (int r = lane; r < nx; r += 32)
        {
            double s = v.in.b[xo + r] - v.x[xo + r];
            double s1 = 0.0, s2 = 0.0;
            for (int c = 0; c < nxp; c++) s1 += A[r + c * nx] * xp[c];
            for (int c = 0; c < nup; c++) s2 += B[r + c * nx] * up[c];
            s += s1;
            s += s2;
            v.res[xo + r] = s;
            v.resMod[xo + r] = s;
            if (v.o.termCondition == 2) part = fmax(part, fabs(s));
            else part += s * s;
        }
    }
    part = (v.o.termCondition == 2) ? warp_max(part) : warp_sum(part);
    if (lane == 0) v.nva[g] = part;
}

/* fixed-shape reduction of a per-node array over each chunk of <=1024 consecutive nodes */
__global__ void __launch_bounds__(256) k_chunk_reduce(View v, const double *src, double *dst, int is_max, int g_min = 0)
{
    __shared__ double sh[256];
    const int c = blockIdx.x;
    /* g_min > 0 (sharded solve, termination norm taken before the cut exchange): nodes of the replicated top are left out */
    const int beg = v.T.chunk_beg[c] > g_min ? v.T.chunk_beg[c] : g_min, end = v.T.chunk_end[c];
    double acc = 0.0;
    for (int g = beg + threadIdx.x; g < end; g += 256)
    {
        const double val = src[g];
        acc = is_max ? fmax(acc, val) : acc + val;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if (threadIdx.x < s)
            sh[threadIdx.x] = is_max ? fmax(sh[threadIdx.x], sh[threadIdx.x + s]) : sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) dst[c] = sh[0];
}

/* per-tree reduction of the chunk values: every lane returns the result */
__device__ __forceinline__ double tree_reduce(const View &v, const double *cv, int t, int lane, bool is_max)
{
    const int c0 = v.T.tchunk0[t], c1 = v.T.tchunk0[t + 1];
    double acc = 0.0;
    for (int c = c0 + lane; c < c1; c += 32) acc = is_max ? fmax(acc, cv[c]) : acc + cv[c];
    return is_max ? warp_max(acc) : warp_sum(acc);
}

/* sharded solve: this rank's part of a per-tree sum / max, to be all-reduced */
__global__ void k_tree_partial(View v, const double *cv, double *out, int is_max)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= v.nT) return;
    const double r = tree_reduce(v, cv, t, lane, is_max != 0);
    if (lane == 0) out[t] = r;
}

/* sharded solve: what a cut node contributes to its parent's block (stage solution, active-set
 * mask, right-hand side part, Schur complement), packed for the all-gather / scattered after it.
 * One warp per cut node; `per` = nx(nx+3) doubles. */
__global__ void k_cut_pack(View v, const int *cut_nodes, int first, int count, int nx, double *buf)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= count) return;
    const int g = cut_nodes[first + w], xo = v.T.xoff[g];
    double *dst = buf + (size_t) w * nx * (nx + 3);
    for (int e = lane; e < nx; e += 32)
    {
        dst[e] = v.x[xo + e];
        dst[nx + e] = v.P[xo + e];
        dst[2 * nx + e] = v.rsub[xo + e];
    }
    const double *S = v.schur + (size_t) g * nx * nx;
    for (int e = lane; e < nx * nx; e += 32) dst[3 * nx + e] = S[e];
}

__global__ void k_cut_unpack(View v, const int *cut_nodes, int count, int nx, const double *buf)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= count) return;
    const int g = cut_nodes[w], xo = v.T.xoff[g];
    const double *src = buf + (size_t) w * nx * (nx + 3);
    for (int e = lane; e < nx; e += 32)
    {
        v.x[xo + e] = src[e];
        v.P[xo + e] = src[nx + e];
        v.rsub[xo + e] = src[2 * nx + e];
    }
    double *S = v.schur + (size_t) g * nx * nx;
    for (int e = lane; e < nx * nx; e += 32) S[e] = src[3 * nx + e];
}

/* ---- the same exchanges through the peer mailboxes (comm.cuh), ONE CTA each ------------------------ */

/* all payload stores of this CTA are done -> publish `seq` to every peer, wait for every peer's `seq` */
__device__ __forceinline__ void p2p_signal_wait(const tqb_p2p_dev &x, unsigned long long seq)
{
    __threadfence_system();
    __syncthreads();
    const int par = (int) (seq & 1);
    if ((int) threadIdx.x < x.nranks)
    {
        const int r = threadIdx.x;
        volatile unsigned long long *theirs = x.flags[r] + par * TQB_P2P_MAXR + x.rank;
        *theirs = seq;
        volatile unsigned long long *mine = x.flags[x.rank] + par * TQB_P2P_MAXR + r;
        const long long t0 = clock64();
        while (*mine != seq)
            if (clock64() - t0 > 6000000000LL) { atomicExch(x.err, 1); break; } /* ~3 s: a peer died */
    }
    __syncthreads();
    __threadfence_system();
}

__device__ __forceinline__ double *p2p_slot(const tqb_p2p_dev &x, int owner, unsigned long long seq, int src)
{
    return x.slots[owner] + ((size_t) (seq & 1) * x.nranks + src) * x.cap;
}

/* cut nodes -> every rank (k_cut_pack + all-gather + k_cut_unpack in one launch) */
/* `tred` != NULL: the ranks' partial termination norms (nT values, the subtrees below the cut) ride along behind the cut
 * nodes and are combined in rank order (max when is_max), saving the separate exchange after the sweep */
__global__ void __launch_bounds__(512) k_cut_exchange(View v, tqb_p2p_dev x, unsigned long long seq, const int *cut_nodes,
                                                      int my_cut0, int my_ncut, int ncut, int nx, double *tred, int nT, int is_max)
{
    const int per = nx * (nx + 3), nmine = my_ncut * per;
    if (tred)
        for (int e = threadIdx.x; e < nT; e += blockDim.x)
        {
            const double val = tred[e];
            for (int r = 0; r < x.nranks; r++) p2p_slot(x, r, seq, x.rank)[nmine + e] = val;
        }
    for (int e = threadIdx.x; e < nmine; e += blockDim.x)
    {
        const int w = e / per, i = e - w * per;
        const int g = cut_nodes[my_cut0 + w], xo = v.T.xoff[g];
        double val;
        if (i < nx) val = v.x[xo + i];
        else if (i < 2 * nx) val = v.P[xo + i - nx];
        else if (i < 3 * nx) val = v.rsub[xo + i - 2 * nx];
        else val = v.schur[(size_t) g * nx * nx + (i - 3 * nx)];
        for (int r = 0; r < x.nranks; r++) p2p_slot(x, r, seq, x.rank)[e] = val;
    }
    p2p_signal_wait(x, seq);
    const int ntot = ncut * per;
    for (int e = threadIdx.x; e < ntot; e += blockDim.x)
    {
        const int q = e / per, i = e - q * per, src = q / my_ncut, w = q - src * my_ncut;
        if (src == x.rank) continue; /* own nodes are already in place */
        const double val = __ldcv(p2p_slot(x, x.rank, seq, src) + w * per + i);
        const int g = cut_nodes[q], xo = v.T.xoff[g];
        if (i < nx) v.x[xo + i] = val;
        else if (i < 2 * nx) v.P[xo + i - nx] = val;
        else if (i < 3 * nx) v.rsub[xo + i - 2 * nx] = val;
        else v.schur[(size_t) g * nx * nx + (i - 3 * nx)] = val;
    }
    if (tred)
        for (int e = threadIdx.x; e < nT; e += blockDim.x)
        {
            double acc = __ldcv(p2p_slot(x, x.rank, seq, 0) + nmine + e);
            for (int r = 1; r < x.nranks; r++)
            {
                const double t = __ldcv(p2p_slot(x, x.rank, seq, r) + nmine + e);
                acc = is_max ? fmax(acc, t) : acc + t;
            }
            tred[e] = acc;
        }
}

/* in-place all-reduce of buf[0..count): the first `nmax` entries with max, the others as sums in rank
 * order (the same association on every rank: all ranks hold bit-identical results) */
__global__ void __launch_bounds__(256) k_red_exchange(tqb_p2p_dev x, unsigned long long seq, double *buf, int count, int nmax)
{
    for (int e = threadIdx.x; e < count; e += blockDim.x)
    {
        const double val = buf[e];
        for (int r = 0; r < x.nranks; r++) p2p_slot(x, r, seq, x.rank)[e] = val;
    }
    p2p_signal_wait(x, seq);
    for (int e = threadIdx.x; e < count; e += blockDim.x)
    {
        double acc = __ldcv(p2p_slot(x, x.rank, seq, 0) + e);
        for (int r = 1; r < x.nranks; r++)
        {
            const double t = __ldcv(p2p_slot(x, x.rank, seq, r) + e);
            acc = e < nmax ? fmax(acc, t) : acc + t;
        }
        buf[e] = acc;
    }
}

/* termination test per tree (dual_Newton_tree.c:542-546) */
__global__ void k_decide_term(View v)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= v.nT) return;
    if (v.status[t] != TQB_RUNNING) { if (lane == 0) v.ls_active[t] = 0; return; }
    double err = v.sharded ? v.t_red[t] : tree_reduce(v, v.cva, t, lane, v.o.termCondition == 2);
    if (v.sharded && v.top_nodes > 0)
    {   /* t_red came with the cut exchange (subtrees below the cut); the replicated top is every rank's own, in node order */
        for (int g = 0; g < v.top_nodes; g++) err = (v.o.termCondition == 2) ? fmax(err, v.nva[g]) : err + v.nva[g];
    }
    if (v.o.termCondition == 1) err = sqrt(err);
    if (lane == 0)
    {
        v.t_err[t] = err;
        if (err < v.o.tol) { v.status[t] = 0; v.ls_active[t] = 0; }
        else { v.ls_active[t] = 1; atomicAdd(&v.counters[0], 1); }
    }
}

/* ------------------------------------------------------------------------------------------
 * Dual Hessian block of every parent p:
 *   W_p[k,s] = A_k P_p A_s' + B_k Rt_p B_s' (+ P_k if k==s), lower triangle, column-major d x d
 *   Ut_p     = -P_p [A_k1' ... A_km']  (nx_p x d), only if p is not a root
 * Bytes per parent: d(nx+nu) + nx + nu + d read; d(d+1)/2 + nx*d written.
 * ---------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) k_hessian(View v)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    /* per warp (doubles): sAB (abmax) | sP (nxmax+numax) | row tables (dmax doubles = 2 int tables)
     * and with dense stage QPs: kid table (dmax/2+1) | P_p (nvmax^2) | CP (dmax*nvmax) */
    const int g = blockIdx.x * wpb + wib;
    if (g >= v.Nn) return;
    const int nk = v.T.nkids[g];
    if (nk == 0) return;
    if (v.status[v.T.tree[g]] != TQB_RUNNING) return;

    double *sAB = smem + (size_t) wib * v.pw_hess;
    double *sP = sAB + v.abmax;
    int *rbase = (int *) (sP + v.nxmax + v.numax);
    int *rstr = rbase + v.dmax;
    int *rkid = rstr + v.dmax;                                 /* only with dense stage QPs */
    double *sPd = (double *) (rbase + 2 * v.dmax) + v.dmax / 2 + 1; /* P_p, nv x nv */
    double *sCP = sPd + v.nvmax * v.nvmax;                     /* C P_p, d x nv */

    const int nx = v.T.nx[g], nu = v.T.nu[g], d = v.T.dim[g], k0 = v.T.kid0[g];
    const int hp = v.has_dense ? v.T.hoff[g] : -1;
    const int na = d * nx, nb = d * nu, nv = nx + nu;
    const double *Ag = v.in.A + v.T.aoff[k0], *Bg = v.in.B + v.T.boff[k0];
    for (int e = lane; e < na; e += 32) sAB[e] = Ag[e];
    for (int e = lane; e < nb; e += 32) sAB[na + e] = Bg[e];
    if (hp >= 0)
        for (int e = lane; e < nv * nv; e += 32) sPd[e] = v.Pd[hp + e];
    else
    {
        for (int e = lane; e < nx; e += 32) sP[e] = v.P[v.T.xoff[g] + e];
        for (int e = lane; e < nu; e += 32) sP[nx + e] = v.Rt[v.T.uoff[g] + e];
    }
    /* row tables: row I of the stacked [A_kids;B_kids] belongs to a kid that starts at stacked row
     * `rows` and has nxj rows: element (I,c) of A is sAB[rows*nx + i + c*nxj], of B is
     * sAB[na + rows*nu + i + c*nxj].  rbase packs (rows<<16 | i). */
    if (lane == 0)
    {
        int rows = 0, I = 0;
        for (int j = 0; j < nk; j++)
        {
            const int nxj = v.T.nx[k0 + j];
            for (int i = 0; i < nxj; i++)
            {
                rbase[I] = rows * 65536 + i; rstr[I] = nxj;
                if (v.has_dense) rkid[I] = j;
                I++;
            }
            rows += nxj;
        }
    }
    __syncwarp();
    if (hp >= 0)
    {   /* CP[I][a] = sum_c C[I][c] P_p[c][a]  (M = C P, set_CmPnCmT qpoases.c:401-426) */
        for (int e = lane; e < d * nv; e += 32)
        {
            const int I = e % d, a = e / d;
            const int rowsI = rbase[I] >> 16, ii = rbase[I] & 65535, nxk = rstr[I];
            const double *aI = sAB + rowsI * nx + ii, *bI = sAB + na + rowsI * nu + ii;
            double acc = 0.0;
            for (int c = 0; c < nx; c++) acc += aI[c * nxk] * sPd[c + a * nv];
            for (int c = 0; c < nu; c++) acc += bI[c * nxk] * sPd[nx + c + a * nv];
            sCP[I + a * d] = acc;
        }
        __syncwarp();
    }
    double *Wp = v.W + v.T.woff[g];
    const double *Pk = v.P + v.T.xoff[k0];
    for (int J = 0; J < d; J++)
    {
        const int rowsJ = rbase[J] >> 16, jj = rbase[J] & 65535, nxs = rstr[J];
        const double *aJ = sAB + rowsJ * nx + jj;
        const double *bJ = sAB + na + rowsJ * nu + jj;
        for (int I = J + lane; I < d; I += 32)
        {
            const int rowsI = rbase[I] >> 16, ii = rbase[I] & 65535, nxk = rstr[I];
            const double *aI = sAB + rowsI * nx + ii;
            const double *bI = sAB + na + rowsI * nu + ii;
            double acc = 0.0, acc2 = 0.0;
            if (hp >= 0)
            {   /* W[I][J] = sum_a C[I][a] (C P)[J][a] */
                for (int c = 0; c < nx; c++) acc += aI[c * nxk] * sCP[J + c * d];
                for (int c = 0; c < nu; c++) acc += bI[c * nxk] * sCP[J + (nx + c) * d];
            }
            else
            {
                for (int c = 0; c < nx; c++) acc += aI[c * nxk] * (aJ[c * nxs] * sP[c]);
                for (int c = 0; c < nu; c++) acc2 += bI[c * nxk] * (bJ[c * nxs] * sP[nx + c]);
                acc += acc2;
            }
            /* the kid's own term E P_k E' (add_EPmE): diagonal for a clipping kid, the xx block of
             * its projector for a dense kid */
            if (v.has_dense)
            {
                if (rkid[I] == rkid[J])
                {
                    const int kid = k0 + rkid[I], hk = v.T.hoff[kid];
                    if (hk >= 0) acc += v.Pd[hk + ii + jj * (nxk + v.T.nu[kid])];
                    else if (I == J) acc += Pk[I];
                }
            }
            else if (I == J) acc += Pk[I];
            Wp[I + J * d] = acc;
        }
    }
    if (v.T.dad[g] >= 0)
    {
        double *Utp = v.Ut + v.T.utoff[g];
        for (int e = lane; e < nx * d; e += 32)
        {
            const int c = e % nx, I = e / nx;
            if (hp >= 0) { Utp[e] = -sCP[I + c * d]; continue; }
            const int rowsI = rbase[I] >> 16, ii = rbase[I] & 65535, nxk = rstr[I];
            Utp[e] = -(sAB[rowsI * nx + ii + c * nxk] * sP[c]);
        }
    }
}

/* One depth level of the leaves-to-root sweep: factor W_p (+ Levenberg-Marquardt policy of
 * dual_Newton_common.c:35-78), y = L^-1 resMod_p, Utilde = Ut_p L^-T, then downdate the parent's
 * diagonal sub-block and right-hand side.  Roots finish with the backward substitution.
 * Bytes per block: d^2 + d + nx*d read, d(d+1)/2 + d + nx*d + d written, nx(nx+1)+2nx RMW. */
/* nlev > 1: depths lvl_begin .. are a chain segment (one kid per parent, `count` parents each, block w of a
 * depth is the parent of block w of the next one): the warp walks its chain from the deepest block up in ONE
 * launch; what a block leaves for its parent (W, resMod in HBM) is re-read by the same warp. */
/* DC > 0: every block of the launch has d == nx == DC (the chains of a tree with one state dimension; the host checks
 * it per depth, device.cu: lvl_udim): loop bounds, strides and the index arithmetic become compile-time constants, the
 * operation order is the one of DC == 0. */
template <int DC>
__global__ void __launch_bounds__(256) k_factor_level(View v, int lvl_begin, int count, int per_warp, int nlev)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (w >= count) return;
    if (v.status[v.T.tree[v.T.lvl_nodes[lvl_begin + w]]] != TQB_RUNNING) return;
#pragma unroll 1
    for (int lev = nlev - 1; lev >= 0; lev--)
    {
    const int g = v.T.lvl_nodes[lvl_begin + lev * count + w];
    __syncwarp();

    const int d = DC ? DC : v.T.dim[g], nx = DC ? DC : v.T.nx[g], k0 = v.T.kid0[g], dad = v.T.dad[g];
    const int ld = d | 1, ldu = nx | 1;
    double *sW = smem + (size_t) wib * per_warp;
    double *sy = sW + (v.dmax | 1) * v.dmax;
    double *sinv = sy + v.dmax;
    double *sU = sinv + v.dmax;
    double *Wp = v.W + v.T.woff[g];
    const int xb = v.T.xoff[k0]; /* start of this block's segment in x-space */

    for (int e = lane; e < d * d; e += 32)
    {
        const int i = e % d, j = e / d;
        double val = Wp[e];
        if (i == j && v.o.regType == 1) val += v.o.regValue;
        sW[i + j * ld] = val;
    }
    for (int e = lane; e < d; e += 32) sy[e] = v.resMod[xb + e];
    double *Utp = v.Ut + v.T.utoff[g];
    if (dad >= 0) /* staged with W: the loads do not wait for the factorisation */
        for (int e = lane; e < nx * d; e += 32) sU[(e % nx) + (e / nx) * ldu] = Utp[e];
    __syncwarp();
    warp_potrf(sW, ld, d, sinv, lane);
    if (v.o.regType == 2)
    {
        bool small = false;
        for (int j = lane; j < d; j += 32) small |= (sW[j + j * ld] <= v.o.regTol);
        if (__any_sync(FULL, small))
        {
            __syncwarp();
            for (int e = lane; e < d * d; e += 32)
            {
                const int i = e % d, j = e / d;
                sW[i + j * ld] = Wp[e] + (i == j ? v.o.regValue : 0.0);
            }
            __syncwarp();
            warp_potrf(sW, ld, d, sinv, lane);
            if (lane == 0) atomicAdd(&v.counters[2], 1);
        }
    }
    /* store the factor (lower triangle) and its inverse diagonal */
    for (int e = lane; e < d * d; e += 32)
    {
        const int i = e % d, j = e / d;
        if (i >= j) Wp[e] = sW[i + j * ld];
    }
    for (int e = lane; e < d; e += 32) v.Linv[xb + e] = sinv[e];

    /* forward substitution y = L^-1 resMod (dtrsv_lnn) */
    for (int j = 0; j < d; j++)
    {
        const double yj = sy[j] * sinv[j];
        __syncwarp();
        if (lane == 0) sy[j] = yj;
        for (int i = j + 1 + lane; i < d; i += 32) sy[i] -= sW[i + j * ld] * yj;
        __syncwarp();
    }

    if (dad >= 0)
    {
        /* Utilde = Ut L^-T (dtrsm_rltn): each lane owns one row of Ut */
        for (int c = lane; c < nx; c += 32)
            for (int j = 0; j < d; j++)
            {
                double s = sU[c + j * ldu];
                for (int t = 0; t < j; t++) s -= sU[c + t * ldu] * sW[j + t * ld];
                sU[c + j * ldu] = s * sinv[j];
            }
        __syncwarp();
        for (int e = lane; e < nx * d; e += 32) Utp[e] = sU[(e % nx) + (e / nx) * ldu];
        /* parent's diagonal sub-block -= Utilde Utilde' (dsyrk_ln), rhs -= Utilde y (dgemv_n) */
        const int dd = v.T.dim[dad], off = v.T.pos[g];
        double *Wd = v.W + v.T.woff[dad];
        for (int e = lane; e < nx * nx; e += 32)
        {
            const int a = e % nx, b = e / nx;
            if (a < b) continue;
            double *wd = Wd + (off + a) + (off + b) * dd;
            const double w0 = *wd; /* requested before the dot product */
            double s = 0.0;
            for (int t = 0; t < d; t++) s += sU[a + t * ldu] * sU[b + t * ldu];
            *wd = w0 - s;
        }
        const int xo = v.T.xoff[g];
        for (int a = lane; a < nx; a += 32)
        {
            const double r0 = v.resMod[xo + a];
            double s = 0.0;
            for (int t = 0; t < d; t++) s += sU[a + t * ldu] * sy[t];
            v.resMod[xo + a] = r0 - s;
        }
    }
    else
        warp_trsv_ltn(sW, ld, d, sinv, sy, lane);
    __syncwarp();
    for (int e = lane; e < d; e += 32) v.dlam[xb + e] = sy[e];
    }
}

/* One depth level of the root-to-leaves sweep:
 * dlam_p <- L_p^-T (y_p - Utilde_p' dlam_dad[segment of p])   (dual_Newton_tree.c:764-773) */
template <int DC> /* DC > 0: d == nx == DC for every block of the launch (see k_factor_level) */
__global__ void __launch_bounds__(256) k_backsolve_level(View v, int lvl_begin, int count, int per_warp, int nlev)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (w >= count) return;
    if (v.status[v.T.tree[v.T.lvl_nodes[lvl_begin + w]]] != TQB_RUNNING) return;
#pragma unroll 1
    for (int lev = 0; lev < nlev; lev++)
    {   /* nlev > 1: chain segment walked top-down by one warp (see k_factor_level) */
    const int g = v.T.lvl_nodes[lvl_begin + lev * count + w];
    __syncwarp();
    const int d = DC ? DC : v.T.dim[g], nx = DC ? DC : v.T.nx[g], k0 = v.T.kid0[g];
    const int ld = d | 1;
    double *sW = smem + (size_t) wib * per_warp;
    double *sy = sW + (v.dmax | 1) * v.dmax;
    double *sinv = sy + v.dmax;
    double *sd = sinv + v.dmax;
    const double *Wp = v.W + v.T.woff[g], *Utp = v.Ut + v.T.utoff[g];
    const int xb = v.T.xoff[k0], xo = v.T.xoff[g];
    for (int e = lane; e < d * d; e += 32)
    {
        const int i = e % d, j = e / d;
        if (i >= j) sW[i + j * ld] = Wp[e];
    }
    for (int e = lane; e < d; e += 32) { sinv[e] = v.Linv[xb + e]; }
    for (int e = lane; e < nx; e += 32) sd[e] = v.dlam[xo + e];
    __syncwarp();
    for (int j = lane; j < d; j += 32)
    {
        double s = 0.0;
        for (int c = 0; c < nx; c++) s += Utp[c + j * nx] * sd[c];
        sy[j] = v.dlam[xb + j] - s;
    }
    __syncwarp();
    warp_trsv_ltn(sW, ld, d, sinv, sy, lane);
    __syncwarp();
    for (int e = lane; e < d; e += 32) v.dlam[xb + e] = sy[e];
    }
}

/* ------------------------------------------------------------------------------------------
 * Chain segments with one block dimension (d == nx == DC at every depth of the segment, at most 32 depths), the
 * arithmetic of k_factor_level<DC> / k_backsolve_level<DC> with the global-memory latency taken off the dependent
 * chain of the walk:
 *   - the tables of ALL depths of the chain are read up front, depth L by lane L, and handed out by shuffles;
 *   - while a block is being processed, the inputs of the NEXT block of the chain travel into registers; what the
 *     current block contributes to it (the Schur complement of the factor sweep, the parent's solution of the
 *     substitution sweep) is applied on chip.  Everything is still written to HBM exactly as before (the
 *     regularisation retry, the substitution sweep and the next Newton iteration read it there).
 * Same operations in the same order as the level kernels: same bits.
 * ---------------------------------------------------------------------------------------- */
template <int DC>
__global__ void __launch_bounds__(256) k_factor_chain_pf(View v, int lvl_begin, int count, int per_warp, int nlev)
{
    constexpr int NW = (DC * DC + 31) / 32, ld = DC | 1, ldu = DC | 1;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (w >= count) return;
    if (v.status[v.T.tree[v.T.lvl_nodes[lvl_begin + w]]] != TQB_RUNNING) return;
    double *sW = smem + (size_t) wib * per_warp;
    double *sy = sW + (v.dmax | 1) * v.dmax;
    double *sinv = sy + v.dmax;
    double *sU = sinv + v.dmax;

    /* tables of depth `lane` of the chain */
    int t_w = 0, t_u = 0, t_xb = 0, t_xo = 0;
    if (lane < nlev)
    {
        const int gl = v.T.lvl_nodes[lvl_begin + lane * count + w];
        t_w = v.T.woff[gl]; t_u = v.T.utoff[gl]; t_xb = v.T.xoff[v.T.kid0[gl]]; t_xo = v.T.xoff[gl];
    }
    const int g0 = v.T.lvl_nodes[lvl_begin + w], dad0 = v.T.dad[g0]; /* top of the chain: its parent is outside */

    /* inputs of the deepest block */
    double pW[NW], pU[NW], pr = 0.0;
    {
        const int wo = __shfl_sync(FULL, t_w, nlev - 1), uo = __shfl_sync(FULL, t_u, nlev - 1), xb = __shfl_sync(FULL, t_xb, nlev - 1);
#pragma unroll
        for (int k = 0; k < NW; k++)
        {
            const int e = lane + 32 * k;
            pW[k] = e < DC * DC ? v.W[wo + e] : 0.0;
            pU[k] = e < DC * DC ? v.Ut[uo + e] : 0.0;
        }
        if (lane < DC) pr = v.resMod[xb + lane];
    }
#pragma unroll 1
    for (int lev = nlev - 1; lev >= 0; lev--)
    {
        const int wo = __shfl_sync(FULL, t_w, lev), uo = __shfl_sync(FULL, t_u, lev);
        const int xb = __shfl_sync(FULL, t_xb, lev), xo = __shfl_sync(FULL, t_xo, lev);
        double *Wp = v.W + wo, *Utp = v.Ut + uo;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NW; k++)
        {
            const int e = lane + 32 * k;
            if (e < DC * DC)
            {
                const int i = e % DC, j = e / DC;
                sW[i + j * ld] = (i == j && v.o.regType == 1) ? pW[k] + v.o.regValue : pW[k];
                sU[i + j * ldu] = pU[k];
            }
        }
        if (lane < DC) sy[lane] = pr;
        /* the next block of the chain (the parent): W and the right-hand side before this block's contribution */
        if (lev > 0)
        {
            const int wn = __shfl_sync(FULL, t_w, lev - 1), un = __shfl_sync(FULL, t_u, lev - 1);
#pragma unroll
            for (int k = 0; k < NW; k++)
            {
                const int e = lane + 32 * k;
                pW[k] = e < DC * DC ? v.W[wn + e] : 0.0;
                pU[k] = e < DC * DC ? v.Ut[un + e] : 0.0;
            }
            if (lane < DC) pr = v.resMod[xo + lane];
        }
        __syncwarp();
        warp_potrf32(sW, ld, DC, sinv, lane); /* DC <= 32: one barrier per column, same operations */
        if (v.o.regType == 2)
        {
            bool small = false;
            for (int j = lane; j < DC; j += 32) small |= (sW[j + j * ld] <= v.o.regTol);
            if (__any_sync(FULL, small))
            {
                __syncwarp();
                for (int e = lane; e < DC * DC; e += 32)
                {
                    const int i = e % DC, j = e / DC;
                    sW[i + j * ld] = Wp[e] + (i == j ? v.o.regValue : 0.0);
                }
                __syncwarp();
                warp_potrf32(sW, ld, DC, sinv, lane); /* DC <= 32: one barrier per column, same operations */
                if (lane == 0) atomicAdd(&v.counters[2], 1);
            }
        }
        for (int e = lane; e < DC * DC; e += 32)
        {
            const int i = e % DC, j = e / DC;
            if (i >= j) Wp[e] = sW[i + j * ld];
        }
        for (int e = lane; e < DC; e += 32) v.Linv[xb + e] = sinv[e];
        {   /* forward substitution, one entry per lane (the operations of warp_trsv_lnn) */
            const double yl = warp_trsv_lnn32(sW, ld, DC, lane < DC ? sinv[lane] : 0.0, lane < DC ? sy[lane] : 0.0, lane);
            __syncwarp();
            if (lane < DC) sy[lane] = yl;
            __syncwarp();
        }
        if (lev > 0 || dad0 >= 0)
        {
            for (int c = lane; c < DC; c += 32)
                for (int j = 0; j < DC; j++)
                {
                    double s = sU[c + j * ldu];
                    for (int t = 0; t < j; t++) s -= sU[c + t * ldu] * sW[j + t * ld];
                    sU[c + j * ldu] = s * sinv[j];
                }
            __syncwarp();
            for (int e = lane; e < DC * DC; e += 32) Utp[e] = sU[(e % DC) + (e / DC) * ldu];
            if (lev > 0)
            {   /* parent inside the chain: its block is DC x DC, this kid is its only one */
                double *Wd = v.W + __shfl_sync(FULL, t_w, lev - 1);
#pragma unroll
                for (int k = 0; k < NW; k++)
                {
                    const int e = lane + 32 * k, a = e % DC, b = e / DC;
                    if (e >= DC * DC || a < b) continue;
                    double s = 0.0;
                    for (int t = 0; t < DC; t++) s += sU[a + t * ldu] * sU[b + t * ldu];
                    pW[k] -= s;
                    Wd[e] = pW[k];
                }
                if (lane < DC)
                {
                    double s = 0.0;
                    for (int t = 0; t < DC; t++) s += sU[lane + t * ldu] * sy[t];
                    pr -= s;
                    v.resMod[xo + lane] = pr;
                }
            }
            else
            {   /* parent above the chain: any shape */
                const int dd = v.T.dim[dad0], off = v.T.pos[g0];
                double *Wd = v.W + v.T.woff[dad0];
                for (int e = lane; e < DC * DC; e += 32)
                {
                    const int a = e % DC, b = e / DC;
                    if (a < b) continue;
                    double *wd = Wd + (off + a) + (off + b) * dd;
                    const double w0 = *wd;
                    double s = 0.0;
                    for (int t = 0; t < DC; t++) s += sU[a + t * ldu] * sU[b + t * ldu];
                    *wd = w0 - s;
                }
                for (int a = lane; a < DC; a += 32)
                {
                    const double r0 = v.resMod[xo + a];
                    double s = 0.0;
                    for (int t = 0; t < DC; t++) s += sU[a + t * ldu] * sy[t];
                    v.resMod[xo + a] = r0 - s;
                }
            }
        }
        else
            warp_trsv_ltn(sW, ld, DC, sinv, sy, lane);
        __syncwarp();
        for (int e = lane; e < DC; e += 32) v.dlam[xb + e] = sy[e];
    }
}

template <int DC>
__global__ void __launch_bounds__(256) k_backsolve_chain_pf(View v, int lvl_begin, int count, int per_warp, int nlev)
{
    constexpr int NW = (DC * DC + 31) / 32, ld = DC | 1;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (w >= count) return;
    if (v.status[v.T.tree[v.T.lvl_nodes[lvl_begin + w]]] != TQB_RUNNING) return;
    double *sW = smem + (size_t) wib * per_warp;
    double *sd = sW + (v.dmax | 1) * v.dmax + 2 * v.dmax;

    int t_w = 0, t_u = 0, t_xb = 0;
    if (lane < nlev)
    {
        const int gl = v.T.lvl_nodes[lvl_begin + lane * count + w];
        t_w = v.T.woff[gl]; t_u = v.T.utoff[gl]; t_xb = v.T.xoff[v.T.kid0[gl]];
    }
    /* inputs of the top block: lane j holds column j of Utilde (DC entries), its y and 1 / L_jj */
    double pW[NW], pU[DC], py = 0.0, pinv = 0.0;
    {
        const int wo = __shfl_sync(FULL, t_w, 0), uo = __shfl_sync(FULL, t_u, 0), xb = __shfl_sync(FULL, t_xb, 0);
#pragma unroll
        for (int k = 0; k < NW; k++)
        {
            const int e = lane + 32 * k;
            pW[k] = e < DC * DC ? v.W[wo + e] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < DC; c++) pU[c] = lane < DC ? v.Ut[uo + c + lane * DC] : 0.0;
        if (lane < DC) { py = v.dlam[xb + lane]; pinv = v.Linv[xb + lane]; }
        const int xo = v.T.xoff[v.T.lvl_nodes[lvl_begin + w]];
        if (lane < DC) sd[lane] = v.dlam[xo + lane]; /* the solution of the parent above the chain */
    }
#pragma unroll 1
    for (int lev = 0; lev < nlev; lev++)
    {
        const int xb = __shfl_sync(FULL, t_xb, lev);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NW; k++)
        {
            const int e = lane + 32 * k;
            if (e < DC * DC) sW[(e % DC) + (e / DC) * ld] = pW[k];
        }
        double yv = 0.0;
        const double inv_cur = pinv;
        if (lane < DC)
        {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < DC; c++) s += pU[c] * sd[c];
            yv = py - s;
        }
        if (lev + 1 < nlev)
        {   /* the kid's inputs travel while this block is solved */
            const int wn = __shfl_sync(FULL, t_w, lev + 1), un = __shfl_sync(FULL, t_u, lev + 1), xn = __shfl_sync(FULL, t_xb, lev + 1);
#pragma unroll
            for (int k = 0; k < NW; k++)
            {
                const int e = lane + 32 * k;
                pW[k] = e < DC * DC ? v.W[wn + e] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < DC; c++) pU[c] = lane < DC ? v.Ut[un + c + lane * DC] : 0.0;
            if (lane < DC) { py = v.dlam[xn + lane]; pinv = v.Linv[xn + lane]; }
        }
        __syncwarp();
        const double r = warp_trsv_ltn32(sW, ld, DC, inv_cur, yv, lane); /* one entry per lane, the operations of warp_trsv_ltn */
        if (lane < DC) { v.dlam[xb + lane] = r; sd[lane] = r; }
    }
}

/* lambda += tau_final * dlam once the line search of a tree is over.  (The reference moves
 * lambda incrementally by (tau - tauPrev)*dlam at every trial, dual_Newton_tree.c:964-968; the
 * end point is the same up to rounding.) */
__global__ void k_lambda_step(View v)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= v.Nn) return;
    const int t = v.T.tree[g];
    if (v.status[t] != TQB_RUNNING || v.T.dad[g] < 0) return;
    const double step = v.t_taufin[t];
    const int xo = v.T.xoff[g], nx = v.T.nx[g];
    for (int i = 0; i < nx; i++) v.lam[xo + i] += step * v.dlam[xo + i];
}

/* the same with 32/NXP nodes per warp (NXP >= largest nx, a power of two): unit-stride accesses */
template <int NXP>
__global__ void __launch_bounds__(256) k_lambda_step_w(View v)
{
    constexpr int LPW = 32 / NXP;
    const int lane = threadIdx.x & 31, slot = lane / NXP, i = lane - slot * NXP;
    const int g = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LPW + slot;
    if (g >= v.Nn) return;
    const int t = v.T.tree[g];
    if (v.status[t] != TQB_RUNNING || v.T.dad[g] < 0 || i >= v.T.nx[g]) return;
    const int e = v.T.xoff[g] + i;
    v.lam[e] += v.t_taufin[t] * v.dlam[e];
}

/* the same as a stream over x-space: one CTA per fixed reduction chunk (a node range of ONE tree, so
 * status and step length are CTA-uniform and nothing is looked up per element); roots carry no
 * multiplier and are skipped.  Needs chunks that cover every node (not a sharded solve). */
__global__ void __launch_bounds__(256) k_lambda_step_chunks(View v)
{
    const int c = blockIdx.x, t = v.T.chunk_tree[c];
    if (v.status[t] != TQB_RUNNING) return;
    const double step = v.t_taufin[t];
    const int gb = v.T.chunk_beg[c];
    const int x0 = v.T.chunk_x0[c] + (v.T.dad[gb] < 0 ? v.T.nx[gb] : 0), x1 = v.T.chunk_x1[c];
    for (int e = x0 + threadIdx.x; e < x1; e += 256) v.lam[e] += step * v.dlam[e];
}

/* ------------------------------------------------------------------------------------------
 * Batched line search.  The Armijo sequence tau, beta*tau, ... is known
 * in advance and a trial only streams the affine model (qmod0, dq, ...), so TQB_LS_NT step lengths
 * are evaluated in ONE pass over the vectors and the Armijo loop of dual_Newton_tree.c:960-1000
 * then runs on the device over the batch: no host round trip per trial.
 *   first = 1: slot 0 is f(0), slots 1.. are tau = 1, beta, beta^2, ...; the chunk sums of
 *              res'*dlam (nvb, written by k_stage<2>) are taken in the same pass.
 *   first = 0: slots 0.. are tau, beta*tau, ... for the trees whose line search is still open.
 * Step lengths are generated by repeated multiplication exactly as k_ls_decide does.
 * Bytes per pass: 6(X+U) + 3 Nn doubles read, (NT+1) doubles per chunk written.
 * ---------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) k_ls_eval(View v, int first)
{
    constexpr int NT = TQB_LS_NT;
    __shared__ double sh[NT + 1][256];
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c == 0 && tid == 0) v.counters[1] = 0;
    const int t = v.T.chunk_tree[c];
    if (first ? (v.status[t] != TQB_RUNNING) : !v.ls_active[t]) return;
    double tau[NT], acc[NT];
    {
        double tm = first ? 1.0 : v.t_tau[t];
#pragma unroll
        for (int m = 0; m < NT; m++)
        {
            acc[m] = 0.0;
            if (first && m == 0) { tau[0] = 0.0; continue; }
            tau[m] = tm;
            tm = v.o.beta * tm;
        }
    }
    int dummy;
    for (int i = v.T.chunk_x0[c] + tid; i < v.T.chunk_x1[c]; i += 256)
    {
        const double q0 = v.qmod0[i], dq = v.dq[i], qi = v.in.Qinv[i], lo = v.in.xmin[i], hi = v.in.xmax[i], Qd = v.in.Qd[i];
#pragma unroll
        for (int m = 0; m < NT; m++)
        {
            const double qm = q0 + tau[m] * dq;
            const double val = clip_mask(qi * qm, lo, hi, dummy);
            acc[m] += -0.5 * ((Qd * val) * val) + qm * val;
        }
    }
    for (int i = v.T.chunk_u0[c] + tid; i < v.T.chunk_u1[c]; i += 256)
    {
        const double r0 = v.rmod0[i], dr = v.dr[i], ri = v.in.Rinv[i], lo = v.in.umin[i], hi = v.in.umax[i], Rd = v.in.Rd[i];
#pragma unroll
        for (int m = 0; m < NT; m++)
        {
            const double rm = r0 + tau[m] * dr;
            const double val = clip_mask(ri * rm, lo, hi, dummy);
            acc[m] += -0.5 * ((Rd * val) * val) + rm * val;
        }
    }
    double accd = 0.0;
    for (int g = v.T.chunk_beg[c] + tid; g < v.T.chunk_end[c]; g += 256)
    {
        const double c0 = v.cmod0[g], dc = v.dc[g];
#pragma unroll
        for (int m = 0; m < NT; m++) acc[m] -= c0 + tau[m] * dc;
        if (first) accd += v.nvb[g];
        if (v.has_dense)
        {   /* dual terms of the dense nodes (k_trial_dense, one slot per step length) */
#pragma unroll
            for (int m = 0; m < NT; m++) acc[m] += v.fdense[(size_t) m * v.Nn + g];
        }
    }
#pragma unroll
    for (int m = 0; m < NT; m++) sh[m][tid] = acc[m];
    sh[NT][tid] = accd;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if (tid < s)
        {
#pragma unroll
            for (int m = 0; m <= NT; m++) sh[m][tid] += sh[m][tid + s];
        }
        __syncthreads();
    }
    if (tid < NT) v.cvb[(size_t) tid * v.nchunk + c] = sh[tid][0];
    if (first && tid == NT) v.cva[c] = sh[NT][0];
}

/* descent test (first) + Armijo loop over the batch of k_ls_eval (dual_Newton_tree.c:940-1000).
 * One CTA per tree, warp m < NT reduces the chunk values of step length m, warp NT those of
 * res'*dlam (same additions as tree_reduce, four independent loads in flight per lane); thread 0
 * then walks the Armijo sequence. */
__global__ void __launch_bounds__(32 * (TQB_LS_NT + 1)) k_ls_decide(View v, int first, int nt)
{
    constexpr int NT = TQB_LS_NT;
    __shared__ double sf[NT + 1];
    const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
    const int t = blockIdx.x;
    if (first ? (v.status[t] != TQB_RUNNING) : !v.ls_active[t]) return;
    if (v.sharded)
    {
        if (lane == 0) sf[m] = m < NT ? v.t_red[(size_t) (1 + m) * v.nT + t] : (first ? v.t_red[t] : 0.0);
    }
    else if (m < NT || first)
    {
        const double *cv = m < NT ? v.cvb + (size_t) m * v.nchunk : v.cva;
        const int c1 = v.T.tchunk0[t + 1];
        int c = v.T.tchunk0[t] + lane;
        double acc = 0.0;
        for (; c + 96 < c1; c += 128)
        {
            const double a0 = cv[c], a1 = cv[c + 32], a2 = cv[c + 64], a3 = cv[c + 96];
            acc += a0; acc += a1; acc += a2; acc += a3;
        }
        for (; c < c1; c += 32) acc += cv[c];
        acc = warp_sum(acc);
        if (lane == 0) sf[m] = acc;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double f[NT];
#pragma unroll
    for (int k = 0; k < NT; k++) f[k] = sf[k];
    int m0 = 0;
    if (first)
    {
        const double dot = -sf[NT];
        v.t_dot[t] = dot;
        v.t_f0[t] = f[0];
        if (dot > 1e-10 || !((dot > 1e-10) || (dot < 1e-10)))
        {
            v.status[t] = 2; /* TREEQP_DN_NOT_DESCENT_DIRECTION */
            v.ls_active[t] = 0;
            return;
        }
        v.t_tau[t] = 1.0;
        v.t_tauprev[t] = 0.0;
        v.ls_cnt[t] = 0;
        v.ls_active[t] = 1;
        m0 = 1;
    }
    const double f0 = v.t_f0[t], tdot = v.t_dot[t];
    const int restart = v.restart_cnt[t];
    double tau = v.t_tau[t], taufin = tau, tauprev = v.t_tauprev[t];
    int ls = v.ls_cnt[t];
    bool done = false;
    int recorded = 0;
#pragma unroll
    for (int k = 0; k < NT; k++)
    {
        if (k < m0 || k >= nt || done) continue;
        ls++;
        recorded = ls;
        taufin = tau; /* lambda of the last evaluated trial */
        if (restart == v.o.lsRestartTrigger) done = true; /* full step when stuck (:973-977) */
        else if (f[k] <= f0 + v.o.gamma * tau * tdot) done = true;
        else
        {
            tauprev = tau;
            tau = v.o.beta * tau;
            if (ls >= v.o.lsMaxIter) { done = true; recorded = v.o.lsMaxIter + 1; }
        }
    }
    v.t_taufin[t] = taufin;
    v.t_tauprev[t] = tauprev;
    v.t_tau[t] = tau;
    v.ls_cnt[t] = ls;
    if (done)
    {
        v.ls_active[t] = 0;
        v.restart_cnt[t] = (recorded >= v.o.lsMaxIter) ? restart + 1 : 0; /* :993-1000 */
        const int it = v.iter[t]++;
        if (t == 0 && it < 256) v.ls_hist[it] = recorded;
        atomicAdd(&v.counters[3], ls);
    }
    else
        atomicAdd(&v.counters[1], 1);
}

/* sharded solve, batched line search: this rank's parts of res'*dlam (slot 0) and of the NT dual
 * function values (slots 1..NT) per tree; grid.y = slot */
__global__ void k_tree_partial_ls(View v, double *out)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= v.nT) return;
    const int y = blockIdx.y;
    const double *cv = y == 0 ? v.cva : v.cvb + (size_t) (y - 1) * v.nchunk;
    const double r = tree_reduce(v, cv, t, lane, false);
    if (lane == 0) out[(size_t) y * v.nT + t] = r;
}

/* trees still running after maxIter iterations (dual_Newton_tree.c:1252) */
__global__ void k_finalize_status(View v)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < v.nT && v.status[t] == TQB_RUNNING) v.status[t] = 1;
}

/* bound multipliers mu = Q.*(xUnc - x) (clipping.c:386-399) */
__global__ void k_finish(View v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.X) v.mux[i] = v.in.Qd[i] * (v.xUnc[i] - v.x[i]);
    if (i < v.U) v.muu[i] = v.in.Rd[i] * (v.uUnc[i] - v.u[i]);
}


/* Largest KKT residual of every tree from the iterates in HBM, with the definition of
 * tree_qp_out_max_KKT_res (tree_qp_common.c:540-788; host restatement in containers.c): infinity norm
 * over stationarity (Q x + S'u + q + mu_x - lam + sum_kids A_k' lam_k ; R u + S x + r + mu_u +
 * sum_kids B_k' lam_k), dynamics (A x_dad + B u_dad + b - x), bound violation and complementarity
 * (mu > 0 <=> upper bound).  One warp per node; the per-tree maximum is taken on the bit patterns of
 * the absolute values (order preserving for non-negative doubles, NaN sorts above everything, so
 * it propagates like in the host version). */
__device__ __forceinline__ unsigned long long kkt_bits(double v) { return (unsigned long long) __double_as_longlong(fabs(v)); }

__global__ void __launch_bounds__(256) k_kkt(View v, unsigned long long *out)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= v.Nn) return;
    const int nx = v.T.nx[g], nu = v.T.nu[g], xo = v.T.xoff[g], uo = v.T.uoff[g];
    const int dad = v.T.dad[g], k0 = v.T.kid0[g], nk = v.T.nkids[g], ho = v.T.hoff[g], nv = nx + nu;
    const double *x = v.x + xo, *u = v.u + uo;
    const double *H = ho >= 0 ? v.in.H + ho : nullptr;
    const int mrows = (v.ncmax > 0 && ho >= 0) ? v.T.ncn[g] : 0;
    const double *Cg = mrows > 0 ? v.in.Cg + v.T.cgoff[g] : nullptr, *md = mrows > 0 ? v.mud + v.T.doff[g] : nullptr;
    unsigned long long err = 0ull;
    for (int r = lane; r < nv; r += 32)
    {   /* stationarity, row r of z = [x; u] */
        const bool isx = r < nx;
        const int i = isx ? r : r - nx;
        double s = isx ? v.in.q[xo + i] + v.mux[xo + i] : v.in.r[uo + i] + v.muu[uo + i];
        if (H)
        {
            for (int c = 0; c < nx; c++) s += H[r + c * nv] * x[c];
            for (int c = 0; c < nu; c++) s += H[r + (nx + c) * nv] * u[c];
        }
        else s += isx ? v.in.Qd[xo + i] * x[i] : v.in.Rd[uo + i] * u[i];
        if (isx && dad >= 0) s -= v.lam[xo + i];
        if (mrows > 0)
            for (int a = 0; a < mrows; a++) s += Cg[a + r * mrows] * md[a]; /* + C' mu_d / D' mu_d */
        for (int j = 0; j < nk; j++)
        {
            const int k = k0 + j, nxk = v.T.nx[k];
            const double *M = isx ? v.in.A + v.T.aoff[k] + (size_t) i * nxk : v.in.B + v.T.boff[k] + (size_t) i * nxk;
            const double *lk = v.lam + v.T.xoff[k];
            for (int c = 0; c < nxk; c++) s += M[c] * lk[c];
        }
        err = max(err, kkt_bits(s));
        /* bounds and complementarity of the same variable */
        double lo, hi;
        if (H) { lo = isx ? v.in.dxmin[xo + i] : v.in.dumin[uo + i]; hi = isx ? v.in.dxmax[xo + i] : v.in.dumax[uo + i]; }
        else { lo = isx ? v.in.xmin[xo + i] : v.in.umin[uo + i]; hi = isx ? v.in.xmax[xo + i] : v.in.umax[uo + i]; }
        const double z = isx ? x[i] : u[i], m = isx ? v.mux[xo + i] : v.muu[uo + i];
        if (z > hi) err = max(err, kkt_bits(z - hi));
        else if (z < lo) err = max(err, kkt_bits(lo - z));
        err = max(err, kkt_bits(m > 0 ? m * (z - hi) : m * (lo - z)));
    }
    for (int a = lane; a < mrows; a += 32)
    {   /* general constraints: feasibility and complementarity (tree_qp_common.c:723-755) */
        double gv = 0.0;
        for (int c = 0; c < nx; c++) gv += Cg[a + c * mrows] * x[c];
        for (int c = 0; c < nu; c++) gv += Cg[a + (nx + c) * mrows] * u[c];
        const double lo = v.in.dmin[v.T.doff[g] + a], hi = v.in.dmax[v.T.doff[g] + a], m = md[a];
        if (gv > hi) err = max(err, kkt_bits(gv - hi));
        else if (gv < lo) err = max(err, kkt_bits(lo - gv));
        err = max(err, kkt_bits(m > 0 ? m * (gv - hi) : m * (lo - gv)));
    }
    if (dad >= 0)
    {   /* dynamics of the edge into this node */
        const int nxp = v.T.nx[dad], nup = v.T.nu[dad];
        const double *xp = v.x + v.T.xoff[dad], *up = v.u + v.T.uoff[dad];
        const double *A = v.in.A + v.T.aoff[g], *B = v.in.B + v.T.boff[g];
        for (int r = lane; r < nx; r += 32)
        {
            double s = v.in.b[xo + r] - x[r];
            for (int c = 0; c < nxp; c++) s += A[r + c * nx] * xp[c];
            for (int c = 0; c < nup; c++) s += B[r + c * nx] * up[c];
            err = max(err, kkt_bits(s));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err = max(err, __shfl_xor_sync(FULL, err, o));
    if (lane == 0) atomicMax(out + v.T.tree[g], err);
}
