/*
 * tdunes_small.cuh -- the whole dual Newton loop of treeqp_tdunes_solve (dual_Newton_tree.c:1166-1228) for
 * SMALL uniform trees as ONE persistent kernel: one CTA per tree of the forest, block barriers instead of kernel
 * launches between the phases, every decision (termination, descent test, Armijo) taken inside the CTA.
 *
 * Why: a small tree (config 1: 85 nodes, config 4: 175 nodes per QP) is a chain of dependent block operations --
 * 2 x 19 levels per Newton iteration for Nh = 20 -- and the level-synchronous driver pays a kernel boundary
 * (>= 4 us) plus one host round trip per iteration for work that takes a warp 1-2 us.  Here a level costs a
 * __syncthreads(), an iteration never leaves the SM, the tree's data stays in L1/L2 after the first iteration,
 * and the trees of a batch run (and finish) independently of each other.
 *
 * The kernel is the driver loop of device.cu re-expressed with the SAME device functions (stage_parent /
 * stage_leaves of tdunes_stage.cuh, bf_block / bf_block_lb / bs_block of tdunes_fused.cuh), so a block is built,
 * factored and substituted with the same arithmetic in the same order; only the reductions (termination norm,
 * res'dlam, the dual function values of a line-search pass) have their own fixed shape: thread-strided partial
 * sums over the tree, a warp butterfly, then the warps in order.
 *
 * Preconditions (checked by tqb_device when it picks this path): uniform forest on the fused path (every tree the
 * same shape), clipping stage QPs, no sharding, (NX, NU) compiled in, at most TQB_SMALL_MAXH depths.
 */
#pragma once
#include "tdunes_stage.cuh"

#define TQB_SMALL_MAXH 64
#define TQB_SMALL_MAXWARPS 16 /* the launch picks 4 (batches: four CTAs per SM) .. 16 (a single tree) warps */

struct SmallPlan
{
    int Nh;                            /* parents live at depths 0 .. Nh-1 */
    int nroots_total;                  /* = nT (one root per tree) */
    int lvl_begin[TQB_SMALL_MAXH];     /* first descriptor of depth l in the forest-wide level order */
    short lvl_cnt[TQB_SMALL_MAXH];     /* parents of ONE tree at depth l */
    signed char lvl_md[TQB_SMALL_MAXH], lvl_kp[TQB_SMALL_MAXH];
    int nleaf;                         /* leaves of ONE tree */
    int nodes, xlen, ulen;             /* nodes / x-space / u-space elements of ONE tree */
    int pnx;                           /* state dimension of the roots (NX or 0) */
    int smem_warp;                     /* doubles of block scratch per warp */
};

/* Deterministic CTA reductions whose shape does not depend on the number of warps the launch picked (a QP solved
 * alone and inside a batch gives the same bits): the values are dealt to TQB_SMALL_VL virtual lanes (element i to
 * lane i mod TQB_SMALL_VL), a warp works through whole virtual warps, butterfly inside a virtual warp, then the
 * virtual warps in order.  blockDim.x must divide TQB_SMALL_VL (4, 8 or 16 warps). */
#define TQB_SMALL_VL 512
#define TQB_SMALL_VW (TQB_SMALL_VL / 32)

/* per-node values src[g0 .. g0+n) -> every thread gets the sum / maximum */
__device__ __forceinline__ double small_reduce_nodes(const double *src, int g0, int n, bool is_max, double *s_part, int tid, int T)
{
    __syncthreads(); /* s_part may still be read */
    for (int vb = 0; vb < TQB_SMALL_VL; vb += T)
    {
        const int vl = vb + tid;
        double part = 0.0;
        for (int g = g0 + vl; g < g0 + n; g += TQB_SMALL_VL) part = is_max ? fmax(part, src[g]) : part + src[g];
        part = is_max ? warp_max(part) : warp_sum(part);
        if ((tid & 31) == 0) s_part[vl >> 5] = part;
    }
    __syncthreads();
    double r = s_part[0];
#pragma unroll
    for (int w = 1; w < TQB_SMALL_VW; w++) r = is_max ? fmax(r, s_part[w]) : r + s_part[w];
    return r;
}

template <int NX, int NU, int MD, int PNX, bool ROOT>
__device__ __forceinline__ void small_bf(const View &v, size_t desc, int kp, double *Ms, int lane)
{
    if constexpr (BfSel<NX, NU, MD, PNX, ROOT>::BLOCKED) bf_block_lb<NX, NU, MD, PNX, ROOT>(v, desc, kp, false, Ms, lane);
    else bf_block<NX, NU, MD, PNX, ROOT>(v, desc, kp, false, Ms, lane);
}

template <int NX, int NU, int MD>
__device__ __forceinline__ void small_bf_md(const View &v, size_t desc, int kp, bool root, int pnx, double *Ms, int lane)
{
    if (!root) small_bf<NX, NU, MD, NX, false>(v, desc, kp, Ms, lane);
    else if (pnx == 0) small_bf<NX, NU, MD, 0, true>(v, desc, kp, Ms, lane);
    else small_bf<NX, NU, MD, NX, true>(v, desc, kp, Ms, lane);
}

/* MDMAX = largest number of kids per parent compiled in for this (NX, NU) (TQB_FUSED_LIST) */
template <int NX, int NU, int MDMAX>
__device__ __forceinline__ void small_factor(const View &v, int md, size_t desc, int kp, bool root, int pnx, double *Ms, int lane)
{
    if (md == 1) small_bf_md<NX, NU, 1>(v, desc, kp, root, pnx, Ms, lane);
    else if (MDMAX >= 2 && md == 2) small_bf_md<NX, NU, (MDMAX >= 2 ? 2 : 1)>(v, desc, kp, root, pnx, Ms, lane);
    else if (MDMAX >= 3 && md == 3) small_bf_md<NX, NU, (MDMAX >= 3 ? 3 : 1)>(v, desc, kp, root, pnx, Ms, lane);
    else if (MDMAX >= 4 && md == 4) small_bf_md<NX, NU, (MDMAX >= 4 ? 4 : 1)>(v, desc, kp, root, pnx, Ms, lane);
}

template <int NX, int MDMAX>
__device__ __forceinline__ void small_backsolve(const View &v, int md, size_t desc, int lane)
{
    if (md == 1) bs_block<NX, 1>(v, desc, false, lane);
    else if (MDMAX >= 2 && md == 2) bs_block<NX, (MDMAX >= 2 ? 2 : 1)>(v, desc, false, lane);
    else if (MDMAX >= 3 && md == 3) bs_block<NX, (MDMAX >= 3 ? 3 : 1)>(v, desc, false, lane);
    else if (MDMAX >= 4 && md == 4) bs_block<NX, (MDMAX >= 4 ? 4 : 1)>(v, desc, false, lane);
}

/* stage QPs / direction model of every node of tree t (MODE 0 / 2 / 3 of k_stage_fused) */
template <int NX, int NU, int MODE>
__device__ __forceinline__ void small_stage(const View &v, const SmallPlan &pl, int t, double *sL, double *sS, int lane, int wib, int nw)
{
    constexpr int NXP = tqb_pow2_ge(NX), LPW = 32 / NXP;
    const int4 *descs = reinterpret_cast<const int4 *>(v.lvl_desc);
    for (int l = 0; l < pl.Nh; l++)
    {
        const int cnt = pl.lvl_cnt[l];
        for (int j = wib; j < cnt; j += nw)
        {
            const size_t e = 3 * ((size_t) pl.lvl_begin[l] + (size_t) t * cnt + j);
            const int4 d0 = __ldg(descs + e), d1 = __ldg(descs + e + 1), d2 = __ldg(descs + e + 2);
            stage_parent<NX, NU, MODE>(v, d0, d1, d2, l == 0, pl.pnx, sL, sS, lane);
        }
    }
    const int slot = lane / NXP, li = lane - slot * NXP;
    const int ngrp = (pl.nleaf + LPW - 1) / LPW;
    for (int grp = wib; grp < ngrp; grp += nw)
    {
        const int lf = grp * LPW + slot;
        int4 ld = make_int4(0, 0, 0, 0);
        if (lf < pl.nleaf) ld = __ldg(reinterpret_cast<const int4 *>(v.leaf_desc) + (size_t) t * pl.nleaf + lf);
        stage_leaves<NX, MODE>(v, ld, lf < pl.nleaf, li);
    }
}

template <int NX, int NU, int MDMAX>
__global__ void __launch_bounds__(32 * TQB_SMALL_MAXWARPS, 1) k_tdunes_small(View v, SmallPlan pl)
{
    constexpr int NT = TQB_LS_NT;
    extern __shared__ double smem[];
    __shared__ double s_red[TQB_SMALL_VW];
    __shared__ double s_fp[TQB_SMALL_VW][TQB_LS_NT];
    __shared__ double s_f[NT + 1];
    __shared__ double sLs[TQB_SMALL_MAXWARPS][32];
    __shared__ double sSs[TQB_SMALL_MAXWARPS][TQB_STAGE_MAXPAIR];
    __shared__ int s_ctl[4]; /* 0: leave the Newton loop, 1: line search still open */
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5, tid = threadIdx.x, T = blockDim.x;
    const int t = blockIdx.x;
    double *Ms = smem + (size_t) wib * pl.smem_warp;
    const int g0 = t * pl.nodes, x0 = t * pl.xlen, u0 = t * pl.ulen;
    const int xs = x0 + pl.pnx; /* the root carries no multiplier */
    const bool is_max = v.o.termCondition == 2;

    for (int it = 0; it < v.o.maxIter; it++)
    {
        small_stage<NX, NU, 0>(v, pl, t, sLs[wib], sSs[wib], lane, wib, nw);
        __syncthreads();
        /* gradient + Hessian blocks + factorisation, leaves to root */
        for (int l = pl.Nh - 1; l >= 0; l--)
        {
            const int cnt = pl.lvl_cnt[l], md = pl.lvl_md[l], kp = pl.lvl_kp[l];
            for (int j = wib; j < cnt; j += nw)
            {
                small_factor<NX, NU, MDMAX>(v, md, (size_t) pl.lvl_begin[l] + (size_t) t * cnt + j, kp, l == 0, pl.pnx, Ms, lane);
                __syncwarp();
            }
            __syncthreads();
        }
        /* termination test (dual_Newton_tree.c:542-546) */
        double err = small_reduce_nodes(v.nva, g0, pl.nodes, is_max, s_red, tid, T);
        if (v.o.termCondition == 1) err = sqrt(err);
        if (tid == 0)
        {
            v.t_err[t] = err;
            s_ctl[0] = 0;
            if (err < v.o.tol) { v.status[t] = 0; s_ctl[0] = 1; }
        }
        __syncthreads();
        if (s_ctl[0]) break;
        /* substitution, root to leaves */
        for (int l = 1; l < pl.Nh; l++)
        {
            const int cnt = pl.lvl_cnt[l], md = pl.lvl_md[l];
            for (int j = wib; j < cnt; j += nw)
            {
                small_backsolve<NX, MDMAX>(v, md, (size_t) pl.lvl_begin[l] + (size_t) t * cnt + j, lane);
                __syncwarp();
            }
            __syncthreads();
        }
        /* affine model of the stage gradients along dlam, res'dlam */
        small_stage<NX, NU, 2>(v, pl, t, sLs[wib], sSs[wib], lane, wib, nw);
        __syncthreads();
        const double dot = -small_reduce_nodes(v.nvb, g0, pl.nodes, false, s_red, tid, T);
        /* batched line search: NT step lengths per pass (k_ls_eval + k_ls_decide) */
        double f0 = 0.0, tau = 1.0, tauprev = 0.0, taufin = 1.0;
        int ls = 0, recorded = 0;
        bool done = false, first = true, bad = false;
        const int restart = v.restart_cnt[t];
        while (!done)
        {
            double taus[NT], acc[NT];
            {
                double tm = tau;
#pragma unroll
                for (int m = 0; m < NT; m++)
                {
                    acc[m] = 0.0;
                    if (first && m == 0) { taus[0] = 0.0; continue; }
                    taus[m] = tm;
                    tm = v.o.beta * tm;
                }
            }
            __syncthreads(); /* s_fp of the previous pass has been read */
            for (int vb = 0; vb < TQB_SMALL_VL; vb += T)
            {
                const int vl = vb + tid;
                int dummy;
#pragma unroll
                for (int m = 0; m < NT; m++) acc[m] = 0.0;
                for (int i = x0 + vl; i < x0 + pl.xlen; i += TQB_SMALL_VL)
                {
                    const double q0 = v.qmod0[i], dq = v.dq[i], qi = v.in.Qinv[i], lo = v.in.xmin[i], hi = v.in.xmax[i], Qd = v.in.Qd[i];
#pragma unroll
                    for (int m = 0; m < NT; m++)
                    {
                        const double qm = q0 + taus[m] * dq;
                        const double val = clip_mask(qi * qm, lo, hi, dummy);
                        acc[m] += -0.5 * ((Qd * val) * val) + qm * val;
                    }
                }
                for (int i = u0 + vl; i < u0 + pl.ulen; i += TQB_SMALL_VL)
                {
                    const double r0 = v.rmod0[i], dr = v.dr[i], ri = v.in.Rinv[i], lo = v.in.umin[i], hi = v.in.umax[i], Rd = v.in.Rd[i];
#pragma unroll
                    for (int m = 0; m < NT; m++)
                    {
                        const double rm = r0 + taus[m] * dr;
                        const double val = clip_mask(ri * rm, lo, hi, dummy);
                        acc[m] += -0.5 * ((Rd * val) * val) + rm * val;
                    }
                }
                for (int g = g0 + vl; g < g0 + pl.nodes; g += TQB_SMALL_VL)
                {
                    const double c0 = v.cmod0[g], dc = v.dc[g];
#pragma unroll
                    for (int m = 0; m < NT; m++) acc[m] -= c0 + taus[m] * dc;
                }
#pragma unroll
                for (int m = 0; m < NT; m++)
                {
                    const double r = warp_sum(acc[m]);
                    if (lane == 0) s_fp[vl >> 5][m] = r;
                }
            }
            __syncthreads();
            if (tid < NT)
            {
                double r = s_fp[0][tid];
#pragma unroll
                for (int w = 1; w < TQB_SMALL_VW; w++) r += s_fp[w][tid];
                s_f[tid] = r;
            }
            __syncthreads();
            /* every thread walks the Armijo sequence on the same numbers (dual_Newton_tree.c:940-1000) */
            int m0 = 0;
            if (first)
            {
                f0 = s_f[0];
                if (dot > 1e-10 || !((dot > 1e-10) || (dot < 1e-10))) { bad = true; break; }
                m0 = 1;
            }
#pragma unroll
            for (int k = 0; k < NT; k++)
            {
                if (k < m0 || done) continue;
                ls++;
                recorded = ls;
                taufin = tau;
                if (restart == v.o.lsRestartTrigger) done = true;
                else if (s_f[k] <= f0 + v.o.gamma * tau * dot) done = true;
                else
                {
                    tauprev = tau;
                    tau = v.o.beta * tau;
                    if (ls >= v.o.lsMaxIter) { done = true; recorded = v.o.lsMaxIter + 1; }
                }
            }
            first = false;
            __syncthreads();
        }
        if (bad)
        {   /* TREEQP_DN_NOT_DESCENT_DIRECTION: lambda stays where it is */
            if (tid == 0) { v.status[t] = 2; v.t_dot[t] = dot; v.t_f0[t] = f0; }
            __syncthreads();
            break;
        }
        /* lambda += tau_final * dlam */
        for (int i = xs + tid; i < x0 + pl.xlen; i += T) v.lam[i] += taufin * v.dlam[i];
        if (tid == 0)
        {
            v.t_dot[t] = dot; v.t_f0[t] = f0; v.t_tau[t] = tau; v.t_tauprev[t] = tauprev; v.t_taufin[t] = taufin;
            v.ls_cnt[t] = ls;
            v.restart_cnt[t] = (recorded >= v.o.lsMaxIter) ? restart + 1 : 0;
            const int itn = v.iter[t]++;
            if (t == 0 && itn < 256) v.ls_hist[itn] = recorded;
            atomicAdd(&v.counters[3], ls);
        }
        __syncthreads();
    }
}
