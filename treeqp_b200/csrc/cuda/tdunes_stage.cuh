/*
 * tdunes_stage.cuh -- compile-time-dimension stage-QP kernel for uniform trees (the same class of
 * trees as tdunes_fused.cuh: every non-root node NX states, every parent NU inputs, clipping
 * stage QPs, leaves without inputs).
 *
 * k_stage_fused<NX,NU,MODE> replaces k_stage<MODE> (tdunes_kernels.cuh) on that class, with the
 * same arithmetic in the same order (the results are bit-identical):
 *   MODE 0  solve_stage_problems (dual_Newton_tree.c:218) + stage_qp_clipping_solve_extended
 *           (dual_Newton_tree_clipping.c:188): qmod, rmod, x, u, xUnc, uUnc, P, Rt, cmod
 *   MODE 2  the same operator applied to dlam: dq, dr, dc and the node's part of res'*dlam
 *           (gradient_trans_times_direction, dual_Newton_tree.c:808)
 *   MODE 3  clip-only solve (stage_qp_clipping_solve, clipping.c:231)
 *
 * Work split.  The generic kernel stages [A_kids B_kids] in shared memory and then walks its
 * columns with a stride of nx_kid doubles (an 8-way bank conflict for nx = 12) and spends a whole
 * warp on every leaf.  Here
 *   - a PARENT warp gives each lane one (kid, column) pair of C = [A_kids B_kids]: the column is
 *     NX contiguous doubles in HBM, read with 16-byte loads straight into registers and folded
 *     with the kid's multiplier by a sequential fma chain (dgemv_t order); the MD partial products
 *     of a column are then combined kid by kid by the lane that owns the column's variable;
 *   - a LEAF warp handles 32/NXP leaves at once (NXP = NX rounded up to a power of two): leaves
 *     are a pure element-wise clip of -q + lambda.
 * Algorithmic bytes are those of k_stage; nothing is staged, nothing is re-read.
 */
#pragma once
#include "tdunes_fused.cuh"

#define TQB_STAGE_WPB 8       /* warps per CTA */
#ifndef TQB_STAGE_MINBLOCKS
#define TQB_STAGE_MINBLOCKS 4 /* resident CTAs per SM asked of the register allocator */
#endif
#define TQB_STAGE_MAXPAIR 64  /* max kids * (NX + NU) per parent */

struct LeafDesc { int g, xo, tree, root; };

__host__ __device__ constexpr int tqb_pow2_ge(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : n <= 16 ? 16 : 32; }

/* one parent: d0,d1,d2 = its LvlDesc record (already in registers) */
template <int NX, int NU, int MODE>
__device__ __forceinline__ void stage_parent(const View &v, const int4 d0, const int4 d1, const int4 d2, bool root, int root_nx,
                                             double *sL, double *sS, int lane)
{
    const double *src = (MODE == 2) ? v.dlam : v.lam;
    const int g = d0.x, xo = d0.z, uo = d0.w, xb = d1.x, md = d2.z;
    const int running = v.status[d2.y] == TQB_RUNNING; /* tested before the first store */
    const int pnx = root ? root_nx : NX, ncol = pnx + NU, D = md * NX, npair = md * ncol;

    /* inputs of the epilogue first (independent of everything below) */
    const bool own = lane < ncol, isx = lane < pnx;
    const int i = isx ? lane : lane - pnx;
    double base = 0.0, winv = 0.0, lo = 0.0, hi = 0.0, rs = 0.0, pold = 0.0;
    if (own)
    {
        if (isx)
        {
            const double l = root ? 0.0 : src[xo + i];
            if (MODE == 2) { base = l; rs = root ? 0.0 : v.res[xo + i]; }
            else
            {
                base = -v.in.q[xo + i] + l;
                winv = v.in.Qinv[xo + i]; lo = v.in.xmin[xo + i]; hi = v.in.xmax[xo + i];
                if (MODE == 0) pold = v.P[xo + i];
            }
        }
        else if (MODE != 2)
        {
            base = -v.in.r[uo + i];
            winv = v.in.Rinv[uo + i]; lo = v.in.umin[uo + i]; hi = v.in.umax[uo + i];
            if (MODE == 0) pold = v.Rt[uo + i];
        }
    }
    /* the kids' multipliers (one contiguous range) and cmod = sum_kids b_j' lambda_j (:892) */
    double cpart = 0.0, lkid = 0.0;
    if (lane < D)
    {
        lkid = src[xb + lane];
        if (MODE != 3) cpart = v.in.b[xb + lane] * lkid;
    }
    /* first (kid, column) pair of this lane: NX contiguous doubles straight into registers */
    double cr[NX];
    int p = lane;
    auto load_col = [&](int pp) {
        const int kid = pp / ncol, c = pp - kid * ncol;
        const double *col = c < pnx ? v.in.A + d1.y + (size_t) (kid * pnx + c) * NX
                                    : v.in.B + d1.z + (size_t) (kid * NU + (c - pnx)) * NX;
        if (NX % 2 == 0)
        {
            const double2 *c2 = reinterpret_cast<const double2 *>(col);
#pragma unroll
            for (int r = 0; r < NX / 2; r++) { const double2 t = __ldg(c2 + r); cr[2 * r] = t.x; cr[2 * r + 1] = t.y; }
        }
        else
        {
#pragma unroll
            for (int r = 0; r < NX; r++) cr[r] = __ldg(col + r);
        }
        return kid;
    };
    int kid = 0;
    if (p < npair) kid = load_col(p);
    if (!running) return; /* warp-uniform */
    if (lane < D) sL[lane] = lkid;
    __syncwarp();
    if (MODE != 3)
    {
        const double c = warp_sum(cpart);
        if (lane == 0) { if (MODE == 0) v.cmod0[g] = c; else v.dc[g] = c; }
    }
    while (p < npair)
    {   /* s = column' * lambda_kid, sequential (dgemv_t order) */
        const double *lk = sL + kid * NX;
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < NX; r++) s += cr[r] * lk[r];
        sS[p] = s;
        p += 32;
        if (p < npair) kid = load_col(p);
    }
    __syncwarp();
    bool changed = false;
    if (own)
    {
        double acc = base;
        for (int j = 0; j < md; j++) acc -= sS[j * ncol + lane];
        if (MODE == 2)
        {
            if (isx) v.dq[xo + i] = acc; else v.dr[uo + i] = acc;
        }
        else
        {
            int mask;
            const double unc = winv * acc;
            const double val = clip_mask(unc, lo, hi, mask);
            const double pnew = mask == 0 ? winv : 0.0;
            changed = pnew != pold;
            if (isx)
            {
                v.x[xo + i] = val;
                if (MODE == 0) { v.qmod0[xo + i] = acc; v.xUnc[xo + i] = unc; v.P[xo + i] = pnew; }
            }
            else
            {
                v.u[uo + i] = val;
                if (MODE == 0) { v.rmod0[uo + i] = acc; v.uUnc[uo + i] = unc; v.Rt[uo + i] = pnew; }
            }
        }
    }
    if (MODE == 0)
    {   /* did the active set of this node move since the last Newton iteration? (k_build_factor re-uses the
         * factorisation of blocks whose whole subtree did not) */
        const bool any = __any_sync(FULL, changed);
        if (lane == 0 && v.chg) v.chg[g] = any;
    }
    if (MODE == 2)
    {   /* node's part of res'*dlam */
        const double dotp = warp_sum((own && isx) ? rs * base : 0.0);
        if (lane == 0) v.nvb[g] = dotp;
    }
    __syncwarp(); /* sL / sS are reused by the warp's next item */
}

/* Two chain parents per warp (one kid each, NX + NU <= 16): lanes 0-15 work on one parent, lanes
 * 16-31 on the next one.  A chain parent keeps only NX + NU of a warp's 32 lanes busy; pairing
 * doubles the bytes a warp has in flight, which is what bounds this kernel.  d0,d1,d2 = the
 * descriptor of THIS lane's parent (valid = it exists).  Same arithmetic as stage_parent with
 * md = 1: the 16-lane butterfly performs the additions of the 32-lane one padded with zeros. */
__device__ __forceinline__ double half_sum(double v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <int NX, int NU, int MODE>
__device__ __forceinline__ void stage_parent_pair(const View &v, const int4 d0, const int4 d1, const int4 d2, bool valid,
                                                  double *sL, int lane)
{
    static_assert(NX + NU <= 16, "a chain parent must fit 16 lanes");
    constexpr int ncol = NX + NU;
    const double *src = (MODE == 2) ? v.dlam : v.lam;
    const int sl = lane & 15, grp = lane >> 4;
    const int g = d0.x, xo = d0.z, uo = d0.w, xb = d1.x;
    const bool act = valid && v.status[d2.y] == TQB_RUNNING;
    const bool own = valid && sl < ncol, isx = sl < NX;
    const int i = isx ? sl : sl - NX;
    double base = 0.0, winv = 0.0, lo = 0.0, hi = 0.0, rs = 0.0;
    if (own)
    {
        if (isx)
        {
            const double l = src[xo + i];
            if (MODE == 2) { base = l; rs = v.res[xo + i]; }
            else
            {
                base = -v.in.q[xo + i] + l;
                winv = v.in.Qinv[xo + i]; lo = v.in.xmin[xo + i]; hi = v.in.xmax[xo + i];
            }
        }
        else if (MODE != 2)
        {
            base = -v.in.r[uo + i];
            winv = v.in.Rinv[uo + i]; lo = v.in.umin[uo + i]; hi = v.in.umax[uo + i];
        }
    }
    double cpart = 0.0, lkid = 0.0;
    if (valid && sl < NX)
    {
        lkid = src[xb + sl];
        if (MODE != 3) cpart = v.in.b[xb + sl] * lkid;
    }
    double cr[NX];
#pragma unroll
    for (int r = 0; r < NX; r++) cr[r] = 0.0;
    if (own)
    {
        const double *col = isx ? v.in.A + d1.y + (size_t) sl * NX : v.in.B + d1.z + (size_t) (sl - NX) * NX;
        if (NX % 2 == 0)
        {
            const double2 *c2 = reinterpret_cast<const double2 *>(col);
#pragma unroll
            for (int r = 0; r < NX / 2; r++) { const double2 t = __ldg(c2 + r); cr[2 * r] = t.x; cr[2 * r + 1] = t.y; }
        }
        else
        {
#pragma unroll
            for (int r = 0; r < NX; r++) cr[r] = __ldg(col + r);
        }
    }
    sL[lane] = lkid; /* lanes sl >= NX publish zeros that nobody reads */
    __syncwarp();
    if (MODE != 3)
    {
        const double c = half_sum(cpart);
        if (sl == 0 && act) { if (MODE == 0) v.cmod0[g] = c; else v.dc[g] = c; }
    }
    const double *lk = sL + 16 * grp;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < NX; r++) s += cr[r] * lk[r];
    if (own && act)
    {
        double acc = base;
        acc -= s;
        if (MODE == 2)
        {
            if (isx) v.dq[xo + i] = acc; else v.dr[uo + i] = acc;
        }
        else
        {
            int mask;
            const double unc = winv * acc;
            const double val = clip_mask(unc, lo, hi, mask);
            if (isx)
            {
                v.x[xo + i] = val;
                if (MODE == 0) { v.qmod0[xo + i] = acc; v.xUnc[xo + i] = unc; v.P[xo + i] = mask == 0 ? winv : 0.0; }
            }
            else
            {
                v.u[uo + i] = val;
                if (MODE == 0) { v.rmod0[uo + i] = acc; v.uUnc[uo + i] = unc; v.Rt[uo + i] = mask == 0 ? winv : 0.0; }
            }
        }
    }
    if (MODE == 2)
    {
        const double dotp = half_sum((own && isx) ? rs * base : 0.0);
        if (sl == 0 && act) v.nvb[g] = dotp;
    }
    __syncwarp(); /* sL is reused by the warp's next item */
}

/* 32/NXP leaves: ld = this lane's LeafDesc (valid = the leaf exists) */
template <int NX, int MODE>
__device__ __forceinline__ void stage_leaves(const View &v, const int4 ld, bool valid, int i)
{
    constexpr int NXP = tqb_pow2_ge(NX);
    const double *src = (MODE == 2) ? v.dlam : v.lam;
    double part = 0.0;
    int g = -1;
    bool changed = false;
    if (valid)
    {
        if (v.status[ld.z] == TQB_RUNNING) g = ld.x;
        if (i < NX && g >= 0)
        {
            const int e = ld.y + i;
            const double l = ld.w ? 0.0 : src[e];
            if (MODE == 2)
            {
                v.dq[e] = l;
                part = ld.w ? 0.0 : v.res[e] * l;
            }
            else
            {
                const double pold = (MODE == 0) ? v.P[e] : 0.0; /* with the other inputs: a load behind the stores would be a second round trip */
                const double acc = -v.in.q[e] + l, winv = v.in.Qinv[e], lo = v.in.xmin[e], hi = v.in.xmax[e];
                const double unc = winv * acc;
                int mask;
                v.x[e] = clip_mask(unc, lo, hi, mask);
                if (MODE == 0)
                {
                    const double pnew = mask == 0 ? winv : 0.0;
                    changed = pnew != pold;
                    v.qmod0[e] = acc; v.xUnc[e] = unc; v.P[e] = pnew;
                }
            }
        }
    }
    if (MODE == 0)
    {   /* per leaf (NXP lanes): any bound entered or left the active set */
        const unsigned b = __ballot_sync(FULL, changed);
        const int lane = threadIdx.x & 31, first = lane & ~(NXP - 1);
        const unsigned grp = NXP == 32 ? 0xffffffffu : (((1u << NXP) - 1u) << first);
        if (g >= 0 && i == 0 && v.chg) v.chg[g] = (b & grp) != 0;
    }
    if (MODE == 2)
    {   /* segmented butterfly: the same additions as a full-warp sum padded with zeros */
#pragma unroll
        for (int o = NXP / 2; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        if (g >= 0 && i == 0) { v.nvb[g] = part; v.dc[g] = 0.0; }
    }
    else if (MODE == 0)
    {
        if (g >= 0 && i == 0) v.cmod0[g] = 0.0;
    }
}

/* Persistent warps: item = a parent (items [0, Np)) or a group of 32/NXP leaves; a warp walks the
 * items w, w + nwarps, ... and fetches the descriptor of its next item before it works on the
 * current one, so the descriptor -> data dependency costs one memory latency per warp, not per item. */
template <int NX, int NU, int MODE>
__global__ void __launch_bounds__(32 * TQB_STAGE_WPB, TQB_STAGE_MINBLOCKS) k_stage_fused(View v, int Np, int nroots, int root_nx, int nleaf, int pair_begin)
{
    /* items: [0, PB) single parents | NPR pairs of chain parents (descriptors PB + 2j, PB + 2j + 1) |
     * groups of 32/NXP leaves */
    constexpr int NXP = tqb_pow2_ge(NX), LPW = 32 / NXP;
    constexpr bool PAIR = NX + NU <= 16 && NX <= 8; /* larger NX: the second code path costs registers the binary-tree case needs */
    __shared__ double sLs[TQB_STAGE_WPB][32];
    __shared__ double sSs[TQB_STAGE_WPB][TQB_STAGE_MAXPAIR];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int PB = PAIR ? pair_begin : Np, NPR = (Np - PB + 1) / 2, NPI = PB + NPR;
    const int nwarps = gridDim.x * TQB_STAGE_WPB, nitems = NPI + (nleaf + LPW - 1) / LPW;
    const int slot = lane / NXP, li = lane - slot * NXP;
    int4 n0 = make_int4(0, 0, 0, 0), n1 = n0, n2 = n0;
    auto fetch = [&](int it) {
        if (it < NPI)
        {
            const int e = it < PB ? it : PB + 2 * (it - PB) + (lane >> 4);
            if (e < Np)
            {
                const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * (size_t) e;
                n0 = __ldg(dp); n1 = __ldg(dp + 1); n2 = __ldg(dp + 2);
            }
        }
        else if (it < nitems)
        {
            const int lf = (it - NPI) * LPW + slot;
            if (lf < nleaf) n0 = __ldg(reinterpret_cast<const int4 *>(v.leaf_desc) + lf);
        }
    };
    int it = blockIdx.x * TQB_STAGE_WPB + wib;
    fetch(it);
    while (it < nitems)
    {
        const int4 d0 = n0, d1 = n1, d2 = n2;
        const int nxt = it + nwarps;
        fetch(nxt);
        if (it < PB) stage_parent<NX, NU, MODE>(v, d0, d1, d2, it < nroots, root_nx, sLs[wib], sSs[wib], lane);
        else if (it < NPI)
        {
            if constexpr (PAIR) stage_parent_pair<NX, NU, MODE>(v, d0, d1, d2, PB + 2 * (it - PB) + (lane >> 4) < Np, sLs[wib], lane);
        }
        else stage_leaves<NX, MODE>(v, d0, (it - NPI) * LPW + slot < nleaf, li);
        it = nxt;
    }
}


/* ---- stage QPs from the affine model: lambda step + next iteration's stage solution in one stream -------------
 * After the line search the multipliers move along dlam by the accepted step: lambda += tau dlam.  The stage
 * gradients are affine in lambda -- qmod(lambda + tau dlam) = qmod0 + tau dq with dq from k_stage_fused<2> -- which
 * is what every line-search trial already evaluates (k_ls_eval).  This kernel applies the step to lambda AND forms
 * the next Newton iteration's stage solution from the model: qmod0 += tau dq, x = clip(Qinv qmod0), xUnc, P, the
 * active-set change flags, cmod0 += tau dc (the same for u) -- vectors only, [A B] is not read: k_stage_fused<0>
 * (a full pass over A and B, 201 MB on config 3) is skipped for that iteration.  Against recomputing from lambda the
 * stage gradient differs by rounding (one fma per element and iteration), so the driver refreshes it from lambda
 * with k_stage_fused<0> every TQB_AFFINE_REFRESH-th iteration (and at iteration 0): the drift never spans more than
 * a few iterations.  LPN lanes per node (>= nx + nu, a power of two): lane e < nx owns x element e, lane nx + e u
 * element e. */
template <int LPN>
__global__ void __launch_bounds__(256) k_step_affine(View v, int with_stage)
{
    constexpr int NPW = 32 / LPN;
    const int lane = threadIdx.x & 31, slot = lane / LPN, e = lane - slot * LPN;
    const int g = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * NPW + slot;
    bool changed = false;
    bool live = g < v.Nn;
    int t = 0;
    if (live) { t = v.T.tree[g]; live = v.status[t] == TQB_RUNNING; }
    if (live)
    {
        const double tau = v.t_taufin[t];
        const int nx = v.T.nx[g], nu = v.T.nu[g];
        const bool root = v.T.dad[g] < 0;
        if (v.ghost && v.ghost[g]) with_stage = 0; /* the multiplier moves everywhere, the stage solution comes from the owner */
        if (e < nx)
        {
            const int i = v.T.xoff[g] + e;
            if (!root) v.lam[i] += tau * v.dlam[i];
            if (with_stage)
            {
                const double pold = v.P[i], winv = v.in.Qinv[i], lo = v.in.xmin[i], hi = v.in.xmax[i];
                const double qm = v.qmod0[i] + tau * v.dq[i];
                const double unc = winv * qm;
                int mask;
                const double val = clip_mask(unc, lo, hi, mask);
                const double pnew = mask == 0 ? winv : 0.0;
                changed = pnew != pold;
                v.qmod0[i] = qm; v.x[i] = val; v.xUnc[i] = unc; v.P[i] = pnew;
            }
        }
        else if (with_stage && e < nx + nu)
        {
            const int i = v.T.uoff[g] + e - nx;
            const double pold = v.Rt[i], winv = v.in.Rinv[i], lo = v.in.umin[i], hi = v.in.umax[i];
            const double rm = v.rmod0[i] + tau * v.dr[i];
            const double unc = winv * rm;
            int mask;
            const double val = clip_mask(unc, lo, hi, mask);
            const double pnew = mask == 0 ? winv : 0.0;
            changed = pnew != pold;
            v.rmod0[i] = rm; v.u[i] = val; v.uUnc[i] = unc; v.Rt[i] = pnew;
        }
        if (with_stage && e == 0) v.cmod0[g] += tau * v.dc[g];
    }
    {
        const unsigned b = __ballot_sync(FULL, changed);
        const unsigned grp = LPN == 32 ? 0xffffffffu : (((1u << LPN) - 1u) << (slot * LPN));
        if (live && with_stage && e == 0 && v.chg) v.chg[g] = (b & grp) != 0;
    }
}
