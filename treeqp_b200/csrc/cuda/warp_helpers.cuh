/*
 * warp_helpers.cuh -- warp-level building blocks shared by the TDUNES and SDUNES kernels:
 * shuffle reductions, BLASFEO's clip and Cholesky pivot rules, and in-shared-memory
 * factorisation / substitution of one small dense block by one warp.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

/* inclusive clip, lower bound wins when both fire (blasfeo_dveccl_mask, AVX branch,
 * external/blasfeo/auxiliary/d_aux_lib4.c:4412-4417) */
__device__ __forceinline__ double clip_mask(double v, double lo, double hi, int &mask)
{
    const bool up = hi <= v, dn = v <= lo;
    mask = dn ? -1 : (up ? 1 : 0);
    return dn ? lo : (up ? hi : v);
}

/* BLASFEO's Cholesky pivot rule (kernel_dgemm_4x4_lib4.c:4633-4646): c > 0 ? (l = sqrt(c), inv = 1/l)
 * : (0, 0).  sqrt and 1/l come from ONE reciprocal-square-root seed (MUFU.RSQ64H) refined by two
 * coupled Goldschmidt steps, a Markstein correction of the root and one Newton step of the
 * reciprocal against the ROUNDED root: 13 FP64 operations instead of the 33 instructions of
 * sqrt() followed by a division, and half the dependent latency of a Cholesky step.  A CPU
 * emulation (fma(), 22-bit seed, 2e7 samples over 26 decades) reproduces correctly rounded
 * sqrt(c) and 1/sqrt(c)-rounded bit for bit; outside [1e-290, 1e290] the input is rescaled by 4^k.
 * -DTQB_IEEE_PIVOT forces the library sqrt and division everywhere. */
/* one rsqrt seed -> correctly rounded sqrt(c) and the rounded reciprocal of that root */
__device__ __forceinline__ void tqb_rsqrt_core(double c, double &root, double &inv)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
    double h = 0.5 * y, g = c * y, r;
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    g = fma(fma(-g, g, c), h, g);            /* correctly rounded sqrt(c) */
    const double yy = h + h;                 /* ~ 1/sqrt(c) */
    inv = fma(yy, fma(-g, yy, 1.0), yy);     /* 1/g */
    root = g;
}
/* rare pivots: non-positive (BLASFEO: l = 0, inv = 0) or outside [1e-290, 1e290] (rescaled by an exact 4^k) */
/* by value: reference parameters of a non-inlined function would force the caller's pivot into local
 * memory (one store per Cholesky step whether or not the rare path runs; ncu showed 76 local-store
 * sectors per 8 x 8 block, a third of the kernel's L1 traffic) */
static __device__ __noinline__ double2 tqb_pivot_rare(double piv)
{
    if (!(piv > 0.0)) return make_double2(0.0, 0.0);
    double sc, isc, inv;
    if (piv < 1.0) { piv *= 0x1p600; sc = 0x1p-300; isc = 0x1p300; }
    else { piv *= 0x1p-600; sc = 0x1p300; isc = 0x1p-300; }
    tqb_rsqrt_core(piv, piv, inv);
    return make_double2(piv * sc, inv * isc);
}
__device__ __forceinline__ void tqb_pivot(double &piv, double &inv)
{
#ifndef TQB_IEEE_PIVOT
    if (piv > 1e-290 && piv < 1e290)
    {   /* the only branch of the common case */
        double root, rinv;
        tqb_rsqrt_core(piv, root, rinv);
        piv = root; inv = rinv;
    }
    else
    {
        const double2 t = tqb_pivot_rare(piv);
        piv = t.x; inv = t.y;
    }
#else
    if (piv > 0.0) { piv = sqrt(piv); inv = 1.0 / piv; }
    else { piv = 0.0; inv = 0.0; }
#endif
}

/* same results as tqb_pivot; the common-case arithmetic runs unconditionally and the rare cases patch
 * it afterwards, so that the caller's independent work (the trailing update of the previous column)
 * sits in the same basic block as the 100-cycle dependent chain and is scheduled into its bubbles */
__device__ __forceinline__ void tqb_pivot_spec(double &piv, double &inv)
{
#ifndef TQB_IEEE_PIVOT
    const double c = piv;
    double root, rinv;
    tqb_rsqrt_core(c, root, rinv);
    piv = root; inv = rinv;
    if (!(c > 1e-290 && c < 1e290))
    {
        const double2 t = tqb_pivot_rare(c);
        piv = t.x; inv = t.y;
    }
#else
    tqb_pivot(piv, inv);
#endif
}

/* ------------------------------------------------------------------------------------------
 * In-shared-memory Cholesky of a d x d block by one warp (left-looking, column by column) with
 * BLASFEO's pivot rule (kernel_dgemm_4x4_lib4.c:4633-4646): pivot>0 ? (l=sqrt,inv=1/l) : (0,0).
 * ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ void warp_potrf(double *sW, int ld, int d, double *sinv, int lane)
{
    for (int j = 0; j < d; j++)
    {
        for (int i = j + lane; i < d; i += 32)
        {
            double s = sW[i + j * ld];
            for (int t = 0; t < j; t++) s -= sW[i + t * ld] * sW[j + t * ld];
            sW[i + j * ld] = s;
        }
        __syncwarp();
        double piv = sW[j + j * ld], inv;
        tqb_pivot(piv, inv);
        __syncwarp();
        for (int i = j + lane; i < d; i += 32)
        {
            if (i == j) { sW[j + j * ld] = piv; sinv[j] = inv; }
            else sW[i + j * ld] *= inv;
        }
        __syncwarp();
    }
}

/* y <- L^-T y in shared memory (dtrsv_ltn with the stored inverse diagonal) */
__device__ __forceinline__ void warp_trsv_ltn(const double *sW, int ld, int d, const double *sinv,
                                              double *sy, int lane)
{
    for (int i = d - 1; i >= 0; i--)
    {
        const double yi = sy[i] * sinv[i];
        __syncwarp();
        if (lane == 0) sy[i] = yi;
        for (int t = lane; t < i; t += 32) sy[t] -= sW[i + t * ld] * yi;
        __syncwarp();
    }
}


/* y <- L^-1 y in shared memory (dtrsv_lnn with the stored inverse diagonal) */
__device__ __forceinline__ void warp_trsv_lnn(const double *sW, int ld, int d, const double *sinv,
                                              double *sy, int lane)
{
    for (int j = 0; j < d; j++)
    {
        const double yj = sy[j] * sinv[j];
        __syncwarp();
        if (lane == 0) sy[j] = yj;
        for (int i = j + 1 + lane; i < d; i += 32) sy[i] -= sW[i + j * ld] * yj;
        __syncwarp();
    }
}

/* ---- d <= 32 variants: one row/element per lane, shuffles instead of shared-memory round trips ---- */

/* left-looking Cholesky, one barrier per column: lane t owns row j+t of column j, lane 0 the pivot */
__device__ __forceinline__ void warp_potrf32(double *sW, int ld, int d, double *sinv, int lane)
{
    for (int j = 0; j < d; j++)
    {
        const int i = j + lane;
        const bool act = i < d;
        double s = 0.0;
        if (act)
        {
            s = sW[i + j * ld];
            for (int t = 0; t < j; t++) s -= sW[i + t * ld] * sW[j + t * ld];
        }
        double piv = __shfl_sync(FULL, s, 0), inv;
        tqb_pivot(piv, inv);
        if (act) sW[i + j * ld] = lane == 0 ? piv : s * inv;
        if (lane == 0) sinv[j] = inv;
        __syncwarp();
    }
}

/* y <- L^-1 y, y held one element per lane (lanes >= d ignored); L in shared memory */
__device__ __forceinline__ double warp_trsv_lnn32(const double *sW, int ld, int d, double inv, double y, int lane)
{
    for (int j = 0; j < d; j++)
    {
        const double yj = __shfl_sync(FULL, y, j) * __shfl_sync(FULL, inv, j);
        if (lane == j) y = yj;
        else if (lane > j && lane < d) y -= sW[lane + j * ld] * yj;
    }
    return y;
}

/* y <- L^-T y, same conventions */
__device__ __forceinline__ double warp_trsv_ltn32(const double *sW, int ld, int d, double inv, double y, int lane)
{
    for (int i = d - 1; i >= 0; i--)
    {
        const double yi = __shfl_sync(FULL, y, i) * __shfl_sync(FULL, inv, i);
        if (lane == i) y = yi;
        else if (lane < i) y -= sW[i + lane * ld] * yi;
    }
    return y;
}

/* Cholesky with the Levenberg-Marquardt policy of treeqp_dpotrf_l_with_reg_opts
 * (dual_Newton_common.c:35-78): regType 1 = always add regValue to the diagonal first,
 * regType 2 = factorise, and if any pivot <= regTol add regValue and factorise again.
 * sW0 is scratch of the same shape (keeps the unfactorised block for the retry).
 * Returns 1 if the block was regularised. */
__device__ __forceinline__ int warp_potrf_reg(double *sW, double *sW0, int ld, int d, double *sinv, int lane,
                                              int regType, double regTol, double regValue)
{
    if (regType == 1)
        for (int j = lane; j < d; j += 32) sW[j + j * ld] += regValue;
    if (regType == 2)
        for (int e = lane; e < d * d; e += 32)
        {
            const int a = e % d, b = e / d;
            sW0[a + b * ld] = sW[a + b * ld];
        }
    __syncwarp();
    if (d <= 32) warp_potrf32(sW, ld, d, sinv, lane); else warp_potrf(sW, ld, d, sinv, lane);
    if (regType == 2)
    {
        bool small = false;
        for (int j = lane; j < d; j += 32) small |= (sW[j + j * ld] <= regTol);
        if (__any_sync(FULL, small))
        {
            __syncwarp();
            for (int e = lane; e < d * d; e += 32)
            {
                const int a = e % d, b = e / d;
                sW[a + b * ld] = sW0[a + b * ld] + (a == b ? regValue : 0.0);
            }
            __syncwarp();
            if (d <= 32) warp_potrf32(sW, ld, d, sinv, lane); else warp_potrf(sW, ld, d, sinv, lane);
            return 1;
        }
        return 0;
    }
    return regType == 1;
}
