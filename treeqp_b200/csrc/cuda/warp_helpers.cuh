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
        if (piv > 0.0) { piv = sqrt(piv); inv = 1.0 / piv; }
        else { piv = 0.0; inv = 0.0; }
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
    warp_potrf(sW, ld, d, sinv, lane);
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
            warp_potrf(sW, ld, d, sinv, lane);
            return 1;
        }
        return 0;
    }
    return regType == 1;
}
