/*
 * sdunes_device.cu -- SDUNES (dual Newton strategy on the scenario decomposition) as ONE persistent
 * sm_100a kernel: the whole Newton loop of treeqp_sdunes_solve (dual_Newton_scenarios.c:1850-2091)
 * runs on the device without returning to the host.
 *
 * Why one kernel: an SDUNES QP is small (config 2: 81 scenarios x 20 stages of 4 states) and its
 * Newton iteration is a chain of dependent phases, one of which (the non-anticipativity system J,
 * scenarios.c:795-1074) is a serial block-tridiagonal recursion over the scenario pairs.  The cost
 * is latency, not bandwidth: the working set (< a few MB) lives in L2, the active block of every
 * warp in shared memory.  The kernel is launched as a single thread-block CLUSTER (up to 8 CTAs on
 * neighbouring SMs); phases are separated by the hardware cluster barrier instead of kernel
 * launches, and every decision (termination, Armijo) is taken redundantly by all threads from
 * fixed-shape reductions, so the run is bit-reproducible.
 *
 * Phase                                   replaces (dual_Newton_scenarios.c)            unit
 *   sd_stage_solve                        solve_stage_problems :311-416 (clipping part)   element
 *   sd_build                              :417-471 (Zbar, LambdaD, LambdaL), calculate_residuals :489,
 *                                         calculate_last_residual :528, error :1452        element
 *   sd_factor_scenario                    factorize_Lambda :590, form_K :692, first half of
 *                                         form_RHS_non_anticipaticity :871                warp / scenario
 *   sd_jay_chain                          form_and_factorize_Jay :795, rest of :871,
 *                                         calculate_delta_lambda :1004                    one warp (serial chain)
 *   sd_delta_mu_scenario                  calculate_delta_mu :1077 + affine model of the
 *                                         stage gradients along the step                  warp / scenario
 *   sd_fval                               evaluate_dual_function :1229 (NEW_FVAL), line_search :1385   element
 *   sd_export                             copy-out :2026-2075                              element
 */
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../host/host_internal.h"
#include "warp_helpers.cuh"

namespace cg = cooperative_groups;

static thread_local char g_sd_err[512] = "";
#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            snprintf(g_sd_err, sizeof g_sd_err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            fprintf(stderr, "[treeqp_b200] CUDA error %s\n", g_sd_err);                          \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

struct SdView
{
    tqb_sd_dims d;
    tqb_sd_tables T;
    tqb_sd_inputs in;
    tqb_sd_outputs out; /* x,u,mu,lambda iterates live in out.xs, out.us, out.mu_s, out.lam_s */
    double *xU, *dxU, *uU, *duU;      /* unconstrained stage solutions and their derivative along the step */
    double *Qc, *Rc;                  /* QinvCal, RinvCal */
    double *res, *resMod, *dmu;
    double *Zbar, *LD, *LL, *CLD, *CLL, *CLDi;
    double *Ut, *K;
    double *CJD, *CJDi, *CJL, *resNA, *rhsNA, *dlam;
    double *red;                      /* [4 slots][2][16] cross-CTA reduction scratch */
    int *info;
    tqb_opts o;
    int reverse;                      /* reverseCholesky = checkLastActiveSet (scenarios.c:1922) */
    int per_warp;                     /* doubles of shared-memory scratch per warp */
};

struct Team
{
    int tid, T, lane, wib, wpb, gw, W, nb, rank, slot;
};

#define NODE(i, k) (v.T.nodeIdx[(i) * (Nh + 1) + (k)])
#define REM(i, k) (v.T.removed[(i) * (Nh + 1) + (k)])

__device__ __forceinline__ void team_sync() { cg::this_cluster().sync(); }

/* fixed-shape reduction over the whole team of (a: sum or max, b: sum); every thread gets both */
__device__ __forceinline__ void team_reduce2(const SdView &v, Team &tm, double &a, double &b, bool a_is_max, double *s_red)
{
    a = a_is_max ? warp_max(a) : warp_sum(a);
    b = warp_sum(b);
    if (tm.lane == 0) { s_red[tm.wib] = a; s_red[32 + tm.wib] = b; }
    __syncthreads();
    if (tm.wib == 0)
    {
        double ta = tm.lane < tm.wpb ? s_red[tm.lane] : 0.0;
        double tb = tm.lane < tm.wpb ? s_red[32 + tm.lane] : 0.0;
        ta = a_is_max ? warp_max(ta) : warp_sum(ta);
        tb = warp_sum(tb);
        if (tm.lane == 0)
        {
            double *g = v.red + tm.slot * 32;
            g[tm.rank] = ta;
            g[16 + tm.rank] = tb;
        }
    }
    team_sync();
    const double *g = v.red + tm.slot * 32;
    double ra = 0.0, rb = 0.0;
    for (int r = 0; r < tm.nb; r++)
    {
        const double xa = __ldcg(g + r), xb = __ldcg(g + 16 + r);
        ra = a_is_max ? fmax(ra, xa) : ra + xa;
        rb += xb;
    }
    a = ra;
    b = rb;
    tm.slot = (tm.slot + 1) & 3;
}

/* ------------------------------------------------------------------------------------------
 * Stage QPs of all (scenario, stage) pairs at the dual point (mu, lambda), one thread per
 * component (scenarios.c:346-416):
 *   x[k+1] = clip(Qinv .* (mu[k] - A[k+2]' mu[k+1] - q)),  u[k] = clip(Rinv .* (-B' mu[k] - r -+ lambda))
 * Nodes whose bounds were "removed" (shared with an earlier scenario) are not clipped.
 * ---------------------------------------------------------------------------------------- */
__device__ void sd_stage_solve(const SdView &v, const Team &tm, bool with_cal)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Ns = v.d.Ns, nv = nx + nu;
    const double *mu = v.out.mu_s, *lam = v.out.lam_s;
    for (int e = tm.tid; e < v.d.SK * nv; e += tm.T)
    {
        const int ik = e / nv, w = e - ik * nv, i = ik / Nh, k = ik - i * Nh;
        const int idxm1 = NODE(i, k), idx = NODE(i, k + 1);
        if (w < nx)
        {
            const int a = w, g = idx * nx + a;
            double s = mu[ik * nx + a];
            if (k < Nh - 1)
            {
                const double *An = v.in.A + (size_t) NODE(i, k + 2) * nx * nx, *mun = mu + (size_t) (ik + 1) * nx;
                double t = 0.0;
                for (int c = 0; c < nx; c++) t += An[c + a * nx] * mun[c];
                s = s - t;
            }
            s = s - v.in.q[g];
            const double xu = v.in.Qinv[g] * s;
            v.xU[ik * nx + a] = xu;
            double xv = xu;
            if (REM(i, k + 1) != 1)
            {
                int m;
                xv = clip_mask(xu, v.in.xmin[g], v.in.xmax[g], m);
                if (with_cal) v.Qc[ik * nx + a] = m == 0 ? v.in.Qinv[g] : 0.0;
            }
            v.out.xs[ik * nx + a] = xv;
        }
        else
        {
            const int a = w - nx, g = idxm1 * nu + a;
            const double *Bk = v.in.B + (size_t) idx * nx * nu, *muk = mu + (size_t) ik * nx;
            double t = 0.0;
            for (int c = 0; c < nx; c++) t += Bk[c + a * nx] * muk[c];
            double s = -v.in.r[g] - t;
            if (i < Ns - 1 && k < v.T.cn[i]) s = s - lam[v.T.lo[i] + k * nu + a];
            if (i > 0 && k < v.T.cn[i - 1]) s = s + lam[v.T.lo[i - 1] + k * nu + a];
            const double uu = v.in.Rinv[g] * s;
            v.uU[ik * nu + a] = uu;
            double uv = uu;
            if (REM(i, k) != 1)
            {
                int m;
                uv = clip_mask(uu, v.in.umin[g], v.in.umax[g], m);
                if (with_cal) v.Rc[ik * nu + a] = m == 0 ? v.in.Rinv[g] : 0.0;
            }
            v.out.us[ik * nu + a] = uv;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Dual gradient (dynamics residual per (i,k), non-anticipativity residual per pair), its norm,
 * and the blocks of the scenario Hessians:
 *   Zbar = B RinvCal,  LambdaD[k] = Zbar B' + QinvCal[k] + A QinvCal[k-1] A',  LambdaL[k-1] = -A QinvCal[k-1]
 * (stored transposed when reverse).  Returns this thread's part of the termination norm.
 * ---------------------------------------------------------------------------------------- */
__device__ double sd_build(const SdView &v, const Team &tm)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, nxx = nx * nx;
    const double *x = v.out.xs, *u = v.out.us;
    const bool is_max = v.o.termCondition == 2;
    double part = 0.0;
    for (int e = tm.tid; e < v.d.SK * nx; e += tm.T)
    {
        const int ik = e / nx, a = e - ik * nx, i = ik / Nh, k = ik - i * Nh;
        const int idx = NODE(i, k + 1);
        const double *Bk = v.in.B + (size_t) idx * nx * nu, *Ak = v.in.A + (size_t) idx * nxx;
        const double *uk = u + (size_t) ik * nu;
        const double bk = v.in.b[idx * nx + a];
        double s;
        if (k == 0)
        {
            double t = 0.0;
            for (int c = 0; c < nu; c++) t += Bk[a + c * nx] * uk[c];
            s = -bk - t;
            s = s + x[e];
        }
        else
        {
            const double *xp = x + (size_t) (ik - 1) * nx;
            double t = 0.0;
            for (int c = 0; c < nx; c++) t += Ak[a + c * nx] * xp[c];
            s = x[e] - t;
            t = 0.0;
            for (int c = 0; c < nu; c++) t += Bk[a + c * nx] * uk[c];
            s = s - t;
            s = s - bk;
        }
        v.res[e] = s;
        part = is_max ? fmax(part, fabs(s)) : part + s * s;
    }
    for (int e = tm.tid; e < v.d.nl; e += tm.T)
    {
        const int i = v.T.pair_of[e], le = e - v.T.lo[i], k = le / nu, a = le - k * nu;
        double s = 0.0;
        s += -1.0 * u[(size_t) (i * Nh + k) * nu + a];
        s += u[(size_t) ((i + 1) * Nh + k) * nu + a];
        v.resNA[e] = s;
        part = is_max ? fmax(part, fabs(s)) : part + s * s;
    }
    for (int e = tm.tid; e < v.d.SK * nx * nu; e += tm.T)
    {
        const int ik = e / (nx * nu), ac = e - ik * nx * nu, c = ac / nx, i = ik / Nh, k = ik - i * Nh;
        const int idx = NODE(i, k + 1);
        v.Zbar[e] = v.in.B[(size_t) idx * nx * nu + ac] * v.Rc[ik * nu + c];
    }
    for (int e = tm.tid; e < v.d.SK * nxx; e += tm.T)
    {
        const int ik = e / nxx, ab = e - ik * nxx, b = ab / nx, a = ab - b * nx, i = ik / Nh, k = ik - i * Nh;
        const int idx = NODE(i, k + 1);
        const double *Bk = v.in.B + (size_t) idx * nx * nu, *Ak = v.in.A + (size_t) idx * nxx;
        const double *Rc = v.Rc + (size_t) ik * nu;
        double s = 0.0;
        for (int c = 0; c < nu; c++) s += (Bk[a + c * nx] * Rc[c]) * Bk[b + c * nx];
        if (a == b) s += v.Qc[ik * nx + a];
        if (k > 0)
        {
            const double *Qp = v.Qc + (size_t) (ik - 1) * nx;
            double s2 = 0.0;
            if (v.reverse)
            {   /* LL = -QinvCal[k-1] A' ; LD -= A LL   (:447-457) */
                for (int c = 0; c < nx; c++) s2 += Ak[a + c * nx] * (-(Qp[c] * Ak[b + c * nx]));
                v.LL[(size_t) (ik - 1) * nxx + ab] = -(Qp[a] * Ak[b + a * nx]);
            }
            else
            {   /* LL = -A QinvCal[k-1] ; LD -= LL A'   (:462-469) */
                for (int c = 0; c < nx; c++) s2 += (-(Ak[a + c * nx] * Qp[c])) * Ak[b + c * nx];
                v.LL[(size_t) (ik - 1) * nxx + ab] = -(Ak[a + b * nx] * Qp[b]);
            }
            s -= s2;
        }
        v.LD[e] = s;
    }
    return part;
}

__device__ __forceinline__ void warp_load_mat(double *dst, int ld, const double *src, int n, int lane)
{
    for (int e = lane; e < n * n; e += 32)
    {
        const int b = e / n, a = e - b * n;
        dst[a + b * ld] = src[e];
    }
}

struct WarpScratch
{
    double *sD, *sD0, *sL, *sinv, *sy, *sz, *tA, *tB;
};

__device__ __forceinline__ WarpScratch warp_scratch(const SdView &v, double *base)
{
    const int nx = v.d.nx, ld = nx | 1;
    WarpScratch s;
    s.sD = base; s.sD0 = s.sD + nx * ld; s.sL = s.sD0 + nx * ld;
    s.sinv = s.sL + nx * ld; s.sy = s.sinv + nx; s.sz = s.sy + nx;
    s.tA = s.sz + nx; s.tB = s.tA + v.d.nK * nx;
    return s;
}

/* banded substitution with the factor of Lambda_i (scenarios.c:893-958 and :1129-1188):
 * out = Lambda_i^-1 rhs.  mode 0: rhs = res_i; mode 1: rhs = -res_i + Zbar*(dlam_{i-1} - dlam_i). */
__device__ void sd_banded_solve(const SdView &v, int i, int mode, double *out_vec, const WarpScratch &S, int lane)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Ns = v.d.Ns, Nr = v.d.Nr, nxx = nx * nx, ld = nx | 1;
    const size_t ik0 = (size_t) i * Nh;
    for (int s = 0; s < Nh; s++)
    {   /* L-solves in factorisation order */
        const int k = v.reverse ? Nh - 1 - s : s;
        warp_load_mat(S.sD, ld, v.CLD + (ik0 + k) * nxx, nx, lane);
        for (int a = lane; a < nx; a += 32) S.sinv[a] = v.CLDi[(ik0 + k) * nx + a];
        if (s > 0) warp_load_mat(S.sL, ld, v.CLL + (ik0 + (v.reverse ? k : k - 1)) * nxx, nx, lane);
        __syncwarp();
        for (int a = lane; a < nx; a += 32)
        {
            double val;
            const double r = v.res[(ik0 + k) * nx + a];
            if (mode == 0) val = r;
            else if (k < Nr)
            {
                const double *Z = v.Zbar + (ik0 + k) * nx * nu;
                double t = 0.0;
                for (int c = 0; c < nu; c++)
                {
                    double vv = 0.0;
                    if (i > 0 && k < v.T.cn[i - 1]) vv = vv + v.dlam[v.T.lo[i - 1] + k * nu + c];
                    if (i < Ns - 1 && k < v.T.cn[i]) vv = vv - v.dlam[v.T.lo[i] + k * nu + c];
                    t += Z[a + c * nx] * vv;
                }
                val = -r + t;
            }
            else val = 0.0 - r;
            if (s > 0)
            {
                double t = 0.0;
                for (int c = 0; c < nx; c++) t += S.sL[a + c * ld] * S.sz[c];
                val = val - t;
            }
            S.sy[a] = val;
        }
        __syncwarp();
        warp_trsv_lnn(S.sD, ld, nx, S.sinv, S.sy, lane);
        for (int a = lane; a < nx; a += 32)
        {
            S.sz[a] = S.sy[a];
            out_vec[k * nx + a] = S.sy[a];
        }
        __syncwarp();
    }
    for (int s = 0; s < Nh; s++)
    {   /* L'-solves in the opposite order */
        const int k = v.reverse ? s : Nh - 1 - s;
        warp_load_mat(S.sD, ld, v.CLD + (ik0 + k) * nxx, nx, lane);
        for (int a = lane; a < nx; a += 32) S.sinv[a] = v.CLDi[(ik0 + k) * nx + a];
        if (s > 0) warp_load_mat(S.sL, ld, v.CLL + (ik0 + (v.reverse ? k - 1 : k)) * nxx, nx, lane);
        __syncwarp();
        for (int a = lane; a < nx; a += 32)
        {
            double val = out_vec[k * nx + a];
            if (s > 0)
            {
                double t = 0.0;
                for (int c = 0; c < nx; c++) t += S.sL[c + a * ld] * S.sz[c];
                val = val - t;
            }
            S.sy[a] = val;
        }
        __syncwarp();
        warp_trsv_ltn(S.sD, ld, nx, S.sinv, S.sy, lane);
        for (int a = lane; a < nx; a += 32)
        {
            S.sz[a] = S.sy[a];
            out_vec[k * nx + a] = S.sy[a];
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------
 * Scenario-local part of the Newton system, one warp per scenario i:
 *   block-tridiagonal Cholesky of Lambda_i (factorize_Lambda :590-689, reverse order when
 *   checkLastActiveSet), Ut_i and K_i = -Ut Ut' + diag(RinvCal) (form_K :692-792),
 *   resMod_i = Lambda_i^-1 res_i (:893-958).
 * ---------------------------------------------------------------------------------------- */
__device__ void sd_factor_scenario(const SdView &v, int i, const WarpScratch &S, int lane)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Nr = v.d.Nr, nK = v.d.nK, nxx = nx * nx, ld = nx | 1;
    const size_t ik0 = (size_t) i * Nh;
    int nreg = 0;
    for (int s = 0; s < Nh; s++)
    {
        const int k = v.reverse ? Nh - 1 - s : s;
        if (s == 0) warp_load_mat(S.sD, ld, v.LD + (ik0 + k) * nxx, nx, lane);
        __syncwarp();
        nreg += warp_potrf_reg(S.sD, S.sD0, ld, nx, S.sinv, lane, v.o.regType, v.o.regTol, v.o.regValue);
        for (int e = lane; e < nxx; e += 32)
        {
            const int b = e / nx, a = e - b * nx;
            v.CLD[(ik0 + k) * nxx + e] = S.sD[a + b * ld];
        }
        for (int a = lane; a < nx; a += 32) v.CLDi[(ik0 + k) * nx + a] = S.sinv[a];
        if (s < Nh - 1)
        {
            const int kl = v.reverse ? k - 1 : k, kn = v.reverse ? k - 1 : k + 1;
            warp_load_mat(S.sL, ld, v.LL + (ik0 + kl) * nxx, nx, lane);
            __syncwarp();
            /* CholLambdaL = LambdaL CholLambdaD^-T (dtrsm_rltn): one row per lane */
            for (int r = lane; r < nx; r += 32)
                for (int j = 0; j < nx; j++)
                {
                    double t = S.sL[r + j * ld];
                    for (int c = 0; c < j; c++) t -= S.sL[r + c * ld] * S.sD[j + c * ld];
                    S.sL[r + j * ld] = t * S.sinv[j];
                }
            __syncwarp();
            for (int e = lane; e < nxx; e += 32)
            {
                const int b = e / nx, a = e - b * nx;
                v.CLL[(ik0 + kl) * nxx + e] = S.sL[a + b * ld];
            }
            __syncwarp();
            /* next diagonal block: LambdaD -= CholLambdaL CholLambdaL' (dsyrk_ln) */
            for (int e = lane; e < nxx; e += 32)
            {
                const int b = e / nx, a = e - b * nx;
                double val = v.LD[(ik0 + kn) * nxx + e];
                if (a >= b)
                {
                    double t = 0.0;
                    for (int c = 0; c < nx; c++) t += S.sL[a + c * ld] * S.sL[b + c * ld];
                    val -= t;
                }
                S.sD[a + b * ld] = val;
            }
        }
        __syncwarp();
    }
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);

    /* Ut_i: row r = j*nu + c of block row j starts from Zbar[j]' at stage j and is propagated
     * along the factor (one row per lane, independent chains) */
    double *Ut = v.Ut + (size_t) i * nK * Nh * nx, *K = v.K + (size_t) i * nK * nK;
    for (int r = lane; r < nK; r += 32)
    {
        const int j = r / nu, c = r - j * nu;
        double *ta = S.tA + r * nx, *tb = S.tB + r * nx;
        const double *Z = v.Zbar + (ik0 + j) * nx * nu;
        for (int a = 0; a < nx; a++) ta[a] = Z[a + c * nx];
        int k = j;
        while (v.reverse ? k >= 0 : k < Nh)
        {
            const double *CLD = v.CLD + (ik0 + k) * nxx, *CLDi = v.CLDi + (ik0 + k) * nx;
            for (int jj = 0; jj < nx; jj++)
            {
                double t = ta[jj];
                for (int q = 0; q < jj; q++) t -= tb[q] * CLD[jj + q * nx];
                tb[jj] = t * CLDi[jj];
                Ut[r + (size_t) (k * nx + jj) * nK] = tb[jj];
            }
            const bool more = v.reverse ? k > 0 : k < Nh - 1;
            if (more)
            {
                const double *CLL = v.CLL + (ik0 + (v.reverse ? k - 1 : k)) * nxx;
                for (int a = 0; a < nx; a++)
                {
                    double t = 0.0;
                    for (int q = 0; q < nx; q++) t += tb[q] * CLL[a + q * nx];
                    ta[a] = -t;
                }
            }
            k += v.reverse ? -1 : 1;
        }
    }
    __syncwarp();
    for (int e = lane; e < nK * nK; e += 32)
    {
        const int b = e / nK, a = e - b * nK;
        if (a < b) continue;
        const int ja = a / nu, jb = b / nu;
        /* columns where both rows are non-zero */
        const int t0 = v.reverse ? 0 : ja * nx, t1 = v.reverse ? (jb + 1) * nx : Nh * nx;
        double t = 0.0;
        for (int q = t0; q < t1; q++) t += Ut[a + (size_t) q * nK] * Ut[b + (size_t) q * nK];
        double val = -t;
        K[b + a * nK] = val;
        if (a == b) val += v.Rc[(ik0 + ja) * nu + (a - ja * nu)];
        K[a + b * nK] = val;
    }
    (void) Nr;
    sd_banded_solve(v, i, 0, v.resMod + ik0 * nx, S, lane);
}

/* ------------------------------------------------------------------------------------------
 * The non-anticipativity system (one warp; a serial block-tridiagonal recursion over the
 * scenario pairs): J_i = K_i + K_{i+1} restricted to the common inputs, sub-diagonal -K_{i+1};
 * Cholesky merged with the forward substitution, then the backward substitution.
 * (form_and_factorize_Jay :795-868, form_RHS_non_anticipaticity :960-1000, calculate_delta_lambda :1004-1074)
 * ---------------------------------------------------------------------------------------- */
__device__ void sd_jay_chain(const SdView &v, double *sm, int lane)
{
    const int nu = v.d.nu, nx = v.d.nx, Ns = v.d.Ns, Nh = v.d.Nh, nK = v.d.nK, m = v.d.mJ, ld = m | 1;
    double *sJ = sm, *sJ0 = sJ + m * ld, *sJL = sJ0 + m * ld, *sCar = sJL + m * ld;
    double *sinv = sCar + m * ld, *sy = sinv + m, *scv = sy + m;
    for (int e = lane; e < v.d.nl; e += 32)
    {
        const int i = v.T.pair_of[e], le = e - v.T.lo[i], k = le / nu, c = le - k * nu;
        double s = v.resNA[e], t = 0.0;
        const double *Z = v.Zbar + (size_t) (i * Nh + k) * nx * nu, *rm = v.resMod + (size_t) (i * Nh + k) * nx;
        for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
        s = s - t;
        Z = v.Zbar + (size_t) ((i + 1) * Nh + k) * nx * nu;
        rm = v.resMod + (size_t) ((i + 1) * Nh + k) * nx;
        t = 0.0;
        for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
        s = s + t;
        v.rhsNA[e] = s;
    }
    __syncwarp();
    int nreg = 0;
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * v.T.cn[i], lo = v.T.lo[i];
        const double *K0 = v.K + (size_t) i * nK * nK, *K1 = K0 + nK * nK;
        for (int e = lane; e < dim * dim; e += 32)
        {
            const int b = e / dim, a = e - b * dim;
            double val = (i > 0 && a >= b) ? sCar[a + b * ld] : 0.0;
            val += K0[a + b * nK];
            val += K1[a + b * nK];
            sJ[a + b * ld] = val;
        }
        for (int e = lane; e < dim; e += 32)
        {
            double rhs = v.rhsNA[lo + e];
            if (i > 0) rhs = rhs + scv[e];
            sy[e] = -1.0 * rhs;
        }
        __syncwarp();
        nreg += warp_potrf_reg(sJ, sJ0, ld, dim, sinv, lane, v.o.regType, v.o.regTol, v.o.regValue);
        for (int e = lane; e < dim * dim; e += 32)
        {
            const int b = e / dim, a = e - b * dim;
            v.CJD[v.T.jdo[i] + e] = sJ[a + b * ld];
        }
        for (int e = lane; e < dim; e += 32) v.CJDi[lo + e] = sinv[e];
        warp_trsv_lnn(sJ, ld, dim, sinv, sy, lane);
        for (int e = lane; e < dim; e += 32) v.rhsNA[lo + e] = sy[e];
        if (i < Ns - 2)
        {
            const int dimN = nu * v.T.cn[i + 1];
            for (int e = lane; e < dimN * dim; e += 32)
            {
                const int b = e / dimN, a = e - b * dimN;
                sJL[a + b * ld] = 0.0 + -1.0 * K1[a + b * nK];
            }
            __syncwarp();
            for (int a = lane; a < dimN; a += 32)
                for (int j = 0; j < dim; j++)
                {
                    double t = sJL[a + j * ld];
                    for (int c = 0; c < j; c++) t -= sJL[a + c * ld] * sJ[j + c * ld];
                    sJL[a + j * ld] = t * sinv[j];
                }
            __syncwarp();
            for (int e = lane; e < dimN * dim; e += 32)
            {
                const int b = e / dimN, a = e - b * dimN;
                v.CJL[v.T.jlo[i] + e] = sJL[a + b * ld];
            }
            for (int e = lane; e < dimN * dimN; e += 32)
            {
                const int b = e / dimN, a = e - b * dimN;
                if (a < b) continue;
                double t = 0.0;
                for (int c = 0; c < dim; c++) t += sJL[a + c * ld] * sJL[b + c * ld];
                sCar[a + b * ld] = 0.0 - t;
            }
            for (int a = lane; a < dimN; a += 32)
            {
                double t = 0.0;
                for (int c = 0; c < dim; c++) t += sJL[a + c * ld] * sy[c];
                scv[a] = t;
            }
        }
        __syncwarp();
    }
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);
    for (int i = Ns - 2; i >= 0; i--)
    {
        const int dim = nu * v.T.cn[i], lo = v.T.lo[i];
        for (int e = lane; e < dim * dim; e += 32)
        {
            const int b = e / dim, a = e - b * dim;
            sJ[a + b * ld] = v.CJD[v.T.jdo[i] + e];
        }
        for (int e = lane; e < dim; e += 32)
        {
            sinv[e] = v.CJDi[lo + e];
            double rhs = v.rhsNA[lo + e];
            if (i < Ns - 2) rhs = rhs - scv[e];
            sy[e] = rhs;
        }
        __syncwarp();
        warp_trsv_ltn(sJ, ld, dim, sinv, sy, lane);
        for (int e = lane; e < dim; e += 32) v.dlam[lo + e] = sy[e];
        if (i > 0)
        {
            const int dimP = nu * v.T.cn[i - 1];
            const double *CJL = v.CJL + v.T.jlo[i - 1]; /* dim x dimP */
            for (int c = lane; c < dimP; c += 32)
            {
                double t = 0.0;
                for (int a = 0; a < dim; a++) t += CJL[a + c * dim] * sy[a];
                scv[c] = t;
            }
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------
 * dmu_i = Lambda_i^-1 (-res_i + Zbar (dlam_{i-1} - dlam_i))  (calculate_delta_mu :1077-1205), then
 * the derivative of the unconstrained stage solutions along (dmu, dlam): along the step the stage
 * gradients are affine in the step length, so a line-search trial is a pure vector pass.
 * Returns this lane's part of res'dmu.
 * ---------------------------------------------------------------------------------------- */
__device__ double sd_delta_mu_scenario(const SdView &v, int i, const WarpScratch &S, int lane)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Ns = v.d.Ns, nv = nx + nu;
    const size_t ik0 = (size_t) i * Nh;
    double *dmu = v.dmu + ik0 * nx;
    sd_banded_solve(v, i, 1, dmu, S, lane);
    double part = 0.0;
    for (int e = lane; e < Nh * nv; e += 32)
    {
        const int k = e / nv, w = e - k * nv;
        const int idxm1 = NODE(i, k), idx = NODE(i, k + 1);
        if (w < nx)
        {
            const int a = w;
            double s = dmu[k * nx + a];
            if (k < Nh - 1)
            {
                const double *An = v.in.A + (size_t) NODE(i, k + 2) * nx * nx, *dn = dmu + (k + 1) * nx;
                double t = 0.0;
                for (int c = 0; c < nx; c++) t += An[c + a * nx] * dn[c];
                s = s - t;
            }
            v.dxU[(ik0 + k) * nx + a] = v.in.Qinv[idx * nx + a] * s;
            part += v.res[(ik0 + k) * nx + a] * dmu[k * nx + a];
        }
        else
        {
            const int a = w - nx;
            const double *Bk = v.in.B + (size_t) idx * nx * nu, *dk = dmu + k * nx;
            double t = 0.0;
            for (int c = 0; c < nx; c++) t += Bk[c + a * nx] * dk[c];
            double s = -t;
            if (i < Ns - 1 && k < v.T.cn[i]) s = s - v.dlam[v.T.lo[i] + k * nu + a];
            if (i > 0 && k < v.T.cn[i - 1]) s = s + v.dlam[v.T.lo[i - 1] + k * nu + a];
            v.duU[(ik0 + k) * nu + a] = v.in.Rinv[idxm1 * nu + a] * s;
        }
    }
    return part;
}

/* dual function at (mu + tau dmu, lambda + tau dlam), NEW_FVAL form (scenarios.c:1340-1355):
 * f = sum -1/2 x Q (x - 2 xUnc) - 1/2 u R (u - 2 uUnc) - b'mu ; returns this thread's part */
__device__ double sd_fval(const SdView &v, const Team &tm, double tau)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh;
    double acc = 0.0;
    for (int e = tm.tid; e < v.d.SK * nx; e += tm.T)
    {
        const int ik = e / nx, a = e - ik * nx, i = ik / Nh, k = ik - i * Nh;
        const int g = NODE(i, k + 1) * nx + a;
        const double xu = v.xU[e] + tau * v.dxU[e];
        double xv = xu;
        int m;
        if (REM(i, k + 1) != 1) xv = clip_mask(xu, v.in.xmin[g], v.in.xmax[g], m);
        const double t = v.in.Q[g] * (xv + -2.0 * xu);
        acc -= 0.5 * (xv * t);
        acc -= v.in.b[g] * (v.out.mu_s[e] + tau * v.dmu[e]);
    }
    for (int e = tm.tid; e < v.d.SK * nu; e += tm.T)
    {
        const int ik = e / nu, a = e - ik * nu, i = ik / Nh, k = ik - i * Nh;
        const int g = NODE(i, k) * nu + a;
        const double uu = v.uU[e] + tau * v.duU[e];
        double uv = uu;
        int m;
        if (REM(i, k) != 1) uv = clip_mask(uu, v.in.umin[g], v.in.umax[g], m);
        const double t = v.in.R[g] * (uv + -2.0 * uu);
        acc -= 0.5 * (uv * t);
    }
    return acc;
}

/* copy-out (scenarios.c:2026-2075): node results from the first scenario through each node,
 * edge multipliers = sum of the scenario multipliers, mu = W .* (unconstrained - clipped) */
__device__ void sd_export(const SdView &v, const Team &tm)
{
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Nn = v.d.Nn;
    for (int e = tm.tid; e < Nn * nx; e += tm.T)
    {
        const int n = e / nx, a = e - n * nx, st = v.T.nstage[n];
        if (st == 0) { v.out.x[e] = 0.0; v.out.lam[e] = 0.0; v.out.mux[e] = 0.0; continue; }
        const int i0 = v.T.nfirst[n], cnt = v.T.ncount[n], k = st - 1;
        const size_t p = ((size_t) i0 * Nh + k) * nx + a;
        const double xv = v.out.xs[p];
        v.out.x[e] = xv;
        v.out.mux[e] = v.in.Q[e] * (v.xU[p] - xv);
        double s = 0.0;
        for (int i = i0; i < i0 + cnt; i++) s += v.out.mu_s[((size_t) i * Nh + k) * nx + a];
        v.out.lam[e] = s;
    }
    for (int e = tm.tid; e < Nn * nu; e += tm.T)
    {
        const int n = e / nu, a = e - n * nu, st = v.T.nstage[n];
        if (st >= Nh) { v.out.u[e] = 0.0; v.out.muu[e] = 0.0; continue; }
        const size_t p = ((size_t) v.T.nfirst[n] * Nh + st) * nu + a;
        const double uv = v.out.us[p];
        v.out.u[e] = uv;
        v.out.muu[e] = v.in.R[e] * (v.uU[p] - uv);
    }
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
/* phase timer (thread 0 of the team): nanoseconds per phase accumulated over the solve */
#define PHASE(idx)                                                             \
    do {                                                                       \
        if (tm.tid == 0) { const unsigned long long t_ = gtime(); v.info[8 + 256 + (idx)] += (int) (t_ - t_phase); t_phase = t_; } \
    } while (0)

__global__ void __launch_bounds__(512) k_sdunes_solve(SdView v)
{
    extern __shared__ double smem[];
    __shared__ double s_red[64];
    Team tm;
    tm.nb = gridDim.x; tm.rank = blockIdx.x;
    tm.tid = blockIdx.x * blockDim.x + threadIdx.x; tm.T = gridDim.x * blockDim.x;
    tm.lane = threadIdx.x & 31; tm.wib = threadIdx.x >> 5; tm.wpb = blockDim.x >> 5;
    tm.gw = tm.tid >> 5; tm.W = tm.T >> 5; tm.slot = 0;
    const int nx = v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Ns = v.d.Ns, SK = v.d.SK, nl = v.d.nl;
    const WarpScratch S = warp_scratch(v, smem + (size_t) tm.wib * v.per_warp);

    /* dual start, QinvCal/RinvCal of nodes with removed bounds never change (:1901-1917) */
    for (int e = tm.tid; e < SK * nx; e += tm.T)
    {
        const int ik = e / nx, a = e - ik * nx, i = ik / Nh, k = ik - i * Nh;
        v.out.mu_s[e] = v.in.mu0[e];
        v.Qc[e] = v.in.Qinv[NODE(i, k + 1) * nx + a];
    }
    for (int e = tm.tid; e < SK * nu; e += tm.T)
    {
        const int ik = e / nu, a = e - ik * nu, i = ik / Nh, k = ik - i * Nh;
        v.Rc[e] = v.in.Rinv[NODE(i, k) * nu + a];
    }
    for (int e = tm.tid; e < nl; e += tm.T) v.out.lam_s[e] = v.in.lam0[e];
    for (int e = tm.tid; e < TQB_SD_INFO_INTS; e += tm.T) v.info[e] = 0;
    team_sync();
    unsigned long long t_phase = gtime();

    int status = 1, it = 0, ls_total = 0;
    for (it = 0; it < v.o.maxIter; it++)
    {
        sd_stage_solve(v, tm, true);
        team_sync();
        PHASE(0);
        double err = sd_build(v, tm), dummy = 0.0;
        team_reduce2(v, tm, err, dummy, v.o.termCondition == 2, s_red);
        PHASE(1);
        if (v.o.termCondition == 1) err = sqrt(err);
        if (err < v.o.tol) { status = 0; break; }

        for (int i = tm.gw; i < Ns; i += tm.W) sd_factor_scenario(v, i, S, tm.lane);
        team_sync();
        PHASE(2);
        if (Ns > 1)
        {
            if (tm.gw == 0) sd_jay_chain(v, smem, tm.lane);
            team_sync();
        }
        PHASE(3);
        double dot = 0.0;
        for (int i = tm.gw; i < Ns; i += tm.W) dot += sd_delta_mu_scenario(v, i, S, tm.lane);
        for (int e = tm.tid; e < nl; e += tm.T) dot += v.resNA[e] * v.dlam[e];
        /* the reduction's barrier also publishes dxU/duU/dmu; f(0) needs only this scenario's own
         * data for x but all of it is team-visible after the barrier, so evaluate f0 afterwards */
        team_reduce2(v, tm, dot, dummy, false, s_red);
        PHASE(4);
        double f0 = sd_fval(v, tm, 0.0);
        team_reduce2(v, tm, f0, dummy, false, s_red);
        PHASE(5);

        double tau = 1.0, tauPrev = 0.0, tauEval = 0.0;
        int ls;
        for (ls = 1; ls <= v.o.lsMaxIter; ls++)
        {
            double f = sd_fval(v, tm, tau);
            team_reduce2(v, tm, f, dummy, false, s_red);
            tauEval = tau;
            if (f <= f0 + v.o.gamma * tau * dot) break;
            tauPrev = tau;
            tau = v.o.beta * tauPrev;
        }
        (void) tauPrev;
        for (int e = tm.tid; e < SK * nx; e += tm.T) v.out.mu_s[e] += tauEval * v.dmu[e];
        for (int e = tm.tid; e < nl; e += tm.T) v.out.lam_s[e] += tauEval * v.dlam[e];
        if (tm.tid == 0 && it < 256) v.info[8 + it] = ls;
        ls_total += ls;
        PHASE(6);
        team_sync();
        PHASE(7);
    }
    if (status == 1)
    {   /* ran out of iterations: x,u of the last line-search trial point (what the reference holds) */
        sd_stage_solve(v, tm, false);
    }
    team_sync();
    sd_export(v, tm);
    PHASE(8);
    if (tm.tid == 0)
    {
        v.info[0] = status;
        v.info[1] = it;
        v.info[2] = ls_total;
    }
}

/* ========================================================================== host driver ==== */

struct tqb_sd_device
{
    tqb_sd_dims dims;
    int *d_int;
    double *d_in, *d_out, *d_work;
    int *d_info;
    SdView v;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int wpb, nblocks;
    size_t smem;
};

static int sd_build_device(tqb_sd_device *D, const int *h_int)
{
    const tqb_sd_dims *d = &D->dims;
    CK(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&D->ev0));
    CK(cudaEventCreate(&D->ev1));
    CK(cudaMalloc(&D->d_int, sizeof(int) * (size_t) d->n_int));
    CK(cudaMemcpy(D->d_int, h_int, sizeof(int) * (size_t) d->n_int, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&D->d_in, sizeof(double) * (size_t) d->n_in));
    CK(cudaMalloc(&D->d_out, sizeof(double) * (size_t) d->n_out));
    CK(cudaMalloc(&D->d_info, sizeof(int) * TQB_SD_INFO_INTS));

    SdView &v = D->v;
    memset(&v, 0, sizeof v);
    v.d = *d;
    tqb_sd_layout_tables(d, D->d_int, &v.T);
    tqb_sd_layout_inputs(d, D->d_in, &v.in);
    tqb_sd_layout_outputs(d, D->d_out, &v.out);
    v.info = D->d_info;

    const size_t SK = (size_t) d->SK, nx = d->nx, nu = d->nu, nxx = nx * nx, nK = d->nK;
    const size_t nl = d->nl > 0 ? d->nl : 1;
    size_t nwork = 5 * SK * nx + 3 * SK * nu /* xU dxU Qc res resMod dmu->(6) */ + SK * nx
                   + SK * nx * nu + 4 * SK * nxx + SK * nx
                   + (size_t) d->Ns * nK * d->Nh * nx + (size_t) d->Ns * nK * nK
                   + (size_t) d->njd + (size_t) d->njl + 5 * nl + 4 * 32 + 64;
    CK(cudaMalloc(&D->d_work, sizeof(double) * nwork));
    CK(cudaMemset(D->d_work, 0, sizeof(double) * nwork));
    double *p = D->d_work;
#define TK(field, n) v.field = p; p += (((size_t) (n)) + 1) / 2 * 2
    TK(xU, SK * nx); TK(dxU, SK * nx); TK(Qc, SK * nx); TK(res, SK * nx); TK(resMod, SK * nx); TK(dmu, SK * nx);
    TK(uU, SK * nu); TK(duU, SK * nu); TK(Rc, SK * nu);
    TK(Zbar, SK * nx * nu); TK(LD, SK * nxx); TK(LL, SK * nxx); TK(CLD, SK * nxx); TK(CLL, SK * nxx); TK(CLDi, SK * nx);
    TK(Ut, (size_t) d->Ns * nK * d->Nh * nx); TK(K, (size_t) d->Ns * nK * nK);
    TK(CJD, d->njd); TK(CJL, d->njl); TK(CJDi, nl); TK(resNA, nl); TK(rhsNA, nl); TK(dlam, nl);
    TK(red, 4 * 32);
#undef TK
    if ((size_t) (p - D->d_work) > nwork) { snprintf(g_sd_err, sizeof g_sd_err, "sdunes work arena overflow"); return 1; }

    /* launch shape: one warp per scenario where possible, one cluster of <= 8 CTAs */
    int dev = 0, smem_max = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t ld = nx | 1, mJ = d->mJ > 0 ? d->mJ : 1, ldm = mJ | 1;
    size_t per_warp = 3 * nx * ld + 3 * nx + 2 * nK * nx;
    per_warp = (per_warp + 1) / 2 * 2;
    const size_t jay = 4 * mJ * ldm + 3 * mJ;
    const size_t lim = (size_t) smem_max - 4096;
    int wpb = 16;
    while (wpb > 1 && (size_t) wpb * per_warp * 8 > lim) wpb >>= 1;
    if ((size_t) wpb * per_warp * 8 > lim || jay * 8 > lim)
    {
        snprintf(g_sd_err, sizeof g_sd_err, "SDUNES blocks (nx=%d, nu*Nr=%d) need more shared memory than one SM has",
                 d->nx, d->nK);
        fprintf(stderr, "[treeqp_b200] %s\n", g_sd_err);
        return 1;
    }
    int nb = (d->Ns + wpb - 1) / wpb;
    if (nb > 8) nb = 8;
    if (nb < 1) nb = 1;
    D->wpb = wpb;
    D->nblocks = nb;
    size_t sm = (size_t) wpb * per_warp;
    if (jay > sm) sm = jay;
    D->smem = sm * 8;
    v.per_warp = (int) per_warp;
    CK(cudaFuncSetAttribute(k_sdunes_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D->smem));
    return 0;
}

extern "C" tqb_sd_device *tqb_sd_device_create(const tqb_sd_dims *dims, const int *h_int)
{
    if (tqb_cuda_device_count() <= 0)
    {
        fprintf(stderr, "[treeqp_b200] no CUDA device visible: libtreeqp_b200 has no CPU fallback\n");
        return NULL;
    }
    tqb_sd_device *D = (tqb_sd_device *) calloc(1, sizeof(tqb_sd_device));
    D->dims = *dims;
    if (sd_build_device(D, h_int)) { tqb_sd_device_destroy(D); return NULL; }
    return D;
}

extern "C" void tqb_sd_device_destroy(tqb_sd_device *D)
{
    if (!D) return;
    if (D->stream) cudaStreamSynchronize(D->stream);
    cudaFree(D->d_int); cudaFree(D->d_in); cudaFree(D->d_out); cudaFree(D->d_work); cudaFree(D->d_info);
    if (D->ev0) cudaEventDestroy(D->ev0);
    if (D->ev1) cudaEventDestroy(D->ev1);
    if (D->stream) cudaStreamDestroy(D->stream);
    free(D);
}

extern "C" int tqb_sd_device_upload(tqb_sd_device *D, const double *h_in, double *ms)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    CK(cudaMemcpyAsync(D->d_in, h_in, sizeof(double) * (size_t) D->dims.n_in, cudaMemcpyHostToDevice, D->stream));
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms) *ms = t;
    return 0;
}

extern "C" int tqb_sd_device_download(tqb_sd_device *D, double *h_out, double *ms)
{
    CK(cudaEventRecord(D->ev0, D->stream));
    CK(cudaMemcpyAsync(h_out, D->d_out, sizeof(double) * (size_t) D->dims.n_out, cudaMemcpyDeviceToHost, D->stream));
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, D->ev0, D->ev1));
    if (ms) *ms = t;
    return 0;
}

extern "C" int tqb_sd_device_solve(tqb_sd_device *D, const tqb_opts *o, int reverse, int *h_info, double *gpu_ms,
                                   long long *launches)
{
    D->v.o = *o;
    D->v.reverse = reverse ? 1 : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned) D->nblocks, 1, 1);
    cfg.blockDim = dim3((unsigned) (32 * D->wpb), 1, 1);
    cfg.dynamicSmemBytes = D->smem;
    cfg.stream = D->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned) D->nblocks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaEventRecord(D->ev0, D->stream));
    CK(cudaLaunchKernelEx(&cfg, k_sdunes_solve, D->v));
    CK(cudaEventRecord(D->ev1, D->stream));
    CK(cudaMemcpyAsync(h_info, D->d_info, sizeof(int) * TQB_SD_INFO_INTS, cudaMemcpyDeviceToHost, D->stream));
    CK(cudaStreamSynchronize(D->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, D->ev0, D->ev1));
    if (gpu_ms) *gpu_ms = ms;
    if (launches) *launches = 1;
    return 0;
}
