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
    double *jarena;                   /* operands of the J chain when they do not fit in shared memory */
    int jay_in_smem;
    int no_small_jay;                 /* TREEQP_B200_GENERIC_JAY=1: always the generic chain walk */
    int serial_jay;                   /* TREEQP_B200_SERIAL_JAY=1: one warp walks the whole chain (no twisted factorisation) */
    int fixed_nx;                     /* 4 / 8 / 16: compile-time-dimension scenario chains; 0: runtime (TREEQP_B200_GENERIC_NX=1) */
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
template <int NX>
__device__ double sd_build(const SdView &v, const Team &tm)
{
    const int nx = NX > 0 ? NX : v.d.nx, nu = v.d.nu, Nh = v.d.Nh, nxx = nx * nx;
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

/* ---- asynchronous global -> shared staging (LDGSTS): the next block of a serial chain is in
 * flight while the current one is factorised, so the L2 latency leaves the dependent chain ---- */
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned) __cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void warp_fetch_mat(double *dst, int ld, const double *src, int n, int lane)
{
    for (int e = lane; e < n * n; e += 32)
    {
        const int b = e / n, a = e - b * n;
        cp_async8(dst + a + b * ld, src + e);
    }
}

struct WarpScratch
{
    double *bD0, *bL0, *sD0, *binv0, *bvec0, *sz, *tA, *tB;
    int ms, vs; /* strides between the two buffers of a pair (matrix / vector) */
    __device__ __forceinline__ double *bD(int q) const { return bD0 + q * ms; }
    __device__ __forceinline__ double *bL(int q) const { return bL0 + q * ms; }
    __device__ __forceinline__ double *binv(int q) const { return binv0 + q * vs; }
    __device__ __forceinline__ double *bvec(int q) const { return bvec0 + q * vs; }
};

__host__ __device__ inline size_t sd_per_warp_doubles(int nx, int nK)
{
    const size_t ld = nx | 1;
    const size_t n = 5 * (size_t) nx * ld + 5 * (size_t) nx + 2 * (size_t) nK * nx;
    return (n + 1) / 2 * 2;
}

__device__ __forceinline__ WarpScratch warp_scratch(const SdView &v, double *base)
{
    const int nx = v.d.nx, ld = nx | 1;
    WarpScratch s;
    s.ms = nx * ld; s.vs = nx;
    s.bD0 = base; s.bL0 = s.bD0 + 2 * s.ms;
    s.sD0 = s.bL0 + 2 * s.ms;
    s.binv0 = s.sD0 + s.ms;
    s.bvec0 = s.binv0 + 2 * nx;
    s.sz = s.bvec0 + 2 * nx;
    s.tA = s.sz + nx; s.tB = s.tA + v.d.nK * nx;
    return s;
}

/* banded substitution with the factor of Lambda_i (scenarios.c:893-958 and :1129-1188):
 * out = Lambda_i^-1 in (in may alias out).  One element per lane (nx <= 32); the factor blocks of
 * the next stage are prefetched while the current stage is substituted. */
#ifndef TQB_SD_FIXED_NX_DEFAULT
#define TQB_SD_FIXED_NX_DEFAULT 1
#endif

template <int NX>
__device__ void sd_banded_solve(const SdView &v, int i, const double *in_vec, double *out_vec, const WarpScratch &S, int lane)
{
    const int nx = NX > 0 ? NX : v.d.nx, Nh = v.d.Nh, nxx = nx * nx, ld = nx | 1;
    const size_t ik0 = (size_t) i * Nh;
    for (int phase = 0; phase < 2; phase++)
    {   /* phase 0: L-solves in factorisation order; phase 1: L'-solves in the opposite order */
        const bool down = (phase == 0) == (v.reverse != 0); /* stage index decreasing */
#define SD_K(s_) (down ? Nh - 1 - (s_) : (s_))
#define SD_FETCH(s_)                                                                              \
        do {                                                                                      \
            const int k_ = SD_K(s_), q_ = (s_) & 1;                                               \
            warp_fetch_mat(S.bD(q_), ld, v.CLD + (ik0 + k_) * nxx, nx, lane);                     \
            if (lane < nx) cp_async8(S.binv(q_) + lane, v.CLDi + (ik0 + k_) * nx + lane);         \
            if ((s_) > 0)                                                                         \
            {   /* sub-diagonal block linking this stage with the previous one of the sweep */    \
                const int kl_ = down ? k_ : k_ - 1;                                               \
                warp_fetch_mat(S.bL(q_), ld, v.CLL + (ik0 + kl_) * nxx, nx, lane);                \
            }                                                                                     \
            if (phase == 0 || (s_) > 0)                                                           \
            {                                                                                     \
                const double *src_ = phase == 0 ? in_vec : out_vec;                               \
                if (lane < nx) cp_async8(S.bvec(q_) + lane, src_ + k_ * nx + lane);               \
            }                                                                                     \
            cp_async_commit();                                                                    \
        } while (0)
        SD_FETCH(0);
        for (int s = 0; s < Nh; s++)
        {
            const int k = SD_K(s), q = s & 1;
            cp_async_wait_all();
            __syncwarp();
            if (s + 1 < Nh) SD_FETCH(s + 1);
            const double *sD = S.bD(q), *sL = S.bL(q);
            double y = 0.0, inv = 0.0;
            if (lane < nx)
            {
                inv = S.binv(q)[lane];
                y = (phase == 1 && s == 0) ? S.sz[lane] : S.bvec(q)[lane];
                if (s > 0)
                {
                    double t = 0.0;
                    if (phase == 0) for (int c = 0; c < nx; c++) t += sL[lane + c * ld] * S.sz[c];
                    else for (int c = 0; c < nx; c++) t += sL[c + lane * ld] * S.sz[c];
                    y = y - t;
                }
            }
            y = phase == 0 ? warp_trsv_lnn32(sD, ld, nx, inv, y, lane) : warp_trsv_ltn32(sD, ld, nx, inv, y, lane);
            __syncwarp();
            if (lane < nx)
            {
                S.sz[lane] = y;
                out_vec[k * nx + lane] = y;
            }
            __syncwarp();
        }
#undef SD_FETCH
#undef SD_K
    }
}

/* ------------------------------------------------------------------------------------------
 * Scenario-local part of the Newton system, one warp per scenario i:
 *   block-tridiagonal Cholesky of Lambda_i (factorize_Lambda :590-689, reverse order when
 *   checkLastActiveSet), Ut_i and K_i = -Ut Ut' + diag(RinvCal) (form_K :692-792),
 *   resMod_i = Lambda_i^-1 res_i (:893-958).
 * ---------------------------------------------------------------------------------------- */
template <int NX>
__device__ void sd_factor_scenario(const SdView &v, int i, const WarpScratch &S, int lane)
{
    const int nx = NX > 0 ? NX : v.d.nx, nu = v.d.nu, Nh = v.d.Nh, nK = v.d.nK, nxx = nx * nx, ld = nx | 1;
    const size_t ik0 = (size_t) i * Nh;
    double *sinv = S.binv(0);
    int nreg = 0;
#define SD_FETCH(s_)                                                                              \
    do {                                                                                          \
        const int k_ = v.reverse ? Nh - 1 - (s_) : (s_), q_ = (s_) & 1;                           \
        warp_fetch_mat(S.bD(q_), ld, v.LD + (ik0 + k_) * nxx, nx, lane);                          \
        if ((s_) < Nh - 1) warp_fetch_mat(S.bL(q_), ld, v.LL + (ik0 + (v.reverse ? k_ - 1 : k_)) * nxx, nx, lane); \
        cp_async_commit();                                                                        \
    } while (0)
    SD_FETCH(0);
    for (int s = 0; s < Nh; s++)
    {
        const int k = v.reverse ? Nh - 1 - s : s, q = s & 1;
        double *sD = S.bD(q), *sL = S.bL(q);
        cp_async_wait_all();
        __syncwarp();
        if (s > 0)
        {   /* LambdaD -= CholLambdaL CholLambdaL' of the previous stage (dsyrk_ln) */
            const double *sLp = S.bL(q ^ 1);
            for (int e = lane; e < nxx; e += 32)
            {
                const int b = e / nx, a = e - b * nx;
                if (a < b) continue;
                double t = 0.0;
                for (int c = 0; c < nx; c++) t += sLp[a + c * ld] * sLp[b + c * ld];
                sD[a + b * ld] -= t;
            }
            __syncwarp();
        }
        if (s + 1 < Nh) SD_FETCH(s + 1);
        nreg += warp_potrf_reg(sD, S.sD0, ld, nx, sinv, lane, v.o.regType, v.o.regTol, v.o.regValue);
        for (int e = lane; e < nxx; e += 32)
        {
            const int b = e / nx, a = e - b * nx;
            v.CLD[(ik0 + k) * nxx + e] = sD[a + b * ld];
        }
        if (lane < nx) v.CLDi[(ik0 + k) * nx + lane] = sinv[lane];
        if (s < Nh - 1)
        {
            const int kl = v.reverse ? k - 1 : k;
            /* CholLambdaL = LambdaL CholLambdaD^-T (dtrsm_rltn): one row per lane */
            if (lane < nx)
                for (int j = 0; j < nx; j++)
                {
                    double t = sL[lane + j * ld];
                    for (int c = 0; c < j; c++) t -= sL[lane + c * ld] * sD[j + c * ld];
                    sL[lane + j * ld] = t * sinv[j];
                }
            __syncwarp();
            for (int e = lane; e < nxx; e += 32)
            {
                const int b = e / nx, a = e - b * nx;
                v.CLL[(ik0 + kl) * nxx + e] = sL[a + b * ld];
            }
        }
        __syncwarp();
    }
#undef SD_FETCH
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
    sd_banded_solve<NX>(v, i, v.res + ik0 * nx, v.resMod + ik0 * nx, S, lane);
}

/* ------------------------------------------------------------------------------------------
 * The non-anticipativity system: J_i = K_i + K_{i+1} restricted to the common inputs, sub-diagonal
 * -K_{i+1}: a serial block-tridiagonal recursion over the scenario pairs
 * (form_and_factorize_Jay :795-868, form_RHS_non_anticipaticity :960-1000, calculate_delta_lambda :1004-1074).
 * All operands are first gathered into one arena (shared memory of CTA 0 when they fit, else HBM)
 * by every warp of the CTA; then ONE warp walks the chain without touching L2: Cholesky merged with
 * the forward substitution, then the backward substitution.
 * ---------------------------------------------------------------------------------------- */
struct JayArena
{
    double *JD0, *JD1, *JL, *rhs, *inv, *car, *w0, *scv, *sy;
    int *cn, *lo, *jdo, *jlo;
};

__host__ __device__ inline size_t sd_pad2(size_t n) { return (n + 1) / 2 * 2; }
__host__ __device__ inline size_t sd_jay_doubles(const tqb_sd_dims &d)
{
    const size_t m = d.mJ > 0 ? d.mJ : 1;
    return 2 * sd_pad2(d.njd) + sd_pad2(d.njl) + 2 * sd_pad2(d.nl) + 2 * sd_pad2(m * m) + 2 * sd_pad2(m) + 2 * (size_t) (d.Ns + 1) + 2;
}
__device__ __forceinline__ JayArena jay_arena(const SdView &v, double *base)
{
    const size_t m = v.d.mJ > 0 ? v.d.mJ : 1;
    JayArena a;
    a.JD0 = base; a.JD1 = a.JD0 + sd_pad2(v.d.njd); a.JL = a.JD1 + sd_pad2(v.d.njd);
    a.rhs = a.JL + sd_pad2(v.d.njl); a.inv = a.rhs + sd_pad2(v.d.nl);
    a.car = a.inv + sd_pad2(v.d.nl); a.w0 = a.car + sd_pad2(m * m);
    a.scv = a.w0 + sd_pad2(m * m); a.sy = a.scv + sd_pad2(m);
    a.cn = (int *) (a.sy + sd_pad2(m)); a.lo = a.cn + (v.d.Ns + 1); a.jdo = a.lo + (v.d.Ns + 1); a.jlo = a.jdo + (v.d.Ns + 1);
    return a;
}

/* gather phase: every warp of CTA 0, one scenario pair per warp */
__device__ void sd_jay_gather(const SdView &v, const JayArena &A, const Team &tm)
{
    const int nu = v.d.nu, nx = v.d.nx, Ns = v.d.Ns, Nh = v.d.Nh, nK = v.d.nK, lane = tm.lane;
    for (int e = threadIdx.x; e < Ns + 1; e += blockDim.x)
    {
        A.cn[e] = e < Ns ? v.T.cn[e] : 0;
        A.lo[e] = v.T.lo[e]; A.jdo[e] = v.T.jdo[e]; A.jlo[e] = v.T.jlo[e];
    }
    for (int i = tm.wib; i < Ns - 1; i += tm.wpb)
    {
        const int dim = nu * v.T.cn[i], lo = v.T.lo[i], jdo = v.T.jdo[i];
        const double *K0 = v.K + (size_t) i * nK * nK, *K1 = K0 + nK * nK;
        for (int e = lane; e < dim * dim; e += 32)
        {
            const int b = e / dim, a = e - b * dim;
            A.JD0[jdo + e] = K0[a + b * nK];
            A.JD1[jdo + e] = K1[a + b * nK];
        }
        if (i < Ns - 2)
        {
            const int dimN = nu * v.T.cn[i + 1], jlo = v.T.jlo[i];
            for (int e = lane; e < dimN * dim; e += 32)
            {
                const int b = e / dimN, a = e - b * dimN;
                A.JL[jlo + e] = 0.0 + -1.0 * K1[a + b * nK];
            }
        }
        for (int e = lane; e < dim; e += 32)
        {   /* right-hand side (scenarios.c:960-1000) */
            const int k = e / nu, c = e - k * nu;
            double s = v.resNA[lo + e], t = 0.0;
            const double *Z = v.Zbar + (size_t) (i * Nh + k) * nx * nu, *rm = v.resMod + (size_t) (i * Nh + k) * nx;
            for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
            s = s - t;
            Z = v.Zbar + (size_t) ((i + 1) * Nh + k) * nx * nu;
            rm = v.resMod + (size_t) ((i + 1) * Nh + k) * nx;
            t = 0.0;
            for (int a = 0; a < nx; a++) t += Z[a + c * nx] * rm[a];
            s = s + t;
            A.rhs[lo + e] = s;
        }
    }
}

__device__ void sd_jay_chain(const SdView &v, const JayArena &A, int lane)
{
    const int nu = v.d.nu, Ns = v.d.Ns;
    int nreg = 0;
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * A.cn[i], lo = A.lo[i];
        double *W = A.JD0 + A.jdo[i];
        const double *W1 = A.JD1 + A.jdo[i];
        for (int e = lane; e < dim * dim; e += 32)
        {
            const int b = e / dim, a = e - b * dim;
            double val = (i > 0 && a >= b) ? A.car[e] : 0.0;
            val += W[e];
            val += W1[e];
            W[e] = val;
        }
        for (int e = lane; e < dim; e += 32)
        {
            double rhs = A.rhs[lo + e];
            if (i > 0) rhs = rhs + A.scv[e];
            A.sy[e] = -1.0 * rhs;
        }
        __syncwarp();
        nreg += warp_potrf_reg(W, A.w0, dim, dim, A.inv + lo, lane, v.o.regType, v.o.regTol, v.o.regValue);
        if (dim <= 32)
        {
            double y = lane < dim ? A.sy[lane] : 0.0;
            const double inv = lane < dim ? A.inv[lo + lane] : 0.0;
            y = warp_trsv_lnn32(W, dim, dim, inv, y, lane);
            if (lane < dim) A.sy[lane] = y;
            __syncwarp();
        }
        else
            warp_trsv_lnn(W, dim, dim, A.inv + lo, A.sy, lane);
        for (int e = lane; e < dim; e += 32) A.rhs[lo + e] = A.sy[e];
        if (i < Ns - 2)
        {
            const int dimN = nu * A.cn[i + 1];
            double *JL = A.JL + A.jlo[i];
            for (int a = lane; a < dimN; a += 32)
                for (int j = 0; j < dim; j++)
                {
                    double t = JL[a + j * dimN];
                    for (int c = 0; c < j; c++) t -= JL[a + c * dimN] * W[j + c * dim];
                    JL[a + j * dimN] = t * A.inv[lo + j];
                }
            __syncwarp();
            for (int e = lane; e < dimN * dimN; e += 32)
            {
                const int b = e / dimN, a = e - b * dimN;
                if (a < b) continue;
                double t = 0.0;
                for (int c = 0; c < dim; c++) t += JL[a + c * dimN] * JL[b + c * dimN];
                A.car[e] = 0.0 - t;
            }
            for (int a = lane; a < dimN; a += 32)
            {
                double t = 0.0;
                for (int c = 0; c < dim; c++) t += JL[a + c * dimN] * A.sy[c];
                A.scv[a] = t;
            }
        }
        __syncwarp();
    }
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);
    for (int i = Ns - 2; i >= 0; i--)
    {
        const int dim = nu * A.cn[i], lo = A.lo[i];
        const double *W = A.JD0 + A.jdo[i];
        for (int e = lane; e < dim; e += 32)
        {
            double rhs = A.rhs[lo + e];
            if (i < Ns - 2) rhs = rhs - A.scv[e];
            A.sy[e] = rhs;
        }
        __syncwarp();
        if (dim <= 32)
        {
            double y = lane < dim ? A.sy[lane] : 0.0;
            const double inv = lane < dim ? A.inv[lo + lane] : 0.0;
            y = warp_trsv_ltn32(W, dim, dim, inv, y, lane);
            __syncwarp();
            if (lane < dim) A.sy[lane] = y;
            __syncwarp();
        }
        else
            warp_trsv_ltn(W, dim, dim, A.inv + lo, A.sy, lane);
        for (int e = lane; e < dim; e += 32) v.dlam[lo + e] = A.sy[e];
        if (i > 0)
        {
            const int dimP = nu * A.cn[i - 1];
            const double *CJL = A.JL + A.jlo[i - 1]; /* dim x dimP */
            for (int c = lane; c < dimP; c += 32)
            {
                double t = 0.0;
                for (int a = 0; a < dim; a++) t += CJL[a + c * dim] * A.sy[a];
                A.scv[c] = t;
            }
        }
        __syncwarp();
    }
}

/* The same chain for small blocks (every block dimension <= DM, DM = 4 or 8), entirely in
 * registers: lane l owns ROW l of the current diagonal block, of the coupling block to the next
 * scenario pair and of the carried Schur complement; rows of other lanes arrive by shuffle.  No
 * shared-memory round trips or warp barriers inside a step (the generic walk above spends ~4500
 * cycles per step on them for 4 x 4 blocks).  Operation order per entry is that of sd_jay_chain
 * (left-looking Cholesky of warp_potrf32, dtrsv/dtrsm with the stored inverse diagonal). */
template <int DM>
__device__ void sd_jay_chain_small(const SdView &v, const JayArena &A, int lane)
{
    const int nu = v.d.nu, Ns = v.d.Ns;
    const int regType = v.o.regType;
    const double regTol = v.o.regTol, regValue = v.o.regValue;
    const int l = lane < DM ? lane : DM - 1; /* idle lanes mirror the last row and store nothing */
    const bool mine = lane < DM;
    int nreg = 0;
    double car[DM], scv = 0.0;
#pragma unroll
    for (int c = 0; c < DM; c++) car[c] = 0.0;
#pragma unroll 1
    for (int i = 0; i < Ns - 1; i++)
    {
        const int dim = nu * A.cn[i], lo = A.lo[i];
        double *W = A.JD0 + A.jdo[i];
        const double *W1 = A.JD1 + A.jdo[i];
        const bool row = mine && l < dim;
        double w[DM], w0[DM], inv[DM];
#pragma unroll
        for (int c = 0; c < DM; c++)
        {
            double val = 0.0;
            if (row && c < dim)
            {
                val = (i > 0 && l >= c) ? car[c] : 0.0;
                val += W[l + c * dim];
                val += W1[l + c * dim];
                if (regType == 1 && c == l) val += regValue;
            }
            w[c] = w0[c] = val;
            inv[c] = 0.0;
        }
        double y = 0.0;
        if (row)
        {
            double rhs = A.rhs[lo + l];
            if (i > 0) rhs = rhs + scv;
            y = -1.0 * rhs;
        }
#pragma unroll 1
        for (int attempt = 0; attempt < 2; attempt++)
        {
            bool small = false;
#pragma unroll
            for (int j = 0; j < DM; j++)
            {
                if (j < dim)
                {   /* column j: s = W[l][j] - sum_{t<j} L[l][t] L[j][t], t ascending */
                    double sacc = w[j];
#pragma unroll
                    for (int t = 0; t < j; t++) sacc -= w[t] * __shfl_sync(FULL, w[t], j);
                    double piv = __shfl_sync(FULL, sacc, j), pinv;
                    tqb_pivot(piv, pinv);
                    if (l == j) w[j] = piv;
                    else if (l > j) w[j] = sacc * pinv;
                    inv[j] = pinv;
                    small |= (piv <= regTol);
                }
            }
            if (!(regType == 2 && attempt == 0 && small)) { nreg += (regType == 1 || attempt == 1); break; }
#pragma unroll
            for (int c = 0; c < DM; c++) w[c] = w0[c] + ((c == l && row) ? regValue : 0.0);
        }
        /* y <- L^-1 y (dtrsv_lnn) */
#pragma unroll
        for (int j = 0; j < DM; j++)
            if (j < dim)
            {
                const double yj = __shfl_sync(FULL, y, j) * inv[j];
                if (l == j) y = yj;
                else if (l > j && l < dim) y -= w[j] * yj;
            }
        if (row)
        {
            A.rhs[lo + l] = y;
#pragma unroll
            for (int c = 0; c < DM; c++)
                if (c < dim && c <= l) W[l + c * dim] = w[c];
        }
        if (mine)
        {
#pragma unroll
            for (int c = 0; c < DM; c++)
                if (c < dim && c == l) A.inv[lo + c] = inv[c];
        }
        if (i < Ns - 2)
        {
            const int dimN = nu * A.cn[i + 1];
            double *JL = A.JL + A.jlo[i];
            const bool rowN = mine && l < dimN;
            double jl[DM];
#pragma unroll
            for (int c = 0; c < DM; c++) jl[c] = (rowN && c < dim) ? JL[l + c * dimN] : 0.0;
            /* JL <- JL L^-T (dtrsm_rltn): row l, column by column */
#pragma unroll
            for (int j = 0; j < DM; j++)
                if (j < dim)
                {
                    double t = jl[j];
#pragma unroll
                    for (int c = 0; c < j; c++) t -= jl[c] * __shfl_sync(FULL, w[c], j);
                    jl[j] = t * inv[j];
                    if (rowN) JL[l + j * dimN] = jl[j];
                }
            /* carried Schur complement (lower) and right-hand side for the next pair */
#pragma unroll
            for (int b = 0; b < DM; b++)
            {
                double t = 0.0;
#pragma unroll
                for (int c = 0; c < DM; c++)
                {
                    const double o = __shfl_sync(FULL, jl[c], b);
                    if (c < dim) t += jl[c] * o;
                }
                car[b] = 0.0 - t;
            }
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < DM; c++)
            {
                const double yc = __shfl_sync(FULL, y, c);
                if (c < dim) t += jl[c] * yc;
            }
            scv = t;
        }
    }
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);
    __syncwarp();
    scv = 0.0;
#pragma unroll 1
    for (int i = Ns - 2; i >= 0; i--)
    {
        const int dim = nu * A.cn[i], lo = A.lo[i];
        const double *W = A.JD0 + A.jdo[i];
        const bool row = mine && l < dim;
        double lc[DM];
#pragma unroll
        for (int r = 0; r < DM; r++) lc[r] = (row && r < dim && r > l) ? W[r + l * dim] : 0.0;
        double y = 0.0, pinv = 0.0;
        if (row)
        {
            double rhs = A.rhs[lo + l];
            if (i < Ns - 2) rhs = rhs - scv;
            y = rhs;
            pinv = A.inv[lo + l];
        }
        /* y <- L^-T y (dtrsv_ltn) */
#pragma unroll
        for (int r = DM - 1; r >= 0; r--)
            if (r < dim)
            {
                const double yi = __shfl_sync(FULL, y, r) * __shfl_sync(FULL, pinv, r);
                if (l == r) y = yi;
                else if (l < r) y -= lc[r] * yi;
            }
        if (row) v.dlam[lo + l] = y;
        if (i > 0)
        {
            const int dimP = nu * A.cn[i - 1];
            const double *CJL = A.JL + A.jlo[i - 1]; /* dim x dimP */
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < DM; a++)
            {
                const double ya = __shfl_sync(FULL, y, a);
                if (a < dim && mine && l < dimP) t += CJL[a + l * dim] * ya;
            }
            scv = t;
        }
    }
    __syncwarp();
}

/* ------------------------------------------------------------------------------------------
 * Twisted ("burn at both ends") factorisation of the same chain for small blocks: warp 0 eliminates
 * the pairs 0 .. m-1 from the left exactly as sd_jay_chain_small does, warp 1 eliminates the pairs
 * Ns-2 .. m+1 from the RIGHT (the same recursion on the transposed coupling blocks), both hand their
 * Schur complement and right-hand side to the middle pair m, which warp 0 factors; then the two warps
 * substitute outwards from the middle at the same time.  The dependent chain is half as long (config 2:
 * 80 block eliminations per Newton iteration on one warp were 60 % of a solve).  Same pivot rule and
 * Levenberg-Marquardt policy per block; the elimination ORDER differs from the reference's for the right
 * half, so the iterates differ in the last bits (parity: iteration counts and 1e-10 on the goldens).
 *   DIR = +1: block i, successor n = i + 1, coupling M = JL_i L_i^-T            (rows of n, dim_n x dim_i), in JL_i's slot
 *   DIR = -1: block i, successor n = i - 1, coupling M = JL_n' L_i^-T           (rows of n, dim_n x dim_i), in JL_n's slot
 *             (element (l, c) of M at [c + l * dim_i], the position of the transposed entry)
 * ---------------------------------------------------------------------------------------- */
template <int DM>
struct JayBlock
{   /* one block elimination in registers: lane l owns row l (see sd_jay_chain_small) */
    double w[DM], inv[DM], y;
};

template <int DM>
__device__ __forceinline__ int sd_jay_factor_block(const SdView &v, const JayArena &A, int i, int l, bool mine, bool has_carry,
                                                   const double (&car)[DM], double scv_in, double scv_in2, JayBlock<DM> &B)
{
    const int nu = v.d.nu, dim = nu * A.cn[i], lo = A.lo[i];
    const int regType = v.o.regType;
    const double regTol = v.o.regTol, regValue = v.o.regValue;
    double *W = A.JD0 + A.jdo[i];
    const double *W1 = A.JD1 + A.jdo[i];
    const bool row = mine && l < dim;
    double w0[DM];
    int nreg = 0;
#pragma unroll
    for (int c = 0; c < DM; c++)
    {
        double val = 0.0;
        if (row && c < dim)
        {
            val = (has_carry && l >= c) ? car[c] : 0.0;
            val += W[l + c * dim];
            val += W1[l + c * dim];
            if (regType == 1 && c == l) val += regValue;
        }
        B.w[c] = w0[c] = val;
        B.inv[c] = 0.0;
    }
    double y = 0.0;
    if (row)
    {
        double rhs = A.rhs[lo + l];
        rhs = rhs + scv_in;
        rhs = rhs + scv_in2;
        y = -1.0 * rhs;
    }
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++)
    {
        bool small = false;
#pragma unroll
        for (int j = 0; j < DM; j++)
        {
            if (j < dim)
            {
                double sacc = B.w[j];
#pragma unroll
                for (int t = 0; t < j; t++) sacc -= B.w[t] * __shfl_sync(FULL, B.w[t], j);
                double piv = __shfl_sync(FULL, sacc, j), pinv;
                tqb_pivot(piv, pinv);
                if (l == j) B.w[j] = piv;
                else if (l > j) B.w[j] = sacc * pinv;
                B.inv[j] = pinv;
                small |= (piv <= regTol);
            }
        }
        if (!(regType == 2 && attempt == 0 && small)) { nreg += (regType == 1 || attempt == 1); break; }
#pragma unroll
        for (int c = 0; c < DM; c++) B.w[c] = w0[c] + ((c == l && row) ? regValue : 0.0);
    }
#pragma unroll
    for (int j = 0; j < DM; j++)
        if (j < dim)
        {
            const double yj = __shfl_sync(FULL, y, j) * B.inv[j];
            if (l == j) y = yj;
            else if (l > j && l < dim) y -= B.w[j] * yj;
        }
    B.y = y;
    if (row)
    {
        A.rhs[lo + l] = y;
#pragma unroll
        for (int c = 0; c < DM; c++)
            if (c < dim && c <= l) W[l + c * dim] = B.w[c];
    }
    if (mine)
    {
#pragma unroll
        for (int c = 0; c < DM; c++)
            if (c < dim && c == l) A.inv[lo + c] = B.inv[c];
    }
    return nreg;
}

/* elimination half: blocks i0, i0 + DIR, ..., up to (not including) the middle block m; leaves the carried Schur
 * complement (row l, columns <= l, at car_out[l * DM + c]) and right-hand side (scv_out[l]) for the middle block */
template <int DM, int DIR>
__device__ void sd_jay_half_eliminate(const SdView &v, const JayArena &A, int lane, int i0, int m, double *car_out, double *scv_out)
{
    const int nu = v.d.nu;
    const int l = lane < DM ? lane : DM - 1;
    const bool mine = lane < DM;
    int nreg = 0;
    double car[DM], scv = 0.0;
#pragma unroll
    for (int c = 0; c < DM; c++) car[c] = 0.0;
#pragma unroll 1
    for (int i = i0; i != m; i += DIR)
    {
        JayBlock<DM> B;
        nreg += sd_jay_factor_block<DM>(v, A, i, l, mine, i != i0, car, i != i0 ? scv : 0.0, 0.0, B);
        const int dim = nu * A.cn[i], n = i + DIR, dimN = nu * A.cn[n];
        double *JL = A.JL + A.jlo[DIR > 0 ? i : n];
        const bool rowN = mine && l < dimN;
        double jl[DM];
#pragma unroll
        for (int c = 0; c < DM; c++) jl[c] = (rowN && c < dim) ? (DIR > 0 ? JL[l + c * dimN] : JL[c + l * dim]) : 0.0;
        /* M <- (coupling) L^-T (dtrsm_rltn): row l, column by column */
#pragma unroll
        for (int j = 0; j < DM; j++)
            if (j < dim)
            {
                double t = jl[j];
#pragma unroll
                for (int c = 0; c < j; c++) t -= jl[c] * __shfl_sync(FULL, B.w[c], j);
                jl[j] = t * B.inv[j];
                if (rowN) { if (DIR > 0) JL[l + j * dimN] = jl[j]; else JL[j + l * dim] = jl[j]; }
            }
        /* carried Schur complement (lower) and right-hand side for the successor */
#pragma unroll
        for (int b = 0; b < DM; b++)
        {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < DM; c++)
            {
                const double o = __shfl_sync(FULL, jl[c], b);
                if (c < dim) t += jl[c] * o;
            }
            car[b] = 0.0 - t;
        }
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < DM; c++)
        {
            const double yc = __shfl_sync(FULL, B.y, c);
            if (c < dim) t += jl[c] * yc;
        }
        scv = t;
    }
    const int mJ = v.d.mJ; /* the scratch regions hold mJ x mJ / mJ doubles */
    if (mine && l < mJ)
    {
#pragma unroll
        for (int c = 0; c < DM; c++)
            if (c < mJ) car_out[l * mJ + c] = car[c];
        scv_out[l] = scv;
    }
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);
}

/* the middle block: both carries in, factor, solve, and the right-hand-side parts its two neighbours subtract */
template <int DM>
__device__ void sd_jay_middle(const SdView &v, const JayArena &A, int lane, int m, const double *carL, const double *scvL,
                              const double *carR, const double *scvR, double *outL, double *outR)
{
    const int nu = v.d.nu, Ns = v.d.Ns, mJ = v.d.mJ;
    const int l = lane < DM ? lane : DM - 1;
    const bool mine = lane < DM, in = mine && l < mJ;
    const int dim = nu * A.cn[m], lo = A.lo[m];
    double car[DM];
#pragma unroll
    for (int c = 0; c < DM; c++) car[c] = (in && c < mJ) ? carL[l * mJ + c] + carR[l * mJ + c] : 0.0;
    const double sl = in ? scvL[l] : 0.0, sr = in ? scvR[l] : 0.0;
    JayBlock<DM> B;
    const int nreg = sd_jay_factor_block<DM>(v, A, m, l, mine, true, car, sl, sr, B);
    if (nreg > 0 && lane == 0) atomicAdd(&v.info[3], nreg);
    __syncwarp(); /* the factor of the block is in the arena */
    /* dlam_m = L^-T y (dtrsv_ltn), as the substitution halves do it */
    const double *W = A.JD0 + A.jdo[m];
    const bool row = mine && l < dim;
    double lc[DM];
#pragma unroll
    for (int r = 0; r < DM; r++) lc[r] = (row && r < dim && r > l) ? W[r + l * dim] : 0.0;
    double y = row ? B.y : 0.0;
    const double pinv = row ? A.inv[lo + l] : 0.0;
#pragma unroll
    for (int r = DM - 1; r >= 0; r--)
        if (r < dim)
        {
            const double yi = __shfl_sync(FULL, y, r) * __shfl_sync(FULL, pinv, r);
            if (l == r) y = yi;
            else if (l < r) y -= lc[r] * yi;
        }
    if (row) v.dlam[lo + l] = y;
    /* what the neighbours subtract from their right-hand sides: M' dlam_m */
    double tl = 0.0, tr = 0.0;
    if (m > 0)
    {   /* left neighbour m-1: M = JL_{m-1} L_{m-1}^-T (dim_m x dim_{m-1}), column l */
        const int dimP = nu * A.cn[m - 1];
        const double *CJL = A.JL + A.jlo[m - 1];
#pragma unroll
        for (int a = 0; a < DM; a++)
        {
            const double ya = __shfl_sync(FULL, y, a);
            if (a < dim && mine && l < dimP) tl += CJL[a + l * dim] * ya;
        }
    }
    if (m < Ns - 2)
    {   /* right neighbour m+1: M(l', c) at JL_m[c + l' * dim_{m+1}], lane c */
        const int dimP = nu * A.cn[m + 1];
        const double *CJL = A.JL + A.jlo[m];
#pragma unroll
        for (int a = 0; a < DM; a++)
        {
            const double ya = __shfl_sync(FULL, y, a);
            if (a < dim && mine && l < dimP) tr += CJL[l + a * dimP] * ya;
        }
    }
    __syncwarp();
    if (in) { outL[l] = tl; outR[l] = tr; }
}

/* substitution half: from the block next to the middle outwards to i0 */
template <int DM, int DIR>
__device__ void sd_jay_half_substitute(const SdView &v, const JayArena &A, int lane, int i0, int m, const double *scv_in)
{
    const int nu = v.d.nu;
    const int l = lane < DM ? lane : DM - 1;
    const bool mine = lane < DM;
    double scv = (mine && l < v.d.mJ) ? scv_in[l] : 0.0;
#pragma unroll 1
    for (int i = m - DIR; ; i -= DIR)
    {
        const int dim = nu * A.cn[i], lo = A.lo[i];
        const double *W = A.JD0 + A.jdo[i];
        const bool row = mine && l < dim;
        double lc[DM];
#pragma unroll
        for (int r = 0; r < DM; r++) lc[r] = (row && r < dim && r > l) ? W[r + l * dim] : 0.0;
        double y = 0.0, pinv = 0.0;
        if (row)
        {
            y = A.rhs[lo + l] - scv;
            pinv = A.inv[lo + l];
        }
#pragma unroll
        for (int r = DM - 1; r >= 0; r--)
            if (r < dim)
            {
                const double yi = __shfl_sync(FULL, y, r) * __shfl_sync(FULL, pinv, r);
                if (l == r) y = yi;
                else if (l < r) y -= lc[r] * yi;
            }
        if (row) v.dlam[lo + l] = y;
        if (i == i0) break;
        const int p = i - DIR, dimP = nu * A.cn[p];
        /* DIR = +1: M_p = JL_p L_p^-T (dim_i x dim_p) in JL_p: t_l = sum_a JL_p[a + l dim_i] y_a
         * DIR = -1: M_p (rows of i, dim_i x dim_p) in JL_i at [c + l' dim_p]: t_c = sum_l' JL_i[c + l' dim_p] y_l' */
        const double *CJL = A.JL + A.jlo[DIR > 0 ? p : i];
        double t = 0.0;
#pragma unroll
        for (int a = 0; a < DM; a++)
        {
            const double ya = __shfl_sync(FULL, y, a);
            if (a < dim && mine && l < dimP) t += (DIR > 0 ? CJL[a + l * dim] : CJL[l + a * dimP]) * ya;
        }
        scv = t;
    }
}

/* ------------------------------------------------------------------------------------------
 * dmu_i = Lambda_i^-1 (-res_i + Zbar (dlam_{i-1} - dlam_i))  (calculate_delta_mu :1077-1205), then
 * the derivative of the unconstrained stage solutions along (dmu, dlam): along the step the stage
 * gradients are affine in the step length, so a line-search trial is a pure vector pass.
 * Returns this lane's part of res'dmu.
 * ---------------------------------------------------------------------------------------- */
template <int NX>
__device__ double sd_delta_mu_scenario(const SdView &v, int i, const WarpScratch &S, int lane)
{
    const int nx = NX > 0 ? NX : v.d.nx, nu = v.d.nu, Nh = v.d.Nh, Ns = v.d.Ns, Nr = v.d.Nr, nv = nx + nu;
    const size_t ik0 = (size_t) i * Nh;
    double *dmu = v.dmu + ik0 * nx;
    for (int e = lane; e < Nh * nx; e += 32)
    {   /* right-hand side -res + Zbar (dlam_{i-1} - dlam_i), :1095-1127 */
        const int k = e / nx, a = e - k * nx;
        const double r = v.res[ik0 * nx + e];
        double val;
        if (k < Nr)
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
        dmu[e] = val;
    }
    __syncwarp();
    sd_banded_solve<NX>(v, i, dmu, dmu, S, lane);
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
        double err = v.fixed_nx == 4 ? sd_build<4>(v, tm) : v.fixed_nx == 8 ? sd_build<8>(v, tm)
                   : v.fixed_nx == 16 ? sd_build<16>(v, tm) : sd_build<0>(v, tm), dummy = 0.0;
        team_reduce2(v, tm, err, dummy, v.o.termCondition == 2, s_red);
        PHASE(1);
        if (v.o.termCondition == 1) err = sqrt(err);
        if (err < v.o.tol) { status = 0; break; }

        /* the scenario-local chains are instantiated for the state dimensions of the reference's
         * examples (spring-mass nx = 4, linear_chain nm = 4 / 8 -> nx = 8 / 16): constant trip counts and
         * index arithmetic, same operation order as the runtime-dimension instance */
        if (v.fixed_nx == 4) for (int i = tm.gw; i < Ns; i += tm.W) sd_factor_scenario<4>(v, i, S, tm.lane);
        else if (v.fixed_nx == 8) for (int i = tm.gw; i < Ns; i += tm.W) sd_factor_scenario<8>(v, i, S, tm.lane);
        else if (v.fixed_nx == 16) for (int i = tm.gw; i < Ns; i += tm.W) sd_factor_scenario<16>(v, i, S, tm.lane);
        else for (int i = tm.gw; i < Ns; i += tm.W) sd_factor_scenario<0>(v, i, S, tm.lane);
        team_sync();
        PHASE(2);
        if (Ns > 1)
        {
            if (tm.rank == 0)
            {
                const JayArena A = jay_arena(v, v.jay_in_smem ? smem : v.jarena);
                sd_jay_gather(v, A, tm);
                __syncthreads();
                if (v.d.mJ <= 8 && !v.no_small_jay && !v.serial_jay && Ns - 1 >= 5 && tm.wpb >= 2)
                {   /* twisted factorisation: warps 0 and 1 eliminate towards the middle pair m */
                    const int m = (Ns - 1) / 2;
                    double *carL = A.car, *carR = A.w0, *scvL = A.scv, *scvR = A.sy; /* scratch of the generic walk, free here */
                    if (v.d.mJ <= 4)
                    {
                        if (tm.wib == 0) sd_jay_half_eliminate<4, +1>(v, A, tm.lane, 0, m, carL, scvL);
                        else if (tm.wib == 1) sd_jay_half_eliminate<4, -1>(v, A, tm.lane, Ns - 2, m, carR, scvR);
                        __syncthreads();
                        if (tm.wib == 0) sd_jay_middle<4>(v, A, tm.lane, m, carL, scvL, carR, scvR, scvL, scvR);
                        __syncthreads();
                        if (tm.wib == 0) sd_jay_half_substitute<4, +1>(v, A, tm.lane, 0, m, scvL);
                        else if (tm.wib == 1) sd_jay_half_substitute<4, -1>(v, A, tm.lane, Ns - 2, m, scvR);
                    }
                    else
                    {
                        if (tm.wib == 0) sd_jay_half_eliminate<8, +1>(v, A, tm.lane, 0, m, carL, scvL);
                        else if (tm.wib == 1) sd_jay_half_eliminate<8, -1>(v, A, tm.lane, Ns - 2, m, carR, scvR);
                        __syncthreads();
                        if (tm.wib == 0) sd_jay_middle<8>(v, A, tm.lane, m, carL, scvL, carR, scvR, scvL, scvR);
                        __syncthreads();
                        if (tm.wib == 0) sd_jay_half_substitute<8, +1>(v, A, tm.lane, 0, m, scvL);
                        else if (tm.wib == 1) sd_jay_half_substitute<8, -1>(v, A, tm.lane, Ns - 2, m, scvR);
                    }
                }
                else if (tm.wib == 0)
                {
                    if (v.d.mJ <= 4 && !v.no_small_jay) sd_jay_chain_small<4>(v, A, tm.lane);
                    else if (v.d.mJ <= 8 && !v.no_small_jay) sd_jay_chain_small<8>(v, A, tm.lane);
                    else sd_jay_chain(v, A, tm.lane);
                }
            }
            team_sync();
        }
        PHASE(3);
        double dot = 0.0;
        if (v.fixed_nx == 4) for (int i = tm.gw; i < Ns; i += tm.W) dot += sd_delta_mu_scenario<4>(v, i, S, tm.lane);
        else if (v.fixed_nx == 8) for (int i = tm.gw; i < Ns; i += tm.W) dot += sd_delta_mu_scenario<8>(v, i, S, tm.lane);
        else if (v.fixed_nx == 16) for (int i = tm.gw; i < Ns; i += tm.W) dot += sd_delta_mu_scenario<16>(v, i, S, tm.lane);
        else for (int i = tm.gw; i < Ns; i += tm.W) dot += sd_delta_mu_scenario<0>(v, i, S, tm.lane);
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
    double *d_in, *d_out, *d_work, *d_jarena;
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
    const size_t per_warp = sd_per_warp_doubles(d->nx, d->nK);
    const size_t jay = sd_jay_doubles(*d);
    const size_t lim = (size_t) smem_max - 4096;
    int wpb = 16;
    while (wpb > 1 && (size_t) wpb * per_warp * 8 > lim) wpb >>= 1;
    if ((size_t) wpb * per_warp * 8 > lim || d->nx > 32)
    {
        snprintf(g_sd_err, sizeof g_sd_err, "SDUNES blocks (nx=%d, nu*Nr=%d) exceed what one SM holds (nx <= 32 and the block scratch must fit in shared memory)",
                 d->nx, d->nK);
        fprintf(stderr, "[treeqp_b200] %s\n", g_sd_err);
        return 1;
    }
    /* team size: one warp per scenario for the chains, enough threads for the element-wise phases */
    int nb = (d->Ns + wpb - 1) / wpb;
    const long long elems = (long long) d->SK * d->nx * d->nx;
    const int nb_el = (int) ((elems + 32LL * wpb * 4 - 1) / (32LL * wpb * 4));
    if (nb_el > nb) nb = nb_el;
    if (nb > 8) nb = 8;
    if (nb < 1) nb = 1;
    D->wpb = wpb;
    D->nblocks = nb;
    size_t sm = (size_t) wpb * per_warp;
    v.jay_in_smem = jay * 8 <= lim;
    {
        const char *e = getenv("TREEQP_B200_GENERIC_JAY");
        v.no_small_jay = (e && e[0] == '1') ? 1 : 0;
        const char *e2 = getenv("TREEQP_B200_SERIAL_JAY");
        v.serial_jay = (e2 && e2[0] == '1') ? 1 : 0;
    }
    {
        /* TREEQP_B200_GENERIC_NX: 1 = runtime-dimension chains, 0 = the nx = 4 / 8 / 16 instances;
         * unset = TQB_SD_FIXED_NX_DEFAULT */
        const char *e = getenv("TREEQP_B200_GENERIC_NX");
        const int fixed = (e && (e[0] == '0' || e[0] == '1')) ? e[0] == '0' : TQB_SD_FIXED_NX_DEFAULT;
        v.fixed_nx = (fixed && (d->nx == 4 || d->nx == 8 || d->nx == 16)) ? d->nx : 0;
    }
    if (v.jay_in_smem && jay > sm) sm = jay;
    D->smem = sm * 8;
    v.per_warp = (int) per_warp;
    CK(cudaMalloc(&D->d_jarena, sizeof(double) * jay));
    v.jarena = D->d_jarena;
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
    cudaFree(D->d_int); cudaFree(D->d_in); cudaFree(D->d_out); cudaFree(D->d_work); cudaFree(D->d_info); cudaFree(D->d_jarena);
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
