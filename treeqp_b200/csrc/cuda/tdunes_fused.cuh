/*
 * tdunes_fused.cuh -- compile-time-dimension fast path of the leaves-to-root / root-to-leaves
 * sweeps for uniform trees (every non-root node has NX states, every parent NU inputs, all
 * parents of one depth have the same number of children MD; D = MD*NX <= 32).
 *
 * k_build_factor<NX,NU,MD,PNX,ROOT>: one warp per parent block p of one depth.  Fuses what the
 * reference does in four places so that the dual Hessian block never travels through HBM:
 *   - dual gradient of the kids   res_k = b_k - x_k + A_k x_p + B_k u_p      (dual_Newton_tree.c:519-539)
 *   - dual Hessian block          W_p = C diag(P_p,Rt_p) C' + blkdiag(P_kids) (:551-615, clipping.c:264-355)
 *                                 Ut_p = -P_p A_kids'                         (:581-589)
 *     minus the kids' Schur complements (the dsyrk/dgemv updates of :726-732, received through
 *     small per-node buffers instead of read-modify-write on W in HBM)
 *   - Cholesky with BLASFEO's pivot rule + Levenberg-Marquardt policy         (dual_Newton_common.c:35-78)
 *   - y = L^-1 resMod, Utilde = Ut L^-T, S_p = Utilde Utilde', s_p = Utilde y  (:712-732)
 *   - roots: dlam = L^-T y                                                    (:748-758)
 * Layout inside the warp: lane r owns ROW r of the augmented matrix [W_p ; Ut_p] in registers
 * (rows >= 32 form a second register row-set on the low lanes); the factorisation is
 * right-looking: at step j the finished column j is published to shared memory and every lane
 * downdates its own row with broadcast reads.  The right-hand side is kept transposed (lane i
 * holds resMod_i) and substituted on the fly with one shuffle per column.
 *
 * k_backsolve_fused<NX,MD>: dlam_p = L_p^-T (y_p - Utilde_p' dlam_dad[p]), lane t owns COLUMN t of L.
 *
 * HBM format of a factored block: L packed by rows (row i at i(i+1)/2, i+1 entries), Utilde
 * row-major NX x D, inverse diagonal in x-space (Linv), Schur complement S_p (NX x NX, entry (a, b)
 * with a >= b at b*NX + a: the consumer's lanes own rows and read a column step by step) and s_p
 * in x-space (rsub).
 */
#pragma once
#include "tdunes_kernels.cuh"

#ifndef TQB_DMMA_MIN_D
#define TQB_DMMA_MIN_D 8 /* smallest block dimension built on the tensor cores */
#endif

/* Programmatic dependent launch (sm_90+): the level kernels of one sweep are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so level l starts while level l+1 still
 * runs, does everything that does not depend on it (descriptor + A/B loads, gradient, the
 * C diag(P,Rt) C' build; for the substitution: the loads of L and Utilde) and only then waits.
 * Every warp waits before it exits (keeps completion transitive along the chain), and releases
 * its own dependents right after its wait. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__host__ __device__ constexpr int tqb_ld2(int n) { return n + ((6 - (n & 3)) & 3); } /* smallest >= n with ld%4==2 */

/* level descriptor: everything a warp needs to find its block, one 48-byte record per parent in
 * level order (built at create, device.cu) -- one dependent load instead of the chain
 * lvl_nodes -> kid0/xoff/... -> xoff[kid0]/aoff[kid0] */
struct LvlDesc
{
    int g, k0, xo, uo;      /* node, first kid, offsets of the node in x-/u-space */
    int xb, ao, bo, wo;     /* x-space offset of the block (= xoff[k0]), A/B of the kids, packed L */
    int uto, tree, md, r1;  /* Ut offset, owning tree, number of kids */
};

template <int NX, int NU, int MD, int PNX, bool ROOT>
struct FusedCfg
{
    static constexpr int D = NX * MD;
    static constexpr int NCOL = PNX + NU;           /* columns of C = [A_kids B_kids] */
    static constexpr int NUT = ROOT ? 0 : NX;       /* rows of Ut */
    static constexpr int R = D + NUT;               /* rows of the augmented matrix */
    static constexpr int NROWL = NUT < 32 - D ? NUT : 32 - D; /* Ut rows that get a lane of their own */
    static constexpr int NCOLU = NUT - NROWL;       /* Ut rows held column-wise by the W lanes */
    static constexpr int RL = D + NROWL;            /* lanes that own a row */
    static constexpr int LDC = tqb_ld2(NCOL > 0 ? NCOL : 2);
    static constexpr int LDR = tqb_ld2(R);
    /* shared memory per warp, in doubles: Ms (D x LDR; its head doubles as Cp, D x LDC, during the
     * build) | p (NCOL) | z (NCOL) | ys (D) (+ 2 D spare) */
    /* FP64 tensor-core (DMMA m8n8k4) build of C diag(p) C' and of the Schur complement for the
     * larger blocks: 8x8 output tiles, operands in fragment layout straight from HBM / shared memory */
    static constexpr bool DMMA = D >= TQB_DMMA_MIN_D;
    static constexpr int TR = (D + 7) / 8;          /* row tiles of W */
    static constexpr int KS = (NCOL + 3) / 4;       /* k-steps of the build */
    static constexpr int LDCD = 20 + 16 * ((NCOL > 20 ? NCOL - 20 : 0) + 15) / 16 * (NCOL > 20); /* >= NCOL, = 4 mod 16 */
    static constexpr int SM_M0 = D * LDR > D * LDC ? D * LDR : D * LDC;
    static constexpr int SM_M = (DMMA && D * LDCD > SM_M0) ? D * LDCD : SM_M0;
    static constexpr int SMEM = SM_M + 2 * (NCOL + 2) + 3 * D + 2;
};

/* D(8x8) += A(8x4) * B(4x8), FP64: lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2(T%4) + {0,1}] */
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

/* -DTQB_BF_TIMING: thread 0 of CTA 0 leaves clock64 / globaltimer stamps of the phases of its block
 * (slot = 0 for the root level, else 1 + log2(gridDim.x)); read back with tqb_debug_bf_timing */
#ifdef TQB_BF_TIMING
__device__ long long g_bf_t[32][16];
__device__ __forceinline__ void bf_stamp(bool root, int i)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        const int slot = root ? 0 : 32 - __clz(gridDim.x);
        g_bf_t[slot][i] = clock64();
        g_bf_t[slot][8 + i] = gt;
    }
}
#define BF_STAMP(i) bf_stamp(ROOT, i)
#else
#define BF_STAMP(i)
#endif

#ifndef TQB_BF_MINBLOCKS
#define TQB_BF_MINBLOCKS 4
#endif
/* small blocks need few registers: ask for more resident CTAs (the kernel is latency bound) */
#ifndef TQB_BF_MINBLOCKS_D8
#define TQB_BF_MINBLOCKS_D8 6
#endif
#ifndef TQB_BF_MINBLOCKS_D16
#define TQB_BF_MINBLOCKS_D16 5
#endif
__host__ __device__ constexpr int tqb_bf_minblocks(int d) { return d <= 8 ? TQB_BF_MINBLOCKS_D8 : d <= 16 ? TQB_BF_MINBLOCKS_D16 : TQB_BF_MINBLOCKS; }

/* one parent block by one warp: `desc` = index of its level descriptor, `Ms` = the warp's shared
 * memory (Cfg::SMEM doubles), `wait_here` = this block is the first thing the warp does in a kernel
 * launched with programmatic stream serialisation (the wait sits between build and factorisation) */
template <int NX, int NU, int MD, int PNX, bool ROOT>
__device__ __forceinline__ void bf_block(const View &v, size_t desc, int kids_are_parents, bool wait_here, double *Ms, int lane)
{
    using Cfg = FusedCfg<NX, NU, MD, PNX, ROOT>;
    constexpr int D = Cfg::D, NCOL = Cfg::NCOL, LDC = Cfg::LDC, LDR = Cfg::LDR;
    constexpr int NROWL = Cfg::NROWL, NCOLU = Cfg::NCOLU, RL = Cfg::RL;
    static_assert(D <= 32, "block dimension must fit one warp");
    static_assert((D & 1) == 0, "even block dimension required (vectorised rows)");
    const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * desc;
    const int4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
    const int g = d0.x, k0 = d0.y, xo = d0.z, uo = d0.w, xb = d1.x;
    if (v.status[d2.y] != TQB_RUNNING)
    {
        if (wait_here) { pdl_wait(); pdl_release(); }
        return;
    }
    BF_STAMP(1);

    /* Ms[j][r] = column j of the factored augmented matrix */
    double *Cp = Ms;                              /* Cp[K][c] = p_c * C[K][c] (build phase only) */
    double *ps = Ms + Cfg::SM_M;                  /* [P_p ; Rt_p] */
    double *zs = ps + NCOL + 2;                   /* [x_p ; u_p] */
    double *ys = zs + NCOL + 2;                   /* y = L^-1 resMod */

    const double *Ag = v.in.A + d1.y, *Bg = v.in.B + d1.z;
    const bool isW = lane < D;
    const int kid = isW ? lane / NX : 0, ii = isW ? lane - kid * NX : 0;

    double wA[D], utc[NCOLU > 0 ? NCOLU : 1];
    double myinv = 0.0, my_y = 0.0, my_d = 0.0, rr = 0.0;
    bool small = false;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++)
    {
        const double shift = (v.o.regType == 1 || attempt == 1) ? v.o.regValue : 0.0;
        double res = 0.0;
        if constexpr (Cfg::DMMA)
        {
            /* ---- tensor-core build: W = C diag(p) C' tile by tile (lower tiles only) ----
             * fragment (i, ks) of this lane = C[8i + gid][4ks + tig]; the B operand of tile column n
             * is the same fragment of row tile n scaled by p (= Cp), so one pass over C in HBM feeds
             * both operands, the gradient and the Ut rows */
            constexpr int TR = Cfg::TR, KS = Cfg::KS, LDCD = Cfg::LDCD;
            const int gid = lane >> 2, tig = lane & 3;
            double pf[KS], zf[KS], af[TR][KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
            {
                const int c = 4 * ks + tig;
                pf[ks] = zf[ks] = 0.0;
                if (c < NCOL)
                {
                    pf[ks] = c < PNX ? v.P[xo + c] : v.Rt[uo + c - PNX];
                    zf[ks] = c < PNX ? v.x[xo + c] : v.u[uo + c - PNX];
                }
            }
#pragma unroll
            for (int i = 0; i < TR; i++)
            {
                const int r = 8 * i + gid, kr = r / NX, ir = r - kr * NX;
#pragma unroll
                for (int ks = 0; ks < KS; ks++)
                {
                    const int c = 4 * ks + tig;
                    double val = 0.0;
                    if (r < D && c < NCOL) val = c < PNX ? Ag[kr * NX * PNX + ir + c * NX] : Bg[kr * NX * NU + ir + (c - PNX) * NX];
                    af[i][ks] = val;
                }
            }
            if (isW) res = v.in.b[xb + lane] - v.x[xb + lane];
            __syncwarp(); /* previous attempt's readers of Ms are done */
            double acc[TR * (TR + 1) / 2][2];
#pragma unroll
            for (int i = 0; i < TR; i++)
            {   /* C z as one more tile column (z in column 0): the tensor core accumulates k in
                 * ascending order with one rounding per term, like the scalar fma chain */
                const int r = 8 * i + gid;
                double sg = 0.0, sg1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++)
                {
                    dmma884(sg, sg1, af[i][ks], gid == 0 ? zf[ks] : 0.0);
                    if (r < D && 4 * ks + tig < NCOL) Cp[r * LDCD + 4 * ks + tig] = af[i][ks] * pf[ks];
                }
                if (tig == 0 && r < D) ys[r] = sg;
            }
#pragma unroll
            for (int i = 0; i < TR; i++)
#pragma unroll
                for (int n = 0; n <= i; n++)
                {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) dmma884(c0, c1, af[i][ks], af[n][ks] * pf[ks]);
                    acc[i * (i + 1) / 2 + n][0] = c0;
                    acc[i * (i + 1) / 2 + n][1] = c1;
                }
            __syncwarp();
            if (isW)
            {
                res += ys[lane];
                v.res[xb + lane] = res;
                if (NCOLU > 0)
                {
#pragma unroll
                    for (int cc = 0; cc < NCOLU; cc++) utc[cc] = -Cp[lane * LDCD + NROWL + cc];
                }
            }
            else if (!ROOT && lane < RL)
            {   /* Ut rows: M[D+c][K] = -P_p[c] * A[K][c] = -Cp[K][c] */
#pragma unroll
                for (int K = 0; K < D; K++) wA[K] = -Cp[K * LDCD + (lane - D)];
            }
            if (attempt == 0)
            {   /* contribution to the termination norm (calculate_error_in_residuals, :412-442) */
                double part = (v.o.termCondition == 2) ? fabs(res) : res * res;
                part = (v.o.termCondition == 2) ? warp_max(part) : warp_sum(part);
                if (lane == 0) v.nva[g] = part;
            }
            __syncwarp(); /* Cp is dead: the W tiles go to the same words, row-major with leading dimension LDR */
#pragma unroll
            for (int i = 0; i < TR; i++)
#pragma unroll
                for (int n = 0; n <= i; n++)
                    if (8 * i + gid < D && 8 * n + 2 * tig < D) /* D is even: a pair never straddles the row end */
                        *reinterpret_cast<double2 *>(Ms + (8 * i + gid) * LDR + 8 * n + 2 * tig) =
                            make_double2(acc[i * (i + 1) / 2 + n][0], acc[i * (i + 1) / 2 + n][1]);
            __syncwarp();
            if (isW)
            {
                const double pk = v.P[xb + lane];
#pragma unroll
                for (int K = 0; K < D; K += 2)
                {
                    const double2 t2 = *reinterpret_cast<const double2 *>(Ms + lane * LDR + K);
                    wA[K] = t2.x + ((K == lane) ? pk + shift : 0.0);
                    wA[K + 1] = t2.y + ((K + 1 == lane) ? pk + shift : 0.0);
                }
            }
        }
        else
        {
        for (int c = lane; c < NCOL; c += 32)
        {
            ps[c] = c < PNX ? v.P[xo + c] : v.Rt[uo + c - PNX];
            zs[c] = c < PNX ? v.x[xo + c] : v.u[uo + c - PNX];
        }
        /* own row of C straight from HBM: kid k = lane/NX, A_k is NX x PNX column-major */
        double crow[NCOL > 0 ? NCOL : 1];
#pragma unroll
        for (int c = 0; c < NCOL; c++)
        {
            double val = 0.0;
            if (isW) val = c < PNX ? Ag[kid * NX * PNX + ii + c * NX] : Bg[kid * NX * NU + ii + (c - PNX) * NX];
            crow[c] = val;
        }
        if (isW) res = v.in.b[xb + lane] - v.x[xb + lane];
        __syncwarp();
        /* gradient of this lane's multiplier component, and its scaled C row for the others */
        if (isW)
        {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NCOL; c++)
            {
                s += crow[c] * zs[c];
                Cp[lane * LDC + c] = crow[c] * ps[c];
            }
            res += s;
            v.res[xb + lane] = res;
        }
        if (attempt == 0)
        {   /* contribution to the termination norm (calculate_error_in_residuals, :412-442) */
            double part = (v.o.termCondition == 2) ? fabs(res) : res * res;
            part = (v.o.termCondition == 2) ? warp_max(part) : warp_sum(part);
            if (lane == 0) v.nva[g] = part;
        }
        __syncwarp();

        /* ---- build the rows: W rows on lanes < D, Ut rows on the lanes after them ---- */
#pragma unroll
        for (int K = 0; K < D; K++)
        {
            double s = 0.0;
            const double2 *cpk = reinterpret_cast<const double2 *>(Cp + K * LDC);
#pragma unroll
            for (int c2 = 0; c2 < NCOL / 2; c2++)
            {
                const double2 t2 = cpk[c2];
                s += crow[2 * c2] * t2.x;
                s += crow[2 * c2 + 1] * t2.y;
            }
            if (NCOL & 1) s += crow[NCOL - 1] * Cp[K * LDC + NCOL - 1];
            wA[K] = s;
        }
        if (NCOLU > 0)
        {   /* Ut rows without a lane: element (row, K = lane) = -P_p[row] * A[K][row] */
#pragma unroll
            for (int cc = 0; cc < NCOLU; cc++) utc[cc] = -(crow[NROWL + cc] * ps[NROWL + cc]);
        }
        if (isW)
        {
            const double pk = v.P[xb + lane];
#pragma unroll
            for (int K = 0; K < D; K++) wA[K] += (K == lane) ? pk + shift : 0.0;
        }
        }
        /* ---- everything below needs the kids' level ---- */
        BF_STAMP(2);
        if (attempt == 0 && wait_here) { pdl_wait(); pdl_release(); }
        BF_STAMP(3);
        rr = res - ((kids_are_parents && isW) ? v.rsub[xb + lane] : 0.0);
        if (isW)
        {
            if (kids_are_parents)
            {   /* minus the Schur complement of the kid this row belongs to (lower triangle) */
                const double *Sk = v.schur + (size_t) (k0 + kid) * NX * NX + ii;
#pragma unroll
                for (int kk = 0; kk < NX; kk++)
                {
                    const double sv = (kk <= ii) ? Sk[kk * NX] : 0.0; /* S[ii][kk], stored by columns */
#pragma unroll
                    for (int m = 0; m < MD; m++) wA[m * NX + kk] -= (m == kid) ? sv : 0.0;
                }
            }
        }
        else if (!Cfg::DMMA && !ROOT && lane < RL)
        {   /* Ut rows: M[D+c][K] = -P_p[c] * A[K][c] = -Cp[K][c] */
#pragma unroll
            for (int K = 0; K < D; K++) wA[K] = -Cp[K * LDC + (lane - D)];
        }
        small = false;
        __syncwarp(); /* Cp is dead, Ms may be written */
        BF_STAMP(4);
        /* ---- right-looking Cholesky of the augmented matrix ----
         * Software-pipelined by one column: the only chain that has to be serial is
         *   diagonal of row j -> shuffle -> rsqrt core (100 cycles) -> scale -> diagonal of row j+1,
         * so lane j+1 forms its next diagonal `dn` from its OWN scaled entry (the same fma the trailing
         * update would do with the broadcast value), and the trailing update with column j-1 (shared-memory
         * broadcast + D fmas) is issued inside the latency of column j's pivot instead of in front of it. */
        double dn = wA[0], ljp = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++)
        {
            /* lanes <= j carry don't-care values in column j and in rr from here on: the diagonal,
             * its inverse and y are published by lane 0 (no per-lane selects in the loop) */
            double piv = __shfl_sync(FULL, dn, j), inv;
            if (j > 0)
            {   /* trailing update with column j-1 (published before the barrier that ended step j-1) */
                const double *col = Ms + (j - 1) * LDR;
                /* wA[k] -= L[lane][j-1] * L[k][j-1], k >= j: column broadcast from shared memory, two at a time */
                if (j & 1) wA[j] -= ljp * col[j];
#pragma unroll
                for (int k = (j + 1) & ~1; k < D; k += 2)
                {
                    const double2 l2 = *reinterpret_cast<const double2 *>(col + k);
                    wA[k] -= ljp * l2.x;
                    wA[k + 1] -= ljp * l2.y;
                }
                if (NCOLU > 0)
                {   /* Ut[c][lane] -= Utilde[c][j-1] * L[lane][j-1] on the lanes still to come */
                    if constexpr ((NCOLU & 1) == 0 && (RL & 1) == 0)
                    {
#pragma unroll
                        for (int cc = 0; cc < NCOLU; cc += 2)
                        {
                            const double2 u2 = *reinterpret_cast<const double2 *>(col + RL + cc);
                            utc[cc] -= u2.x * ljp;
                            utc[cc + 1] -= u2.y * ljp;
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int cc = 0; cc < NCOLU; cc++) utc[cc] -= col[RL + cc] * ljp;
                    }
                }
            }
            tqb_pivot_spec(piv, inv);
            small |= (piv <= v.o.regTol);
            wA[j] *= inv;
            if (j + 1 < D) dn = fma(-wA[j], wA[j], wA[j + 1]); /* lane j+1: its diagonal after column j */
            /* forward substitution of the right-hand side on the fly (dtrsv_lnn) */
            const double yj = __shfl_sync(FULL, rr, j) * inv;
            rr = fma(-wA[j], yj, rr);
            if (lane < RL) Ms[j * LDR + lane] = wA[j];
            if (NCOLU > 0)
            {   /* lane j finishes its elements of the column-wise Ut rows (its registers are dead afterwards) */
                double us[NCOLU > 0 ? NCOLU : 1];
#pragma unroll
                for (int cc = 0; cc < NCOLU; cc++) us[cc] = utc[cc] * inv;
                if (lane == j)
                {
                    if constexpr ((NCOLU & 1) == 0 && (RL & 1) == 0)
                    {
#pragma unroll
                        for (int cc = 0; cc < NCOLU; cc += 2)
                            *reinterpret_cast<double2 *>(Ms + j * LDR + RL + cc) = make_double2(us[cc], us[cc + 1]);
                    }
                    else
                    {
#pragma unroll
                        for (int cc = 0; cc < NCOLU; cc++) Ms[j * LDR + RL + cc] = us[cc];
                    }
                }
            }
            /* piv, inv and yj are warp-uniform; lane j keeps them (the kernel is bound by L1/shared
             * wavefronts: three uniform stores per step cost more than three selects) */
            if (lane == j) { my_y = yj; my_d = piv; myinv = inv; }
            ljp = wA[j];
            __syncwarp();
        }
        BF_STAMP(5);
        if (!(v.o.regType == 2 && attempt == 0 && small)) break;
        if (lane == 0) atomicAdd(&v.counters[2], 1);
        __syncwarp();
    }

    /* ---- results to HBM ---- */
    /* The factor leaves through shared memory: Ms[j][r] already holds column j of [L ; Utilde] (only
     * the diagonal is missing), so row r of the packed L and the rows of Utilde go out as contiguous
     * warp stores.  Writing them from the row-owning lanes' registers costs one sector per lane and
     * store instruction (ncu: 242 + 96 L1 requests per block against 726 shared-memory wavefronts for
     * the whole factorisation; the kernel is bound by L1 throughput). */
    if (isW)
    {
        rr = my_y;
        ys[lane] = my_y; /* read below by the lanes that form s_p = Utilde y */
        Ms[lane * LDR + lane] = my_d;
        v.Linv[xb + lane] = myinv;
    }
    __syncwarp();
    {
        double *Lg = v.W + d1.w;
#pragma unroll
        for (int r = 0; r < D; r++)
            if (lane <= r) Lg[(r * (r + 1)) / 2 + lane] = Ms[lane * LDR + r];
    }
    BF_STAMP(6);
    if (!ROOT)
    {
        if (isW) v.dlam[xb + lane] = rr;
        double *Up = v.Ut + d2.x;
#pragma unroll
        for (int e0 = 0; e0 < NX * D; e0 += 32)
        {   /* Utilde row-major NX x D: element a*D + j is Ms[j][D + a] */
            const int e = e0 + lane, a = e / D, j = e - a * D;
            if (e < NX * D) Up[e] = Ms[j * LDR + D + a];
        }
        /* Schur complement for the parent's block: S = Utilde Utilde' (lower), s = Utilde y.
         * Lower-triangle entries enumerated as a (NX/2) x (NX+1) rectangle: rows p and NX-1-p share a line */
        double *Sg = v.schur + (size_t) g * NX * NX;
        constexpr int NE = (NX % 2 == 0) ? (NX / 2) * (NX + 1) : NX * NX;
        if constexpr (Cfg::DMMA)
        {   /* tensor-core S = Utilde Utilde': fragment (i, ks) = Utilde[8i + gid][4ks + tig] is both operands */
            constexpr int TRU = (NX + 7) / 8, KSU = (D + 3) / 4;
            const int gid = lane >> 2, tig = lane & 3;
            double uf[TRU][KSU];
#pragma unroll
            for (int i = 0; i < TRU; i++)
#pragma unroll
                for (int ks = 0; ks < KSU; ks++)
                    uf[i][ks] = (8 * i + gid < NX && 4 * ks + tig < D) ? Ms[(4 * ks + tig) * LDR + D + 8 * i + gid] : 0.0;
#pragma unroll
            for (int i = 0; i < TRU; i++)
#pragma unroll
                for (int n = 0; n <= i; n++)
                {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KSU; ks++) dmma884(c0, c1, uf[i][ks], uf[n][ks]);
                    const int a = 8 * i + gid, b = 8 * n + 2 * tig;
                    if (a < NX && b < NX) Sg[b * NX + a] = c0;
                    if (a < NX && b + 1 < NX) Sg[(b + 1) * NX + a] = c1;
                }
        }
        else
        for (int e = lane; e < NE; e += 32)
        {
            int a, b;
            if (NX % 2 == 0)
            {
                const int p = e / (NX + 1), q = e - p * (NX + 1);
                a = q <= p ? p : NX - 1 - p;
                b = q <= p ? q : q - p - 1;
            }
            else
            {
                a = e / NX; b = e - a * NX;
                if (b > a) continue;
            }
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; j++) s += Ms[j * LDR + D + a] * Ms[j * LDR + D + b];
            Sg[b * NX + a] = s;
        }
        __syncwarp(); /* ys */
        for (int a = lane; a < NX; a += 32)
        {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; j++) s += Ms[j * LDR + D + a] * ys[j];
            v.rsub[xo + a] = s;
        }
    }
    else
    {   /* root: backward substitution dlam = L^-T y; lane t owns column t of L = Ms[t][*] */
#pragma unroll
        for (int i = D - 1; i >= 0; i--)
        {
            const double yi = __shfl_sync(FULL, rr, i) * __shfl_sync(FULL, myinv, i);
            if (lane == i) rr = yi;
            else if (lane < i) rr -= Ms[lane * LDR + i] * yi;
        }
        if (isW) v.dlam[xb + lane] = rr;
    }
    BF_STAMP(7);
}

/* ---- blocked variant for block dimensions D >= 16 -----------------------------------------------
 * Same arithmetic per entry, in the same order, as bf_block (so the iterates are bit-identical), but
 * organised as a LEFT-LOOKING Cholesky on 8-column panels with the augmented matrix [W ; Ut] resident in
 * shared memory, column-major with leading dimension LDM (LDM % 8 == 4: the 8 x 4 DMMA fragment loads are
 * bank-conflict free):
 *   build    W = C diag(p) C' on the FP64 tensor cores, tiles stored straight into Ms; the Ut rows are
 *            the scaled fragments themselves (no staging of C in shared memory)
 *   panel p  (a) M[:, p] -= L[:, 0..p) L[p, 0..p)'  as DMMA m8n8k4 products accumulating INTO the tile
 *                (k ascending, one rounding per term: what the right-looking scalar update does);
 *            (b) the panel's 8 columns are pulled into registers, one ROW per lane (rows >= 32 form a
 *                second, short row-set) and factorised right-looking inside the panel only: per column
 *                one shuffle, one pivot, at most 7 fmas per row instead of D;
 *   the column broadcast inside a panel is software-pipelined behind the pivot chain as in bf_block.
 * Against bf_block<12,4,2>: shared-memory wavefronts per block 726 -> ~330, DFMA instructions in the
 * factorisation 24 x 28 -> 24 x 8 + 20 DMMA. */
template <int NX, int NU, int MD, int PNX, bool ROOT>
struct BlockedCfg
{
    static constexpr int D = NX * MD;
    static constexpr int NCOL = PNX + NU;
    static constexpr int NUT = ROOT ? 0 : NX;
    static constexpr int R = D + NUT;                 /* rows of [W ; Ut] */
    static constexpr int RT = (R + 7) / 8;            /* row tiles */
    static constexpr int NP = (D + 7) / 8;            /* column panels */
    static constexpr int NA = R < 32 ? R : 32;        /* rows of the first row-set (lane = row) */
    static constexpr int NB = R > 32 ? R - 32 : 0;    /* rows of the second row-set */
    static constexpr int LDM = 8 * RT + 4;            /* % 8 == 4 */
    static constexpr int TR = (D + 7) / 8;
    static constexpr int KS = (NCOL + 3) / 4;
    static constexpr int SM_M = D * LDM;
    static constexpr int SMEM = SM_M + 2 * D + 4;     /* Ms | ys (D) | spare */
};

/* 16-byte asynchronous copies global -> shared (LDGSTS): the factor of an unchanged block is staged under its loads */
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

/* Re-use of unchanged blocks -- the GPU form of the reference's checkLastActiveSet (dual_Newton_tree.c:334-371, 563,
 * 703).  The reference skips the re-factorisation of the SUFFIX of the parent numbering behind the last changed active
 * set; measured on config 3 that suffix is 0.2 % of the blocks for most iterations, while 94-100 % of ALL blocks have an
 * unchanged subtree from the third Newton iteration on (profiles/r02_active_set_change_config3.txt).  A block whose
 * own, kids' and descendants' active sets did not move has the W_p, L_p, Utilde_p and Schur complement of the last
 * iteration bit for bit (same inputs, same operations): only the right-hand side is redone -- the gradient res, the
 * termination norm, y = L^-1 (res - s_kids) with L from HBM (staged by cp.async under the gradient's loads), and
 * s_p = Utilde y -- with the operations of the full path in the same order, so the iterates do not change by a bit.
 * `try_skip` (wide levels only): returns true when the block was handled here; false: the caller runs the full path
 * (`waited` tells it whether the programmatic-launch wait has been done). */
template <int NX, int NU, int MD, int PNX, bool ROOT>
__device__ __forceinline__ bool bf_block_reuse(const View &v, const int4 d0, const int4 d1, const int4 d2, int kids_are_parents,
                                               bool wait_here, bool &waited, double *Ms, int lane, int tree_status)
{
    using Cfg = BlockedCfg<NX, NU, MD, PNX, ROOT>;
    constexpr int D = Cfg::D, NCOL = Cfg::NCOL, TR = Cfg::TR, KS = Cfg::KS;
    constexpr int NL = D * (D + 1) / 2;          /* packed L */
    constexpr int LDU = D + 2;                   /* staged Utilde rows: even (16-byte copies), spreads the NX reader lanes over the banks */
    constexpr int OFFU = (NL + 1) & ~1;
    static_assert(!ROOT, "roots always take the full path");
    static_assert((D & 1) == 0 && OFFU + NX * LDU <= Cfg::SM_M, "staging must fit the block's shared memory");
    const int g = d0.x, k0 = d0.y, xo = d0.z, uo = d0.w, xb = d1.x;
    /* Everything is requested at once, before the flags are looked at: the flags of the node and its kids
     * (k_stage_fused<0> finished before this sweep started), last iteration's factor (written by this block one sweep
     * ago, nobody touches it in between), and the gradient's operands.  A changed block (a few per cent) wastes the
     * staging copies. */
    const int c_own = (lane == 0) ? v.chg[g] : (lane <= MD ? v.chg[k0 + lane - 1] : 0);
    const int was_reg = (lane == 0) ? v.breg[g] : 0; /* with the other requests: a load behind the wait is one more round trip */
    double *sL = Ms, *sU = Ms + OFFU, *ys = Ms + Cfg::SM_M;
    {
        const double *Lg = v.W + d1.w, *Up = v.Ut + d2.x;
        for (int e = lane; e < (NL + 1) / 2; e += 32) cp_async16(sL + 2 * e, Lg + 2 * e); /* odd NL: one double of the next block rides along */
        for (int e = lane; e < NX * D / 2; e += 32)
        {
            const int a = (2 * e) / D, j = 2 * e - a * D;
            cp_async16(sU + a * LDU + j, Up + 2 * e);
        }
    }
    const double *Ag = v.in.A + d1.y, *Bg = v.in.B + d1.z;
    const bool isW = lane < D;
    const int gid = lane >> 2, tig = lane & 3;
    const double inv = isW ? v.Linv[xb + lane] : 0.0;
    double res = 0.0;
    {   /* gradient: exactly the tile column of the full path (C z on the tensor cores, k ascending) */
        double zf[KS], af[TR][KS];
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
        {
            const int cc = 4 * ks + tig;
            zf[ks] = 0.0;
            if (cc < NCOL) zf[ks] = cc < PNX ? v.x[xo + cc] : v.u[uo + cc - PNX];
        }
#pragma unroll
        for (int i = 0; i < TR; i++)
        {
            const int r = 8 * i + gid, kr = r / NX, ir = r - kr * NX;
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
            {
                const int cc = 4 * ks + tig;
                double val = 0.0;
                if (r < D && cc < NCOL) val = cc < PNX ? Ag[kr * NX * PNX + ir + cc * NX] : Bg[kr * NX * NU + ir + (cc - PNX) * NX];
                af[i][ks] = val;
            }
        }
        if (isW) res = v.in.b[xb + lane] - v.x[xb + lane];
        if (__any_sync(FULL, c_own != 0) || tree_status != TQB_RUNNING)
        {   /* own or a kid's active set moved: full path (its loads now hit L1); a finished tree: the caller returns */
            cp_async_commit_wait_all();
            __syncwarp();
            return false;
        }
#pragma unroll
        for (int i = 0; i < TR; i++)
        {
            const int r = 8 * i + gid;
            double sg = 0.0, sg1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) dmma884(sg, sg1, af[i][ks], gid == 0 ? zf[ks] : 0.0);
            if (tig == 0 && r < D) ys[r] = sg;
        }
        __syncwarp();
        if (isW) res += ys[lane];
    }
    if (wait_here) { pdl_wait(); pdl_release(); }
    waited = true;
    /* a kid whose block was re-factored in this sweep hands up a new Schur complement: full path */
    const int c_kid = (kids_are_parents && lane < MD) ? v.bdirty[k0 + lane] : 0;
    double rr = res - ((kids_are_parents && isW) ? v.rsub[xb + lane] : 0.0);
    cp_async_commit_wait_all();
    if (__any_sync(FULL, c_kid != 0)) { __syncwarp(); return false; }
    if (isW) v.res[xb + lane] = res;
    {
        double part = (v.o.termCondition == 2) ? fabs(res) : res * res;
        part = (v.o.termCondition == 2) ? warp_max(part) : warp_sum(part);
        if (lane == 0)
        {
            v.nva[g] = part;
            v.bdirty[g] = 0;
            if (was_reg) atomicAdd(&v.counters[2], 1); /* the re-used factor is a regularised one: counted as the full path would */
        }
    }
    __syncwarp(); /* staged factor visible to the whole warp */
    double lrow[D - 1];
#pragma unroll
    for (int j = 0; j < D - 1; j++) lrow[j] = (isW && j < lane) ? sL[(lane * (lane + 1)) / 2 + j] : 0.0;
    double my_y = 0.0;
#pragma unroll
    for (int J = 0; J < D; J++)
    {   /* forward substitution, column by column as the factorisation does it on the fly */
        const double yj = __shfl_sync(FULL, rr, J) * __shfl_sync(FULL, inv, J);
        if (J < D - 1) rr = fma(-lrow[J], yj, rr);
        if (lane == J) my_y = yj;
    }
    if (isW) { ys[lane] = my_y; v.dlam[xb + lane] = my_y; }
    __syncwarp();
    for (int a = lane; a < NX; a += 32)
    {   /* s = Utilde y */
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) s += sU[a * LDU + j] * ys[j];
        v.rsub[xo + a] = s;
    }
    __syncwarp(); /* Ms is the next block's */
    return true;
}

template <int NX, int NU, int MD, int PNX, bool ROOT>
__device__ __forceinline__ void bf_block_lb(const View &v, size_t desc, int kids_are_parents, bool wait_here, double *Ms, int lane,
                                            bool try_skip = false)
{
    using Cfg = BlockedCfg<NX, NU, MD, PNX, ROOT>;
    constexpr int D = Cfg::D, NCOL = Cfg::NCOL, RT = Cfg::RT, NP = Cfg::NP, NA = Cfg::NA, NB = Cfg::NB, LDM = Cfg::LDM;
    constexpr int TR = Cfg::TR, KS = Cfg::KS;
    static_assert(D <= 32 && (D & 1) == 0, "block dimension: even, at most one warp of rows");
    const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * desc;
    const int4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
    const int g = d0.x, k0 = d0.y, xo = d0.z, uo = d0.w, xb = d1.x;
    /* requested here, looked at behind the first batch of loads: a test right away would put one more dependent
     * round trip (descriptor -> status -> data) in front of every block */
    const int tree_status = v.status[d2.y];
    BF_STAMP(1);
    if constexpr (!ROOT)
    {
        if (try_skip)
        {
            bool waited = false;
            if (bf_block_reuse<NX, NU, MD, PNX, ROOT>(v, d0, d1, d2, kids_are_parents, wait_here, waited, Ms, lane, tree_status)) return;
            if (waited) wait_here = false;
        }
    }
    double *ys = Ms + Cfg::SM_M;                  /* C z, later y = L^-1 resMod */
    const double *Ag = v.in.A + d1.y, *Bg = v.in.B + d1.z;
    const bool isW = lane < D;
    const int kid = isW ? lane / NX : 0, ii = isW ? lane - kid * NX : 0;
    const int gid = lane >> 2, tig = lane & 3;

    double myinv = 0.0, my_y = 0.0, my_d = 0.0, rr = 0.0;
    bool small = false;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++)
    {
        const double shift = (v.o.regType == 1 || attempt == 1) ? v.o.regValue : 0.0;
        double res = 0.0;
        {
            /* ---- tensor-core build (see bf_block): fragment (i, ks) = C[8i + gid][4ks + tig] ---- */
            double pf[KS], zf[KS], af[TR][KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
            {
                const int c = 4 * ks + tig;
                pf[ks] = zf[ks] = 0.0;
                if (c < NCOL)
                {
                    pf[ks] = c < PNX ? v.P[xo + c] : v.Rt[uo + c - PNX];
                    zf[ks] = c < PNX ? v.x[xo + c] : v.u[uo + c - PNX];
                }
            }
#pragma unroll
            for (int i = 0; i < TR; i++)
            {
                const int r = 8 * i + gid, kr = r / NX, ir = r - kr * NX;
#pragma unroll
                for (int ks = 0; ks < KS; ks++)
                {
                    const int c = 4 * ks + tig;
                    double val = 0.0;
                    if (r < D && c < NCOL) val = c < PNX ? Ag[kr * NX * PNX + ir + c * NX] : Bg[kr * NX * NU + ir + (c - PNX) * NX];
                    af[i][ks] = val;
                }
            }
            if (isW) res = v.in.b[xb + lane] - v.x[xb + lane];
            if (tree_status != TQB_RUNNING)
            {   /* a finished tree of the forest: nothing to do (warp-uniform) */
                if (wait_here) { pdl_wait(); pdl_release(); }
                return;
            }
            __syncwarp(); /* previous attempt's readers of Ms / ys are done */
#pragma unroll
            for (int i = 0; i < TR; i++)
            {   /* C z as one more tile column (z in column 0) */
                const int r = 8 * i + gid;
                double sg = 0.0, sg1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) dmma884(sg, sg1, af[i][ks], gid == 0 ? zf[ks] : 0.0);
                if (tig == 0 && r < D) ys[r] = sg;
            }
#pragma unroll
            for (int i = 0; i < TR; i++)
            {
#pragma unroll
                for (int n = 0; n <= i; n++)
                {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) dmma884(c0, c1, af[i][ks], af[n][ks] * pf[ks]);
                    /* W tile (i, n) -> Ms, column-major */
                    const int r = 8 * i + gid, c = 8 * n + 2 * tig;
                    if (r < D && c < D) { Ms[c * LDM + r] = c0; Ms[(c + 1) * LDM + r] = c1; }
                }
                if (!ROOT)
                {   /* Ut[c][K] = -P_p[c] * A[K][c]: the scaled fragments, row D + c of column K = 8i + gid */
                    const int K = 8 * i + gid;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++)
                        if (4 * ks < PNX)
                        {
                            const int c = 4 * ks + tig;
                            if (K < D && c < PNX) Ms[K * LDM + D + c] = -(af[i][ks] * pf[ks]);
                        }
                }
            }
            __syncwarp();
            if (isW)
            {
                res += ys[lane];
                v.res[xb + lane] = res;
                Ms[lane * LDM + lane] += v.P[xb + lane] + shift;
            }
            if (attempt == 0)
            {   /* contribution to the termination norm (calculate_error_in_residuals, :412-442) */
                double part = (v.o.termCondition == 2) ? fabs(res) : res * res;
                part = (v.o.termCondition == 2) ? warp_max(part) : warp_sum(part);
                if (lane == 0) v.nva[g] = part;
            }
        }
        /* ---- everything below needs the kids' level ---- */
        BF_STAMP(2);
        if (attempt == 0 && wait_here) { pdl_wait(); pdl_release(); }
        BF_STAMP(3);
        rr = res - ((kids_are_parents && isW) ? v.rsub[xb + lane] : 0.0);
        if (kids_are_parents && isW)
        {   /* minus the Schur complement of the kid this row belongs to (lower triangle, row `lane`) */
            const double *Sk = v.schur + (size_t) (k0 + kid) * NX * NX + ii;
            double sv[NX];
#pragma unroll
            for (int kk = 0; kk < NX; kk++) sv[kk] = (kk <= ii) ? Sk[kk * NX] : 0.0; /* S[ii][kk], stored by columns */
#pragma unroll
            for (int kk = 0; kk < NX; kk++)
                if (kk <= ii) Ms[(kid * NX + kk) * LDM + lane] -= sv[kk];
        }
        small = false;
        __syncwarp();
        BF_STAMP(4);
        /* ---- left-looking blocked Cholesky of the augmented matrix ---- */
#pragma unroll
        for (int p = 0; p < NP; p++)
        {
            constexpr int PWMAX = 8;
            const int PW = (D - 8 * p) < PWMAX ? (D - 8 * p) : PWMAX; /* compile-time after unrolling */
            if (p > 0)
            {   /* (a) panel update on the tensor cores: tile (i, p) -= sum_k L[8i.., k] L[8p.., k] */
                double bfr[2 * (NP - 1)];
#pragma unroll
                for (int ks = 0; ks < 2 * p; ks++) bfr[ks] = Ms[(4 * ks + tig) * LDM + 8 * p + gid];
#pragma unroll
                for (int i = p; i < RT; i++)
                {
                    const int r = 8 * i + gid, c = 8 * p + 2 * tig;
                    double c0 = 0.0, c1 = 0.0;
                    if (c < D) { c0 = Ms[c * LDM + r]; c1 = Ms[(c + 1) * LDM + r]; }
#pragma unroll
                    for (int ks = 0; ks < 2 * p; ks++)
                    {
                        const double a = (i == p) ? bfr[ks] : Ms[(4 * ks + tig) * LDM + r];
                        dmma884(c0, c1, -a, bfr[ks]);
                    }
                    if (c < D) { Ms[c * LDM + r] = c0; Ms[(c + 1) * LDM + r] = c1; }
                }
                __syncwarp();
            }
            /* (b) the panel in registers, one row per lane */
            double pa[PWMAX], pb[PWMAX];
#pragma unroll
            for (int k = 0; k < PWMAX; k++)
            {
                pa[k] = pb[k] = 0.0;
                if (k < PW)
                {
                    if (NA == 32 || lane < NA) pa[k] = Ms[(8 * p + k) * LDM + lane]; /* lanes without a row compute on zeros */
                    if (NB > 0 && lane < NB) pb[k] = Ms[(8 * p + k) * LDM + 32 + lane];
                }
            }
            /* (c) right-looking inside the panel, pivot chain decoupled from the column broadcast */
            double dn = pa[0], ljp = 0.0, ljpb = 0.0;
#pragma unroll
            for (int j = 0; j < PWMAX; j++)
            {
                if (j < PW)
                {
                    const int J = 8 * p + j;
                    double piv = __shfl_sync(FULL, dn, J), inv;
                    if (j > 0)
                    {   /* trailing update inside the panel with column J-1 (published before the last barrier) */
                        const double *col = Ms + (J - 1) * LDM + 8 * p;
                        if (j & 1)
                        {
                            const double l = col[j];
                            pa[j] -= ljp * l;
                            if (NB > 0) pb[j] -= ljpb * l;
                        }
#pragma unroll
                        for (int k = (j + 1) & ~1; k < PW; k += 2)
                        {
                            const double2 l2 = *reinterpret_cast<const double2 *>(col + k);
                            pa[k] -= ljp * l2.x;
                            pa[k + 1] -= ljp * l2.y;
                            if (NB > 0) { pb[k] -= ljpb * l2.x; pb[k + 1] -= ljpb * l2.y; }
                        }
                    }
                    tqb_pivot_spec(piv, inv);
                    small |= (piv <= v.o.regTol);
                    pa[j] *= inv;
                    if (NB > 0) pb[j] *= inv;
                    if (j + 1 < PW) dn = fma(-pa[j], pa[j], pa[j + 1]); /* lane J+1: its diagonal after column J */
                    /* forward substitution of the right-hand side on the fly (dtrsv_lnn) */
                    const double yj = __shfl_sync(FULL, rr, J) * inv;
                    rr = fma(-pa[j], yj, rr);
                    if (NA == 32 || lane < NA) Ms[J * LDM + lane] = pa[j];
                    if (NB > 0 && lane < NB) Ms[J * LDM + 32 + lane] = pb[j];
                    if (lane == J) { my_y = yj; my_d = piv; myinv = inv; }
                    ljp = pa[j]; ljpb = pb[j];
                    __syncwarp();
                }
            }
        }
        BF_STAMP(5);
        if (!(v.o.regType == 2 && attempt == 0 && small))
        {
            if (lane == 0 && v.bdirty) { v.bdirty[g] = 1; v.breg[g] = attempt; }
            break;
        }
        if (lane == 0) atomicAdd(&v.counters[2], 1);
        __syncwarp();
    }

    /* ---- results to HBM (as bf_block: Ms[j][r] holds column j of [L ; Utilde] but the diagonal) ---- */
    if (isW)
    {
        rr = my_y;
        ys[lane] = my_y;
        Ms[lane * LDM + lane] = my_d;
        v.Linv[xb + lane] = myinv;
    }
    __syncwarp();
    {
        double *Lg = v.W + d1.w;
#pragma unroll
        for (int r = 0; r < D; r++)
            if (lane <= r) Lg[(r * (r + 1)) / 2 + lane] = Ms[lane * LDM + r];
    }
    BF_STAMP(6);
    if (!ROOT)
    {
        if (isW) v.dlam[xb + lane] = rr;
        double *Up = v.Ut + d2.x;
#pragma unroll
        for (int e0 = 0; e0 < NX * D; e0 += 32)
        {   /* Utilde row-major NX x D: element a*D + j is Ms[j][D + a] */
            const int e = e0 + lane, a = e / D, j = e - a * D;
            if (e < NX * D) Up[e] = Ms[j * LDM + D + a];
        }
        /* Schur complement for the parent's block: S = Utilde Utilde' (lower), s = Utilde y */
        double *Sg = v.schur + (size_t) g * NX * NX;
        constexpr int TRU = (NX + 7) / 8, KSU = (D + 3) / 4;
        double uf[TRU][KSU];
#pragma unroll
        for (int i = 0; i < TRU; i++)
#pragma unroll
            for (int ks = 0; ks < KSU; ks++)
                uf[i][ks] = (8 * i + gid < NX && 4 * ks + tig < D) ? Ms[(4 * ks + tig) * LDM + D + 8 * i + gid] : 0.0;
#pragma unroll
        for (int i = 0; i < TRU; i++)
#pragma unroll
            for (int n = 0; n <= i; n++)
            {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < KSU; ks++) dmma884(c0, c1, uf[i][ks], uf[n][ks]);
                const int a = 8 * i + gid, b = 8 * n + 2 * tig;
                if (a < NX && b < NX) Sg[b * NX + a] = c0;
                if (a < NX && b + 1 < NX) Sg[(b + 1) * NX + a] = c1;
            }
        for (int a = lane; a < NX; a += 32)
        {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < D; j++) s += Ms[j * LDM + D + a] * ys[j];
            v.rsub[xo + a] = s;
        }
    }
    else
    {   /* root: backward substitution dlam = L^-T y; lane t owns column t of L = Ms[t][*] */
#pragma unroll
        for (int i = D - 1; i >= 0; i--)
        {
            const double yi = __shfl_sync(FULL, rr, i) * __shfl_sync(FULL, myinv, i);
            if (lane == i) rr = yi;
            else if (lane < i) rr -= Ms[lane * LDM + i] * yi;
        }
        if (isW) v.dlam[xb + lane] = rr;
    }
    BF_STAMP(7);
}

#ifndef TQB_BF_BLOCKED_MIN_D
#define TQB_BF_BLOCKED_MIN_D 16 /* smallest block dimension that takes the blocked left-looking variant */
#endif
template <int NX, int NU, int MD, int PNX, bool ROOT>
struct BfSel
{
    static constexpr bool BLOCKED = NX * MD >= TQB_BF_BLOCKED_MIN_D;
    static constexpr int SMEM = BLOCKED ? BlockedCfg<NX, NU, MD, PNX, ROOT>::SMEM : FusedCfg<NX, NU, MD, PNX, ROOT>::SMEM;
};

template <int NX, int NU, int MD, int PNX, bool ROOT>
__global__ void __launch_bounds__(128, tqb_bf_minblocks(NX * MD)) k_build_factor(View v, int lvl_begin, int count, int kids_are_parents, int pre_ok, int try_skip)
{
    using Sel = BfSel<NX, NU, MD, PNX, ROOT>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    BF_STAMP(0);
    /* pre_ok = 0: the launch before this one wrote data read by the build phase (k_stage<0> before
     * the bottom level, the cut-node exchange of a sharded solve): wait first */
    if (!pre_ok) { pdl_wait(); pdl_release(); }
    if (w >= count) { pdl_wait(); pdl_release(); return; }
    if constexpr (Sel::BLOCKED)
        bf_block_lb<NX, NU, MD, PNX, ROOT>(v, (size_t) (lvl_begin + w), kids_are_parents, true, smem + (size_t) wib * Sel::SMEM, lane, try_skip != 0);
    else
        bf_block<NX, NU, MD, PNX, ROOT>(v, (size_t) (lvl_begin + w), kids_are_parents, true, smem + (size_t) wib * Sel::SMEM, lane);
}

/* Chain segments (consecutive depths whose parents have ONE kid each: the scenario chains below the
 * robust horizon of a multistage tree): warp w walks chain w from its deepest parent up to the
 * top of the segment in ONE launch instead of one launch per depth.  Breadth-first numbering keeps
 * the order of the chains from depth to depth, so block w of depth l-1 is the parent of block w
 * of depth l; the Schur complement travels through v.schur / v.rsub (written and re-read by the
 * same warp, L1/L2 hits).  Same arithmetic per block as k_build_factor<NX,NU,1,NX,false>. */
template <int NX, int NU>
__global__ void __launch_bounds__(128, tqb_bf_minblocks(NX)) k_build_factor_chain(View v, int top_begin, int count, int nlev, int bottom_kids_are_parents, int pre_ok)
{
    using Cfg = FusedCfg<NX, NU, 1, NX, false>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (!pre_ok) { pdl_wait(); pdl_release(); }
    if (w >= count) { pdl_wait(); pdl_release(); return; }
    double *Ms = smem + (size_t) wib * Cfg::SMEM;
#pragma unroll 1
    for (int l = nlev - 1; l >= 0; l--)
    {
        const bool first = l == nlev - 1;
        bf_block<NX, NU, 1, NX, false>(v, (size_t) top_begin + (size_t) l * count + w, first ? bottom_kids_are_parents : 1, first, Ms, lane);
        __syncwarp(); /* S, s of this block are visible to the whole warp; Ms may be reused */
    }
}

/* ---- chains of small blocks: TWO chains per warp ------------------------------------------------
 * A chain block of dimension D = NX <= 8 has 2 NX <= 16 rows in its augmented matrix [W ; Ut]: half
 * of a warp idles in bf_block and the chain kernel is bound by instruction issue.  Here lanes 0-15
 * own the rows of chain 2w, lanes 16-31 those of chain 2w+1 (sub-lane sl = lane & 15, shuffles of
 * width 16, one shared-memory region per half).  Same operations in the same order per block as
 * bf_block<NX,NU,1,NX,false>; the build is the scalar fma chain, which the tensor-core build
 * reproduces bit for bit (tools/micro/dmma_order.cu).  A half whose chain does not exist or has
 * finished computes on zeros and stores nothing. */
template <int NX, int NU>
struct HalfCfg
{
    static constexpr int D = NX, NCOL = NX + NU, R = 2 * NX;
    static constexpr int LDC = tqb_ld2(NCOL), LDR = tqb_ld2(R);
    static constexpr int SM_M = D * LDR > D * LDC ? D * LDR : D * LDC;
    static constexpr int SMEM = SM_M + 2 * (NCOL + 2) + 3 * D + 2; /* doubles per half-warp */
};

__device__ __forceinline__ double half_sum16(double v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double half_max16(double v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

template <int NX, int NU>
__device__ __forceinline__ void bf_block_half(const View &v, size_t desc, bool valid, int kids_are_parents, bool wait_here, double *Ms, int lane)
{
    using Cfg = HalfCfg<NX, NU>;
    constexpr int D = Cfg::D, NCOL = Cfg::NCOL, LDC = Cfg::LDC, LDR = Cfg::LDR, PNX = NX;
    static_assert(2 * NX <= 16 && (NX & 1) == 0, "two chains per warp need 2 NX <= 16, NX even");
    const int sl = lane & 15;
    int4 d0 = make_int4(0, 0, 0, 0), d1 = d0, d2 = d0;
    bool act = false;
    if (valid)
    {
        const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * desc;
        d0 = __ldg(dp); d1 = __ldg(dp + 1); d2 = __ldg(dp + 2);
        act = v.status[d2.y] == TQB_RUNNING;
    }
    if (!__any_sync(FULL, act))
    {
        if (wait_here) { pdl_wait(); pdl_release(); }
        return;
    }
    const int g = d0.x, k0 = d0.y, xo = d0.z, uo = d0.w, xb = d1.x;
    double *Cp = Ms;
    double *ps = Ms + Cfg::SM_M, *zs = ps + NCOL + 2, *ys = zs + NCOL + 2;
    const double *Ag = v.in.A + d1.y, *Bg = v.in.B + d1.z;
    const bool isW = sl < D, isU = sl >= D && sl < 2 * D; /* Ut row sl - D */
    const bool ldW = act && isW;

    double wA[D];
    double rr = 0.0, my_y = 0.0, my_d = 0.0, my_inv = 0.0;
    bool small = false;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++)
    {
        /* the half that does not need the retry repeats its first attempt (same shift, same result) */
        const double shift = (v.o.regType == 1 || (attempt == 1 && small)) ? v.o.regValue : 0.0;
        double res = 0.0;
        for (int c = sl; c < NCOL; c += 16)
        {
            ps[c] = act ? (c < PNX ? v.P[xo + c] : v.Rt[uo + c - PNX]) : 0.0;
            zs[c] = act ? (c < PNX ? v.x[xo + c] : v.u[uo + c - PNX]) : 0.0;
        }
        double crow[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; c++)
        {
            double val = 0.0;
            if (ldW) val = c < PNX ? Ag[sl + c * NX] : Bg[sl + (c - PNX) * NX];
            crow[c] = val;
        }
        if (ldW) res = v.in.b[xb + sl] - v.x[xb + sl];
        __syncwarp();
        if (isW)
        {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NCOL; c++)
            {
                s += crow[c] * zs[c];
                Cp[sl * LDC + c] = crow[c] * ps[c];
            }
            res += s;
            if (act) v.res[xb + sl] = res;
        }
        if (attempt == 0)
        {   /* contribution to the termination norm (calculate_error_in_residuals, :412-442) */
            double part = (v.o.termCondition == 2) ? fabs(res) : res * res;
            part = (v.o.termCondition == 2) ? half_max16(part) : half_sum16(part);
            if (sl == 0 && act) v.nva[g] = part;
        }
        __syncwarp();
#pragma unroll
        for (int K = 0; K < D; K++)
        {
            double s = 0.0;
            const double2 *cpk = reinterpret_cast<const double2 *>(Cp + K * LDC);
#pragma unroll
            for (int c2 = 0; c2 < NCOL / 2; c2++)
            {
                const double2 t2 = cpk[c2];
                s += crow[2 * c2] * t2.x;
                s += crow[2 * c2 + 1] * t2.y;
            }
            if (NCOL & 1) s += crow[NCOL - 1] * Cp[K * LDC + NCOL - 1];
            wA[K] = s;
        }
        if (isW)
        {
            const double pk = act ? v.P[xb + sl] : 0.0;
#pragma unroll
            for (int K = 0; K < D; K++) wA[K] += (K == sl) ? pk + shift : 0.0;
        }
        /* ---- everything below needs the kids' level ---- */
        if (attempt == 0 && wait_here) { pdl_wait(); pdl_release(); }
        rr = res - ((kids_are_parents && ldW) ? v.rsub[xb + sl] : 0.0);
        if (isW)
        {
            if (kids_are_parents && act)
            {   /* minus the Schur complement of the kid (lower triangle) */
                const double *Sk = v.schur + (size_t) k0 * NX * NX + sl;
#pragma unroll
                for (int kk = 0; kk < NX; kk++) wA[kk] -= (kk <= sl) ? Sk[kk * NX] : 0.0;
            }
        }
        else if (isU)
        {   /* Ut rows: M[D+c][K] = -P_p[c] * A[K][c] = -Cp[K][c] */
#pragma unroll
            for (int K = 0; K < D; K++) wA[K] = -Cp[K * LDC + (sl - D)];
        }
        small = false;
        __syncwarp(); /* Cp is dead, Ms may be written */
#pragma unroll
        for (int j = 0; j < D; j++)
        {
            double piv = __shfl_sync(FULL, wA[j], j, 16), inv;
            tqb_pivot(piv, inv);
            small |= (piv <= v.o.regTol);
            wA[j] *= inv;
            const double yj = __shfl_sync(FULL, rr, j, 16) * inv;
            rr = fma(-wA[j], yj, rr);
            if (sl < 2 * D) Ms[j * LDR + sl] = wA[j];
            if (sl == j) { my_y = yj; my_d = piv; my_inv = inv; }
            __syncwarp();
            const double lj = wA[j];
            const double *col = Ms + j * LDR;
            if (j + 1 < D)
            {
                if ((j + 1) & 1) wA[j + 1] -= lj * col[j + 1];
#pragma unroll
                for (int k = (j + 2) & ~1; k < D; k += 2)
                {
                    const double2 l2 = *reinterpret_cast<const double2 *>(col + k);
                    wA[k] -= lj * l2.x;
                    wA[k + 1] -= lj * l2.y;
                }
            }
        }
        small = small && act;
        if (!(v.o.regType == 2 && attempt == 0 && __any_sync(FULL, small))) break;
        if (sl == 0 && small) atomicAdd(&v.counters[2], 1);
        __syncwarp();
    }

    /* ---- results to HBM ---- */
    if (ldW)
    {
        double *Lp = v.W + d1.w + (sl * (sl + 1)) / 2;
#pragma unroll
        for (int t = 0; t < D; t++)
            if (t < sl) Lp[t] = wA[t];
        Lp[sl] = my_d;
        v.Linv[xb + sl] = my_inv;
        v.dlam[xb + sl] = my_y;
    }
    if (isW) ys[sl] = my_y;
    __syncwarp();
    if (isU && act)
    {
        double2 *dst = reinterpret_cast<double2 *>(v.Ut + d2.x + (sl - D) * D);
#pragma unroll
        for (int t = 0; t < D / 2; t++) dst[t] = make_double2(wA[2 * t], wA[2 * t + 1]);
    }
    /* Schur complement for the parent's block: S = Utilde Utilde' (lower), s = Utilde y */
    double *Sg = v.schur + (size_t) g * NX * NX;
    constexpr int NE = (NX / 2) * (NX + 1);
    for (int e = sl; e < NE; e += 16)
    {
        const int p = e / (NX + 1), q = e - p * (NX + 1);
        const int a = q <= p ? p : NX - 1 - p;
        const int b = q <= p ? q : q - p - 1;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) s += Ms[j * LDR + D + a] * Ms[j * LDR + D + b];
        if (act) Sg[b * NX + a] = s;
    }
    if (sl < NX)
    {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) s += Ms[j * LDR + D + sl] * ys[j];
        if (act) v.rsub[xo + sl] = s;
    }
}

template <int NX, int NU>
__global__ void __launch_bounds__(128, 6) k_build_factor_chain2(View v, int top_begin, int count, int nlev, int bottom_kids_are_parents, int pre_ok)
{
    using Cfg = HalfCfg<NX, NU>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int w = blockIdx.x * wpb + wib;
    if (!pre_ok) { pdl_wait(); pdl_release(); }
    if (2 * w >= count) { pdl_wait(); pdl_release(); return; }
    const int chain = 2 * w + (lane >> 4);
    double *Ms = smem + (size_t) (2 * wib + (lane >> 4)) * Cfg::SMEM;
#pragma unroll 1
    for (int l = nlev - 1; l >= 0; l--)
    {
        const bool first = l == nlev - 1;
        bf_block_half<NX, NU>(v, (size_t) top_begin + (size_t) l * count + chain, chain < count, first ? bottom_kids_are_parents : 1, first, Ms, lane);
        __syncwarp();
    }
}

/* substitution along two chains per warp (see k_backsolve_chain) */
template <int NX>
__global__ void __launch_bounds__(128) k_backsolve_chain2(View v, int top_begin, int count, int nlev)
{
    constexpr int D = NX;
    const int lane = threadIdx.x & 31, sl = lane & 15;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (2 * w >= count) { pdl_wait(); pdl_release(); return; }
    const int chain = 2 * w + (lane >> 4);
    const bool valid = chain < count;
#pragma unroll 1
    for (int l = 0; l < nlev; l++)
    {
        int4 d0 = make_int4(0, 0, 0, 0), d1 = d0, d2 = d0;
        if (valid)
        {
            const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * ((size_t) top_begin + (size_t) l * count + chain);
            d0 = __ldg(dp); d1 = __ldg(dp + 1); d2 = __ldg(dp + 2);
        }
        const int xb = d1.x, xo = d0.z;
        const double *Lp = v.W + d1.w, *Up = v.Ut + d2.x;
        const bool ld = valid && sl < D;
        double lcol[D], ucol[NX];
#pragma unroll
        for (int i = 0; i < D; i++) lcol[i] = (ld && sl <= i) ? Lp[(i * (i + 1)) / 2 + sl] : 0.0;
#pragma unroll
        for (int c = 0; c < NX; c++) ucol[c] = ld ? Up[c * D + sl] : 0.0;
        double y = ld ? v.dlam[xb + sl] : 0.0;
        const double inv = ld ? v.Linv[xb + sl] : 0.0;
        if (l == 0) { pdl_wait(); pdl_release(); }
        const bool act = ld && v.status[d2.y] == TQB_RUNNING;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < NX; c++) s += ucol[c] * (ld ? v.dlam[xo + c] : 0.0);
        y -= s;
#pragma unroll
        for (int i = D - 1; i >= 0; i--)
        {
            const double yi = __shfl_sync(FULL, y, i, 16) * __shfl_sync(FULL, inv, i, 16);
            if (sl == i) y = yi;
            else if (sl < i) y -= lcol[i] * yi;
        }
        if (act) v.dlam[xb + sl] = y;
        __syncwarp();
    }
}

/* one block of the root-to-leaves substitution; `wait_here` as in bf_block */
template <int NX, int MD>
__device__ __forceinline__ void bs_block(const View &v, size_t desc, bool wait_here, int lane)
{
    constexpr int D = NX * MD;
    const int4 *dp = reinterpret_cast<const int4 *>(v.lvl_desc) + 3 * desc;
    const int4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
    const int xb = d1.x, xo = d0.z;
    const double *Lp = v.W + d1.w, *Up = v.Ut + d2.x;
    const bool act = lane < D;
    /* the factor of this block is several launches old: load it before waiting for the parents */
    double lcol[D], ucol[NX];
#pragma unroll
    for (int i = 0; i < D; i++) lcol[i] = (act && lane <= i) ? Lp[(i * (i + 1)) / 2 + lane] : 0.0;
#pragma unroll
    for (int c = 0; c < NX; c++) ucol[c] = act ? Up[c * D + lane] : 0.0;
    double y = act ? v.dlam[xb + lane] : 0.0;
    const double inv = act ? v.Linv[xb + lane] : 0.0;
    if (wait_here)
    {
        pdl_wait();
        pdl_release();
    }
    if (v.status[d2.y] != TQB_RUNNING) return;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NX; c++) s += ucol[c] * v.dlam[xo + c];
    y -= s;
#pragma unroll
    for (int i = D - 1; i >= 0; i--)
    {
        const double yi = __shfl_sync(FULL, y, i) * __shfl_sync(FULL, inv, i);
        if (lane == i) y = yi;
        else if (lane < i) y -= lcol[i] * yi;
    }
    if (act) v.dlam[xb + lane] = y;
}

template <int NX, int MD>
__global__ void __launch_bounds__(128) k_backsolve_fused(View v, int lvl_begin, int count)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= count) { pdl_wait(); pdl_release(); return; }
    bs_block<NX, MD>(v, (size_t) (lvl_begin + w), true, lane);
}

/* ---- the same substitution for the FAT levels: persistent warps fed by the bulk-copy engine ------------
 * A block of the root-to-leaves sweep streams 4.7 KB (packed L, Utilde, y, inverse diagonal) for ~1 us of
 * dependent arithmetic, and k_backsolve_fused has every warp wait for its own loads (ncu, round 1: 42 % of
 * the stall samples on the first use of a load, 3.7 TB/s at the bottom level).  Here a warp walks blocks
 * w, w + nwarps, ... and lane 0 hands the NEXT block's five contiguous pieces to the TMA unit
 * (cp.async.bulk global -> shared, completion on a warp-private mbarrier) before the warp starts on the
 * current one: two 5 KB buffers per warp, loads of block i+1 in flight under the arithmetic of block i.
 * Same arithmetic in the same order as bs_block (bit-identical). */
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int NX, int MD>
struct BsTmaCfg
{
    static constexpr int D = NX * MD;
    static constexpr int LB = ((D * (D + 1) / 2) * 8 + 15) / 16 * 16;  /* packed L, rounded up (the block owns D*D doubles) */
    static constexpr int UB = NX * D * 8, YB = D * 8, DB = ((NX * 8) + 15) / 16 * 16;
    static constexpr int OFF_U = LB, OFF_Y = LB + UB, OFF_I = OFF_Y + YB, OFF_D = OFF_I + YB;
    static constexpr int BUF = OFF_D + DB;                               /* bytes per buffer */
    static constexpr bool OK = (D % 2 == 0) && (NX % 2 == 0);           /* 16-byte alignment of every piece */
};

template <int NX, int MD>
__global__ void __launch_bounds__(128) k_backsolve_tma(View v, int lvl_begin, int count)
{
    using Cfg = BsTmaCfg<NX, MD>;
    constexpr int D = Cfg::D;
    extern __shared__ __align__(16) unsigned char bsm[];
    __shared__ unsigned long long bars[4][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * 4;
    int w = blockIdx.x * 4 + wib;
    unsigned char *buf0 = bsm + (size_t) wib * 2 * Cfg::BUF;
    if (lane == 0) { mbar_init(&bars[wib][0], 1); mbar_init(&bars[wib][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int4 *descs = reinterpret_cast<const int4 *>(v.lvl_desc);
    auto issue = [&](const int4 d0, const int4 d1, const int4 d2, int s) {
        /* lane 0: the five pieces of one block -> buffer s */
        unsigned char *b = buf0 + (size_t) s * Cfg::BUF;
        unsigned long long *bar = &bars[wib][s];
        mbar_expect_tx(bar, (unsigned) Cfg::BUF);
        bulk_g2s(b, v.W + d1.w, Cfg::LB, bar);
        bulk_g2s(b + Cfg::OFF_U, v.Ut + d2.x, Cfg::UB, bar);
        bulk_g2s(b + Cfg::OFF_Y, v.dlam + d1.x, Cfg::YB, bar);
        bulk_g2s(b + Cfg::OFF_I, v.Linv + d1.x, Cfg::YB, bar);
        bulk_g2s(b + Cfg::OFF_D, v.dlam + d0.z, Cfg::DB, bar);
    };
    /* everything this kernel reads that the previous level wrote is dlam of the parents: wait first */
    pdl_wait();
    pdl_release();
    if (w >= count) return;
    int4 c0 = __ldg(descs + 3 * (size_t) (lvl_begin + w)), c1 = __ldg(descs + 3 * (size_t) (lvl_begin + w) + 1),
         c2 = __ldg(descs + 3 * (size_t) (lvl_begin + w) + 2);
    int4 n0 = c0, n1 = c1, n2 = c2;
    if (w + nwarps < count)
    {
        const size_t e = 3 * (size_t) (lvl_begin + w + nwarps);
        n0 = __ldg(descs + e); n1 = __ldg(descs + e + 1); n2 = __ldg(descs + e + 2);
    }
    if (lane == 0) issue(c0, c1, c2, 0);
#pragma unroll 1
    for (int it = 0; w < count; w += nwarps, it++)
    {
        const int s = it & 1;
        const unsigned parity = (unsigned) (it >> 1) & 1u;
        const bool has_next = w + nwarps < count;
        if (has_next && lane == 0) issue(n0, n1, n2, s ^ 1);
        int4 m0 = n0, m1 = n1, m2 = n2; /* descriptor two blocks ahead: in flight under this block's arithmetic */
        if (w + 2 * nwarps < count)
        {
            const size_t e = 3 * (size_t) (lvl_begin + w + 2 * nwarps);
            m0 = __ldg(descs + e); m1 = __ldg(descs + e + 1); m2 = __ldg(descs + e + 2);
        }
        const bool running = v.status[c2.y] == TQB_RUNNING;
        mbar_wait(&bars[wib][s], parity);
        const double *Ls = reinterpret_cast<const double *>(buf0 + (size_t) s * Cfg::BUF);
        const double *Us = Ls + Cfg::OFF_U / 8, *ys = Ls + Cfg::OFF_Y / 8, *is = Ls + Cfg::OFF_I / 8, *dd = Ls + Cfg::OFF_D / 8;
        const bool act = lane < D;
        double y = act ? ys[lane] : 0.0;
        const double inv = act ? is[lane] : 0.0;
        double sacc = 0.0;
#pragma unroll
        for (int c = 0; c < NX; c++) sacc += (act ? Us[c * D + lane] : 0.0) * dd[c];
        y -= sacc;
#pragma unroll
        for (int i = D - 1; i >= 0; i--)
        {
            const double yi = __shfl_sync(FULL, y, i) * __shfl_sync(FULL, inv, i);
            if (lane == i) y = yi;
            else if (lane < i) y -= Ls[(i * (i + 1)) / 2 + lane] * yi;
        }
        if (act && running) v.dlam[c1.x + lane] = y;
        __syncwarp(); /* the buffer is free for the block after the next one */
        c0 = n0; c1 = n1; c2 = n2;
        n0 = m0; n1 = m1; n2 = m2;
    }
}

/* root-to-leaves substitution along chain w (see k_build_factor_chain): dlam of the block just
 * finished is the parent value of the next one */
template <int NX>
__global__ void __launch_bounds__(128) k_backsolve_chain(View v, int top_begin, int count, int nlev)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= count) { pdl_wait(); pdl_release(); return; }
#pragma unroll 1
    for (int l = 0; l < nlev; l++)
    {
        bs_block<NX, 1>(v, (size_t) top_begin + (size_t) l * count + w, l == 0, lane);
        __syncwarp();
    }
}
