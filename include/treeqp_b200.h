/*
 * treeqp_b200.h -- C ABI of the B200-native dual-Newton engine behind treeQP's C API.
 *
 * This header is the drop-in boundary.  A program written against the reference headers
 *   treeqp/src/tree_qp_common.h, treeqp/src/dual_Newton_tree.h, treeqp/utils/tree.h,
 *   treeqp/utils/types.h  (treeQP @5bffe146)
 * and BLASFEO's blasfeo_common.h (fa39ce5, LA_HIGH_PERFORMANCE, panel size 4) can be linked
 * against libtreeqp_b200.so instead of libtreeqp.a's dual_Newton_tree.o without recompiling:
 * every struct below has the same field order, types and size as its reference counterpart
 * (checked by tests/test_abi.py against the compiled reference), and every `treeqp_tdunes_*`
 * function has the reference's exact signature and return codes.
 *
 * The ABI corresponds to the reference's default build: PROFILE=0 (no `timings` member in the
 * workspace, dual_Newton_tree.h:146-148) and BLASFEO HIGH_PERFORMANCE (panel-major matrices); compiled with
 * -DPROFILE=n (1..3) this header is the PROFILE = n layout, served by libtreeqp_b200_profile<n>.so (`make -C treeqp_b200/csrc
 * profile PL=n`; section 4, timers).
 *
 * Sections
 *   1. BLASFEO-compatible matrix/vector views          (blasfeo_common.h:46-89)
 *   2. enums / return codes                            (treeqp/utils/types.h:37-85,
 *                                                       treeqp/src/dual_Newton_common.h:41-52)
 *   3. tree + QP containers                            (treeqp/utils/tree.h:41-51,
 *                                                       treeqp/src/tree_qp_common.h:43-133)
 *   4. TDUNES options / workspace / solve  -- THE HOT PATH (dual_Newton_tree.h:48-172)
 *   4b. SDUNES options / workspace / solve               (dual_Newton_scenarios.h:49-163)
 *   5. container helpers (the subset of tree_qp_common.h:139-323 a standalone user needs)
 *   6. additive B200 entry points (device-resident solves, batches of independent QPs)
 */
#ifndef TREEQP_B200_H_
#define TREEQP_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. BLASFEO-compatible views.  Define TREEQP_B200_EXTERNAL_BLASFEO before including this file
 *    if blasfeo_common.h is already in the translation unit.
 * ---------------------------------------------------------------------------------------- */
#ifndef TREEQP_B200_EXTERNAL_BLASFEO

#define TQB_PS 4 /* panel height of a panel-major double matrix (D_PS, blasfeo_block_size.h:37) */

/* panel-major matrix: element (i,j) lives at pA[(i - i%4)*cn + j*4 + i%4] */
struct blasfeo_dmat
{
    int m;       /* rows */
    int n;       /* columns */
    int pm;      /* rows rounded up to a multiple of 4 */
    int cn;      /* columns rounded up to a multiple of 4 */
    double *pA;  /* pm*cn doubles */
    double *dA;  /* scratch for an inverse diagonal (BLASFEO-internal) */
    int use_dA;  /* BLASFEO-internal validity flag of dA */
    int memsize; /* bytes owned by this view */
};

struct blasfeo_dvec
{
    int m;       /* length */
    int pm;      /* length rounded up to a multiple of 4 */
    double *pa;  /* pm doubles, element i at pa[i] */
    int memsize; /* bytes owned by this view */
};

#define BLASFEO_DMATEL(sA, ai, aj) \
    ((sA)->pA[((ai) - ((ai) & (TQB_PS - 1))) * (sA)->cn + (aj) * TQB_PS + ((ai) & (TQB_PS - 1))])
#define BLASFEO_DVECEL(sa, ai) ((sa)->pa[ai])

#endif /* TREEQP_B200_EXTERNAL_BLASFEO */

/* ------------------------------------------------------------------------------------------
 * 2. enums (values must equal the reference's, callers compare against them numerically)
 * ---------------------------------------------------------------------------------------- */
#ifndef TREEQP_UTILS_TYPES_H_ /* reference types.h not included */

#define TREEQP_INF 1e12

typedef enum { YES, NO } answer_t;

typedef enum
{
    TREEQP_SUMSQUAREDERRORS = 0,
    TREEQP_TWONORM,
    TREEQP_INFNORM,
} termination_t;

typedef enum
{
    TREEQP_OPTIMAL_SOLUTION_FOUND,      /* 0 */
    TREEQP_MAXIMUM_ITERATIONS_REACHED,  /* 1 */
    TREEQP_DN_NOT_DESCENT_DIRECTION,    /* 2 */
    TREEQP_DN_STAGE_QP_INIT_FAILED,     /* 3 */
    TREEQP_DN_STAGE_QP_SOLVE_FAILED,    /* 4 */
    TREEQP_IP_MIN_STEP,                 /* 5 */
    TREEQP_IP_UNKNOWN_FLAG,             /* 6 */
    TREEQP_OK,                          /* 7 */
    TREEQP_FAILURE,                     /* 8 */
    TREEQP_INVALID_OPTION,              /* 9 */
    TREEQP_ERROR_OPENING_FILE,          /* 10 */
    TREEQP_UNKNOWN_ERROR,               /* 11: also used for CUDA runtime failures */
} return_t;

typedef enum
{
    TREEQP_CLIPPING_SOLVER = 0,
    TREEQP_QPOASES_SOLVER,
} stage_qp_t;

#endif /* TREEQP_UTILS_TYPES_H_ */

#ifndef DUAL_NEWTON_COMMON_H_
typedef enum
{
    TREEQP_UNKNOWN_REGULARIZATION,
    TREEQP_NO_REGULARIZATION = 0, /* same value as UNKNOWN: quirk of the reference kept on purpose */
    TREEQP_ALWAYS_LEVENBERG_MARQUARDT,
    TREEQP_ON_THE_FLY_LEVENBERG_MARQUARDT,
} regType_t;
#endif

/* ------------------------------------------------------------------------------------------
 * 3. tree and QP containers
 * ---------------------------------------------------------------------------------------- */
#ifndef TREEQP_UTILS_TREE_H_
struct node
{
    int *kids;  /* indices of the children (contiguous: nodes are numbered breadth first) */
    int idx;    /* own index */
    int dad;    /* parent index, -1 at the root */
    int nkids;  /* number of children */
    int stage;  /* depth */
    int real;   /* realisation id (used by the LTI filler only) */
    int idxkid; /* position among the siblings */
};
#endif

#ifndef TREEQP_SRC_TREE_OCP_QP_COMMON_H_

typedef struct treeqp_info_t_
{
    int Nn;
    int iter;
    double total_time;
    double solver_time;
    double interface_time;
} treeqp_info_t;

/* scratch the container keeps so that x0 can be (re-)eliminated, tree_qp_common.h:58-80 */
typedef struct qp_internal_t_
{
    int *is_A_initialized;
    int *is_b_initialized;
    int is_C_initialized;
    int is_dmin_initialized;
    int is_dmax_initialized;
    int is_S_initialized;
    int is_r_initialized;
    struct blasfeo_dvec x0;
    struct blasfeo_dmat *A0;
    struct blasfeo_dvec *b0;
    struct blasfeo_dmat C0;
    struct blasfeo_dvec dmax0;
    struct blasfeo_dvec dmin0;
    struct blasfeo_dmat S0;
    struct blasfeo_dvec r0;
} qp_internal_t;

/* min sum_n 1/2 [x;u]'[Q S';S R][x;u] + q'x + r'u
 * s.t. x_n = A[n-1] x_dad(n) + B[n-1] u_dad(n) + b[n-1],  xmin<=x<=xmax, umin<=u<=umax,
 *      dmin <= C x + D u <= dmax                                    (tree_qp_common.h:85-116) */
typedef struct tree_qp_in_
{
    int N;   /* number of nodes */
    int *nx; /* states per node */
    int *nu; /* inputs per node */
    int *nc; /* general constraints per node */

    struct blasfeo_dmat *A; /* per edge (= node-1) */
    struct blasfeo_dmat *B;
    struct blasfeo_dvec *b;

    struct blasfeo_dmat *Q; /* per node */
    struct blasfeo_dmat *R;
    struct blasfeo_dmat *S;
    struct blasfeo_dvec *q;
    struct blasfeo_dvec *r;

    struct blasfeo_dvec *xmin;
    struct blasfeo_dvec *xmax;
    struct blasfeo_dvec *umin;
    struct blasfeo_dvec *umax;

    struct blasfeo_dmat *C;
    struct blasfeo_dmat *D;
    struct blasfeo_dvec *dmin;
    struct blasfeo_dvec *dmax;

    struct node *tree;

    qp_internal_t internal_memory;
} tree_qp_in;

typedef struct tree_qp_out_
{
    treeqp_info_t info;
    struct blasfeo_dvec *x;
    struct blasfeo_dvec *u;
    struct blasfeo_dvec *lam;  /* per edge */
    struct blasfeo_dvec *mu_x; /* + upper bound active, - lower bound active */
    struct blasfeo_dvec *mu_u;
    struct blasfeo_dvec *mu_d;
} tree_qp_out;

#endif /* TREEQP_SRC_TREE_OCP_QP_COMMON_H_ */

/* ------------------------------------------------------------------------------------------
 * 4. TDUNES: dual Newton strategy on the tree formulation  (THE HOT PATH)
 * ---------------------------------------------------------------------------------------- */
#ifndef TREEQP_SRC_DUAL_NEWTON_TREE_H_

/* The reference dispatches the stage QPs through a per-node table of 13 function pointers
 * (dual_Newton_tree.h:48-63).  On the GPU the stage QPs of all nodes run inside one batched
 * kernel, so the table is never filled; only its size matters for `calculate_size`. */
typedef struct stage_qp_fcn_ptrs_
{
    void *reserved[13];
} stage_qp_fcn_ptrs;

/* dual_Newton_tree.h:67-87 */
typedef struct treeqp_tdunes_opts_t_
{
    int maxIter;
    stage_qp_t *qp_solver; /* per node */
    int checkLastActiveSet; /* accepted; the GPU path always rebuilds (bit-identical result) */
    double stationarityTolerance;
    termination_t termCondition;
    regType_t regType;
    double regTol;
    double regValue;
    int lineSearchMaxIter;
    double lineSearchGamma;
    double lineSearchBeta;
    int lineSearchRestartTrigger;
} treeqp_tdunes_opts_t;

#if defined(PROFILE) && PROFILE > 0
/* utils/profiling.h:38-68, field for field (which members exist depends on the PROFILE level, as in the reference) */
typedef struct treeqp_profiling_t_
{
    int num_iter;
    int run_indx;
    double total_time;      /* PROFILE > 0: seconds of the last solve, minimum over the runs, trials of the first run */
    double min_total_time;
    int total_ls_iter;
#if PROFILE > 1
    double *iter_times;     /* per Newton iteration (NaN beyond the last one) */
    double *min_iter_times;
    int *ls_iters;
#endif
#if PROFILE > 2
    double *stage_qps_times;
    double *min_stage_qps_times;
    double *build_dual_times;
    double *min_build_dual_times;
    double *newton_direction_times;
    double *min_newton_direction_times;
    double *line_search_times;
    double *min_line_search_times;
#endif
} treeqp_profiling_t;
#endif

/* dual_Newton_tree.h:91-150.  Field order is the reference's; on this
 * implementation only the members marked [host] are valid host memory after `create`, the
 * iterates live in HBM and are reached through the context stored in stage_qp_data[0]. */
typedef struct treeqp_tdunes_workspace_
{
    int Nn; /* [host] */
    int Np; /* [host] */

    int lsIter;                   /* [host] line-search trials of the last Newton iteration */
    int lineSearchRestartCounter; /* [host] */

    int *npar;   /* [host] 1 x Nh nodes per depth */
    int *idxpos; /* [host] 1 x Nn offset of a node inside its parent's multiplier block */

    int *xasChanged;
    int *uasChanged;
    int *blockChanged;

    double *fval; /* [host] 1 x Nn (scratch) */
    double *cmod; /* [host] 1 x Nn (scratch) */

    stage_qp_fcn_ptrs *stage_qp_ptrs;
    void **stage_qp_data; /* [host] slot 0 holds the device context (see tdunes.c) */

    struct blasfeo_dvec *sqmod;
    struct blasfeo_dvec *srmod;

    struct blasfeo_dmat *sAB;
    struct blasfeo_dmat *sM;
    struct blasfeo_dmat *sW;
    struct blasfeo_dmat *sCholW;
    struct blasfeo_dmat *sUt;
    struct blasfeo_dmat *sCholUt;
    struct blasfeo_dvec *sres;
    struct blasfeo_dvec *sresMod;
    struct blasfeo_dvec *slambda;
    struct blasfeo_dvec *sDeltalambda;

    struct blasfeo_dvec *sx;
    struct blasfeo_dvec *su;
    struct blasfeo_dvec *sxas;
    struct blasfeo_dvec *suas;
    struct blasfeo_dvec *sxUnc;
    struct blasfeo_dvec *suUnc;

    struct blasfeo_dvec *sxasPrev;
    struct blasfeo_dvec *suasPrev;
    struct blasfeo_dmat *sWdiag;
#if defined(PROFILE) && PROFILE > 0
    treeqp_profiling_t timings; /* dual_Newton_tree.h:146-148: only in the PROFILE > 0 build (libtreeqp_b200_profile3.so) */
#endif
} treeqp_tdunes_workspace;

#endif /* TREEQP_SRC_DUAL_NEWTON_TREE_H_ */

/* replaces dual_Newton_tree.c:71 */
int treeqp_tdunes_opts_calculate_size(int Nn);
/* replaces dual_Newton_tree.c:79 */
void treeqp_tdunes_opts_create(int Nn, treeqp_tdunes_opts_t *opts, void *ptr);
/* replaces dual_Newton_tree.c:92 (same defaults) */
void treeqp_tdunes_opts_set_default(int Nn, treeqp_tdunes_opts_t *opts);
/* replaces dual_Newton_tree.c:1291: HOST bytes the caller must hand to `create` */
int treeqp_tdunes_calculate_size(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts);
/* replaces dual_Newton_tree.c:1411: carves the host buffer AND allocates the device arena.
 * Exits (like the reference, clipping.c:70-74) if the clipping solver is selected for a node
 * whose Q/R are not diagonal, S is not zero or nc>0. */
void treeqp_tdunes_create(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                          treeqp_tdunes_workspace *work, void *ptr);
/* replaces dual_Newton_tree.c:1654: lambda is the concatenation of the edge multipliers in node
 * order (node 1 .. Nn-1) */
void treeqp_tdunes_set_dual_initialization(const double *lambda, treeqp_tdunes_workspace *work);
/* replaces dual_Newton_tree.c:1104.  Reads qp_in from HOST memory, uploads it, runs the dual
 * Newton iterations on the GPU, writes qp_out on the host.  Returns the reference's codes;
 * a CUDA failure returns TREEQP_UNKNOWN_ERROR.  There is no CPU fallback. */
return_t treeqp_tdunes_solve(const tree_qp_in *qp_in, tree_qp_out *qp_out,
                             const treeqp_tdunes_opts_t *opts, treeqp_tdunes_workspace *work);
/* replaces dual_Newton_tree.c:1023 (public header dual_Newton_tree.h:172; called unconditionally by
 * examples/spring_mass_dual_newton_tree.c:142).  Dumps the iterates of the last solve -- read back from
 * HBM -- into the reference's files, same names and format ("%.16e" per line):
 * examples/spring_mass_utils/{x_opt,u_opt,deltalambda_opt,lambda_opt,iter}.txt (like the reference,
 * deltalambda_opt.txt receives lambda, :1056).  Np and tree are unused (as good as in the reference). */
void write_solution_to_txt(const tree_qp_in *qp_in, int Np, int iter, struct node *tree,
                           treeqp_tdunes_workspace *work);

/* PROFILE > 0 (Makefile.rule:112-119): the reference's workspace structs grow a trailing `timings` member and its solve fills
 * it (utils/profiling.h:38-68).  The default library is the PROFILE = 0 ABI: a caller compiled with PROFILE > 0 that links it
 * binds timers_print / timers_write_to_txt here and is stopped with a message rather than allowed to print timers nobody
 * filled.  `make -C treeqp_b200/csrc profile PL=n` builds libtreeqp_b200_profile<n>.so with -DPROFILE=n (n = 3: `make profile3`):
 * the same solvers, the PROFILE = n struct layout, and `timings` fed from CUDA events around every launch (which serialises the launches, as the
 * reference's own timers perturb its run): per Newton iteration the four phase times, the iteration time, the line-search
 * trial count; total_time and the minima over repeated runs as utils/profiling.c keeps them. */
#if defined(PROFILE) && PROFILE > 0
int timers_calculate_size(int num_iter);
void timers_create(int num_iter, treeqp_profiling_t *timings, void *ptr);
void timers_initialize(treeqp_profiling_t *timings);
void timers_update(treeqp_profiling_t *timings);
void timers_print(treeqp_profiling_t *timings);
void timers_write_to_txt(treeqp_profiling_t *timings);
#else
void timers_print(void *timings);
void timers_write_to_txt(void *timings);
#endif

/* ------------------------------------------------------------------------------------------
 * 4b. SDUNES: dual Newton strategy on the scenario decomposition (dual_Newton_scenarios.h:49-163)
 * ---------------------------------------------------------------------------------------- */
#ifndef TREEQP_SRC_DUAL_NEWTON_SCENARIOS_H_

/* dual_Newton_scenarios.h:49-66 */
typedef struct treeqp_sdunes_opts_t_
{
    int maxIter;
    int checkLastActiveSet; /* also selects the order of the banded factorisation (reverseCholesky) */
    double stationarityTolerance;
    termination_t termCondition;
    regType_t regType;
    double regTol;
    double regValue;
    int lineSearchMaxIter;
    double lineSearchGamma;
    double lineSearchBeta;
} treeqp_sdunes_opts_t;

/* dual_Newton_scenarios.h:70-139 (PROFILE=0).  Field order is the reference's; on this
 * implementation the members marked [host] are valid host memory after `create` / `solve` (they
 * are what the reference's callers read, examples/solve_qp_json.cpp:165-188, 470-471); the other
 * iterates live in HBM.  sTmpMats holds the device context. */
typedef struct treeqp_sdunes_workspace_
{
    int Ns; /* [host] scenarios */
    int Nh; /* [host] prediction horizon */
    int Nr; /* [host] robust horizon */
    int md; /* [host] realisations */

    int reverseCholesky; /* [host] */

    int **nodeIdx;       /* [host] Ns x (Nh+1) tree node of every scenario node */
    int **boundsRemoved; /* [host] Ns x (Nh+1) */

    int **xasChanged;
    int **uasChanged;

    int *commonNodes; /* [host] Ns-1 */
    double *fvals;    /* [host] Ns (scratch) */

    struct blasfeo_dmat *sJayD;
    struct blasfeo_dmat *sJayL;
    struct blasfeo_dmat *sCholJayD;
    struct blasfeo_dmat *sCholJayL;
    struct blasfeo_dmat *sUt;
    struct blasfeo_dmat *sK;

    struct blasfeo_dmat *sTmpMats; /* device context (opaque) */
    struct blasfeo_dvec *sTmpVecs;

    struct blasfeo_dvec *sResNonAnticip;
    struct blasfeo_dvec *sRhsNonAnticip;
    struct blasfeo_dvec *slambda; /* [host] Ns-1 multipliers of the non-anticipativity constraints */
    struct blasfeo_dvec *sDeltalambda;

    struct blasfeo_dmat **sZbar;
    struct blasfeo_dmat **sLambdaD;
    struct blasfeo_dmat **sLambdaL;
    struct blasfeo_dmat **sCholLambdaD;
    struct blasfeo_dmat **sCholLambdaL;

    struct blasfeo_dvec *sQ;
    struct blasfeo_dvec *sR;
    struct blasfeo_dvec *sq;
    struct blasfeo_dvec *sr;
    struct blasfeo_dvec *sQinv;
    struct blasfeo_dvec *sRinv;

    struct blasfeo_dvec **sQinvCal;
    struct blasfeo_dvec **sRinvCal;

    struct blasfeo_dvec **sx; /* [host] Ns x Nh scenario states after the last solve */
    struct blasfeo_dvec **su; /* [host] Ns x Nh scenario inputs */
    struct blasfeo_dvec **sxas;
    struct blasfeo_dvec **suas;
    struct blasfeo_dvec **sxUnc;
    struct blasfeo_dvec **suUnc;

    struct blasfeo_dvec **sresk;
    struct blasfeo_dvec **sreskMod;
    struct blasfeo_dvec **smu; /* [host] Ns x Nh multipliers of the scenario dynamics */
    struct blasfeo_dvec **sDeltamu;

    struct blasfeo_dvec **sxasPrev;
    struct blasfeo_dvec **suasPrev;
    struct blasfeo_dmat **sTmpLambdaD;
#if defined(PROFILE) && PROFILE > 0
    treeqp_profiling_t timings; /* dual_Newton_scenarios.h:135-137 */
#endif
} treeqp_sdunes_workspace;

#endif /* TREEQP_SRC_DUAL_NEWTON_SCENARIOS_H_ */

/* replaces dual_Newton_scenarios.c:66 */
int treeqp_sdunes_opts_calculate_size(int Nn);
/* replaces dual_Newton_scenarios.c:73 */
void treeqp_sdunes_opts_create(int Nn, treeqp_sdunes_opts_t *opts, void *ptr);
/* replaces dual_Newton_scenarios.c:80 (same defaults) */
void treeqp_sdunes_opts_set_default(int Nn, treeqp_sdunes_opts_t *opts);
/* replaces dual_Newton_scenarios.c:1495: HOST bytes the caller must hand to `create` */
int treeqp_sdunes_calculate_size(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts);
/* replaces dual_Newton_scenarios.c:1607: carves the host buffer AND allocates the device arena.
 * Exits if the QP violates the preconditions of the scenario decomposition (x0 eliminated, uniform
 * nx/nu, diagonal weights, bounds only, scenarios.c:1498-1504). */
void treeqp_sdunes_create(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work,
                          void *ptr);
/* replaces dual_Newton_scenarios.c:99 */
int treeqp_sdunes_calculate_dual_dimension(int Nr, int md, int nu);
/* replaces dual_Newton_scenarios.c:2094: lam over the scenario pairs, mu over (scenario, stage) */
void treeqp_sdunes_set_dual_initialization(double *lam, double *mu, treeqp_sdunes_workspace *work);
/* replaces dual_Newton_scenarios.c:1850.  Host buffers in, host buffers out; the whole Newton loop
 * is one persistent kernel.  Returns 0 / 1 like the reference, TREEQP_UNKNOWN_ERROR on a CUDA
 * failure.  No CPU fallback. */
return_t treeqp_sdunes_solve(tree_qp_in *qp_in, tree_qp_out *qp_out, treeqp_sdunes_opts_t *opts,
                             treeqp_sdunes_workspace *work);
/* replaces dual_Newton_scenarios.c:204 (public header dual_Newton_scenarios.h:163; called unconditionally
 * by examples/spring_mass_dual_newton_scenarios.c:149): scenario iterates and duals of the last solve to
 * examples/spring_mass_utils/{lambdaIter,muIter,xIter,uIter,iter}.txt */
void write_scenarios_solution_to_txt(int Ns, int Nh, int Nr, int md, int nx, int nu, int NewtonIter,
                                     treeqp_sdunes_workspace *work);

/* ------------------------------------------------------------------------------------------
 * 5. container helpers (same names and semantics as tree_qp_common.h / utils/tree.h)
 * ---------------------------------------------------------------------------------------- */
int calculate_number_of_nodes(int md, int Nr, int Nh);                /* utils/tree.c:36  */
void setup_multistage_tree(int md, int Nr, int Nh, int *nk);          /* utils/tree.c:247 */
int number_of_nodes_from_nkids(const int *nkids);                     /* utils/tree.c:105 */

int tree_qp_in_calculate_size(int Nn, const int *nx, const int *nu, const int *nc, const int *nk);
void tree_qp_in_create(int Nn, const int *nx, const int *nu, const int *nc, const int *nk,
                       tree_qp_in *qp_in, void *ptr);
int tree_qp_out_calculate_size(const int Nn, const int *const nx, const int *const nu,
                               const int *const nc);
void tree_qp_out_create(const int Nn, const int *const nx, const int *const nu,
                        const int *const nc, tree_qp_out *const qp_out, void *ptr);

void tree_qp_in_set_edge_dynamics_colmajor(const double *const A, const double *const B,
                                           const double *const b, tree_qp_in *const qp_in,
                                           const int indx);
void tree_qp_in_set_node_objective_colmajor(const double *const Q, const double *const R,
                                            const double *const S, const double *const q,
                                            const double *const r, tree_qp_in *const qp_in,
                                            const int indx);
void tree_qp_in_set_node_objective_diag(const double *const Qd, const double *const Rd,
                                        const double *const q, const double *const r,
                                        tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_node_bounds(const double *const xmin, const double *const xmax,
                                const double *const umin, const double *const umax,
                                tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_inf_bounds(tree_qp_in *qp_in);
/* general stage constraints dmin <= C x + D u <= dmax (tree_qp_common.h:249-265; lda < 0 = tight).  TDUNES
 * takes them on nodes that use the dense stage solver (TREEQP_QPOASES_SOLVER), nx + nu + nc <= 32, nc <= 8 */
void tree_qp_in_set_node_C_colmajor(const double *const C, const int lda, tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_node_D_colmajor(const double *const D, const int lda, tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_node_dmin(const double *const dmin, tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_node_dmax(const double *const dmax, tree_qp_in *const qp_in, const int indx);
void tree_qp_in_set_node_general_constraints(const double *const C, const double *const D, const double *const dmin,
                                             const double *const dmax, tree_qp_in *const qp_in, const int indx);
void tree_qp_out_get_node_mu_d(double *const mu, const tree_qp_out *const qp_out, const int indx);
/* initial state: pinned through the root's bounds, or eliminated (nx[0] becomes 0; only b of the
 * root's children, r[0], dmin/dmax[0] depend on x0)          tree_qp_common.h:167-169, 314-316 */
void tree_qp_in_eliminate_x0(tree_qp_in *const qp_in);
void tree_qp_out_eliminate_x0(tree_qp_out *const qp_out);
void tree_qp_in_set_x0_strvec(tree_qp_in *qp_in, struct blasfeo_dvec *sx0);
void tree_qp_in_set_x0_colmaj(tree_qp_in *qp_in, double *x0);
/* concatenated per-edge / per-node data, tree_qp_common.c:1959-2100 */
void tree_qp_in_set_ltv_dynamics_colmajor(double *A, double *B, double *b, tree_qp_in *qp_in);
void tree_qp_in_set_ltv_objective_diag(double *Qd, double *Rd, double *q, double *r,
                                       tree_qp_in *qp_in);
void tree_qp_in_set_ltv_bounds(double *xmin, double *xmax, double *umin, double *umax,
                               tree_qp_in *qp_in);

void tree_qp_out_get_node_x(double *const x, const tree_qp_out *const qp_out, const int indx);
void tree_qp_out_get_node_u(double *const u, const tree_qp_out *const qp_out, const int indx);
void tree_qp_out_get_edge_lam(double *const lam, const tree_qp_out *const qp_out, const int indx);
void tree_qp_out_get_node_mu_x(double *const mu_x, const tree_qp_out *const qp_out, const int indx);
void tree_qp_out_get_node_mu_u(double *const mu_u, const tree_qp_out *const qp_out, const int indx);

/* infinity norm of the KKT residuals (stationarity, dynamics, bound feasibility,
 * complementarity), tree_qp_common.c:540-788; bounds/diagonal or dense objective, nc == 0 */
double tree_qp_out_max_KKT_res(const tree_qp_in *const qp_in, const tree_qp_out *const qp_out);

/* ------------------------------------------------------------------------------------------
 * 6. additive B200 entry points (not in the reference; the calls above keep their signatures)
 * ---------------------------------------------------------------------------------------- */

/* frees the device arena allocated by treeqp_tdunes_create (the reference never frees
 * because its workspace lives entirely in the caller's buffer) */
void treeqp_tdunes_destroy(treeqp_tdunes_workspace *work);

/* per-solve statistics of the last solve on this workspace */
typedef struct treeqp_b200_stats_
{
    int iter;              /* Newton iterations */
    int ls_total;          /* line-search trials over all iterations */
    int ls_iters[256];     /* trials per Newton iteration (first 256) */
    int n_reg;             /* number of blocks that received on-the-fly regularisation */
    double gpu_ms;         /* device time of the Newton loop (CUDA events) */
    double h2d_ms, d2h_ms; /* device time of the copies */
    double pack_ms;        /* host time spent packing/unpacking BLASFEO views */
    long long h2d_bytes, d2h_bytes;
    long long kernel_launches; /* kernels launched by the last solve */
    /* per kernel class (see treeqp_b200_kernel_class_name), filled when profiling is on:
     * device time from CUDA events recorded on the solver's stream around every launch */
    double kernel_ms[16];
    int kernel_count[16];
} treeqp_b200_stats;
void treeqp_tdunes_get_stats(const treeqp_tdunes_workspace *work, treeqp_b200_stats *stats);

/* Device-resident protocol, for callers that re-solve without touching the data
 * (and for benchmarking the kernels with inputs already in HBM):
 *   upload:  pack qp_in and copy it to the device arena (what `solve` does first)
 *   solve_resident: Newton loop only; the dual start is the one last set by
 *                   treeqp_tdunes_set_dual_initialization (kept on the device, re-applied
 *                   at the start of every call)
 *   download: copy x,u,lam,mu into qp_out (what `solve` does last) */
return_t treeqp_tdunes_upload(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                              treeqp_tdunes_workspace *work);
return_t treeqp_tdunes_solve_resident(const treeqp_tdunes_opts_t *opts,
                                      treeqp_tdunes_workspace *work);
return_t treeqp_tdunes_download(tree_qp_out *qp_out, treeqp_tdunes_workspace *work);
/* MPC re-solve fast path (benchmark/treeqp_main.m:235-237: between consecutive QPs of a closed loop
 * only x0 changes).  After tree_qp_in_set_x0_colmaj/strvec on an uploaded QP, update_x0 re-sends only
 * what depends on x0 -- the root and its children, a few hundred bytes -- instead of the whole
 * problem; warm_start_resident makes the multipliers of the last solve, still in HBM, the dual
 * start of the next one (what treeqp_tdunes_set_dual_initialization(lambda_prev) would do through
 * the host). */
return_t treeqp_tdunes_update_x0(const tree_qp_in *qp_in, treeqp_tdunes_workspace *work);
return_t treeqp_tdunes_warm_start_resident(treeqp_tdunes_workspace *work);
/* tree_qp_out_max_KKT_res (tree_qp_common.c:769) of the last solve, evaluated on the device from the
 * resident iterates: no O(Nn) host pass and no download needed for a convergence certificate.
 * The batch form fills kkt[n], one value per QP. */
return_t treeqp_tdunes_kkt_resident(treeqp_tdunes_workspace *work, double *kkt);

/* Batch of n independent tree QPs solved as one forest: every kernel launch covers the same
 * tree level of all QPs.  All QPs share `opts` (qp_solver is read from opts for the first Nn
 * nodes of each QP).  status[i] receives the per-QP return code; the return value is
 * TREEQP_OK if the batch ran, TREEQP_UNKNOWN_ERROR on a CUDA failure. */
typedef struct treeqp_tdunes_batch_ treeqp_tdunes_batch;
treeqp_tdunes_batch *treeqp_tdunes_batch_create(int n, const tree_qp_in *const *qp_in,
                                                const treeqp_tdunes_opts_t *opts);
void treeqp_tdunes_batch_set_dual_initialization(treeqp_tdunes_batch *batch, int i,
                                                 const double *lambda);
return_t treeqp_tdunes_batch_upload(treeqp_tdunes_batch *batch, const tree_qp_in *const *qp_in);
/* closed-loop step of the whole batch: after tree_qp_in_set_x0_colmaj / _strvec on every qp_in, re-send only what
 * depends on x0 (root + its children of every tree; one strided copy per array for a batch of equal shapes) */
return_t treeqp_tdunes_batch_update_x0(treeqp_tdunes_batch *batch, const tree_qp_in *const *qp_in);
return_t treeqp_tdunes_batch_solve_resident(treeqp_tdunes_batch *batch, return_t *status);
return_t treeqp_tdunes_batch_download(treeqp_tdunes_batch *batch, tree_qp_out *const *qp_out);
/* upload + solve_resident + download */
return_t treeqp_tdunes_solve_batch(treeqp_tdunes_batch *batch, const tree_qp_in *const *qp_in,
                                   tree_qp_out *const *qp_out, return_t *status);
return_t treeqp_tdunes_batch_kkt_resident(treeqp_tdunes_batch *batch, double *kkt);
void treeqp_tdunes_batch_get_stats(const treeqp_tdunes_batch *batch, treeqp_b200_stats *stats);
void treeqp_tdunes_batch_destroy(treeqp_tdunes_batch *batch);

/* ONE tree QP sharded by subtree over the GPUs of a box (one process per GPU).  The tree is cut
 * at the first depth with as many nodes as ranks; every rank factors its subtrees locally, the cut
 * nodes' Schur complements are all-gathered once per Newton iteration, the few levels above the
 * cut are factored redundantly, and the line-search scalars are all-reduced (NCCL over NVLink).
 * All ranks pass the SAME qp_in; each rank uploads only its share and writes into qp_out only the
 * nodes it owns (its subtrees; rank 0 also the levels above the cut).
 *   comm:  rank 0 calls treeqp_b200_comm_unique_id and hands the 128 bytes to the other ranks
 *          (MPI, torch.distributed, a file ...); every rank then calls treeqp_b200_comm_create. */
typedef struct treeqp_b200_comm_ treeqp_b200_comm;
int treeqp_b200_comm_unique_id(char *id128);
treeqp_b200_comm *treeqp_b200_comm_create(int rank, int nranks, const char *id128, int device);
void treeqp_b200_comm_destroy(treeqp_b200_comm *comm);

typedef struct treeqp_tdunes_sharded_ treeqp_tdunes_sharded;
treeqp_tdunes_sharded *treeqp_tdunes_sharded_create(const tree_qp_in *qp_in, const treeqp_tdunes_opts_t *opts,
                                                    treeqp_b200_comm *comm);
/* lambda: the full vector of treeqp_tdunes_set_dual_initialization */
void treeqp_tdunes_sharded_set_dual_initialization(treeqp_tdunes_sharded *sh, const double *lambda,
                                                   const tree_qp_in *qp_in);
return_t treeqp_tdunes_sharded_upload(treeqp_tdunes_sharded *sh, const tree_qp_in *qp_in);
return_t treeqp_tdunes_sharded_solve_resident(treeqp_tdunes_sharded *sh);
return_t treeqp_tdunes_sharded_download(treeqp_tdunes_sharded *sh, tree_qp_out *qp_out);
return_t treeqp_tdunes_sharded_solve(treeqp_tdunes_sharded *sh, const tree_qp_in *qp_in, tree_qp_out *qp_out);
void treeqp_tdunes_sharded_get_stats(const treeqp_tdunes_sharded *sh, treeqp_b200_stats *stats);
void treeqp_tdunes_sharded_get_plan(const treeqp_tdunes_sharded *sh, int *cut_depth, int *ncut, int *n_local,
                                    int *n_owned);
void treeqp_tdunes_sharded_destroy(treeqp_tdunes_sharded *sh);
/* the sharding plan alone (host logic, no GPU): owned_global / ghost_global are [N] 0/1 masks */
int treeqp_b200_shard_plan(const tree_qp_in *qp_in, int rank, int nranks, int *cut_depth, int *ncut, int *n_local,
                           int *owned_global, int *ghost_global);

/* SDUNES: same additive protocol as TDUNES (destroy, statistics, device-resident solves) */
void treeqp_sdunes_destroy(treeqp_sdunes_workspace *work);
void treeqp_sdunes_get_stats(const treeqp_sdunes_workspace *work, treeqp_b200_stats *stats);
return_t treeqp_sdunes_upload(tree_qp_in *qp_in, treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work);
return_t treeqp_sdunes_solve_resident(treeqp_sdunes_opts_t *opts, treeqp_sdunes_workspace *work);
return_t treeqp_sdunes_download(tree_qp_out *qp_out, treeqp_sdunes_workspace *work);
/* the SDUNES solve is ONE persistent kernel; treeqp_b200_stats.kernel_ms[k] then holds the time
 * spent in phase k of that kernel (device globaltimer), named by treeqp_sdunes_phase_name */
int treeqp_sdunes_num_phases(void);
const char *treeqp_sdunes_phase_name(int k);
/* scenario duals after the last solve: lam [dual dimension], mu [Ns*Nh*nx] */
void treeqp_sdunes_get_duals(double *lam, double *mu, const treeqp_sdunes_workspace *work);

/* per-launch CUDA-event timing of every kernel class (adds two event records per launch);
 * off by default, also switched on by the environment variable TREEQP_B200_PROFILE=1 */
void treeqp_b200_set_profiling(int on);
int treeqp_b200_num_kernel_classes(void);
const char *treeqp_b200_kernel_class_name(int k);

/* library / device introspection */
const char *treeqp_b200_version(void);
/* solver contexts (device arena + pinned staging) alive in this process.  The reference API has create but no
 * destroy: a create over caller memory that already carries a workspace releases that workspace's context first,
 * treeqp_*_destroy releases it explicitly. */
int treeqp_b200_live_contexts(void);
/* Page-lock (cudaHostRegister) a buffer the caller owns -- typically the `ptr` given to tree_qp_in_create and to
 * tree_qp_out_create -- so that treeqp_tdunes_solve moves large trees by DMA straight out of / into it: the [A | B]
 * blocks and the vectors of tree_qp_in travel as strided copies in the caller's layout and are converted on the
 * device, the results are assembled on the device in tree_qp_out's layout and written in place (config 3: upload
 * 12.0 -> 7.0 ms, download 2.9 -> 0.9 ms).  One call per buffer before the solve loop (tens of milliseconds for a few
 * hundred MB); treeqp_b200_host_unregister(ptr) BEFORE the buffer is freed.  Only the pages that lie entirely inside
 * [ptr, ptr + bytes) are locked (a page shared with another allocation has to stay pageable), so allocate the buffer
 * page-aligned with a size rounded up to the page size (posix_memalign(&ptr, 4096, ...)).  Without the call -- or for a
 * tree whose views are not wholly inside registered pages -- every solve goes through the library's own pinned
 * staging block (host pass over every node): same results, slower copies.  Return 0 on success. */
int treeqp_b200_host_register(void *ptr, size_t bytes);
int treeqp_b200_host_unregister(void *ptr); /* NULL: every registered range */
int treeqp_b200_device_count(void);
int treeqp_b200_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* TREEQP_B200_H_ */
