/*
 * host_internal.h -- internal interface between the C host layer (containers.c, forest.c,
 * tdunes.c) and the CUDA layer (cuda/device.cu, cuda/kernels.cu).  Plain C, no CUDA types.
 *
 * Vocabulary
 *   forest   nT independent tree QPs laid out back to back ("tree-major"); a single
 *            treeqp_tdunes_solve is a forest of one tree, a batch of MPC QPs is a forest of n.
 *   node     global index g = tnode0[t] + local breadth-first index
 *   x-space  all per-state vectors (x, q, xmin, lambda, res, ...) concatenated in node order;
 *            the multiplier lambda_k of node k's dynamics sits at xoff[k] (unused for roots), so
 *            the reference's per-parent block vector slambda[p] (dual_Newton_tree.c:177-194) is the
 *            contiguous range [xoff[kid0[p]], +dim[p]).
 *   u-space  same for inputs.
 *   block    a parent node p: dual Hessian block W_p (dim[p] x dim[p], dim = sum of the kids' nx)
 *            and, if p is not a root, the coupling Ut_p (nx[p] x dim[p]).
 */
#ifndef TQB_HOST_INTERNAL_H_
#define TQB_HOST_INTERNAL_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- integer tables (one allocation; same layout on host and device) ---- */
typedef struct tqb_tables
{
    int *nx, *nu;        /* [Nn] */
    int *dad;            /* [Nn] global index of the parent, -1 for roots */
    int *kid0, *nkids;   /* [Nn] first child / number of children */
    int *tree;           /* [Nn] owning tree */
    int *depth;          /* [Nn] */
    int *pos;            /* [Nn] offset of lambda_k inside its parent's block (idxpos) */
    int *dim;            /* [Nn] block dimension (0 for leaves) */
    int *xoff, *uoff;    /* [Nn] offsets into x-/u-space */
    int *aoff, *boff;    /* [Nn] offsets of A_k (nx[k] x nx[dad]) / B_k (nx[k] x nu[dad]), column-major */
    int *woff, *utoff;   /* [Nn] offsets of W_p / Ut_p */
    int *hoff;           /* [Nn] dense stage QP: offset of H_k / P_k ((nx+nu)^2 doubles), -1 = clipping node */
    int *ncn, *doff, *cgoff; /* [Nn] general stage constraints: rows, offset into d-space (dmin, dmax, mu_d),
                              * offset of Cg_k = [C_k D_k] (nc x (nx+nu) column-major) */
    int *lvl_nodes;      /* [Np] parents sorted by depth (level lists) */
    int *tnode0;         /* [nT+1] node range of each tree */
    int *chunk_beg, *chunk_end, *chunk_tree; /* [nchunk] fixed-shape reduction chunks (node ranges) */
    int *chunk_x0, *chunk_x1, *chunk_u0, *chunk_u1; /* [nchunk] the same chunks as x-/u-space ranges */
    int *tchunk0;        /* [nT+1] chunk range of each tree */
} tqb_tables;

typedef struct tqb_dims
{
    int nT, Nn, Np, Nh;  /* trees, nodes, parents, max depth */
    int X, U;            /* x-/u-space sizes */
    long long szA, szB, szW, szUt;
    long long szH;       /* doubles in the dense stage Hessians (0: every node uses clipping) */
    int has_dense, nvmax; /* any dense stage QP; largest nx+nu among them */
    int NC, ncmax;       /* general stage constraints: rows in total (d-space), most rows of one node */
    long long szCg;      /* doubles in the [C D] blocks */
    int dmax, nxmax, numax, nkmax, abmax; /* abmax = max over parents of dim*(nx+nu) */
    int nchunk;
    int n_int;           /* ints in the table block */
    long long n_in;      /* doubles in the input block */
} tqb_dims;

/* ---- input block (doubles; one allocation, one H2D copy) ---- */
typedef struct tqb_inputs
{
    double *A, *B;       /* [szA], [szB] */
    double *b;           /* [X] b_k at xoff[k] (unused for roots) */
    double *q, *Qd, *Qinv, *xmin, *xmax; /* [X] */
    double *r, *Rd, *Rinv, *umin, *umax; /* [U] */
    double *lam0;        /* [X] dual start, lambda_k at xoff[k] */
    /* dense stage QPs (only allocated when dims.has_dense): H_k = [Q S'; S R] column-major at
     * hoff[k], and the bounds of the dense nodes (their slots in xmin.. hold +-1e300 so that the
     * vector-only line-search kernel sees an inactive, weightless variable) */
    double *H;           /* [szH] */
    double *dxmin, *dxmax, *dumin, *dumax; /* [X] [X] [U] [U] */
    /* general stage constraints dmin <= C x + D u <= dmax of the dense nodes (only when dims.NC > 0) */
    double *Cg, *dmin, *dmax; /* [szCg] [NC] [NC] */
} tqb_inputs;

#define TQB_NC_MAX 8 /* general constraint rows per node the warp active-set solver handles */

/* ---- options forwarded to the device ---- */
typedef struct tqb_opts
{
    int maxIter, termCondition, regType, lsMaxIter, lsRestartTrigger;
    double tol, regTol, regValue, gamma, beta;
} tqb_opts;

/* opaque device context (cuda/device.cu) */
typedef struct tqb_device tqb_device;

/* host-side forest: tables + pinned staging for the inputs and the outputs */
/* Runs of nodes whose views sit at a constant stride in the caller's tree_qp_in / tree_qp_out (tree_qp_in_create carves
 * [A B Q R S C D] node by node and then [b q r xmin xmax umin umax dmin dmax] node by node; tree_qp_out_create carves
 * [x u mu_x mu_u mu_d lam] node by node): once the caller's buffer is page-locked such a run travels as strided DMA
 * copies, in the caller's layout, and is converted on the device (k_unpack_run / k_gather_run). */
typedef struct tqb_in_run
{
    int k0, k1;                       /* LOCAL nodes [k0, k1) of the forest's tree, k0 >= 1 (sharded: node gmap[k0] .. of the caller's tree) */
    const char *mbase;                /* A of node k0 */
    long long mpitch, mwidth, b_off;  /* bytes: node to node, A..end of B, A to B */
    int ma, na, cna, mb, nb, cnb;     /* BLASFEO panel-major views of A and B */
    const char *vbase;                /* b of node k0 */
    long long vpitch, vwidth;         /* bytes */
    int voff[7];                      /* doubles from vbase: b q r xmin xmax umin umax */
    int nx, nu;
} tqb_in_run;
typedef struct tqb_out_run
{
    int k0, k1;
    char *base;                       /* x of node k0 */
    long long pitch, width;           /* bytes */
    int off[5];                       /* doubles from base: x u mu_x mu_u lam */
    int nx, nu;
} tqb_out_run;
#define TQB_MAX_RUNS 8

typedef struct tqb_forest
{
    tqb_dims dims;
    int *lvl_ptr;            /* [Nh+1] host only: ranges of lvl_nodes per depth */
    int *h_int;              /* host table block (n_int ints) */
    tqb_tables ht;           /* pointers into h_int */
    double *h_in;            /* pinned input staging (n_in doubles) */
    tqb_inputs hin;          /* pointers into h_in */
    double *h_out;           /* pinned output staging: x[X] u[U] lam[X] mu_x[X] mu_u[U] mu_d[NC] */
    int *h_status, *h_iter;  /* pinned [nT] */
    tqb_device *dev;
    /* statistics of the last solve */
    int ls_iters[256];
    int iter_max, ls_total, n_reg;
    double gpu_ms, h2d_ms, d2h_ms, pack_ms;
    long long h2d_bytes, d2h_bytes, launches;
    double kernel_ms[16];
    int kernel_count[16];
    double comm_ms;          /* sharded solves: device time spent in collectives */
    int n_collectives;
    /* strided DMA of the caller's own buffers (large single trees): runs of the last upload / download and what they
     * were derived from */
    int n_in_runs, n_out_runs;
    tqb_in_run in_runs[TQB_MAX_RUNS];
    tqb_out_run out_runs[TQB_MAX_RUNS];
    const void *in_key[4], *out_key[4];
    int in_direct;           /* the last upload used in_runs */
    int lam0_zero;           /* the host dual start is all +0.0 (set_dual_initialization of a large single tree) */
} tqb_forest;

/* ---- registry.c: contexts by the caller's workspace memory (a re-create on the same ptr frees the old arena) ---- */
void tqb_registry_replace(const void *key, void *ctx, void (*dtor)(void *));
void tqb_registry_remove(const void *ctx);
int tqb_registry_size(void);

/* ---- forest.c ---- */
struct tree_qp_in_;
struct tree_qp_out_;
/* `solver`: stage QP solver per node (stage_qp_t values, the same for every tree of the forest), NULL = clipping */
tqb_forest *tqb_forest_create(int nT, const struct tree_qp_in_ *const *qps, const int *solver);
/* general form: `ghost` marks nodes to be treated as parents whose block lives elsewhere, `owned`
 * restricts the reduction chunks (both per node, NULL = none / all) */
tqb_forest *tqb_forest_create_ex(int nT, const struct tree_qp_in_ *const *qps, const int *ghost, const int *owned,
                                 const int *solver);
void tqb_forest_destroy(tqb_forest *f);
/* BLASFEO views -> pinned staging (reciprocals of the weights are formed here,
 * dual_Newton_tree_clipping.c:149-170) */
void tqb_forest_pack(tqb_forest *f, int t, const struct tree_qp_in_ *qp);
/* pack + H2D as a pipeline over node ranges (single tree) or tree ranges (batch); tail_ms = the
 * wait after the last range was packed.  Returns 0 on success. */
int tqb_forest_pack_upload(tqb_forest *f, const struct tree_qp_in_ *const *qp, double *tail_ms);
/* re-pack and re-send nodes [j0, j1) of tree t only (what an x0 update touches) / the dual start only */
int tqb_forest_pack_upload_mapped(tqb_forest *f, const struct tree_qp_in_ *full_qp, const int *gmap, double *tail_ms);
/* results of a large single tree straight into the caller's (page-locked) tree_qp_out; 0 = not applicable */
int tqb_forest_download_direct(tqb_forest *f, struct tree_qp_out_ *out, const int *gmap, const int *owned);
int tqb_forest_update_nodes(tqb_forest *f, int t, const struct tree_qp_in_ *qp, int j0, int j1);
int tqb_forest_update_x0_batch(tqb_forest *f, const struct tree_qp_in_ *const *qp, long long *bytes);
int tqb_forest_upload_lambda0(tqb_forest *f);
int tqb_forest_warm_start(tqb_forest *f);
void tqb_forest_set_lambda0(tqb_forest *f, int t, const double *lambda);
void tqb_forest_unpack(const tqb_forest *f, int t, struct tree_qp_out_ *out);
void tqb_layout_tables(const tqb_dims *d, int *base, tqb_tables *t);
void tqb_layout_inputs(const tqb_dims *d, double *base, tqb_inputs *in);
/* returns YES(0)/NO(1) like stage_qp_clipping_is_applicable (clipping.c:45-77) */
int tqb_clipping_applicable(const struct tree_qp_in_ *qp, int node);
double tqb_wall_ms(void);

/* ---- cuda/device.cu ---- */
int tqb_cuda_device_count(void);
int tqb_cuda_set_device(int dev);
void *tqb_pinned_alloc(size_t bytes);
void tqb_pinned_free(void *p);
tqb_device *tqb_device_create(const tqb_dims *dims, const int *h_int, const int *lvl_ptr, const int *ghost);
void tqb_device_destroy(tqb_device *d);
/* async H2D of the input block, returns 0 on success */
int tqb_device_upload(tqb_device *d, const double *h_in, long long n_in, double *ms);
int tqb_device_upload_begin(tqb_device *d);
int tqb_device_upload_range(tqb_device *d, const double *h_in, long long off, long long n);
int tqb_device_upload_end(tqb_device *d, double *ms_total);
/* `rows` slices of `width` doubles at stride `pitch` starting at offset `off` of the input block */
int tqb_device_upload_2d(tqb_device *d, const double *h_in, long long off, long long pitch, long long width, int rows);
int tqb_device_warm_start(tqb_device *d, long long lam0_off, long long n);
/* Caller-owned host memory as a DMA source / target (device.cu): register / unregister by the owner
 * (treeqp_b200_host_register), tqb_host_pin = "is this range page-locked?" */
int tqb_host_register(const void *base, size_t bytes);
int tqb_host_unregister(const void *base);
int tqb_host_pin(const void *base, size_t bytes);
void tqb_host_unpin_all(void);
int tqb_device_stage_reserve(tqb_device *d, size_t bytes);
int tqb_device_upload_run(tqb_device *d, const tqb_in_run *r, int g_first, size_t stage_off);
int tqb_device_download_run(tqb_device *d, const tqb_out_run *r, int g_first, size_t stage_off);
/* D2H of the slices of x, u, lam, mu_x, mu_u that belong to nodes [ga, gb) into the pinned output block */
int tqb_device_download_nodes(tqb_device *d, double *h_out, long long x0, long long x1, long long u0, long long u1);
/* after a DMA-mode upload: Qinv = 1 / Qd, Rinv = 1 / Rd on the device, lam0 = 0 by memset when the host start is zero */
int tqb_device_finish_direct_upload(tqb_device *d, long long lam0_off, int lam0_zero);
int tqb_device_download_begin(tqb_device *d);
int tqb_device_download_end(tqb_device *d, double *ms);
int tqb_device_kkt(tqb_device *d, double *h_kkt);
int tqb_device_iter_phase_times(tqb_device *d, double *ph, int n);
/* the whole Newton loop; fills h_status/h_iter (pinned, nT) */
int tqb_device_solve(tqb_device *d, const tqb_opts *o, int *h_status, int *h_iter,
                     int *ls_iters256, int *ls_total, int *n_reg, double *gpu_ms, long long *launches,
                     double *kernel_ms16, int *kernel_count16);
void tqb_set_profiling(int on);
int tqb_num_kernel_classes(void);
const char *tqb_kernel_class_name(int k);
/* async D2H of x,u,lam,mu_x,mu_u into h_out */
int tqb_device_download(tqb_device *d, double *h_out, double *ms);
const char *tqb_cuda_last_error(void);


/* =========================================================================================
 * Subtree sharding of ONE tree QP over the GPUs of a box (TDUNES)
 *
 * The tree is cut at depth c (the first depth with at least `nranks` nodes).  The nodes of depth
 * <= c form the TOP, replicated on every rank; the nodes of depth c are the CUT nodes, dealt to the
 * ranks in equal contiguous groups; a rank's LOCAL tree is the top plus the subtrees below its own
 * cut nodes.  Cut nodes of other ranks are GHOSTS: leaves of the local tree whose stage solution and
 * Schur complement arrive by all-gather.  Every reduction counts a node on exactly one rank
 * (`owned`): the top on rank 0, a subtree (with its cut node) on its rank.
 * ========================================================================================= */
struct node;
typedef struct tqb_shard_plan
{
    int rank, nranks;
    int cut_depth, ncut;       /* cut nodes = all nodes of depth cut_depth */
    int cut_first;             /* global (= local) index of the first cut node */
    int my_cut0, my_ncut;      /* this rank's cut nodes: cut_first + my_cut0 .. + my_ncut */
    int Nl;                    /* local nodes */
    int *gmap;                 /* [Nl] local -> global node index */
    int *ghost;                /* [Nl] cut node of another rank */
    int *owned;                /* [Nl] counted by this rank in reductions / reported in qp_out */
    struct node *ltree;        /* [Nl] local tree (kids point into lkids) */
    int *lkids;
    int *lnx, *lnu, *lnc;      /* [Nl] */
    long long *lam_off;        /* [Nl] offset of local node j's multiplier in the caller's full lambda vector (filled on first use) */
} tqb_shard_plan;

/* returns NULL (with a message on stderr) if the tree cannot be cut for this number of ranks */
tqb_shard_plan *tqb_shard_plan_create(const struct tree_qp_in_ *qp, int rank, int nranks);
void tqb_shard_plan_destroy(tqb_shard_plan *p);

/* forest of ONE local tree described by `plan`; data is read from the FULL qp through plan->gmap */
tqb_forest *tqb_forest_create_sharded(const struct tree_qp_in_ *full_qp, const tqb_shard_plan *plan, void *comm);
void tqb_forest_pack_sharded(tqb_forest *f, const struct tree_qp_in_ *full_qp, const tqb_shard_plan *plan);
void tqb_forest_set_lambda0_sharded(tqb_forest *f, const double *lambda_full, const struct tree_qp_in_ *full_qp,
                                    const tqb_shard_plan *plan);
void tqb_forest_unpack_sharded(const tqb_forest *f, const tqb_shard_plan *plan, struct tree_qp_out_ *out_full);

/* ---- cuda/comm.cu: NCCL (loaded with dlopen) ---- */
typedef struct tqb_comm tqb_comm;
int tqb_comm_unique_id(char *id128);
tqb_comm *tqb_comm_create(int rank, int nranks, const char *id128, int device);
void tqb_comm_destroy(tqb_comm *c);
int tqb_comm_rank(const tqb_comm *c);
int tqb_comm_size(const tqb_comm *c);
/* device side of a sharded forest: cut-node exchange + scalar all-reduces on the solver stream */
int tqb_device_enable_sharding(tqb_device *d, tqb_comm *comm, int cut_depth, int ncut, int my_cut0, int my_ncut,
                               const int *cut_nodes_local);

/* =========================================================================================
 * SDUNES (dual Newton strategy on the scenario decomposition, dual_Newton_scenarios.c)
 *
 * Vocabulary
 *   scenario   one root-to-leaf path i of the multistage tree (Ns of them, Nh stages each)
 *   (i,k)      stage k of scenario i: state of tree node nodeIdx[i][k+1], input of node
 *              nodeIdx[i][k]; all per-(i,k) vectors are stored scenario-major, ik = i*Nh + k
 *   pair i     neighbouring scenarios (i, i+1): non-anticipativity multiplier lambda_i of
 *              dimension nu*cn[i] (cn = number of common nodes), stored at lo[i]
 *   node slot  per-node inputs use uniform slots: element a of node n at n*nx + a (n*nu + a)
 * ========================================================================================= */
typedef struct tqb_sd_dims
{
    int Nn, Np, Ns, Nh, Nr, md, nx, nu;
    int nK;              /* nu*Nr: rows of Ut_i, dimension of K_i */
    int nl;              /* total dimension of lambda */
    int SK;              /* Ns*Nh */
    int njd, njl, mJ;    /* doubles in the J diagonal / sub-diagonal blocks, largest J block */
    int n_int;
    long long n_in, n_out;
} tqb_sd_dims;

typedef struct tqb_sd_tables
{
    int *nodeIdx, *removed;   /* [Ns*(Nh+1)] tree node of (i,k), bounds-removed flag (scenarios.c:1650-1668) */
    int *cn;                  /* [Ns] common nodes of pair i (last entry unused) */
    int *lo, *jdo, *jlo;      /* [Ns+1] offsets of lambda_i / J diagonal block i / J sub-diagonal block i */
    int *pair_of;             /* [nl] pair owning each lambda element */
    int *nfirst, *ncount;     /* [Nn] first scenario through a node, number of scenarios through it */
    int *nstage;              /* [Nn] depth of the node */
} tqb_sd_tables;

typedef struct tqb_sd_inputs
{
    double *A, *B, *b;                       /* node slots: nx*nx, nx*nu, nx */
    double *Q, *q, *Qinv, *xmin, *xmax;      /* node slots nx (weights un-scaled, scenarios.c:1885-1898) */
    double *R, *r, *Rinv, *umin, *umax;      /* node slots nu */
    double *lam0, *mu0;                      /* [nl], [SK*nx] dual start */
} tqb_sd_inputs;

/* output block (one D2H copy): node results, then the scenario duals and iterates */
typedef struct tqb_sd_outputs
{
    double *x, *u, *lam, *mux, *muu;         /* node slots nx / nu */
    double *lam_s, *mu_s;                    /* [nl], [SK*nx] */
    double *xs, *us;                         /* [SK*nx], [SK*nu] scenario iterates */
} tqb_sd_outputs;

void tqb_sd_layout_tables(const tqb_sd_dims *d, int *base, tqb_sd_tables *t);
void tqb_sd_layout_inputs(const tqb_sd_dims *d, double *base, tqb_sd_inputs *in);
void tqb_sd_layout_outputs(const tqb_sd_dims *d, double *base, tqb_sd_outputs *out);

#define TQB_SD_INFO_INTS (8 + 256 + 16) /* status, iter, ls_total, n_reg, 4 spare, trials per iteration, ns per phase */
#define TQB_SD_NPHASE 10

typedef struct tqb_sd_device tqb_sd_device;
tqb_sd_device *tqb_sd_device_create(const tqb_sd_dims *dims, const int *h_int);
void tqb_sd_device_destroy(tqb_sd_device *d);
int tqb_sd_device_upload(tqb_sd_device *d, const double *h_in, double *ms);
/* the whole Newton loop in ONE persistent kernel; h_info is pinned [TQB_SD_INFO_INTS] */
int tqb_sd_device_solve(tqb_sd_device *d, const tqb_opts *o, int reverse, int *h_info, double *gpu_ms,
                        long long *launches);
int tqb_sd_device_download(tqb_sd_device *d, double *h_out, double *ms);

int tqb_verbose(void); /* TREEQP_B200_VERBOSE=1 */
void tqb_set_profiling(int on); /* device.cu: CUDA events around every launch */
#if defined(PROFILE) && PROFILE > 0
/* profiling.c: one solve's per-iteration phase seconds (4 x n), iterations run, trials per iteration, total seconds */
struct treeqp_profiling_t_;
void tqb_timers_fill(struct treeqp_profiling_t_ *t, const double *ph, int n, int nit, const int *ls_iters, double total_time);
#endif
/* "%.16e" / "%d" per line, as utils/utils.c:98-143 */
int tqb_write_doubles_txt(const double *v, long long n, const char *filename);
int tqb_write_ints_txt(const int *v, long long n, const char *filename);

#ifdef __cplusplus
}
#endif
#endif
