/*
 * mpc_loop_registered.c -- the reference's call sequence (examples/spring_mass_dual_newton_tree.c:52-158) on a large random
 * tree QP, written against include/treeqp_b200.h only, with the one addition INTEGRATION.md suggests for a loop over large
 * trees: the two buffers the caller owns (the `ptr` of tree_qp_in_create / tree_qp_out_create) are allocated page-aligned
 * and page-locked once with treeqp_b200_host_register, so that every treeqp_tdunes_solve moves the tree by DMA straight out
 * of / into them.
 *
 *   gcc -O2 -std=c99 -Iinclude examples/mpc_loop_registered.c -Ltreeqp_b200/lib -ltreeqp_b200 -lm -Wl,-rpath,$PWD/treeqp_b200/lib
 *   ./a.out [depth = 13] [samples = 5]
 *
 * Solves the QP `samples` times through the pageable path, then registers the buffers and solves again: same bits, KKT
 * residual below the reference's 1e-8 bar, and the two wall times per solve.  Exit code 0 only if all of that holds.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "treeqp_b200.h"

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}

static unsigned long long rng_state = 88172645463325252ULL;
static double urand(void)
{   /* xorshift64: uniform in [0, 1) */
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double) (rng_state >> 11) / 9007199254740992.0;
}

static void *page_alloc(size_t bytes, size_t *rounded)
{
    void *p = NULL;
    *rounded = (bytes + 4095) / 4096 * 4096;
    if (posix_memalign(&p, 4096, *rounded)) return NULL;
    memset(p, 0, *rounded);
    return p;
}

int main(int argc, char **argv)
{
    const int depth = argc > 1 ? atoi(argv[1]) : 13, samples = argc > 2 ? atoi(argv[2]) : 5;
    const int NX = 12, NU = 4, Nn = (1 << (depth + 1)) - 1, Np = (1 << depth) - 1;
    if (treeqp_b200_device_count() <= 0) { fprintf(stderr, "no CUDA device\n"); return 2; }

    /* binary tree: every node of depth < `depth` has two children */
    int *nk = malloc(sizeof(int) * Nn), *nx = malloc(sizeof(int) * Nn), *nu = malloc(sizeof(int) * Nn);
    for (int k = 0; k < Nn; k++) { nk[k] = k < Np ? 2 : 0; nx[k] = NX; nu[k] = k < Np ? NU : 0; }

    size_t in_bytes, out_bytes;
    void *in_mem = page_alloc((size_t) tree_qp_in_calculate_size(Nn, nx, nu, NULL, nk), &in_bytes);
    void *out_mem = page_alloc((size_t) tree_qp_out_calculate_size(Nn, nx, nu, NULL), &out_bytes);
    tree_qp_in qp_in;
    tree_qp_out qp_out;
    tree_qp_in_create(Nn, nx, nu, NULL, nk, &qp_in, in_mem);
    tree_qp_out_create(Nn, nx, nu, NULL, &qp_out, out_mem);

    /* data: A = (0.96 / nx) U(0,1), B, b, q, r ~ U(0,1), Q in [1, 10], R in [0.5, 1.5], |u| <= 0.2, x free */
    const size_t X = (size_t) Nn * NX, U = (size_t) Np * NU, E = (size_t) Nn - 1;
    double *A = malloc(sizeof(double) * E * NX * NX), *B = malloc(sizeof(double) * E * NX * NU), *b = malloc(sizeof(double) * E * NX);
    double *Qd = malloc(sizeof(double) * X), *Rd = malloc(sizeof(double) * U), *q = malloc(sizeof(double) * X), *r = malloc(sizeof(double) * U);
    double *xmin = malloc(sizeof(double) * X), *xmax = malloc(sizeof(double) * X), *umin = malloc(sizeof(double) * U), *umax = malloc(sizeof(double) * U);
    for (size_t i = 0; i < E * NX * NX; i++) A[i] = 0.96 / NX * urand();
    for (size_t i = 0; i < E * NX * NU; i++) B[i] = urand();
    for (size_t i = 0; i < E * NX; i++) b[i] = urand();
    for (size_t i = 0; i < X; i++) { Qd[i] = 1.0 + 9.0 * urand(); q[i] = urand(); xmin[i] = -1e12; xmax[i] = 1e12; }
    for (size_t i = 0; i < U; i++) { Rd[i] = 0.5 + urand(); r[i] = urand(); umin[i] = -0.2; umax[i] = 0.2; }
    tree_qp_in_set_ltv_dynamics_colmajor(A, B, b, &qp_in);
    tree_qp_in_set_ltv_objective_diag(Qd, Rd, q, r, &qp_in);
    tree_qp_in_set_ltv_bounds(xmin, xmax, umin, umax, &qp_in);

    /* the reference's septet */
    treeqp_tdunes_opts_t opts;
    void *opts_mem = malloc((size_t) treeqp_tdunes_opts_calculate_size(Nn));
    treeqp_tdunes_opts_create(Nn, &opts, opts_mem);
    treeqp_tdunes_opts_set_default(Nn, &opts);
    opts.maxIter = 200;
    treeqp_tdunes_workspace work;
    void *work_mem = malloc((size_t) treeqp_tdunes_calculate_size(&qp_in, &opts));
    treeqp_tdunes_create(&qp_in, &opts, &work, work_mem);
    double *lam0 = calloc(E * NX, sizeof(double));

    double *x_ref = malloc(sizeof(double) * X);
    int iter_ref = -1, ok = 1;
    double ms[2] = {1e30, 1e30};
    for (int pass = 0; pass < 2; pass++)
    {
        if (pass == 1 && (treeqp_b200_host_register(in_mem, in_bytes) || treeqp_b200_host_register(out_mem, out_bytes)))
        {
            fprintf(stderr, "treeqp_b200_host_register failed\n");
            return 3;
        }
        for (int s = 0; s < samples; s++)
        {
            treeqp_tdunes_set_dual_initialization(lam0, &work);
            const double t0 = now_ms();
            const return_t st = treeqp_tdunes_solve(&qp_in, &qp_out, &opts, &work);
            const double t = now_ms() - t0;
            if (t < ms[pass]) ms[pass] = t;
            if (st != TREEQP_OPTIMAL_SOLUTION_FOUND) { fprintf(stderr, "status %d\n", (int) st); ok = 0; }
            for (int k = 0; k < Nn; k++)
            {
                double *xk = x_ref + (size_t) k * NX, cur[12];
                tree_qp_out_get_node_x(cur, &qp_out, k);
                if (pass == 0 && s == 0) memcpy(xk, cur, sizeof cur);
                else if (memcmp(xk, cur, sizeof cur)) { if (ok) fprintf(stderr, "x of node %d differs (pass %d, sample %d)\n", k, pass, s); ok = 0; }
            }
            if (iter_ref < 0) iter_ref = qp_out.info.iter;
            else if (iter_ref != qp_out.info.iter) ok = 0;
        }
    }
    const double kkt = tree_qp_out_max_KKT_res(&qp_in, &qp_out);
    printf("binary tree depth %d (%d nodes): %d Newton iterations, KKT %.2e, %.2f ms per solve pageable, %.2f ms page-locked, results %s\n",
           depth, Nn, iter_ref, kkt, ms[0], ms[1], ok ? "bit-identical" : "DIFFERENT");
    if (!(kkt < 1e-8)) ok = 0;

    treeqp_tdunes_destroy(&work);
    treeqp_b200_host_unregister(in_mem);   /* before the buffers are freed */
    treeqp_b200_host_unregister(out_mem);
    free(in_mem); free(out_mem);
    return ok ? 0 : 1;
}
