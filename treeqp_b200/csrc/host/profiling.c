/*
 * profiling.c -- the reference's timers (treeqp/utils/profiling.c, treeqp/utils/print.c:312-345) for the PROFILE > 0 build
 * of the library (libtreeqp_b200_profile3.so): storage carved from the caller's workspace memory, NaN-initialised
 * per-iteration arrays, minima over repeated runs, the reference's report.  The numbers come from CUDA events around every
 * launch (tqb_device_iter_phase_times), not from host timers around BLAS calls.  Empty in the default (PROFILE = 0) build.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "treeqp_b200.h"
#include "host_internal.h"

#if defined(PROFILE) && PROFILE > 0

/* bytes behind a treeqp_profiling_t: 2 (PROFILE > 1) + 8 (PROFILE > 2) double arrays and one int array of num_iter entries */
int timers_calculate_size(int num_iter)
{
    int bytes = 0;
#if PROFILE > 1
    bytes += (int) (2 * num_iter * sizeof(double) + num_iter * sizeof(int));
#endif
#if PROFILE > 2
    bytes += (int) (8 * num_iter * sizeof(double));
#endif
    return bytes;
}

void timers_create(int num_iter, treeqp_profiling_t *t, void *ptr)
{
    double *d = (double *) ptr;
    (void) d;
    t->num_iter = num_iter;
#if PROFILE > 1
    t->iter_times = d; d += num_iter;
    t->min_iter_times = d; d += num_iter;
#endif
#if PROFILE > 2
    double **arr[8] = {&t->stage_qps_times, &t->build_dual_times, &t->newton_direction_times, &t->line_search_times,
                       &t->min_stage_qps_times, &t->min_build_dual_times, &t->min_newton_direction_times, &t->min_line_search_times};
    for (int i = 0; i < 8; i++) { *arr[i] = d; d += num_iter; }
#endif
#if PROFILE > 1
    t->ls_iters = (int *) d;
#endif
}

void timers_initialize(treeqp_profiling_t *t)
{
    t->total_time = NAN;
    t->run_indx = 0;
#if PROFILE > 1
    for (int i = 0; i < t->num_iter; i++) { t->iter_times[i] = NAN; t->ls_iters[i] = 0; }
#endif
#if PROFILE > 2
    for (int i = 0; i < t->num_iter; i++)
        t->stage_qps_times[i] = t->build_dual_times[i] = t->newton_direction_times[i] = t->line_search_times[i] = NAN;
#endif
}

/* keep the minimum of a timer over the runs; fmin would drop the NaN of iterations that did not run, the reference's MIN
 * macro (a < b ? a : b) keeps whatever sits in `b` then: reproduced */
static void keep_min(double *mn, const double *cur, int n, int first)
{
    for (int i = 0; i < n; i++) mn[i] = first ? cur[i] : (mn[i] < cur[i] ? mn[i] : cur[i]);
}

void timers_update(treeqp_profiling_t *t)
{
    const int first = t->run_indx == 0;
    keep_min(&t->min_total_time, &t->total_time, 1, first);
#if PROFILE > 1
    keep_min(t->min_iter_times, t->iter_times, t->num_iter, first);
    if (first)
    {
        t->total_ls_iter = 0;
        for (int i = 0; i < t->num_iter; i++) t->total_ls_iter += t->ls_iters[i];
    }
#endif
#if PROFILE > 2
    keep_min(t->min_stage_qps_times, t->stage_qps_times, t->num_iter, first);
    keep_min(t->min_build_dual_times, t->build_dual_times, t->num_iter, first);
    keep_min(t->min_newton_direction_times, t->newton_direction_times, t->num_iter, first);
    keep_min(t->min_line_search_times, t->line_search_times, t->num_iter, first);
#endif
    t->run_indx++;
}

void timers_print(treeqp_profiling_t *t)
{
#if PROFILE > 1
    int nit = 0;
    while (nit < t->num_iter && !isnan(t->iter_times[nit])) nit++;
    printf("\nTotal time:\n\n> > > algorithm converged in (%d it):\t %10.4f ms\n\n", nit, t->min_total_time * 1e3);
    printf("\nTimings per iteration:\n\n");
    for (int j = 0; j < nit; j++)
        printf("Iteration #%3d - %7.3f ms  (%3d ls iters. )\n", j + 1, t->min_iter_times[j] * 1e3, t->ls_iters[j]);
#else
    printf("\nTotal time:\n\n> > > algorithm converged in:\t %10.4f ms\n\n", t->min_total_time * 1e3);
#endif
#if PROFILE > 2
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < nit; j++)
    {
        s[0] += t->min_stage_qps_times[j]; s[1] += t->min_build_dual_times[j];
        s[2] += t->min_newton_direction_times[j]; s[3] += t->min_line_search_times[j];
    }
    const double all = s[0] + s[1] + s[2] + s[3];
    printf("\nTimings per operation (CUDA events around every launch, B200):\n\n");
    printf("> > > solved stage QPs in:\t\t %10.4f ms (%5.2f %%)\n", s[0] * 1e3, 100 * s[0] / all);
    printf("> > > built dual problem in:\t\t %10.4f ms (%5.2f %%)\n", s[1] * 1e3, 100 * s[1] / all);
    printf("> > > calculated Newton direction in: \t %10.4f ms (%5.2f %%)\n", s[2] * 1e3, 100 * s[2] / all);
    printf("> > > performed line-search (%d it) in:\t %10.4f ms (%5.2f %%)\n", t->total_ls_iter, s[3] * 1e3, 100 * s[3] / all);
    printf("> > > sum all of the above:\t\t %10.4f ms\n", all * 1e3);
#endif
}

void timers_write_to_txt(treeqp_profiling_t *t)
{   /* the files of utils/print.c:312-345 (same names; cputime / iter_times only below PROFILE 3, as there) */
#if PROFILE > 1
    tqb_write_ints_txt(t->ls_iters, t->num_iter, "examples/spring_mass_utils/ls_iters.txt");
#endif
#if PROFILE < 3
    tqb_write_doubles_txt(&t->min_total_time, 1, "examples/spring_mass_utils/cputime.txt");
#if PROFILE > 1
    tqb_write_doubles_txt(t->min_iter_times, t->num_iter, "examples/spring_mass_utils/iter_times.txt");
#endif
#endif
#if PROFILE > 2
    tqb_write_doubles_txt(t->min_stage_qps_times, t->num_iter, "examples/spring_mass_utils/stage_qps_times.txt");
    tqb_write_doubles_txt(t->min_build_dual_times, t->num_iter, "examples/spring_mass_utils/build_dual_times.txt");
    tqb_write_doubles_txt(t->min_newton_direction_times, t->num_iter, "examples/spring_mass_utils/newton_direction_times.txt");
    tqb_write_doubles_txt(t->min_line_search_times, t->num_iter, "examples/spring_mass_utils/line_search_times.txt");
#endif
}

/* one solve's numbers into `t` (what the reference's solve does around its phases, dual_Newton_tree.c:1168-1226, :1257-1260):
 * ph = 4 x n phase seconds per iteration, nit = iterations that ran */
void tqb_timers_fill(treeqp_profiling_t *t, const double *ph, int n, int nit, const int *ls_iters, double total_time)
{
    (void) ph; (void) n; (void) ls_iters;
#if PROFILE > 1
    for (int i = 0; i < nit && i < t->num_iter; i++)
    {
        t->iter_times[i] = ph[i] + ph[n + i] + ph[2 * n + i] + ph[3 * n + i];
        t->ls_iters[i] = (ls_iters && i < 256) ? ls_iters[i] : 0;
    }
#endif
#if PROFILE > 2
    for (int i = 0; i < nit && i < t->num_iter; i++)
    {
        t->stage_qps_times[i] = ph[i]; t->build_dual_times[i] = ph[n + i];
        t->newton_direction_times[i] = ph[2 * n + i]; t->line_search_times[i] = ph[3 * n + i];
    }
#endif
    (void) nit;
    t->total_time = total_time;
    timers_update(t);
}

#endif /* PROFILE > 0 */
