/*
 * registry.c -- device arenas by the caller's workspace memory.
 *
 * The reference API has create but no destroy (dual_Newton_tree.h:154-168, dual_Newton_scenarios.h:143-163): a
 * caller that re-creates its workspace in a loop over the same `ptr` -- the benchmark harnesses do, once per QP --
 * must not leak a device arena and pinned staging buffers per create.  Every create registers its context under
 * the caller's `ptr`; a create on a `ptr` that is already registered releases the previous context first (the
 * workspace that lived there is necessarily dead: its memory is being carved again).  treeqp_*_destroy removes
 * the entry.
 */
#include <pthread.h>
#include <stdlib.h>

#include "host_internal.h"

typedef struct { const void *key; void *ctx; void (*dtor)(void *); } tqb_reg_entry;
static tqb_reg_entry *g_reg = NULL;
static int g_reg_n = 0, g_reg_cap = 0;
static pthread_mutex_t g_reg_mu = PTHREAD_MUTEX_INITIALIZER;

void tqb_registry_replace(const void *key, void *ctx, void (*dtor)(void *))
{
    void *old = NULL;
    void (*old_dtor)(void *) = NULL;
    pthread_mutex_lock(&g_reg_mu);
    int i = 0;
    for (; i < g_reg_n; i++)
        if (g_reg[i].key == key) break;
    if (i < g_reg_n) { old = g_reg[i].ctx; old_dtor = g_reg[i].dtor; }
    else
    {
        if (g_reg_n == g_reg_cap)
        {
            const int cap = g_reg_cap ? 2 * g_reg_cap : 16;
            tqb_reg_entry *r = (tqb_reg_entry *) realloc(g_reg, sizeof(tqb_reg_entry) * (size_t) cap);
            if (!r)
            {   /* out of host memory: the context stays unregistered (released by treeqp_*_destroy only) */
                pthread_mutex_unlock(&g_reg_mu);
                return;
            }
            g_reg = r; g_reg_cap = cap;
        }
        g_reg_n++;
    }
    g_reg[i].key = key; g_reg[i].ctx = ctx; g_reg[i].dtor = dtor;
    pthread_mutex_unlock(&g_reg_mu);
    if (old && old != ctx && old_dtor) old_dtor(old);
}

void tqb_registry_remove(const void *ctx)
{
    pthread_mutex_lock(&g_reg_mu);
    for (int i = 0; i < g_reg_n; i++)
        if (g_reg[i].ctx == ctx) { g_reg[i] = g_reg[--g_reg_n]; break; }
    pthread_mutex_unlock(&g_reg_mu);
}

int tqb_registry_size(void)
{
    pthread_mutex_lock(&g_reg_mu);
    const int n = g_reg_n;
    pthread_mutex_unlock(&g_reg_mu);
    return n;
}
