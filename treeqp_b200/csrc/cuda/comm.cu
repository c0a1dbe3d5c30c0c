/*
 * comm.cu -- NCCL communicator of the sharded solves.  One process per GPU; the unique id is
 * created on rank 0 and handed to the other ranks by whatever the host program already uses
 * (torch.distributed in bench.py / the tests, MPI or a file in a C program).
 *
 * NCCL is loaded with dlopen at communicator creation, not linked: libtreeqp_b200.so keeps
 * loading on machines without NCCL (single-GPU use), and inside a process that already loaded
 * NCCL (PyTorch) the same instance is reused.
 *
 * Traffic per Newton iteration of a sharded solve: ONE all-gather of nx(nx+3) doubles
 * per cut node (stage solution, active-set mask, Schur complement and right-hand side of the cut
 * node's block) and one all-reduce of <= 2 doubles per decision (termination norm; descent
 * product + f(0); f(tau) per line-search trial).  All messages are < 10 KB: latency bound, issued
 * on the solver's stream between the level kernels.
 */
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "comm.cuh"

struct NcclApi
{
    void *handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int load_nccl(void)
{
    if (g_nccl.handle) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = NULL;
    for (int i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h)
    {
        fprintf(stderr, "[treeqp_b200] NCCL not found (%s): sharded solves are unavailable\n", dlerror());
        return 1;
    }
#define SYM(field, name)                                                             \
    *(void **) (&g_nccl.field) = dlsym(h, name);                                     \
    if (!g_nccl.field) { fprintf(stderr, "[treeqp_b200] NCCL symbol %s missing\n", name); dlclose(h); return 1; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return 0;
}

struct tqb_comm
{
    ncclComm_t comm;
    int rank, nranks, device;
};

#define NCK(call)                                                                                   \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            fprintf(stderr, "[treeqp_b200] NCCL error %s:%d %s: %s\n", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

extern "C" int tqb_comm_unique_id(char *id128)
{
    if (load_nccl()) return 1;
    ncclUniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof id == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof id);
    return 0;
}

extern "C" tqb_comm *tqb_comm_create(int rank, int nranks, const char *id128, int device)
{
    if (load_nccl()) return NULL;
    if (cudaSetDevice(device) != cudaSuccess)
    {
        fprintf(stderr, "[treeqp_b200] cudaSetDevice(%d) failed\n", device);
        return NULL;
    }
    tqb_comm *c = (tqb_comm *) calloc(1, sizeof *c);
    c->rank = rank; c->nranks = nranks; c->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess)
    {
        fprintf(stderr, "[treeqp_b200] ncclCommInitRank failed: %s\n", g_nccl.GetErrorString(r));
        free(c);
        return NULL;
    }
    return c;
}

extern "C" void tqb_comm_destroy(tqb_comm *c)
{
    if (!c) return;
    if (c->comm) g_nccl.CommDestroy(c->comm);
    free(c);
}

extern "C" int tqb_comm_rank(const tqb_comm *c) { return c ? c->rank : 0; }
extern "C" int tqb_comm_size(const tqb_comm *c) { return c ? c->nranks : 1; }

int tqb_comm_allgather(tqb_comm *c, const double *send, double *recv, size_t count, cudaStream_t s)
{
    NCK(g_nccl.AllGather(send, recv, count, ncclDouble, c->comm, s));
    return 0;
}

int tqb_comm_allreduce(tqb_comm *c, double *buf, size_t count, int is_max, cudaStream_t s)
{
    NCK(g_nccl.AllReduce(buf, buf, count, ncclDouble, is_max ? ncclMax : ncclSum, c->comm, s));
    return 0;
}
