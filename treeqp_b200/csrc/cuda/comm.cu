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
    /* peer mailboxes (tqb_comm_p2p_setup) */
    int p2p_ready;
    tqb_p2p_dev p2p;
    void *box;                       /* this rank's mailbox allocation */
    void *peer_box[TQB_P2P_MAXR];    /* IPC mappings of the peers' */
    unsigned long long seq;
    int *h_err;
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

/* ---- peer mailboxes over NVLink (see comm.cuh) ---- */
#define P2P_HEADER_BYTES 4096 /* flags [2][MAXR] u64 + error word, then the data slots */

int tqb_comm_p2p_setup(tqb_comm *c, size_t cap_doubles, cudaStream_t s)
{
    if (c->p2p_ready) return (size_t) c->p2p.cap >= cap_doubles ? 0 : 1;
    const char *e = getenv("TREEQP_B200_NO_P2P");
    int ok = !(e && e[0] == '1') && c->nranks <= TQB_P2P_MAXR && c->nranks > 1;
    const size_t bytes = P2P_HEADER_BYTES + sizeof(double) * 2 * (size_t) c->nranks * cap_doubles;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaMalloc(&c->box, bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(c->box, 0, bytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, c->box) != cudaSuccess) ok = 0;
    /* every rank takes part in the hand-shake whatever its own outcome: handle + ok flag, all-gathered */
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    char *h_send = (char *) calloc(1, rec), *h_all = (char *) calloc((size_t) c->nranks, rec);
    memcpy(h_send, &mine, sizeof mine);
    h_send[sizeof mine] = (char) ok;
    char *d_send = NULL, *d_all = NULL;
    if (cudaMalloc(&d_send, rec) != cudaSuccess || cudaMalloc(&d_all, rec * (size_t) c->nranks) != cudaSuccess) return 1;
    cudaMemcpyAsync(d_send, h_send, rec, cudaMemcpyHostToDevice, s);
    NCK(g_nccl.AllGather(d_send, d_all, rec, ncclChar, c->comm, s));
    cudaMemcpyAsync(h_all, d_all, rec * (size_t) c->nranks, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    for (int r = 0; r < c->nranks; r++) ok = ok && h_all[(size_t) r * rec + sizeof mine];
    for (int r = 0; r < c->nranks && ok; r++)
    {
        if (r == c->rank) { c->peer_box[r] = c->box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_all + (size_t) r * rec, sizeof h);
        if (cudaIpcOpenMemHandle(&c->peer_box[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
            fprintf(stderr, "[treeqp_b200] rank %d: cudaIpcOpenMemHandle(rank %d) failed (%s): NCCL collectives stay on the data path\n",
                    c->rank, r, cudaGetErrorString(cudaGetLastError()));
            ok = 0;
        }
    }
    /* agree on the outcome: one more tiny all-gather (a rank that failed to map a peer vetoes) */
    h_send[0] = (char) ok;
    cudaMemcpyAsync(d_send, h_send, 1, cudaMemcpyHostToDevice, s);
    NCK(g_nccl.AllGather(d_send, d_all, 1, ncclChar, c->comm, s));
    cudaMemcpyAsync(h_all, d_all, (size_t) c->nranks, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    for (int r = 0; r < c->nranks; r++) ok = ok && h_all[r];
    cudaFree(d_send); cudaFree(d_all); free(h_send); free(h_all);
    if (!ok) return 1;
    for (int r = 0; r < c->nranks; r++)
    {
        c->p2p.flags[r] = (unsigned long long *) c->peer_box[r];
        c->p2p.slots[r] = (double *) ((char *) c->peer_box[r] + P2P_HEADER_BYTES);
    }
    c->p2p.err = (int *) ((char *) c->box + sizeof(unsigned long long) * 2 * TQB_P2P_MAXR);
    c->p2p.rank = c->rank; c->p2p.nranks = c->nranks; c->p2p.cap = (int) cap_doubles;
    if (cudaMallocHost(&c->h_err, sizeof(int)) != cudaSuccess) return 1;
    *c->h_err = 0;
    c->seq = 0;
    c->p2p_ready = 1;
    return 0;
}

const tqb_p2p_dev *tqb_comm_p2p(const tqb_comm *c) { return (c && c->p2p_ready) ? &c->p2p : NULL; }
unsigned long long tqb_comm_p2p_next_seq(tqb_comm *c) { return ++c->seq; }

/* stream-ordered read of the time-out word (call before a synchronisation, check *after* it) */
int tqb_comm_p2p_error(tqb_comm *c, cudaStream_t s)
{
    if (!c->p2p_ready) return 0;
    if (*c->h_err) return 1;
    cudaMemcpyAsync(c->h_err, c->p2p.err, sizeof(int), cudaMemcpyDeviceToHost, s);
    return 0;
}

extern "C" void tqb_comm_destroy(tqb_comm *c)
{
    if (!c) return;
    if (c->p2p_ready)
    {
        cudaDeviceSynchronize();
        for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && c->peer_box[r]) cudaIpcCloseMemHandle(c->peer_box[r]);
        cudaFreeHost(c->h_err);
    }
    if (c->box) cudaFree(c->box);
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
