/* comm.cuh -- the exchanges the sharded TDUNES solve issues on its stream (cuda/comm.cu):
 * NCCL collectives (set-up, fallback) and the peer-mailbox exchange over NVLink that replaces them on
 * the data path (all messages of a sharded solve are < 12 KB: an NCCL launch costs 20-40 us, a store
 * into a peer's memory through NVSwitch plus a flag costs a few). */
#pragma once
#include <cuda_runtime.h>
#include "../host/host_internal.h"

/* all-gather of `count` doubles per rank */
int tqb_comm_allgather(tqb_comm *c, const double *send, double *recv, size_t count, cudaStream_t s);
/* in-place all-reduce (sum or max) of `count` doubles */
int tqb_comm_allreduce(tqb_comm *c, double *buf, size_t count, int is_max, cudaStream_t s);

/* ---- peer mailboxes -------------------------------------------------------------------------
 * Every rank owns one mailbox in its HBM: 2 (parity) x nranks (source) data slots of `cap` doubles and
 * as many 64-bit flags.  The mailboxes are mapped into every peer with CUDA IPC.  One exchange =
 * every rank stores its payload into slot [seq & 1][my rank] of EVERY mailbox (its own included),
 * fences, stores `seq` into the matching flag, then polls the flags of its OWN mailbox until all
 * sources show `seq`.  Two parities suffice: a rank can only start exchange s + 2 after it finished
 * s + 1, which needs every peer's contribution to s + 1, which a peer sends after it consumed s. */
#define TQB_P2P_MAXR 16
struct tqb_p2p_dev
{
    double *slots[TQB_P2P_MAXR];               /* data base of rank r's mailbox, in THIS rank's address space */
    unsigned long long *flags[TQB_P2P_MAXR];   /* [2][TQB_P2P_MAXR] flags of rank r's mailbox */
    int *err;                                  /* local: set when a peer did not show up in time */
    int rank, nranks, cap;
};
/* allocate + map the mailboxes (collective: every rank calls it with the same cap); 0 = ready,
 * 1 = not available (no peer access, IPC refused, TREEQP_B200_NO_P2P=1): the caller keeps NCCL */
int tqb_comm_p2p_setup(tqb_comm *c, size_t cap_doubles, cudaStream_t s);
const tqb_p2p_dev *tqb_comm_p2p(const tqb_comm *c);
unsigned long long tqb_comm_p2p_next_seq(tqb_comm *c);
int tqb_comm_p2p_error(tqb_comm *c, cudaStream_t s);
