/* comm.cuh -- the two collectives the sharded TDUNES solve issues on its stream (cuda/comm.cu) */
#pragma once
#include <cuda_runtime.h>
#include "../host/host_internal.h"

/* all-gather of `count` doubles per rank */
int tqb_comm_allgather(tqb_comm *c, const double *send, double *recv, size_t count, cudaStream_t s);
/* in-place all-reduce (sum or max) of `count` doubles */
int tqb_comm_allreduce(tqb_comm *c, double *buf, size_t count, int is_max, cudaStream_t s);
