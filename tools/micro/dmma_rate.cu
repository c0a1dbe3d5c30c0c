// Microbenchmark: FP64 throughput of DFMA vs mma.sync.m8n8k4.f64 (DMMA) on one GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double acc[8];
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    double s = 0; for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 0.5 * i; }
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(double *out, int iters, double a, double b)
{   // 1 DMMA + 4 DFMA interleaved: do they share a pipe?
    double c[4][2], acc[8];
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 0.5 * i; }
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            acc[2 * i] = fma(acc[2 * i], a, b);
            acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
        }
    }
    double s = 0; for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148 * 4, block = 512;
    float ms;
    for (int rep = 0; rep < 2; rep++)
    {
        cudaEventRecord(e0); k_dfma<<<grid, block>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA : %.2f TFLOP/s (%.3f ms)\n", 2.0 * grid * block * 8.0 * iters / (ms * 1e9), ms);
        cudaEventRecord(e0); k_dmma<<<grid, block>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA : %.2f TFLOP/s (%.3f ms)  [m8n8k4 = 512 flop/warp-instr]\n", 512.0 * grid * (block / 32) * 8.0 * iters / (ms * 1e9), ms);
        cudaEventRecord(e0); k_mix<<<grid, block>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("MIX  : %.3f ms for %d x (4 DMMA + 8 DFMA) per warp; DMMA-only would be %.3f, DFMA-only %.3f\n", ms, iters, 0.0, 0.0);
    }
    // latency: single warp, dependent chain
    cudaEventRecord(e0); k_dmma<<<1, 32>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA 1 warp, 8 independent chains: %.1f ns per DMMA\n", ms * 1e6 / (8.0 * iters));
    cudaEventRecord(e0); k_dfma<<<1, 32>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA 1 warp, 8 independent chains: %.1f ns per DFMA\n", ms * 1e6 / (8.0 * iters));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
