// Which scalar expression does mma.sync.m8n8k4.f64 reproduce bit for bit?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_order dmma_order.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double *A, const double *B, const double *C, double *Dm, int ntile)
{   // one warp per tile: A 8x4 row-major, B 4x8 row-major (k x n), C/D 8x8 row-major
    int t = blockIdx.x, lane = threadIdx.x, gid = lane >> 2, tig = lane & 3;
    const double *a = A + t * 32, *b = B + t * 32, *c = C + t * 64;
    double c0 = c[gid * 8 + 2 * tig], c1 = c[gid * 8 + 2 * tig + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a[gid * 4 + tig]), "d"(b[tig * 8 + gid]));
    Dm[t * 64 + gid * 8 + 2 * tig] = c0; Dm[t * 64 + gid * 8 + 2 * tig + 1] = c1;
}
int main()
{
    const int nt = 4096;
    double *A = (double *) malloc(nt * 32 * 8), *B = (double *) malloc(nt * 32 * 8), *C = (double *) malloc(nt * 64 * 8), *D = (double *) malloc(nt * 64 * 8);
    srand(1);
    for (int i = 0; i < nt * 32; i++) { A[i] = (rand() / (double) RAND_MAX - 0.5) * exp2(rand() % 20 - 10); B[i] = (rand() / (double) RAND_MAX - 0.5) * exp2(rand() % 20 - 10); }
    for (int i = 0; i < nt * 64; i++) C[i] = (i & 1) ? 0.0 : (rand() / (double) RAND_MAX - 0.5);
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, nt * 32 * 8); cudaMalloc(&dB, nt * 32 * 8); cudaMalloc(&dC, nt * 64 * 8); cudaMalloc(&dD, nt * 64 * 8);
    cudaMemcpy(dA, A, nt * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B, nt * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dC, C, nt * 64 * 8, cudaMemcpyHostToDevice);
    k<<<nt, 32>>>(dA, dB, dC, dD, nt);
    cudaMemcpy(D, dD, nt * 64 * 8, cudaMemcpyDeviceToHost);
    long n = 0, m_fwd = 0, m_rev = 0, m_pair = 0, m_unfused = 0, m_sumfirst = 0;
    double maxrel = 0;
    for (int t = 0; t < nt; t++)
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++)
            {
                const double *a = A + t * 32 + i * 4, c = C[t * 64 + i * 8 + j];
                double b[4]; for (int q = 0; q < 4; q++) b[q] = B[t * 32 + q * 8 + j];
                double d = D[t * 64 + i * 8 + j];
                double fwd = fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], fma(a[0], b[0], c))));
                double rev = fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], c))));
                double pair = (fma(a[1], b[1], a[0] * b[0]) + fma(a[3], b[3], a[2] * b[2])) + c;
                double unf = (((c + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3];
                double sf = c + fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
                n++; m_fwd += d == fwd; m_rev += d == rev; m_pair += d == pair; m_unfused += d == unf; m_sumfirst += d == sf;
                double rel = fabs(d - fwd) / (fabs(fwd) + 1e-300); if (rel > maxrel) maxrel = rel;
            }
    printf("elements %ld: == fma chain k=0..3: %ld, k=3..0: %ld, pairwise: %ld, unfused chain: %ld, products first then +c: %ld; max rel diff to fwd chain %.3e\n",
           n, m_fwd, m_rev, m_pair, m_unfused, m_sumfirst, maxrel);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
