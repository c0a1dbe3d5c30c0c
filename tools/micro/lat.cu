// Microbenchmark: dependent-chain latencies (cycles, one warp) and per-SM throughputs of the
// operations the TDUNES block kernels are made of, on one GPU.  Answers "what does one Cholesky
// step cost" before a kernel is redesigned.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define N 512

__device__ __forceinline__ void rsqrt_core(double c, double &root, double &inv)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(c));
    double h = 0.5 * y, g = c * y, r;
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    g = fma(fma(-g, g, c), h, g);
    const double yy = h + h;
    inv = fma(yy, fma(-g, yy, 1.0), yy);
    root = g;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode selects the chain; out[0] = cycles per operation * 1000
__global__ void k_lat(int mode, long long *cyc, double *sink, const int *chase, double seed)
{
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double x = seed + 1e-6 * lane, y = 0.999, z = 1e-3;
    double c0 = x, c1 = y;
    int idx = lane;
    long long t0 = clock64();
    switch (mode)
    {
    case 0: for (int i = 0; i < N; i++) x = fma(x, y, z); break;                         // DFMA
    case 1: for (int i = 0; i < N; i++) x = x + y; break;                                // DADD
    case 2: for (int i = 0; i < N; i++) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); break; // MUFU.RSQ64H
    case 3: for (int i = 0; i < N; i++) { double r, v; rsqrt_core(x, r, v); x = r + v; } break; // pivot core (+1 DADD)
    case 4: for (int i = 0; i < N; i++) x = __shfl_sync(FULL, x, (lane + 1) & 31); break; // SHFL.64
    case 5: for (int i = 0; i < N; i++) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); } break; // STS -> LDS
    case 6: for (int i = 0; i < N; i++) { const double2 t = *reinterpret_cast<const double2 *>(sm + 2 * (idx & 255)); idx = (int) (t.x) + idx + 1; } x = idx; break; // LDS.128 broadcast-ish, address dependent
    case 7: for (int i = 0; i < N; i++) dmma(c0, c1, y, z); x = c0 + c1; break;          // DMMA, accumulator dependent
    case 8: for (int i = 0; i < N; i++) { dmma(c0, c1, y, z); y = c0; } x = c0 + c1; break; // DMMA -> A operand dependent
    case 9: for (int i = 0; i < N; i++) idx = chase[idx]; x = idx; break;                // global load, L2/L1 hit pointer chase
    case 10: for (int i = 0; i < N; i++) x = sqrt(x) + 1.0; break;                       // library sqrt
    case 11: for (int i = 0; i < N; i++) x = 1.0 / x + 1e-3; break;                      // library division
    case 12: for (int i = 0; i < N; i++) { int v = (int) x; v = __shfl_sync(FULL, v, (lane + 1) & 31); x = v; } break; // SHFL.32 + 2 cvt
    case 13: for (int i = 0; i < N; i++) x = x * y; break;                               // DMUL
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[mode] = (t1 - t0) * 1000 / N;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x + c1;
}

// throughput kernels: every SM, `warps` warps per SM, ILP 8
__global__ void k_tp(int mode, double *sink, int iters, double seed)
{
    __shared__ double sm[4096];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = seed + i + 1e-6 * threadIdx.x;
    double *mine = sm + (threadIdx.x >> 5) * 256;
    for (int it = 0; it < iters; it++)
    {
        switch (mode)
        {
        case 0:
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], 0.999, 1e-3);
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < 4; i++) dmma(a[2 * i], a[2 * i + 1], 0.999, 1e-3);
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __shfl_sync(FULL, a[i], (lane + 1 + i) & 31);
            break;
        case 3: // uniform 16-byte broadcast loads (column broadcast of the right-looking Cholesky)
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const double2 t = *reinterpret_cast<const double2 *>(mine + 2 * ((it + i) & 63));
                a[2 * i] += t.x; a[2 * i + 1] += t.y;
            }
            break;
        case 4: // fragment loads: lane T reads [T/4][T%4] of a row-major tile with leading dimension 38
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] += mine[(lane >> 2) * 26 + (lane & 3) + ((it + i) & 7)];
            break;
        case 5: // uniform 8-byte broadcast loads
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] += mine[(it + i) & 127];
            break;
        case 6: // per-lane 8-byte stores, unit stride
#pragma unroll
            for (int i = 0; i < 8; i++) mine[lane + 32 * (i & 3)] = a[i];
            break;
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    long long *cyc; double *sink; int *chase;
    cudaMallocManaged(&cyc, 64 * sizeof(long long));
    cudaMalloc(&sink, sizeof(double) * 148 * 64 * 1024);
    const int nchase = 1 << 16;
    cudaMallocManaged(&chase, nchase * sizeof(int));
    for (int i = 0; i < nchase; i++) chase[i] = (i + 4099) & (nchase - 1);
    const char *names[] = {"DFMA", "DADD", "MUFU.RSQ64H", "pivot core (rsqrt + 12 FP64) + DADD", "SHFL 64-bit", "STS+syncwarp+LDS+syncwarp",
                           "LDS.128 address-dependent", "DMMA acc-dependent", "DMMA acc->A dependent", "LDG chase (256 KB set, L2)",
                           "sqrt() + DADD", "1.0/x + DADD", "SHFL 32-bit + 2 cvt", "DMUL"};
    int dev = 0, clk = 0, sms = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    printf("device: %d SMs, clock %d kHz\n", sms, clk);
    for (int rep = 0; rep < 2; rep++)
        for (int m = 0; m < 14; m++)
        {
            k_lat<<<1, 32>>>(m, cyc, sink, chase, 1.5);
            cudaDeviceSynchronize();
            if (rep) printf("latency  %-40s %8.1f cycles/op\n", names[m], cyc[m] / 1000.0);
        }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *tn[] = {"DFMA (warp instr)", "DMMA m8n8k4", "SHFL 64-bit", "LDS.128 uniform broadcast", "LDS.64 fragment pattern", "LDS.64 uniform broadcast", "STS.64 unit stride"};
    const double per_iter[] = {8, 4, 8, 4, 8, 8, 8};
    for (int wps = 4; wps <= 16; wps *= 2)
        for (int m = 0; m < 7; m++)
        {
            const int iters = 4000;
            float ms = 0;
            k_tp<<<sms, 32 * wps>>>(m, sink, iters, 1.5);
            cudaEventRecord(e0); k_tp<<<sms, 32 * wps>>>(m, sink, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = per_iter[m] * iters * wps;        // warp instructions per SM
            const double cycles = ms * 1e-3 * clk * 1e3;
            printf("throughput %2d warps/SM  %-28s %7.3f warp-instr/cycle/SM (%.1f cycles per instr)\n", wps, tn[m], instr / cycles, cycles / instr);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
