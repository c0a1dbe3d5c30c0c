// H2D bandwidth of the copies the upload path uses: linear from cudaMallocHost memory, linear and strided (2D) from
// cudaHostRegister'ed malloc memory, one or two streams.   nvcc -O2 -arch=sm_100a h2d2d.cu -o h2d2d
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t rows = 131070, pitch = 3200, width = 1536, lin = rows * width;
    char *hp, *hm = (char *) malloc(rows * pitch + 4096), *d, *d2;
    CK(cudaMallocHost(&hp, lin)); memset(hp, 1, lin); memset(hm, 2, rows * pitch);
    CK(cudaMalloc(&d, rows * pitch)); CK(cudaMalloc(&d2, lin));
    double t0 = now(); CK(cudaHostRegister(hm, rows * pitch, cudaHostRegisterDefault)); printf("cudaHostRegister of %.0f MB: %.1f ms\n", rows * pitch / 1e6, now() - t0);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *what, double bytes) { float ms; cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("%-64s %7.3f ms  %6.1f GB/s\n", what, ms, bytes / ms / 1e6); };
    for (int rep = 0; rep < 2; rep++)
    {
        cudaEventRecord(e0, s1); CK(cudaMemcpyAsync(d, hp, lin, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("linear, cudaMallocHost, 201 MB", lin);
        cudaEventRecord(e0, s1); CK(cudaMemcpyAsync(d, hm, lin, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("linear, registered malloc, 201 MB", lin);
        cudaEventRecord(e0, s1); CK(cudaMemcpyAsync(d, hm, rows * pitch, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("linear, registered malloc, whole 419 MB", rows * pitch);
        cudaEventRecord(e0, s1); CK(cudaMemcpy2DAsync(d, width, hm, pitch, width, rows, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("2D 1536 of 3200 B x 131070 rows, registered", lin);
        cudaEventRecord(e0, s1); CK(cudaMemcpy2DAsync(d, 1152, hm, pitch, 1152, rows, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("2D 1152 of 3200 B", rows * 1152.0);
        cudaEventRecord(e0, s1); CK(cudaMemcpy2DAsync(d, 384, hm, pitch, 384, rows, cudaMemcpyHostToDevice, s1)); cudaEventRecord(e1, s1); report("2D 384 of 3200 B", rows * 384.0);
        // two halves on two streams
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
        CK(cudaMemcpy2DAsync(d, width, hm, pitch, width, rows / 2, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpy2DAsync(d + (rows / 2) * width, width, hm + (rows / 2) * pitch, pitch, width, rows - rows / 2, cudaMemcpyHostToDevice, s2));
        cudaEvent_t e2; cudaEventCreate(&e2); cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); report("2D 1536 of 3200 B, two halves on two streams", lin);
        // 2D on s2 concurrently with linear pinned on s1 (the vectors)
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
        CK(cudaMemcpyAsync(d2, hp, lin / 2, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpy2DAsync(d, width, hm, pitch, width, rows, cudaMemcpyHostToDevice, s2));
        cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); report("2D 201 MB on one stream + linear 100 MB on another", lin * 1.5);
        cudaEventRecord(e0, s1); CK(cudaMemcpyAsync(hp, d, 42e6, cudaMemcpyDeviceToHost, s1)); cudaEventRecord(e1, s1); report("D2H linear 42 MB", 42e6);
    }
    return 0;
}
