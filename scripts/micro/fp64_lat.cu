// FP64 pipe microbenchmark for B200: dependent-issue latency and per-SMSP issue interval of DFMA,
// MUFU.RCP64H latency, LDS.64 latency.  One block of W warps per SM on one SM only.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dep(double* out, int iters, long long* cyc, int chains)
{
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    const double m = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    if (chains == 1) for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); }
    else if (chains == 2) for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); }
    else for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp(double* out, int iters, long long* cyc)
{
    double a = threadIdx.x * 1e-9 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { double s; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(a)); a = s; }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, int iters, long long* cyc)
{
    __shared__ unsigned idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 8 + 8) % 1024 * 1;   // pointer chase in words
    __syncthreads();
    unsigned p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) p = idx[p & 1023];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const int iters = 100000;
    for (int chains : {1, 2, 4})
        for (int warps : {1, 2, 4, 8, 16}) {
            k_dep<<<1, warps * 32>>>(out, iters, cyc, chains); cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("DFMA chains/warp %d warps/SM %2d : %.2f cycles per loop trip (%.2f per DFMA per warp; SM-wide %.3f DFMA warp-instr/cycle)\n",
                   chains, warps, (double)h / iters, (double)h / iters / chains, (double)chains * warps * iters / h);
        }
    k_rcp<<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RCP64H dependent: %.2f cycles\n", (double)h / iters);
    k_lds<<<1, 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.32 pointer chase: %.2f cycles\n", (double)h / iters);
    return 0;
}
