// Micro-benchmark: dependent DADD / DMUL chain latency and independent-chain throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double *out, int iters, long long *cyc)
{
    double a[CHAINS];
    for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x * 1e-3 + c;
    const double b = out[0];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) a[c] = __dadd_rn(a[c], b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += a[c];
    out[1 + threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int threads, double *d, long long *dc)
{
    const int iters = 4096;
    k<CHAINS><<<1, threads>>>(d, iters, dc);
    cudaDeviceSynchronize();
    k<CHAINS><<<1, threads>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("chains/thread %d threads %4d : %.2f cycles per DADD step (per chain-step %.2f)\n", CHAINS, threads,
           (double)c / iters, (double)c / iters / CHAINS);
}
int main()
{
    double *d; long long *dc;
    cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20); cudaMalloc(&dc, 8);
    run<1>(32, d, dc); run<2>(32, d, dc); run<4>(32, d, dc); run<8>(32, d, dc);
    run<1>(128, d, dc); run<1>(256, d, dc); run<1>(512, d, dc); run<1>(1024, d, dc);
    run<4>(512, d, dc); run<6>(256, d, dc);
    return 0;
}
