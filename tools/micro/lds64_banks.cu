// Micro-benchmark: cycles per LDS.64 warp instruction for different lane -> address patterns (which
// 16 lanes form a conflict group?).  One warp, dependent-free loads, clock64 around the loop.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const int *pat, int npat, long long *cyc, double *sink)
{
    __shared__ double s[4096];
    for (int i = threadIdx.x; i < 4096; i += 32) s[i] = i;
    __syncwarp();
    for (int p = 0; p < npat; p++) {
        const int cell = pat[p * 32 + threadIdx.x];
        const double *a = s + cell;
        double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 256; it++) {
            acc0 += *(volatile const double *)(a + 0 * 16);
            acc1 += *(volatile const double *)(a + 1 * 16);
            acc2 += *(volatile const double *)(a + 2 * 16);
            acc3 += *(volatile const double *)(a + 3 * 16);
            acc0 += *(volatile const double *)(a + 4 * 16);
            acc1 += *(volatile const double *)(a + 5 * 16);
            acc2 += *(volatile const double *)(a + 6 * 16);
            acc3 += *(volatile const double *)(a + 7 * 16);
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[p] = t1 - t0;
        sink[p * 32 + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
    }
}
int main()
{
    const int NP = 8;
    int h[NP][32];
    const char *name[NP] = {"contiguous", "stride 7", "halves overlap (0-15 == 16-31 + 16 cells)", "8+8 rows 8 apart (each half warp)",
                            "8+8 rows 0 apart (conflict in half warp)", "even lanes set A / odd lanes same set (conflict if half = consecutive)",
                            "lanes 0-7,16-23 distinct; 8-15,24-31 same as those", "all same address"};
    for (int l = 0; l < 32; l++) {
        h[0][l] = l;
        h[1][l] = 7 * l;
        h[2][l] = 7 * (l % 16) + (l / 16) * 16;
        h[3][l] = 7 * (l % 8) + ((l / 8) % 2) * (56 + 8 * 0) + (l / 16) * 256 + (((l / 8) % 2) ? 8 - 56 % 16 + 0 : 0);
        h[4][l] = 7 * (l % 8) + ((l / 8) % 2) * 112 + (l / 16) * 256;
        h[5][l] = 7 * (l / 2) + (l % 2) * 16 * 8;
        h[6][l] = 7 * ((l % 8) + 8 * (l / 16)) + ((l / 8) % 2) * 16 * 8;
        h[7][l] = 3;
    }
    // pattern 3: row b = row a + 56 cells (56 = 8 mod 16)
    for (int l = 0; l < 32; l++) h[3][l] = 7 * (l % 8) + ((l / 8) % 2) * 56 + (l / 16) * 256;
    int *d; long long *dc; double *ds;
    cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&dc, NP * 8); cudaMalloc(&ds, NP * 32 * 8);
    k<<<1, 32>>>(d, NP, dc, ds); cudaDeviceSynchronize();
    k<<<1, 32>>>(d, NP, dc, ds); cudaDeviceSynchronize();
    long long c[NP];
    cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    for (int p = 0; p < NP; p++) printf("%-75s %.2f cycles per LDS.64\n", name[p], (double)c[p] / (256 * 8));
    return 0;
}
