// Dev microbenchmark: can ONE warp keep several independent FP64 operations in flight on B200?
// (cycles per DFMA for 1, 2, 4, 8 independent chains in one warp; and for 1..4 warps on one scheduler)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chains(double* out, long long* cyc, int n, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = __fma_rn(x[k], b, a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP> void run(int threads)
{
    double* out; long long* cyc; cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64 * 8);
    int n = 2000;
    chains<ILP><<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
    chains<ILP><<<1, threads>>>(out, cyc, n, 1.0000001, 0.9999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads=%4d (warps/scheduler %.2f)  ILP=%d : %.2f cycles per DFMA per warp, %.2f cycles per warp-DFMA issued on the scheduler\n",
           threads, threads / 128.0, ILP, (double)c / (n * 8.0 * ILP), (double)c / (n * 8.0 * ILP) / (threads > 128 ? threads / 128.0 : 1.0));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {32, 128, 256, 512, 1024}) { run<1>(threads); run<2>(threads); run<4>(threads); run<8>(threads); }
    return 0;
}
