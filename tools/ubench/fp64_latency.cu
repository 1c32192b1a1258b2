// Dev microbenchmark: dependent-issue latency of FP64 ops on B200 (one warp, clock64 around a chain).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, int n, double a, double b)
{
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __fma_rn(x, b, a);
            else if (OP == 1) x = __dmul_rn(x, b);
            else if (OP == 2) x = __dadd_rn(x, a);
            else if (OP == 3) x = 1.0 / x + a;      // IEEE reciprocal (+ add to keep it finite)
            else if (OP == 4) x = sqrt(x) + a;
            else if (OP == 5) x = x / b;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, int lanes)
{
    double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    int n = 2000;
    chain<OP><<<1, lanes>>>(out, cyc, n, 1.0000001, 0.9999999);
    chain<OP><<<1, lanes>>>(out, cyc, n, 1.0000001, 0.9999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s lanes=%2d : %.1f cycles per dependent op\n", name, lanes, (double)c / (n * 16.0));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int lanes : {32, 8}) {
        run<0>("DFMA", lanes); run<1>("DMUL", lanes); run<2>("DADD", lanes);
        run<3>("1.0/x + a (IEEE rcp + DADD)", lanes); run<4>("sqrt(x) + a", lanes); run<5>("x / b (IEEE div)", lanes);
    }
    return 0;
}
