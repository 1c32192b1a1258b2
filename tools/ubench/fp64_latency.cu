// Dev microbenchmark: dependent-issue latency of the operations on the ray ODE's critical path on B200
// (one warp, clock64 around a dependent chain).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, int n, double a, double b, const double* tab, int mask)
{
    double x = a + threadIdx.x * 1e-9;
    float xf = (float)x;
    int idx = threadIdx.x & mask;
    __shared__ double sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
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
            else if (OP == 6) xf = __fmaf_rn(xf, (float)b, (float)a);
            else if (OP == 7) x = (double)(float)x;                       // F2F down + up
            else if (OP == 8) x = (double)__double2int_rz(x) * 1e-9 + a;  // F2I + I2F + DFMA(as mul+add)
            else if (OP == 9) { idx = (int)__ldg(tab + idx) ; x += idx; } // dependent LDG (L1 hit) + I2F..
            else if (OP == 10) { idx = (int)sm[idx]; x += idx; }           // dependent LDS
            else if (OP == 11) x = (x > a) ? x * b : x + a;               // DSETP + select chain
            else if (OP == 12) xf = 1.0f / sqrtf(xf) + (float)a;         // IEEE sqrtf + div
            else if (OP == 13) xf = rsqrtf(xf) + (float)a;               // MUFU.RSQ
            else if (OP == 14) x = __hiloint2double(__double2hiint(x) + 1, __double2loint(x)); // int op on hi word
            else if (OP == 15) x = fabs(x) * b;                           // DMUL with |x| modifier
            else if (OP == 16) x = fmax(x, a) * b;                        // DMNMX?
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + xf + idx;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, int lanes)
{
    double* out; long long* cyc; double* tab;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&tab, 256 * 8);
    double h[256]; for (int i = 0; i < 256; ++i) h[i] = (double)((i * 7 + 3) & 255);
    cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
    int n = 2000;
    chain<OP><<<1, lanes>>>(out, cyc, n, 1.0000001, 0.9999999, tab, 255);
    chain<OP><<<1, lanes>>>(out, cyc, n, 1.0000001, 0.9999999, tab, 255);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s lanes=%2d : %.1f cycles per dependent op\n", name, lanes, (double)c / (n * 16.0));
    cudaFree(out); cudaFree(cyc); cudaFree(tab);
}
int main()
{
    for (int lanes : {32, 8}) {
        run<0>("DFMA", lanes); run<1>("DMUL", lanes); run<2>("DADD", lanes);
        run<3>("1.0/x + a (IEEE rcp + DADD)", lanes); run<4>("sqrt(x) + a", lanes); run<5>("x / b (IEEE div)", lanes);
        run<6>("FFMA", lanes); run<7>("F2F f64->f32->f64", lanes); run<8>("F2I.f64 + I2F.f64 + DMUL + DADD", lanes);
        run<9>("LDG(L1 hit) + F2I + I2F + DADD", lanes); run<10>("LDS + F2I + I2F + DADD", lanes);
        run<11>("DSETP + 2 alt ops + select", lanes); run<12>("1/sqrtf IEEE + FADD", lanes);
        run<13>("MUFU.RSQ + FADD", lanes); run<14>("IADD on hi word", lanes); run<15>("DMUL |x|", lanes);
        run<16>("fmax + DMUL", lanes);
    }
    return 0;
}
