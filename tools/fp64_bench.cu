// Microbenchmark: per-SM issue rates of the float64 operations the integrate kernel leans on.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, double a, double b, int iters)
{
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#define STEP(x)                                                                         \
    if (OP == 0) x = __fma_rn(x, b, a);                                                 \
    else if (OP == 1) x = __dadd_rn(x, b);                                              \
    else if (OP == 2) x = __dmul_rn(x, b);                                              \
    else if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; } \
    else if (OP == 4) x = __ddiv_rn(a, x);                                              \
    else if (OP == 5) x = rint(x * b);                                                  \
    else if (OP == 6) { x = (x > b) ? x - 1.0 : x + 1.0; }
        STEP(x0) STEP(x1) STEP(x2) STEP(x3) STEP(x4) STEP(x5) STEP(x6) STEP(x7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int OP> __global__ void kf(float *out, float a, float b, int iters)
{
    float x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#define STEPF(x) if (OP == 0) x = __fmaf_rn(x, b, a); else x = __fdiv_rn(a, x);
        STEPF(x0) STEPF(x1) STEPF(x2) STEPF(x3) STEPF(x4) STEPF(x5) STEPF(x6) STEPF(x7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// max relative error of the rcp.approx.ftz.f64 seed, scanned over 2^28 mantissas x a few exponents
__global__ void k_rcp_err(double *out)
{
    double worst = 0.0;
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    for (int rep = 0; rep < 256; ++rep) {
        const unsigned long long m = (tid * 256ull + rep) * 0x9E3779B97F4A7C15ull;
        const int ex = 1023 - 40 + (int)(m % 80);
        const double z = __longlong_as_double(((long long)ex << 52) | (long long)(m >> 12));
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
        const double err = fabs(__fma_rn(-z, r, 1.0));
        worst = fmax(worst, err);
    }
    for (int o = 16; o; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)out, (unsigned long long)__double_as_longlong(worst));
}
template <typename F> float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d kHz\n", p.name, sms, clk);
    double *o; cudaMalloc(&o, sizeof(double) * sms * 8 * 1024);
    const int iters = 4096, blocks = sms * 8, threads = 512;
    const char *names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H", "DDIV(ieee)", "DMUL+rint", "DSETP+DADD(sel)"};
    float ms;
#define RUN(OP) ms = timeit([&] { k<OP><<<blocks, threads>>>(o, 1.0000001, 0.9999999, iters); }); \
    printf("%-16s %8.3f ms  -> %7.2f ops/clk/SM (at max clock)\n", names[OP], ms, (double)blocks * threads * iters * 8 / (ms * 1e-3) / ((double)clk * 1e3) / sms);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    ms = timeit([&] { kf<0><<<blocks, threads>>>((float *)o, 1.0000001f, 0.9999999f, iters); });
    printf("%-16s %8.3f ms  -> %7.2f ops/clk/SM\n", "FFMA", ms, (double)blocks * threads * iters * 8 / (ms * 1e-3) / ((double)clk * 1e3) / sms);
    ms = timeit([&] { kf<1><<<blocks, threads>>>((float *)o, 1.0000001f, 0.9999999f, iters); });
    printf("%-16s %8.3f ms  -> %7.2f ops/clk/SM\n", "FDIV(ieee)", ms, (double)blocks * threads * iters * 8 / (ms * 1e-3) / ((double)clk * 1e3) / sms);
    cudaMemset(o, 0, 8);
    k_rcp_err<<<4096, 256>>>(o);
    double w; cudaMemcpy(&w, o, 8, cudaMemcpyDeviceToHost);
    printf("rcp.approx.ftz.f64 seed: max |1 - z*r| over 2.7e8 samples = %.3e = 2^%.2f\n", w, log2(w));
    return 0;
}
