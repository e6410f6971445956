// Microbenchmark (measurement tool): latency and issue interval of DFMA and of mma.m8n8k4.f64 on one SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
__global__ void dmma_chain(double* out, long long* cyc, int iters) {
    double c[2 * CHAINS];
    for (int i = 0; i < 2 * CHAINS; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0000001, b = 0.9999999;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[2 * j]), "+d"(c[2 * j + 1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 2 * CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F>
static void run(const char* name, F kern, int threads, int chains, int iters, double* out, long long* cyc) {
    kern<<<1, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    kern<<<1, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-6s threads/CTA %4d (warps per sub-partition %d) chains %2d: %.2f cycles per instruction per warp, %.2f cycles per dependent step\n",
           name, threads, (threads / 32 + 3) / 4, chains, (double)h / ((double)iters * chains), (double)h / iters);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int it = 20000;
    for (int threads : {32, 128, 256, 512}) {
        run("DFMA", dfma_chain<1>, threads, 1, it, out, cyc);
        run("DFMA", dfma_chain<2>, threads, 2, it, out, cyc);
        run("DFMA", dfma_chain<4>, threads, 4, it, out, cyc);
        run("DFMA", dfma_chain<8>, threads, 8, it, out, cyc);
        run("DFMA", dfma_chain<16>, threads, 16, it, out, cyc);
    }
    for (int threads : {32, 128, 256}) {
        run("DMMA", dmma_chain<1>, threads, 1, it, out, cyc);
        run("DMMA", dmma_chain<2>, threads, 2, it, out, cyc);
        run("DMMA", dmma_chain<4>, threads, 4, it, out, cyc);
        run("DMMA", dmma_chain<8>, threads, 8, it, out, cyc);
        run("DMMA", dmma_chain<13>, threads, 13, it, out, cyc);
    }
    return 0;
}
