// Raw FP64 throughput of one B200: DFMA (scalar pipe) vs DMMA.8x8x4 (mma.sync m8n8k4 f64),
// register-resident operands, ILP-many independent accumulators per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dmma(double *out, int iters)
{
    double d0[ILP], d1[ILP];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < ILP; i++) d0[i] = d1[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d0[i]), "+d"(d1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += d0[i] + d1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters)
{
    double d[ILP];
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < ILP; i++) d[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) d[i] = fma(d[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 32 * 1024 * sizeof(double));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = 148 * (warps * 32 / threads);
        float ms = timeit([&] { k_dmma<16><<<blocks, threads>>>(out, iters); });
        double fl = (double)blocks * (threads / 32) * iters * 16 * 512.0;
        printf("DMMA.8x8x4 ILP16  warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        ms = timeit([&] { k_dmma<4><<<blocks, threads>>>(out, iters); });
        fl = (double)blocks * (threads / 32) * iters * 4 * 512.0;
        printf("DMMA.8x8x4 ILP4   warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
        fl = (double)blocks * threads * iters * 16 * 2.0;
        printf("DFMA       ILP16  warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    return 0;
}
