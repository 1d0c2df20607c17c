// FP64 tensor-core shapes on B200, register-resident operands, a 4x4-block warp tile as in the
// panel kernels: which mma.sync f64 shape does the pipe run fastest, and does a realistic operand
// pattern (distinct A / B fragments per instruction) reach the same rate as the ILP microbenchmark?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_shapes dmma_shapes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}

// 32 x 32 warp tile, k advanced by 4 per step: 16 m8n8k4 per step with 4 + 4 distinct fragments
__global__ void k884(double *out, int iters)
{
    double acc[4][4][2] = {};
    double a[4], b[4];
    for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-3 + i, b[i] = 1.0 + threadIdx.x * 1e-6 * i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) mma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] += 1e-9, b[i] -= 1e-9; // new fragments every step
    }
    double s = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 32 x 32 warp tile with m16n8k4: 2 x 4 blocks of 16 x 8
__global__ void k1684(double *out, int iters)
{
    double acc[2][4][4] = {};
    double a[2][2], b[4];
    for (int i = 0; i < 2; i++) a[i][0] = threadIdx.x * 1e-3 + i, a[i][1] = a[i][0] + 0.5;
    for (int j = 0; j < 4; j++) b[j] = 1.0 + threadIdx.x * 1e-6 * j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) mma1684(acc[i][j], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 2; i++) a[i][0] += 1e-9, a[i][1] += 1e-9;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] -= 1e-9;
    }
    double s = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) s += acc[i][j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k1688(double *out, int iters)
{
    double acc[2][4][4] = {};
    double a[2][4], b[4][2];
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 4; k++) a[i][k] = threadIdx.x * 1e-3 + i + k;
    for (int j = 0; j < 4; j++) b[j][0] = 1.0 + threadIdx.x * 1e-6 * j, b[j][1] = b[j][0] * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) mma1688(acc[i][j], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 2; i++) a[i][0] += 1e-9, a[i][3] += 1e-9;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j][0] -= 1e-9;
    }
    double s = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) s += acc[i][j][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k16816(double *out, int iters)
{
    double acc[2][4][4] = {};
    double a[2][8], b[4][4];
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 8; k++) a[i][k] = threadIdx.x * 1e-3 + i + k;
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++) b[j][k] = 1.0 + threadIdx.x * 1e-6 * j + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) mma16816(acc[i][j], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 2; i++) a[i][0] += 1e-9, a[i][7] += 1e-9;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j][0] -= 1e-9;
    }
    double s = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) s += acc[i][j][k];
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
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(double));
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        const int threads = warps * 32 > 512 ? 512 : warps * 32;
        const int blocks = 148 * (warps * 32 / threads);
        const double nw = (double)blocks * (threads / 32);
        float ms = timeit([&] { k884<<<blocks, threads>>>(out, iters); });
        printf("warps/SM=%2d  m8n8k4   4x4 tile: %6.2f TFLOP/s\n", warps, nw * iters * 16 * 512.0 / ms / 1e9);
        ms = timeit([&] { k1684<<<blocks, threads>>>(out, iters); });
        printf("warps/SM=%2d  m16n8k4  2x4 tile: %6.2f TFLOP/s\n", warps, nw * iters * 8 * 1024.0 / ms / 1e9);
        ms = timeit([&] { k1688<<<blocks, threads>>>(out, iters); });
        printf("warps/SM=%2d  m16n8k8  2x4 tile: %6.2f TFLOP/s\n", warps, nw * iters * 8 * 2048.0 / ms / 1e9);
        ms = timeit([&] { k16816<<<blocks, threads>>>(out, iters); });
        printf("warps/SM=%2d  m16n8k16 2x4 tile: %6.2f TFLOP/s\n", warps, nw * iters * 8 * 4096.0 / ms / 1e9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
