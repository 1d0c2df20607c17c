// stream_read.cu -- what a read-only stream can reach on this GPU: the ceiling of hm_stream_kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_read stream_read.cu && ./stream_read
// Variants: CTAs of 256 threads, U independent 16-byte loads in flight per thread, persistent grid of
// 148 * C CTAs (grid-stride) or one CTA per 256 * U * 16 * K bytes chunk (like one slab per CTA).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ double2 ld_cs(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_nc(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int U, int MODE>
__global__ void __launch_bounds__(256) read_kernel(const double2 *__restrict__ a, size_t n2, double *out)
{
    double s = 0.0;
    const size_t stride = (size_t)gridDim.x * 256 * U;
    for (size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x; i + 255 * 0 < n2; i += stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t j = i + (size_t)u * 256;
            v[u] = j < n2 ? (MODE == 0 ? ld_cs(a + j) : MODE == 1 ? ld_nc(a + j) : a[j]) : make_double2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) s += v[u].x + v[u].y;
    }
    if (s == 1.2345e-300) out[0] = s;
}

template <int U, int MODE>
float run(const double2 *a, size_t n2, double *out, int grid)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) read_kernel<U, MODE><<<grid, 256>>>(a, n2, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) read_kernel<U, MODE><<<grid, 256>>>(a, n2, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    const size_t bytes = (size_t)12 << 30;
    const size_t n2 = bytes / 16;
    double2 *a;
    double *out;
    cudaMalloc(&a, bytes);
    cudaMalloc(&out, 8);
    cudaMemset(a, 0, bytes);
    const char *mn[3] = {"ld.cs", "ld.nc.no_allocate", "plain"};
    for (int mode = 0; mode < 3; mode++)
        for (int c : {2, 4, 6, 8}) {
            for (int big = 0; big < 2; big++) {
                const int grid = big ? (int)((n2 + 256 * 8 * 64 - 1) / (256 * 8 * 64)) : 148 * c; // big: ~2 MB per CTA, many CTAs
                float m4 = mode == 0 ? run<4, 0>(a, n2, out, grid) : mode == 1 ? run<4, 1>(a, n2, out, grid) : run<4, 2>(a, n2, out, grid);
                float m8 = mode == 0 ? run<8, 0>(a, n2, out, grid) : mode == 1 ? run<8, 1>(a, n2, out, grid) : run<8, 2>(a, n2, out, grid);
                float m16 = mode == 0 ? run<16, 0>(a, n2, out, grid) : mode == 1 ? run<16, 1>(a, n2, out, grid) : run<16, 2>(a, n2, out, grid);
                printf("%-18s grid=%6d (%s)  U=4: %7.1f GB/s  U=8: %7.1f GB/s  U=16: %7.1f GB/s\n", mn[mode], grid,
                       big ? "2 MB per CTA" : "persistent", bytes / m4 / 1e6, bytes / m8 / 1e6, bytes / m16 / 1e6);
                if (big) break;
            }
            if (c == 8) {
                const int grid = (int)((n2 + 256 * 8 * 64 - 1) / (256 * 8 * 64));
                float m8 = mode == 0 ? run<8, 0>(a, n2, out, grid) : mode == 1 ? run<8, 1>(a, n2, out, grid) : run<8, 2>(a, n2, out, grid);
                printf("%-18s grid=%6d (one CTA per 2 MB)  U=8: %7.1f GB/s\n", mn[mode], grid, bytes / m8 / 1e6);
            }
        }
    return 0;
}
