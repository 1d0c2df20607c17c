// hm_device.cuh -- small device functions shared by the matrix-free kernels (hm_kernels.cu,
// hm_free_panel.cu).
#pragma once
#include <cuda_runtime.h>

// 1/d to about one ulp: the hardware's 2^-23 approximation r0 and one cubic (Halley) step,
// 1/d = r0 (1 + e + e^2 + ...), e = 1 - d r0, truncated after e^2 (error e^3 ~ 2^-69): 3 DFMA,
// against the ~3x longer correctly rounded __drcp_rn.  The matrix-free kernels spend their time here.
__device__ __forceinline__ double frcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// the four kernels of hm_assemble_kernel (cauchy, coulomb, coulomb', log; examples/Kernel.jl:34-37)
__device__ __forceinline__ double kernel_eval_fast(int id, double x, double y)
{
    const double d = __dsub_rn(x, y);
    switch (id) {
    case 0: return frcp(d);
    case 1: return frcp(__dmul_rn(d, d));
    case 2: return frcp(__dmul_rn(__dmul_rn(d, d), d));
    default: return log(fabs(d));
    }
}
