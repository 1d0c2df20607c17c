// hm_nest_dev.cuh -- device helper shared by the subtree kernels of hm_nest.cu and hm_nest_panel.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "hm_nest.h"

// The schedule of a subtree (box ids depth by depth and what the pass needs of each box) is staged in
// shared memory once per CTA: without it every depth costs three dependent global round trips (group
// bounds -> box id -> node record) before the first useful load, ~5 us per depth, which is all the time
// the upper tiers take.  Subtrees beyond the caps read the tables from global memory as before.
constexpr int HM_SUBCAP = 192, HM_GCAP = 48;
struct HmSubSched {
    int id[HM_SUBCAP];
    int aux[HM_SUBCAP]; // up: first half (child0); down: parent * 2 + which (no parent: -1)
    int fin[HM_SUBCAP]; // down, panel form: index among the finest boxes, or -1
    int grp[HM_GCAP + 1];
    int cached;
};

// every thread of the CTA calls it; the caller's __syncthreads() publishes it
template <bool DOWN>
__device__ __forceinline__ void hm_stage_schedule(HmSubSched &S, const HmNestNode *__restrict__ nodes,
                                                  const int32_t *__restrict__ order, const int32_t *__restrict__ grp,
                                                  const int32_t *__restrict__ fin, int g0, int g1)
{
    const int eb = grp[g0], ne = grp[g1] - eb, ng = g1 - g0;
    const bool ok = ne <= HM_SUBCAP && ng <= HM_GCAP;
    if (ok) {
        for (int i = threadIdx.x; i < ne; i += blockDim.x) {
            const int id = order[eb + i];
            S.id[i] = id;
            if (DOWN) {
                const int par = nodes[id].parent;
                S.aux[i] = par >= 0 ? par * 2 + nodes[id].which : -1;
                S.fin[i] = fin ? fin[id] : -1;
            } else {
                S.aux[i] = nodes[id].child0;
            }
        }
        for (int i = threadIdx.x; i <= ng; i += blockDim.x) S.grp[i] = grp[g0 + i];
    }
    if (threadIdx.x == 0) S.cached = ok;
}

// Programmatic dependent launch (as in hm_kernels.cu): a kernel launched with the attribute may start while
// its predecessor in the stream still runs; hm_pdl_wait() blocks until the predecessor has completed and its
// writes are visible.  The tree passes are chains of short dependent launches: each kernel stages its
// schedule and maps (plan-time tables nobody writes) before it waits, under the tail of the previous one.
__device__ __forceinline__ void hm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void hm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#include <cstdlib>
inline bool hm_pdl_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HMB200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// only kernels that call hm_pdl_wait() before they touch what the predecessor wrote (and before their own
// first global write) may be launched through this
template <class... KArgs, class... Args>
inline cudaError_t hm_launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = hm_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
