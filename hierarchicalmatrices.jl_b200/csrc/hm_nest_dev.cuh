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
