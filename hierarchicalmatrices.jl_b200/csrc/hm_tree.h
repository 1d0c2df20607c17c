// hm_tree.h -- host side of on-device KernelMatrix assembly: the tree of index
// ranges and interpolation boxes (no matrix entries are computed on the host).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "hm_types.h"

// BLOCKRANK(Float64), BLOCKSIZE(Float64) -- /root/reference/src/HierarchicalMatrices.jl:5-7
int hm_blockrank_double();
int hm_blocksize_double();

// chebyshevpoints(Float64, n) and chebyshevbarycentricweights(Float64, n), first kind
// -- /root/reference/src/BarycentricMatrix.jl:92-136
void hm_cheb_nodes_weights(int n, double *nodes, double *weights);

// KernelMatrix(f, x, y, a, b, c, d) -- /root/reference/src/KernelMatrix.jl:47-116.
// Produces the leaves in walk order with the absolute offsets the reference's
// mul! walk (KernelMatrix.jl:24-41) would pass to each leaf.  Returns "" or an
// error message (where the reference would throw).
std::string hm_kernel_tree(const double *x, int64_t nx, const double *y, int64_t ny, double a,
                           double b, double c, double d, std::vector<HmLeaf> &leaves,
                           int64_t &nrows, int64_t &ncols);
