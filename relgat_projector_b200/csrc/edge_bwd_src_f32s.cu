// by-source backward pass, instantiation for <float, 1> (see edge_bwd_src.cuh)
#include "edge_bwd_src.cuh"

namespace relgat {
template int run_src<float, 1>(const void*, long long, const void*, const float*, const float*, const float*,
                          const float*, const int*, const int*, const int*, const int*, const int4*, int, const int2*,
                          const int*, const int*, int, float*, float*, void*, void*, float*, const uint32_t*, float, const uint32_t*, const int*, int, int,
                          long long, int, int, int, int, int*, cudaStream_t);
}  // namespace relgat
