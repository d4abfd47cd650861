// by-source backward pass, instantiation for <__nv_bfloat16, 8> (see edge_bwd_src.cuh)
#include "edge_bwd_src.cuh"

namespace relgat {
template int run_src<__nv_bfloat16, 8>(const void*, long long, const void*, const float*, const float*, const float*,
                          const float*, const int*, const int*, const int*, const int*, const int4*, int, const int2*,
                          const int*, const int*, int, float*, float*, void*, void*, float*, const uint32_t*, float, const uint32_t*, const int*, int, int,
                          long long, int, int, int, int, int*, cudaStream_t);
}  // namespace relgat
