// Graph index construction on the device: the reference keeps the message-passing graph as
// COO int64 (reference dataset/relgat_dataset.py:123-137); every kernel here consumes the
// stable by-destination ordering of that COO (CSR), its by-source twin (CSC) and a
// by-relation ordering of the CSR slots.  "Stable" = original edge order inside a bucket,
// which fixes the result uniquely (bit-exact vs oracle.graph_index_np).
//
// One-off per graph (not on the per-step path).  The stable LSD radix sort is CUB's (CUDA
// toolkit header library); bucket boundaries come from a binary search over the sorted keys.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace relgat {

__global__ void narrow_keys_kernel(const long long* __restrict__ key64, int* __restrict__ key32,
                                   int* __restrict__ iota, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    key32[i] = static_cast<int>(key64[i]);
    iota[i] = static_cast<int>(i);
  }
}

__global__ void iota_kernel(int* __restrict__ iota, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) iota[i] = static_cast<int>(i);
}

__global__ void gather64_kernel(const long long* __restrict__ a, const long long* __restrict__ b,
                                const int* __restrict__ perm, int* __restrict__ oa, int* __restrict__ ob, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const int p = perm[i];
    oa[i] = static_cast<int>(a[p]);
    ob[i] = static_cast<int>(b[p]);
  }
}

__global__ void gather32_kernel(const int* __restrict__ a, const int* __restrict__ b, const int* __restrict__ perm,
                                int* __restrict__ oa, int* __restrict__ ob, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const int p = perm[i];
    oa[i] = a[p];
    if (b) ob[i] = b[p];
  }
}

// ptr[k] = first position whose sorted key is >= k, k in [0, n_keys]
__global__ void bucket_ptr_kernel(const int* __restrict__ sorted, long long n, int* __restrict__ ptr, long long n_keys) {
  const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k > n_keys) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (sorted[mid] < k) lo = mid + 1; else hi = mid;
  }
  ptr[k] = static_cast<int>(lo);
}

static int bits_for(long long n) {
  int b = 1;
  while (b < 31 && (1ll << b) < n) ++b;
  return b;
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static size_t cub_bytes(long long E) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, static_cast<const int*>(nullptr), static_cast<int*>(nullptr),
                                  static_cast<const int*>(nullptr), static_cast<int*>(nullptr),
                                  static_cast<int>(E), 0, 31, static_cast<cudaStream_t>(0));
  return tmp;
}

}  // namespace relgat

using namespace relgat;

extern "C" long long relgat_graph_index_workspace_bytes(long long E) {
  if (E < 0) return RG_ERR_ARG;
  return static_cast<long long>(3 * align256(static_cast<size_t>(E) * 4 + 4) + align256(cub_bytes(E)) + 256);
}

extern "C" int relgat_graph_index_build(const long long* src, const long long* dst, const long long* rel,
                                        long long E, long long N, long long N_src, long long R,
                                        int* rowptr, int* csr_perm, int* csr_src, int* csr_rel, int* csr_dst,
                                        int* colptr, int* csc_slot, int* csc_dst, int* csc_rel,
                                        int* relptr, int* rel_slot,
                                        void* workspace, long long workspace_bytes, void* stream) {
  if (E < 0 || N <= 0 || N_src <= 0 || R <= 0 || E >= (1ll << 31) || N >= (1ll << 31) || N_src >= (1ll << 31))
    return RG_ERR_ARG;
  if (!rowptr || !colptr || !relptr) return RG_ERR_ARG;
  if (E > 0 && (!src || !dst || !rel || !csr_perm || !csr_src || !csr_rel || !csr_dst || !csc_slot || !csc_dst ||
                !csc_rel || !rel_slot || !workspace))
    return RG_ERR_ARG;
  if (workspace_bytes < relgat_graph_index_workspace_bytes(E)) return RG_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int th = 256;
  auto nb = [&](long long n) { return static_cast<unsigned>((n + th - 1) / th); };
  if (E == 0) {
    cudaMemsetAsync(rowptr, 0, (N + 1) * sizeof(int), s);
    cudaMemsetAsync(colptr, 0, (N_src + 1) * sizeof(int), s);
    cudaMemsetAsync(relptr, 0, (R + 1) * sizeof(int), s);
    return cuda_status(cudaGetLastError());
  }
  char* w = static_cast<char*>(workspace);
  w = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(w)));
  const size_t seg = align256(static_cast<size_t>(E) * 4 + 4);
  int* key_in = reinterpret_cast<int*>(w);
  int* key_out = reinterpret_cast<int*>(w + seg);
  int* val_in = reinterpret_cast<int*>(w + 2 * seg);
  void* tmp = w + 3 * seg;
  size_t tmp_bytes = cub_bytes(E);
  const int n_e = static_cast<int>(E);

  // 1. CSR: stable sort of edge ids by destination
  narrow_keys_kernel<<<nb(E), th, 0, s>>>(dst, key_in, val_in, E);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, csr_dst, val_in, csr_perm, n_e, 0, bits_for(N), s);
  if (e != cudaSuccess) return cuda_status(e);
  gather64_kernel<<<nb(E), th, 0, s>>>(src, rel, csr_perm, csr_src, csr_rel, E);
  bucket_ptr_kernel<<<nb(N + 1), th, 0, s>>>(csr_dst, E, rowptr, N);

  // 2. CSC: stable sort of CSR slots by source
  iota_kernel<<<nb(E), th, 0, s>>>(val_in, E);
  e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, csr_src, key_out, val_in, csc_slot, n_e, 0, bits_for(N_src), s);
  if (e != cudaSuccess) return cuda_status(e);
  gather32_kernel<<<nb(E), th, 0, s>>>(csr_dst, csr_rel, csc_slot, csc_dst, csc_rel, E);
  bucket_ptr_kernel<<<nb(N_src + 1), th, 0, s>>>(key_out, E, colptr, N_src);

  // 3. by relation: stable sort of CSR slots by relation id
  e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, csr_rel, key_out, val_in, rel_slot, n_e, 0, bits_for(R), s);
  if (e != cudaSuccess) return cuda_status(e);
  bucket_ptr_kernel<<<nb(R + 1), th, 0, s>>>(key_out, E, relptr, R);
  return cuda_status(cudaGetLastError());
}
