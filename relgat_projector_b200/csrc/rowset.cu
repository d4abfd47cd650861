// Row sets as bitmaps -> compact numbering.
//
// The backward of a training step only has non-zero gradient rows inside the batch's in-neighbourhood (mask.cu builds
// the bitmaps: relgat_mark_rows / relgat_mark_sources).  This turns a bitmap into what the compacted backward needs:
//   rank[i]  = position of row i among the marked rows (ascending), -1 for unmarked rows   (where a kernel writes row i)
//   list[k]  = k-th marked row                                                              (what a gather reads)
//   count    = number of marked rows                                                        (GEMM sizes; read by the host)
// Three launches: popcount per word, exclusive scan of the word counts (CUB), expand.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace relgat {

__global__ void word_popc_kernel(const uint32_t* __restrict__ bits, int n_words, int n_rows, int* __restrict__ cnt) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t v = __ldg(bits + w);
  const int rem = n_rows - w * 32;
  if (rem < 32) v &= (1u << rem) - 1u;  // bits beyond the last row do not count
  cnt[w] = __popc(v);
}

__global__ void rank_expand_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ cnt,
                                   const int* __restrict__ before, int n_words, int n_rows, int* __restrict__ rank,
                                   long long* __restrict__ list, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int w = i >> 5, b = i & 31;
  const uint32_t v = __ldg(bits + w);
  int r = -1;
  if ((v >> b) & 1u) {
    r = __ldg(before + w) + __popc(v & ((1u << b) - 1u));
    if (list) list[r] = i;
  }
  if (rank) rank[i] = r;
  if (i == n_rows - 1) *count = __ldg(before + n_words - 1) + __ldg(cnt + n_words - 1);
}

static size_t align256r(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static size_t rank_scan_bytes(int n_words) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, static_cast<const int*>(nullptr), static_cast<int*>(nullptr), n_words,
                                static_cast<cudaStream_t>(0));
  return tmp;
}

}  // namespace relgat

using namespace relgat;

extern "C" long long relgat_bitmap_ranks_workspace_bytes(long long n_rows) {
  if (n_rows <= 0) return 256;
  const int n_words = static_cast<int>((n_rows + 31) / 32);
  return static_cast<long long>(align256r(rank_scan_bytes(n_words)) + 2 * align256r(sizeof(int) * static_cast<size_t>(n_words)));
}

extern "C" int relgat_bitmap_ranks(const unsigned int* bits, long long n_rows, int* rank, long long* list, int* count,
                                   void* workspace, long long workspace_bytes, void* stream) {
  if (n_rows < 0 || n_rows >= (1ll << 31) || !count) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_rows == 0) return cuda_status(cudaMemsetAsync(count, 0, sizeof(int), s));
  if (!bits || !workspace) return RG_ERR_ARG;
  if (workspace_bytes < relgat_bitmap_ranks_workspace_bytes(n_rows)) return RG_ERR_WORKSPACE;
  const int n = static_cast<int>(n_rows), n_words = (n + 31) / 32;
  char* w = static_cast<char*>(workspace);
  size_t tmp_bytes = rank_scan_bytes(n_words);
  void* tmp = w;
  w += align256r(tmp_bytes);
  int* cnt = reinterpret_cast<int*>(w);
  w += align256r(sizeof(int) * static_cast<size_t>(n_words));
  int* before = reinterpret_cast<int*>(w);
  word_popc_kernel<<<(n_words + 255) / 256, 256, 0, s>>>(bits, n_words, n, cnt);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, before, n_words, s);
  if (e != cudaSuccess) return cuda_status(e);
  rank_expand_kernel<<<(n + 255) / 256, 256, 0, s>>>(bits, cnt, before, n_words, n, rank, list, count);
  return cuda_status(cudaGetLastError());
}
