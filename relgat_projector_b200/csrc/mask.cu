// Dropout masks for the fused layer (reference core/model/layer.py:296-297 attention dropout, :321-322 feature
// dropout) and a row-clearing helper for the sparse hand-over of the batch gradient.
//
// A mask is a bit array: element i lives in word i >> 5, bit i & 31; bit = 1 means KEEP.  Feature masks index
// i = row * (32 * words_per_row) + column (rows start on a word boundary), attention masks i = csr_slot * H + head.
// The bits come from Philox4x32-10 keyed by a seed the host draws from torch's generator (so torch.manual_seed
// governs them), or are supplied by the caller (tests inject a mask and replay it in the oracle).
#include "common.cuh"

namespace relgat {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// One thread per 32-bit word: four Philox calls = 32 draws of 16 bits; keep when draw >= threshold.
__global__ void bernoulli_bits_kernel(uint32_t* __restrict__ bits, long long n_words, uint32_t threshold, uint64_t seed) {
  const long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t word = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[4];
    philox4x32_10(seed, static_cast<uint64_t>(w) * 4 + q, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      word |= ((r[i] & 0xffffu) >= threshold ? 1u : 0u) << (q * 8 + i * 2);
      word |= ((r[i] >> 16) >= threshold ? 1u : 0u) << (q * 8 + i * 2 + 1);
    }
  }
  bits[w] = word;
}

__global__ void zero_rows_kernel(float* __restrict__ table, long long ld, const long long* __restrict__ ids, long long n,
                                 int D) {
  const bool vec = (D & 3) == 0 && (ld & 3) == 0 && reinterpret_cast<uintptr_t>(table) % 16 == 0;
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
    const long long id = ids[i];
    if (id < 0) continue;  // entry switched off by the caller
    float* row = table + id * ld;
    if (vec) {
      for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) *reinterpret_cast<float4*>(row + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int c = threadIdx.x; c < D; c += blockDim.x) row[c] = 0.f;
    }
  }
}

// row-set bitmaps of the backward's exact-zero hint (relgat_layer_bwd_src, dst_nz_bits): bit j = row j of the
// gradient table may be non-zero.  atomicOr is idempotent, so the result does not depend on the thread order.
__global__ void mark_rows_kernel(const long long* __restrict__ ids, long long n, long long n_rows,
                                 uint32_t* __restrict__ bits) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const long long j = ids[i];
  if (j >= 0 && j < n_rows) atomicOr(bits + (j >> 5), 1u << (j & 31));
}

// sources of the edges into marked destinations: the rows of dP — hence of the gradient handed to the layer below —
// that can be non-zero
__global__ void mark_sources_kernel(const uint32_t* __restrict__ dst_bits, const int* __restrict__ rowptr,
                                    const int* __restrict__ csr_src, int n_dst, uint32_t* __restrict__ src_bits) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_dst) return;
  if (!((__ldg(dst_bits + (j >> 5)) >> (j & 31)) & 1u)) return;
  const int e1 = __ldg(rowptr + j + 1);
  for (int e = __ldg(rowptr + j); e < e1; ++e) {
    const int i = __ldg(csr_src + e);
    atomicOr(src_bits + (i >> 5), 1u << (i & 31));
  }
}

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_mark_rows(const long long* ids, long long n, long long n_rows, unsigned int* bits, void* stream) {
  if (n < 0 || n_rows < 0) return RG_ERR_ARG;
  if (n == 0) return RG_OK;
  if (!ids || !bits) return RG_ERR_ARG;
  mark_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, n, n_rows,
                                                                                                       bits);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_mark_sources(const unsigned int* dst_bits, const int* rowptr, const int* csr_src, int n_dst,
                                   unsigned int* src_bits, void* stream) {
  if (n_dst < 0) return RG_ERR_ARG;
  if (n_dst == 0) return RG_OK;
  if (!dst_bits || !rowptr || !csr_src || !src_bits) return RG_ERR_ARG;
  mark_sources_kernel<<<(n_dst + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst_bits, rowptr, csr_src,
                                                                                         n_dst, src_bits);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_bernoulli_bits(unsigned int* bits, long long n_words, float p_drop, unsigned long long seed,
                                     void* stream) {
  if (!bits || n_words < 0 || !(p_drop >= 0.f) || p_drop > 1.f) return RG_ERR_ARG;
  if (n_words == 0) return RG_OK;
  // P(drop) = threshold / 65536 (16-bit draws: |p - P(drop)| <= 7.6e-6)
  uint32_t threshold = static_cast<uint32_t>(p_drop * 65536.f + 0.5f);
  if (threshold > 65536u) threshold = 65536u;
  const int th = 256;
  bernoulli_bits_kernel<<<static_cast<unsigned>((n_words + th - 1) / th), th, 0, static_cast<cudaStream_t>(stream)>>>(
      bits, n_words, threshold, seed);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_zero_rows(float* table, long long ld, const long long* ids, long long n, int D, void* stream) {
  if (n < 0 || D <= 0) return RG_ERR_ARG;
  if (n == 0) return RG_OK;
  if (!table || !ids) return RG_ERR_ARG;
  const long long cap = 148 * 16;  // a grid-stride loop: long lists with most entries switched off stay cheap
  zero_rows_kernel<<<static_cast<unsigned>(n < cap ? n : cap), 128, 0, static_cast<cudaStream_t>(stream)>>>(table, ld, ids, n, D);
  return cuda_status(cudaGetLastError());
}
