// GELU + LayerNorm of the ProjectionHead in one pass per direction (reference core/model/projection.py:48-67: every
// hidden block is Linear -> GELU -> LayerNorm).  The linears run on the tcgen05 GEMM; this kernel
// replaces the two ATen element-wise kernels between them (one read + one write of the [M, D] activations per
// direction instead of two of each), with fixed-order reductions.
//
//   forward : g = gelu(h) (exact, erf form: nn.GELU() default), mu / rstd over the row (biased variance, eps inside
//             the root like F.layer_norm), y = (g - mu) * rstd * gamma + beta; mu and rstd are saved.
//   backward: ghat = (g - mu) * rstd, dghat = dy * gamma,
//             dg = rstd * (dghat - mean(dghat) - ghat * mean(dghat * ghat)),  dh = dg * gelu'(h),
//             dgamma = sum_rows dy * ghat, dbeta = sum_rows dy  (per-CTA partials over kRowsPerCta rows, then an
//             ordered reduction).
#include "common.cuh"

namespace relgat {

constexpr int kLnThreads = 128;
constexpr int kRowsPerCta = 32;

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// sum over the block, result broadcast to every thread; fixed tree => reproducible
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kLnThreads / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(kLnThreads)
gelu_ln_fwd_kernel(const float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int M, int D, float eps) {
  extern __shared__ float row[];  // D floats: gelu(h) of the row
  __shared__ float red[kLnThreads / 32];
  const int r = blockIdx.x;
  if (r >= M) return;
  const float* hr = h + static_cast<long long>(r) * D;
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += kLnThreads) {
    const float g = gelu_f(hr[c]);
    row[c] = g;
    s += g;
  }
  const float mu = block_sum(s, red) / D;
  float q = 0.f;
  for (int c = threadIdx.x; c < D; c += kLnThreads) {
    const float d = row[c] - mu;
    q = fmaf(d, d, q);
  }
  const float rs = rsqrtf(block_sum(q, red) / D + eps);
  float* yr = y + static_cast<long long>(r) * D;
  for (int c = threadIdx.x; c < D; c += kLnThreads)
    yr[c] = fmaf((row[c] - mu) * rs, gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f);
  if (threadIdx.x == 0) { mean[r] = mu; rstd[r] = rs; }
}

// one CTA per group of kRowsPerCta rows: dh rows + this group's partial dgamma / dbeta
__global__ void __launch_bounds__(kLnThreads)
gelu_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ h, const float* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dh,
                   float* __restrict__ part_g, float* __restrict__ part_b, int M, int D) {
  extern __shared__ float sm[];  // [D] ghat of the row, [D] dghat, [D] + [D] this CTA's column partials
  float* ghat = sm;
  float* dgh = sm + D;
  float* pg = sm + 2 * D;
  float* pb = sm + 3 * D;
  __shared__ float red[kLnThreads / 32];
  const int r0 = blockIdx.x * kRowsPerCta;
  const int r1 = min(M, r0 + kRowsPerCta);
  for (int c = threadIdx.x; c < D; c += kLnThreads) { pg[c] = 0.f; pb[c] = 0.f; }
  for (int r = r0; r < r1; ++r) {
    const float mu = mean[r], rs = rstd[r];
    const float* hr = h + static_cast<long long>(r) * D;
    const float* dyr = dy + static_cast<long long>(r) * D;
    float s1 = 0.f, s2 = 0.f;
    __syncthreads();  // the previous row's buffers are free
    for (int c = threadIdx.x; c < D; c += kLnThreads) {
      const float gh = (gelu_f(hr[c]) - mu) * rs;
      const float d = dyr[c] * (gamma ? gamma[c] : 1.f);
      ghat[c] = gh;
      dgh[c] = d;
      s1 += d;
      s2 = fmaf(d, gh, s2);
      // the same thread owns column c for every row: plain read-modify-write, fixed row order
      pg[c] = fmaf(dyr[c], gh, pg[c]);
      pb[c] += dyr[c];
    }
    const float m1 = block_sum(s1, red) / D;
    const float m2 = block_sum(s2, red) / D;
    float* dhr = dh + static_cast<long long>(r) * D;
    for (int c = threadIdx.x; c < D; c += kLnThreads)
      dhr[c] = rs * (dgh[c] - m1 - ghat[c] * m2) * gelu_grad_f(hr[c]);
  }
  for (int c = threadIdx.x; c < D; c += kLnThreads) {
    part_g[static_cast<long long>(blockIdx.x) * D + c] = pg[c];
    part_b[static_cast<long long>(blockIdx.x) * D + c] = pb[c];
  }
}

__global__ void gelu_ln_param_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, int groups,
                                            int D, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float sg = 0.f, sb = 0.f;
  for (int g = 0; g < groups; ++g) {
    sg += part_g[static_cast<long long>(g) * D + c];
    sb += part_b[static_cast<long long>(g) * D + c];
  }
  if (dgamma) dgamma[c] = sg;
  if (dbeta) dbeta[c] = sb;
}

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_gelu_layernorm_fwd(const float* h, const float* gamma, const float* beta, float* y, float* mean,
                                         float* rstd, int M, int D, float eps, void* stream) {
  if (!h || !y || !mean || !rstd || M < 0 || D <= 0) return RG_ERR_ARG;
  if (M == 0) return RG_OK;
  const size_t smem = static_cast<size_t>(D) * sizeof(float);
  if (smem > 200 * 1024) return RG_ERR_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gelu_ln_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_status(e);
  }
  gelu_ln_fwd_kernel<<<M, kLnThreads, smem, static_cast<cudaStream_t>(stream)>>>(h, gamma, beta, y, mean, rstd, M, D, eps);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_gelu_layernorm_groups(int M) { return M <= 0 ? 0 : (M + kRowsPerCta - 1) / kRowsPerCta; }

// part_g / part_b: float [relgat_gelu_layernorm_groups(M) * D] scratch each
extern "C" int relgat_gelu_layernorm_bwd(const float* dy, const float* h, const float* gamma, const float* mean,
                                         const float* rstd, float* dh, float* part_g, float* part_b, float* dgamma,
                                         float* dbeta, int M, int D, void* stream) {
  if (!dy || !h || !mean || !rstd || !dh || !part_g || !part_b || M < 0 || D <= 0) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = relgat_gelu_layernorm_groups(M);
  if (M > 0) {
    const size_t smem = 4 * static_cast<size_t>(D) * sizeof(float);
    if (smem > 200 * 1024) return RG_ERR_SHAPE;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(gelu_ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_status(e);
    }
    gelu_ln_bwd_kernel<<<groups, kLnThreads, smem, s>>>(dy, h, gamma, mean, rstd, dh, part_g, part_b, M, D);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_status(e);
  }
  gelu_ln_param_reduce_kernel<<<(D + 127) / 128, 128, 0, s>>>(part_g, part_b, groups, D, dgamma, dbeta);
  return cuda_status(cudaGetLastError());
}
