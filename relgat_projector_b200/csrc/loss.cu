// S2 — ranking and reconstruction losses fused behind the gather-score kernels (BASELINE.json north_star:
// "scoring, with ranking and reconstruction losses"; replaces reference core/loss/relgat_loss.py:32-71,
// core/loss/multi_objective_loss.py:47-83, core/loss/cosine.py:4-13, core/loss/mse.py:4-10 and the score
// sanitisation of trainer/relgat_projector.py:578-584, 647-648).
//
// Every kernel produces the loss value AND its gradient with respect to its inputs in the same pass, so the
// backward of the scorer needs no further loss arithmetic.  All reductions run in a fixed order (per-row
// partials, then one block with a fixed tree) => bitwise reproducible.
#include "common.cuh"

namespace relgat {

enum : int { RANK_MARGIN = 0, RANK_SELF_ADV = 1 };

// log(sigmoid(x)) = min(x, 0) - log1p(exp(-|x|))   (the form torch uses)
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// nan -> 0, +-inf -> +-1e9 (trainer:584, 647-648); *finite tells whether a gradient may flow back
__device__ __forceinline__ float sanitize_score(float x, int on, bool* finite) {
  *finite = true;
  if (!on) return x;
  if (x != x) { *finite = false; return 0.f; }
  if (isinf(x)) { *finite = false; return x > 0.f ? 1e9f : -1e9f; }
  return x;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

// pos[b], neg[b*sb + k*sk]; dpos / dneg use the same addressing.  Single block: B*K is a few thousand scalars.
__global__ void __launch_bounds__(256)
rank_loss_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int B, int K, long long sb,
                 long long sk, int type, float margin, float alpha, int sanitize, float* __restrict__ loss,
                 float* __restrict__ dpos, float* __restrict__ dneg) {
  __shared__ float red[256];
  const int tid = threadIdx.x;
  const float inv_b = B > 0 ? 1.f / static_cast<float>(B) : 0.f;
  float acc = 0.f;
  if (type == RANK_MARGIN) {
    // mean_{b,k} relu(margin + neg[b,k] - pos[b])   (relgat_loss.py:51-54)
    const float inv = (B > 0 && K > 0) ? 1.f / (static_cast<float>(B) * K) : 0.f;
    for (int b = tid; b < B; b += blockDim.x) {
      bool pf;
      const float p = sanitize_score(pos[b], sanitize, &pf);
      float dp = 0.f;
      for (int k = 0; k < K; ++k) {
        const long long o = b * sb + k * sk;
        bool nf;
        const float v = margin + sanitize_score(neg[o], sanitize, &nf) - p;
        const bool on = v > 0.f;
        acc += on ? v : 0.f;
        dneg[o] = (on && nf) ? inv : 0.f;
        dp -= on ? inv : 0.f;
      }
      dpos[b] = pf ? dp : 0.f;
    }
    const float tot = block_sum_256(acc, red);
    // K == 0: the reference takes the mean of an empty tensor (nan)
    if (tid == 0) loss[0] = (K > 0 && B > 0) ? tot * inv : __int_as_float(0x7fc00000);
    return;
  }
  // self-adversarial (relgat_loss.py:56-71): -mean_b logsig(pos) - mean_b sum_k softmax_k(alpha*neg).detach() * logsig(-neg)
  for (int b = tid; b < B; b += blockDim.x) {
    bool pf;
    const float p = sanitize_score(pos[b], sanitize, &pf);
    acc -= log_sigmoid(p);
    dpos[b] = pf ? -sigmoid(-p) * inv_b : 0.f;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      bool nf;
      m = fmaxf(m, alpha * sanitize_score(neg[b * sb + k * sk], sanitize, &nf));
    }
    float den = 0.f;
    for (int k = 0; k < K; ++k) {
      bool nf;
      den += expf(alpha * sanitize_score(neg[b * sb + k * sk], sanitize, &nf) - m);
    }
    const float iden = den > 0.f ? 1.f / den : 0.f;
    for (int k = 0; k < K; ++k) {
      const long long o = b * sb + k * sk;
      bool nf;
      const float n = sanitize_score(neg[o], sanitize, &nf);
      const float w = expf(alpha * n - m) * iden;
      acc -= w * log_sigmoid(-n);
      dneg[o] = nf ? w * sigmoid(n) * inv_b : 0.f;
    }
  }
  const float tot = block_sum_256(acc, red);
  if (tid == 0) loss[0] = B > 0 ? tot * inv_b : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------------
// reconstruction terms: one warp per positive b
//   c_pos[b]   = cos(tr_b, dst_b)          c_neg[b,k] = cos(tr_b, nd_{b,k})       (F.normalize, eps 1e-12)
//   values     = (mean_b (1 - c_pos), mean_{k,b} (1 - c_neg), mean_{b,d} (tr - dst)^2)
//   gradients of  S = w_pos * values[0] + w_neg * (1 - values[1]) + w_mse * values[2]
// nd_{b,k} lives at negdst + b*nsb + k*nsk (element strides), row length D.
// ------------------------------------------------------------------------------------------------------
constexpr int kReconWarps = 4;
constexpr int kReconMaxK = 256;
constexpr float kCosEps = 1e-12f;

struct ReconArgs {
  const float* tr;
  const float* dst;
  const float* negdst;
  int B, K, D;
  long long nsb, nsk;
  float w_pos, w_neg, w_mse;
  float* partial;   // [B, 3] per-positive sums: (1 - c_pos), sum_k (1 - c_neg), sum_d (tr - dst)^2
  float* d_tr;      // [B, D]
  float* d_dst;     // [B, D]
  float* d_negdst;  // same addressing as negdst, or nullptr
};

template <int V>
__global__ void __launch_bounds__(kReconWarps * 32) recon_loss_kernel(const ReconArgs a) {
  __shared__ float sm_c[kReconWarps][kReconMaxK];
  __shared__ float sm_in[kReconWarps][kReconMaxK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kReconWarps + warp;
  if (b >= a.B) return;
  const float* p = a.tr + static_cast<long long>(b) * a.D;
  const float* t = a.dst + static_cast<long long>(b) * a.D;
  float pp = 0.f, tt = 0.f, pt = 0.f, se = 0.f;
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float pv[V], tv[V];
    RowVec<float, V>::load_cached(p + c, pv);
    RowVec<float, V>::load_cached(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      pp = fmaf(pv[v], pv[v], pp);
      tt = fmaf(tv[v], tv[v], tt);
      pt = fmaf(pv[v], tv[v], pt);
      const float d = pv[v] - tv[v];
      se = fmaf(d, d, se);
    }
  }
  pp = warp_sum(pp); tt = warp_sum(tt); pt = warp_sum(pt); se = warp_sum(se);
  const float np_raw = sqrtf(pp), nt_raw = sqrtf(tt);
  const float inp = 1.f / fmaxf(np_raw, kCosEps), int_ = 1.f / fmaxf(nt_raw, kCosEps);
  const bool p_unit = np_raw > kCosEps, t_unit = nt_raw > kCosEps;  // below eps normalize() is linear (x / eps)
  const float c_pos = pt * inp * int_;
  float neg_sum = 0.f;
  for (int k = 0; k < a.K; ++k) {
    const float* nd = a.negdst + b * a.nsb + k * a.nsk;
    float nn = 0.f, pn = 0.f;
    for (int c = lane * V; c < a.D; c += 32 * V) {
      float pv[V], nv[V];
      RowVec<float, V>::load_cached(p + c, pv);
      RowVec<float, V>::load_cached(nd + c, nv);
#pragma unroll
      for (int v = 0; v < V; ++v) { nn = fmaf(nv[v], nv[v], nn); pn = fmaf(pv[v], nv[v], pn); }
    }
    nn = warp_sum(nn); pn = warp_sum(pn);
    const float nraw = sqrtf(nn);
    const float inn = 1.f / fmaxf(nraw, kCosEps);
    const float ck = pn * inp * inn;
    neg_sum += 1.f - ck;
    if (lane == 0) {
      sm_c[warp][k] = ck;
      sm_in[warp][k] = nraw > kCosEps ? inn : -inn;  // sign carries the "unit norm" flag
    }
  }
  __syncwarp();
  if (lane == 0) {
    a.partial[b * 3 + 0] = 1.f - c_pos;
    a.partial[b * 3 + 1] = neg_sum;
    a.partial[b * 3 + 2] = se;
  }
  // gradients.  d(1 - c)/dp = -(t_n - [p unit] c p_n) / max(|p|, eps), same with the roles swapped.
  const float g_pos = -a.w_pos / static_cast<float>(a.B);                                     // dS / dc_pos
  const float g_neg = a.K > 0 ? a.w_neg / (static_cast<float>(a.B) * a.K) : 0.f;             // dS / dc_neg
  const float g_mse = a.w_mse * 2.f / (static_cast<float>(a.B) * a.D);
  // coefficient of p_n in d_tr from all cosine terms: -(g_pos c_pos + g_neg sum_k c_k) when p has unit norm
  float csum = g_pos * c_pos;
  for (int k = 0; k < a.K; ++k) csum = fmaf(g_neg, sm_c[warp][k], csum);
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float pv[V], tv[V], dp[V], dt[V];
    RowVec<float, V>::load_cached(p + c, pv);
    RowVec<float, V>::load_cached(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float pn_ = pv[v] * inp, tn_ = tv[v] * int_;
      const float diff = pv[v] - tv[v];
      dp[v] = g_pos * tn_ - (p_unit ? csum * pn_ : 0.f);
      dt[v] = g_pos * (pn_ - (t_unit ? c_pos * tn_ : 0.f)) * int_ - g_mse * diff;
    }
    for (int k = 0; k < a.K; ++k) {
      const long long off = b * a.nsb + k * a.nsk + c;
      float nv[V], dn[V];
      RowVec<float, V>::load_cached(a.negdst + off, nv);
      const float ck = sm_c[warp][k];
      const float inn_s = sm_in[warp][k];
      const float inn = fabsf(inn_s);
      const bool n_unit = inn_s > 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float nn_ = nv[v] * inn;
        dp[v] = fmaf(g_neg, nn_, dp[v]);
        dn[v] = g_neg * (pv[v] * inp - (n_unit ? ck * nn_ : 0.f)) * inn;
      }
      if (a.d_negdst) RowVec<float, V>::store(a.d_negdst + off, dn);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) dp[v] = fmaf(dp[v], inp, g_mse * (pv[v] - tv[v]));
    RowVec<float, V>::store(a.d_tr + static_cast<long long>(b) * a.D + c, dp);
    RowVec<float, V>::store(a.d_dst + static_cast<long long>(b) * a.D + c, dt);
  }
}

__global__ void __launch_bounds__(256)
recon_reduce_kernel(const float* __restrict__ partial, int B, int K, int D, float* __restrict__ values) {
  __shared__ float red[256];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    s0 += partial[b * 3 + 0];
    s1 += partial[b * 3 + 1];
    s2 += partial[b * 3 + 2];
  }
  s0 = block_sum_256(s0, red);
  s1 = block_sum_256(s1, red);
  s2 = block_sum_256(s2, red);
  if (threadIdx.x == 0) {
    const float nan = __int_as_float(0x7fc00000);
    values[0] = B > 0 ? s0 / B : nan;
    values[1] = (B > 0 && K > 0) ? s1 / (static_cast<float>(B) * K) : nan;
    values[2] = (B > 0 && D > 0) ? s2 / (static_cast<float>(B) * D) : nan;
  }
}

static inline bool al16(const void* p) { return !p || reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_rank_loss(const float* pos, const float* neg, int B, int K, long long stride_b,
                                long long stride_k, int type, float margin, float alpha, int sanitize, float* loss,
                                float* dpos, float* dneg, void* stream) {
  if (!pos || !loss || !dpos || B < 0 || K < 0) return RG_ERR_ARG;
  if (K > 0 && (!neg || !dneg)) return RG_ERR_ARG;
  if (type != RANK_MARGIN && type != RANK_SELF_ADV) return RG_ERR_ARG;
  rank_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, neg, B, K, stride_b, stride_k, type, margin,
                                                                     alpha, sanitize, loss, dpos, dneg);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_recon_loss(const float* tr, const float* dst, const float* negdst, int B, int K, int D,
                                 long long neg_stride_b, long long neg_stride_k, float w_pos, float w_neg, float w_mse,
                                 float* values, float* partial, float* d_tr, float* d_dst, float* d_negdst,
                                 void* stream) {
  if (!tr || !dst || !values || !partial || !d_tr || !d_dst || B < 0 || K < 0 || D <= 0) return RG_ERR_ARG;
  if (K > 0 && !negdst) return RG_ERR_ARG;
  if (K > kReconMaxK) return RG_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B > 0) {
    ReconArgs a{tr, dst, negdst, B, K, D, neg_stride_b, neg_stride_k, w_pos, w_neg, w_mse, partial, d_tr, d_dst, d_negdst};
    const unsigned blocks = (B + kReconWarps - 1) / kReconWarps;
    const bool v4 = D % 4 == 0 && al16(tr) && al16(dst) && al16(negdst) && al16(d_tr) && al16(d_dst) && al16(d_negdst) &&
                    neg_stride_b % 4 == 0 && neg_stride_k % 4 == 0;
    if (v4) recon_loss_kernel<4><<<blocks, kReconWarps * 32, 0, s>>>(a);
    else recon_loss_kernel<1><<<blocks, kReconWarps * 32, 0, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_status(e);
  }
  recon_reduce_kernel<<<1, 256, 0, s>>>(partial, B, K, D, values);
  return cuda_status(cudaGetLastError());
}
