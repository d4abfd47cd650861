// E1 — fused RelGAT edge forward (replaces reference layer.py:220-318, ops K2-K11 of
// SURVEY.md §2.2): per-edge logit + LeakyReLU, destination-segmented softmax, attention-
// weighted aggregation and relation bias in ONE pass over the CSR (by-destination) edges.
//
// One warp = one (destination j, head-group) task.  The projected source rows P[src] are
// gathered with 128-bit streaming loads and read exactly once; the softmax is the online
// (running max / running sum) form so nothing of size [E, F] is ever materialised.
// Segment order = CSR order = original edge order inside a destination (deterministic).
//
// HBM-bound: algorithmic bytes per launch = E*(C*s + 8) + N*(C*s_out + 4) + 2*E*H*4.
#include "common.cuh"

namespace relgat {

constexpr int kFwdWarps = 4;

template <typename T, int V>
struct FwdArgs {
  const T* P;            // [N_src, H*F] projected features (row stride = ldp elements)
  const float* A;        // [H, R, F] attention vectors
  const float* beta;     // [R] relation bias or nullptr
  const int* rowptr;     // [N+1]
  const int* csr_src;    // [E]
  const int* csr_rel;    // [E]
  float* out;            // [N, H*F] fp32 layer output (pre-activation), may be nullptr
  __nv_bfloat16* act_hi; // [N, H*F] optional bf16 copy of act(out) (hi part)
  __nv_bfloat16* act_lo; // [N, H*F] optional residual (lo part); nullptr = hi only
  float* alpha;          // [E, H] attention weights (CSR order)
  float* z;              // [E, H] pre-activation logits (CSR order)
  float* bias_out;       // [N] sum of relation biases per destination
  int N, H, F, R, hg;
  long long ldp;         // row stride of P in elements
  int apply_elu;         // act = ELU (reference model.py:286-287) else identity
  int max_deg;           // rows with more in-edges are left to the hub path (<=0: no limit)
};

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

template <typename T, int V>
__global__ void __launch_bounds__(kFwdWarps * 32, 3)
edge_fwd_kernel(const FwdArgs<T, V> a) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long task = static_cast<long long>(blockIdx.x) * kFwdWarps + warp;
  if (task >= static_cast<long long>(a.N) * groups) return;
  const int j = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(j) * groups);
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;

  const int lo = a.rowptr[j];
  const int hi = a.rowptr[j + 1];
  if (a.max_deg > 0 && hi - lo > a.max_deg) return;  // hub: handled by the split path

  float acc[kMaxVecPerLane][V];
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[k][v] = 0.f;
  float m = -INFINITY, l = 0.f, bsum = 0.f;

  for (int base = lo; base < hi; base += 32) {
    const int cnt = min(32, hi - base);
    int my_src = 0, my_rel = 0;
    if (lane < cnt) {
      my_src = __ldg(a.csr_src + base + lane);
      my_rel = __ldg(a.csr_rel + base + lane);
    }
    for (int t = 0; t < cnt; t += 2) {
      const bool two = (t + 1 < cnt);
      const int i0 = __shfl_sync(0xffffffffu, my_src, t);
      const int r0 = __shfl_sync(0xffffffffu, my_rel, t);
      const int i1 = __shfl_sync(0xffffffffu, my_src, two ? t + 1 : t);
      const int r1 = __shfl_sync(0xffffffffu, my_rel, two ? t + 1 : t);
      const T* p0 = a.P + static_cast<long long>(i0) * a.ldp + lm.head_off;
      const T* p1 = a.P + static_cast<long long>(i1) * a.ldp + lm.head_off;
      const float* a0 = a.A + (static_cast<long long>(lm.hh) * a.R + r0) * a.F;
      const float* a1 = a.A + (static_cast<long long>(lm.hh) * a.R + r1) * a.F;
      float x0[kMaxVecPerLane][V], x1[kMaxVecPerLane][V];
      // issue both row gathers before any arithmetic (two rows in flight per warp)
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) RowVec<T, V>::load_stream(p0 + q * V, x0[k]);
      }
      if (two) {
#pragma unroll
        for (int k = 0; k < kMaxVecPerLane; ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) RowVec<T, V>::load_stream(p1 + q * V, x1[k]);
        }
      }
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) {
          float av[V];
          RowVec<float, V>::load_cached(a0 + q * V, av);
#pragma unroll
          for (int v = 0; v < V; ++v) d0 = fmaf(x0[k][v], av[v], d0);
          if (two) {
            RowVec<float, V>::load_cached(a1 + q * V, av);
#pragma unroll
            for (int v = 0; v < V; ++v) d1 = fmaf(x1[k][v], av[v], d1);
          }
        }
      }
      d0 = head_sum(d0, lm.lph);
      d1 = head_sum(d1, lm.lph);
      if (lm.sub == 0) {
        a.z[static_cast<long long>(base + t) * a.H + lm.hh] = d0;
        if (two) a.z[static_cast<long long>(base + t + 1) * a.H + lm.hh] = d1;
      }
      // edge t
      {
        const float e = d0 > 0.f ? d0 : kLeakySlope * d0;
        const float mn = fmaxf(m, e);
        const float sc = expf(m - mn);
        const float w = expf(e - mn);
        l = fmaf(l, sc, w);
#pragma unroll
        for (int k = 0; k < kMaxVecPerLane; ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[k][v] = fmaf(acc[k][v], sc, w * x0[k][v]);
          }
        }
        m = (e != e) ? e : mn;  // NaN logits poison the row like the reference does
        if (a.beta) bsum += __ldg(a.beta + r0);
      }
      if (two) {
        const float e = d1 > 0.f ? d1 : kLeakySlope * d1;
        const float mn = fmaxf(m, e);
        const float sc = expf(m - mn);
        const float w = expf(e - mn);
        l = fmaf(l, sc, w);
#pragma unroll
        for (int k = 0; k < kMaxVecPerLane; ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[k][v] = fmaf(acc[k][v], sc, w * x1[k][v]);
          }
        }
        m = (e != e) ? e : mn;
        if (a.beta) bsum += __ldg(a.beta + r1);
      }
    }
  }

  const bool empty = (hi == lo);
  const float inv = empty ? 0.f : 1.f / fmaxf(l, 1e-16f);  // reference layer.py:291 clamp
  // out = acc / den + bias  (bias is added to every head and channel, layer.py:313-318)
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      float o[V];
#pragma unroll
      for (int v = 0; v < V; ++v) o[v] = empty ? 0.f : fmaf(acc[k][v], inv, bsum);
      const long long off = static_cast<long long>(j) * C + lm.head_off + q * V;
      if (a.out) RowVec<float, V>::store(a.out + off, o);
      if (a.act_hi) {
        if (a.apply_elu) {
#pragma unroll
          for (int v = 0; v < V; ++v) o[v] = elu1(o[v]);
        }
        store_split_bf16<V>(a.act_hi + off, a.act_lo ? a.act_lo + off : nullptr, o);
      }
    }
  }
  if (g == 0 && lane == 0 && a.bias_out) a.bias_out[j] = empty ? 0.f : bsum;

  // attention weights: alpha = exp(eps - m) / den for every (edge, head of this group)
  __syncwarp();
  const int items = (hi - lo) * a.hg;
  for (int it = 0; it < items; it += 32) {
    const int idx = it + lane;
    const int hgi = idx % a.hg;
    const float mh = __shfl_sync(0xffffffffu, m, hgi * lm.lph);
    const float ih = __shfl_sync(0xffffffffu, inv, hgi * lm.lph);
    if (idx < items) {
      const long long o = static_cast<long long>(lo + idx / a.hg) * a.H + g * a.hg + hgi;
      const float zz = a.z[o];
      const float e = zz > 0.f ? zz : kLeakySlope * zz;
      a.alpha[o] = expf(e - mh) * ih;
    }
  }
}

template <typename T, int V>
static int launch_fwd(const FwdArgs<T, V>& a, cudaStream_t stream) {
  const long long tasks = static_cast<long long>(a.N) * (a.H / a.hg);
  if (tasks == 0) return RG_OK;
  const long long blocks = (tasks + kFwdWarps - 1) / kFwdWarps;
  edge_fwd_kernel<T, V><<<static_cast<unsigned>(blocks), kFwdWarps * 32, 0, stream>>>(a);
  return cuda_status(cudaGetLastError());
}

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_layer_fwd(
    const void* P, int p_is_bf16, long long ldp, const float* A, const float* beta,
    const int* rowptr, const int* csr_src, const int* csr_rel,
    float* out, void* act_hi, void* act_lo, int apply_elu,
    float* alpha, float* z, float* bias_out,
    int N, int H, int F, int R, int max_deg, void* stream) {
  if (!P || !A || !rowptr || !alpha || !z || N < 0 || H <= 0 || F <= 0 || R <= 0) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(P) % 16 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                      (!out || reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (p_is_bf16) {
    if (F % 8 != 0 || ldp % 8 != 0) return RG_ERR_SHAPE;
    if (!vec_ok) return RG_ERR_ALIGN;
    const int hg = pick_heads_per_warp(H, F, 8);
    if (!hg) return RG_ERR_SHAPE;
    // A is read with the fp32 vector type of the same element count: 8 floats = two float4
    return RG_ERR_DTYPE;  // bf16 feature storage is wired in edge_fwd_bf16.cu (not built yet)
  }
  const bool v4 = (F % 4 == 0) && (ldp % 4 == 0) && vec_ok;
  if (v4) {
    const int hg = pick_heads_per_warp(H, F, 4);
    if (!hg) return RG_ERR_SHAPE;
    FwdArgs<float, 4> a{static_cast<const float*>(P), A, beta, rowptr, csr_src, csr_rel, out,
                        static_cast<__nv_bfloat16*>(act_hi), static_cast<__nv_bfloat16*>(act_lo),
                        alpha, z, bias_out, N, H, F, R, hg, ldp, apply_elu, max_deg};
    return launch_fwd(a, s);
  }
  const int hg = pick_heads_per_warp(H, F, 1);
  if (!hg) return RG_ERR_SHAPE;
  FwdArgs<float, 1> a{static_cast<const float*>(P), A, beta, rowptr, csr_src, csr_rel, out,
                      static_cast<__nv_bfloat16*>(act_hi), static_cast<__nv_bfloat16*>(act_lo),
                      alpha, z, bias_out, N, H, F, R, hg, ldp, apply_elu, max_deg};
  return launch_fwd(a, s);
}
