// E2/E3 — RelGAT edge backward (replaces the autograd replay K17-K19 of SURVEY.md §2.2).
//
// Closed form, with one algebraic step that removes a whole gather pass:
//   t[j,h] = sum_e alpha*dalpha = <G[j,h,:], out[j,h,:] - bias_j>
// (the softmax-backward row term equals the dot product of the incoming gradient with the
// un-biased forward output, so it needs no edge traversal).  With t known per destination
// the per-edge quantities no longer need a destination-segmented reduction, and the edge
// traversal order is free.  We pick the order that makes each reduction local and
// deterministic:
//   bwd_prep : per node   — G = dY * act'(out); t[j,h]; hsum[j,h] = sum_f G[j,h,f]
//   bwd_src  : by SOURCE  — dP[i] = sum_{e: src=i} (alpha*G[dst] + dz*A[rel]);  writes dz[e,h]
//   bwd_rel  : by RELATION (fixed-size chunks) — partial dA[h,r,:] = sum dz*P[src]; partial dbeta
//   bwd_rel_reduce : ordered sum of the chunk partials
// HBM-bound: per launch bwd_src gathers E*C*s bytes of G rows, bwd_rel gathers E*C*s of P rows.
#include "common.cuh"

namespace relgat {

constexpr int kBwdWarps = 4;

// ------------------------------------------------------------------------------------
// bwd_prep
// ------------------------------------------------------------------------------------
template <int V>
struct PrepArgs {
  const float* dY;     // [N, C] gradient w.r.t. the layer's (activated) output
  const float* out;    // [N, C] forward pre-activation output
  const float* bias;   // [N] forward bias_out
  float* G;            // [N, C] gradient w.r.t. out (may alias dY when apply_elu == 0)
  float* t;            // [N, H]
  float* hsum;         // [N, H]
  int N, H, F, hg, apply_elu;
};

template <int V>
__global__ void __launch_bounds__(kBwdWarps * 32) bwd_prep_kernel(const PrepArgs<V> a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long task = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  if (task >= static_cast<long long>(a.N) * groups) return;
  const int j = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(j) * groups);
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const long long row = static_cast<long long>(j) * a.H * a.F + lm.head_off;
  const float b = a.bias ? __ldg(a.bias + j) : 0.f;
  float tt = 0.f, hs = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      float dy[V], o[V];
      RowVec<float, V>::load_stream(a.dY + row + q * V, dy);
      RowVec<float, V>::load_stream(a.out + row + q * V, o);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // ELU'(x) = 1 (x > 0) else exp(x)  (reference model.py:286-287, torch ELU alpha = 1)
        const float gg = a.apply_elu ? (o[v] > 0.f ? dy[v] : dy[v] * expf(o[v])) : dy[v];
        dy[v] = gg;
        tt = fmaf(gg, o[v] - b, tt);
        hs += gg;
      }
      if (a.G != a.dY || a.apply_elu) RowVec<float, V>::store(a.G + row + q * V, dy);
    }
  }
  tt = head_sum(tt, lm.lph);
  hs = head_sum(hs, lm.lph);
  if (lm.sub == 0) {
    a.t[static_cast<long long>(j) * a.H + lm.hh] = tt;
    a.hsum[static_cast<long long>(j) * a.H + lm.hh] = hs;
  }
}

// ------------------------------------------------------------------------------------
// bwd_src
// ------------------------------------------------------------------------------------
template <int V>
struct SrcArgs {
  const float* P;       // [N_src, C]   (row stride ldp)
  const float* G;       // [N_dst, C]
  const float* A;       // [H, R, F]
  const float* alpha;   // [E, H] CSR order
  const float* z;       // [E, H] CSR order
  const float* t;       // [N_dst, H]
  const int* colptr;    // [N_src+1]
  const int* csc_slot;  // [E] CSR slot of each by-source edge
  const int* csc_dst;   // [E]
  const int* csc_rel;   // [E]
  float* dP;            // [N_src, C] fp32 (may be nullptr when only the bf16 split is wanted)
  __nv_bfloat16* dP_hi; // optional bf16 split of dP for the tensor-core GEMMs
  __nv_bfloat16* dP_lo;
  float* dz;            // [E, H] CSR order
  int N, H, F, R, hg;
  long long ldp;
  int max_deg;
};

template <int V>
__global__ void __launch_bounds__(kBwdWarps * 32, 3) bwd_src_kernel(const SrcArgs<V> a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long task = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  if (task >= static_cast<long long>(a.N) * groups) return;
  const int i = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(i) * groups);
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int lo = a.colptr[i], hi = a.colptr[i + 1];
  if (a.max_deg > 0 && hi - lo > a.max_deg) return;

  float p[kMaxVecPerLane][V], acc[kMaxVecPerLane][V];
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int q = lm.sub + lm.lph * k;
#pragma unroll
    for (int v = 0; v < V; ++v) { acc[k][v] = 0.f; p[k][v] = 0.f; }
    if (hi > lo && q < lm.vph)
      RowVec<float, V>::load_stream(a.P + static_cast<long long>(i) * a.ldp + lm.head_off + q * V, p[k]);
  }

  for (int base = lo; base < hi; base += 32) {
    const int cnt = min(32, hi - base);
    int my_slot = 0, my_dst = 0, my_rel = 0;
    if (lane < cnt) {
      my_slot = __ldg(a.csc_slot + base + lane);
      my_dst = __ldg(a.csc_dst + base + lane);
      my_rel = __ldg(a.csc_rel + base + lane);
    }
    for (int tix = 0; tix < cnt; tix += 2) {
      const bool two = (tix + 1 < cnt);
      const int u1 = two ? tix + 1 : tix;
      const int s0 = __shfl_sync(0xffffffffu, my_slot, tix), s1 = __shfl_sync(0xffffffffu, my_slot, u1);
      const int j0 = __shfl_sync(0xffffffffu, my_dst, tix), j1 = __shfl_sync(0xffffffffu, my_dst, u1);
      const int r0 = __shfl_sync(0xffffffffu, my_rel, tix), r1 = __shfl_sync(0xffffffffu, my_rel, u1);
      const float* g0 = a.G + static_cast<long long>(j0) * C + lm.head_off;
      const float* g1 = a.G + static_cast<long long>(j1) * C + lm.head_off;
      float x0[kMaxVecPerLane][V], x1[kMaxVecPerLane][V];
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) RowVec<float, V>::load_stream(g0 + q * V, x0[k]);
      }
      if (two) {
#pragma unroll
        for (int k = 0; k < kMaxVecPerLane; ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) RowVec<float, V>::load_stream(g1 + q * V, x1[k]);
        }
      }
      const long long e0 = static_cast<long long>(s0) * a.H + lm.hh;
      const long long e1 = static_cast<long long>(s1) * a.H + lm.hh;
      const float al0 = __ldg(a.alpha + e0), z0 = __ldg(a.z + e0), t0 = __ldg(a.t + static_cast<long long>(j0) * a.H + lm.hh);
      const float al1 = __ldg(a.alpha + e1), z1 = __ldg(a.z + e1), t1 = __ldg(a.t + static_cast<long long>(j1) * a.H + lm.hh);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            d0 = fmaf(x0[k][v], p[k][v], d0);
            if (two) d1 = fmaf(x1[k][v], p[k][v], d1);
          }
        }
      }
      d0 = head_sum(d0, lm.lph);  // dalpha
      d1 = head_sum(d1, lm.lph);
      const float dz0 = al0 * (d0 - t0) * (z0 > 0.f ? 1.f : kLeakySlope);
      const float dz1 = al1 * (d1 - t1) * (z1 > 0.f ? 1.f : kLeakySlope);
      if (lm.sub == 0) {
        a.dz[e0] = dz0;
        if (two) a.dz[e1] = dz1;
      }
      const float* a0 = a.A + (static_cast<long long>(lm.hh) * a.R + r0) * a.F;
      const float* a1 = a.A + (static_cast<long long>(lm.hh) * a.R + r1) * a.F;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) {
          float av[V];
          RowVec<float, V>::load_cached(a0 + q * V, av);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[k][v] = fmaf(al0, x0[k][v], fmaf(dz0, av[v], acc[k][v]));
          if (two) {
            RowVec<float, V>::load_cached(a1 + q * V, av);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[k][v] = fmaf(al1, x1[k][v], fmaf(dz1, av[v], acc[k][v]));
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      const long long off = static_cast<long long>(i) * C + lm.head_off + q * V;
      if (a.dP) RowVec<float, V>::store(a.dP + off, acc[k]);
      if (a.dP_hi) store_split_bf16<V>(a.dP_hi + off, a.dP_lo ? a.dP_lo + off : nullptr, acc[k]);
    }
  }
}

// ------------------------------------------------------------------------------------
// bwd_rel: chunk partials of dA and dbeta, then ordered reduce
// ------------------------------------------------------------------------------------
template <int V>
struct RelArgs {
  const float* P;        // [N_src, C]
  const float* dz;       // [E, H] CSR order
  const float* hsum;     // [N_dst, H]
  const int* rel_slot;   // [E] CSR slots in by-relation order
  const int* csr_src;    // [E]
  const int* csr_dst;    // [E]
  const int* chunk_lo;   // [n_chunks] range in rel_slot
  const int* chunk_hi;   // [n_chunks]
  float* partA;          // [n_chunks, C]
  float* partB;          // [n_chunks]
  int n_chunks, H, F, hg;
  long long ldp;
};

template <int V>
__global__ void __launch_bounds__(kBwdWarps * 32, 3) bwd_rel_kernel(const RelArgs<V> a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long task = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  if (task >= static_cast<long long>(a.n_chunks) * groups) return;
  const int c = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(c) * groups);
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int lo = a.chunk_lo[c], hi = a.chunk_hi[c];

  float acc[kMaxVecPerLane][V];
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[k][v] = 0.f;
  float bacc = 0.f;

  for (int base = lo; base < hi; base += 32) {
    const int cnt = min(32, hi - base);
    int my_slot = 0, my_src = 0;
    float hs = 0.f;
    if (lane < cnt) {
      my_slot = __ldg(a.rel_slot + base + lane);
      my_src = __ldg(a.csr_src + my_slot);
      if (g == 0) {  // dbeta partial: sum over heads of hsum[dst]
        const int jd = __ldg(a.csr_dst + my_slot);
        for (int h = 0; h < a.H; ++h) hs += __ldg(a.hsum + static_cast<long long>(jd) * a.H + h);
      }
    }
    if (g == 0) bacc += warp_sum(hs);  // fixed xor-tree: reproducible run to run
    for (int tix = 0; tix < cnt; tix += 2) {
      const bool two = (tix + 1 < cnt);
      const int u1 = two ? tix + 1 : tix;
      const int s0 = __shfl_sync(0xffffffffu, my_slot, tix), s1 = __shfl_sync(0xffffffffu, my_slot, u1);
      const int i0 = __shfl_sync(0xffffffffu, my_src, tix), i1 = __shfl_sync(0xffffffffu, my_src, u1);
      const float* p0 = a.P + static_cast<long long>(i0) * a.ldp + lm.head_off;
      const float* p1 = a.P + static_cast<long long>(i1) * a.ldp + lm.head_off;
      float x0[kMaxVecPerLane][V], x1[kMaxVecPerLane][V];
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) RowVec<float, V>::load_stream(p0 + q * V, x0[k]);
      }
      if (two) {
#pragma unroll
        for (int k = 0; k < kMaxVecPerLane; ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) RowVec<float, V>::load_stream(p1 + q * V, x1[k]);
        }
      }
      const float dz0 = __ldg(a.dz + static_cast<long long>(s0) * a.H + lm.hh);
      const float dz1 = two ? __ldg(a.dz + static_cast<long long>(s1) * a.H + lm.hh) : 0.f;
#pragma unroll
      for (int k = 0; k < kMaxVecPerLane; ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            acc[k][v] = fmaf(dz0, x0[k][v], acc[k][v]);
            if (two) acc[k][v] = fmaf(dz1, x1[k][v], acc[k][v]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxVecPerLane; ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) RowVec<float, V>::store(a.partA + static_cast<long long>(c) * C + lm.head_off + q * V, acc[k]);
  }
  if (g == 0 && lane == 0) a.partB[c] = bacc;
}

// dA[h, r, f] = sum over the chunks of relation r (ascending) of partA[c, h*F + f]
__global__ void bwd_rel_reduce_kernel(const float* __restrict__ partA, const float* __restrict__ partB,
                                      const int* __restrict__ rel_chunk_ptr, float* __restrict__ dA,
                                      float* __restrict__ dbeta, int R, int H, int F) {
  const int C = H * F;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < static_cast<long long>(R) * C) {
    const int r = static_cast<int>(idx / C);
    const int col = static_cast<int>(idx - static_cast<long long>(r) * C);
    const int lo = rel_chunk_ptr[r], hi = rel_chunk_ptr[r + 1];
    float s = 0.f;
    for (int c = lo; c < hi; ++c) s += partA[static_cast<long long>(c) * C + col];
    const int h = col / F, f = col - h * F;
    dA[(static_cast<long long>(h) * R + r) * F + f] = s;
  }
  if (dbeta && idx < R) {
    const int r = static_cast<int>(idx);
    float s = 0.f;
    for (int c = rel_chunk_ptr[r]; c < rel_chunk_ptr[r + 1]; ++c) s += partB[c];
    dbeta[r] = s;
  }
}

template <typename Args, typename K>
static int launch_tasks(K kernel, const Args& a, long long tasks, cudaStream_t s) {
  if (tasks == 0) return RG_OK;
  const long long blocks = (tasks + kBwdWarps - 1) / kBwdWarps;
  kernel<<<static_cast<unsigned>(blocks), kBwdWarps * 32, 0, s>>>(a);
  return cuda_status(cudaGetLastError());
}

static inline bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_layer_bwd_prep(const float* dY, const float* out, const float* bias, float* G, float* t,
                                     float* hsum, int N, int H, int F, int apply_elu, void* stream) {
  if (!dY || !out || !G || !t || !hsum || N < 0 || H <= 0 || F <= 0) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (F % 4 == 0 && al16(dY) && al16(out) && al16(G)) {
    const int hg = pick_heads_per_warp(H, F, 4);
    if (!hg) return RG_ERR_SHAPE;
    PrepArgs<4> a{dY, out, bias, G, t, hsum, N, H, F, hg, apply_elu};
    return launch_tasks(bwd_prep_kernel<4>, a, static_cast<long long>(N) * (H / hg), s);
  }
  const int hg = pick_heads_per_warp(H, F, 1);
  if (!hg) return RG_ERR_SHAPE;
  PrepArgs<1> a{dY, out, bias, G, t, hsum, N, H, F, hg, apply_elu};
  return launch_tasks(bwd_prep_kernel<1>, a, static_cast<long long>(N) * (H / hg), s);
}

extern "C" int relgat_layer_bwd_src(const float* P, long long ldp, const float* G, const float* A,
                                    const float* alpha, const float* z, const float* t,
                                    const int* colptr, const int* csc_slot, const int* csc_dst, const int* csc_rel,
                                    float* dP, void* dP_hi, void* dP_lo, float* dz,
                                    int N_src, int H, int F, int R, int max_deg, void* stream) {
  if (!P || !G || !A || !alpha || !z || !t || !colptr || !dz || N_src < 0 || H <= 0 || F <= 0 || R <= 0)
    return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (F % 4 == 0 && ldp % 4 == 0 && al16(P) && al16(G) && al16(A) && (!dP || al16(dP))) {
    const int hg = pick_heads_per_warp(H, F, 4);
    if (!hg) return RG_ERR_SHAPE;
    SrcArgs<4> a{P, G, A, alpha, z, t, colptr, csc_slot, csc_dst, csc_rel, dP,
                 static_cast<__nv_bfloat16*>(dP_hi), static_cast<__nv_bfloat16*>(dP_lo), dz,
                 N_src, H, F, R, hg, ldp, max_deg};
    return launch_tasks(bwd_src_kernel<4>, a, static_cast<long long>(N_src) * (H / hg), s);
  }
  const int hg = pick_heads_per_warp(H, F, 1);
  if (!hg) return RG_ERR_SHAPE;
  SrcArgs<1> a{P, G, A, alpha, z, t, colptr, csc_slot, csc_dst, csc_rel, dP,
               static_cast<__nv_bfloat16*>(dP_hi), static_cast<__nv_bfloat16*>(dP_lo), dz,
               N_src, H, F, R, hg, ldp, max_deg};
  return launch_tasks(bwd_src_kernel<1>, a, static_cast<long long>(N_src) * (H / hg), s);
}

extern "C" int relgat_layer_bwd_rel(const float* P, long long ldp, const float* dz, const float* hsum,
                                    const int* rel_slot, const int* csr_src, const int* csr_dst,
                                    const int* chunk_lo, const int* chunk_hi, const int* rel_chunk_ptr,
                                    int n_chunks, float* partA, float* partB, float* dA, float* dbeta,
                                    int H, int F, int R, void* stream) {
  if (!P || !dz || !hsum || !rel_chunk_ptr || !dA || n_chunks < 0 || H <= 0 || F <= 0 || R <= 0) return RG_ERR_ARG;
  if (n_chunks > 0 && (!rel_slot || !csr_src || !csr_dst || !chunk_lo || !chunk_hi || !partA || !partB))
    return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (F % 4 == 0 && ldp % 4 == 0 && al16(P) && al16(partA)) {
    const int hg = pick_heads_per_warp(H, F, 4);
    if (!hg) return RG_ERR_SHAPE;
    RelArgs<4> a{P, dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, partA, partB, n_chunks, H, F, hg, ldp};
    rc = launch_tasks(bwd_rel_kernel<4>, a, static_cast<long long>(n_chunks) * (H / hg), s);
  } else {
    const int hg = pick_heads_per_warp(H, F, 1);
    if (!hg) return RG_ERR_SHAPE;
    RelArgs<1> a{P, dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, partA, partB, n_chunks, H, F, hg, ldp};
    rc = launch_tasks(bwd_rel_kernel<1>, a, static_cast<long long>(n_chunks) * (H / hg), s);
  }
  if (rc != RG_OK) return rc;
  const long long total = static_cast<long long>(R) * H * F;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  bwd_rel_reduce_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(partA, partB, rel_chunk_ptr, dA, dbeta, R, H, F);
  return cuda_status(cudaGetLastError());
}
