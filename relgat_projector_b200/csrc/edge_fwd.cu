// E1 — fused RelGAT edge forward (replaces reference layer.py:220-318, ops K2-K11 of
// SURVEY.md §2.2): per-edge logit + LeakyReLU, destination-segmented softmax, attention-
// weighted aggregation and relation bias in ONE pass over the CSR (by-destination) edges.
//
// Work decomposition: the CSR edge array is cut, at destination boundaries, into chunks of
// ~32 edges (graph.py: StreamChunks; destinations with more than 512 in-edges are split into 256-edge
// parts merged by edge_fwd_merge_kernel).  Persistent CTAs (one per SM, 12 warps) own one head-group
// and keep its attention vectors in shared memory; a warp claims one chunk at a time (atomic work
// counter: the tail of a static assignment cost more than the claims) and streams it: the source
// rows P[src] are gathered with 128-bit streaming loads, two rows in flight per warp, and the
// pipeline does not drain at destination boundaries — a destination is "finalised" (normalise,
// add bias, write the row, save the softmax statistics) when the stream crosses into the next
// one.  Short segments (average in-degree ~4.5 on the named graphs) therefore cost no extra
// dependent-load round trips, which is what bounded the first (warp-per-destination) version
// at 39% of HBM peak (profiles/r01_launches_c2_fp32_v1.md).
//
// Softmax is the online (running max / running sum) form; segment order = CSR order =
// original edge order inside a destination (deterministic).
// HBM-bound: algorithmic bytes per launch = E*(C*s + 8) + N*(C*s_out + 12) + E*H*4 (+ planes).
#include "common.cuh"

namespace relgat {

constexpr int kFwdWarps = 12;  // paired variant: 12 warps x <=170 registers, two rows in flight per warp
constexpr int kFwdWarpsSingle = 16;  // single-row variant: 16 warps x <=128 registers (latency hidden by the L2 prefetch)

constexpr int kPrefetchDist = 2;
constexpr int kPrefetchDistSingle = 3;
constexpr int kFwdRowsDefault = 2;  // default number of edges ahead whose source rows are pulled into L2

// warp-cooperative L2 prefetch of one row slice [ptr, ptr + bytes): lane l touches line l
__device__ __forceinline__ void prefetch_l2(const char* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T, int V>
struct FwdArgs {
  const T* P;            // [N_src, H*F] projected features (row stride = ldp elements)
  const float* A;        // [H, R, F] attention vectors
  const float* beta;     // [R] relation bias or nullptr
  const int* rowptr;     // [N+1]
  const int* csr_src;    // [E]
  const int* csr_rel;    // [E]
  const int4* chunks;    // [n_chunks] (first destination, count <= 64, part slot or -1, 0)
  const int2* parts;     // [n_parts] (first edge, end edge) of the parts of split (long) destinations
  float* part_ml;        // [n_parts, H, 2] partial (max, sum)
  float* part_b;         // [n_parts] partial bias sums
  float* part_acc;       // [n_parts, H*F] partial un-normalised accumulators
  float* out;            // [N, H*F] fp32 layer output (pre-activation), may be nullptr
  __nv_bfloat16* act_hi; // [N, H*F] optional bf16 copy of act(out) (hi part)
  __nv_bfloat16* act_lo; // [N, H*F] optional residual (lo part); nullptr = hi only
  float* alpha;          // [E, H] attention weights (CSR order) or nullptr
  float* z;              // [E, H] pre-activation logits (CSR order)
  float* minv;           // [N, H, 2] softmax statistics (max, 1/denominator) saved for backward
  float* bias_out;       // [N] sum of relation biases per destination
  int n_chunks, H, F, R, hg;
  long long ldp;         // row stride of P in elements
  int apply_elu;         // act = ELU (reference model.py:286-287) else identity
  int a_in_smem;         // the head-group's slice of A (hg*R*F floats) is staged in shared memory
  int pf_dist;           // L2 prefetch distance in edges (0 = off)
  int* work_counter;     // zeroed device int: warps claim chunks dynamically (nullptr = static round-robin)
  // dropout (training): keep-bit masks, nullptr = off.  Feature dropout (layer.py:321-322) multiplies the finished
  // row; attention dropout (layer.py:296-297) multiplies alpha AFTER the softmax normalisation.
  const uint32_t* drop_bits;  // [N, drop_words] words
  int drop_words;
  float drop_scale;           // 1 / (1 - p)
  const uint32_t* edge_bits;  // bit index = csr slot * H + head
  float edge_scale;
  // compacted sources (receptive-field step): P holds only the rows the listed destinations read, row src_row[i] for
  // source i (relgat_bitmap_ranks); nullptr = row i
  const int* src_row;
};

// ELU(x) = x (x > 0) else exp(x) - 1.  __expf keeps the absolute error at ~1e-7 (the inputs are
// O(1) activations and the parity budget is 1e-4 of the tensor's max); expm1f costs ~20 instructions
// per element and made the epilogue the largest instruction consumer of the kernel (ncu, round 1).
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : fast_exp(x) - 1.f; }  // NaN stays NaN

// grid = (CTAs per head-group, head-groups); every CTA is persistent and owns one head-group, so
// it stages only that group's attention vectors (ncu on the first version showed the per-edge
// A-row reads missing L1 ~40-70% of the time and doubling the L2->SM traffic).
// DROP: a dropout mask is active (training); false compiles every mask test out of the hot loop
template <typename T, int V, int KV, bool ASM, int NP, int LPHC, bool DROP>
__global__ void __launch_bounds__((NP == 2 ? kFwdWarps : kFwdWarpsSingle) * 32, 1)
edge_fwd_kernel(const FwdArgs<T, V> a) {
  constexpr int kWarps = NP == 2 ? kFwdWarps : kFwdWarpsSingle;
  extern __shared__ __align__(16) float a_sm[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int hl = lm.hh - g * a.hg;  // head index inside the group

  // lane geometry: vectors k < KV-1 are always inside the head, only the last one needs a guard
  // elements between a lane's consecutive vectors; a compile-time constant on the specialised paths
  // (LPHC > 0), which turns every per-vector address into base + immediate
  const int kstride = (LPHC > 0 ? LPHC : lm.lph) * V;
  const int lane_off = lm.head_off + lm.sub * V;       // first element of this lane inside a node row
  const bool last_ok = lm.sub + lm.lph * (KV - 1) < lm.vph;
  const int row_bytes = a.hg * a.F * static_cast<int>(sizeof(T));  // this head-group's slice of a row
  const unsigned long long row_stride_b = static_cast<unsigned long long>(a.ldp) * sizeof(T);
  const char* p_lane = reinterpret_cast<const char*>(a.P + lane_off);              // lane's first vector of row 0
  const char* p_pf = reinterpret_cast<const char*>(a.P + g * a.hg * a.F) + lane * 128;  // lane's prefetch line
  const bool pf_lane_ok = lane * 128 < row_bytes;
#define RG_VALID(k_) ((k_) < KV - 1 || last_ok)

  const float* a_base;  // rows of this lane's head: a_base + r * F
  if (ASM) {
    const float* src = a.A + static_cast<long long>(g) * a.hg * a.R * a.F;
    const int n = a.hg * a.R * a.F;
    if ((n & 3) == 0) {
      for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(a_sm + i) = __ldg(reinterpret_cast<const float4*>(src + i));
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) a_sm[i] = __ldg(src + i);
    }
    __syncthreads();
    a_base = a_sm + hl * a.R * a.F + lm.sub * V;
  } else {
    a_base = a.A + static_cast<long long>(lm.hh) * a.R * a.F + lm.sub * V;
  }

  // chunk claim: dynamic (one atomic per chunk, per head-group counter) keeps all warps busy until the
  // last chunk; static round-robin left a tail of up to one chunk per warp (CE sweep in DESIGN.md)
  int* counter = a.work_counter ? a.work_counter + g : nullptr;
  for (int c = claim_chunk(counter, lane, blockIdx.x * kWarps + warp); c < a.n_chunks;
       c = counter ? claim_chunk(counter, lane, 0) : c + gridDim.x * kWarps) {
    const int4 ch = __ldg(a.chunks + c);
    const int n_lo = ch.x;
    const int nn = ch.y;     // 1..64 destinations
    const int part = ch.z;   // >= 0: this chunk is one part of a split (long) destination
    // rowptr window of the chunk, spread over the lanes (nn + 1 <= 65 entries)
    int rp0 = 0, rp1 = 0, rp2 = 0;
    if (lane <= nn) rp0 = __ldg(a.rowptr + n_lo + lane);
    if (32 + lane <= nn) rp1 = __ldg(a.rowptr + n_lo + 32 + lane);
    if (64 + lane <= nn) rp2 = __ldg(a.rowptr + n_lo + 64 + lane);
#define RG_RP(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, rp0, (k_) & 31)       \
                             : ((k_) < 64 ? __shfl_sync(0xffffffffu, rp1, (k_) & 31) \
                                          : __shfl_sync(0xffffffffu, rp2, (k_) & 31)))
    int e_lo = RG_RP(0);
    int e_hi = RG_RP(nn);
    if (part >= 0) {  // sub-range of the destination's segment
      const int2 pe = __ldg(a.parts + part);
      e_lo = pe.x;
      e_hi = pe.y;
    }

    float acc[KV][V];
#pragma unroll
    for (int k = 0; k < KV; ++k)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[k][v] = 0.f;
    float m = -INFINITY, l = 0.f, bsum = 0.f;
    int kn = 0;  // destination cursor inside the chunk
    int seg_start = e_lo;
    int seg_end = part >= 0 ? e_hi : RG_RP(1);
    int base = e_lo - 32;  // (src, rel) window [base, base + 32) held across the lanes
    int my_src = 0, my_rel = 0;
    float my_beta = 0.f;

    // one edge: online softmax update + weighted accumulate (logit d_ already reduced)
#define RG_EDGE(x_, d_, b_, e_)                                                                    \
  {                                                                                                \
    if (lm.sub == 0) a.z[static_cast<long long>(e_) * a.H + lm.hh] = (d_);                         \
    const float ev = (d_) > 0.f ? (d_) : kLeakySlope * (d_);                                       \
    const float mn = fmaxf(m, ev);                                                                 \
    const float sc = __expf(m - mn);                                                               \
    const float w0 = __expf(ev - mn);                                                              \
    l = fmaf(l, sc, w0); /* the denominator counts every edge; attention dropout scales the kept terms */ \
    const float w = (DROP && a.edge_bits) ? w0 * keep_scale1(a.edge_bits, static_cast<long long>(e_) * a.H + lm.hh, a.edge_scale) : w0; \
    _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                               \
      _Pragma("unroll") for (int v = 0; v < V; ++v) acc[k][v] = fmaf(acc[k][v], sc, w * x_[k][v]); \
    }                                                                                              \
    m = (ev != ev) ? ev : mn; /* NaN logits poison the row like the reference does */              \
    bsum += (b_);                                                                                  \
  }

    int e = e_lo;
    while (true) {
      if (e < e_hi && e >= base + 32) {
        base = e;
        const int idx = base + lane;
        if (idx < e_hi) {
          my_src = __ldg(a.csr_src + idx);
          if (a.src_row) my_src = __ldg(a.src_row + my_src);
          my_rel = __ldg(a.csr_rel + idx);
          my_beta = a.beta ? __ldg(a.beta + my_rel) : 0.f;
        }
        // warm L2 with the first rows of the new window; the rolling prefetch below keeps
        // kPrefetchDist rows ahead, so the gathers mostly pay L2 instead of HBM latency
        for (int pf = 0; pf < a.pf_dist; ++pf) {
          const int ip = __shfl_sync(0xffffffffu, my_src, pf);
          if (base + pf < e_hi && pf_lane_ok) prefetch_l2(p_pf + static_cast<unsigned long long>(ip) * row_stride_b);
        }
      }
      const int npair = min(NP, min(e_hi - e, base + 32 - e));  // 0 only when the chunk is exhausted
      float x0[KV][V], x1[KV][V];
      float d0 = 0.f, d1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { x0[KV - 1][v] = 0.f; x1[KV - 1][v] = 0.f; }  // the only guarded vector
      if (npair > 0) {
        const bool two = NP == 2 && npair == 2;
        const int t = e - base;
        const int i0 = __shfl_sync(0xffffffffu, my_src, t);
        const int r0 = __shfl_sync(0xffffffffu, my_rel, t);
        b0 = __shfl_sync(0xffffffffu, my_beta, t);
        const int i1 = __shfl_sync(0xffffffffu, my_src, two ? t + 1 : t);
        const int r1 = __shfl_sync(0xffffffffu, my_rel, two ? t + 1 : t);
        b1 = __shfl_sync(0xffffffffu, my_beta, two ? t + 1 : t);
        const T* p0 = reinterpret_cast<const T*>(p_lane + static_cast<unsigned long long>(i0) * row_stride_b);
        const T* p1 = reinterpret_cast<const T*>(p_lane + static_cast<unsigned long long>(i1) * row_stride_b);
#pragma unroll
        for (int pf = 0; pf < NP; ++pf) {  // rolling L2 prefetch, pf_dist edges ahead (same window)
          const int tp = t + a.pf_dist + pf;
          const int ip = __shfl_sync(0xffffffffu, my_src, tp & 31);
          if (pf_lane_ok && a.pf_dist > 0 && tp < 32 && base + tp < e_hi)
            prefetch_l2(p_pf + static_cast<unsigned long long>(ip) * row_stride_b);
        }
        // issue both row gathers before any arithmetic (two rows in flight per warp)
#pragma unroll
        for (int k = 0; k < KV; ++k)
          if (RG_VALID(k)) RowVec<T, V>::load_stream(p0 + k * kstride, x0[k]);
        if (two) {
#pragma unroll
          for (int k = 0; k < KV; ++k)
            if (RG_VALID(k)) RowVec<T, V>::load_stream(p1 + k * kstride, x1[k]);
        }
        const float* a0 = a_base + r0 * a.F;
        const float* a1 = a_base + r1 * a.F;
        // V independent partial sums per edge: the logit's FMA chain is KV long instead of KV*V
        float s0[V], s1[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { s0[v] = 0.f; s1[v] = 0.f; }
#pragma unroll
        for (int k = 0; k < KV; ++k) {
          if (RG_VALID(k)) {
            float av[V];
            if (ASM) RowVec<float, V>::load_shared(a0 + k * kstride, av); else RowVec<float, V>::load_cached(a0 + k * kstride, av);
#pragma unroll
            for (int v = 0; v < V; ++v) s0[v] = fmaf(x0[k][v], av[v], s0[v]);
            if (two) {
              if (ASM) RowVec<float, V>::load_shared(a1 + k * kstride, av); else RowVec<float, V>::load_cached(a1 + k * kstride, av);
#pragma unroll
              for (int v = 0; v < V; ++v) s1[v] = fmaf(x1[k][v], av[v], s1[v]);
            }
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { d0 += s0[v]; d1 += s1[v]; }
        if constexpr (LPHC > 0) { d0 = head_sum_c<LPHC>(d0); d1 = head_sum_c<LPHC>(d1); }
        else { d0 = head_sum(d0, lm.lph); d1 = head_sum(d1, lm.lph); }
      }
      // consume the pair; destinations are finalised (ONE code site) whenever the edge cursor
      // reaches the end of the current segment, including empty and trailing destinations
      int u = 0;
      while (true) {
        const int cur = e + u;
        if (kn < nn && cur == seg_end && part >= 0) {
          // split destination: park the running (max, sum, bias, accumulator) for the merge kernel
          const long long prow = static_cast<long long>(part) * C + lane_off;
#pragma unroll
          for (int k = 0; k < KV; ++k)
            if (RG_VALID(k)) RowVec<float, V>::store(a.part_acc + prow + k * kstride, acc[k]);
          if (lm.sub == 0)
            *reinterpret_cast<float2*>(a.part_ml + (static_cast<long long>(part) * a.H + lm.hh) * 2) = make_float2(m, l);
          if (g == 0 && lane == 0) a.part_b[part] = bsum;
          ++kn;
          continue;
        }
        if (kn < nn && cur == seg_end) {
          const int j = n_lo + kn;
          const bool empty = (seg_end == seg_start);
          const float inv = empty ? 0.f : 1.f / fmaxf(l, 1e-16f);  // reference layer.py:291 clamp
          if (empty) bsum = 0.f;  // acc is 0 too, so the row below comes out exactly 0
          const long long row_off = static_cast<long long>(j) * C + lane_off;
#pragma unroll
          for (int k = 0; k < KV; ++k) {
            if (RG_VALID(k)) {
              float o[V];
#pragma unroll
              for (int v = 0; v < V; ++v) {
                o[v] = fmaf(acc[k][v], inv, bsum);  // bias on every head/channel, :313-318
                acc[k][v] = 0.f;
              }
              if (DROP && a.drop_bits) {  // feature dropout on the finished row (layer.py:321-322), before the activation
                float ms[V];
                keep_scale<V>(a.drop_bits + static_cast<long long>(j) * a.drop_words, lane_off + k * kstride, a.drop_scale, ms);
#pragma unroll
                for (int v = 0; v < V; ++v) o[v] *= ms[v];
              }
              const long long off = row_off + k * kstride;
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
          if (g == 0 && lane == 0 && a.bias_out) a.bias_out[j] = bsum;
          if (lm.sub == 0 && a.minv) {
            *reinterpret_cast<float2*>(a.minv + (static_cast<long long>(j) * a.H + lm.hh) * 2) =
                make_float2(empty ? 0.f : m, inv);
          }
          if (a.alpha) {  // alpha = exp(eps - m) / den for every (edge, head of this group)
            __syncwarp();
            const int items = (seg_end - seg_start) * a.hg;
            for (int it = 0; it < items; it += 32) {
              const int idx = it + lane;
              const int hgi = idx % a.hg;
              const float mh = __shfl_sync(0xffffffffu, m, hgi * lm.lph);
              const float ih = __shfl_sync(0xffffffffu, inv, hgi * lm.lph);
              if (idx < items) {
                const long long o = static_cast<long long>(seg_start + idx / a.hg) * a.H + g * a.hg + hgi;
                const float zz = a.z[o];
                const float ee = zz > 0.f ? zz : kLeakySlope * zz;
                a.alpha[o] = __expf(ee - mh) * ih * (a.edge_bits ? keep_scale1(a.edge_bits, o, a.edge_scale) : 1.f);
              }
            }
          }
          m = -INFINITY; l = 0.f; bsum = 0.f;
          ++kn;
          seg_start = seg_end;
          if (kn < nn) seg_end = RG_RP(kn + 1);
          continue;
        }
        if (u == npair) break;
        if (u == 0) RG_EDGE(x0, d0, b0, cur) else RG_EDGE(x1, d1, b1, cur);
        ++u;
      }
      e += npair;
      if (e >= e_hi) break;
    }
#undef RG_EDGE
#undef RG_RP
  }
#undef RG_VALID
}

// Combines the parts of split destinations (flash-decoding style) and writes the same outputs as
// the main kernel's epilogue.  One CTA of kMergeWarps warps per (split destination, head-group):
// warp w folds parts p_lo+w, p_lo+w+kMergeWarps, ... in ascending order, the per-warp states are
// then folded in warp order — a fixed reduction tree, so results are reproducible.
constexpr int kMergeWarps = 8;

template <typename T, int V>
__global__ void __launch_bounds__(kMergeWarps * 32)
edge_fwd_merge_kernel(const FwdArgs<T, V> a, const int* __restrict__ long_node,
                      const int* __restrict__ long_part_ptr, int n_long) {
  __shared__ __align__(16) float sm_acc[kMergeWarps][max_vec<V>() * 32 * V];
  __shared__ float sm_m[kMergeWarps][32], sm_l[kMergeWarps][32], sm_b[kMergeWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const int task = blockIdx.x;
  if (task >= n_long * groups) return;
  const int li = task / groups, g = task - li * groups;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int j = long_node[li];
  const int p_lo = long_part_ptr[li], p_hi = long_part_ptr[li + 1];

  // pass 1 (per warp): online merge of this warp's parts
  float M = -INFINITY, L = 0.f, bsum = 0.f;
  float acc[max_vec<V>()][V];
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[k][v] = 0.f;
  for (int p = p_lo + warp; p < p_hi; p += kMergeWarps) {
    const float2 ml = *reinterpret_cast<const float2*>(a.part_ml + (static_cast<long long>(p) * a.H + lm.hh) * 2);
    const float mn = (ml.x != ml.x) ? ml.x : fmaxf(M, ml.x);
    const float s_old = __expf(M - mn), s_new = __expf(ml.x - mn);
    L = fmaf(L, s_old, ml.y * s_new);
    bsum += a.part_b[p];
#pragma unroll
    for (int k = 0; k < max_vec<V>(); ++k) {
      const int q = lm.sub + lm.lph * k;
      if (q < lm.vph) {
        float x[V];
        RowVec<float, V>::load_cached(a.part_acc + static_cast<long long>(p) * C + lm.head_off + q * V, x);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[k][v] = fmaf(acc[k][v], s_old, x[v] * s_new);
      }
    }
    M = mn;
  }
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k) RowVec<float, V>::store(&sm_acc[warp][(k * 32 + lane) * V], acc[k]);
  sm_m[warp][lane] = M;
  sm_l[warp][lane] = L;
  if (lane == 0) sm_b[warp] = bsum;
  __syncthreads();
  if (warp != 0) return;

  // pass 2 (warp 0): fold the per-warp states in warp order
  const int nw = min(kMergeWarps, p_hi - p_lo);
  for (int w = 1; w < nw; ++w) {
    const float mw = sm_m[w][lane], lw = sm_l[w][lane];
    const float mn = (mw != mw) ? mw : fmaxf(M, mw);
    const float s_old = __expf(M - mn), s_new = __expf(mw - mn);
    L = fmaf(L, s_old, lw * s_new);
    bsum += sm_b[w];
#pragma unroll
    for (int k = 0; k < max_vec<V>(); ++k) {
      float x[V];
      RowVec<float, V>::load_shared(&sm_acc[w][(k * 32 + lane) * V], x);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[k][v] = fmaf(acc[k][v], s_old, x[v] * s_new);
    }
    M = mn;
  }
  const float inv = 1.f / fmaxf(L, 1e-16f);
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      float o[V];
#pragma unroll
      for (int v = 0; v < V; ++v) o[v] = fmaf(acc[k][v], inv, bsum);
      if (a.drop_bits) {
        float ms[V];
        keep_scale<V>(a.drop_bits + static_cast<long long>(j) * a.drop_words, lm.head_off + q * V, a.drop_scale, ms);
#pragma unroll
        for (int v = 0; v < V; ++v) o[v] *= ms[v];
      }
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
  if (g == 0 && lane == 0 && a.bias_out) a.bias_out[j] = bsum;
  if (lm.sub == 0 && a.minv)
    *reinterpret_cast<float2*>(a.minv + (static_cast<long long>(j) * a.H + lm.hh) * 2) = make_float2(M, inv);
  if (a.alpha) {
    const int e0 = a.rowptr[j], e1 = a.rowptr[j + 1];
    const int items = (e1 - e0) * a.hg;
    for (int it = 0; it < items; it += 32) {
      const int idx = it + lane;
      const int hgi = idx % a.hg;
      const float mh = __shfl_sync(0xffffffffu, M, hgi * lm.lph);
      const float ih = __shfl_sync(0xffffffffu, inv, hgi * lm.lph);
      if (idx < items) {
        const long long o = static_cast<long long>(e0 + idx / a.hg) * a.H + g * a.hg + hgi;
        const float zz = a.z[o];
        const float ee = zz > 0.f ? zz : kLeakySlope * zz;
        a.alpha[o] = __expf(ee - mh) * ih * (a.edge_bits ? keep_scale1(a.edge_bits, o, a.edge_scale) : 1.f);
      }
    }
  }
}

template <typename T, int V, int KV, int NP, int LPHC, bool DROP>
static int launch_fwd_drop(FwdArgs<T, V> a, int sm_count, cudaStream_t stream) {
  constexpr int kWarps = NP == 2 ? kFwdWarps : kFwdWarpsSingle;
  const int groups = a.H / a.hg;
  if (sm_count <= 0) sm_count = 148;
  int ctas = sm_count / groups;
  if (ctas < 1) ctas = 1;
  const int need = (a.n_chunks + kWarps - 1) / kWarps;
  if (ctas > need) ctas = need;
  const size_t a_bytes = static_cast<size_t>(a.hg) * a.R * a.F * sizeof(float);
  a.a_in_smem = a_bytes <= kSmemBudgetA ? 1 : 0;
  {
    const char* pv = getenv("RELGAT_PF_DIST");
    a.pf_dist = pv ? atoi(pv) : (NP == 2 ? kPrefetchDist : kPrefetchDistSingle);
    if (a.pf_dist < 0) a.pf_dist = 0;
    if (a.pf_dist > 30) a.pf_dist = 30;
  }
  if (a.a_in_smem) {
    cudaError_t e = cudaFuncSetAttribute(edge_fwd_kernel<T, V, KV, true, NP, LPHC, DROP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudgetA));
    if (e != cudaSuccess) return cuda_status(e);
    edge_fwd_kernel<T, V, KV, true, NP, LPHC, DROP><<<dim3(ctas, groups), kWarps * 32, a_bytes, stream>>>(a);
  } else {
    edge_fwd_kernel<T, V, KV, false, NP, LPHC, DROP><<<dim3(ctas, groups), kWarps * 32, 0, stream>>>(a);
  }
  return cuda_status(cudaGetLastError());
}

template <typename T, int V, int KV, int NP, int LPHC>
static int launch_fwd_np(const FwdArgs<T, V>& a, int sm_count, cudaStream_t stream) {
  if (a.drop_bits || a.edge_bits) return launch_fwd_drop<T, V, KV, NP, LPHC, true>(a, sm_count, stream);
  return launch_fwd_drop<T, V, KV, NP, LPHC, false>(a, sm_count, stream);
}

template <typename T, int V, int KV>
static int launch_fwd_kv(const FwdArgs<T, V>& a, int sm_count, cudaStream_t stream) {
  const char* v = getenv("RELGAT_FWD_ROWS");  // experiment knob: rows in flight per warp (1 or 2)
  const int np = v ? atoi(v) : kFwdRowsDefault;
  const int lph = 32 / a.hg;
  // compile-time lane stride for the shapes of the named configs (F = 200: 4 heads per warp in fp32
  // and bf16, or one head per warp when R is large); everything else takes the generic path
  constexpr bool kSpec8 = (V == 4 && KV == 7) || (V == 8 && KV == 4);
  constexpr bool kSpec32 = (V == 4 && KV == 2);
  if (np == 1) return launch_fwd_np<T, V, KV, 1, 0>(a, sm_count, stream);
  if (kSpec8 && lph == 8) return launch_fwd_np<T, V, KV, 2, kSpec8 ? 8 : 0>(a, sm_count, stream);
  if (kSpec32 && lph == 32) return launch_fwd_np<T, V, KV, 2, kSpec32 ? 32 : 0>(a, sm_count, stream);
  return launch_fwd_np<T, V, KV, 2, 0>(a, sm_count, stream);
}

// KV = 128-bit vectors per lane: specialised so unused accumulator registers are not allocated
template <typename T, int V>
static int launch_fwd(const FwdArgs<T, V>& a, int sm_count, cudaStream_t stream) {
  if (a.n_chunks == 0) return RG_OK;
  const int kv = vectors_per_lane(a.F / V, a.hg);
  switch (kv) {  // exact count: only the last vector of a lane needs a bounds guard
    case 1: return launch_fwd_kv<T, V, 1>(a, sm_count, stream);
    case 2: return launch_fwd_kv<T, V, 2>(a, sm_count, stream);
    case 3: return launch_fwd_kv<T, V, 3>(a, sm_count, stream);
    case 4: return launch_fwd_kv<T, V, 4>(a, sm_count, stream);
    default: break;
  }
  if constexpr (V != 8) {
    switch (kv) {
      case 5: return launch_fwd_kv<T, V, 5>(a, sm_count, stream);
      case 6: return launch_fwd_kv<T, V, 6>(a, sm_count, stream);
      case 7: return launch_fwd_kv<T, V, 7>(a, sm_count, stream);
      default: return launch_fwd_kv<T, V, 8>(a, sm_count, stream);
    }
  }
  return RG_ERR_SHAPE;
}

}  // namespace relgat

using namespace relgat;

template <typename T, int V>
static int run_fwd(const void* P, long long ldp, const float* A, const float* beta, const int* rowptr,
                   const int* csr_src, const int* csr_rel, const int4* ch, int n_chunks, const int2* pt,
                   const int* long_node, const int* long_part_ptr, int n_long, float* part_ml, float* part_b,
                   float* part_acc, float* out, void* act_hi, void* act_lo, int apply_elu, float* alpha, float* z,
                   float* minv, float* bias_out, const uint32_t* drop_bits, int drop_words, float drop_scale,
                   const uint32_t* edge_bits, float edge_scale, const int* src_row, int H, int F, int R, int sm_count,
                   int* work_counter, cudaStream_t s) {
  const int hg = pick_heads_per_warp(H, F, V, R, smem_budget_override("RELGAT_FWD_BUDGET_KB", kSmemBudgetA));
  if (!hg) return RG_ERR_SHAPE;
  if (H / hg > 32) return RG_ERR_SHAPE;  // work_counter holds 32 ints (one per head-group)
  if (work_counter) {
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int) * (H / hg), s);
    if (e != cudaSuccess) return cuda_status(e);
  }
  FwdArgs<T, V> a{static_cast<const T*>(P), A, beta, rowptr, csr_src, csr_rel, ch, pt, part_ml, part_b,
                  part_acc, out, static_cast<__nv_bfloat16*>(act_hi), static_cast<__nv_bfloat16*>(act_lo),
                  alpha, z, minv, bias_out, n_chunks, H, F, R, hg, ldp, apply_elu, 0, 0, work_counter,
                  drop_bits, drop_words, drop_scale, edge_bits, edge_scale, src_row};
  int rc = launch_fwd(a, sm_count, s);
  if (rc != RG_OK || n_long == 0) return rc;
  const int tasks = n_long * (H / hg);
  edge_fwd_merge_kernel<T, V><<<tasks, kMergeWarps * 32, 0, s>>>(a, long_node, long_part_ptr, n_long);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_layer_fwd(
    const void* P, int p_is_bf16, long long ldp, const float* A, const float* beta,
    const int* rowptr, const int* csr_src, const int* csr_rel,
    const int* chunks, int n_chunks, const int* parts, int n_parts,
    const int* long_node, const int* long_part_ptr, int n_long,
    float* part_ml, float* part_b, float* part_acc,
    float* out, void* act_hi, void* act_lo, int apply_elu,
    float* alpha, float* z, float* minv, float* bias_out,
    const unsigned int* drop_bits, int drop_words, float drop_scale, const unsigned int* edge_bits, float edge_scale,
    const int* src_row, int H, int F, int R, int sm_count, int* work_counter, void* stream) {
  if (!P || !A || !rowptr || n_chunks < 0 || n_parts < 0 || n_long < 0 || H <= 0 || F <= 0 || R <= 0) return RG_ERR_ARG;
  if (drop_bits && drop_words * 32 < H * F) return RG_ERR_ARG;
  if (n_chunks > 0 && !chunks) return RG_ERR_ARG;
  if (n_parts > 0 && (!parts || !long_node || !long_part_ptr || !part_ml || !part_b || !part_acc)) return RG_ERR_ARG;
  if (n_chunks == 0) return RG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto al = [](const void* p, int n) { return !p || reinterpret_cast<uintptr_t>(p) % n == 0; };
  const bool vec_ok = al(P, 16) && al(A, 16) && al(out, 16) && al(act_hi, 16) && al(act_lo, 16) && al(part_acc, 16);
  const int4* ch = reinterpret_cast<const int4*>(chunks);
  const int2* pt = reinterpret_cast<const int2*>(parts);
  if (p_is_bf16) {  // bf16 feature storage: 8-element vectors
    if (F % 8 != 0 || ldp % 8 != 0) return RG_ERR_SHAPE;
    if (!vec_ok) return RG_ERR_ALIGN;
    return run_fwd<__nv_bfloat16, 8>(P, ldp, A, beta, rowptr, csr_src, csr_rel, ch, n_chunks, pt, long_node,
                                     long_part_ptr, n_long, part_ml, part_b, part_acc, out, act_hi, act_lo, apply_elu,
                                     alpha, z, minv, bias_out, drop_bits, drop_words, drop_scale, edge_bits, edge_scale, src_row, H, F, R, sm_count,
                                     work_counter, s);
  }
  if ((F % 4 == 0) && (ldp % 4 == 0) && vec_ok)
    return run_fwd<float, 4>(P, ldp, A, beta, rowptr, csr_src, csr_rel, ch, n_chunks, pt, long_node, long_part_ptr,
                             n_long, part_ml, part_b, part_acc, out, act_hi, act_lo, apply_elu, alpha, z, minv,
                             bias_out, drop_bits, drop_words, drop_scale, edge_bits, edge_scale, src_row, H, F, R, sm_count,
                             work_counter, s);
  return run_fwd<float, 1>(P, ldp, A, beta, rowptr, csr_src, csr_rel, ch, n_chunks, pt, long_node, long_part_ptr,
                           n_long, part_ml, part_b, part_acc, out, act_hi, act_lo, apply_elu, alpha, z, minv,
                           bias_out, drop_bits, drop_words, drop_scale, edge_bits, edge_scale, src_row, H, F, R, sm_count,
                           work_counter, s);
}
