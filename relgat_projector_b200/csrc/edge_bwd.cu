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
//   bwd_src  : by SOURCE  — dP[i] = sum_{e: src=i} (alpha*G[dst] + dz*A[rel]);  writes dz[e,h];
//              alpha is recomputed from the saved logits and softmax statistics (max, 1/den)
//   bwd_rel  : by RELATION (fixed-size chunks) — partial dA[h,r,:] = sum dz*P[src]; partial dbeta
//   bwd_rel_reduce : ordered sum of the chunk partials
// HBM-bound: per launch bwd_src gathers E*C*s bytes of G rows, bwd_rel gathers E*C*s of P rows.
#include "common.cuh"

namespace relgat {

constexpr int kBwdWarps = 4;
constexpr long long kPrepBlocks = 148 * 4;  // resident blocks of the grid-stride prep kernel (4 x 128 threads per SM)

// ------------------------------------------------------------------------------------
// bwd_prep
// ------------------------------------------------------------------------------------
template <typename TG, int V>
struct PrepArgs {
  const float* dY;     // [N, C] gradient w.r.t. the layer's (activated) output
  const float* out;    // [N, C] forward pre-activation output
  const float* bias;   // [N] forward bias_out
  TG* G;               // [N, C] gradient w.r.t. out, fp32 (may alias dY when apply_elu == 0) or bf16
  float* t;            // [N, H]
  float* hsum;         // [N, H]
  int N, H, F, hg, apply_elu;
  const long long* row_ids;  // optional: only these rows of dY are non-zero (N = their count); t / hsum pre-zeroed
  // feature dropout of the forward (keep bits, nullptr = off): `out` then holds the POST-dropout rows y = out*m*s
  const uint32_t* drop_bits;
  int drop_words;
  float drop_scale;
  __nv_bfloat16* G_export;  // optional bf16 copy of G (what the peers pull on the partitioned path); fp32 G only
  int dy_compact;           // with row_ids: row k of dY belongs to node row_ids[k] (dY holds the listed rows only)
};

template <typename TG, int V>
__global__ void __launch_bounds__(kBwdWarps * 32) bwd_prep_kernel(const PrepArgs<TG, V> a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long n_tasks = static_cast<long long>(a.N) * groups;
  // grid-stride over (row, head-group) tasks: a capped grid of resident warps streams the table (one short-lived block
  // per four rows — 75 k blocks at config 2 — ran at 0.68 of the copy bandwidth)
  for (long long task = static_cast<long long>(blockIdx.x) * kBwdWarps + warp; task < n_tasks;
       task += static_cast<long long>(gridDim.x) * kBwdWarps) {
  const int jt = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(jt) * groups);
  const long long j = a.row_ids ? __ldg(a.row_ids + jt) : jt;
  // in-place dropout scaling must touch a row once: a row named twice (adjacent in the sorted list) is skipped
  if (a.row_ids && a.drop_bits && static_cast<const void*>(a.G) == static_cast<const void*>(a.dY) && jt > 0 &&
      __ldg(a.row_ids + jt - 1) == j)
    continue;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const long long row = j * a.H * a.F + lm.head_off;
  const long long row_y = a.dy_compact ? static_cast<long long>(jt) * a.H * a.F + lm.head_off : row;
  const float b = a.bias ? __ldg(a.bias + j) : 0.f;
  float tt = 0.f, hs = 0.f;
  // every load of the row first (2 x KMAX 128-bit requests in flight per lane), then the arithmetic: with the loads
  // issued pair by pair inside the loop the pass ran at 0.68 of the copy bandwidth (a warp serves one row and leaves)
  float dyv[max_vec<V>()][V], ov[max_vec<V>()][V];
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      RowVec<float, V>::load_stream(a.dY + row_y + q * V, dyv[k]);
      RowVec<float, V>::load_stream(a.out + row + q * V, ov[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k) {
    const int q = lm.sub + lm.lph * k;
    if (q < lm.vph) {
      float ms[V];
      float (&dy)[V] = dyv[k];
      float (&o)[V] = ov[k];
#pragma unroll
      for (int v = 0; v < V; ++v) ms[v] = 1.f;
      if (a.drop_bits) keep_scale<V>(a.drop_bits + j * a.drop_words, lm.head_off + q * V, a.drop_scale, ms);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // ELU'(x) = 1 (x > 0) else exp(x)  (reference model.py:286-287, torch ELU alpha = 1)
        // with dropout: y = out*ms is what was stored and activated; G = d/d out = dy*ELU'(y)*ms and
        // <G, out - b> = sum dy*ELU'(y)*(y - b*ms)
        // (MUFU exp: ncu counted ~1050 warp instructions per row with the library expf — issue-bound, 0.57 of DRAM peak)
        const float gy = a.apply_elu ? (o[v] > 0.f ? dy[v] : dy[v] * fast_exp(o[v])) : dy[v];
        float gg = gy * ms[v];
        if (sizeof(TG) == 2) gg = bf16_round(gg);  // t / hsum consistent with the stored (rounded) G
        dy[v] = gg;
        tt = a.drop_bits ? fmaf(gy, o[v] - b * ms[v], tt) : fmaf(gg, o[v] - b, tt);
        hs += gg;
      }
      if (static_cast<const void*>(a.G) != static_cast<const void*>(a.dY) || a.apply_elu || a.drop_bits)
        RowVec<TG, V>::store(a.G + row + q * V, dy);
      if (a.G_export) store_split_bf16<V>(a.G_export + row + q * V, nullptr, dy);
    }
  }
  tt = head_sum(tt, lm.lph);
  hs = head_sum(hs, lm.lph);
  if (lm.sub == 0) {
    a.t[j * a.H + lm.hh] = tt;
    a.hsum[j * a.H + lm.hh] = hs;
  }
  }
}

// ------------------------------------------------------------------------------------
// bwd_rel: chunk partials of dA and dbeta, then ordered reduce
// ------------------------------------------------------------------------------------
template <typename T, int V>
struct RelArgs {
  const T* P;            // [N_src, C] fp32 or bf16
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

template <typename T, int V>
__global__ void __launch_bounds__(kBwdWarps * 32, 3) bwd_rel_kernel(const RelArgs<T, V> a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = a.H / a.hg;
  const long long task = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  if (task >= static_cast<long long>(a.n_chunks) * groups) return;
  const int c = static_cast<int>(task / groups);
  const int g = static_cast<int>(task - static_cast<long long>(c) * groups);
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int lo = a.chunk_lo[c], hi = a.chunk_hi[c];

  float acc[max_vec<V>()][V];
#pragma unroll
  for (int k = 0; k < max_vec<V>(); ++k)
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
      const T* p0 = a.P + static_cast<long long>(i0) * a.ldp + lm.head_off;
      const T* p1 = a.P + static_cast<long long>(i1) * a.ldp + lm.head_off;
      float x0[max_vec<V>()][V], x1[max_vec<V>()][V];
#pragma unroll
      for (int k = 0; k < max_vec<V>(); ++k) {
        const int q = lm.sub + lm.lph * k;
        if (q < lm.vph) RowVec<T, V>::load_stream(p0 + q * V, x0[k]);
      }
      if (two) {
#pragma unroll
        for (int k = 0; k < max_vec<V>(); ++k) {
          const int q = lm.sub + lm.lph * k;
          if (q < lm.vph) RowVec<T, V>::load_stream(p1 + q * V, x1[k]);
        }
      }
      const float dz0 = __ldg(a.dz + static_cast<long long>(s0) * a.H + lm.hh);
      const float dz1 = two ? __ldg(a.dz + static_cast<long long>(s1) * a.H + lm.hh) : 0.f;
#pragma unroll
      for (int k = 0; k < max_vec<V>(); ++k) {
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
  for (int k = 0; k < max_vec<V>(); ++k) {
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


// dbeta[r] = sum_{e: rel = r} sum_h hsum[dst_e, h] over the by-relation chunks (the part of the by-relation pass that
// does not need P; used when dA comes from the widened dW GEMM).  One warp per chunk, fixed xor-tree per 32 edges,
// chunk partials folded in chunk order by bwd_beta_reduce_kernel: reproducible.
__global__ void __launch_bounds__(kBwdWarps * 32)
bwd_beta_kernel(const float* __restrict__ hsum, const int* __restrict__ rel_slot, const int* __restrict__ csr_dst,
                const int* __restrict__ chunk_lo, const int* __restrict__ chunk_hi, int n_chunks, int H,
                float* __restrict__ partB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kBwdWarps + warp;
  if (c >= n_chunks) return;
  float bacc = 0.f;
  for (int base = chunk_lo[c]; base < chunk_hi[c]; base += 32) {
    float hs = 0.f;
    if (base + lane < chunk_hi[c]) {
      const int jd = __ldg(csr_dst + __ldg(rel_slot + base + lane));
      for (int h = 0; h < H; ++h) hs += __ldg(hsum + static_cast<long long>(jd) * H + h);
    }
    bacc += warp_sum(hs);
  }
  if (lane == 0) partB[c] = bacc;
}

__global__ void bwd_beta_reduce_kernel(const float* __restrict__ partB, const int* __restrict__ rel_chunk_ptr,
                                       float* __restrict__ dbeta, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int c = rel_chunk_ptr[r]; c < rel_chunk_ptr[r + 1]; ++c) s += partB[c];
  dbeta[r] = s;
}

template <typename Args, typename K>
static int launch_tasks(K kernel, const Args& a, long long tasks, cudaStream_t s) {
  if (tasks == 0) return RG_OK;
  const long long blocks = (tasks + kBwdWarps - 1) / kBwdWarps;
  kernel<<<static_cast<unsigned>(blocks), kBwdWarps * 32, 0, s>>>(a);
  return cuda_status(cudaGetLastError());
}

// grid-stride kernels: at most `cap` blocks
template <typename K, typename Args>
static int launch_tasks_capped(K kernel, const Args& a, long long tasks, long long cap, cudaStream_t s) {
  if (tasks == 0) return RG_OK;
  long long blocks = (tasks + kBwdWarps - 1) / kBwdWarps;
  if (blocks > cap) blocks = cap;
  kernel<<<static_cast<unsigned>(blocks), kBwdWarps * 32, 0, s>>>(a);
  return cuda_status(cudaGetLastError());
}

static inline bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace relgat

using namespace relgat;

// by-source pass: kernel and launch ladder live in edge_bwd_src.cuh, instantiated per storage type in
// edge_bwd_src_{f32,f32s,bf16}.cu (parallel compilation)
namespace relgat {
template <typename T, int V>
int run_src(const void* P, long long ldp, const void* G, const float* A, const float* z, const float* minv,
            const float* t, const int* colptr, const int* csc_slot, const int* csc_dst, const int* csc_rel,
            const int4* ch, int n_chunks, const int2* pt, const int* long_node, const int* long_part_ptr,
            int n_long, float* part_acc, float* dP, void* dP_hi, void* dP_lo, float* dz,
            const uint32_t* edge_bits, float edge_scale, const uint32_t* nz_bits, const int* src_row, int p_compact,
            int ds_on, long long ldo, int H, int F, int R, int sm_count, int* work_counter, cudaStream_t s);
extern template int run_src<float, 4>(const void*, long long, const void*, const float*, const float*, const float*,
                                      const float*, const int*, const int*, const int*, const int*, const int4*, int,
                                      const int2*, const int*, const int*, int, float*, float*, void*, void*, float*,
                                      const uint32_t*, float, const uint32_t*, const int*, int, int, long long, int, int, int, int, int*, cudaStream_t);
extern template int run_src<float, 1>(const void*, long long, const void*, const float*, const float*, const float*,
                                      const float*, const int*, const int*, const int*, const int*, const int4*, int,
                                      const int2*, const int*, const int*, int, float*, float*, void*, void*, float*,
                                      const uint32_t*, float, const uint32_t*, const int*, int, int, long long, int, int, int, int, int*, cudaStream_t);
extern template int run_src<__nv_bfloat16, 8>(const void*, long long, const void*, const float*, const float*,
                                              const float*, const float*, const int*, const int*, const int*, const int*,
                                              const int4*, int, const int2*, const int*, const int*, int, float*, float*,
                                              void*, void*, float*, const uint32_t*, float, const uint32_t*, const int*, int, int, long long, int, int, int,
                                              int, int*, cudaStream_t);
}  // namespace relgat

extern "C" int relgat_layer_bwd_prep(const float* dY, const float* out, const float* bias, void* G, int g_is_bf16,
                                     float* t, float* hsum, int N, int H, int F, int apply_elu,
                                     const long long* row_ids, int n_rows,
                                     const unsigned int* drop_bits, int drop_words, float drop_scale,
                                     void* G_export_bf16, int dy_compact, void* stream) {
  if (!dY || !out || !G || !t || !hsum || N < 0 || H <= 0 || F <= 0 || n_rows < 0) return RG_ERR_ARG;
  if (dy_compact && (!row_ids || g_is_bf16 || static_cast<const void*>(G) == static_cast<const void*>(dY))) return RG_ERR_ARG;
  if (drop_bits && drop_words * 32 < H * F) return RG_ERR_ARG;
  if (G_export_bf16 && (g_is_bf16 || F % 4 != 0 || reinterpret_cast<uintptr_t>(G_export_bf16) % 8 != 0)) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rows = N;
  if (row_ids) {
    // sparse dY (the loss touches few rows), no activation: only the listed rows are read; G either aliases dY or is a
    // buffer whose other rows the caller keeps at zero (peer tables); t / hsum of the other rows are zero
    // (compact dY: G is a separate table, so the activation derivative may be applied)
    if ((apply_elu && !dy_compact) || g_is_bf16) return RG_ERR_ARG;
    cudaError_t e = cudaMemsetAsync(t, 0, sizeof(float) * static_cast<size_t>(N) * H, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(hsum, 0, sizeof(float) * static_cast<size_t>(N) * H, s);
    if (e != cudaSuccess) return cuda_status(e);
    rows = n_rows;
  }
  if (g_is_bf16) {
    if (F % 8 != 0) return RG_ERR_SHAPE;
    if (!al16(dY) || !al16(out) || !al16(G)) return RG_ERR_ALIGN;
    const int hg = pick_heads_per_warp(H, F, 8);
    if (!hg) return RG_ERR_SHAPE;
    PrepArgs<__nv_bfloat16, 8> a{dY, out, bias, static_cast<__nv_bfloat16*>(G), t, hsum, rows, H, F, hg, apply_elu, row_ids, drop_bits, drop_words, drop_scale,
                                      static_cast<__nv_bfloat16*>(G_export_bf16), dy_compact};
    return launch_tasks_capped(bwd_prep_kernel<__nv_bfloat16, 8>, a, static_cast<long long>(rows) * (H / hg), kPrepBlocks, s);
  }
  float* Gf = static_cast<float*>(G);
  if (F % 4 == 0 && al16(dY) && al16(out) && al16(G)) {
    const int hg = pick_heads_per_warp(H, F, 4);
    if (!hg) return RG_ERR_SHAPE;
    PrepArgs<float, 4> a{dY, out, bias, Gf, t, hsum, rows, H, F, hg, apply_elu, row_ids, drop_bits, drop_words, drop_scale,
                                      static_cast<__nv_bfloat16*>(G_export_bf16), dy_compact};
    return launch_tasks_capped(bwd_prep_kernel<float, 4>, a, static_cast<long long>(rows) * (H / hg), kPrepBlocks, s);
  }
  const int hg = pick_heads_per_warp(H, F, 1);
  if (!hg) return RG_ERR_SHAPE;
  PrepArgs<float, 1> a{dY, out, bias, Gf, t, hsum, rows, H, F, hg, apply_elu, row_ids, drop_bits, drop_words, drop_scale,
                                      static_cast<__nv_bfloat16*>(G_export_bf16), dy_compact};
  return launch_tasks_capped(bwd_prep_kernel<float, 1>, a, static_cast<long long>(rows) * (H / hg), kPrepBlocks, s);
}

// dP rows of split sources: ordered sum of their parts.
extern "C" int relgat_layer_bwd_src(const void* P, long long ldp, const void* G, int feat_is_bf16, const float* A,
                                    const float* z, const float* minv, const float* t,
                                    const int* colptr, const int* csc_slot, const int* csc_dst, const int* csc_rel,
                                    const int* chunks, int n_chunks, const int* parts, int n_parts,
                                    const int* long_node, const int* long_part_ptr, int n_long, float* part_acc,
                                    float* dP, void* dP_hi, void* dP_lo, float* dz,
                                    const unsigned int* edge_bits, float edge_scale, const unsigned int* dst_nz_bits,
                                    const int* src_row, int p_compact, int want_ds, long long ldo, int H, int F, int R, int sm_count,
                                    int* work_counter,
                                    void* stream) {
  if (!P || !G || !A || !colptr || n_chunks < 0 || n_parts < 0 || n_long < 0 || H <= 0 || F <= 0 || R <= 0)
    return RG_ERR_ARG;
  if (ldo <= 0) ldo = static_cast<long long>(H) * F;
  if (ldo < static_cast<long long>(H) * F + (want_ds ? static_cast<long long>(H) * R : 0)) return RG_ERR_ARG;
  if (n_chunks > 0 && !chunks) return RG_ERR_ARG;
  if (n_parts > 0 && (!parts || !long_node || !long_part_ptr || !part_acc)) return RG_ERR_ARG;
  if (n_chunks == 0) return RG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool ok16 = al16(P) && al16(G) && al16(A) && (!dP || al16(dP)) && (!dP_hi || al16(dP_hi)) &&
                    (!dP_lo || al16(dP_lo)) && (!part_acc || al16(part_acc));
  const int4* ch = reinterpret_cast<const int4*>(chunks);
  const int2* pt = reinterpret_cast<const int2*>(parts);
  if (feat_is_bf16) {
    if (F % 8 != 0 || ldp % 8 != 0 || ldo % 8 != 0) return RG_ERR_SHAPE;
    if (!ok16) return RG_ERR_ALIGN;
    return run_src<__nv_bfloat16, 8>(P, ldp, G, A, z, minv, t, colptr, csc_slot, csc_dst, csc_rel, ch, n_chunks, pt,
                                     long_node, long_part_ptr, n_long, part_acc, dP, dP_hi, dP_lo, dz, edge_bits, edge_scale,
                                     dst_nz_bits, src_row, p_compact, want_ds, ldo, H, F, R, sm_count, work_counter, s);
  }
  if (F % 4 == 0 && ldp % 4 == 0 && ldo % 4 == 0 && ok16)
    return run_src<float, 4>(P, ldp, G, A, z, minv, t, colptr, csc_slot, csc_dst, csc_rel, ch, n_chunks, pt, long_node,
                             long_part_ptr, n_long, part_acc, dP, dP_hi, dP_lo, dz, edge_bits, edge_scale, dst_nz_bits, src_row, p_compact, want_ds, ldo, H, F, R,
                             sm_count, work_counter, s);
  return run_src<float, 1>(P, ldp, G, A, z, minv, t, colptr, csc_slot, csc_dst, csc_rel, ch, n_chunks, pt, long_node,
                           long_part_ptr, n_long, part_acc, dP, dP_hi, dP_lo, dz, edge_bits, edge_scale, dst_nz_bits, src_row, p_compact, want_ds, ldo, H, F, R,
                             sm_count, work_counter, s);
}

template <typename T, int V>
static int run_rel(const void* P, long long ldp, const float* dz, const float* hsum, const int* rel_slot,
                   const int* csr_src, const int* csr_dst, const int* chunk_lo, const int* chunk_hi, int n_chunks,
                   float* partA, float* partB, int H, int F, cudaStream_t s) {
  const int hg = pick_heads_per_warp(H, F, V);
  if (!hg) return RG_ERR_SHAPE;
  RelArgs<T, V> a{static_cast<const T*>(P), dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, partA, partB,
                  n_chunks, H, F, hg, ldp};
  return launch_tasks(bwd_rel_kernel<T, V>, a, static_cast<long long>(n_chunks) * (H / hg), s);
}

extern "C" int relgat_layer_bwd_rel(const void* P, int p_is_bf16, long long ldp, const float* dz, const float* hsum,
                                    const int* rel_slot, const int* csr_src, const int* csr_dst,
                                    const int* chunk_lo, const int* chunk_hi, const int* rel_chunk_ptr,
                                    int n_chunks, float* partA, float* partB, float* dA, float* dbeta,
                                    int H, int F, int R, void* stream) {
  if (!P || !hsum || !rel_chunk_ptr || !dA || n_chunks < 0 || H <= 0 || F <= 0 || R <= 0) return RG_ERR_ARG;
  if (n_chunks > 0 && !dz) return RG_ERR_ARG;
  if (n_chunks > 0 && (!rel_slot || !csr_src || !csr_dst || !chunk_lo || !chunk_hi || !partA || !partB))
    return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (p_is_bf16) {
    if (F % 8 != 0 || ldp % 8 != 0) return RG_ERR_SHAPE;
    if (!al16(P) || !al16(partA)) return RG_ERR_ALIGN;
    rc = run_rel<__nv_bfloat16, 8>(P, ldp, dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, n_chunks, partA,
                                   partB, H, F, s);
  } else if (F % 4 == 0 && ldp % 4 == 0 && al16(P) && al16(partA)) {
    rc = run_rel<float, 4>(P, ldp, dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, n_chunks, partA, partB,
                           H, F, s);
  } else {
    rc = run_rel<float, 1>(P, ldp, dz, hsum, rel_slot, csr_src, csr_dst, chunk_lo, chunk_hi, n_chunks, partA, partB,
                           H, F, s);
  }
  if (rc != RG_OK) return rc;
  const long long total = static_cast<long long>(R) * H * F;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  bwd_rel_reduce_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(partA, partB, rel_chunk_ptr, dA, dbeta, R, H, F);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_layer_bwd_beta(const float* hsum, const int* rel_slot, const int* csr_dst, const int* chunk_lo,
                                     const int* chunk_hi, const int* rel_chunk_ptr, int n_chunks, float* partB,
                                     float* dbeta, int H, int R, void* stream) {
  if (!hsum || !rel_chunk_ptr || !dbeta || n_chunks < 0 || H <= 0 || R <= 0) return RG_ERR_ARG;
  if (n_chunks > 0 && (!rel_slot || !csr_dst || !chunk_lo || !chunk_hi || !partB)) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_chunks > 0) {
    bwd_beta_kernel<<<(n_chunks + kBwdWarps - 1) / kBwdWarps, kBwdWarps * 32, 0, s>>>(hsum, rel_slot, csr_dst, chunk_lo,
                                                                                    chunk_hi, n_chunks, H, partB);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_status(e);
  }
  bwd_beta_reduce_kernel<<<(R + 127) / 128, 128, 0, s>>>(partB, rel_chunk_ptr, dbeta, R);
  return cuda_status(cudaGetLastError());
}
