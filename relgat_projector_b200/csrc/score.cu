// S1 — fused gather-score kernels for the KG scorers (replaces reference core/scorer.py:58-94,
// 154-201 plus the row gathers of model.py:136-137, ops K14-K16 of SURVEY.md §2.2).
//
//   DistMult: score = sum_d s*r*t           transform = s (.) r
//   TransE  : score = -|| n(s) + n(r) - n(t) ||_2 with n(v) = v / max(||v||, 1e-12) when
//             normalize != 0 (model.py:93-95 always asks for it); transform = n(s) + n(r)
//
// One warp per triple; source/destination rows are read through optional index vectors so
// x[src_ids] / x[dst_ids] are never materialised unless the caller asks for the gathered
// copies.  The backward writes one gradient row per triple and role; the ordered
// (deterministic) reduction of those rows per node / per relation is index_add_sorted.
#include "common.cuh"

namespace relgat {

constexpr int kScoreWarps = 4;
constexpr float kNormEps = 1e-12f;  // torch.nn.functional.normalize default eps

enum : int { SCORER_DISTMULT = 0, SCORER_TRANSE = 1 };

struct ScoreArgs {
  const float* xs;         // source matrix  [*, D]
  const float* xd;         // destination matrix [*, D]
  const long long* src_idx;  // [B] or nullptr (row b)
  const long long* dst_idx;  // [B] or nullptr
  const float* rel_emb;    // [R, D]
  const long long* rel_ids;  // [B]
  int B, D, kind, normalize;
  // forward outputs
  float* score;            // [B]
  float* transform;        // [Bt, D] or nullptr
  int Bt;
  float* src_vec;          // [B, D] gathered copies or nullptr
  float* dst_vec;          // [B, D] or nullptr
  // backward inputs / outputs
  const float* dscore;     // [B] or nullptr
  const float* dtransform; // [Bt, D] or nullptr
  float* d_src;            // [B, D]
  float* d_dst;            // [B, D]
  float* d_rel;            // [B, D]
};

template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&v)[V]) { RowVec<float, V>::load_cached(p, v); }

template <int V>
__global__ void __launch_bounds__(kScoreWarps * 32) score_fwd_kernel(const ScoreArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kScoreWarps + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const long long si = a.src_idx ? a.src_idx[b] : b;
  const long long di = a.dst_idx ? a.dst_idx[b] : b;
  const float* s = a.xs + si * a.D;
  const float* t = a.xd + di * a.D;
  const float* r = a.rel_emb + a.rel_ids[b] * a.D;
  const bool want_tr = a.transform && b < a.Bt;
  if (a.kind == SCORER_DISTMULT) {
    float acc = 0.f;
    for (int c = lane * V; c < a.D; c += 32 * V) {
      float sv[V], rv[V], tv[V];
      ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
      float tr[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { tr[v] = sv[v] * rv[v]; acc = fmaf(tr[v], tv[v], acc); }
      if (want_tr) RowVec<float, V>::store(a.transform + static_cast<long long>(b) * a.D + c, tr);
      if (a.src_vec) RowVec<float, V>::store(a.src_vec + static_cast<long long>(b) * a.D + c, sv);
      if (a.dst_vec) RowVec<float, V>::store(a.dst_vec + static_cast<long long>(b) * a.D + c, tv);
    }
    acc = warp_sum(acc);
    if (lane == 0) a.score[b] = acc;
    return;
  }
  // TransE
  float is = 1.f, ir = 1.f, it = 1.f;
  if (a.normalize) {
    float ns = 0.f, nr = 0.f, nt = 0.f;
    for (int c = lane * V; c < a.D; c += 32 * V) {
      float sv[V], rv[V], tv[V];
      ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
      for (int v = 0; v < V; ++v) { ns = fmaf(sv[v], sv[v], ns); nr = fmaf(rv[v], rv[v], nr); nt = fmaf(tv[v], tv[v], nt); }
    }
    is = 1.f / fmaxf(sqrtf(warp_sum(ns)), kNormEps);
    ir = 1.f / fmaxf(sqrtf(warp_sum(nr)), kNormEps);
    it = 1.f / fmaxf(sqrtf(warp_sum(nt)), kNormEps);
  }
  float acc = 0.f;
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float sv[V], rv[V], tv[V], tr[V];
    ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      tr[v] = sv[v] * is + rv[v] * ir;
      const float u = tr[v] - tv[v] * it;
      acc = fmaf(u, u, acc);
    }
    if (want_tr) RowVec<float, V>::store(a.transform + static_cast<long long>(b) * a.D + c, tr);
    if (a.src_vec) RowVec<float, V>::store(a.src_vec + static_cast<long long>(b) * a.D + c, sv);
    if (a.dst_vec) RowVec<float, V>::store(a.dst_vec + static_cast<long long>(b) * a.D + c, tv);
  }
  acc = warp_sum(acc);
  if (lane == 0) a.score[b] = -sqrtf(acc);
}

template <int V>
__global__ void __launch_bounds__(kScoreWarps * 32) score_bwd_kernel(const ScoreArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kScoreWarps + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const long long si = a.src_idx ? a.src_idx[b] : b;
  const long long di = a.dst_idx ? a.dst_idx[b] : b;
  const float* s = a.xs + si * a.D;
  const float* t = a.xd + di * a.D;
  const float* r = a.rel_emb + a.rel_ids[b] * a.D;
  const float gs = a.dscore ? a.dscore[b] : 0.f;
  const float* gtr = (a.dtransform && b < a.Bt) ? a.dtransform + static_cast<long long>(b) * a.D : nullptr;
  const long long row = static_cast<long long>(b) * a.D;
  if (a.kind == SCORER_DISTMULT) {
    for (int c = lane * V; c < a.D; c += 32 * V) {
      float sv[V], rv[V], tv[V], gv[V], o1[V], o2[V], o3[V];
      ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
      for (int v = 0; v < V; ++v) gv[v] = 0.f;
      if (gtr) ldv<V>(gtr + c, gv);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        o1[v] = gs * rv[v] * tv[v] + gv[v] * rv[v];  // d/ds
        o2[v] = gs * sv[v] * rv[v];                  // d/dt
        o3[v] = gs * sv[v] * tv[v] + gv[v] * sv[v];  // d/dr
      }
      if (a.d_src) RowVec<float, V>::store(a.d_src + row + c, o1);
      if (a.d_dst) RowVec<float, V>::store(a.d_dst + row + c, o2);
      if (a.d_rel) RowVec<float, V>::store(a.d_rel + row + c, o3);
    }
    return;
  }
  // TransE: pass 1 norms, pass 2 distance and the projections needed by normalize-backward
  float ns = 0.f, nr = 0.f, nt = 0.f;
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float sv[V], rv[V], tv[V];
    ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) { ns = fmaf(sv[v], sv[v], ns); nr = fmaf(rv[v], rv[v], nr); nt = fmaf(tv[v], tv[v], nt); }
  }
  ns = sqrtf(warp_sum(ns)); nr = sqrtf(warp_sum(nr)); nt = sqrtf(warp_sum(nt));
  float is = 1.f, ir = 1.f, it = 1.f;
  if (a.normalize) { is = 1.f / fmaxf(ns, kNormEps); ir = 1.f / fmaxf(nr, kNormEps); it = 1.f / fmaxf(nt, kNormEps); }
  float d2 = 0.f;
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float sv[V], rv[V], tv[V];
    ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) { const float u = sv[v] * is + rv[v] * ir - tv[v] * it; d2 = fmaf(u, u, d2); }
  }
  const float dist = sqrtf(warp_sum(d2));
  const float coef = dist > 0.f ? -gs / dist : 0.f;  // d(-dist)/du = -u/dist (0 at dist == 0 like torch.norm)
  // <v_hat, dv_hat> for the three normalisations
  float ps = 0.f, pr = 0.f, pt = 0.f;
  if (a.normalize) {
    for (int c = lane * V; c < a.D; c += 32 * V) {
      float sv[V], rv[V], tv[V], gv[V];
      ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
      for (int v = 0; v < V; ++v) gv[v] = 0.f;
      if (gtr) ldv<V>(gtr + c, gv);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float u = sv[v] * is + rv[v] * ir - tv[v] * it;
        const float dsh = coef * u + gv[v];
        ps = fmaf(sv[v] * is, dsh, ps);
        pr = fmaf(rv[v] * ir, dsh, pr);
        pt = fmaf(tv[v] * it, -coef * u, pt);
      }
    }
    ps = warp_sum(ps); pr = warp_sum(pr); pt = warp_sum(pt);
  }
  const bool fs = a.normalize && ns > kNormEps, fr = a.normalize && nr > kNormEps, ft = a.normalize && nt > kNormEps;
  for (int c = lane * V; c < a.D; c += 32 * V) {
    float sv[V], rv[V], tv[V], gv[V], o1[V], o2[V], o3[V];
    ldv<V>(s + c, sv); ldv<V>(r + c, rv); ldv<V>(t + c, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) gv[v] = 0.f;
    if (gtr) ldv<V>(gtr + c, gv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float u = sv[v] * is + rv[v] * ir - tv[v] * it;
      const float dsh = coef * u + gv[v];   // grad wrt s_hat and r_hat
      const float dth = -coef * u;          // grad wrt t_hat
      o1[v] = (dsh - (fs ? sv[v] * is * ps : 0.f)) * is;
      o3[v] = (dsh - (fr ? rv[v] * ir * pr : 0.f)) * ir;
      o2[v] = (dth - (ft ? tv[v] * it * pt : 0.f)) * it;
    }
    if (a.d_src) RowVec<float, V>::store(a.d_src + row + c, o1);
    if (a.d_dst) RowVec<float, V>::store(a.d_dst + row + c, o2);
    if (a.d_rel) RowVec<float, V>::store(a.d_rel + row + c, o3);
  }
}

// Ordered segmented row sum over entries sorted by key (stable sort => deterministic order):
//   out[key, :] (+)= sum_{p : sorted_keys[p] == key} rows[perm[p], :]
// One warp per entry position; only the warp sitting on the first entry of a key run works, so
// no segment table (and no host sync to size one) is needed.
template <int V>
__global__ void __launch_bounds__(kScoreWarps * 32)
index_add_sorted_kernel(const float* __restrict__ rows, const long long* __restrict__ perm,
                        const long long* __restrict__ sorted_keys, float* __restrict__ out, int M, int D,
                        int accumulate) {
  const int lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * kScoreWarps + (threadIdx.x >> 5);
  if (p0 >= M) return;
  const long long key = sorted_keys[p0];
  if (p0 > 0 && sorted_keys[p0 - 1] == key) return;
  int p1 = p0 + 1;  // end of the key run: 32 positions per probe (all lanes are still active here)
  while (true) {
    const int q = p1 + lane;
    const unsigned same = __ballot_sync(0xffffffffu, q < M && sorted_keys[q] == key);
    if (same == 0xffffffffu) { p1 += 32; continue; }
    p1 += __ffs(~same) - 1;
    break;
  }
  // blockIdx.y selects a 32*V-column slab, so a long key run (few relations, many triples) is
  // spread over D / (32*V) warps instead of one
  const int c = (blockIdx.y * 32 + lane) * V;
  if (c >= D) return;
  float* o = out + key * D;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  if (accumulate) ldv<V>(o + c, acc);
  // the sum stays strictly in run order (reproducible, same result as a serial loop); only the loads of
  // kRunUnroll rows are issued together — a run of ~100 rows (few relations) is latency-bound otherwise
  constexpr int kRunUnroll = 8;
  int p = p0;
  for (; p + kRunUnroll <= p1; p += kRunUnroll) {
    long long r[kRunUnroll];
    float x[kRunUnroll][V];
#pragma unroll
    for (int u = 0; u < kRunUnroll; ++u) r[u] = perm[p + u];
#pragma unroll
    for (int u = 0; u < kRunUnroll; ++u) ldv<V>(rows + r[u] * D + c, x[u]);
#pragma unroll
    for (int u = 0; u < kRunUnroll; ++u)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] += x[u][v];
  }
  for (; p < p1; ++p) {
    float x[V];
    ldv<V>(rows + perm[p] * D + c, x);
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] += x[v];
  }
  RowVec<float, V>::store(o + c, acc);
}

// Margin-ranking loss fused on the scores (reference core/loss/relgat_loss.py:51-54 on the split of
// trainer/relgat_projector.py:657-676 / :628-630):  loss = mean_{b,k} relu(margin + neg[b,k] - pos[b]).
// score layout: positives [0, B), then B*K negatives; neg[b,k] = score[B + k*B + b] (K-major blocks,
// no-projection path) or score[B + b*K + k] (the projection path's view(B, K)).
// Also writes dloss/dscore so the backward of the scorer needs no further loss arithmetic.
// Single block, fixed summation order => reproducible.
__global__ void __launch_bounds__(256)
margin_loss_kernel(const float* __restrict__ score, int B, int K, float margin, int bk_layout,
                   float* __restrict__ loss, float* __restrict__ dscore) {
  __shared__ float red[256];
  const int tid = threadIdx.x;
  const float inv = (B > 0 && K > 0) ? 1.f / (static_cast<float>(B) * K) : 0.f;
  float acc = 0.f;
  for (int b = tid; b < B; b += blockDim.x) {
    const float pos = score[b];
    float dpos = 0.f;
    for (int k = 0; k < K; ++k) {
      const int o = bk_layout ? B + b * K + k : B + k * B + b;
      const float v = margin + score[o] - pos;
      const bool on = v > 0.f;
      acc += on ? v : 0.f;
      dscore[o] = on ? inv : 0.f;
      dpos -= on ? inv : 0.f;
    }
    dscore[b] = dpos;
  }
  red[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0) loss[0] = red[0] * inv;
}

static inline bool al16(const void* p) { return !p || reinterpret_cast<uintptr_t>(p) % 16 == 0; }

static bool score_vec4(const ScoreArgs& a) {
  return a.D % 4 == 0 && al16(a.xs) && al16(a.xd) && al16(a.rel_emb) && al16(a.transform) && al16(a.src_vec) &&
         al16(a.dst_vec) && al16(a.dtransform) && al16(a.d_src) && al16(a.d_dst) && al16(a.d_rel);
}

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_score_fwd(int kind, int normalize, const float* xs, const long long* src_idx, const float* xd,
                                const long long* dst_idx, const float* rel_emb, const long long* rel_ids, int B, int D,
                                float* score, float* transform, int Bt, float* src_vec, float* dst_vec, void* stream) {
  if (!xs || !xd || !rel_emb || !rel_ids || !score || B < 0 || D <= 0) return RG_ERR_ARG;
  if (kind != SCORER_DISTMULT && kind != SCORER_TRANSE) return RG_ERR_ARG;
  if (B == 0) return RG_OK;
  ScoreArgs a{};
  a.xs = xs; a.xd = xd; a.src_idx = src_idx; a.dst_idx = dst_idx; a.rel_emb = rel_emb; a.rel_ids = rel_ids;
  a.B = B; a.D = D; a.kind = kind; a.normalize = normalize;
  a.score = score; a.transform = transform; a.Bt = Bt; a.src_vec = src_vec; a.dst_vec = dst_vec;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (B + kScoreWarps - 1) / kScoreWarps;
  if (score_vec4(a)) score_fwd_kernel<4><<<blocks, kScoreWarps * 32, 0, s>>>(a);
  else score_fwd_kernel<1><<<blocks, kScoreWarps * 32, 0, s>>>(a);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_score_bwd(int kind, int normalize, const float* xs, const long long* src_idx, const float* xd,
                                const long long* dst_idx, const float* rel_emb, const long long* rel_ids, int B, int D,
                                const float* dscore, const float* dtransform, int Bt,
                                float* d_src, float* d_dst, float* d_rel, void* stream) {
  if (!xs || !xd || !rel_emb || !rel_ids || B < 0 || D <= 0) return RG_ERR_ARG;
  if (kind != SCORER_DISTMULT && kind != SCORER_TRANSE) return RG_ERR_ARG;
  if (B == 0) return RG_OK;
  ScoreArgs a{};
  a.xs = xs; a.xd = xd; a.src_idx = src_idx; a.dst_idx = dst_idx; a.rel_emb = rel_emb; a.rel_ids = rel_ids;
  a.B = B; a.D = D; a.kind = kind; a.normalize = normalize;
  a.dscore = dscore; a.dtransform = dtransform; a.Bt = Bt; a.d_src = d_src; a.d_dst = d_dst; a.d_rel = d_rel;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (B + kScoreWarps - 1) / kScoreWarps;
  if (score_vec4(a)) score_bwd_kernel<4><<<blocks, kScoreWarps * 32, 0, s>>>(a);
  else score_bwd_kernel<1><<<blocks, kScoreWarps * 32, 0, s>>>(a);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_index_add_sorted(const float* rows, const long long* perm, const long long* sorted_keys,
                                       float* out, int M, int D, int accumulate, void* stream) {
  if (M < 0 || D <= 0) return RG_ERR_ARG;
  if (M == 0) return RG_OK;
  if (!rows || !perm || !sorted_keys || !out) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (M + kScoreWarps - 1) / kScoreWarps;
  if (D % 4 == 0 && al16(rows) && al16(out))
    index_add_sorted_kernel<4><<<dim3(blocks, (D + 127) / 128), kScoreWarps * 32, 0, s>>>(rows, perm, sorted_keys, out, M, D, accumulate);
  else
    index_add_sorted_kernel<1><<<dim3(blocks, (D + 31) / 32), kScoreWarps * 32, 0, s>>>(rows, perm, sorted_keys, out, M, D, accumulate);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_margin_loss(const float* score, int B, int K, float margin, int bk_layout, float* loss,
                                  float* dscore, void* stream) {
  if (!score || !loss || !dscore || B < 0 || K < 0) return RG_ERR_ARG;
  margin_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(score, B, K, margin, bk_layout, loss, dscore);
  return cuda_status(cudaGetLastError());
}
