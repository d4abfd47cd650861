// bwd_src, second generation — the by-source pass of the RelGAT edge backward for the training path
// (fp32 rows, 128-bit vectors, bf16 (hi, lo) planes out, logit-table gradient dS beside dP).
//
// Why a second kernel: the ncu source view of the first one (profiles/r02_ncu_edge_kernels.md) shows 398 warp
// instructions per edge of which 85 are FFMAs; the rest is the item state machine (OWN / EDGE / ZERO / END items,
// window refills, register moves between the two ring slots), three dependent scalar gathers per edge (z[slot],
// t[dst], (max, 1/den)[dst]) and the exp that rebuilds alpha — at 12 warps per SM every one of those instructions
// costs ~9 cycles of warp time, so the kernel is bound by its instruction count, not by HBM.  Here:
//   * a pre-pass (bwd_coef_kernel, one thread per edge and head, fully parallel) turns the three gathers and the exp
//     into three coefficients per edge and head, stored in by-source order as one 16-byte record:
//         c0 = alpha * m            (weight of G[dst] in dP; m = attention-dropout keep scale or 1)
//         c1 = alpha * slope * m,   c2 = alpha * slope * t[dst]        =>  dz = c1 * <G[dst], P[src]> - c2
//     the main loop reads them as a sequential stream;
//   * plain nested control flow (sources of a chunk, edges of a source) instead of an item stream; the own row of
//     the NEXT source and the G rows of the next TWO edges are in flight in registers, and no row buffer is ever
//     copied (the edge loop is unrolled by two with the buffers' roles fixed);
//   * the (dst, rel) window of the next 32 edges is loaded one window ahead.
#include "edge_bwd_src.cuh"

namespace relgat {

struct CoefArgs {
  const float* z;        // [E, H] CSR order
  const float* minv;     // [N_dst, H, 2]
  const float* t;        // [N_dst, H]
  const int* csc_slot;   // [E]
  const int* csc_dst;    // [E]
  const uint32_t* edge_bits;
  float edge_scale;
  float* coef;           // [E, H, 4] = (c0, c1, c2, 0)
  long long n;           // E * H
  int H;
};

__global__ void __launch_bounds__(256) bwd_coef_kernel(const CoefArgs a) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= a.n) return;
  const long long p = i / a.H;
  const int h = static_cast<int>(i - p * a.H);
  const int slot = __ldg(a.csc_slot + p);
  const int j = __ldg(a.csc_dst + p);
  const float zz = __ldg(a.z + static_cast<long long>(slot) * a.H + h);
  const float2 mi = __ldg(reinterpret_cast<const float2*>(a.minv) + static_cast<long long>(j) * a.H + h);
  const float tt = __ldg(a.t + static_cast<long long>(j) * a.H + h);
  const float ee = zz > 0.f ? zz : kLeakySlope * zz;
  const float al = __expf(ee - mi.x) * mi.y;
  const float sl = zz > 0.f ? 1.f : kLeakySlope;
  const float ek = a.edge_bits ? keep_scale1(a.edge_bits, static_cast<long long>(slot) * a.H + h, a.edge_scale) : 1.f;
  reinterpret_cast<float4*>(a.coef)[i] = make_float4(al * ek, al * sl * ek, al * sl * tt, 0.f);  // i == p * H + h
}

constexpr int kSrc2Warps = 12;
constexpr int kSrc2Prefetch = 4;  // L2 prefetch distance in edges (two of them are already in registers)

template <int KV, bool ASM, int LPHC>
__global__ void __launch_bounds__(kSrc2Warps * 32, 1)
bwd_src2_kernel(const SrcArgs<float, 4> a, const float* __restrict__ coef) {
  constexpr int V = 4;
  extern __shared__ __align__(16) float dyn_sm[];
  constexpr int kOwnFloats = KV * 32 * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int hl = lm.hh - g * a.hg;
  float* p_own = dyn_sm + warp * kOwnFloats + lane * V;  // this lane's pieces: p_own + k * 32 * V
  float* a_sm = dyn_sm + kSrc2Warps * kOwnFloats;
  const int ds_n = a.hg * a.R;
  float* ds_sm = a_sm + (ASM ? a.hg * a.R * a.F : 0) + warp * ds_n;
  const int ds_col0 = C + g * ds_n;
  const int kstride = (LPHC > 0 ? LPHC : lm.lph) * V;
  const int lane_off = lm.head_off + lm.sub * V;
  const bool last_ok = lm.sub + lm.lph * (KV - 1) < lm.vph;
  const int row_bytes = a.hg * a.F * 4;
  const unsigned long long p_stride_b = static_cast<unsigned long long>(a.ldp) * 4;
  const unsigned long long g_stride_b = static_cast<unsigned long long>(C) * 4;
  const char* p_lane = reinterpret_cast<const char*>(a.P + lane_off);
  const char* g_lane = reinterpret_cast<const char*>(a.G + lane_off);
  const char* g_pf = reinterpret_cast<const char*>(a.G + g * a.hg * a.F) + lane * 128;
  const char* p_pf = reinterpret_cast<const char*>(a.P + g * a.hg * a.F) + lane * 128;
  const bool pf_lane_ok = lane * 128 < row_bytes;
  const float4* cf_lane = reinterpret_cast<const float4*>(coef) + lm.hh;  // + e * H: one 16-byte load per edge
#define RG_VALID(k_) ((k_) < KV - 1 || last_ok)

  const float* a_base;
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
  for (int i = lane; i < ds_n; i += 32) ds_sm[i] = 0.f;
  __syncwarp();

  // the three row buffers live for the whole kernel: the guarded last vector is zeroed ONCE (a lane whose last
  // vector lies outside its head never loads into it)
  float x0[KV][V], x1[KV][V];
#pragma unroll
  for (int v = 0; v < V; ++v) { x0[KV - 1][v] = 0.f; x1[KV - 1][v] = 0.f; }

  int* counter = a.work_counter ? a.work_counter + g : nullptr;
  for (int c = claim_chunk(counter, lane, blockIdx.x * kSrc2Warps + warp); c < a.n_chunks;
       c = counter ? claim_chunk(counter, lane, 0) : c + gridDim.x * kSrc2Warps) {
    const int4 ch = __ldg(a.chunks + c);
    const int n_lo = ch.x, nn = ch.y, part = ch.z;
    int cp0 = 0, cp1 = 0, cp2 = 0;
    if (lane <= nn) cp0 = __ldg(a.colptr + n_lo + lane);
    if (32 + lane <= nn) cp1 = __ldg(a.colptr + n_lo + 32 + lane);
    if (64 + lane <= nn) cp2 = __ldg(a.colptr + n_lo + 64 + lane);
#define RG_CP(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, cp0, (k_) & 31)       \
                             : ((k_) < 64 ? __shfl_sync(0xffffffffu, cp1, (k_) & 31) \
                                          : __shfl_sync(0xffffffffu, cp2, (k_) & 31)))
    int e_lo = RG_CP(0), e_hi = RG_CP(nn);
    if (part >= 0) {
      const int2 pe = __ldg(a.parts + part);
      e_lo = pe.x;
      e_hi = pe.y;
    }

    // (dst, rel) windows: w0 covers [base, base + 32), w1 the 32 edges after it
    int base = e_lo;
    int w_dst0 = 0, w_rel0 = 0, w_dst1 = 0, w_rel1 = 0;
    if (base + lane < e_hi) { w_dst0 = __ldg(a.csc_dst + base + lane); w_rel0 = __ldg(a.csc_rel + base + lane); }
    if (base + 32 + lane < e_hi) { w_dst1 = __ldg(a.csc_dst + base + 32 + lane); w_rel1 = __ldg(a.csc_rel + base + 32 + lane); }
#define RG_WIN(arr_, e_) __shfl_sync(0xffffffffu, ((e_) - base < 32) ? arr_##0 : arr_##1, ((e_) - base) & 31)

    // issue the loads of edge e_ into buffer x_ (row of G[dst]) and its coefficients; L2 prefetch further ahead
#define RG_ISSUE(e_, x_, q_)                                                        \
  if ((e_) < e_hi) {                                                                           \
    const int jd = RG_WIN(w_dst, e_);                                                          \
    const float* rowp = reinterpret_cast<const float*>(g_lane + static_cast<unsigned long long>(jd) * g_stride_b); \
    _Pragma("unroll") for (int k = 0; k < KV; ++k)                                             \
      if (RG_VALID(k)) RowVec<float, V>::load_stream(rowp + k * kstride, x_[k]);               \
    q_ = __ldg(cf_lane + static_cast<long long>(e_) * a.H);                                    \
    const int ep = (e_) + kSrc2Prefetch - 2;                                                   \
    if (ep < e_hi && ep - base < 64) {                                                         \
      const int jp = RG_WIN(w_dst, ep);                                                        \
      if (pf_lane_ok) prefetch_l2(g_pf + static_cast<unsigned long long>(jp) * g_stride_b);    \
    }                                                                                          \
  }


    // write the finished row of source k_ (dP planes + dS columns, or the partial row of a split source), clear acc
#define RG_CLOSE(k_)                                                                           \
  {                                                                                            \
    const long long rbase = (part >= 0 ? static_cast<long long>(part) : static_cast<long long>(n_lo + (k_))) * a.ldo; \
    if (part >= 0) {                                                                           \
      _Pragma("unroll") for (int kk = 0; kk < KV; ++kk)                                        \
        if (RG_VALID(kk)) RowVec<float, V>::store(a.part_acc + rbase + lane_off + kk * kstride, acc[kk]); \
    } else {                                                                                   \
      _Pragma("unroll") for (int kk = 0; kk < KV; ++kk)                                        \
        if (RG_VALID(kk)) {                                                                    \
          const long long off = rbase + lane_off + kk * kstride;                               \
          store_split_bf16<V>(a.dP_hi + off, a.dP_lo ? a.dP_lo + off : nullptr, acc[kk]);      \
        }                                                                                      \
    }                                                                                          \
    _Pragma("unroll") for (int kk = 0; kk < KV; ++kk)                                          \
      _Pragma("unroll") for (int v = 0; v < V; ++v) acc[kk][v] = 0.f;                          \
    __syncwarp();                                                                              \
    const long long drow = rbase + ds_col0;                                                    \
    if (part < 0 && (ds_n & 7) == 0 && (ds_col0 & 7) == 0 && (a.ldo & 7) == 0) {               \
      for (int i = lane * 8; i < ds_n; i += 256) {                                             \
        float dv[8];                                                                           \
        RowVec<float, 8>::load_shared(ds_sm + i, dv);                                          \
        *reinterpret_cast<float4*>(ds_sm + i) = make_float4(0.f, 0.f, 0.f, 0.f);               \
        *reinterpret_cast<float4*>(ds_sm + i + 4) = make_float4(0.f, 0.f, 0.f, 0.f);           \
        store_split_bf16<8>(a.dP_hi + drow + i, a.dP_lo ? a.dP_lo + drow + i : nullptr, dv);   \
      }                                                                                        \
    } else {                                                                                   \
      for (int i = lane; i < ds_n; i += 32) {                                                  \
        const float dv = ds_sm[i];                                                             \
        ds_sm[i] = 0.f;                                                                        \
        if (part >= 0) a.part_acc[drow + i] = dv;                                              \
        else {                                                                                 \
          const float hv = bf16_round(dv);                                                     \
          a.dP_hi[drow + i] = __float2bfloat16_rn(hv);                                         \
          if (a.dP_lo) a.dP_lo[drow + i] = __float2bfloat16_rn(dv - hv);                       \
        }                                                                                      \
      }                                                                                        \
    }                                                                                          \
    __syncwarp();                                                                              \
  }

    // make source k_ current: its own row (pulled into L2 when the previous source was opened) goes to the
    // lane-private shared-memory slots; the next source's row is prefetched.  Sources are consecutive rows of P.
#define RG_OPEN(k_)                                                                            \
  {                                                                                            \
    const char* rown = p_lane + static_cast<unsigned long long>(n_lo + (k_)) * p_stride_b;     \
    if ((k_) + 1 < nn && pf_lane_ok)                                                           \
      prefetch_l2(p_pf + static_cast<unsigned long long>(n_lo + (k_) + 1) * p_stride_b);       \
    float ow[KV][V];                                                                           \
    _Pragma("unroll") for (int v = 0; v < V; ++v) ow[KV - 1][v] = 0.f;                         \
    _Pragma("unroll") for (int kk = 0; kk < KV; ++kk)                                          \
      if (RG_VALID(kk)) RowVec<float, V>::load_stream(reinterpret_cast<const float*>(rown) + kk * kstride, ow[kk]); \
    _Pragma("unroll") for (int kk = 0; kk < KV; ++kk)                                          \
      RowVec<float, V>::store(p_own + kk * 32 * V, ow[kk]);                                    \
  }

    // one edge: dalpha = <G[dst], P[src]>, dz from the coefficients, dP += c0 * G + dz * A[rel], dS[rel] += dz
#define RG_EDGE(e_, x_, q_)                                                         \
  {                                                                                            \
    while ((e_) == seg_end) { /* the edge belongs to a later source: finish this one */        \
      RG_CLOSE(k_cur)                                                                          \
      ++k_cur;                                                                                 \
      RG_OPEN(k_cur)                                                                           \
      seg_end = RG_CP(k_cur + 1);                                                              \
    }                                                                                          \
    const int rl = RG_WIN(w_rel, e_);                                                          \
    float sd[V];                                                                               \
    _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = 0.f;                                 \
    _Pragma("unroll") for (int kk = 0; kk < KV; ++kk) {                                        \
      float pv[V];                                                                             \
      RowVec<float, V>::load_shared(p_own + kk * 32 * V, pv);                                  \
      _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = fmaf(x_[kk][v], pv[v], sd[v]);     \
    }                                                                                          \
    float dd = (sd[0] + sd[1]) + (sd[2] + sd[3]);                                              \
    if constexpr (LPHC > 0) dd = head_sum_c<LPHC>(dd); else dd = head_sum(dd, lm.lph);         \
    const float dzv = fmaf(q_.y, dd, -q_.z);                                                   \
    if (lm.sub == 0) ds_sm[hl * a.R + rl] += dzv;                                              \
    const float* ar = a_base + rl * a.F;                                                       \
    _Pragma("unroll") for (int kk = 0; kk < KV; ++kk) {                                        \
      if (RG_VALID(kk)) {                                                                      \
        float av[V];                                                                           \
        if (ASM) RowVec<float, V>::load_shared(ar + kk * kstride, av);                         \
        else RowVec<float, V>::load_cached(ar + kk * kstride, av);                             \
        _Pragma("unroll") for (int v = 0; v < V; ++v)                                          \
          acc[kk][v] = fmaf(q_.x, x_[kk][v], fmaf(dzv, av[v], acc[kk][v]));                    \
      }                                                                                        \
    }                                                                                          \
  }

    float acc[KV][V];
#pragma unroll
    for (int kk = 0; kk < KV; ++kk)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[kk][v] = 0.f;
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;

    int k_cur = 0;
    int seg_end = part >= 0 ? e_hi : RG_CP(1);
    RG_ISSUE(e_lo, x0, q0)
    RG_ISSUE(e_lo + 1, x1, q1)
    RG_OPEN(0)
    int e = e_lo;
    while (e < e_hi) {
      if (e - base >= 32) {  // slide the windows: w1 becomes current, the window after it is loaded
        base += 32;
        w_dst0 = w_dst1; w_rel0 = w_rel1;
        w_dst1 = 0; w_rel1 = 0;
        if (base + 32 + lane < e_hi) { w_dst1 = __ldg(a.csc_dst + base + 32 + lane); w_rel1 = __ldg(a.csc_rel + base + 32 + lane); }
      }
      RG_EDGE(e, x0, q0)
      RG_ISSUE(e + 2, x0, q0)
      if (e + 1 < e_hi) {
        RG_EDGE(e + 1, x1, q1)
        RG_ISSUE(e + 3, x1, q1)
      }
      e += 2;
    }
    // trailing sources of the chunk (the current one, then any without out-edges)
    while (k_cur < nn) {
      RG_CLOSE(k_cur)
      ++k_cur;
      if (k_cur < nn && part < 0) { RG_OPEN(k_cur) }
      if (part >= 0) break;
    }
#undef RG_EDGE
#undef RG_OPEN
#undef RG_CLOSE
#undef RG_ISSUE
#undef RG_WIN
#undef RG_CP
  }
#undef RG_VALID
}

template <int KV, int LPHC>
static int launch_src2(SrcArgs<float, 4> a, const float* coef, int sm_count, cudaStream_t s) {
  const int groups = a.H / a.hg;
  if (sm_count <= 0) sm_count = 148;
  int ctas = sm_count / groups;
  if (ctas < 1) ctas = 1;
  const int need = (a.n_chunks + kSrc2Warps - 1) / kSrc2Warps;
  if (ctas > need) ctas = need;
  const size_t ds_bytes = static_cast<size_t>(kSrc2Warps) * a.hg * a.R * sizeof(float);
  const size_t own_bytes = static_cast<size_t>(kSrc2Warps) * KV * 32 * 4 * sizeof(float);
  const size_t a_bytes = static_cast<size_t>(a.hg) * a.R * a.F * sizeof(float);
  const bool asm_ok = a_bytes <= kSmemBudgetA && own_bytes + a_bytes + ds_bytes <= 227 * 1024;
  if (own_bytes + ds_bytes > 227 * 1024) return RG_ERR_SHAPE;
  if (asm_ok) {
    cudaError_t e = cudaFuncSetAttribute(bwd_src2_kernel<KV, true, LPHC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return cuda_status(e);
    bwd_src2_kernel<KV, true, LPHC><<<dim3(ctas, groups), kSrc2Warps * 32, own_bytes + a_bytes + ds_bytes, s>>>(a, coef);
  } else {
    cudaError_t e = cudaFuncSetAttribute(bwd_src2_kernel<KV, false, LPHC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return cuda_status(e);
    bwd_src2_kernel<KV, false, LPHC><<<dim3(ctas, groups), kSrc2Warps * 32, own_bytes + ds_bytes, s>>>(a, coef);
  }
  return cuda_status(cudaGetLastError());
}

}  // namespace relgat

using namespace relgat;

// Training-path variant of relgat_layer_bwd_src (fp32 P / G rows with F % 4 == 0, planes out, want_ds semantics:
// rows ldo >= H*F + H*R wide, dS in the columns behind dP, no dz).  coef: float [E * H * 4] scratch (16-byte aligned).
// Returns RG_ERR_SHAPE for layouts it does not cover — the caller then uses relgat_layer_bwd_src.
extern "C" int relgat_layer_bwd_src2(const float* P, long long ldp, const float* G, const float* A, const float* z,
                                     const float* minv, const float* t, const int* colptr, const int* csc_slot,
                                     const int* csc_dst, const int* csc_rel, const int* chunks, int n_chunks,
                                     const int* parts, int n_parts, const int* long_node, const int* long_part_ptr,
                                     int n_long, float* part_acc, void* dP_hi, void* dP_lo, float* coef, long long E,
                                     const unsigned int* edge_bits, float edge_scale, long long ldo, int H, int F, int R,
                                     int sm_count, int* work_counter, void* stream) {
  if (!P || !G || !A || !colptr || !dP_hi || n_chunks < 0 || n_parts < 0 || n_long < 0 || H <= 0 || F <= 0 || R <= 0 || E < 0)
    return RG_ERR_ARG;
  if (E > 0 && (!z || !minv || !t || !csc_slot || !csc_dst || !csc_rel || !coef)) return RG_ERR_ARG;
  if (n_chunks > 0 && !chunks) return RG_ERR_ARG;
  if (n_parts > 0 && (!parts || !long_node || !long_part_ptr || !part_acc)) return RG_ERR_ARG;
  if (ldo < static_cast<long long>(H) * F + static_cast<long long>(H) * R) return RG_ERR_ARG;
  if (n_chunks == 0) return RG_OK;
  auto al = [](const void* p) { return !p || reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (F % 4 != 0 || ldp % 4 != 0 || ldo % 4 != 0) return RG_ERR_SHAPE;
  if (!al(P) || !al(G) || !al(A) || !al(dP_hi) || !al(dP_lo) || !al(part_acc)) return RG_ERR_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int hg = pick_heads_per_warp(H, F, 4, R, kSmemBudgetA);
  if (!hg || H / hg > 32) return RG_ERR_SHAPE;
  if (E > 0) {
    CoefArgs ca{z, minv, t, csc_slot, csc_dst, edge_bits, edge_scale, coef, E * H, H};
    bwd_coef_kernel<<<static_cast<unsigned>((E * H + 255) / 256), 256, 0, s>>>(ca);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_status(e);
  }
  if (work_counter) {
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int) * (H / hg), s);
    if (e != cudaSuccess) return cuda_status(e);
  }
  SrcArgs<float, 4> a{P, G, A, z, minv, t, colptr, csc_slot, csc_dst, csc_rel, reinterpret_cast<const int4*>(chunks),
                      reinterpret_cast<const int2*>(parts), part_acc, nullptr, static_cast<__nv_bfloat16*>(dP_hi),
                      static_cast<__nv_bfloat16*>(dP_lo), nullptr, n_chunks, H, F, R, hg, ldp, 0, 0, work_counter,
                      nullptr, 1.f, 1, ldo};
  const int kv = vectors_per_lane(F / 4, hg);
  const int lph = 32 / hg;
  int rc;
  switch (kv) {
    case 1: rc = launch_src2<1, 0>(a, coef, sm_count, s); break;
    case 2: rc = lph == 32 ? launch_src2<2, 32>(a, coef, sm_count, s) : launch_src2<2, 0>(a, coef, sm_count, s); break;
    case 3: rc = launch_src2<3, 0>(a, coef, sm_count, s); break;
    case 4: rc = launch_src2<4, 0>(a, coef, sm_count, s); break;
    case 5: rc = launch_src2<5, 0>(a, coef, sm_count, s); break;
    case 6: rc = launch_src2<6, 0>(a, coef, sm_count, s); break;
    case 7: rc = lph == 8 ? launch_src2<7, 8>(a, coef, sm_count, s) : launch_src2<7, 0>(a, coef, sm_count, s); break;
    default: rc = launch_src2<8, 0>(a, coef, sm_count, s); break;
  }
  if (rc != RG_OK || n_long == 0) return rc;
  bwd_src_merge_kernel<float, 4><<<n_long, 128, 0, s>>>(a, long_node, long_part_ptr, n_long);
  return cuda_status(cudaGetLastError());
}
