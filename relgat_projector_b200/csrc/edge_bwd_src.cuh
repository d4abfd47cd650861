// bwd_src — by-source pass of the RelGAT edge backward (kernel + launch ladder).  Included by the three translation
// units edge_bwd_src_{f32,f32s,bf16}.cu, one per (storage type, vector width), so that they compile in parallel.
#pragma once
#include "common.cuh"

namespace relgat {

// ------------------------------------------------------------------------------------
// bwd_src — edge-balanced streaming over the by-source (CSC) order.
// The CSC edge array is cut at source boundaries into chunks of ~32 edges (graph.py: StreamChunks;
// sources with more than 512 out-edges are split into 256-edge parts merged by bwd_src_merge_kernel).
// Persistent CTAs (one per SM, 12 warps) own one head-group and keep its attention vectors in shared
// memory; a warp claims one chunk at a time (atomic work counter) and streams it as a sequence of row
// "items": OWN(i) = the source's own P row (needed for dalpha = <G[dst], P[i]>; parked in a
// lane-private shared-memory slot), followed by one EDGE item per out-edge (the gathered G[dst]
// row).  Two items are in flight per warp and the pipeline does not drain at source boundaries.
// ------------------------------------------------------------------------------------
constexpr int kSrcWarps = 12;      // plain variant
constexpr int kSrcWarpsPipe = 12;  // two-slot ring variant (same register budget as the plain one)
constexpr int kSrcPrefetchDist = 2;
constexpr int kSrcPipeDefault = 1;  // two-slot ring (1.44 -> 1.38 ms on config 2); 0 = load two items, consume two

__device__ __forceinline__ void prefetch_l2(const char* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T, int V>
struct SrcArgs {
  const T* P;           // [N_src, C]   (row stride ldp), fp32 or bf16
  const T* G;           // [N_dst, C]   same storage type as P
  const float* A;       // [H, R, F]
  const float* z;       // [E, H] CSR order
  const float* minv;    // [N_dst, H, 2] forward softmax statistics (max, 1/den)
  const float* t;       // [N_dst, H]
  const int* colptr;    // [N_src+1]
  const int* csc_slot;  // [E] CSR slot of each by-source edge
  const int* csc_dst;   // [E]
  const int* csc_rel;   // [E]
  const int4* chunks;   // [n_chunks] (first source, count <= 64, part slot or -1, 0)
  const int2* parts;    // [n_parts] (first edge, end edge) of the parts of split (high out-degree) sources
  float* part_acc;      // [n_parts, C] partial dP rows of split sources
  float* dP;            // [N_src, C] fp32 (may be nullptr when only the bf16 split is wanted)
  __nv_bfloat16* dP_hi; // optional bf16 split of dP for the tensor-core GEMMs
  __nv_bfloat16* dP_lo;
  float* dz;            // [E, H] CSR order
  int n_chunks, H, F, R, hg;
  long long ldp;
  int a_in_smem;
  int pf_dist;  // L2 prefetch distance in edges (0 = off)
  int* work_counter;  // zeroed device ints (one per head-group): dynamic chunk claim; nullptr = static
  const uint32_t* edge_bits;  // attention-dropout keep bits (index = csr slot * H + head) or nullptr
  float edge_scale;
  // logit-table gradient: with ds_on the dP rows are ldo = C + H*R wide and columns C + h*R + r
  // receive dS[i, h, r] = sum_{e: src = i, rel = r} dz[e, h].  The dW GEMM over the widened rows then also yields
  // dS^T X, from which dA = (dS^T X) W^T follows without a third gather of P (the by-relation pass).
  int ds_on;
  long long ldo;        // row stride of dP / dP_hi / dP_lo / part_acc in elements (C, or C + H*R with ds_on)
  // exact-zero hint (ds_on only): bit j set = row j of G may be non-zero.  An edge into a row whose bit is clear has
  // G[dst] = 0 and t[dst] = 0, hence dz = 0 and no contribution to dP or dS: it is skipped without touching memory.
  // nullptr = every row may be non-zero.
  const uint32_t* nz_bits;
  // compacted output (ds_on only): src_row[i] = row of dP / dP_hi / dP_lo that receives source i, or -1 = source i has
  // no edge into a non-zero row: it is skipped altogether (no read of its P row, no output row).  nullptr = row i.
  const int* src_row;
  int p_compact;  // with src_row: P holds the rows of the kept sources only, source i at row src_row[i]
};

// DS: logit-table gradient columns on (a.ds_on), compile-time so that the plain variants carry none of its code.
template <typename T, int V, int KV, bool ASM, int PIPE, int LPHC, bool DS>
__global__ void __launch_bounds__((PIPE ? kSrcWarpsPipe : kSrcWarps) * 32, 1) bwd_src_kernel(const SrcArgs<T, V> a) {
  constexpr int kWarps = PIPE ? kSrcWarpsPipe : kSrcWarps;
  extern __shared__ __align__(16) float dyn_sm[];
  constexpr int kOwnFloats = KV * 32 * V;  // lane-private slots of one warp's own row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int hl = lm.hh - g * a.hg;
  float* p_own = dyn_sm + warp * kOwnFloats;
  float* a_sm = dyn_sm + kWarps * kOwnFloats;
  const int ds_n = a.hg * a.R;  // dS entries of this warp's head-group for one source
  float* ds_sm = a_sm + (ASM ? a.hg * a.R * a.F : 0) + warp * ds_n;
  const int ds_col0 = C + g * ds_n;  // first dS column of this head-group inside an output row

  const int kstride = (LPHC > 0 ? LPHC : lm.lph) * V;  // compile-time on the specialised paths
  const int lane_off = lm.head_off + lm.sub * V;
  const bool last_ok = lm.sub + lm.lph * (KV - 1) < lm.vph;
  const int row_bytes = a.hg * a.F * static_cast<int>(sizeof(T));
  const int grp_off = g * a.hg * a.F;
  const unsigned long long p_stride_b = static_cast<unsigned long long>(a.ldp) * sizeof(T);
  const unsigned long long g_stride_b = static_cast<unsigned long long>(C) * sizeof(T);
  const char* p_lane = reinterpret_cast<const char*>(a.P + lane_off);
  const char* g_lane = reinterpret_cast<const char*>(a.G + lane_off);
  const char* p_pf = reinterpret_cast<const char*>(a.P + grp_off) + lane * 128;
  const char* g_pf = reinterpret_cast<const char*>(a.G + grp_off) + lane * 128;
  const bool pf_lane_ok = lane * 128 < row_bytes;
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

  enum { IT_NONE = 0, IT_OWN = 1, IT_EDGE = 2, IT_ZERO = 3, IT_END = 4 };

  int* counter = a.work_counter ? a.work_counter + g : nullptr;
  for (int c = claim_chunk(counter, lane, blockIdx.x * kWarps + warp); c < a.n_chunks;
       c = counter ? claim_chunk(counter, lane, 0) : c + gridDim.x * kWarps) {
    const int4 ch = __ldg(a.chunks + c);
    const int n_lo = ch.x;
    const int nn = ch.y;     // 1..64 sources
    const int part = ch.z;   // >= 0: one part of a split (high out-degree) source
    int cp0 = 0, cp1 = 0, cp2 = 0;
    if (lane <= nn) cp0 = __ldg(a.colptr + n_lo + lane);
    if (32 + lane <= nn) cp1 = __ldg(a.colptr + n_lo + 32 + lane);
    if (64 + lane <= nn) cp2 = __ldg(a.colptr + n_lo + 64 + lane);
    int sr0 = 0, sr1 = 0;  // output rows of the chunk's sources (nn <= 64)
    if (DS && a.src_row) {
      if (lane < nn) sr0 = __ldg(a.src_row + n_lo + lane);
      if (32 + lane < nn) sr1 = __ldg(a.src_row + n_lo + 32 + lane);
    }
#define RG_SR(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, sr0, (k_) & 31) : __shfl_sync(0xffffffffu, sr1, (k_) & 31))
#define RG_OROW(k_) ((DS && a.src_row) ? RG_SR(k_) : n_lo + (k_))
#define RG_PROW(k_) ((DS && a.src_row && a.p_compact) ? RG_SR(k_) : n_lo + (k_))
#define RG_CP(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, cp0, (k_) & 31)       \
                             : ((k_) < 64 ? __shfl_sync(0xffffffffu, cp1, (k_) & 31) \
                                          : __shfl_sync(0xffffffffu, cp2, (k_) & 31)))
    int e_lo = RG_CP(0);
    int e_hi = RG_CP(nn);
    if (part >= 0) {
      const int2 pe = __ldg(a.parts + part);
      e_lo = pe.x;
      e_hi = pe.y;
    }

    float acc[KV][V];
#pragma unroll
    for (int k = 0; k < KV; ++k)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[k][v] = 0.f;
    if (DS) {
      for (int i = lane; i < ds_n; i += 32) ds_sm[i] = 0.f;
      __syncwarp();
    }

    // fetch cursor (item generation)
    int fk = 0;             // source whose items are being generated
    int fe = e_lo;          // next edge to hand out
    int f_end = part >= 0 ? e_hi : RG_CP(1);   // end of source fk's edges
    bool own_done = false;  // OWN(fk) already handed out
    bool end_done = false;  // the closing IT_END item already handed out
    int base = e_lo - 32;   // edge-metadata window [base, base + 32) held across the lanes
    int my_slot = 0, my_dst = 0, my_rel = 0;
    unsigned nzmask = 0xffffffffu;  // window lanes whose destination row may be non-zero (a.nz_bits)
    int cur = -1;           // source being accumulated by the consumer (-1: none yet)

    // next item of the stream: OWN(k) | ZERO(k) (source without out-edges) | EDGE | END | NONE
#define RG_NEXT(ty_, nd_, sl_, ds_, rl_)                                                       \
  {                                                                                            \
    ty_ = IT_NONE; nd_ = 0; sl_ = 0; ds_ = 0; rl_ = 0;                                         \
    while (true) {                                                                             \
      if (fk >= nn) {                                                                          \
        if (!end_done) { end_done = true; ty_ = IT_END; }                                      \
        break;                                                                                 \
      }                                                                                        \
      if (!own_done && DS && a.src_row && RG_SR(fk) < 0) { /* no edge into a non-zero row */    \
        fe = f_end; ++fk;                                                                      \
        if (fk < nn) f_end = RG_CP(fk + 1);                                                    \
        continue;                                                                              \
      }                                                                                        \
      if (!own_done) {                                                                         \
        own_done = true; nd_ = fk;                                                             \
        ty_ = (f_end == fe) ? IT_ZERO : IT_OWN;                                                \
        if (a.pf_dist > 0 && fk + 2 < nn) { /* own rows: keep two ahead in L2 */                \
          const bool wanted = !(DS && a.src_row) || RG_SR(fk + 2) >= 0;                        \
          const int prow = RG_PROW(fk + 2); /* shuffles: every lane takes part */               \
          if (pf_lane_ok && wanted)                                                            \
            prefetch_l2(p_pf + static_cast<unsigned long long>(prow) * p_stride_b);            \
        }                                                                                      \
        break;                                                                                 \
      }                                                                                        \
      if (fe < f_end) {                                                                        \
        if (fe >= base + 32) {                                                                 \
          base = fe;                                                                           \
          const int idx = base + lane;                                                         \
          int my_nz = 0;                                                                       \
          if (idx < e_hi) {                                                                    \
            my_slot = __ldg(a.csc_slot + idx);                                                 \
            my_dst = __ldg(a.csc_dst + idx);                                                   \
            my_rel = __ldg(a.csc_rel + idx);                                                   \
            my_nz = (DS && a.nz_bits) ? ((__ldg(a.nz_bits + (my_dst >> 5)) >> (my_dst & 31)) & 1u) : 1; \
          }                                                                                    \
          nzmask = __ballot_sync(0xffffffffu, my_nz);                                          \
          for (int pf = 0; pf < a.pf_dist; ++pf) { /* warm L2 with the window's first rows */  \
            const int jp = __shfl_sync(0xffffffffu, my_dst, pf);                               \
            if (pf_lane_ok && ((nzmask >> pf) & 1u))                                           \
              prefetch_l2(g_pf + static_cast<unsigned long long>(jp) * g_stride_b);            \
          }                                                                                    \
        }                                                                                      \
        if (DS && a.nz_bits) { /* jump over the edges whose gradient row is an exact zero */   \
          const unsigned rem = nzmask >> (fe - base);                                          \
          const int skip = rem ? __ffs(rem) - 1 : 32;                                          \
          if (skip) {                                                                          \
            fe = min(fe + skip, min(base + 32, f_end));                                        \
            continue;                                                                          \
          }                                                                                    \
        }                                                                                      \
        if (a.pf_dist > 0) { /* rolling prefetch, pf_dist edges ahead inside the window */      \
          const int tp = fe - base + a.pf_dist;                                                \
          const int jp = __shfl_sync(0xffffffffu, my_dst, tp & 31);                            \
          if (pf_lane_ok && tp < 32 && ((nzmask >> tp) & 1u))                                  \
            prefetch_l2(g_pf + static_cast<unsigned long long>(jp) * g_stride_b);              \
        }                                                                                      \
        sl_ = __shfl_sync(0xffffffffu, my_slot, fe - base);                                    \
        ds_ = __shfl_sync(0xffffffffu, my_dst, fe - base);                                     \
        rl_ = __shfl_sync(0xffffffffu, my_rel, fe - base);                                     \
        ty_ = IT_EDGE; nd_ = fk; ++fe;                                                         \
        break;                                                                                 \
      }                                                                                        \
      ++fk; own_done = false;                                                                  \
      if (fk < nn) f_end = RG_CP(fk + 1);                                                      \
    }                                                                                          \
  }

#define RG_ISSUE(ty_, nd_, ds_, x_)                                                            \
  _Pragma("unroll") for (int v = 0; v < V; ++v) x_[KV - 1][v] = 0.f;                           \
  if (ty_ == IT_OWN || ty_ == IT_EDGE) {                                                       \
    const T* rowp = reinterpret_cast<const T*>((ty_ == IT_OWN)                                 \
        ? p_lane + static_cast<unsigned long long>(RG_PROW(nd_)) * p_stride_b                  \
        : g_lane + static_cast<unsigned long long>(ds_) * g_stride_b);                         \
    _Pragma("unroll") for (int k = 0; k < KV; ++k)                                             \
      if (RG_VALID(k)) RowVec<T, V>::load_stream(rowp + k * kstride, x_[k]);                   \
  }

#define RG_EDGE_ITEM(sl_, rl_, x_, zz_, mi_, tt_)                                              \
  {                                                                                            \
    float dd = 0.f;                                                                            \
    float sd[V];  /* V independent partial sums: short FMA dependency chains */                 \
    _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = 0.f;                                 \
    _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                           \
      float pv[V];                                                                             \
      RowVec<float, V>::load_shared(p_own + (k * 32 + lane) * V, pv);                          \
      _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = fmaf(x_[k][v], pv[v], sd[v]);      \
    }                                                                                          \
    _Pragma("unroll") for (int v = 0; v < V; ++v) dd += sd[v];                                 \
    if constexpr (LPHC > 0) dd = head_sum_c<LPHC>(dd); else dd = head_sum(dd, lm.lph); /* dalpha */ \
    const float ee = zz_ > 0.f ? zz_ : kLeakySlope * zz_;                                      \
    float al = __expf(ee - mi_.x) * mi_.y;                                                     \
    /* attention dropout: out used alpha*m*s, so dalpha = m*s*<G,P> and the G term carries alpha*m*s */ \
    const float ek = a.edge_bits ? keep_scale1(a.edge_bits, static_cast<long long>(sl_) * a.H + lm.hh, a.edge_scale) : 1.f; \
    const float dzv = al * (dd * ek - tt_) * (zz_ > 0.f ? 1.f : kLeakySlope);                  \
    al *= ek;                                                                                  \
    if (lm.sub == 0) {                                                                         \
      if (!DS) { if (a.dz) a.dz[static_cast<long long>(sl_) * a.H + lm.hh] = dzv; }            \
      else ds_sm[hl * a.R + (rl_)] += dzv; /* one lane per head owns the slot */               \
    }                                                                                          \
    const float* ar = a_base + (rl_) * a.F;                                                    \
    _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                           \
      if (RG_VALID(k)) {                                                                       \
        float av[V];                                                                           \
        if (ASM) RowVec<float, V>::load_shared(ar + k * kstride, av);                          \
        else RowVec<float, V>::load_cached(ar + k * kstride, av);                              \
        _Pragma("unroll") for (int v = 0; v < V; ++v)                                          \
          acc[k][v] = fmaf(al, x_[k][v], fmaf(dzv, av[v], acc[k][v]));                         \
      }                                                                                        \
    }                                                                                          \
  }

    // EDGE: accumulate.  OWN / ZERO / END: close the source being accumulated (write its dP row),
    // then open the next one (OWN parks the freshly loaded row in the lane-private smem slots).
#define RG_CONSUME(ty_, nd_, sl_, rl_, x_, zz_, mi_, tt_)                                      \
  if (ty_ == IT_EDGE) {                                                                        \
    RG_EDGE_ITEM(sl_, rl_, x_, zz_, mi_, tt_)                                                  \
  } else {                                                                                     \
    if (DS && cur >= 0) { /* dS columns of the source being closed, then clear the slots */    \
      __syncwarp();                                                                            \
      const long long drow = (part >= 0 ? static_cast<long long>(part) : static_cast<long long>(RG_OROW(cur))) * a.ldo + ds_col0; \
      if (part < 0 && !a.dP && a.dP_hi && (ds_n & 7) == 0 && (ds_col0 & 7) == 0 && (a.ldo & 7) == 0) { \
        /* planes only (the training path): 8 values -> one 16-byte store per plane */          \
        for (int i = lane * 8; i < ds_n; i += 256) {                                           \
          float dv[8];                                                                         \
          RowVec<float, 8>::load_shared(ds_sm + i, dv);                                        \
          *reinterpret_cast<float4*>(ds_sm + i) = make_float4(0.f, 0.f, 0.f, 0.f);             \
          *reinterpret_cast<float4*>(ds_sm + i + 4) = make_float4(0.f, 0.f, 0.f, 0.f);         \
          store_split_bf16<8>(a.dP_hi + drow + i, a.dP_lo ? a.dP_lo + drow + i : nullptr, dv);  \
        }                                                                                      \
      } else {                                                                                 \
        for (int i = lane; i < ds_n; i += 32) {                                                \
          const float dv = ds_sm[i];                                                           \
          ds_sm[i] = 0.f;                                                                      \
          if (part >= 0) a.part_acc[drow + i] = dv;                                            \
          else {                                                                               \
            if (a.dP) a.dP[drow + i] = dv;                                                     \
            if (a.dP_hi) {                                                                     \
              const float hv = bf16_round(dv);                                                 \
              a.dP_hi[drow + i] = __float2bfloat16_rn(hv);                                     \
              if (a.dP_lo) a.dP_lo[drow + i] = __float2bfloat16_rn(dv - hv);                   \
            }                                                                                  \
          }                                                                                    \
        }                                                                                      \
      }                                                                                        \
      __syncwarp();                                                                            \
    }                                                                                          \
    if (cur >= 0 && part >= 0) { /* split source: park the partial row for the merge kernel */  \
      const long long row_off = static_cast<long long>(part) * a.ldo + lane_off;               \
      _Pragma("unroll") for (int k = 0; k < KV; ++k)                                           \
        if (RG_VALID(k)) RowVec<float, V>::store(a.part_acc + row_off + k * kstride, acc[k]);  \
    } else if (cur >= 0) {                                                                     \
      const long long row_off = static_cast<long long>(RG_OROW(cur)) * a.ldo + lane_off;       \
      _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                         \
        if (RG_VALID(k)) {                                                                     \
          const long long off = row_off + k * kstride;                                         \
          if (a.dP) RowVec<float, V>::store(a.dP + off, acc[k]);                               \
          if (a.dP_hi) store_split_bf16<V>(a.dP_hi + off, a.dP_lo ? a.dP_lo + off : nullptr, acc[k]); \
        }                                                                                      \
      }                                                                                        \
    }                                                                                          \
    cur = (ty_ == IT_END) ? -1 : (nd_);                                                        \
    _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                           \
      _Pragma("unroll") for (int v = 0; v < V; ++v) acc[k][v] = 0.f;                           \
      if (ty_ == IT_OWN) RowVec<float, V>::store(p_own + (k * 32 + lane) * V, x_[k]);          \
    }                                                                                          \
  }

    // FETCH: hand out the next two items, issue their row loads and per-edge scalars
#define RG_DECL(S_)                                                                            \
    int ty##S_##0 = IT_NONE, nd##S_##0 = 0, sl##S_##0 = 0, ds##S_##0 = 0, rl##S_##0 = 0;       \
    int ty##S_##1 = IT_NONE, nd##S_##1 = 0, sl##S_##1 = 0, ds##S_##1 = 0, rl##S_##1 = 0;       \
    float x##S_##0[KV][V], x##S_##1[KV][V];                                                    \
    float z##S_##0 = 0.f, z##S_##1 = 0.f, t##S_##0 = 0.f, t##S_##1 = 0.f;                      \
    float2 mi##S_##0 = make_float2(0.f, 0.f), mi##S_##1 = make_float2(0.f, 0.f);
#define RG_FETCH(S_)                                                                           \
    RG_NEXT(ty##S_##0, nd##S_##0, sl##S_##0, ds##S_##0, rl##S_##0);                            \
    if (ty##S_##0 != IT_NONE) { RG_NEXT(ty##S_##1, nd##S_##1, sl##S_##1, ds##S_##1, rl##S_##1); } \
    else { ty##S_##1 = IT_NONE; }                                                              \
    RG_ISSUE(ty##S_##0, nd##S_##0, ds##S_##0, x##S_##0);                                       \
    RG_ISSUE(ty##S_##1, nd##S_##1, ds##S_##1, x##S_##1);                                       \
    if (ty##S_##0 == IT_EDGE) {                                                                \
      z##S_##0 = __ldg(a.z + static_cast<long long>(sl##S_##0) * a.H + lm.hh);                 \
      t##S_##0 = __ldg(a.t + static_cast<long long>(ds##S_##0) * a.H + lm.hh);                 \
      mi##S_##0 = __ldg(reinterpret_cast<const float2*>(a.minv) + static_cast<long long>(ds##S_##0) * a.H + lm.hh); \
    }                                                                                          \
    if (ty##S_##1 == IT_EDGE) {                                                                \
      z##S_##1 = __ldg(a.z + static_cast<long long>(sl##S_##1) * a.H + lm.hh);                 \
      t##S_##1 = __ldg(a.t + static_cast<long long>(ds##S_##1) * a.H + lm.hh);                 \
      mi##S_##1 = __ldg(reinterpret_cast<const float2*>(a.minv) + static_cast<long long>(ds##S_##1) * a.H + lm.hh); \
    }
#define RG_FETCH1(S_, I_)                                                                      \
    RG_NEXT(ty##S_##I_, nd##S_##I_, sl##S_##I_, ds##S_##I_, rl##S_##I_);                       \
    RG_ISSUE(ty##S_##I_, nd##S_##I_, ds##S_##I_, x##S_##I_);                                   \
    if (ty##S_##I_ == IT_EDGE) {                                                               \
      z##S_##I_ = __ldg(a.z + static_cast<long long>(sl##S_##I_) * a.H + lm.hh);               \
      t##S_##I_ = __ldg(a.t + static_cast<long long>(ds##S_##I_) * a.H + lm.hh);               \
      mi##S_##I_ = __ldg(reinterpret_cast<const float2*>(a.minv) + static_cast<long long>(ds##S_##I_) * a.H + lm.hh); \
    }
#define RG_DRAIN(S_)                                                                           \
    RG_CONSUME(ty##S_##0, nd##S_##0, sl##S_##0, rl##S_##0, x##S_##0, z##S_##0, mi##S_##0, t##S_##0); \
    if (ty##S_##1 != IT_NONE)                                                                  \
      RG_CONSUME(ty##S_##1, nd##S_##1, sl##S_##1, rl##S_##1, x##S_##1, z##S_##1, mi##S_##1, t##S_##1);

    if (PIPE) {
      // two-slot ring: a slot is refilled right after it has been consumed, so the row loads of one item are
      // in flight while the other item is consumed (the plain variant loads two items, then consumes both:
      // 25 % of its stall samples sit on the first use of the loaded row — profiles/r01_summary.md)
      RG_DECL(A)
      RG_FETCH1(A, 0)
      while (true) {
        RG_FETCH1(A, 1)
        if (tyA0 == IT_NONE) break;
        RG_CONSUME(tyA0, ndA0, slA0, rlA0, xA0, zA0, miA0, tA0);
        RG_FETCH1(A, 0)
        if (tyA1 == IT_NONE) break;
        RG_CONSUME(tyA1, ndA1, slA1, rlA1, xA1, zA1, miA1, tA1);
      }
    } else {
      while (true) {
        RG_DECL(A)
        RG_FETCH(A)
        if (tyA0 == IT_NONE) break;
        RG_DRAIN(A)
      }
    }
#undef RG_FETCH1
#undef RG_DRAIN
#undef RG_FETCH
#undef RG_DECL
#undef RG_CONSUME
#undef RG_EDGE_ITEM
#undef RG_ISSUE
#undef RG_NEXT
#undef RG_CP
#undef RG_PROW
#undef RG_OROW
#undef RG_SR
  }
#undef RG_VALID
}


template <typename T, int V>
__global__ void __launch_bounds__(128)
bwd_src_merge_kernel(const SrcArgs<T, V> a, const int* __restrict__ long_node, const int* __restrict__ long_part_ptr,
                     int n_long) {
  const int C = a.H * a.F;
  const int li = blockIdx.x;
  if (li >= n_long) return;
  const int node = long_node[li];
  const int i = a.src_row ? a.src_row[node] : node;  // output row
  if (i < 0) return;                                  // skipped source: its parts were not written
  const int p_lo = long_part_ptr[li], p_hi = long_part_ptr[li + 1];
  for (int c = threadIdx.x * V; c < C; c += blockDim.x * V) {
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    for (int p = p_lo; p < p_hi; ++p) {
      float x[V];
      RowVec<float, V>::load_cached(a.part_acc + static_cast<long long>(p) * a.ldo + c, x);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] += x[v];
    }
    const long long off = static_cast<long long>(i) * a.ldo + c;
    if (a.dP) RowVec<float, V>::store(a.dP + off, acc);
    if (a.dP_hi) store_split_bf16<V>(a.dP_hi + off, a.dP_lo ? a.dP_lo + off : nullptr, acc);
  }
  if (a.ds_on) {  // the dS columns of the split source: ordered sum of its parts
    for (int c = C + threadIdx.x; c < C + a.H * a.R; c += blockDim.x) {
      float acc = 0.f;
      for (int p = p_lo; p < p_hi; ++p) acc += a.part_acc[static_cast<long long>(p) * a.ldo + c];
      const long long off = static_cast<long long>(i) * a.ldo + c;
      if (a.dP) a.dP[off] = acc;
      if (a.dP_hi) {
        const float hv = bf16_round(acc);
        a.dP_hi[off] = __float2bfloat16_rn(hv);
        if (a.dP_lo) a.dP_lo[off] = __float2bfloat16_rn(acc - hv);
      }
    }
  }
}

template <typename T, int V, int KV, int PIPE, int LPHC, bool DS>
static int launch_src_flags(SrcArgs<T, V> a, int sm_count, cudaStream_t s) {
  constexpr int kWarps = PIPE ? kSrcWarpsPipe : kSrcWarps;
  const int groups = a.H / a.hg;
  if (sm_count <= 0) sm_count = 148;
  int ctas = sm_count / groups;
  if (ctas < 1) ctas = 1;
  const int need = (a.n_chunks + kWarps - 1) / kWarps;
  if (ctas > need) ctas = need;
  const size_t ds_bytes = a.ds_on ? static_cast<size_t>(kWarps) * a.hg * a.R * sizeof(float) : 0;
  const size_t own_bytes = static_cast<size_t>(kWarps) * KV * 32 * V * sizeof(float);
  const size_t a_bytes = static_cast<size_t>(a.hg) * a.R * a.F * sizeof(float);
  a.a_in_smem = (a_bytes <= kSmemBudgetA && own_bytes + a_bytes + ds_bytes <= 227 * 1024) ? 1 : 0;
  if (own_bytes + ds_bytes > 227 * 1024) return RG_ERR_SHAPE;
  {
    const char* pv = getenv("RELGAT_SRC_PF_DIST");
    a.pf_dist = pv ? atoi(pv) : kSrcPrefetchDist;
    if (a.pf_dist < 0) a.pf_dist = 0;
    if (a.pf_dist > 30) a.pf_dist = 30;
  }
  if (a.a_in_smem) {
    cudaError_t e = cudaFuncSetAttribute(bwd_src_kernel<T, V, KV, true, PIPE, LPHC, DS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_status(e);
    bwd_src_kernel<T, V, KV, true, PIPE, LPHC, DS><<<dim3(ctas, groups), kWarps * 32, own_bytes + a_bytes + ds_bytes, s>>>(a);
  } else {
    if (own_bytes + ds_bytes > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(bwd_src_kernel<T, V, KV, false, PIPE, LPHC, DS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return cuda_status(e);
    }
    bwd_src_kernel<T, V, KV, false, PIPE, LPHC, DS><<<dim3(ctas, groups), kWarps * 32, own_bytes + ds_bytes, s>>>(a);
  }
  return cuda_status(cudaGetLastError());
}

template <typename T, int V, int KV, int PIPE, int LPHC>
static int launch_src_pipe(const SrcArgs<T, V>& a, int sm_count, cudaStream_t s) {
  if (a.ds_on) return launch_src_flags<T, V, KV, PIPE, LPHC, true>(a, sm_count, s);
  return launch_src_flags<T, V, KV, PIPE, LPHC, false>(a, sm_count, s);
}

template <typename T, int V, int KV>
static int launch_src_kv(const SrcArgs<T, V>& a, int sm_count, cudaStream_t s) {
  const char* v = getenv("RELGAT_SRC_PIPE");  // experiment knob
  // the ring pays off for fp32 rows (1.44 -> 1.38 ms on config 2); with bf16 rows it measured slower (1.46 -> 1.68 ms)
  const int pipe = v ? atoi(v) : (sizeof(T) == 4 ? kSrcPipeDefault : 0);
  const int lph = 32 / a.hg;
  constexpr bool kSpec8 = (V == 4 && KV == 7) || (V == 8 && KV == 4);  // F = 200, 4 heads per warp
  constexpr bool kSpec32 = (V == 4 && KV == 2);                        // F = 200, one head per warp
  if (pipe) {
    if (kSpec8 && lph == 8) return launch_src_pipe<T, V, KV, 1, kSpec8 ? 8 : 0>(a, sm_count, s);
    return launch_src_pipe<T, V, KV, 1, 0>(a, sm_count, s);
  }
  if (kSpec8 && lph == 8) return launch_src_pipe<T, V, KV, 0, kSpec8 ? 8 : 0>(a, sm_count, s);
  if (kSpec32 && lph == 32) return launch_src_pipe<T, V, KV, 0, kSpec32 ? 32 : 0>(a, sm_count, s);
  return launch_src_pipe<T, V, KV, 0, 0>(a, sm_count, s);
}

template <typename T, int V>
static int launch_src(const SrcArgs<T, V>& a, int sm_count, cudaStream_t s) {
  if (a.n_chunks == 0) return RG_OK;
  const int kv = vectors_per_lane(a.F / V, a.hg);
  switch (kv) {
    case 1: return launch_src_kv<T, V, 1>(a, sm_count, s);
    case 2: return launch_src_kv<T, V, 2>(a, sm_count, s);
    case 3: return launch_src_kv<T, V, 3>(a, sm_count, s);
    case 4: return launch_src_kv<T, V, 4>(a, sm_count, s);
    default: break;
  }
  if constexpr (V != 8) {
    switch (kv) {
      case 5: return launch_src_kv<T, V, 5>(a, sm_count, s);
      case 6: return launch_src_kv<T, V, 6>(a, sm_count, s);
      case 7: return launch_src_kv<T, V, 7>(a, sm_count, s);
      default: return launch_src_kv<T, V, 8>(a, sm_count, s);
    }
  }
  return RG_ERR_SHAPE;
}

template <typename T, int V>
int run_src(const void* P, long long ldp, const void* G, const float* A, const float* z, const float* minv,
                   const float* t, const int* colptr, const int* csc_slot, const int* csc_dst, const int* csc_rel,
                   const int4* ch, int n_chunks, const int2* pt, const int* long_node, const int* long_part_ptr,
                   int n_long, float* part_acc, float* dP, void* dP_hi, void* dP_lo, float* dz,
                   const uint32_t* edge_bits, float edge_scale, const uint32_t* nz_bits, const int* src_row, int p_compact,
                   int ds_on, long long ldo, int H, int F, int R, int sm_count, int* work_counter, cudaStream_t s) {
  const int hg = pick_heads_per_warp(H, F, V, R, smem_budget_override("RELGAT_SRC_BUDGET_KB", kSmemBudgetA));
  if (!hg) return RG_ERR_SHAPE;
  if (H / hg > 32) return RG_ERR_SHAPE;  // work_counter holds 32 ints (one per head-group)
  if (work_counter) {
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int) * (H / hg), s);
    if (e != cudaSuccess) return cuda_status(e);
  }
  SrcArgs<T, V> a{static_cast<const T*>(P), static_cast<const T*>(G), A, z, minv, t, colptr, csc_slot, csc_dst,
                  csc_rel, ch, pt, part_acc, dP, static_cast<__nv_bfloat16*>(dP_hi),
                  static_cast<__nv_bfloat16*>(dP_lo), dz, n_chunks, H, F, R, hg, ldp, 0, 0, work_counter,
                  edge_bits, edge_scale, ds_on, ldo, ds_on ? nz_bits : nullptr,
                  ds_on ? src_row : nullptr, p_compact};
  int rc = launch_src(a, sm_count, s);
  if (rc != RG_OK || n_long == 0) return rc;
  bwd_src_merge_kernel<T, V><<<n_long, 128, 0, s>>>(a, long_node, long_part_ptr, n_long);
  return cuda_status(cudaGetLastError());
}


}  // namespace relgat
