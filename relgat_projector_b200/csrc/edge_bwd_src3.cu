// bwd_src3 — by-source pass of the RelGAT edge backward, third generation: the gathered rows travel
// global -> shared memory as BULK ASYNC COPIES (cp.async.bulk + mbarrier complete_tx), D rows deep per warp,
// instead of through two register slots per warp.
//
// Why (profiles/r02_summary.md, "by-source pass"): the first generation holds every row in flight in registers
// (2 rows x 28 registers per lane, 12 warps x 168 registers = the whole register file) and keeps the head-group's
// attention vectors A in 160 KB of shared memory for the per-edge term dz * A[rel]; 24 rows in flight per SM are
// too few for the loaded DRAM latency (~2.3 us) and the pass ran at 0.70 of the copy bandwidth.  Here
//   * the per-edge term is gone: the kernel writes dPa[i] = sum_e alpha_e G[dst_e]  (WITHOUT + dz * A[rel]) beside
//     dS[i, h, r] = sum_{e: rel = r} dz, and the caller folds  dP = dPa + dS·A  into the GEMMs that consume the rows
//     ([dPa | dS]ᵀ·X gives dW after adding Aᵀ(dSᵀX);  dX = [dPa | dS]·[W ; A·W]) — so no A in shared memory;
//   * the freed shared memory is a per-warp ring of D row slots filled by cp.async.bulk: (D - 1) rows in flight per
//     warp while one is consumed, no registers held by loads in flight, no L2 prefetch heuristics.
// The item stream (OWN row / EDGE rows / ZERO / END per source, chunks claimed dynamically, split high-degree
// sources, exact-zero rows skipped, compact output rows) is the first generation's (edge_bwd_src.cuh).
// fp32 rows, F % 4 == 0, logit-table-gradient (want_ds) layout only; everything else stays on relgat_layer_bwd_src.
//
// ONE copy of the fetch / consume code with a run-time slot index: the first cut unrolled the ring (slot state in
// registers, one expansion of fetch + consume per slot) and got SLOWER with depth — 1.37 / 2.03 / 3.00 ms at depth
// 2 / 3 / 4 on config 2, 62 KB -> 120 KB of SASS: twelve warps at different places of a loop that no longer fits the
// instruction cache.  Now the per-slot state lives in shared memory too: the item record (type, source, csr slot,
// relation) is written by lane 0, and the per-edge scalars (z, t, softmax max and 1/den of this edge and head) arrive
// by cp.async (LDGSTS) whose completion is counted on the slot's mbarrier together with the row's bulk copy.
#include "common.cuh"

namespace relgat {

constexpr int kSrc3Warps = 12;

struct Src3Args {
  const float* P;       // [N_src, C] (row stride ldp)
  const float* G;       // [N_dst, C]
  const float* z;       // [E, H] CSR order
  const float* minv;    // [N_dst, H, 2]
  const float* t;       // [N_dst, H]
  const int* colptr;
  const int* csc_slot;
  const int* csc_dst;
  const int* csc_rel;
  const int4* chunks;
  const int2* parts;
  float* part_acc;      // [n_parts, ldo]
  float* dP;            // optional fp32 rows [*, ldo]
  __nv_bfloat16* dP_hi; // optional bf16 planes
  __nv_bfloat16* dP_lo;
  int n_chunks, H, F, R, hg;
  int depth;            // ring slots per warp (2..8)
  long long ldp, ldo;
  int* work_counter;
  const uint32_t* edge_bits;
  float edge_scale;
  const uint32_t* nz_bits;
  const int* src_row;
  int p_compact;
};

__device__ __forceinline__ uint32_t s3_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// 33 arrivals complete a slot: lane 0's expect_tx and every lane's "my cp.asyncs have landed"
__device__ __forceinline__ void s3_bar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 33;" ::"r"(s3_smem_u32(bar)));
}
template <int BYTES>
__device__ __forceinline__ void s3_cp_async(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s3_smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
// arrives on the barrier once all cp.async issued by this thread so far have completed
__device__ __forceinline__ void s3_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s3_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void s3_bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s3_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s3_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// one contiguous row (bytes % 16 == 0, both addresses 16-byte aligned): global -> this CTA's shared memory
__device__ __forceinline__ void s3_bulk_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s3_smem_u32(dst)), "l"(src), "r"(bytes), "r"(s3_smem_u32(bar)) : "memory");
}

// KV: 128-bit vectors per lane and row; LPHC: compile-time lanes per head (0 = runtime)
template <int KV, int LPHC>
__global__ void __launch_bounds__(kSrc3Warps * 32, 1) bwd_src3_kernel(const Src3Args a) {
  constexpr int V = 4;
  constexpr int kOwnFloats = KV * 32 * V;
  extern __shared__ __align__(128) float dyn_sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const LaneMap lm = make_lane_map<V>(lane, g, a.hg, a.F);
  const int C = a.H * a.F;
  const int hl = lm.hh - g * a.hg;
  const int row_floats = a.hg * a.F;                       // this head-group's slice of a row
  const int slot_floats = (row_floats + 31) & ~31;          // 128-byte slot pitch
  const uint32_t row_bytes = static_cast<uint32_t>(row_floats) * 4u;
  const int ds_n = a.hg * a.R;
  const int ds_pad = (ds_n + 3) & ~3;
  const int D = a.depth;
  // per warp: ring | per-slot scalars [D][hg][4] | per-slot item record [D][4] | own row | dS | D barriers; 128-byte pitch
  const int warp_floats = (D * slot_floats + D * a.hg * 4 + D * 4 + 2 * D + kOwnFloats + ds_pad + 31) & ~31;
  float* ring = dyn_sm + static_cast<size_t>(warp) * warp_floats;
  float4* meta_f = reinterpret_cast<float4*>(ring + D * slot_floats);
  int4* meta_i = reinterpret_cast<int4*>(meta_f + D * a.hg);
  float* p_own = reinterpret_cast<float*>(meta_i + D);
  float* ds_sm = p_own + kOwnFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_sm + ds_pad);  // every block before it is a multiple of 16 bytes
  const int ds_col0 = C + g * ds_n;

  const int kstride = (LPHC > 0 ? LPHC : lm.lph) * V;
  const int lane_off = lm.head_off + lm.sub * V;            // column of this lane's first vector in a full row
  const int grp_off = g * row_floats;
  const int slot_off = lane_off - grp_off;                  // ... and inside a ring slot
  const bool last_ok = lm.sub + lm.lph * (KV - 1) < lm.vph;
#define S3_VALID(k_) ((k_) < KV - 1 || last_ok)

  if (lane == 0) {
    for (int i = 0; i < D; ++i) s3_bar_init(&bars[i]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase_bits = 0;  // bit i = parity slot i's barrier completes next

  enum { IT_NONE = 0, IT_OWN = 1, IT_EDGE = 2, IT_ZERO = 3, IT_END = 4 };

  int* counter = a.work_counter ? a.work_counter + g : nullptr;
  for (int c = claim_chunk(counter, lane, blockIdx.x * kSrc3Warps + warp); c < a.n_chunks;
       c = counter ? claim_chunk(counter, lane, 0) : c + gridDim.x * kSrc3Warps) {
    const int4 ch = __ldg(a.chunks + c);
    const int n_lo = ch.x;
    const int nn = ch.y;     // 1..64 sources
    const int part = ch.z;   // >= 0: one part of a split (high out-degree) source
    int cp0 = 0, cp1 = 0, cp2 = 0;
    if (lane <= nn) cp0 = __ldg(a.colptr + n_lo + lane);
    if (32 + lane <= nn) cp1 = __ldg(a.colptr + n_lo + 32 + lane);
    if (64 + lane <= nn) cp2 = __ldg(a.colptr + n_lo + 64 + lane);
    int sr0 = 0, sr1 = 0;  // output rows of the chunk's sources (nn <= 64)
    if (a.src_row) {
      if (lane < nn) sr0 = __ldg(a.src_row + n_lo + lane);
      if (32 + lane < nn) sr1 = __ldg(a.src_row + n_lo + 32 + lane);
    }
#define S3_SR(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, sr0, (k_) & 31) : __shfl_sync(0xffffffffu, sr1, (k_) & 31))
#define S3_OROW(k_) (a.src_row ? S3_SR(k_) : n_lo + (k_))
#define S3_PROW(k_) ((a.src_row && a.p_compact) ? S3_SR(k_) : n_lo + (k_))
#define S3_CP(k_) ((k_) < 32 ? __shfl_sync(0xffffffffu, cp0, (k_) & 31)       \
                             : ((k_) < 64 ? __shfl_sync(0xffffffffu, cp1, (k_) & 31) \
                                          : __shfl_sync(0xffffffffu, cp2, (k_) & 31)))
    int e_lo = S3_CP(0);
    int e_hi = S3_CP(nn);
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
    for (int i = lane; i < ds_n; i += 32) ds_sm[i] = 0.f;
    __syncwarp();

    // item cursor (identical to the first generation's)
    int fk = 0;
    int fe = e_lo;
    int f_end = part >= 0 ? e_hi : S3_CP(1);
    bool own_done = false;
    bool end_done = false;
    int base = e_lo - 32;   // edge-metadata window [base, base + 32) held across the lanes
    int my_slot = 0, my_dst = 0, my_rel = 0;
    unsigned nzmask = 0xffffffffu;
    int cur = -1;           // source being accumulated (-1: none yet)

    // next item of the stream into slot I_: generate it, start its row copy, load its per-edge scalars
#define S3_FETCH(I_)                                                                           \
  {                                                                                            \
    int ty_ = IT_NONE, nd_ = 0, sl_ = 0, ds_ = 0, rl_ = 0;                                     \
    while (true) {                                                                             \
      if (fk >= nn) {                                                                          \
        if (!end_done) { end_done = true; ty_ = IT_END; }                                      \
        break;                                                                                 \
      }                                                                                        \
      if (!own_done && a.src_row && S3_SR(fk) < 0) { /* no edge into a non-zero row */         \
        fe = f_end; ++fk;                                                                      \
        if (fk < nn) f_end = S3_CP(fk + 1);                                                    \
        continue;                                                                              \
      }                                                                                        \
      if (!own_done) {                                                                         \
        own_done = true; nd_ = fk;                                                             \
        ty_ = (f_end == fe) ? IT_ZERO : IT_OWN;                                                \
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
            my_nz = a.nz_bits ? ((__ldg(a.nz_bits + (my_dst >> 5)) >> (my_dst & 31)) & 1u) : 1; \
          }                                                                                    \
          nzmask = __ballot_sync(0xffffffffu, my_nz);                                          \
        }                                                                                      \
        if (a.nz_bits) { /* jump over the edges whose gradient row is an exact zero */         \
          const unsigned rem = nzmask >> (fe - base);                                          \
          const int skip = rem ? __ffs(rem) - 1 : 32;                                          \
          if (skip) {                                                                          \
            fe = min(fe + skip, min(base + 32, f_end));                                        \
            continue;                                                                          \
          }                                                                                    \
        }                                                                                      \
        sl_ = __shfl_sync(0xffffffffu, my_slot, fe - base);                                    \
        ds_ = __shfl_sync(0xffffffffu, my_dst, fe - base);                                     \
        rl_ = __shfl_sync(0xffffffffu, my_rel, fe - base);                                     \
        ty_ = IT_EDGE; nd_ = fk; ++fe;                                                         \
        break;                                                                                 \
      }                                                                                        \
      ++fk; own_done = false;                                                                  \
      if (fk < nn) f_end = S3_CP(fk + 1);                                                      \
    }                                                                                          \
    if (lane == 0) meta_i[I_] = make_int4(ty_, nd_, sl_, rl_);                                 \
    if (ty_ == IT_OWN || ty_ == IT_EDGE) {                                                     \
      const int prow = S3_PROW(nd_); /* shuffles: every lane takes part */                      \
      if (lane == 0) {                                                                         \
        const float* src = (ty_ == IT_OWN) ? a.P + static_cast<long long>(prow) * a.ldp + grp_off \
                                           : a.G + static_cast<long long>(ds_) * C + grp_off;  \
        s3_bar_expect(&bars[I_], row_bytes);                                                   \
        s3_bulk_row(ring + (I_) * slot_floats, src, row_bytes, &bars[I_]);                     \
      }                                                                                        \
      if (ty_ == IT_EDGE && lm.sub == 0) { /* one lane per head: this edge's scalars -> the slot */ \
        float* mf = reinterpret_cast<float*>(meta_f + (I_) * a.hg + hl);                       \
        s3_cp_async<4>(mf, a.z + static_cast<long long>(sl_) * a.H + lm.hh);                   \
        s3_cp_async<4>(mf + 1, a.t + static_cast<long long>(ds_) * a.H + lm.hh);               \
        s3_cp_async<8>(mf + 2, reinterpret_cast<const float2*>(a.minv) + static_cast<long long>(ds_) * a.H + lm.hh); \
      }                                                                                        \
      s3_cp_async_arrive(&bars[I_]);                                                           \
    }                                                                                          \
  }

    // close the source being accumulated: its dS columns and its dPa row
#define S3_CLOSE()                                                                             \
  if (cur >= 0) {                                                                              \
    __syncwarp();                                                                              \
    const int orow = S3_OROW(cur);                                                             \
    const long long drow = (part >= 0 ? static_cast<long long>(part) : static_cast<long long>(orow)) * a.ldo + ds_col0; \
    if (part < 0 && !a.dP && a.dP_hi && (ds_n & 7) == 0 && (ds_col0 & 7) == 0 && (a.ldo & 7) == 0) { \
      for (int i = lane * 8; i < ds_n; i += 256) {                                             \
        float dv[8];                                                                           \
        RowVec<float, 8>::load_shared(ds_sm + i, dv);                                          \
        *reinterpret_cast<float4*>(ds_sm + i) = make_float4(0.f, 0.f, 0.f, 0.f);               \
        *reinterpret_cast<float4*>(ds_sm + i + 4) = make_float4(0.f, 0.f, 0.f, 0.f);           \
        store_split_bf16<8>(a.dP_hi + drow + i, a.dP_lo ? a.dP_lo + drow + i : nullptr, dv);    \
      }                                                                                        \
    } else {                                                                                   \
      for (int i = lane; i < ds_n; i += 32) {                                                  \
        const float dv = ds_sm[i];                                                             \
        ds_sm[i] = 0.f;                                                                        \
        if (part >= 0) a.part_acc[drow + i] = dv;                                              \
        else {                                                                                 \
          if (a.dP) a.dP[drow + i] = dv;                                                       \
          if (a.dP_hi) {                                                                       \
            const float hv = bf16_round(dv);                                                   \
            a.dP_hi[drow + i] = __float2bfloat16_rn(hv);                                       \
            if (a.dP_lo) a.dP_lo[drow + i] = __float2bfloat16_rn(dv - hv);                     \
          }                                                                                    \
        }                                                                                      \
      }                                                                                        \
    }                                                                                          \
    __syncwarp();                                                                              \
    const long long row_off = (part >= 0 ? static_cast<long long>(part) : static_cast<long long>(orow)) * a.ldo + lane_off; \
    _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                           \
      if (S3_VALID(k)) {                                                                       \
        const long long off = row_off + k * kstride;                                           \
        if (part >= 0) RowVec<float, V>::store(a.part_acc + off, acc[k]);                      \
        else {                                                                                 \
          if (a.dP) RowVec<float, V>::store(a.dP + off, acc[k]);                               \
          if (a.dP_hi) store_split_bf16<V>(a.dP_hi + off, a.dP_lo ? a.dP_lo + off : nullptr, acc[k]); \
        }                                                                                      \
      }                                                                                        \
    }                                                                                          \
  }

#define S3_CONSUME(I_, REC_)                                                                   \
  {                                                                                            \
    const int ty_ = REC_.x;                                                                    \
    float x_[KV][V];                                                                           \
    if (ty_ == IT_OWN || ty_ == IT_EDGE) {                                                     \
      s3_bar_wait(&bars[I_], (phase_bits >> (I_)) & 1u);                                       \
      phase_bits ^= 1u << (I_);                                                                \
      const float* row = ring + (I_) * slot_floats + slot_off;                                 \
      _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                         \
        if (S3_VALID(k)) RowVec<float, V>::load_shared(row + k * kstride, x_[k]);              \
        else { _Pragma("unroll") for (int v = 0; v < V; ++v) x_[k][v] = 0.f; }                 \
      }                                                                                        \
    }                                                                                          \
    if (ty_ == IT_EDGE) {                                                                      \
      float sd[V];                                                                             \
      _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = 0.f;                               \
      _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                         \
        float pv[V];                                                                           \
        RowVec<float, V>::load_shared(p_own + (k * 32 + lane) * V, pv);                        \
        _Pragma("unroll") for (int v = 0; v < V; ++v) sd[v] = fmaf(x_[k][v], pv[v], sd[v]);    \
      }                                                                                        \
      float dd = (sd[0] + sd[1]) + (sd[2] + sd[3]);                                            \
      if constexpr (LPHC > 0) dd = head_sum_c<LPHC>(dd); else dd = head_sum(dd, lm.lph);       \
      const float4 sc = meta_f[(I_) * a.hg + hl]; /* z, t, max, 1/den */                        \
      const float z_ = sc.x;                                                                   \
      const float ee = z_ > 0.f ? z_ : kLeakySlope * z_;                                       \
      float al = __expf(ee - sc.z) * sc.w;                                                     \
      const float ek = a.edge_bits ? keep_scale1(a.edge_bits, static_cast<long long>(REC_.z) * a.H + lm.hh, a.edge_scale) : 1.f; \
      const float dzv = al * (dd * ek - sc.y) * (z_ > 0.f ? 1.f : kLeakySlope);                \
      al *= ek;                                                                                \
      if (lm.sub == 0) ds_sm[hl * a.R + REC_.w] += dzv; /* one lane per head owns the slot */  \
      _Pragma("unroll") for (int k = 0; k < KV; ++k)                                           \
        _Pragma("unroll") for (int v = 0; v < V; ++v) acc[k][v] = fmaf(al, x_[k][v], acc[k][v]); \
    } else {                                                                                   \
      S3_CLOSE()                                                                               \
      cur = (ty_ == IT_END) ? -1 : REC_.y;                                                     \
      _Pragma("unroll") for (int k = 0; k < KV; ++k) {                                         \
        _Pragma("unroll") for (int v = 0; v < V; ++v) acc[k][v] = 0.f;                         \
        if (ty_ == IT_OWN) RowVec<float, V>::store(p_own + (k * 32 + lane) * V, x_[k]);        \
      }                                                                                        \
      if (ty_ == IT_OWN) __syncwarp();                                                         \
    }                                                                                          \
  }

    // fill the ring, then: consume the oldest slot, refill it (its lanes have all read it: __syncwarp)
#pragma unroll 1
    for (int i = 0; i < D; ++i) S3_FETCH(i)
    __syncwarp();
    int slot = 0;
#pragma unroll 1
    while (true) {
      const int4 rec = meta_i[slot];  // (type, source, csr slot, relation), written by lane 0 when the slot was filled
      if (rec.x == IT_NONE) break;
      S3_CONSUME(slot, rec)
      __syncwarp();
      S3_FETCH(slot)
      __syncwarp();
      slot = (slot + 1 == D) ? 0 : slot + 1;
    }
#undef S3_CONSUME
#undef S3_CLOSE
#undef S3_FETCH
#undef S3_CP
#undef S3_PROW
#undef S3_OROW
#undef S3_SR
  }
#undef S3_VALID
}

// dP row (and dS columns) of a split source: ordered sum of its parts
__global__ void __launch_bounds__(128)
bwd_src3_merge_kernel(const Src3Args a, const int* __restrict__ long_node, const int* __restrict__ long_part_ptr, int n_long) {
  const int W = a.H * a.F + a.H * a.R;
  const int li = blockIdx.x;
  if (li >= n_long) return;
  const int node = long_node[li];
  const int i = a.src_row ? a.src_row[node] : node;
  if (i < 0) return;
  const int p_lo = long_part_ptr[li], p_hi = long_part_ptr[li + 1];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
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

template <int KV, int LPHC>
static int launch_src3(Src3Args a, int sm_count, cudaStream_t s) {
  const int groups = a.H / a.hg;
  int ctas = sm_count / groups;
  if (ctas < 1) ctas = 1;
  const int need = (a.n_chunks + kSrc3Warps - 1) / kSrc3Warps;
  if (ctas > need) ctas = need;
  const int row_floats = a.hg * a.F;
  const int slot_floats = (row_floats + 31) & ~31;
  const int ds_pad = (a.hg * a.R + 3) & ~3;
  int depth = 4;
  if (const char* v = getenv("RELGAT_SRC3_DEPTH")) depth = atoi(v);  // experiment knob: ring depth 2..8
  if (depth > 8) depth = 8;
  size_t smem = 0;
  for (; depth >= 2; --depth) {  // deepest ring that fits 227 KB of shared memory
    const size_t warp_floats = (static_cast<size_t>(depth) * (slot_floats + a.hg * 4 + 4 + 2) + KV * 32 * 4 + ds_pad + 31) &
                               ~static_cast<size_t>(31);
    smem = warp_floats * kSrc3Warps * sizeof(float);
    if (smem <= 227 * 1024) break;
  }
  if (depth < 2) return RG_ERR_SHAPE;
  a.depth = depth;
  cudaError_t e = cudaFuncSetAttribute(bwd_src3_kernel<KV, LPHC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_status(e);
  bwd_src3_kernel<KV, LPHC><<<dim3(ctas, groups), kSrc3Warps * 32, smem, s>>>(a);
  return cuda_status(cudaGetLastError());
}

}  // namespace relgat

using namespace relgat;

// By-source pass WITHOUT the per-edge attention-vector term (see the header of this file): rows [dPa | dS], ldo >=
// H*F + H*R wide; fp32 P / G rows, F % 4 == 0, 16-byte aligned rows.  Other arguments as relgat_layer_bwd_src with
// want_ds = 1.  RG_ERR_SHAPE: layout not covered (use relgat_layer_bwd_src and its complete dP rows).
extern "C" int relgat_layer_bwd_src3(const float* P, long long ldp, const float* G, const float* z, const float* minv,
                                     const float* t, const int* colptr, const int* csc_slot, const int* csc_dst,
                                     const int* csc_rel, const int* chunks, int n_chunks, const int* parts, int n_parts,
                                     const int* long_node, const int* long_part_ptr, int n_long, float* part_acc,
                                     float* dP, void* dP_hi, void* dP_lo, const unsigned int* edge_bits,
                                     float edge_scale, const unsigned int* dst_nz_bits, const int* src_row,
                                     int p_compact, long long ldo, int H, int F, int R, int sm_count,
                                     int* work_counter, void* stream) {
  // (z / minv / t / csc_* are only read for edges: a graph without edges passes empty arrays)
  if (!P || !G || !colptr || (!dP && !dP_hi) || H <= 0 || F <= 0 || R <= 0 || n_chunks < 0) return RG_ERR_ARG;
  if (n_chunks > 0 && !chunks) return RG_ERR_ARG;
  if (n_parts > 0 && (!parts || !part_acc)) return RG_ERR_ARG;
  if (F % 4 != 0 || ldp % 4 != 0 || ldo < static_cast<long long>(H) * F + static_cast<long long>(H) * R) return RG_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(P) % 16 || reinterpret_cast<uintptr_t>(G) % 16) return RG_ERR_ALIGN;
  if (n_chunks == 0) return RG_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int hg = pick_heads_per_warp(H, F, 4);
  if (!hg || H / hg > 32) return RG_ERR_SHAPE;
  if (sm_count <= 0) sm_count = 148;
  if (work_counter) {
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int) * (H / hg), s);
    if (e != cudaSuccess) return cuda_status(e);
  }
  Src3Args a{P, G, z, minv, t, colptr, csc_slot, csc_dst, csc_rel, reinterpret_cast<const int4*>(chunks),
             reinterpret_cast<const int2*>(parts), part_acc, dP, static_cast<__nv_bfloat16*>(dP_hi),
             static_cast<__nv_bfloat16*>(dP_lo), n_chunks, H, F, R, hg, 0, ldp, ldo, work_counter, edge_bits, edge_scale,
             dst_nz_bits, src_row, p_compact};
  const int kv = vectors_per_lane(F / 4, hg);
  const int lph = 32 / hg;
  int rc = RG_ERR_SHAPE;
  switch (kv) {
    case 1: rc = launch_src3<1, 0>(a, sm_count, s); break;
    case 2: rc = launch_src3<2, 0>(a, sm_count, s); break;
    case 3: rc = launch_src3<3, 0>(a, sm_count, s); break;
    case 4: rc = launch_src3<4, 0>(a, sm_count, s); break;
    case 5: rc = launch_src3<5, 0>(a, sm_count, s); break;
    case 6: rc = launch_src3<6, 0>(a, sm_count, s); break;
    case 7: rc = (lph == 8) ? launch_src3<7, 8>(a, sm_count, s) : launch_src3<7, 0>(a, sm_count, s); break;
    case 8: rc = (lph == 8) ? launch_src3<8, 8>(a, sm_count, s) : launch_src3<8, 0>(a, sm_count, s); break;
    default: break;
  }
  if (rc != RG_OK || n_long == 0) return rc;
  bwd_src3_merge_kernel<<<n_long, 128, 0, s>>>(a, long_node, long_part_ptr, n_long);
  return cuda_status(cudaGetLastError());
}
