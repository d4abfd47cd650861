// Shared device helpers for the RelGAT sm_100a kernels.
//
// Row layout convention used by every edge kernel: a node row is [H, F] elements
// (column = h*F + f, the head-major concat order of reference layer.py:321).  One warp
// serves one (node, head-group) task: the group holds `hg` heads (power of two, divides H),
// each head is owned by lph = 32/hg consecutive lanes, and a lane walks its head in
// 128-bit vectors q = sub + lph*k, k < KMAX.  Because a lane only ever touches one head,
// per-head reductions are log2(lph) xor-shuffles shared by all heads of the group.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace relgat {

constexpr int kMaxVecPerLane = 8;  // KMAX: 128-bit vectors a lane may own per row (fp32 layout)
// a lane owns at most 32 feature elements: 8 vectors of 4 fp32, or 4 vectors of 8 bf16
template <int V>
__host__ __device__ constexpr int max_vec() { return V == 8 ? 4 : 8; }
inline int max_vec_rt(int V) { return V == 8 ? 4 : 8; }
constexpr float kLeakySlope = 0.2f;  // reference layer.py:233
// shared-memory budget for a head-group's attention vectors (hg*R*F fp32) in the edge kernels;
// leaves room for the by-source kernel's per-warp own-row slots (12 x 4 KB) inside 227 KB
constexpr size_t kSmemBudgetA = 176 * 1024;

// ---- error codes of the C ABI (include/relgat_b200.h) ------------------------------
enum : int {
  RG_OK = 0,
  RG_ERR_ARG = -1,        // null pointer / negative size
  RG_ERR_SHAPE = -2,      // shape not supported by the lane mapping
  RG_ERR_ALIGN = -3,      // pointer not 16-byte aligned
  RG_ERR_WORKSPACE = -4,  // workspace too small
  RG_ERR_DTYPE = -5,
  RG_ERR_DRIVER = -6,     // driver entry point (tensor-map encode) unavailable
};

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? RG_OK : static_cast<int>(e); }

// ---- 128-bit (or scalar) row-vector access, values widened to fp32 ------------------
template <typename T, int V>
struct RowVec;

template <>
struct RowVec<float, 4> {
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
    // gathered rows are touched ~degree times chip-wide, never twice by one SM: keep them out of L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  }
  static __device__ __forceinline__ void load_cached(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void load_shared(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  // generic pointer (shared or global): plain 128-bit load
  static __device__ __forceinline__ void load_any(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

template <>
struct RowVec<float, 8> {  // fp32 data addressed with the 8-element vectors of the bf16 feature layout
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p + 4));
  }
  static __device__ __forceinline__ void load_cached(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load_shared(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void load_any(const float* p, float (&v)[8]) { load_shared(p, v); }
};

template <>
struct RowVec<float, 1> {
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void load_cached(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
  static __device__ __forceinline__ void load_shared(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void load_any(const float* p, float (&v)[1]) { v[0] = *p; }
};

template <>
struct RowVec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    unpack(t, v);
  }
  static __device__ __forceinline__ void load_cached(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    unpack(t, v);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// ---- fp32 -> bf16 (hi, lo) split stores: x ~= hi + lo with hi = rn(x), lo = rn(x - hi) --------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// (a, b) -> packed bf16x2 with one cvt; also returns the rounded values widened back to fp32
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // low half = a, high half = b
  return r;
}
__device__ __forceinline__ float bf16_lo_as_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_as_f32(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// Lean split store: per PAIR of elements one cvt for hi, two integer ops to widen it back, two
// subtractions and one cvt for lo (ncu: the scalar conversion helpers made the per-destination
// epilogue 35% of the forward kernel's instructions).
template <int V>
__device__ __forceinline__ void store_split_bf16(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[V]) {
  if constexpr (V % 2 == 0) {
    uint32_t h[V / 2], l[V / 2];
#pragma unroll
    for (int i = 0; i < V / 2; ++i) {
      h[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
      l[i] = cvt_bf16x2(v[2 * i] - bf16_lo_as_f32(h[i]), v[2 * i + 1] - bf16_hi_as_f32(h[i]));
    }
    if constexpr (V == 8) {
      *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
      if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
    } else if constexpr (V == 4) {
      *reinterpret_cast<uint2*>(hi) = make_uint2(h[0], h[1]);
      if (lo) *reinterpret_cast<uint2*>(lo) = make_uint2(l[0], l[1]);
    } else {
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        reinterpret_cast<uint32_t*>(hi)[i] = h[i];
        if (lo) reinterpret_cast<uint32_t*>(lo)[i] = l[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float h = bf16_round(v[i]);
      hi[i] = __float2bfloat16_rn(h);
      if (lo) lo[i] = __float2bfloat16_rn(v[i] - h);
    }
  }
}

// exp(x) for x <= 0-ish activations: one multiply + MUFU.EX2 (flush-to-zero, no range fix-up)
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// ---- dropout keep-bit masks (csrc/mask.cu): element i -> word i >> 5, bit i & 31; 1 = keep ------------
// multiplier of V consecutive feature columns starting at col0 (V | 32 on the vector paths, so they share a word)
template <int V>
__device__ __forceinline__ void keep_scale(const uint32_t* __restrict__ row_bits, int col0, float scale, float (&m)[V]) {
  if constexpr (V == 4 || V == 8) {
    const uint32_t w = __ldg(row_bits + (col0 >> 5)) >> (col0 & 31);
#pragma unroll
    for (int v = 0; v < V; ++v) m[v] = ((w >> v) & 1u) ? scale : 0.f;
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = col0 + v;
      m[v] = ((__ldg(row_bits + (c >> 5)) >> (c & 31)) & 1u) ? scale : 0.f;
    }
  }
}
__device__ __forceinline__ float keep_scale1(const uint32_t* __restrict__ bits, long long i, float scale) {
  return ((__ldg(bits + (i >> 5)) >> (i & 31)) & 1u) ? scale : 0.f;
}

// ---- lane mapping ------------------------------------------------------------------
struct LaneMap {
  int hh;        // absolute head owned by this lane
  int sub;       // lane index inside the head's lane set
  int lph;       // lanes per head
  int vph;       // vectors per head (F / V)
  int head_off;  // hh * F (elements)
};

template <int V>
__device__ __forceinline__ LaneMap make_lane_map(int lane, int group, int hg, int F) {
  LaneMap m;
  m.lph = 32 / hg;
  const int hl = lane / m.lph;
  m.sub = lane - hl * m.lph;
  m.hh = group * hg + hl;
  m.vph = F / V;
  m.head_off = m.hh * F;
  return m;
}

// Next chunk for this warp: dynamic claim from a device counter (lane 0 does the atomic, the warp gets the
// value), or the static start index when no counter is given.
__device__ __forceinline__ int claim_chunk(int* counter, int lane, int static_index) {
  if (!counter) return static_index;
  int c = 0;
  if (lane == 0) c = atomicAdd(counter, 1);
  return __shfl_sync(0xffffffffu, c, 0);
}

// Sum over the lanes that own the same head (lph is a power of two <= 32).
__device__ __forceinline__ float head_sum(float x, int lph) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float y = __shfl_xor_sync(0xffffffffu, x, o);
    if (o < lph) x += y;
  }
  return x;
}

// same with a compile-time lane count: only log2(LPH) shuffles are issued
template <int LPH>
__device__ __forceinline__ float head_sum_c(float x) {
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Host side: heads per warp for (H, F, V).  Largest power of two dividing H such that a head's
// F/V vectors fit kMaxVecPerLane per lane and, when R is given, the group's attention vectors
// (hg*R*F fp32) fit the shared-memory budget; if no group fits the budget the largest
// register-feasible group is used and the kernels read A through L1/L2 instead.
// Returns 0 when no mapping exists.
inline int pick_heads_per_warp(int H, int F, int V, int R = 0, size_t budget = kSmemBudgetA) {
  if (H <= 0 || F <= 0 || F % V != 0) return 0;
  const int vph = F / V;
  int best = 0;
  for (int hg = 32; hg >= 1; hg >>= 1) {
    if (H % hg != 0) continue;
    const int lph = 32 / hg;
    if ((vph + lph - 1) / lph > max_vec_rt(V)) continue;
    if (!best) best = hg;
    if (R <= 0 || static_cast<size_t>(hg) * R * F * sizeof(float) <= budget) return hg;
  }
  return best;
}

// experiment knob (host side): RELGAT_<name>_BUDGET_KB overrides the attention-vector smem budget
inline size_t smem_budget_override(const char* name, size_t dflt) {
  const char* v = getenv(name);
  if (!v) return dflt;
  const long kb = atol(v);
  return kb > 0 ? static_cast<size_t>(kb) * 1024 : dflt;
}

inline int vectors_per_lane(int vph, int hg) {
  const int lph = 32 / hg;
  return (vph + lph - 1) / lph;
}

}  // namespace relgat
