// G1/G2/G3 — the dense feature transforms of the RelGAT layer on the 5th-gen tensor cores:
//   G1  P  = X  · Wᵀ   (reference layer.py:220, all heads concatenated: W = cat_h proj[h].weight)
//   G2  dW = dPᵀ · X   (autograd of G1, K20 in SURVEY.md §2.2) — split-K, both operands MN-major
//   G3  dX = dP · W    (layers >= 1)                           — B operand MN-major
//
// One persistent, warp-specialised kernel (sm_100a only):
//   warp 0   : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx)
//   warp 1   : MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, fp32 accum in TMEM)
//   warp 2   : TMEM allocator (512 columns = two accumulator stages of up to 256 columns)
//   warps 4-7: epilogue       (tcgen05.ld 32x32b -> registers -> 128-bit global stores)
//
// fp32-parity mode ("split"): each fp32 operand is carried as two bf16 planes (hi = rn(x),
// lo = rn(x - hi)); the kernel issues hi·hi + hi·lo + lo·hi into the same fp32 accumulator,
// which reproduces an fp32 GEMM to ~2^-16 relative per product (the dropped lo·lo term),
// well inside the 1e-4 budget, at 3 tensor-core passes over ONE smem fill.
#include <cuda.h>

#include "common.cuh"

namespace relgat {

constexpr int kBM = 128;       // UMMA M (cta_group::1)
constexpr int kBK = 64;        // one 128-byte swizzle atom of bf16 along the contiguous dim
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;

struct GemmParams {
  int M, N, K;           // D[M,N] = sum_k A[m,k] * B[n,k]
  int BN;                // N tile (multiple of 16, <= 256)
  int a_mn, b_mn;        // operand major-ness: 0 = K-major (K contiguous), 1 = MN-major
  int split;             // 1: operands carry hi and lo planes (3 MMA passes), 0: single plane
  int stages;
  int splits_k;          // split-K factor (partials written at d + s * d_split_stride)
  int kb_per_split;      // k-blocks per split
  float* d;              // fp32 output (or split-K partials)
  __nv_bfloat16* d_bf16; // bf16 output instead of d (no split-K)
  long long ldd;
  long long d_split_stride;
  int a_tile_bytes, b_tile_bytes;   // per plane
  int b_boxes;           // MN-major B: number of 64-wide boxes per tile
  int staged;            // 1: the epilogue transposes 32x16 (fp32) / 32x32 (bf16) blocks through shared memory
  // fused "backward prep" epilogue of the dX GEMM (replaces relgat_layer_bwd_prep for a hidden layer): the tile is
  // dX = d loss / d act(y); written instead: G = dX * act'(y) * m, and per (row, tile) partial sums of
  // t = <dX * act'(y), y - bias * m> and hsum = sum G for the (at most two) heads the tile touches (BN <= F).
  const float* epi_y;    // [M, N] post-dropout pre-activation rows of the layer below (nullptr = plain GEMM)
  const float* epi_bias; // [M]
  float* epi_tpart;      // [M, n_tiles, 2]
  float* epi_hpart;      // [M, n_tiles, 2]
  const uint32_t* epi_drop_bits;  // keep bits [M, epi_drop_words] or nullptr
  int epi_drop_words;
  float epi_drop_scale;
  int epi_F;             // head width
  int epi_elu;           // act = ELU, else identity
  int dbg_epilogue;      // experiments (RELGAT_GEMM_EPI): 0 normal, 1 = no global stores, 2 = no epilogue work at all
  int tma_store;         // 1: the staged blocks leave shared memory through TMA bulk stores (map_d), not st.global
  // N tiles per work unit (1 or 2).  With 2 the unit is 256 (128) rows x 2·BN columns: both tiles use the SAME A stage —
  // per k-block a CTA receives (128 + 2·BN/cg) operand rows for two tiles' MMAs instead of 2·(128 + BN/cg) — and their
  // accumulators sit side by side in tensor memory (2 x <=256 columns = all of it: the epilogue of a unit is no longer
  // hidden behind the next unit's main loop).
  int nt_unit;
  // k-block depth: 64 (one 128-byte swizzle atom along K), or 32 when BOTH operands are MN-major (their boxes are
  // [bk k-rows x 64 MN-columns], any multiple of 8 rows works): half-size stages, so that two-tile units still get a
  // four-deep pipeline
  int bk;
};

constexpr int kEpiStageBytes = 4 * 32 * 64;  // per epilogue warp: 32 rows x 64 bytes

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of one cluster (one TPC) run ONE 256-row MMA; each CTA stages its own
// 128 rows of A and HALF of the B tile, the leader (cluster rank 0) issues the MMAs and reads both CTAs' shared memory.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory object in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncthreads();  // reconverges every warp of this CTA: the .aligned cluster barrier below needs whole warps
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier of any CTA of the cluster (address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are signalled on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of the pair's MMAs arrives on the barrier at this offset in every CTA of `mask` (cluster ranks)
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// TMA load delivered to the same shared-memory offset of every CTA in `mask`; each receiver's own mbarrier (same offset)
// is credited with the bytes
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {  // pairs with a remote release.cluster arrive
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major  tile [rows][64 bf16]: SBO = 1024 B between 8-row groups, LBO unused (1)
//   MN-major tile [k rows][64 bf16] x boxes: SBO = 1024 B between 8-k groups, LBO = bytes between 64-wide MN boxes
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Coalesced epilogue of one tile for one epilogue warp (TMEM lane quarter `ew`).  tcgen05.ld hands every lane ONE row
// (TMEM lane) and 16 / 32 consecutive columns, so a direct store scatters 16-byte pieces over 32 rows per
// instruction.  Each warp therefore transposes its 32 x 64-byte block through its own 2 KB of shared memory (16-byte
// pieces XOR-swizzled by (row >> 1) & 3: conflict-free both ways) and stores 8 rows x 64 contiguous bytes per
// instruction (whole 32-byte sectors).  Kept out of line so that its registers do not add to the kernel's (the
// by-relation edge kernel runs beside the dW GEMM and needs the register file's other half).
template <bool BF16>
__device__ __noinline__ void epilogue_staged(const GemmParams& p, const CUtensorMap* map_d, uint8_t* stg, uint32_t taddr,
                                             int row0, int col0, int ks, bool empty_k, int lane, int pf_row0,
                                             int pf_col0) {
  const int rsub = lane >> 2, q_rd = lane & 3;
  constexpr int cols_per = BF16 ? 32 : 16;  // columns per 64-byte row piece
  constexpr int elems16 = BF16 ? 8 : 4;     // elements per 16-byte piece
  float epi_t0 = 0.f, epi_t1 = 0.f, epi_h0 = 0.f, epi_h1 = 0.f, epi_b = 0.f;
  if (!BF16 && p.epi_y && p.epi_bias && row0 + lane < p.M) epi_b = __ldg(p.epi_bias + row0 + lane);
  if (!BF16 && p.epi_y && pf_row0 >= 0 && pf_row0 + lane < p.M) {
    // the y rows of this CTA's NEXT tile: pull them into L2 now (one row piece of BN floats per lane), so that the
    // per-chunk loads below pay an L2 hit instead of a DRAM round trip each
    const char* yl = reinterpret_cast<const char*>(p.epi_y + static_cast<long long>(pf_row0 + lane) * p.N + pf_col0);
    for (int b = 0; b < p.BN * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(yl + b));
  }
  for (int c0 = 0; c0 < p.BN; c0 += cols_per) {
    uint32_t r[BF16 ? 32 : 16];
    const int width = min(cols_per, p.BN - c0);  // bf16: BN % 32 may be 16
    if constexpr (BF16) {
      if (width == 32) {
        tmem_ld32(taddr + c0, r);
      } else {
        uint32_t t16[32];
        tmem_ld16(taddr + c0, t16);
#pragma unroll
        for (int v = 0; v < 16; ++v) { r[v] = t16[v]; r[v + 16] = 0u; }
      }
#pragma unroll
      for (int v = 0; v < 16; ++v)
        r[v] = empty_k ? 0u : pack_bf16x2(__uint_as_float(r[2 * v]), __uint_as_float(r[2 * v + 1]));
    } else {
      const int row_me = row0 + lane;
      const int colb = col0 + c0;
      const bool epi_on = p.epi_y && row_me < p.M && colb + 16 <= p.N;
      float y[16];
      if (epi_on) {  // issued ahead of the tensor-memory load: both latencies overlap
        const float4* yp = reinterpret_cast<const float4*>(p.epi_y + static_cast<long long>(row_me) * p.N + colb);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t4 = __ldg(yp + q);
          y[4 * q] = t4.x; y[4 * q + 1] = t4.y; y[4 * q + 2] = t4.z; y[4 * q + 3] = t4.w;
        }
      }
      uint32_t t16[32];
      tmem_ld16(taddr + c0, t16);
#pragma unroll
      for (int v = 0; v < 16; ++v) r[v] = empty_k ? 0u : t16[v];
      if (p.epi_y) {
        // this lane holds dX[row, col .. col+15]; fold in act'(y), the dropout mask and the row sums of backward prep
        if (epi_on) {
          uint32_t keep = 0xffffu;
          if (p.epi_drop_bits)
            keep = (__ldg(p.epi_drop_bits + static_cast<long long>(row_me) * p.epi_drop_words + (colb >> 5)) >> (colb & 31)) & 0xffffu;
          const int split = (col0 / p.epi_F + 1) * p.epi_F;  // first column of the tile's second head
#pragma unroll
          for (int v = 0; v < 16; ++v) {
            const float dx = __uint_as_float(r[v]);
            const float gy = p.epi_elu ? (y[v] > 0.f ? dx : dx * expf(y[v])) : dx;
            const float ms = p.epi_drop_bits ? (((keep >> v) & 1u) ? p.epi_drop_scale : 0.f) : 1.f;
            const float gg = gy * ms;
            const float tv = gy * (y[v] - epi_b * ms);
            if (colb + v < split) { epi_t0 += tv; epi_h0 += gg; } else { epi_t1 += tv; epi_h1 += gg; }
            r[v] = __float_as_uint(gg);
          }
        }
      }
    }
    if (p.tma_store) {
      // the previous block's bulk store must have finished READING the staging buffer before it is overwritten
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();  // the previous block has been read back by every lane
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    const int col = col0 + c0;
    if (p.tma_store) {
      // 32 rows x 64 bytes, laid out exactly as a SWIZZLE_64B box: one TMA store per block; rows / columns beyond the
      // matrix are clipped by the tensor map.  The warp does not wait for the data to reach memory.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && p.dbg_epilogue != 1) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(reinterpret_cast<uint64_t>(map_d)), "r"(col), "r"(row0), "r"(smem_u32(stg)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      continue;
    }
    __syncwarp();
    const bool piece_ok = col + (q_rd + 1) * elems16 <= p.N && q_rd * elems16 < width;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rl = 8 * j + rsub;
      const int row = row0 + rl;
      const uint4 vv = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((q_rd ^ ((rl >> 1) & 3)) << 4));
      if (row < p.M && piece_ok && p.dbg_epilogue != 1) {
        if constexpr (BF16)
          *reinterpret_cast<uint4*>(p.d_bf16 + static_cast<long long>(row) * p.ldd + col + q_rd * 8) = vv;
        else
          *reinterpret_cast<uint4*>(p.d + static_cast<long long>(ks) * p.d_split_stride +
                                    static_cast<long long>(row) * p.ldd + col + q_rd * 4) = vv;
      }
    }
  }
  if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if constexpr (!BF16) {
    if (p.epi_y && row0 + lane < p.M) {
      const int n_tiles = (p.N + p.BN - 1) / p.BN;
      const long long o = (static_cast<long long>(row0 + lane) * n_tiles + col0 / p.BN) * 2;
      p.epi_tpart[o] = epi_t0; p.epi_tpart[o + 1] = epi_t1;
      p.epi_hpart[o] = epi_h0; p.epi_hpart[o + 1] = epi_h1;
    }
  }
}

// ---------------------------------------------------------------------------- the kernel
// CG = 1: one CTA per 128-row tile (cta_group::1).  CG = 2: launched as clusters of two CTAs; the pair owns a 256-row
// tile (rank r: rows r*128..), each CTA loads its A rows and B columns [r*BN/2, (r+1)*BN/2), the leader issues
// tcgen05.mma.cta_group::2 (M = 256) whose completion is multicast to both CTAs' barriers.  Per CTA and k-block this
// moves (128 + BN/2) operand rows through L2 -> shared memory instead of (128 + BN).
//
// MC = 2 (needs CG = 2): clusters of FOUR CTAs = two pairs working on the same 256 rows and two neighbouring N tiles.
// The A rows of CTA r of pair 0 and of CTA r of pair 1 are the same: each of the two loads HALF of them (64 rows) and
// TMA-multicasts its half into both, so every CTA pulls (64 + BN/2) operand rows per k-block through L2 instead of
// (128 + BN/2).  Barriers: every CTA counts the bytes landing in its OWN shared memory; the odd CTA of a pair forwards
// "my stage is full" to its leader (one remote arrive per stage); a stage is free when BOTH pairs have consumed it
// (the commits of both leaders are multicast to all four CTAs).
template <int CG, int MC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                         const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                         const __grid_constant__ CUtensorMap map_d, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], peer_full[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int planes = p.split ? 2 : 1;
  const int ntu = p.nt_unit;
  const int stage_bytes = planes * (p.a_tile_bytes + ntu * p.b_tile_bytes);
  const int n_acc = ntu == 2 ? 1 : 2;  // accumulator stages in tensor memory

  static_assert(MC == 1 || CG == 2, "operand multicast is built on CTA pairs");
  constexpr int kCluster = CG * MC;
  const uint32_t cl_rank = (kCluster > 1) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cl_rank & (CG - 1);     // which 128 rows of the pair's tile
  const uint32_t pair_in_cl = cl_rank >> 1;          // MC = 2: which of the cluster's two N tiles
  const uint32_t leader = cl_rank & ~1u;             // cluster rank of this CTA's MMA issuer
  const int group_id = blockIdx.x / kCluster;        // persistent cluster (or CTA) index
  const int n_groups = gridDim.x / kCluster;
  constexpr int kTileM = kBM * CG;            // rows of one unit
  const int bn_cta = p.BN / CG;               // B rows (output columns) this CTA stages
  const int mt = (p.M + kTileM - 1) / kTileM;
  const int nt_tiles = (p.N + p.BN - 1) / p.BN;
  // N positions of the unit grid (MC = 2: pairs of N tiles, the host checked evenness; ntu = 2: two tiles per unit)
  const int nt = MC == 2 ? nt_tiles / 2 : (nt_tiles + ntu - 1) / ntu;
  const int bk = p.bk;
  const int total_kb = (p.K + bk - 1) / bk;
  const long long units = static_cast<long long>(mt) * nt * p.splits_k;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)));
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)));
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_d)));
    if (p.split) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_lo)));
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_lo)));
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MC); mbar_init(&peer_full[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CG == 2) {  // the same columns are reserved in both CTAs' tensor memory
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long u = group_id; u < units; u += n_groups) {
        const int n_tile = MC == 2 ? static_cast<int>(u % nt) * 2 + static_cast<int>(pair_in_cl) : static_cast<int>(u % nt) * ntu;
        const int tiles_here = MC == 2 ? 1 : min(ntu, nt_tiles - n_tile);
        const int m_tile = static_cast<int>((u / nt) % mt);
        const int ks = static_cast<int>(u / (static_cast<long long>(nt) * mt));
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int a_row0 = m_tile * kTileM + static_cast<int>(cta_rank) * kBM;      // this CTA's A rows
        const int b_row0 = n_tile * p.BN + static_cast<int>(cta_rank) * bn_cta;     // this CTA's share of the (first) B tile
        const int fill_bytes = planes * (p.a_tile_bytes + tiles_here * p.b_tile_bytes);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
          // pair: both CTAs' bytes are counted on the leader's barrier (a peer's bytes may land before the leader's
          // expect_tx of the same phase: the pending arrival keeps the phase open)
          if constexpr (MC == 2) mbar_expect_tx(&full_bar[stage], fill_bytes);  // what lands in THIS CTA's smem
          else if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], fill_bytes * CG);
          const uint32_t bar_leader = (CG == 2 && MC == 1) ? mapa_shared(smem_u32(&full_bar[stage]), 0) : 0u;
          auto load = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if constexpr (CG == 2 && MC == 1) tma_load_2d_pair(dst, m, bar_leader, c0, c1);
            else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
          };
          for (int pl = 0; pl < planes; ++pl) {
            const CUtensorMap* ma = pl ? &map_a_lo : &map_a_hi;
            const CUtensorMap* mb = pl ? &map_b_lo : &map_b_hi;
            uint8_t* sa = st + pl * p.a_tile_bytes;
            if constexpr (MC == 2) {
              // my half (64 rows = one 8 KB box in either layout) of the A rows I share with CTA cta_rank of the
              // other pair, delivered to both of us
              const uint16_t both = static_cast<uint16_t>((1u << cta_rank) | (1u << (cta_rank + 2)));
              uint8_t* dst = sa + pair_in_cl * (bk * 128);
              const int r0 = a_row0 + static_cast<int>(pair_in_cl) * 64;
              if (!p.a_mn) tma_load_2d_mcast(dst, ma, &full_bar[stage], kb * bk, r0, both);
              else tma_load_2d_mcast(dst, ma, &full_bar[stage], r0, kb * bk, both);
            } else if (!p.a_mn) {
              load(sa, ma, kb * bk, a_row0);
            } else {
              for (int bx = 0; bx < kBM / 64; ++bx) load(sa + bx * (bk * 128), ma, a_row0 + bx * 64, kb * bk);
            }
            for (int t = 0; t < tiles_here; ++t) {  // the unit's B tiles, one after the other behind the A planes
              uint8_t* sb = st + planes * p.a_tile_bytes + (t * planes + pl) * p.b_tile_bytes;
              const int br = b_row0 + t * p.BN;
              if (!p.b_mn) {
                load(sb, mb, kb * bk, br);
              } else {
                for (int bx = 0; bx < p.b_boxes; ++bx) load(sb + bx * (bk * 128), mb, br + bx * 64, kb * bk);
              }
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.a_mn) << 15) |
                             (static_cast<uint32_t>(p.b_mn) << 16) | (static_cast<uint32_t>(p.BN >> 3) << 17) |
                             (static_cast<uint32_t>(kTileM >> 4) << 24);
      const uint32_t a_lbo = p.a_mn ? bk * 128 : 16, b_lbo = p.b_mn ? bk * 128 : 16;
      // descriptor address-field increment per UMMA_K step (bytes >> 4)
      const uint64_t a_dstep = (p.a_mn ? kUmmaK * 128 : kUmmaK * 2) >> 4;
      const uint64_t b_dstep = (p.b_mn ? kUmmaK * 128 : kUmmaK * 2) >> 4;
      const int ksteps = bk / kUmmaK;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (long long u = group_id; u < units; u += n_groups) {
        const int ks = static_cast<int>(u / (static_cast<long long>(nt) * mt));
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int tiles_here = MC == 2 ? 1 : min(ntu, nt_tiles - static_cast<int>(u % nt) * ntu);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if constexpr (MC == 2) mbar_wait_cluster(&peer_full[stage], phase);  // the odd CTA's stage, forwarded
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          // The ONE issuing thread must stay ahead of the tensor pipe (an MMA of this size retires in 80-130 clocks): the
          // descriptors of a stage are built once, the k-steps add a constant to their address field (the smem window
          // is < 256 KB, so the 14-bit field never carries), and the 3 x 4 issue sequence is fully unrolled.  (A
          // run-time trip count here cost 15 % on the 300k-row GEMMs.)
          const int passes = p.split ? 3 : 1;
          const uint64_t da_hi = make_desc(st, a_lbo, 1024), da_lo = make_desc(st + p.a_tile_bytes, a_lbo, 1024);
          const uint32_t first_acc = kb > kb0 ? 1u : 0u;  // the unit's very first MMA overwrites the accumulator
          for (int t = 0; t < tiles_here; ++t) {  // the unit's N tiles: same A stage, accumulators side by side
            const uint32_t tmem_d = tmem_base + (acc + t) * 256;
            const uint32_t sb_hi = st + planes * p.a_tile_bytes + t * planes * p.b_tile_bytes;
            const uint64_t db_hi = make_desc(sb_hi, b_lbo, 1024), db_lo = make_desc(sb_hi + p.b_tile_bytes, b_lbo, 1024);
#pragma unroll
            for (int ps = 0; ps < 3; ++ps) {
              if (ps < passes) {
                const uint64_t da = (ps == 2) ? da_lo : da_hi;   // hi·hi, hi·lo, lo·hi
                const uint64_t db = (ps == 1) ? db_lo : db_hi;
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                  if (k < ksteps) {
                    const uint32_t accumulate = (ps | k) ? 1u : first_acc;
                    if constexpr (CG == 2) umma_bf16_pair(tmem_d, da + k * a_dstep, db + k * b_dstep, idesc, accumulate);
                    else umma_bf16(tmem_d, da + k * a_dstep, db + k * b_dstep, idesc, accumulate);
                  }
                }
              }
            }
          }
          // frees the smem stage (in both CTAs of a pair) when the MMAs above retire
          if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage], MC == 2 ? 0xF : 0x3); else umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0) {
          // empty K range (possible for the last split): publish a zero tile via the epilogue flag
        }
        if constexpr (CG == 2) umma_commit_pair(&tmem_full[acc], static_cast<uint16_t>(0x3u << leader)); else umma_commit(&tmem_full[acc]);
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (MC == 2 && warp == 2 && cta_rank == 1) {
    // ===================== odd CTA of a pair: tell the leader when a stage of MY shared memory is full ==========
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long u = group_id; u < units; u += n_groups) {
        const int ks = static_cast<int>(u / (static_cast<long long>(nt) * mt));
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          mbar_arrive_cluster(mapa_shared(smem_u32(&peer_full[stage]), leader));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    for (long long u = group_id; u < units; u += n_groups) {
      const int n_tile0 = MC == 2 ? static_cast<int>(u % nt) * 2 + static_cast<int>(pair_in_cl) : static_cast<int>(u % nt) * ntu;
      const int tiles_here = MC == 2 ? 1 : min(ntu, nt_tiles - n_tile0);
      const int m_tile = static_cast<int>((u / nt) % mt);
      const int ks = static_cast<int>(u / (static_cast<long long>(nt) * mt));
      const bool empty_k = (ks * p.kb_per_split >= total_kb);
      const int row_base = m_tile * kTileM + static_cast<int>(cta_rank) * kBM + ew * 32;  // this warp's 32 rows
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int t = 0; t < tiles_here; ++t) {
      const int n_tile = n_tile0 + t;
      const uint32_t taddr = tmem_base + (acc + t) * 256 + (static_cast<uint32_t>(ew * 32) << 16);
      if (p.dbg_epilogue == 2) {
        // measurement only: hand the accumulator straight back (mainloop-only time)
      } else if (p.staged) {
        uint8_t* stg = smem + static_cast<size_t>(p.stages) * stage_bytes + ew * (32 * 64);
        int pf_row0 = -1, pf_col0 = 0;
        if (p.epi_y && u + n_groups < units) {
          const long long un = u + n_groups;
          pf_row0 = static_cast<int>((un / nt) % mt) * kTileM + static_cast<int>(cta_rank) * kBM + ew * 32;
          pf_col0 = static_cast<int>(un % nt) * p.BN;
        }
        if (p.d_bf16) epilogue_staged<true>(p, &map_d, stg, taddr, row_base, n_tile * p.BN, ks, empty_k, lane, -1, 0);
        else epilogue_staged<false>(p, &map_d, stg, taddr, row_base, n_tile * p.BN, ks, empty_k, lane, pf_row0, pf_col0);
      } else {
      const int row = row_base + lane;
      float* drow = p.d ? p.d + static_cast<long long>(ks) * p.d_split_stride + static_cast<long long>(row) * p.ldd
                        : nullptr;
      __nv_bfloat16* hrow = p.d_bf16 ? p.d_bf16 + static_cast<long long>(row) * p.ldd : nullptr;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        const int width = min(32, p.BN - c0);   // BN is a multiple of 16: width is 32 or 16
        if (width == 32) tmem_ld32(taddr + c0, r); else tmem_ld16(taddr + c0, r);
        const int col = n_tile * p.BN + c0;
        if (row < p.M) {
          if (hrow) {  // bf16 output: 8 elements per 128-bit store
            if (col + width <= p.N && (p.ldd & 7) == 0) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                if (v * 8 < width) {
                  uint4 o;
                  o.x = empty_k ? 0u : pack_bf16x2(__uint_as_float(r[8 * v]), __uint_as_float(r[8 * v + 1]));
                  o.y = empty_k ? 0u : pack_bf16x2(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]));
                  o.z = empty_k ? 0u : pack_bf16x2(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]));
                  o.w = empty_k ? 0u : pack_bf16x2(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]));
                  *reinterpret_cast<uint4*>(hrow + col + 8 * v) = o;
                }
              }
            } else {
#pragma unroll
              for (int v = 0; v < 32; ++v)
                if (v < width && col + v < p.N) hrow[col + v] = __float2bfloat16_rn(empty_k ? 0.f : __uint_as_float(r[v]));
            }
          } else if (col + width <= p.N && (p.ldd & 3) == 0) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              if (v * 4 < width) {
                float4 o = empty_k ? make_float4(0.f, 0.f, 0.f, 0.f)
                                   : make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                 __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                *reinterpret_cast<float4*>(drow + col + 4 * v) = o;
              }
            }
          } else {
#pragma unroll
            for (int v = 0; v < 32; ++v)
              if (v < width && col + v < p.N) drow[col + v] = empty_k ? 0.f : __uint_as_float(r[v]);
          }
        }
      }
      }
      }  // tiles of the unit
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {  // the leader's issuer waits for the epilogue warps of both CTAs
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), leader));
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (p.tma_store && warp >= 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while its peer still works
  else __syncthreads();
  if (warp == 2) {
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ordered split-K reduction: out[i] = sum_s part[s, i]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int splits,
                                     long long stride) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[k * stride + i];
  out[i] = s;
}

// t[j, h], hsum[j, h] from the per-tile partials of the fused prep epilogue, tiles in ascending order (fixed order)
__global__ void prep_combine_kernel(const float* __restrict__ tpart, const float* __restrict__ hpart, float* __restrict__ t,
                                    float* __restrict__ hsum, long long rows, int n_tiles, int BN, int F, int H) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const long long j = i / H;
  const int h = static_cast<int>(i - j * H);
  float ts = 0.f, hs = 0.f;
  for (int tl = 0; tl < n_tiles; ++tl) {
    const int h0 = (tl * BN) / F;  // head of the tile's first column; slot 1 is head h0 + 1
    const long long o = (j * n_tiles + tl) * 2;
    if (h0 == h) { ts += tpart[o]; hs += hpart[o]; }
    else if (h0 + 1 == h) { ts += tpart[o + 1]; hs += hpart[o + 1]; }
  }
  t[i] = ts;
  hsum[i] = hs;
}

// fp32 -> bf16 hi (+ lo residual) planes
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    store_split_bf16<4>(hi + i, lo ? lo + i : nullptr, f);
  } else {
    for (long long k = i; k < n; ++k) {
      const float f[1] = {x[k]};
      store_split_bf16<1>(hi + k, lo ? lo + k : nullptr, f);
    }
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 matrix [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle
static int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return RG_ERR_DRIVER;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RG_OK : RG_ERR_DRIVER;
}

// output matrix [rows, cols] (fp32 or bf16, row stride ld elements): box = 32 rows x 64 bytes, 64-byte swizzle — the
// layout the epilogue warps stage their blocks in
static int make_map_d(CUtensorMap* m, const void* base, bool bf16, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return RG_ERR_DRIVER;
  const int esz = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(64 / esz), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RG_OK : RG_ERR_DRIVER;
}

// N tile of the fused-prep dX GEMM (single CTA per tile): the widest divisor of N, so that a tile stays inside two heads
static int pick_bn_prep(int N) {
  if (const char* v = getenv("RELGAT_GEMM_BN")) {  // experiment knob: N tile (multiple of 16, <= 256)
    const int bn = atoi(v);
    if (bn >= 16 && bn <= 256 && bn % 16 == 0 && N > bn) return bn;
  }
  if (N <= 256) return (N + 15) / 16 * 16;
  for (int bn = 256; bn >= 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;
}

static int gemm_cta_group(int M, bool prep_epilogue, int sm_count) {
  if (const char* v = getenv("RELGAT_GEMM_CG")) { if (atoi(v) == 1) return 1; }
  return (M > kBM && !prep_epilogue && sm_count >= 2) ? 2 : 1;
}

// Cost of ONE k-block of a CTA's tile in SM clocks: the slower of the MMA pipe (fp32-parity mode: 12 instructions of
// 128 x bn / 256 clocks) and the L2 -> shared-memory operand stream.  Measured (tools/gemm_sweep.py, 300k x 800 x 1024,
// N tiles 128 / 160 / 208 / 256: 1.42 / 1.18 / 1.07 / 1.14 ms): the two-plane GEMMs run at the pace of that stream,
// ≈ 30 bytes per clock and SM (8.5 TB/s over the chip), not of the tensor pipe — so tiles are chosen by bytes moved.
static double kblock_cost(int bn, int cg, int mc, int b_mn, int ntu = 1) {
  const int bn_cta = bn / cg;
  const double a_bytes = 2.0 * kBM * kBK * 2 / mc;  // multicast: each CTA fetches half of its A rows
  const double b_bytes = 2.0 * (b_mn ? ((bn_cta + 63) / 64) * kBK * 128 : bn_cta * kBK * 2);
  // ntu N tiles of a unit share the A stage
  const double mma = 6.0 * bn * ntu, l2 = (a_bytes + ntu * b_bytes) / 30.0;
  return mma > l2 ? mma : l2;
}

struct GemmPlan {
  int cg;       // CTAs per 128-row-pair tile: 2 = tcgen05 cta_group::2
  int mc;       // 2: clusters of two pairs on neighbouring N tiles, A halves multicast
  int bn;       // N tile
  int ntu;      // N tiles per work unit (2: they share the A stage, accumulators side by side in tensor memory)
  double cost;  // modelled SM clocks of one k-block over all tiles
};

// Clusters of four CTAs cannot use every SM (GPCs hold 16 / 18 / 20 SMs): the kernel reports what fits (132 of 148 on
// B200); the plan charges the stranded SMs to the multicast variant.
constexpr double kClusterOf4SmShare = 132.0 / 148.0;

static GemmPlan plan_gemm(int M, int N, int b_mn, bool prep_epilogue, int sm_count, bool bf16_out = false) {
  GemmPlan g{};
  g.cg = gemm_cta_group(M, prep_epilogue, sm_count);
  g.mc = 1;
  g.ntu = 1;
  const long long mt = (M + kBM * g.cg - 1) / (kBM * g.cg);
  if (prep_epilogue) { g.bn = pick_bn_prep(N); g.cost = static_cast<double>(mt * ((N + g.bn - 1) / g.bn)) * g.cg * kblock_cost(g.bn, g.cg, 1, b_mn); return g; }
  int forced_bn = 0;
  if (const char* v = getenv("RELGAT_GEMM_BN")) {  // experiment knob: N tile (multiple of 16, <= 256)
    const int bn = atoi(v);
    if (bn >= 16 && bn <= 256 && bn % 16 == 0 && N > bn) forced_bn = bn;
  }
  // Operand multicast over clusters of two pairs is OFF unless RELGAT_GEMM_MC=2: measured slower on every config-2
  // shape (300k x 800 x 1024: 1.19 vs 1.07 ms; [dP|dS]^T X: 1.47 vs 1.22 ms) — the bytes each SM receives, not the bytes
  // L2 sends, set the pace, so sharing the fetch saves nothing and clusters of four leave 16 of the 148 SMs idle.
  bool mc_ok = false;
  if (const char* v = getenv("RELGAT_GEMM_MC")) { mc_ok = atoi(v) == 2 && g.cg == 2 && sm_count >= 4; }
  // two N tiles per unit (RELGAT_GEMM_NTU=1 switches it off): needs two pipeline stages of A + 2 B tiles in shared memory
  bool ntu_ok = true;
  if (const char* v = getenv("RELGAT_GEMM_NTU")) { if (atoi(v) == 1) ntu_ok = false; }
  bool have = false;
  auto consider = [&](int bn, int mc) {
    const long long nt = (N + bn - 1) / bn;
    if (mc == 2 && (nt % 2 != 0 || mt * (nt / 2) < 8)) return;  // pairs of N tiles; not worth a cluster launch for a handful of tiles
    double c = static_cast<double>(mt * nt) * g.cg * kblock_cost(bn, g.cg, mc, b_mn);
    if (mc == 2) c /= kClusterOf4SmShare;
    if (!have || c < g.cost * 0.999) { have = true; g.bn = bn; g.mc = mc; g.ntu = 1; g.cost = c; }
    // measured (tools/gemm_sweep.py): the weight-gradient shapes gain 3-4 % (1.255 -> 1.209 ms), the 300k-row GEMMs with
    // a full-size output lose 6 % (their epilogue is exposed and only two pipeline stages fit): few-row-tile shapes only
    if (mc == 1 && ntu_ok && nt >= 2 && mt <= 16) {
      const int bn_cta = bn / g.cg;
      const long long b_tile = b_mn ? ((bn_cta + 63) / 64) * kBK * 128 : bn_cta * kBK * 2;
      const long long stage = 2 * (kBM * kBK * 2 + 2 * b_tile);  // two planes
      if (2 * stage <= 227 * 1024 - 2048 - kEpiStageBytes) {
        // (+4 % for the unit's epilogue, no longer hidden behind the next unit's main loop)
        const double c2 = static_cast<double>(mt * ((nt + 1) / 2)) * g.cg * kblock_cost(bn, g.cg, 1, b_mn, 2) * 1.04;
        if (c2 < g.cost * 0.999) { g.bn = bn; g.mc = 1; g.ntu = 2; g.cost = c2; }
      }
    }
  };
  if (forced_bn) { if (mc_ok) consider(forced_bn, 2); consider(forced_bn, 1); return g; }
  if (N <= 256) { consider((N + 15) / 16 * 16, 1); return g; }
  // bf16 output: the epilogue's TMA store boxes are 32 columns wide, a 16-column tail would spill into the next tile
  const int step = bf16_out ? 32 : 16;
  for (int bn = 256; bn >= 128; bn -= step) {  // widest first: ties go to the wider tile
    if (mc_ok) consider(bn, 2);
    consider(bn, 1);
  }
  return g;
}

// how many clusters of `cluster` CTAs of this kernel the device runs at once (cached per cluster size)
template <int CG, int MC>
static int max_active_clusters(int smem_bytes, int sm_count) {
  static int cached = -1;
  if (cached > 0) return cached;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(sm_count / (CG * MC) * (CG * MC)));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * MC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_bf16_tcgen05_kernel<CG, MC>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = static_cast<int>(sm_count / (CG * MC) * (MC == 2 ? kClusterOf4SmShare : 1.0) + 0.5);
    return n > 0 ? n : 1;  // not cached: the query may succeed later
  }
  cached = n;
  return n;
}

}  // namespace relgat

using namespace relgat;

extern "C" int relgat_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream) {
  if (!x || !hi || n < 0) return RG_ERR_ARG;
  if (n == 0) return RG_OK;
  if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(hi) % 8 || (lo && reinterpret_cast<uintptr_t>(lo) % 8))
    return RG_ERR_ALIGN;
  const int th = 256;
  const long long blocks = ((n + 3) / 4 + th - 1) / th;
  split_bf16_kernel<<<static_cast<unsigned>(blocks), th, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n);
  return cuda_status(cudaGetLastError());
}

extern "C" long long relgat_gemm_workspace_bytes(int M, int N, int K, int a_mn, int b_mn, int splits_k) {
  (void)K; (void)a_mn; (void)b_mn;
  if (splits_k <= 1) return 0;
  return static_cast<long long>(splits_k) * M * N * 4;
}

// D[M,N] (fp32, row stride ldd) = A · Bᵀ with bf16 operands, fp32 accumulation.
//   a_mn == 0: A is [M, K] row-major (stride lda);  a_mn == 1: A is stored [K, M] row-major (stride lda)
//   b_mn == 0: B is [N, K] row-major (stride ldb);  b_mn == 1: B is stored [K, N] row-major (stride ldb)
//   a_lo / b_lo: residual planes of the fp32 split (both or neither); nullptr = plain bf16 GEMM
struct PrepEpilogue {
  const float* y = nullptr;
  const float* bias = nullptr;
  float* tpart = nullptr;
  float* hpart = nullptr;
  const uint32_t* drop_bits = nullptr;
  int drop_words = 0;
  float drop_scale = 1.f;
  int F = 0, elu = 0;
};

static int gemm_launch(const void* a_hi, const void* a_lo, long long lda, int a_mn,
                       const void* b_hi, const void* b_lo, long long ldb, int b_mn,
                       void* d_out, int d_is_bf16, long long ldd, int M, int N, int K, int splits_k,
                       void* workspace, long long workspace_bytes, int sm_count, void* stream, const PrepEpilogue* epi) {
  if (!a_hi || !b_hi || !d_out || M <= 0 || N <= 0 || K <= 0 || splits_k < 1) return RG_ERR_ARG;
  if (d_is_bf16 && splits_k > 1) return RG_ERR_ARG;  // split-K partials are fp32
  float* d = d_is_bf16 ? nullptr : static_cast<float*>(d_out);
  if ((a_lo == nullptr) != (b_lo == nullptr)) return RG_ERR_ARG;
  if (lda % 8 || ldb % 8) return RG_ERR_ALIGN;  // TMA: 16-byte global strides
  if (reinterpret_cast<uintptr_t>(a_hi) % 16 || reinterpret_cast<uintptr_t>(b_hi) % 16 ||
      reinterpret_cast<uintptr_t>(a_lo) % 16 || reinterpret_cast<uintptr_t>(b_lo) % 16)
    return RG_ERR_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  if (sm_count <= 0) sm_count = 148;
  // CTA pairs (cta_group::2, 256-row tiles) whenever there is more than one 128-row tile; the fused-prep epilogue keeps
  // the single-CTA form.  RELGAT_GEMM_CG=1 is the A/B knob.
  const GemmPlan plan = plan_gemm(M, N, b_mn ? 1 : 0, epi != nullptr, sm_count, d_is_bf16 != 0);
  const int cg = plan.cg, mc = plan.mc;
  p.BN = plan.bn;
  p.nt_unit = plan.ntu;
  p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.split = a_lo ? 1 : 0;
  // half-depth k-blocks for the two-tile units of the MN-major (weight-gradient) GEMMs (four pipeline stages instead
  // of two) are opt-in (RELGAT_GEMM_BK32=1): measured SLOWER (1.43 vs 1.27 ms on [dP|dS]^T X) — twice as many, half as
  // large TMA boxes per byte cost more than the deeper pipeline returns
  p.bk = (p.nt_unit == 2 && p.a_mn && p.b_mn && getenv("RELGAT_GEMM_BK32")) ? 32 : kBK;
  const int total_kb = (K + p.bk - 1) / p.bk;
  if (splits_k > total_kb) splits_k = total_kb;
  p.splits_k = splits_k;
  p.kb_per_split = (total_kb + splits_k - 1) / splits_k;
  const int bn_cta = p.BN / cg;  // BN is a multiple of 16: each CTA's share is a multiple of 8 rows
  p.a_tile_bytes = kBM * p.bk * 2;
  p.b_boxes = (bn_cta + 63) / 64;
  p.b_tile_bytes = p.b_mn ? p.b_boxes * p.bk * 128 : bn_cta * p.bk * 2;
  const int planes = p.split ? 2 : 1;
  const int stage_bytes = planes * (p.a_tile_bytes + p.nt_unit * p.b_tile_bytes);
  // the coalesced epilogue needs whole 16-byte pieces: N and the row stride multiples of the piece width
  const int piece = d_is_bf16 ? 8 : 4;
  p.staged = (N % piece == 0 && (splits_k > 1 ? N : ldd) % piece == 0 &&
              reinterpret_cast<uintptr_t>(splits_k > 1 ? workspace : d_out) % 16 == 0) ? 1 : 0;
  if (getenv("RELGAT_GEMM_DIRECT_EPILOGUE")) p.staged = 0;  // experiment knob: the old one-row-per-lane stores
  if (const char* v = getenv("RELGAT_GEMM_EPI")) p.dbg_epilogue = atoi(v);  // measurement knob, results are wrong
  if (epi) {
    // the fused prep epilogue needs: fp32 output rows of width N (ldd == N), whole 16-column pieces, a tile inside
    // at most two heads, no split-K
    if (!p.staged || d_is_bf16 || splits_k != 1 || ldd != N || N % 16 != 0 || epi->F <= 0 || p.BN > epi->F ||
        N % epi->F != 0 || !epi->y || !epi->tpart || !epi->hpart)
      return RG_ERR_SHAPE;
    p.epi_y = epi->y; p.epi_bias = epi->bias; p.epi_tpart = epi->tpart; p.epi_hpart = epi->hpart;
    p.epi_drop_bits = epi->drop_bits; p.epi_drop_words = epi->drop_words; p.epi_drop_scale = epi->drop_scale;
    p.epi_F = epi->F; p.epi_elu = epi->elu;
  }
  const int smem_budget = 227 * 1024 - 2048 - (p.staged ? kEpiStageBytes : 0);
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* v = getenv("RELGAT_GEMM_STAGES")) {
    const int st = atoi(v);
    if (st >= 2 && st <= stages) stages = st;
  }
  if (stages < 2) return RG_ERR_SHAPE;
  p.stages = stages;
  if (splits_k > 1) {
    if (!workspace || workspace_bytes < relgat_gemm_workspace_bytes(M, N, K, a_mn, b_mn, splits_k)) return RG_ERR_WORKSPACE;
    p.d = static_cast<float*>(workspace);
    p.ldd = N;
    p.d_split_stride = static_cast<long long>(M) * N;
  } else {
    p.d = d; p.d_bf16 = d_is_bf16 ? static_cast<__nv_bfloat16*>(d_out) : nullptr; p.ldd = ldd; p.d_split_stride = 0;
  }
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  // K-major: matrix [MN rows, K cols], box rows = tile MN extent.  MN-major: matrix [K rows, MN cols], box rows = BK.
  const long long a_rows = p.a_mn ? K : M, a_cols = p.a_mn ? M : K;
  const long long b_rows = p.b_mn ? K : N, b_cols = p.b_mn ? N : K;
  const int a_box = p.a_mn ? p.bk : (mc == 2 ? kBM / 2 : kBM), b_box = p.b_mn ? p.bk : bn_cta;  // multicast: half tiles
  if ((rc = make_map(&ma_hi, a_hi, a_rows, a_cols, lda, a_box)) != RG_OK) return rc;
  if ((rc = make_map(&mb_hi, b_hi, b_rows, b_cols, ldb, b_box)) != RG_OK) return rc;
  if (p.split) {
    if ((rc = make_map(&ma_lo, a_lo, a_rows, a_cols, lda, a_box)) != RG_OK) return rc;
    if ((rc = make_map(&mb_lo, b_lo, b_rows, b_cols, ldb, b_box)) != RG_OK) return rc;
  } else {
    ma_lo = ma_hi; mb_lo = mb_hi;
  }
  // TMA bulk stores for the staged epilogue of the (non split-K) GEMMs with a large output: measured on the P / dX
  // GEMMs, the st.global stream of the epilogue warps, not the MMA, set the pace (1.62 ms with stores, 1.15 without)
  CUtensorMap md = ma_hi;
  p.tma_store = 0;
  if (p.staged && splits_k == 1 && !epi && !getenv("RELGAT_GEMM_NO_TMA_STORE") && (ldd * (d_is_bf16 ? 2 : 4)) % 16 == 0) {
    if (make_map_d(&md, d_out, d_is_bf16 != 0, M, N, ldd) == RG_OK) p.tma_store = 1;
  }
  const int tile_m = kBM * cg;
  const long long n_pos = mc == 2 ? ((N + p.BN - 1) / p.BN) / 2 : (((N + p.BN - 1) / p.BN) + p.nt_unit - 1) / p.nt_unit;
  const long long units = static_cast<long long>((M + tile_m - 1) / tile_m) * n_pos * splits_k;
  const int smem_bytes = stages * stage_bytes + 1024 + (p.staged ? kEpiStageBytes : 0);
  cudaError_t e;
  if (cg == 2) {
    auto kernel = mc == 2 ? gemm_bf16_tcgen05_kernel<2, 2> : gemm_bf16_tcgen05_kernel<2, 1>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cuda_status(e);
    const int slots = mc == 2 ? max_active_clusters<2, 2>(smem_bytes, sm_count) : sm_count / 2;
    const int groups = static_cast<int>(units < slots ? units : slots);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(groups * cg * mc));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // the two CTAs of a pair share one TPC
    attr[0].val.clusterDim.x = static_cast<unsigned>(cg * mc); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, ma_hi, ma_lo, mb_hi, mb_lo, md, p);
    if (e != cudaSuccess) return cuda_status(e);
  } else {
    const int groups = static_cast<int>(units < sm_count ? units : sm_count);
    e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cuda_status(e);
    gemm_bf16_tcgen05_kernel<1, 1><<<groups, kGemmThreads, smem_bytes, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, md, p);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_status(e);
  if (splits_k > 1) {
    const long long n = static_cast<long long>(M) * N;
    if (ldd != N) return RG_ERR_SHAPE;
    splitk_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p.d, d, n, splits_k, p.d_split_stride);
    return cuda_status(cudaGetLastError());
  }
  return RG_OK;
}

extern "C" int relgat_gemm_bf16(const void* a_hi, const void* a_lo, long long lda, int a_mn,
                                const void* b_hi, const void* b_lo, long long ldb, int b_mn,
                                void* d_out, int d_is_bf16, long long ldd, int M, int N, int K, int splits_k,
                                void* workspace, long long workspace_bytes, int sm_count, void* stream) {
  return gemm_launch(a_hi, a_lo, lda, a_mn, b_hi, b_lo, ldb, b_mn, d_out, d_is_bf16, ldd, M, N, K, splits_k, workspace,
                     workspace_bytes, sm_count, stream, nullptr);
}

extern "C" int relgat_gemm_tile_n(int N) { return N > 0 ? pick_bn_prep(N) : RG_ERR_ARG; }

// Tile shape relgat_gemm_bf16 uses for an [M, N] output (b_mn: B stored [K, N]) and the modelled cost of one k-block
// over all tiles (SM clocks; comparable between the two orientations of a weight-gradient GEMM).
extern "C" long long relgat_gemm_plan(int M, int N, int b_mn, int sm_count, int* tile_m, int* tile_n, int* slots) {
  if (M <= 0 || N <= 0) return RG_ERR_ARG;
  if (sm_count <= 0) sm_count = 148;
  const GemmPlan g = plan_gemm(M, N, b_mn ? 1 : 0, false, sm_count);
  if (tile_m) *tile_m = kBM * g.cg;
  if (tile_n) *tile_n = g.bn * g.mc * g.ntu;  // a unit may span two N tiles
  if (slots) *slots = g.mc == 2 ? static_cast<int>(sm_count / 4 * kClusterOf4SmShare + 0.5) : sm_count / g.cg;
  return static_cast<long long>(g.cost);
}

// dX GEMM (A = dP [M, K] planes, B = W^T [N, K] planes, both K-major) with the backward prep of the layer below fused
// into its epilogue: G [M, N] = (A·B^T) * act'(y) * m is written instead of dX, t / hsum [M, H] follow from per-tile
// partials (tpart / hpart: float [M, n_tiles, 2] scratch, n_tiles = ceil(N / relgat_gemm_tile_n(N))).
extern "C" int relgat_gemm_dx_prep(const void* a_hi, const void* a_lo, long long lda, const void* b_hi, const void* b_lo,
                                   long long ldb, float* G, int M, int N, int K, const float* y, const float* bias,
                                   const unsigned int* drop_bits, int drop_words, float drop_scale, int H, int F,
                                   int apply_elu, float* tpart, float* hpart, float* t, float* hsum, int sm_count,
                                   void* stream) {
  if (!G || !y || !tpart || !hpart || !t || !hsum || H <= 0 || F <= 0 || N != H * F) return RG_ERR_ARG;
  if (drop_bits && drop_words * 32 < N) return RG_ERR_ARG;
  PrepEpilogue e;
  e.y = y; e.bias = bias; e.tpart = tpart; e.hpart = hpart; e.drop_bits = drop_bits; e.drop_words = drop_words;
  e.drop_scale = drop_scale; e.F = F; e.elu = apply_elu;
  int rc = gemm_launch(a_hi, a_lo, lda, 0, b_hi, b_lo, ldb, 0, G, 0, N, M, N, K, 1, nullptr, 0, sm_count, stream, &e);
  if (rc != RG_OK) return rc;
  const int bn = pick_bn_prep(N);
  const long long total = static_cast<long long>(M) * H;
  prep_combine_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tpart, hpart, t, hsum, M, (N + bn - 1) / bn, bn, F, H);
  return cuda_status(cudaGetLastError());
}
