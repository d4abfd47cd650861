// Peer tables: one row table per GPU, all of them mapped back to back into every process's address
// space (CUDA virtual memory management), so an edge kernel on GPU a gathers rows owned by GPU b with
// plain loads over NVLink/NVSwitch — no pack / exchange / unpack step and no staging copy.
//
// Replaces the feature all-gather of the destination-range partition (BASELINE.json north_star,
// "transformed source features are exchanged ... over NVLink"; the reference itself is single-device,
// SURVEY.md §2.5).  Host-side entry points; the device work is the unchanged edge kernels reading the
// mapped range.  The driver API is reached through cudaGetDriverEntryPoint (no link-time libcuda).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"

namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  if (q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<Fn>(f);
}

using GranFn = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using CreateFn = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using ExportFn = CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
using ImportFn = CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
using ReserveFn = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using MapFn = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using AccessFn = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using UnmapFn = CUresult (*)(CUdeviceptr, size_t);
using ReleaseFn = CUresult (*)(CUmemGenericAllocationHandle);
using FreeVaFn = CUresult (*)(CUdeviceptr, size_t);

CUmemAllocationProp table_prop(int device) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int g_last_driver_error = 0;  // CUresult of the last failed driver call (diagnostics only)
int drv(CUresult r) {
  if (r == CUDA_SUCCESS) return 0;
  g_last_driver_error = static_cast<int>(r);
  return relgat::RG_ERR_DRIVER;
}

}  // namespace

extern "C" int relgat_peer_table_last_driver_error(void) { return g_last_driver_error; }

extern "C" int relgat_peer_table_granularity(int device, unsigned long long* granularity) {
  if (!granularity) return relgat::RG_ERR_ARG;
  auto fn = driver_fn<GranFn>("cuMemGetAllocationGranularity");
  if (!fn) return relgat::RG_ERR_DRIVER;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return relgat::RG_ERR_DRIVER;
  const CUmemAllocationProp prop = table_prop(device);
  size_t g = 0;
  if (int rc = drv(fn(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED))) return rc;
  *granularity = g;
  return 0;
}

extern "C" int relgat_peer_table_create(int device, unsigned long long bytes, unsigned long long* handle, int* fd) {
  if (!handle || !fd || bytes == 0) return relgat::RG_ERR_ARG;
  auto create = driver_fn<CreateFn>("cuMemCreate");
  auto exportfn = driver_fn<ExportFn>("cuMemExportToShareableHandle");
  if (!create || !exportfn) return relgat::RG_ERR_DRIVER;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return relgat::RG_ERR_DRIVER;
  const CUmemAllocationProp prop = table_prop(device);
  CUmemGenericAllocationHandle h = 0;
  if (int rc = drv(create(&h, bytes, &prop, 0))) return rc;
  int out_fd = -1;
  if (int rc = drv(exportfn(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))) return rc;
  *handle = h;
  *fd = out_fd;
  return 0;
}

// slot s of the mapped range holds the table of rank (rank + s) % world: every process sees its own table first
extern "C" int relgat_peer_table_map(int device, int world, int rank, unsigned long long own_handle,
                                     const int* peer_fds, unsigned long long bytes, void** base) {
  if (!base || world < 1 || rank < 0 || rank >= world || bytes == 0 || (world > 1 && !peer_fds)) return relgat::RG_ERR_ARG;
  auto importfn = driver_fn<ImportFn>("cuMemImportFromShareableHandle");
  auto reserve = driver_fn<ReserveFn>("cuMemAddressReserve");
  auto map = driver_fn<MapFn>("cuMemMap");
  auto access = driver_fn<AccessFn>("cuMemSetAccess");
  auto release = driver_fn<ReleaseFn>("cuMemRelease");
  if (!importfn || !reserve || !map || !access || !release) return relgat::RG_ERR_DRIVER;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return relgat::RG_ERR_DRIVER;
  auto unmap = driver_fn<UnmapFn>("cuMemUnmap");
  auto freeva = driver_fn<FreeVaFn>("cuMemAddressFree");
  if (!unmap || !freeva) return relgat::RG_ERR_DRIVER;
  CUdeviceptr va = 0;
  if (int rc = drv(reserve(&va, bytes * world, 0, 0, 0))) return rc;
  int mapped = 0;
  // every failure path gives back what has been taken so far: the slots already mapped and the reserved range
  auto undo = [&](int rc) {
    for (int s = 0; s < mapped; ++s) unmap(va + static_cast<CUdeviceptr>(s) * bytes, bytes);
    freeva(va, bytes * world);
    return rc;
  };
  for (int s = 0; s < world; ++s) {
    const int owner = (rank + s) % world;
    CUmemGenericAllocationHandle h = own_handle;
    if (owner != rank) {
      if (int rc = drv(importfn(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(peer_fds[owner])),
                                CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR)))
        return undo(rc);
    }
    const int rc = drv(map(va + static_cast<CUdeviceptr>(s) * bytes, bytes, 0, h, 0));
    if (owner != rank) release(h);  // the mapping keeps the peer's allocation alive
    if (rc) return undo(rc);
    ++mapped;
  }
  CUmemAccessDesc desc = {};
  desc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  desc.location.id = device;
  desc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (int rc = drv(access(va, bytes * world, &desc, 1))) return undo(rc);
  *base = reinterpret_cast<void*>(va);
  return 0;
}

extern "C" int relgat_peer_table_unmap(void* base, int world, unsigned long long bytes, unsigned long long own_handle) {
  auto unmap = driver_fn<UnmapFn>("cuMemUnmap");
  auto release = driver_fn<ReleaseFn>("cuMemRelease");
  auto freeva = driver_fn<FreeVaFn>("cuMemAddressFree");
  if (!unmap || !release || !freeva) return relgat::RG_ERR_DRIVER;
  int rc = 0;
  if (base) {
    const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(base);
    rc |= drv(unmap(va, bytes * world));
    rc |= drv(freeva(va, bytes * world));
  }
  if (own_handle) rc |= drv(release(own_handle));
  return rc;
}

// ------------------------------------------------------------------------------------
// halo pull: out[i, :] = table[ids[i], :] — the rows of other ranks that this rank's edges reference,
// read from the mapped peer tables (NVLink) into local memory ahead of the edge kernel.  Each thread
// keeps kPullUnroll independent 16-byte loads in flight: the link is latency-bound (~3 us), the
// kernel needs megabytes in flight to fill it.  (`prefetch.global.L2` on a peer address is NOT an
// option: measured 70x slower than plain loads — profiles/r01_summary.md.)
// ------------------------------------------------------------------------------------
namespace relgat {

static inline bool al16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

constexpr int kPullThreads = 256;

template <int V, int kPullUnroll>
__global__ void __launch_bounds__(kPullThreads)
pull_rows_kernel(const float* __restrict__ table, long long ld, const long long* __restrict__ ids,
                 const long long* __restrict__ out_ids, long long n, int D, float* __restrict__ out, long long ldo) {
  const int dv = D / V;  // vectors per row
  const long long total = n * dv;
  const long long stride = static_cast<long long>(gridDim.x) * kPullThreads;
  for (long long base = static_cast<long long>(blockIdx.x) * kPullThreads + threadIdx.x; base < total;
       base += stride * kPullUnroll) {
    float v[kPullUnroll][V];
    long long row[kPullUnroll];
    int q[kPullUnroll];
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u) {
      const long long idx = base + u * stride;
      row[u] = -1;
      if (idx < total) {
        row[u] = idx / dv;
        q[u] = static_cast<int>(idx - row[u] * dv);
        const long long src_row = __ldg(ids + row[u]);
        if (src_row < 0) row[u] = -1;  // entry switched off by the caller (a row known to be all zero at its owner)
        else RowVec<float, V>::load_stream(table + src_row * ld + q[u] * V, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u)
      if (row[u] >= 0) RowVec<float, V>::store(out + (out_ids ? __ldg(out_ids + row[u]) : row[u]) * ldo + q[u] * V, v[u]);
  }
}

// bf16 rows over the link, fp32 rows in local memory: the owner exports its rows rounded to bf16 (half the NVLink
// bytes of a pull), the puller widens them while storing into the [own | pulled] fp32 table the edge kernels read.
template <int kPullUnroll>
__global__ void __launch_bounds__(kPullThreads)
pull_rows_bf16_kernel(const __nv_bfloat16* __restrict__ table, long long ld, const long long* __restrict__ ids,
                      const long long* __restrict__ out_ids, long long n, int D, float* __restrict__ out, long long ldo) {
  const int dv = D / 8;  // 16-byte pieces of 8 bf16 per row
  const long long total = n * dv;
  const long long stride = static_cast<long long>(gridDim.x) * kPullThreads;
  for (long long base = static_cast<long long>(blockIdx.x) * kPullThreads + threadIdx.x; base < total;
       base += stride * kPullUnroll) {
    float v[kPullUnroll][8];
    long long row[kPullUnroll];
    int q[kPullUnroll];
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u) {
      const long long idx = base + u * stride;
      row[u] = -1;
      if (idx < total) {
        row[u] = idx / dv;
        q[u] = static_cast<int>(idx - row[u] * dv);
        const long long src_row = __ldg(ids + row[u]);
        if (src_row < 0) row[u] = -1;  // entry switched off by the caller
        else RowVec<__nv_bfloat16, 8>::load_stream(table + src_row * ld + q[u] * 8, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kPullUnroll; ++u)
      if (row[u] >= 0) RowVec<float, 8>::store(out + (out_ids ? __ldg(out_ids + row[u]) : row[u]) * ldo + q[u] * 8, v[u]);
  }
}

}  // namespace relgat

extern "C" int relgat_pull_rows_bf16(const void* table, long long ld, const long long* ids, const long long* out_ids,
                                     long long n, int D, float* out, long long ldo, int sm_count, void* stream) {
  using namespace relgat;
  if (n < 0 || D <= 0 || ld < D || ldo < D) return RG_ERR_ARG;
  if (n == 0) return RG_OK;
  if (!table || !ids || !out) return RG_ERR_ARG;
  if (D % 8 != 0 || ld % 8 != 0 || ldo % 4 != 0) return RG_ERR_SHAPE;
  if (!al16(table) || !al16(out)) return RG_ERR_ALIGN;
  const long long total = n * (D / 8);
  const long long per_cta = static_cast<long long>(kPullThreads) * 4;
  const long long want = (total + per_cta - 1) / per_cta;
  const long long cap = static_cast<long long>(sm_count > 0 ? sm_count : 148) * 4;
  const unsigned blocks = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
  static const cudaError_t carve = cudaFuncSetAttribute(pull_rows_bf16_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                        cudaSharedmemCarveoutMaxShared);
  (void)carve;
  pull_rows_bf16_kernel<4><<<blocks, kPullThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(table), ld, ids, out_ids, n, D, out, ldo);
  return cuda_status(cudaGetLastError());
}

extern "C" int relgat_pull_rows(const float* table, long long ld, const long long* ids, const long long* out_ids,
                                long long n, int D, float* out, long long ldo, int sm_count, void* stream) {
  using namespace relgat;
  if (n < 0 || D <= 0 || ld < D || ldo < D) return RG_ERR_ARG;
  if (n == 0) return RG_OK;
  if (!table || !ids || !out) return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = D % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && al16(table) && al16(out);
  static const int unroll = []() {
    const char* v = getenv("RELGAT_PULL_UNROLL");  // experiments; 2, 4 and 8 loads in flight per thread measured alike (link-bound)
    const int u = v ? atoi(v) : 4;
    return (u == 2 || u == 4 || u == 8) ? u : 4;
  }();
  const long long total = n * (vec ? D / 4 : D);
  const long long per_cta = static_cast<long long>(kPullThreads) * unroll;
  const long long want = (total + per_cta - 1) / per_cta;
  // 4 CTAs of 256 threads x 4 loads per SM: ~10 MB in flight (the link needs ~3).  1, 2, 4 and 8 CTAs per SM were
  // swept on 2 GPUs: 2..8 are alike (link-bound), 1 is 20 % slower
  static const int per_sm = []() { const char* v = getenv("RELGAT_PULL_CTAS"); const int c = v ? atoi(v) : 4; return c > 0 && c <= 8 ? c : 4; }();
  const long long cap = static_cast<long long>(sm_count > 0 ? sm_count : 148) * per_sm;
  const unsigned blocks = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
  // same shared-memory carve-out as the tcgen05 GEMM (max shared): kernels with different carve-outs cannot share
  // an SM, and the pull is meant to run beside the GEMM of the next row block
#define RG_PULL(V_, U_)                                                                                     \
  do {                                                                                                      \
    static const cudaError_t carve = cudaFuncSetAttribute(                                                  \
        pull_rows_kernel<V_, U_>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    (void)carve;                                                                                            \
    pull_rows_kernel<V_, U_><<<blocks, kPullThreads, 0, s>>>(table, ld, ids, out_ids, n, D, out, ldo);               \
  } while (0)
  if (vec) {
    if (unroll == 8) RG_PULL(4, 8);
    else if (unroll == 4) RG_PULL(4, 4);
    else RG_PULL(4, 2);
  } else {
    RG_PULL(1, 4);
  }
#undef RG_PULL
  return cuda_status(cudaGetLastError());
}
