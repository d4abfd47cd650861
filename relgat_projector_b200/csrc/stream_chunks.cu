// Work tables of the streaming edge kernels (edge_fwd / edge_bwd_src) built on the device.
//
// A CSR pointer array (segments = destinations for the forward pass, sources for the by-source pass) is cut into
// warp-sized chunks: a run of whole segments covering ~chunk_edges edges and at most chunk_nodes segments; a segment
// longer than long_segment stands alone and is cut into parts of part_edges edges, one chunk per part, merged afterwards
// in part order.  Table layout and chunk order are those of relgat_projector_b200/graph.py::StreamChunks (the host
// formulation used once per full graph; tests compare the two): parts first, in segment and part order, then the
// ordinary chunks in segment order.
//
// This variant exists for the per-step receptive-field blocks (blocks.py), where a table is built every step and the
// ~40 small tensor operations of the host formulation would cost more than the kernels that consume the table: here it
// is three launches and no host read (the caller fetches the three counts together with whatever else it reads back).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace relgat {

struct ChunkScan {
  int ord;    // ordinary chunk heads
  int lng;    // split (long) segments
  int parts;  // parts of split segments
  int edges;  // edges of the covered segments
};

struct ChunkScanSum {
  __device__ __forceinline__ ChunkScan operator()(const ChunkScan& a, const ChunkScan& b) const {
    return ChunkScan{a.ord + b.ord, a.lng + b.lng, a.parts + b.parts, a.edges + b.edges};
  }
};

// rows != nullptr: the table covers the listed segments only (the destinations a batch needs), one chunk per listed
// segment (or its parts); n is then the length of the list
// (n_dev: the list's true length lives on the device — n is its upper bound, entries beyond it count as nothing)
__global__ void chunk_flags_kernel(const int* __restrict__ ptr, const long long* __restrict__ rows, int n,
                                   const int* __restrict__ n_dev, int chunk_edges, int chunk_nodes, int long_segment,
                                   int part_edges, ChunkScan* __restrict__ flags, unsigned char* __restrict__ is_head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_dev && i >= __ldg(n_dev)) {
    is_head[i] = 0;
    flags[i] = ChunkScan{0, 0, 0, 0};
    return;
  }
  const int seg = rows ? static_cast<int>(rows[i]) : i;
  const int start = __ldg(ptr + seg);
  const int deg = __ldg(ptr + seg + 1) - start;
  const bool lng = deg > long_segment;
  bool head = i == 0 || lng || rows != nullptr;
  if (i > 0 && !rows) {
    const int pstart = __ldg(ptr + i - 1);
    head = head  || (start - pstart) > long_segment               // the segment before is a split one
           || (start / chunk_edges) != (pstart / chunk_edges)    // crosses an edge-count boundary
           || (i / chunk_nodes) != ((i - 1) / chunk_nodes);      // segment-count cap
  }
  is_head[i] = head ? 1 : 0;
  flags[i] = ChunkScan{(head && !lng) ? 1 : 0, lng ? 1 : 0, lng ? (deg + part_edges - 1) / part_edges : 0, deg};
}

__global__ void chunk_emit_kernel(const int* __restrict__ ptr, const long long* __restrict__ rows, int n, int chunk_nodes,
                                  int part_edges,
                                  const ChunkScan* __restrict__ flags, const ChunkScan* __restrict__ before,
                                  const unsigned char* __restrict__ is_head, int4* __restrict__ chunks, int max_chunks,
                                  int2* __restrict__ parts, int max_parts, int* __restrict__ long_node,
                                  int* __restrict__ long_part_ptr, int max_long, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ChunkScan last_b = before[n - 1], last_f = flags[n - 1];
  const int n_ord = last_b.ord + last_f.ord, n_long = last_b.lng + last_f.lng, n_parts = last_b.parts + last_f.parts;
  const bool fits = n_ord + n_parts <= max_chunks && n_parts <= max_parts && n_long <= max_long;
  if (i == n - 1) {
    counts[0] = fits ? n_ord + n_parts : -1;
    counts[1] = n_parts;
    counts[2] = n_long;
    counts[3] = last_b.edges + last_f.edges;
    if (fits) long_part_ptr[n_long] = n_parts;
  }
  if (!fits || !is_head[i]) return;
  const ChunkScan b = before[i], f = flags[i];
  const int seg = rows ? static_cast<int>(rows[i]) : i;
  if (f.lng) {
    const int start = __ldg(ptr + seg), end = __ldg(ptr + seg + 1);
    long_node[b.lng] = seg;
    long_part_ptr[b.lng] = b.parts;
    for (int p = 0; p < f.parts; ++p) {
      const int lo = start + p * part_edges;
      parts[b.parts + p] = make_int2(lo, min(lo + part_edges, end));
      chunks[b.parts + p] = make_int4(seg, 1, b.parts + p, 0);
    }
  } else {
    int j = i + 1;
    while (!rows && j < n && j - i < chunk_nodes && !is_head[j]) ++j;
    chunks[n_parts + b.ord] = make_int4(seg, j - i, -1, 0);
  }
}

static size_t align256c(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static size_t chunk_scan_bytes(int n) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveScan(nullptr, tmp, static_cast<const ChunkScan*>(nullptr), static_cast<ChunkScan*>(nullptr),
                                 ChunkScanSum(), ChunkScan{0, 0, 0, 0}, n, static_cast<cudaStream_t>(0));
  return tmp;
}

}  // namespace relgat

using namespace relgat;

extern "C" long long relgat_stream_chunks_workspace_bytes(int n) {
  if (n <= 0) return 256;
  return static_cast<long long>(align256c(chunk_scan_bytes(n)) + 2 * align256c(sizeof(ChunkScan) * static_cast<size_t>(n)) +
                                align256c(static_cast<size_t>(n)));
}

static int stream_chunks_run(const int* ptr, const long long* rows, const int* n_dev, int n, int chunk_edges,
                             int chunk_nodes,
                             int long_segment, int part_edges, int* chunks, int max_chunks, int* parts, int max_parts,
                             int* long_node, int* long_part_ptr, int max_long, int* counts, void* workspace,
                             long long workspace_bytes, void* stream);

extern "C" int relgat_stream_chunks_build(const int* ptr, int n, int chunk_edges, int chunk_nodes, int long_segment,
                                          int part_edges, int* chunks, int max_chunks, int* parts, int max_parts,
                                          int* long_node, int* long_part_ptr, int max_long, int* counts,
                                          void* workspace, long long workspace_bytes, void* stream) {
  return stream_chunks_run(ptr, nullptr, nullptr, n, chunk_edges, chunk_nodes, long_segment, part_edges, chunks, max_chunks, parts,
                           max_parts, long_node, long_part_ptr, max_long, counts, workspace, workspace_bytes, stream);
}

extern "C" int relgat_stream_chunks_for_rows(const int* ptr, const long long* rows, int n_rows, const int* n_rows_dev,
                                             int long_segment,
                                             int part_edges, int* chunks, int max_chunks, int* parts, int max_parts,
                                             int* long_node, int* long_part_ptr, int max_long, int* counts,
                                             void* workspace, long long workspace_bytes, void* stream) {
  if (n_rows > 0 && !rows) return RG_ERR_ARG;
  return stream_chunks_run(ptr, rows, n_rows_dev, n_rows, 1, 1, long_segment, part_edges, chunks, max_chunks, parts, max_parts,
                           long_node, long_part_ptr, max_long, counts, workspace, workspace_bytes, stream);
}

static int stream_chunks_run(const int* ptr, const long long* rows, const int* n_dev, int n, int chunk_edges,
                             int chunk_nodes,
                             int long_segment, int part_edges, int* chunks, int max_chunks, int* parts, int max_parts,
                             int* long_node, int* long_part_ptr, int max_long, int* counts, void* workspace,
                             long long workspace_bytes, void* stream) {
  if (n < 0 || chunk_edges <= 0 || chunk_nodes <= 0 || chunk_nodes > 64 || long_segment <= 0 || part_edges <= 0 ||
      max_chunks < 0 || max_parts < 0 || max_long < 0 || !counts || !long_part_ptr)
    return RG_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(long_part_ptr, 0, sizeof(int), s);
    return cuda_status(e);
  }
  if (!ptr || !chunks || !workspace || (max_parts > 0 && !parts) || (max_long > 0 && !long_node)) return RG_ERR_ARG;
  if (workspace_bytes < relgat_stream_chunks_workspace_bytes(n)) return RG_ERR_WORKSPACE;
  char* w = static_cast<char*>(workspace);
  size_t tmp_bytes = chunk_scan_bytes(n);
  void* tmp = w;
  w += align256c(tmp_bytes);
  ChunkScan* flags = reinterpret_cast<ChunkScan*>(w);
  w += align256c(sizeof(ChunkScan) * static_cast<size_t>(n));
  ChunkScan* before = reinterpret_cast<ChunkScan*>(w);
  w += align256c(sizeof(ChunkScan) * static_cast<size_t>(n));
  unsigned char* is_head = reinterpret_cast<unsigned char*>(w);
  const int th = 256, bl = (n + th - 1) / th;
  chunk_flags_kernel<<<bl, th, 0, s>>>(ptr, rows, n, n_dev, chunk_edges, chunk_nodes, long_segment, part_edges, flags, is_head);
  cudaError_t e = cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, flags, before, ChunkScanSum(), ChunkScan{0, 0, 0, 0}, n, s);
  if (e != cudaSuccess) return cuda_status(e);
  chunk_emit_kernel<<<bl, th, 0, s>>>(ptr, rows, n, chunk_nodes, part_edges, flags, before, is_head,
                                      reinterpret_cast<int4*>(chunks), max_chunks, reinterpret_cast<int2*>(parts),
                                      max_parts, long_node, long_part_ptr, max_long, counts);
  return cuda_status(cudaGetLastError());
}
