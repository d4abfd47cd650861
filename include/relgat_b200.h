/* relgat_b200.h — C ABI of librelgat_b200.so: the B200-native RelGAT message-passing hot path.
 *
 * The reference (radlab-dev-group/relgat-projector v0.2.1) has no FFI or plugin interface: its
 * hot path is Python calling torch and torch_scatter.  Each entry point below therefore names
 * the reference code it replaces (file:line under the reference root); INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions (every function):
 *   - all pointers are DEVICE pointers unless stated; the caller owns every buffer (inputs,
 *     outputs, workspaces); the library keeps no device memory and no mutable global state;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - returns 0 on success, a negative RG_ERR_* code for argument errors detected on the host,
 *     or a positive cudaError_t if a launch failed; no C++ exception crosses the boundary;
 *   - node features are row-major [rows, H*F] with column = h*F + f (head-major concat,
 *     reference core/model/layer.py:321); edge-wise arrays are in CSR slot order.
 */
#ifndef RELGAT_B200_H_
#define RELGAT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RG_OK 0
#define RG_ERR_ARG (-1)
#define RG_ERR_SHAPE (-2)
#define RG_ERR_ALIGN (-3)
#define RG_ERR_WORKSPACE (-4)
#define RG_ERR_DTYPE (-5)
#define RG_ERR_DRIVER (-6)

#define RG_SCORER_DISTMULT 0
#define RG_SCORER_TRANSE 1

/* ABI version (bumped on any signature change). */
int relgat_abi_version(void);

/* ---- graph index ------------------------------------------------------------------------
 * Replaces nothing in the reference (it keeps COO int64: dataset/relgat_dataset.py:123-137);
 * defines the stable by-destination (CSR), by-source (CSC) and by-relation orderings of that
 * COO which every kernel below consumes.  src/dst/rel: int64[E].  All outputs int32.
 * N = number of destination rows, N_src = number of source rows (equal on one GPU; a
 * destination-range partition owns N local destinations but reads all N_src global sources).
 *   rowptr[N+1], csr_perm[E] (original edge id of each CSR slot), csr_src/csr_rel/csr_dst[E];
 *   colptr[N_src+1], csc_slot[E] (CSR slot of the t-th by-source edge), csc_dst/csc_rel[E];
 *   relptr[R+1], rel_slot[E] (CSR slots ordered by relation). */
long long relgat_graph_index_workspace_bytes(long long E);
int relgat_graph_index_build(const long long* src, const long long* dst, const long long* rel,
                             long long E, long long N, long long N_src, long long R,
                             int* rowptr, int* csr_perm, int* csr_src, int* csr_rel, int* csr_dst,
                             int* colptr, int* csc_slot, int* csc_dst, int* csc_rel,
                             int* relptr, int* rel_slot,
                             void* workspace, long long workspace_bytes, void* stream);

/* Work tables of the streaming edge kernels (the chunks / parts / long_node / long_part_ptr arguments of
 * relgat_layer_fwd and relgat_layer_bwd_src) from a CSR pointer array ptr[n+1], built on the device without a host
 * read (the per-step receptive-field blocks build one per layer and step).  A chunk is a run of whole segments
 * covering ~chunk_edges edges and <= chunk_nodes (<= 64) segments; a segment with more than long_segment edges stands
 * alone and is cut into parts of part_edges edges, one chunk per part.  Order: parts first (segment, then part order),
 * then the ordinary chunks in segment order.
 *   chunks int32[max_chunks][4] = (first segment, segments, part slot or -1, 0); parts int32[max_parts][2] = (first edge,
 *   end edge); long_node[max_long]; long_part_ptr[max_long + 1]; counts int32[4] = (n_chunks, n_parts, n_long, edges
 *   covered) on the device — n_chunks = -1 if a table was too small (nothing else is then valid).  Safe sizes for E edges:
 *   max_long = E / (long_segment + 1) + 1, max_parts = E / part_edges + max_long, max_chunks = n + max_parts. */
long long relgat_stream_chunks_workspace_bytes(int n);
int relgat_stream_chunks_build(const int* ptr, int n, int chunk_edges, int chunk_nodes, int long_segment,
                               int part_edges, int* chunks, int max_chunks, int* parts, int max_parts,
                               int* long_node, int* long_part_ptr, int max_long, int* counts,
                               void* workspace, long long workspace_bytes, void* stream);
/* The same tables restricted to a list of segments (the destinations one batch needs, ascending): one chunk per listed
 * segment, split segments as above.  rows int64[n_rows]; n_rows_dev (optional, device): the list's true length when
 * only an upper bound n_rows is known to the host (entries beyond it are ignored).  Workspace as for n = n_rows;
 * safe sizes: max_chunks = n_rows + max_parts. */
int relgat_stream_chunks_for_rows(const int* ptr, const long long* rows, int n_rows, const int* n_rows_dev,
                                  int long_segment, int part_edges, int* chunks, int max_chunks, int* parts,
                                  int max_parts, int* long_node, int* long_part_ptr, int max_long, int* counts,
                                  void* workspace, long long workspace_bytes, void* stream);

/* ---- dense feature transform (tcgen05 + TMA) ---------------------------------------------
 * Replaces `lin(node_emb)` of core/model/layer.py:220 (all heads in one GEMM) and its autograd
 * GEMMs.  D[M,N] = A·Bᵀ, bf16 operands, fp32 accumulation in tensor memory; D is written as fp32,
 * or as bf16 when d_is_bf16 (bf16 feature-storage mode; not with split-K).
 *   a_mn/b_mn = 0: operand stored [rows = M|N, K] (K contiguous); 1: stored [K, M|N].
 *   a_lo/b_lo != NULL selects the fp32-parity mode: operands are (hi, lo) bf16 planes of an
 *   fp32 matrix (relgat_split_bf16) and hi·hi + hi·lo + lo·hi is accumulated.
 *   splits_k > 1: split-K with an ordered reduction (needs relgat_gemm_workspace_bytes). */
int relgat_split_bf16(const float* x, void* hi, void* lo, long long n, void* stream);
long long relgat_gemm_workspace_bytes(int M, int N, int K, int a_mn, int b_mn, int splits_k);
int relgat_gemm_bf16(const void* a_hi, const void* a_lo, long long lda, int a_mn,
                     const void* b_hi, const void* b_lo, long long ldb, int b_mn,
                     void* d, int d_is_bf16, long long ldd, int M, int N, int K, int splits_k,
                     void* workspace, long long workspace_bytes, int sm_count, void* stream);

/* N-tile width the GEMM uses for an N-column output, and the dX GEMM of a layer (A = dP [M, K] hi/lo planes, B = W^T
 * [N, K] planes, both K-major, fp32-parity or single-plane mode) with relgat_layer_bwd_prep of the layer below FUSED
 * into its epilogue (the tile is d loss / d act(y) while it sits in tensor memory): writes G [M, N] = dX * act'(y) * m
 * instead of dX, and t / hsum [M, H] from per-(row, tile) partials (tpart / hpart: float [M * n_tiles * 2] scratch,
 * n_tiles = ceil(N / relgat_gemm_tile_n(N))).  y = the layer below's (post-dropout) pre-activation rows [M, N = H*F],
 * bias [M] its relation-bias sums, drop_* its feature-dropout mask (NULL = off), apply_elu: act = ELU else identity.
 * Needs relgat_gemm_tile_n(N) <= F (a tile inside at most two heads), else RG_ERR_SHAPE (use the unfused pair). */
int relgat_gemm_tile_n(int N);
/* (Host-side planning for the GEMMs that replace `lin(node_emb)` of reference core/model/layer.py:220 and its autograd.)
 * Work-unit shape relgat_gemm_bf16 uses for an [M, N] output on a device with sm_count SMs (b_mn as in
 * relgat_gemm_bf16): tile_m = 256 rows when the kernel runs as CTA pairs (tcgen05 cta_group::2, M > 128), else 128;
 * tile_n = columns of one unit (the N tile, chosen by the bytes a tile moves from L2 to shared memory; twice the N tile
 * when clusters of two pairs share their A rows by TMA multicast); slots = units in flight at once.  Returns the
 * modelled cost of one k-block over all tiles in SM clocks.  Host-side split-K and orientation choices use it.
 * (relgat_gemm_tile_n is the N tile of relgat_gemm_dx_prep only.) */
long long relgat_gemm_plan(int M, int N, int b_mn, int sm_count, int* tile_m, int* tile_n, int* slots);
int relgat_gemm_dx_prep(const void* a_hi, const void* a_lo, long long lda, const void* b_hi, const void* b_lo,
                        long long ldb, float* G, int M, int N, int K, const float* y, const float* bias,
                        const unsigned int* drop_bits, int drop_words, float drop_scale, int H, int F,
                        int apply_elu, float* tpart, float* hpart, float* t, float* hsum, int sm_count, void* stream);

/* ---- RelGAT layer, edge part, forward ------------------------------------------------------
 * Replaces core/model/layer.py:220 (the [src] gather) through :318: logits + LeakyReLU(0.2),
 * per-destination stable softmax (torch_scatter.scatter_max / scatter_add), weighted
 * aggregation and the relation bias.  P: projected features [N_src, H*F] (row stride ldp),
 * A: [H, R, F] (stacked attn_vec), beta: [R] or NULL.
 * Work tables (built once per graph, see relgat_projector_b200/graph.py StreamChunks):
 *   chunks int32[n_chunks][4] = (first destination, count <= 64, part slot or -1, 0): the CSR edge
 *     array cut at destination boundaries into ~32-edge chunks; one warp streams one chunk, so
 *     short segments do not drain the load pipeline;
 *   parts int32[n_parts][2] = (first edge, end edge): a destination with more than 512 in-edges is
 *     split into 256-edge parts (one chunk each) whose partial softmax states (part_ml [n_parts,H,2],
 *     part_b [n_parts], part_acc [n_parts,H*F], caller-allocated) are merged in part order;
 *   long_node int32[n_long], long_part_ptr int32[n_long+1]: the split destinations.
 * The kernels are persistent (sm_count CTAs; <= 0 means 148) and stage the attention vectors in
 * shared memory.  work_counter: >= 32 device ints of scratch (zeroed by the call) from which warps
 * claim chunks dynamically; NULL = static round-robin assignment.
 * Outputs: out [N, H*F] fp32 pre-activation (may be NULL), optional bf16 (hi, lo) planes of
 * act(out) for the next layer's GEMM (act = ELU if apply_elu, reference model.py:286-287),
 * z [E, H] raw logits and minv [N, H, 2] = (segment max, 1/denominator) saved for backward,
 * alpha [E, H] attention weights (optional, NULL to skip), bias_out [N].
 * Dropout (training; reference layer.py:296-297 attention dropout, :321-322 feature dropout): keep-bit masks,
 * element i in word i >> 5, bit i & 31, 1 = keep (relgat_bernoulli_bits, or supplied by the caller); NULL = off.
 *   drop_bits uint32[N][drop_words] (bit = column h*F + f), drop_scale = 1/(1-p): the finished row (bias included)
 *     is multiplied before the activation; `out` then holds the POST-dropout row (what backward needs);
 *   edge_bits uint32[ceil(E*H/32)] (bit = csr slot * H + head), edge_scale: alpha is multiplied after the softmax
 *     normalisation (the denominator counts every edge). */
int relgat_layer_fwd(const void* P, int p_is_bf16, long long ldp, const float* A, const float* beta,
                     const int* rowptr, const int* csr_src, const int* csr_rel,
                     const int* chunks, int n_chunks, const int* parts, int n_parts,
                     const int* long_node, const int* long_part_ptr, int n_long,
                     float* part_ml, float* part_b, float* part_acc,
                     float* out, void* act_hi, void* act_lo, int apply_elu,
                     float* alpha, float* z, float* minv, float* bias_out,
                     const unsigned int* drop_bits, int drop_words, float drop_scale,
                     const unsigned int* edge_bits, float edge_scale,
                     const int* src_row /* NULL, or: P holds only the rows the processed destinations read, source i at
                        row src_row[i] (relgat_bitmap_ranks); with a chunk table that lists only the destinations a
                        batch needs (relgat_stream_chunks_for_rows) this is the receptive-field forward: the other rows
                        of out / act / minv / bias_out and the other edges' z are NOT written */,
                     int H, int F, int R, int sm_count, int* work_counter, void* stream);

/* ---- RelGAT layer, edge part, backward (replaces the autograd replay of layer.py:220-318) ---
 * Feature storage: P and G are fp32, or both bf16 when the *_is_bf16 flags are set (F % 8 == 0).
 * bwd_prep: G = dY * act'(out) (fp32: in place allowed; or written as bf16), t[N,H] = <G, out - bias>,
 *           hsum[N,H] = sum_f G.  row_ids (int64[n_rows], may repeat; NULL = all rows): the caller knows that
 *           only these rows of dY are non-zero (the loss reads B*(2+K) rows: reference model.py:136-137) — then
 *           apply_elu must be 0 and G fp32; only those rows are read (and, when G does not alias dY, written: the
 *           caller keeps every other row of G at zero); t / hsum of all other rows are set to 0.
 *           drop_bits / drop_words / drop_scale: the forward's feature-dropout mask (NULL = off); `out` is then the
 *           post-dropout row y = out*m*s, G = dY*act'(y)*m*s and t = <dY*act'(y), y - bias*m*s>.
 * bwd_src : by-source pass over chunks of the CSC order (work tables as in fwd, over sources;
 *           part_acc [n_parts, H*F] holds the partial rows of split sources):
 *           dP [N_src, H*F] (fp32 and/or bf16 hi/lo planes) and dz [E, H]; the attention weights
 *           are recomputed from z and minv.  edge_bits / edge_scale: the forward's attention-dropout mask.
 *           want_ds != 0 (logit-table gradient, replaces the by-relation pass): the dP rows are ldo >= H*F + H*R wide
 *           and columns H*F + h*R + r receive dS[i,h,r] = sum_{e: src=i, rel=r} dz[e,h]; dW_ext = dP_ext^T X then holds
 *           dW in its first H*F rows and dS^T X below, and dA[h] = (dS_h^T X) W_h^T — no third gather of P.  dz may then
 *           be NULL.  ldo = row stride (elements) of dP / dP_hi / dP_lo / part_acc (<= 0: H*F).
 * bwd_beta: dbeta [R] alone (the P-free part of bwd_rel), partB [n_chunks] scratch.
 * bwd_rel : by-relation pass over chunks [chunk_lo, chunk_hi) of rel_slot (a chunk never spans
 *           two relations; rel_chunk_ptr[R+1] gives each relation's chunk range):
 *           dA [H, R, F] and dbeta [R] (NULL to skip) with an ordered reduction of the partials
 *           partA [n_chunks, H*F], partB [n_chunks]. */
int relgat_layer_bwd_prep(const float* dY, const float* out, const float* bias, void* G, int g_is_bf16,
                          float* t, float* hsum, int N, int H, int F, int apply_elu,
                          const long long* row_ids, int n_rows,
                          const unsigned int* drop_bits, int drop_words, float drop_scale,
                          void* G_export_bf16 /* optional bf16 copy of G for the peers' pulls; NULL = none */,
                          int dy_compact /* with row_ids: dY holds the listed rows only (row k <-> node row_ids[k]); G is
                             a separate table whose other rows the caller keeps at zero; apply_elu allowed */,
                          void* stream);
int relgat_layer_bwd_src(const void* P, long long ldp, const void* G, int feat_is_bf16, const float* A,
                         const float* z, const float* minv, const float* t,
                         const int* colptr, const int* csc_slot, const int* csc_dst, const int* csc_rel,
                         const int* chunks, int n_chunks, const int* parts, int n_parts,
                         const int* long_node, const int* long_part_ptr, int n_long, float* part_acc,
                         float* dP, void* dP_hi, void* dP_lo, float* dz,
                         const unsigned int* edge_bits, float edge_scale,
                         const unsigned int* dst_nz_bits /* want_ds only: bit j = row j of G may be non-zero (edges into
                            other rows are skipped: their G[dst], t[dst] and hence dz are exact zeros); NULL = all rows */,
                         const int* src_row /* want_ds only: output row of each source, -1 = skip the source (it has no
                            edge into a non-zero row); from relgat_bitmap_ranks; NULL = row i for source i */,
                         int p_compact /* with src_row: P itself holds the kept sources' rows only (row src_row[i]) */,
                         int want_ds, long long ldo, int H, int F, int R, int sm_count, int* work_counter, void* stream);
/* Training-path variant of relgat_layer_bwd_src (second generation): fp32 P / G rows with F % 4 == 0, bf16 planes out,
 * want_ds semantics (rows ldo >= H*F + H*R wide, dS behind dP, no dz).  A pre-pass turns the per-edge gathers of z,
 * t and the softmax statistics (and the exp) into three coefficients per edge and head (coef: float [E*H*4] scratch,
 * by-source order), which the main loop streams; see csrc/edge_bwd_src2.cu.  Returns RG_ERR_SHAPE for layouts it does
 * not cover (use relgat_layer_bwd_src). */
int relgat_layer_bwd_src2(const float* P, long long ldp, const float* G, const float* A, const float* z,
                          const float* minv, const float* t, const int* colptr, const int* csc_slot,
                          const int* csc_dst, const int* csc_rel, const int* chunks, int n_chunks,
                          const int* parts, int n_parts, const int* long_node, const int* long_part_ptr, int n_long,
                          float* part_acc, void* dP_hi, void* dP_lo, float* coef, long long E,
                          const unsigned int* edge_bits, float edge_scale, long long ldo, int H, int F, int R,
                          int sm_count, int* work_counter, void* stream);
/* Third-generation by-source pass (csrc/edge_bwd_src3.cu; same role as relgat_layer_bwd_src: the autograd replay of
 * reference core/model/layer.py:238-318 seen from the source rows): the gathered rows reach shared memory as bulk async copies
 * (cp.async.bulk + mbarrier), several rows deep per warp, and the per-edge term dz * A[rel] is NOT added: the rows are
 * [dPa | dS] with dPa[i] = sum_e alpha_e G[dst_e]; the caller folds dP = dPa + dS·A into the GEMMs that consume them
 * (dW = dPa^T X + A^T (dS^T X); dX = [dPa | dS] · [W ; A·W]), which keeps the attention vectors out of shared memory.
 * fp32 P / G rows, F % 4 == 0; arguments as relgat_layer_bwd_src with want_ds = 1 (no A, no dz).  RG_ERR_SHAPE:
 * layout not covered. */
int relgat_layer_bwd_src3(const float* P, long long ldp, const float* G, const float* z, const float* minv,
                          const float* t, const int* colptr, const int* csc_slot, const int* csc_dst,
                          const int* csc_rel, const int* chunks, int n_chunks, const int* parts, int n_parts,
                          const int* long_node, const int* long_part_ptr, int n_long, float* part_acc,
                          float* dP, void* dP_hi, void* dP_lo, const unsigned int* edge_bits, float edge_scale,
                          const unsigned int* dst_nz_bits, const int* src_row, int p_compact, long long ldo,
                          int H, int F, int R, int sm_count, int* work_counter, void* stream);
int relgat_layer_bwd_beta(const float* hsum, const int* rel_slot, const int* csr_dst, const int* chunk_lo,
                          const int* chunk_hi, const int* rel_chunk_ptr, int n_chunks, float* partB, float* dbeta,
                          int H, int R, void* stream);
int relgat_layer_bwd_rel(const void* P, int p_is_bf16, long long ldp, const float* dz, const float* hsum,
                         const int* rel_slot, const int* csr_src, const int* csr_dst,
                         const int* chunk_lo, const int* chunk_hi, const int* rel_chunk_ptr,
                         int n_chunks, float* partA, float* partB, float* dA, float* dbeta,
                         int H, int F, int R, void* stream);

/* ---- scorers (replaces core/scorer.py:58-94 DistMult, :154-201 TransE, and the row gathers
 * x[src_ids] / x[dst_ids] of core/model/model.py:136-137) -------------------------------------
 * xs/xd: [*, D]; src_idx/dst_idx: int64[B] or NULL (row b); rel_ids: int64[B]; rel_emb [R, D].
 * fwd: score [B]; optional transform [Bt, D] for the first Bt triples (scorer.transform);
 *      optional gathered copies src_vec/dst_vec [B, D].
 * bwd: given dscore [B] (NULL = 0) and dtransform [Bt, D] (NULL = 0) writes one gradient row per
 *      triple: d_src, d_dst, d_rel [B, D] (any may be NULL).
 * relgat_index_add_sorted: ordered segmented row sum used to fold those rows per node / per
 *      relation.  sorted_keys int64[M] ascending (stable sort of the keys), perm int64[M] the
 *      matching row ids: out[key, :] (+)= sum_{p: sorted_keys[p]==key} rows[perm[p], :]. */
int relgat_score_fwd(int kind, int normalize, const float* xs, const long long* src_idx, const float* xd,
                     const long long* dst_idx, const float* rel_emb, const long long* rel_ids, int B, int D,
                     float* score, float* transform, int Bt, float* src_vec, float* dst_vec, void* stream);
int relgat_score_bwd(int kind, int normalize, const float* xs, const long long* src_idx, const float* xd,
                     const long long* dst_idx, const float* rel_emb, const long long* rel_ids, int B, int D,
                     const float* dscore, const float* dtransform, int Bt,
                     float* d_src, float* d_dst, float* d_rel, void* stream);
int relgat_index_add_sorted(const float* rows, const long long* perm, const long long* sorted_keys,
                            float* out, int M, int D, int accumulate, void* stream);

/* Margin-ranking loss fused on the flat scores (core/loss/relgat_loss.py:51-54 applied to the split of
 * trainer/relgat_projector.py:657-676, or :628-630 when bk_layout != 0): score float[B*(1+K)] ->
 * loss float[1] = mean relu(margin + neg - pos) and dscore float[B*(1+K)] = d loss / d score. */
int relgat_margin_loss(const float* score, int B, int K, float margin, int bk_layout, float* loss,
                       float* dscore, void* stream);

/* Ranking loss on pos[B] and neg[b*stride_b + k*stride_k] (element strides: both negative layouts of the reference
 * trainer — view(K,B).T of trainer/relgat_projector.py:657-676 and view(B,K) of :628-630 — are views of the flat
 * score vector).  type 0: margin ranking (core/loss/relgat_loss.py:51-54); type 1: self-adversarial
 * (relgat_loss.py:56-71: -mean logsig(pos) - mean_b sum_k softmax_k(alpha*neg).detach()*logsig(-neg)).
 * sanitize != 0 applies nan_to_num(nan=0, +-inf=+-1e9) first (trainer:584, 647-648; no gradient through a
 * non-finite score).  Outputs: loss[1], dpos[B], dneg (neg's addressing) = d loss / d score. */
#define RG_RANK_MARGIN 0
#define RG_RANK_SELF_ADVERSARIAL 1
int relgat_rank_loss(const float* pos, const float* neg, int B, int K, long long stride_b, long long stride_k,
                     int type, float margin, float alpha, int sanitize, float* loss, float* dpos, float* dneg,
                     void* stream);

/* Reconstruction terms of MultiObjectiveRelLoss (core/loss/multi_objective_loss.py:47-83 with core/loss/cosine.py:4-13
 * and core/loss/mse.py:4-10): tr = f_r(A) [B, D], dst [B, D], negative destinations nd(b,k) at
 * negdst + b*neg_stride_b + k*neg_stride_k (the reference passes [K, B, D]: strides B*D... any layout works; K <= 256).
 *   values[3] = (mean_b (1 - cos(tr_b, dst_b)), mean_{k,b} (1 - cos(tr_b, nd(b,k))), mean (tr - dst)^2)
 *   d_tr, d_dst [B, D], d_negdst (negdst's addressing; may be NULL): gradient of
 *     w_pos*values[0] + w_neg*(1 - values[1]) + w_mse*values[2]
 * partial: float[B*3] scratch.  cos uses F.normalize's eps 1e-12. */
int relgat_recon_loss(const float* tr, const float* dst, const float* negdst, int B, int K, int D,
                      long long neg_stride_b, long long neg_stride_k, float w_pos, float w_neg, float w_mse,
                      float* values, float* partial, float* d_tr, float* d_dst, float* d_negdst, void* stream);

/* Keep-bit masks for the fused dropout: bits[w] bit j = 1 with probability 1 - p_drop (16-bit resolution), from
 * Philox4x32-10(seed, counter = word index * 4 + q).  relgat_zero_rows: table[ids[i], 0:D] = 0 (clears the rows of
 * the batch gradient that relgat_index_add_sorted scattered into a persistent zero table); entries with ids[i] < 0 are
 * skipped. */
int relgat_bernoulli_bits(unsigned int* bits, long long n_words, float p_drop, unsigned long long seed, void* stream);
int relgat_zero_rows(float* table, long long ld, const long long* ids, long long n, int D, void* stream);

/* Row-set bitmaps for the backward's exact-zero hint (dst_nz_bits of relgat_layer_bwd_src).  The loss reads
 * <= B*(2+K) rows of the stack's output (model.py:136-137), so the gradient of the last layer's output is zero outside
 * them, and the gradient of layer l's output is zero outside the sources of the edges into layer l+1's non-zero rows.
 * bits: uint32[ceil(n_rows / 32)], caller-zeroed, bit j = row j may be non-zero.
 *   relgat_mark_rows:    bits |= {ids[i]}                       (ids outside [0, n_rows) are ignored)
 *   relgat_mark_sources: src_bits |= {csr_src[e] : rowptr[j] <= e < rowptr[j+1], j marked in dst_bits} */
int relgat_mark_rows(const long long* ids, long long n, long long n_rows, unsigned int* bits, void* stream);
int relgat_mark_sources(const unsigned int* dst_bits, const int* rowptr, const int* csr_src, int n_dst,
                        unsigned int* src_bits, void* stream);
/* Compact numbering of a row bitmap: rank[i] = position of row i among the marked rows (ascending) or -1, list[k] = the
 * k-th marked row, *count = their number (all on the device; rank / list may be NULL).  The compacted backward writes
 * dP rows at rank[src] (src_row of relgat_layer_bwd_src), gathers the matching input rows by list and sizes its GEMMs
 * by count. */
long long relgat_bitmap_ranks_workspace_bytes(long long n_rows);
int relgat_bitmap_ranks(const unsigned int* bits, long long n_rows, int* rank, long long* list, int* count,
                        void* workspace, long long workspace_bytes, void* stream);

/* ---- ProjectionHead hidden block (core/model/projection.py:48-67: Linear -> GELU -> LayerNorm) ------------------
 * GELU (exact, erf form) + LayerNorm (biased variance, eps inside the root) of h [M, D] in one pass per direction; the
 * linears around it run on relgat_gemm_bf16.  fwd: y [M, D], mean / rstd [M] saved.  bwd: dh [M, D], dgamma / dbeta [D]
 * (NULL to skip) through per-group partials part_g / part_b: float [relgat_gelu_layernorm_groups(M) * D] scratch each,
 * reduced in group order (reproducible).  gamma / beta may be NULL (no affine). */
int relgat_gelu_layernorm_fwd(const float* h, const float* gamma, const float* beta, float* y, float* mean,
                              float* rstd, int M, int D, float eps, void* stream);
int relgat_gelu_layernorm_groups(int M);
int relgat_gelu_layernorm_bwd(const float* dy, const float* h, const float* gamma, const float* mean,
                              const float* rstd, float* dh, float* part_g, float* part_b, float* dgamma,
                              float* dbeta, int M, int D, void* stream);

/* ---- host-side batch construction (HOST pointers; no GPU involved) --------------------------
 * Replaces the per-sample Python loop of dataset/edge.py:71-115 + trainer/components/
 * relgat_batching.py:5-19 bit for bit: `state` is CPython's MT19937 state (624 words + position, i.e.
 * random.getstate()[1]) and is advanced exactly as B*K random.choice calls (with the reference's
 * redraw-while-equal-to-tail loop) would advance it.  edges int64[n_edges][3] = (src, dst, rel);
 * idxs int64[B]; outputs int64[B*(1+K)], positives then K-major negative blocks.
 * relgat_host_shuffle: random.shuffle of a permutation (dataset/relgat_dataset.py:72). */
int relgat_host_sample_batch(unsigned int* state, const long long* edges, long long n_edges,
                             const long long* idxs, int B, int K, long long n_nodes,
                             long long* src_out, long long* rel_out, long long* dst_out);
int relgat_host_shuffle(unsigned int* state, long long* perm, long long n);

/* ---- peer tables (host side; multi-GPU, one process per GPU) ----------------------------------
 * The destination-range partition of BASELINE.json's north_star needs the transformed source rows
 * of other ranks ("exchanged ... over NVLink"; the reference is single-device, its gather is
 * core/model/layer.py:238-239 `Wh[src_ids]`).  Instead of a pack / all-gather / unpack step every rank
 * keeps its rows in a peer table: one physical allocation per rank (cuMemCreate, exported as a POSIX
 * file descriptor) and, in every process, ONE contiguous virtual range that maps all ranks' allocations
 * back to back.  relgat_layer_fwd / relgat_layer_bwd_src / relgat_layer_bwd_rel then read that range
 * through their ordinary row pointer: rows of a peer arrive by plain loads over NVLink / NVSwitch.
 *   granularity: allocation sizes must be multiples of it.
 *   create     : this rank's allocation of `bytes`; *fd is to be passed to the peers (SCM_RIGHTS).
 *   map        : peer_fds[world] (entry `rank` ignored) -> *base, a range of world*bytes in which
 *                slot s holds the table of rank (rank + s) % world (own table first), read/write.
 *   unmap      : unmaps the range and releases this rank's allocation.
 * The caller orders accesses across ranks (a stream-ordered NCCL collective between the writer's and
 * the readers' kernels is enough).  last_driver_error: CUresult of the last failed driver call. */
int relgat_peer_table_granularity(int device, unsigned long long* granularity);
int relgat_peer_table_create(int device, unsigned long long bytes, unsigned long long* handle, int* fd);
int relgat_peer_table_map(int device, int world, int rank, unsigned long long own_handle, const int* peer_fds,
                          unsigned long long bytes, void** base);
int relgat_peer_table_unmap(void* base, int world, unsigned long long bytes, unsigned long long own_handle);
int relgat_peer_table_last_driver_error(void);
/* halo pull (device): out[o(i), :] = table[ids[i], :] for i < n, o(i) = out_ids[i] (NULL: o(i) = i); ids int64 row
 * numbers of the mapped range (a peer's rows arrive over NVLink), D floats per row, ld / ldo row strides in
 * floats.  Run between the writers' rendezvous and the edge kernel that consumes [own rows | pulled rows].
 * Rows with equal out_ids must carry equal data (the batch may name a node twice).  An entry with ids[i] < 0 is
 * skipped (the caller knows that row to be all zero at its owner and keeps its local copy at zero). */
int relgat_pull_rows(const float* table, long long ld, const long long* ids, const long long* out_ids,
                     long long n, int D, float* out, long long ldo, int sm_count, void* stream);
/* same pull from a table whose rows the owner exported rounded to bf16 (half the NVLink bytes): rows are widened to
 * fp32 while they are stored (D % 8 == 0).  Stated tolerance of a step with bf16 halo rows: 2e-2 relative. */
int relgat_pull_rows_bf16(const void* table, long long ld, const long long* ids, const long long* out_ids,
                          long long n, int D, float* out, long long ldo, int sm_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RELGAT_B200_H_ */
