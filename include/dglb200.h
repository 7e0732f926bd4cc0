/*
 * dglb200.h -- C-ABI of the B200-native sparse message-passing library.
 *
 * This is the drop-in boundary for the ONE hot path of dglai/dgl-0.5-benchmark:
 * generalized SpMM / SDDMM / edge_softmax (+ fused GAT) as reached from
 *   kernel/dgl-new.py:20   dgl.ops.gspmm(g, op, reduce, nfeat, efeat)
 *   kernel/dgl-new.py:39   dgl.ops.gsddmm(g, op, ufeat, vfeat)
 *   end_to_end/full_graph/node_classification/main_dgl_citation_sage.py:75-77 (update_all copy_src sum|mean)
 *   end_to_end/full_graph/node_classification/main_dgl_arxiv_gat.py:9        (dgl.nn.pytorch.GATConv)
 *
 * The arithmetic behind those call sites lives in the un-vendored pip dependency
 * DGL v0.6.1 (docker/build.dockerfile:14, README.md:6).  Each entry point below names
 * the upstream packed function / kernel it replaces (paths are dmlc/dgl@0.6.1).
 *
 * Conventions
 *   - plain C, no C++ / torch types; every pointer is a DEVICE pointer unless the
 *     parameter name ends in `_host`.
 *   - ids are int32 (every in-scope script calls g.int(): kernel/dgl-new.py:63,
 *     main_dgl_citation_sage.py:191, main_dgl_product_sage.py:158); feature offsets are
 *     computed in 64 bit (N*D and E*D exceed 2^31 at reddit/products scale).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, does
 *     no hidden synchronisation and no allocation: the caller owns all buffers,
 *     including workspaces whose size is returned by the matching *_workspace_bytes().
 *   - return value: 0 = ok, negative = DGLB_E_* below; dglb_last_error() gives the
 *     message for the calling thread.
 *   - the library holds no per-graph state.  The only per-graph metadata is the
 *     caller-owned "hub row" description dglb_hub_t (rows whose nnz exceeds a threshold, produced by
 *     dglb_csr_find_hub_rows(), cut into segments of bounded length).
 */
#ifndef DGLB200_H_
#define DGLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGLB_ABI_VERSION 3

/* status codes */
#define DGLB_OK 0
#define DGLB_E_INVALID (-1)     /* bad argument (shape mismatch, null pointer, unknown op) */
#define DGLB_E_UNSUPPORTED (-2) /* combination not implemented on the device              */
#define DGLB_E_CUDA (-3)        /* CUDA runtime error; message in dglb_last_error()       */
#define DGLB_E_WORKSPACE (-4)   /* workspace too small                                    */

/* binary ops: upstream src/array/kernel_decl.h / cpu/spmm_binary_ops.h (Add/Sub/Mul/Div/CopyLhs/CopyRhs/Dot) */
#define DGLB_OP_ADD 0
#define DGLB_OP_SUB 1
#define DGLB_OP_MUL 2
#define DGLB_OP_DIV 3
#define DGLB_OP_COPY_LHS 4
#define DGLB_OP_COPY_RHS 5
#define DGLB_OP_DOT 6 /* gsddmm only */

/* reducers: upstream cpu/spmm_binary_ops.h (Max/Min) and "sum".  "mean" is composed by the
 * caller exactly like upstream python/dgl/ops/spmm.py (sum then divide by clamp(in_deg,1)),
 * or fused through dglb_gspmm_csr's `row_scale` epilogue. */
#define DGLB_REDUCE_SUM 0
#define DGLB_REDUCE_MAX 1
#define DGLB_REDUCE_MIN 2

/* gsddmm operand targets: upstream python/dgl/sparse.py target_mapping {u:0, e:1, v:2} */
#define DGLB_TARGET_U 0
#define DGLB_TARGET_E 1
#define DGLB_TARGET_V 2

/* element types */
#define DGLB_F32 0
#define DGLB_BF16 1 /* bf16 storage, fp32 accumulate, one rounding at the store: gspmm copy_lhs x sum
                       (+ row_scale) and gsddmm_csr u_dot_v; other entry points return DGLB_E_UNSUPPORTED */

#define DGLB_MAX_BCAST_NDIM 5

/* ---------------------------------------------------------------- library / device info */

int dglb_abi_version(void);
/* message of the last failing call on this thread ("" if none) */
const char* dglb_last_error(void);
/* fills sm_count, compute capability, L2 bytes of the current device */
int dglb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);
/* make `device` current for the calling thread (the library links its own CUDA runtime) */
int dglb_set_device(int device);

/* ---------------------------------------------------------------- sparse format build
 * replaces upstream src/array/cpu/spmat_op_impl_coo.cc::COOToCSR (order oracle) and
 * src/array/cuda/coo2csr.cu / coo_sort.cu (device path).  Result is the STABLE sort of the
 * edges by `row`: indptr[r+1]-indptr[r] = #edges with row r; within a row, entries appear in
 * increasing edge id; indices[j] = col of that edge, data[j] = its edge id.
 * To build the CSC used by SpMM pass row=dst, col=src; for the CSR (reverse graph) row=src.
 */
size_t dglb_coo_to_csr_workspace_bytes(int64_t n_rows, int64_t nnz);
int dglb_coo_to_csr(int64_t n_rows, int64_t nnz,
                    const int32_t* row, const int32_t* col,
                    int32_t* indptr /* n_rows+1 */, int32_t* indices /* nnz */, int32_t* data /* nnz */,
                    void* workspace, size_t workspace_bytes, void* stream);

/* degrees = diff(indptr): upstream UnitGraph::InDegrees / OutDegrees */
int dglb_csr_degrees(int64_t n_rows, const int32_t* indptr, int32_t* deg, void* stream);

/* 1 if data[j]==j for all j (row-sorted COO fast path of COOToCSR), written to *flag (device int32) */
int dglb_is_identity_perm(int64_t n, const int32_t* data, int32_t* flag, void* stream);

/* hub rows: rows with nnz > threshold, in no particular order (every hub row is processed
 * independently, so results do not depend on it).  *n_hub (device int32) receives the count,
 * hub_rows (capacity `cap`) the ids (writes beyond cap are dropped, the count stays exact). */
int dglb_csr_find_hub_rows(int64_t n_rows, const int32_t* indptr, int32_t threshold,
                           int32_t* hub_rows, int64_t cap, int32_t* n_hub, void* stream);
/* thresholds the library recommends for a given feature width (elements): the first for the
 * segmented hub path of gspmm / gsddmm (64-256 edges: a row-task's duration follows its edge count),
 * the second for the hub path of the fused GAT kernels (128 edges: measured optimum of a sweep on
 * power-law graphs), the third for the segmented hub path of edge_softmax
 * (64-1024 edges by head count: one warp walks a segment in ~32 trips per pass) */
int32_t dglb_default_hub_threshold(int64_t out_len);
int32_t dglb_default_row_hub_threshold(int64_t out_len);
int32_t dglb_default_softmax_hub_threshold(int64_t n_heads);

/* Hub-row metadata handed to the compute entry points (NULL = treat every row as an ordinary row).
 * Rows with nnz > threshold are skipped by the row-per-group kernels.  gspmm / gsddmm cut each hub
 * row into segments of at most seg_len entries, one CTA per segment, so a 20 000-edge hub is spread
 * over many SMs; gspmm writes per-segment partial results to `workspace` and a second kernel
 * combines the segments of a row in segment order (deterministic, no atomics; max/min ties still
 * resolve to the first CSR entry).  edge_softmax uses the same segments (a warp per segment, three
 * launches: segment stats, per-row combine, apply; workspace dglb_edge_softmax_workspace_bytes).
 * The fused GAT kernels process every segment like an ordinary row and combine the partial results
 * (workspace dglb_gat_hub_workspace_bytes); without a workspace they use one CTA per hub row.
 *   rows     [n_hub]    hub row ids
 *   seg_ptr  [n_hub+1]  first segment of each hub row (prefix sum of ceil(nnz/seg_len))
 *   seg_hub  [n_seg]    index into rows[] of the hub row a segment belongs to
 *   workspace           device scratch of >= dglb_hub_workspace_bytes(n_seg, out_len, with_args)
 *   light_indptr [n_rows+1]  (optional, may be NULL) prefix sum of the nnz of the NON-hub rows (hub rows count 0): the
 *                       persistent ring kernels (wide-row gspmm copy_lhs/sum, gsddmm u_dot_v) cut the rows into equal
 *                       shares of THEIR work with it; without it they balance on indptr, which on a hub-heavy graph
 *                       leaves the warps whose range is mostly hub edges idle (2x on a power-law reddit shape)
 *   row_order [n_rows]  (optional, may be NULL) the row ids sorted by non-increasing nnz (stable).  The row-per-group
 *                       kernels of gspmm / gsddmm then hand the k-th group of lanes row_order[k] instead of row k: the
 *                       4-32 rows a warp walks together have (nearly) equal lengths, so no group idles until the
 *                       warp's longest row is done, and the longest rows start first (degree-binned nnz balance; every
 *                       row is still summed by one group in CSR order: results are bit-identical with or without it)
 * All arrays are device pointers owned by the caller. */
typedef struct dglb_hub_t {
  const int32_t* rows;
  const int32_t* seg_ptr;
  const int32_t* seg_hub;
  int32_t n_hub, n_seg, seg_len, threshold;
  void* workspace;
  size_t workspace_bytes;
  const int32_t* light_indptr;
  const int32_t* row_order;
} dglb_hub_t;
size_t dglb_hub_workspace_bytes(int64_t n_seg, int64_t out_len, int with_args);

/* ---------------------------------------------------------------- staged edge order
 * (new; upstream addresses per-edge operands as efeat[data[j]] / out[data[j]] directly: cuda/spmm.cuh, sddmm.cuh.)
 * When the edges of a graph were not created in CSC (CSR) order, `eids` is a random permutation and every narrow
 * per-edge access W[eids[j]] costs a 128-byte DRAM line.  dglb_edge_stage_plan derives from `eids` (the `data` array
 * of dglb_coo_to_csr) a second permutation that is the identity up to a shuffle INSIDE buckets of 2^log2_bucket
 * consecutive CSR positions:
 *   stage_pos[e]  staged slot of edge id e            (bucket of slot == bucket of the edge's CSR position)
 *   slot[j]       staged slot of CSR position j       (= stage_pos[eids[j]]; inside a bucket, slots are in edge-id order)
 * dglb_edge_stage moves a per-edge tensor of `row_bytes`-byte rows between edge-id order and staged order in one pass
 * (to_staged = 1: dst[stage_pos[e]] = src[e]; 0: dst[e] = src[stage_pos[e]]) at ~3 x row_bytes of DRAM traffic per edge.
 * The compute entry points are then called with `slot` as their `eids` argument and the staged tensor as the per-edge
 * operand / output: inside a kernel every per-edge access stays within a 2^log2_bucket-slot window of its CSR position.
 * Entry points that also use `eids` as a NAME (arg_e of max/min, the dropout counter of the fused GAT kernels) must be
 * given the real edge ids.  Recommended log2_bucket: 15.
 */
size_t dglb_edge_stage_plan_workspace_bytes(int64_t nnz, int log2_bucket);
int dglb_edge_stage_plan(int64_t nnz, const int32_t* eids, int log2_bucket,
                         int32_t* stage_pos /* nnz, by edge id */, int32_t* slot /* nnz, by CSR position */,
                         void* workspace, size_t workspace_bytes, void* stream);
int dglb_edge_stage(int to_staged, int64_t nnz, int64_t row_bytes, const int32_t* stage_pos,
                    const void* src, void* dst, void* stream);

/* ---------------------------------------------------------------- generalized SpMM
 * replaces upstream FFI `_CAPI_DGLKernelSpMM` (src/array/kernel.cc::SpMM ->
 * cuda/spmm.cu::SpMMCsr / CusparseCsrmm2 / cuda/spmm.cuh::SpMMCsrKernel).
 *
 *   out[r, k] = REDUCE_{j in [indptr[r], indptr[r+1])}  op( ufeat[indices[j], lk], efeat[eid(j), rk] )
 *   eid(j) = eids ? eids[j] : j ;  (lk, rk) follow numpy broadcasting of the trailing
 *   shapes lhs_shape / rhs_shape (right-aligned, `ndim` entries each, host arrays; for
 *   copy_lhs / copy_rhs pass the used operand's shape for both).
 *
 *  - sum: `out` is fully written (empty rows -> 0), callers need not zero it.
 *  - max/min: strict compare in row order => the FIRST entry in CSR order wins ties;
 *    arg_u[r,k] = indices[j*], arg_e[r,k] = eid(j*) (either may be NULL); empty rows ->
 *    out = -/+inf, args = 0 -- or out = 0 with DGLB_SPMM_ZERO_INF, which folds upstream's Python
 *    post-pass (python/dgl/ops/spmm.py: where(isinf(out), 0, out)) into the store.  arg_u is
 *    recorded only when the op reads the node operand, arg_e only when it reads the edge operand
 *    (what upstream allocates); the other one, if passed, is filled with 0.
 *  - row_scale (may be NULL): fused epilogue out[r,:] = out[r,:] / row_scale[r] (IEEE division),
 *    used for reducer "mean" with row_scale = float(clamp(in_deg,1)).
 *  - flags: bit mask.  DGLB_SPMM_ACCUMULATE (reducer sum only): out[r,:] += result instead of
 *    out[r,:] = result; lets a row-partitioned caller aggregate one source shard at a time while the
 *    next shard is in flight.  DGLB_SPMM_ZERO_INF (max/min only): see above.
 *  - hub (may be NULL): rows listed there (nnz > hub->threshold) are processed by the segmented
 *    split-row path instead of the row-per-group path; the list must come from
 *    dglb_csr_find_hub_rows(indptr, hub->threshold).
 */
#define DGLB_SPMM_ACCUMULATE 1
#define DGLB_SPMM_ZERO_INF 2
int dglb_gspmm_csr(int op, int reduce, int dtype,
                   int64_t n_rows, int64_t n_cols, int64_t nnz,
                   const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                   const void* ufeat, const void* efeat,
                   int ndim, const int64_t* lhs_shape_host, const int64_t* rhs_shape_host,
                   void* out, int32_t* arg_u, int32_t* arg_e,
                   const float* row_scale, int flags,
                   const dglb_hub_t* hub,
                   void* stream);

/* ---------------------------------------------------------------- generalized SDDMM
 * replaces upstream FFI `_CAPI_DGLKernelSDDMM` (src/array/kernel.cc::SDDMM ->
 * cuda/sddmm.cuh::SDDMMCooKernel / SDDMMCsrKernel).
 *
 *   out[e, k] = op( lhs[sel(lhs_target, e), lk], rhs[sel(rhs_target, e), rk] )      (edge-id order)
 *   sel(U,e)=src[e], sel(E,e)=e, sel(V,e)=dst[e];  op DOT reduces the LAST dim (which must match
 *   in both shapes) and the output's last dim becomes 1.
 *
 * Two traversals:
 *   _csr : destination-major over the CSC (indptr over dst, indices = src, eids = edge id).
 *          (U,V) targets with equal shapes take the vectorised path in which the dst row is read
 *          once per row; everything else a generic kernel (row found by binary search).
 *   _coo : edge-parallel over (src[e], dst[e]); any targets / broadcast.
 */
int dglb_gsddmm_csr(int op, int dtype, int lhs_target, int rhs_target,
                    int64_t n_dst, int64_t n_src, int64_t nnz,
                    const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                    const void* lhs, const void* rhs,
                    int ndim, const int64_t* lhs_shape_host, const int64_t* rhs_shape_host,
                    void* out,
                    const dglb_hub_t* hub,
                    void* stream);

int dglb_gsddmm_coo(int op, int dtype, int lhs_target, int rhs_target,
                    int64_t n_src, int64_t n_dst, int64_t nnz,
                    const int32_t* src, const int32_t* dst,
                    const void* lhs, const void* rhs,
                    int ndim, const int64_t* lhs_shape_host, const int64_t* rhs_shape_host,
                    void* out, void* stream);

/* ---------------------------------------------------------------- edge_softmax (norm_by='dst')
 * replaces the 4+1 (fwd) / 2+2 (bwd) launch composite of upstream
 * python/dgl/backend/pytorch/sparse.py::EdgeSoftmax with one kernel each (plus three small ones
 * when hub rows are listed).
 *   fwd: out[e,h] = exp(logits[e,h] - max_v) / sum_v  over the in-edges of v = dst[e]
 *   bwd: grad_logits[e,h] = out*g - out * sum_{in(v)} (out*g)
 * logits/out/grad are (E, H) in edge-id order; indptr is the CSC over dst, eids its data.
 * hub (may be NULL): rows with more than hub->threshold in-edges are processed per SEGMENT; needs the
 * segment lists and hub->workspace >= dglb_edge_softmax_workspace_bytes(n_seg, n_hub, n_heads).
 */
size_t dglb_edge_softmax_workspace_bytes(int64_t n_seg, int64_t n_hub, int64_t n_heads);
int dglb_edge_softmax_fwd(int dtype, int64_t n_dst, int64_t nnz, int64_t n_heads,
                          const int32_t* indptr, const int32_t* eids,
                          const void* logits, void* out,
                          const dglb_hub_t* hub,
                          void* stream);
int dglb_edge_softmax_bwd(int dtype, int64_t n_dst, int64_t nnz, int64_t n_heads,
                          const int32_t* indptr, const int32_t* eids,
                          const void* out, const void* grad_out, void* grad_logits,
                          const dglb_hub_t* hub,
                          void* stream);

/* ---------------------------------------------------------------- fused GAT attention
 * replaces, for dgl.nn.pytorch.GATConv.forward (upstream python/dgl/nn/pytorch/conv/gatconv.py;
 * written-out twin: main_pyg_arxiv_gat.py:98-111), the chain
 *   apply_edges(u_add_v) -> leaky_relu -> edge_softmax -> attn_drop -> update_all(u_mul_e, sum)
 * with ONE forward kernel that never writes a per-edge tensor:
 *   e_j   = leaky_relu(el[src_j,h] + er[v,h], slope);  a_j = softmax_j(e)
 *   rst[v,h,:] = sum_j a_j * drop_j * ft[src_j,h,:]
 * and saves per-destination row_max[v,h], row_sum[v,h] for the backward.  drop_j is the
 * attention-dropout factor (0 or 1/(1-p)) derived from a counter-based hash of
 * (seed, edge id, head), so the backward regenerates it; dropout_p = 0 disables it.
 * Backward is two kernels that recompute the scores (dd_j = drop_j * ft[src_j,h,:].grad_rst[v,h,:]):
 *   _bwd_dst (CSC over dst): s1[v,h] = sum_j a_j*dd_j ; grad_er[v,h] = sum_j a_j*(dd_j - s1)*lrelu'_j ;
 *                            also writes row_pack[v,h] = {er, row_max, row_sum, s1} (one 16-byte record
 *                            per (node, head) so the src pass fetches it with a single load per edge)
 *   _bwd_src (CSR over src): grad_ft[u,h,:] = sum_{u->v} a*drop*grad_rst[v,h,:] ;
 *                            grad_el[u,h]   = sum_{u->v} a*(dd - s1[v,h])*lrelu'
 * optional `edge_scores` (E,H) in edge-id order receives a_j (before dropout); pass NULL on the
 * training path.  n_heads <= 8.  hub lists refer to the matrix each kernel traverses.
 * Hub rows: with the segment lists and hub->workspace >= dglb_gat_hub_workspace_bytes(n_seg, n_heads,
 * head_dim) (16-byte aligned) every segment of a hub row is processed like an ordinary row and small
 * combine kernels fold the partial results in segment order (all three passes are linear in the edges
 * once the row's max / sum are known); without a workspace one CTA handles each hub row.
 */
size_t dglb_gat_hub_workspace_bytes(int64_t n_seg, int64_t n_heads, int64_t head_dim);
int dglb_gat_fused_fwd(int dtype, int64_t n_dst, int64_t n_src, int64_t nnz,
                       int64_t n_heads, int64_t head_dim, float negative_slope,
                       float dropout_p, uint64_t seed,
                       const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                       const void* ft, const void* el, const void* er,
                       void* rst, float* row_max, float* row_sum,
                       void* edge_scores,
                       const dglb_hub_t* hub,
                       void* stream);

int dglb_gat_fused_bwd_dst(int dtype, int64_t n_dst, int64_t n_src, int64_t nnz,
                           int64_t n_heads, int64_t head_dim, float negative_slope,
                           float dropout_p, uint64_t seed,
                           const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                           const void* ft, const void* el, const void* er,
                           const float* row_max, const float* row_sum,
                           const void* grad_rst,
                           float* row_pack /* (n_dst,H,4), 16-byte aligned */, void* grad_er /* (n_dst,H) */,
                           const dglb_hub_t* hub,
                           void* stream);

int dglb_gat_fused_bwd_src(int dtype, int64_t n_src, int64_t n_dst, int64_t nnz,
                           int64_t n_heads, int64_t head_dim, float negative_slope,
                           float dropout_p, uint64_t seed,
                           const int32_t* indptr_csr, const int32_t* indices_csr /* dst ids */,
                           const int32_t* eids_csr,
                           const void* ft, const void* el,
                           const float* row_pack /* from _bwd_dst */,
                           const void* grad_rst,
                           void* grad_ft /* (n_src,H,F) */, void* grad_el /* (n_src,H) */,
                           const dglb_hub_t* hub,
                           void* stream);

/* ---------------------------------------------------------------- batched small graphs (BASELINE config 5)
 * Fused GCN message + sum of the graph-classification scripts (main_dgl_molhiv_gcn.py:46,50-52 -- upstream runs the
 * Python UDF `message` with torch ops on (E, D) tensors, then update_all(copy_e, sum)):
 *   fwd (CSC over dst): out[v,:]    = sum_{e=(u->v), CSC order} (c_src[u] * c_dst[v]) * relu(x[u,:] + w[eid(e),:])
 *   bwd (CSR over src): grad_w[e,:] = (x[u,:] + w[e,:] > 0) ? grad_out[v,:] * (c_src[u] * c_dst[v]) : 0     (every edge once)
 *                       grad_x[u,:] = sum_{e=(u->v), CSR order} grad_w[e,:]
 * x (n_src, D), w (nnz, D) in edge-id order, c_src (n_src), c_dst (n_dst), fp32.  Products and sums are rounded one by one
 * (no FMA contraction), so the forward equals the unfused composite bit for bit.  grad_w rows of edges that appear in no
 * CSR row (padding slots of a fixed-size batch) are not written: the caller zero-fills grad_w when such slots exist.
 * Rows are walked sequentially by one thread per 4 columns: meant for the short rows of batched small graphs.
 */
int dglb_gcn_msg_sum_fwd(int64_t n_dst, int64_t n_src, int64_t nnz, int64_t feat_len,
                         const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                         const float* x, const float* w, const float* c_src, const float* c_dst,
                         float* out, void* stream);
int dglb_gcn_msg_sum_bwd(int64_t n_src, int64_t n_dst, int64_t nnz, int64_t feat_len,
                         const int32_t* indptr_csr, const int32_t* indices_csr, const int32_t* eids_csr,
                         const float* x, const float* w, const float* c_src, const float* c_dst,
                         const float* grad_out, float* grad_x, float* grad_w, void* stream);

/* Sum of categorical embeddings (ogb.graphproppred.mol_encoder AtomEncoder / BondEncoder as used by
 * main_dgl_molhiv_gcn.py:28,72): x (n_rows, n_columns) int64 codes, `table` (offsets_host[n_columns], feat_len) the
 * columns' embedding tables stacked, offsets_host[k] (HOST array, n_columns + 1 entries) the first row of column k's table.
 *   fwd: out[i,:]        = table[off[0] + x[i,0],:] + table[off[1] + x[i,1],:] + ...      (in column order: bit-identical to
 *                          the sequential `out = out + emb_k(x[:, k])` loop of the encoder)
 *   bwd: grad_table[r,:] = sum over rows i with off[k] + x[i,k] == r of grad_out[i,:]   (deterministic: rows are summed in
 *                          order inside 512-row chunks, chunks in order; no atomics, no sort.  workspace >=
 *                          dglb_cat_embed_sum_bwd_workspace_bytes(n_rows, offsets_host[n_columns], feat_len); feat_len <= 1024)
 */
#define DGLB_MAX_CAT_COLUMNS 16
int dglb_cat_embed_sum_fwd(int64_t n_rows, int64_t n_columns, int64_t feat_len, const int64_t* x,
                           const int32_t* offsets_host, const float* table, float* out, void* stream);
size_t dglb_cat_embed_sum_bwd_workspace_bytes(int64_t n_rows, int64_t n_table_rows, int64_t feat_len);
int dglb_cat_embed_sum_bwd(int64_t n_rows, int64_t n_columns, int64_t feat_len, const int64_t* x,
                           const int32_t* offsets_host, const float* grad_out, float* grad_table,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Device-side dgl.batch (replaces upstream python/dgl/batch.py::batch and the per-batch COO -> CSC / CSR conversions of
 * the training loop, main_dgl_molhiv_gcn.py:101,163).  The dataset is ONE union graph on the device (member graph g owns
 * nodes [node_ptr[g], node_ptr[g+1]) and edges [edge_ptr[g], edge_ptr[g+1]); its CSC / CSR come from dglb_coo_to_csr).
 * Node ranges of member graphs are disjoint and increasing, so the stable CSC / CSR of a batch is the concatenation of
 * the members' slices with shifted ids: bit-identical to dglb_coo_to_csr on the batched COO, without a sort.
 *   dglb_batch_offsets: out_node_ptr / out_edge_ptr [n_sel + 2] = exclusive prefix sums of the selected graphs' node /
 *     edge counts, then the totals, then the padded sizes (so rows [0, n_sel] of out_node_ptr are the CSC indptr of the
 *     "node -> member graph" relation, the last row being the padding graph); status[0] (may be NULL) = 1 when the
 *     batch exceeds the padded sizes.  One CTA.
 *   dglb_batch_gather: fills the fixed-size buffers of dglb_batch_io_t (any output may be NULL).  Nodes / edge slots
 *     beyond the batch's real counts are padding: isolated nodes of the padding member graph; edge slots that no
 *     indptr range covers (their COO entries point at node n_nodes_pad - 1).
 * Both are asynchronous and allocation-free: a whole training step including batch construction can be captured in one
 * CUDA graph and replayed with only `graph_ids` changing.
 */
typedef struct dglb_batch_io_t {
  int32_t n_sel, n_nodes_pad, n_edges_pad;
  /* selection and the union graph (inputs) */
  const int32_t* graph_ids;      /* [n_sel]      member graphs of the batch, in batch order            */
  const int32_t* node_ptr;       /* [G + 1]                                                            */
  const int32_t* edge_ptr;       /* [G + 1]                                                            */
  const int32_t* out_node_ptr;   /* [n_sel + 2]  from dglb_batch_offsets                               */
  const int32_t* out_edge_ptr;   /* [n_sel + 2]                                                        */
  const int32_t* u_src;          /* [E_union]    COO of the union graph, edge-id order                 */
  const int32_t* u_dst;
  const int32_t* u_csc_indptr;   /* [N_union+1]  CSC (by dst) of the union graph                       */
  const int32_t* u_csc_indices;
  const int32_t* u_csc_eids;     /* NULL = identity                                                    */
  const int32_t* u_csr_indptr;   /* CSR (by src)                                                       */
  const int32_t* u_csr_indices;
  const int32_t* u_csr_eids;
  /* the batch (outputs, fixed size) */
  int32_t* src;                  /* [n_edges_pad]                                                      */
  int32_t* dst;
  int32_t* csc_indptr;           /* [n_nodes_pad + 1]                                                  */
  int32_t* csc_indices;          /* [n_edges_pad]                                                      */
  int32_t* csc_eids;
  int32_t* csr_indptr;
  int32_t* csr_indices;
  int32_t* csr_eids;
  int32_t* node_graph;           /* [n_nodes_pad] member-graph slot of every node (padding: n_sel)     */
  int32_t* node_map;             /* [n_nodes_pad] union-graph node of every batch node (feature gather) */
  int32_t* edge_map;             /* [n_edges_pad] union-graph edge id of every batch edge              */
} dglb_batch_io_t;
int dglb_batch_offsets(int64_t n_sel, const int32_t* graph_ids, const int32_t* node_ptr, const int32_t* edge_ptr,
                       int32_t* out_node_ptr, int32_t* out_edge_ptr,
                       int64_t n_nodes_pad, int64_t n_edges_pad, int32_t* status, void* stream);
int dglb_batch_gather(const dglb_batch_io_t* io, void* stream);

/* ---------------------------------------------------------------- halo exchange of the 1-D row partition
 * (new: the reference is single-GPU; BASELINE.json north_star asks for "halo and full feature all-gather".)
 *   dst[idx[i], :] = src[idx[i], :]   for i in [0, n_idx), rows of row_bytes bytes (a multiple of 4)
 * src is typically a peer GPU's buffer mapped over NVLink (symmetric memory), dst the local gather buffer's slot for that
 * peer, idx the sorted unique rows of the peer that the rank's CSC / CSR slice references.  Rows keep their position, so
 * nothing downstream changes and results stay bit-identical to a full exchange.
 */
int dglb_copy_rows_indexed(int64_t n_idx, const int32_t* idx, int64_t row_bytes, const void* src, void* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGLB200_H_ */
