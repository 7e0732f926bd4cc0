// capi.cu -- the extern "C" boundary declared in include/dglb200.h: argument checks (the role of
// upstream src/array/kernel.cc::CheckCtx/CheckShape/CheckContiguous), broadcast analysis (the role
// of include/dgl/bcast.h::CalcBcastOff) and dispatch into the sm_100a kernels.
#include <cstdarg>

#include "kernels.cuh"

namespace dglb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return DGLB_E_CUDA;
}

int make_bcast(int op, int ndim, const int64_t* lhs_shape, const int64_t* rhs_shape, BcastShape* b,
               int64_t* reduce_size) {
  if (ndim < 1 || ndim > DGLB_MAX_BCAST_NDIM || !lhs_shape || !rhs_shape) {
    set_error("broadcast: ndim must be in [1,%d] and shapes non-null", DGLB_MAX_BCAST_NDIM);
    return DGLB_E_INVALID;
  }
  b->ndim = ndim;
  b->lhs_len = b->rhs_len = b->out_len = 1;
  *reduce_size = 1;
  for (int d = 0; d < ndim; ++d) {
    const int64_t l = lhs_shape[d], r = rhs_shape[d];
    if (l < 0 || r < 0 || (l != r && l != 1 && r != 1)) {
      set_error("broadcast: cannot broadcast dim %d (%lld vs %lld)", d, (long long)l, (long long)r);
      return DGLB_E_INVALID;
    }
    b->lhs[d] = l; b->rhs[d] = r; b->out[d] = l > r ? l : r;
    b->lhs_len *= l; b->rhs_len *= r;
  }
  if (op == DGLB_OP_DOT) {
    if (lhs_shape[ndim - 1] != rhs_shape[ndim - 1]) {
      set_error("dot: last dims differ (%lld vs %lld)", (long long)lhs_shape[ndim - 1], (long long)rhs_shape[ndim - 1]);
      return DGLB_E_INVALID;
    }
    *reduce_size = lhs_shape[ndim - 1];
    b->out[ndim - 1] = 1;
  }
  for (int d = 0; d < ndim; ++d) b->out_len *= b->out[d];
  return DGLB_OK;
}

static bool valid_op(int op, bool allow_dot) {
  return op >= DGLB_OP_ADD && (op <= DGLB_OP_COPY_RHS || (allow_dot && op == DGLB_OP_DOT));
}

}  // namespace dglb

using namespace dglb;

extern "C" {

int dglb_abi_version(void) { return DGLB_ABI_VERSION; }

const char* dglb_last_error(void) { return g_err; }

int dglb_set_device(int device) {
  DGLB_CUDA(cudaSetDevice(device));
  return DGLB_OK;
}

int dglb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
  int dev = 0;
  DGLB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DGLB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (l2_bytes) *l2_bytes = prop.l2CacheSize;
  return DGLB_OK;
}

size_t dglb_coo_to_csr_workspace_bytes(int64_t n_rows, int64_t nnz) {
  return coo_to_csr_workspace_bytes(n_rows, nnz);
}

int dglb_coo_to_csr(int64_t n_rows, int64_t nnz, const int32_t* row, const int32_t* col, int32_t* indptr,
                    int32_t* indices, int32_t* data, void* workspace, size_t workspace_bytes, void* stream) {
  DGLB_CHECK_ARG(n_rows >= 0 && nnz >= 0, "coo_to_csr: negative size");
  DGLB_CHECK_ARG(indptr != nullptr, "coo_to_csr: indptr is null");
  DGLB_CHECK_ARG(nnz == 0 || (row && col && indices && data), "coo_to_csr: null array with nnz > 0");
  return coo_to_csr(n_rows, nnz, row, col, indptr, indices, data, workspace, workspace_bytes,
                    static_cast<cudaStream_t>(stream));
}

int dglb_csr_degrees(int64_t n_rows, const int32_t* indptr, int32_t* deg, void* stream) {
  DGLB_CHECK_ARG(n_rows >= 0 && indptr && (deg || n_rows == 0), "csr_degrees: bad arguments");
  return csr_degrees(n_rows, indptr, deg, static_cast<cudaStream_t>(stream));
}

int dglb_is_identity_perm(int64_t n, const int32_t* data, int32_t* flag, void* stream) {
  DGLB_CHECK_ARG(n >= 0 && flag && (data || n == 0), "is_identity_perm: bad arguments");
  return is_identity_perm(n, data, flag, static_cast<cudaStream_t>(stream));
}

int dglb_csr_find_hub_rows(int64_t n_rows, const int32_t* indptr, int32_t threshold, int32_t* hub_rows,
                           int64_t cap, int32_t* n_hub, void* stream) {
  DGLB_CHECK_ARG(n_rows >= 0 && indptr && n_hub && (hub_rows || cap == 0), "find_hub_rows: bad arguments");
  return csr_find_hub_rows(n_rows, indptr, threshold, hub_rows, cap, n_hub, static_cast<cudaStream_t>(stream));
}

size_t dglb_edge_stage_plan_workspace_bytes(int64_t nnz, int log2_bucket) {
  if (log2_bucket < 5 || log2_bucket > 24) return 0;
  return edge_stage_plan_workspace_bytes(nnz, log2_bucket);
}

int dglb_edge_stage_plan(int64_t nnz, const int32_t* eids, int log2_bucket, int32_t* stage_pos, int32_t* slot,
                         void* workspace, size_t workspace_bytes, void* stream) {
  DGLB_CHECK_ARG(nnz >= 0 && log2_bucket >= 5 && log2_bucket <= 24, "edge_stage_plan: bad nnz / log2_bucket");
  DGLB_CHECK_ARG(nnz == 0 || (eids && stage_pos && slot), "edge_stage_plan: null array");
  return edge_stage_plan(nnz, eids, log2_bucket, stage_pos, slot, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int dglb_edge_stage(int to_staged, int64_t nnz, int64_t row_bytes, const int32_t* stage_pos, const void* src, void* dst,
                    void* stream) {
  DGLB_CHECK_ARG(nnz >= 0 && row_bytes >= 0 && row_bytes % 4 == 0, "edge_stage: rows must be a whole number of 4-byte words");
  DGLB_CHECK_ARG(nnz == 0 || row_bytes == 0 || (stage_pos && src && dst && src != dst), "edge_stage: null or aliased buffers");
  return edge_stage_move(to_staged ? 1 : 0, nnz, row_bytes / 4, stage_pos, src, dst, static_cast<cudaStream_t>(stream));
}

size_t dglb_hub_workspace_bytes(int64_t n_seg, int64_t out_len, int with_args) {
  if (n_seg <= 0 || out_len <= 0) return 0;
  return (size_t)n_seg * (size_t)out_len * 4 * (with_args ? 3 : 1);
}

int32_t dglb_default_row_hub_threshold(int64_t out_len) {
  // One CTA per hub row (fused GAT).  Swept on power-law reddit / products graphs for (H,F) = (1,16), (4,16),
  // (4,40), (1,64): monotonically faster down to the smallest cut-off tried (profiles/r01_notes.md section 11:
  // e.g. (1,16) forward 3.61 ms at the old byte-based cut-off of 8192 edges, 0.62 ms at 128) -- a long row on a
  // 4..32-lane group is a serial chain of gather batches, a CTA splits it 8..64 ways.
  (void)out_len;
  return 128;
}

int32_t dglb_default_softmax_hub_threshold(int64_t n_heads) {
  // a hub segment is walked by one warp whose lanes are (edge slot, head): ~32 trips per pass
  int64_t hp = 1;
  while (hp < n_heads && hp < 32) hp <<= 1;
  int64_t t = 1024 / hp;
  if (t < 64) t = 64;
  return (int32_t)t;
}

size_t dglb_edge_softmax_workspace_bytes(int64_t n_seg, int64_t n_hub, int64_t n_heads) {
  return edge_softmax_workspace_bytes(n_seg, n_hub, n_heads);
}

int32_t dglb_default_hub_threshold(int64_t out_len) {
  // A row-task is a chain of dependent batches (G edges per round trip), so its duration is set by its
  // EDGE count, not its bytes: measured on a power-law reddit graph the best cut-off is ~160-400 edges
  // from D=64 to D=602 (profiles/r01_notes.md).  Narrow rows (few lanes per row) get a lower cut-off.
  if (out_len < 1) out_len = 1;
  const int64_t ncols = (out_len + 3) / 4;
  int64_t g = 1;
  while (g < 32 && g < ncols) g <<= 1;
  int64_t t = 16 * g;
  if (t < 64) t = 64;
  if (t > 256) t = 256;
  return (int32_t)t;
}

int dglb_gspmm_csr(int op, int reduce, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz,
                   const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* ufeat,
                   const void* efeat, int ndim, const int64_t* lhs_shape_host, const int64_t* rhs_shape_host,
                   void* out, int32_t* arg_u, int32_t* arg_e, const float* row_scale, int flags,
                   const dglb_hub_t* hub, void* stream) {
  DGLB_CHECK_ARG(valid_op(op, false), "gspmm: unknown op %d", op);
  DGLB_CHECK_ARG(reduce >= DGLB_REDUCE_SUM && reduce <= DGLB_REDUCE_MIN, "gspmm: unknown reducer %d", reduce);
  if (dtype != DGLB_F32 && dtype != DGLB_BF16) { set_error("gspmm: unknown dtype %d", dtype); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "gspmm: negative size");
  DGLB_CHECK_ARG(indptr && (indices || nnz == 0) && out, "gspmm: null graph/out pointer");
  DGLB_CHECK_ARG(op == DGLB_OP_COPY_RHS || ufeat || nnz == 0, "gspmm: op needs lhs (node) data");
  DGLB_CHECK_ARG(op == DGLB_OP_COPY_LHS || efeat || nnz == 0, "gspmm: op needs rhs (edge) data");
  DGLB_CHECK_ARG((flags & ~(DGLB_SPMM_ACCUMULATE | DGLB_SPMM_ZERO_INF)) == 0, "gspmm: unknown flag bits %d", flags);
  DGLB_CHECK_ARG(!(flags & DGLB_SPMM_ACCUMULATE) || reduce == DGLB_REDUCE_SUM,
                 "gspmm: accumulate is only defined for reducer sum");
  BcastShape b;
  int64_t rs;
  int rc = make_bcast(op, ndim, lhs_shape_host, rhs_shape_host, &b, &rs);
  if (rc != DGLB_OK) return rc;
  if (op == DGLB_OP_COPY_LHS) { for (int d = 0; d < b.ndim; ++d) { b.rhs[d] = b.lhs[d]; b.out[d] = b.lhs[d]; } b.rhs_len = b.out_len = b.lhs_len; }
  if (op == DGLB_OP_COPY_RHS) { for (int d = 0; d < b.ndim; ++d) { b.lhs[d] = b.rhs[d]; b.out[d] = b.rhs[d]; } b.lhs_len = b.out_len = b.rhs_len; }
  if (dtype == DGLB_BF16)
    return spmm_csr_bf16(op, reduce, n_rows, n_cols, nnz, indptr, indices, ufeat, b.out_len, out, row_scale, flags, hub,
                         static_cast<cudaStream_t>(stream));
  return spmm_csr_f32(op, reduce, n_rows, n_cols, nnz, indptr, indices, eids, static_cast<const float*>(ufeat),
                      static_cast<const float*>(efeat), b, static_cast<float*>(out), arg_u, arg_e, row_scale,
                      flags, hub, static_cast<cudaStream_t>(stream));
}

static int sddmm_common(int op, int dtype, int lhs_target, int rhs_target, int ndim, const int64_t* ls,
                        const int64_t* rs_, const void* lhs, const void* rhs, void* out, int64_t nnz,
                        BcastShape* b, int64_t* reduce_size) {
  DGLB_CHECK_ARG(valid_op(op, true), "gsddmm: unknown op %d", op);
  if (dtype != DGLB_F32 && dtype != DGLB_BF16) { set_error("gsddmm: unknown dtype %d", dtype); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(lhs_target >= 0 && lhs_target <= 2 && rhs_target >= 0 && rhs_target <= 2, "gsddmm: bad target");
  DGLB_CHECK_ARG(nnz >= 0 && (out || nnz == 0), "gsddmm: bad nnz/out");
  DGLB_CHECK_ARG(op == DGLB_OP_COPY_RHS || lhs || nnz == 0, "gsddmm: op needs lhs data");
  DGLB_CHECK_ARG(op == DGLB_OP_COPY_LHS || rhs || nnz == 0, "gsddmm: op needs rhs data");
  int rc = make_bcast(op, ndim, ls, rs_, b, reduce_size);
  if (rc != DGLB_OK) return rc;
  if (op == DGLB_OP_COPY_LHS) { for (int d = 0; d < b->ndim; ++d) { b->rhs[d] = b->lhs[d]; b->out[d] = b->lhs[d]; } b->rhs_len = b->out_len = b->lhs_len; }
  if (op == DGLB_OP_COPY_RHS) { for (int d = 0; d < b->ndim; ++d) { b->lhs[d] = b->rhs[d]; b->out[d] = b->rhs[d]; } b->lhs_len = b->out_len = b->rhs_len; }
  return DGLB_OK;
}

int dglb_gsddmm_csr(int op, int dtype, int lhs_target, int rhs_target, int64_t n_dst, int64_t n_src, int64_t nnz,
                    const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* lhs,
                    const void* rhs, int ndim, const int64_t* lhs_shape_host, const int64_t* rhs_shape_host,
                    void* out, const dglb_hub_t* hub, void* stream) {
  BcastShape b;
  int64_t rs;
  int rc = sddmm_common(op, dtype, lhs_target, rhs_target, ndim, lhs_shape_host, rhs_shape_host, lhs, rhs, out, nnz, &b, &rs);
  if (rc != DGLB_OK) return rc;
  DGLB_CHECK_ARG(indptr && (indices || nnz == 0), "gsddmm_csr: null graph pointer");
  if (nnz == 0 || n_dst == 0) return DGLB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (lhs_target == DGLB_TARGET_U && rhs_target == DGLB_TARGET_V) {
    rc = sddmm_csr_fast_f32(op, n_dst, n_src, nnz, indptr, indices, eids, static_cast<const float*>(lhs),
                            static_cast<const float*>(rhs), b, rs, static_cast<float*>(out), hub, st, dtype);
    if (rc != DGLB_E_UNSUPPORTED) return rc;
  }
  if (dtype != DGLB_F32) {
    set_error("gsddmm: bf16 storage is implemented for u_dot_v on the CSC (single head or power-of-two head segments)");
    return DGLB_E_UNSUPPORTED;
  }
  GenericSddmmParams g;
  g.src = nullptr; g.dst = nullptr; g.indptr = indptr; g.indices = indices; g.eids = eids;
  g.L = static_cast<const float*>(lhs); g.R = static_cast<const float*>(rhs); g.out = static_cast<float*>(out);
  g.nnz = nnz; g.n_rows = n_dst; g.op = op; g.lhs_target = lhs_target; g.rhs_target = rhs_target;
  g.reduce_size = rs; g.b = b;
  return sddmm_generic_f32(g, st);
}

int dglb_gsddmm_coo(int op, int dtype, int lhs_target, int rhs_target, int64_t n_src, int64_t n_dst, int64_t nnz,
                    const int32_t* src, const int32_t* dst, const void* lhs, const void* rhs, int ndim,
                    const int64_t* lhs_shape_host, const int64_t* rhs_shape_host, void* out, void* stream) {
  (void)n_src; (void)n_dst;
  BcastShape b;
  int64_t rs;
  int rc = sddmm_common(op, dtype, lhs_target, rhs_target, ndim, lhs_shape_host, rhs_shape_host, lhs, rhs, out, nnz, &b, &rs);
  if (rc != DGLB_OK) return rc;
  DGLB_CHECK_ARG((src && dst) || nnz == 0, "gsddmm_coo: null src/dst");
  if (dtype != DGLB_F32) { set_error("gsddmm_coo: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  if (nnz == 0) return DGLB_OK;
  rc = sddmm_coo_narrow_f32(op, lhs_target, rhs_target, nnz, src, dst, static_cast<const float*>(lhs),
                            static_cast<const float*>(rhs), b, rs, static_cast<float*>(out), static_cast<cudaStream_t>(stream));
  if (rc != DGLB_E_UNSUPPORTED) return rc;   // rows of <= 8 floats with equal shapes: one thread per edge
  GenericSddmmParams g;
  g.src = src; g.dst = dst; g.indptr = nullptr; g.indices = nullptr; g.eids = nullptr;
  g.L = static_cast<const float*>(lhs); g.R = static_cast<const float*>(rhs); g.out = static_cast<float*>(out);
  g.nnz = nnz; g.n_rows = 0; g.op = op; g.lhs_target = lhs_target; g.rhs_target = rhs_target;
  g.reduce_size = rs; g.b = b;
  return sddmm_generic_f32(g, static_cast<cudaStream_t>(stream));
}

int dglb_edge_softmax_fwd(int dtype, int64_t n_dst, int64_t nnz, int64_t n_heads, const int32_t* indptr,
                          const int32_t* eids, const void* logits, void* out, const dglb_hub_t* hub, void* stream) {
  if (dtype != DGLB_F32) { set_error("edge_softmax: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_dst >= 0 && nnz >= 0 && n_heads >= 0 && indptr, "edge_softmax_fwd: bad sizes / null indptr");
  DGLB_CHECK_ARG(nnz == 0 || (logits && out), "edge_softmax_fwd: null data");
  if (nnz == 0) return DGLB_OK;
  return edge_softmax_f32(false, n_dst, nnz, n_heads, indptr, eids, static_cast<const float*>(logits), nullptr,
                          static_cast<float*>(out), hub, static_cast<cudaStream_t>(stream));
}

int dglb_edge_softmax_bwd(int dtype, int64_t n_dst, int64_t nnz, int64_t n_heads, const int32_t* indptr,
                          const int32_t* eids, const void* out, const void* grad_out, void* grad_logits,
                          const dglb_hub_t* hub, void* stream) {
  if (dtype != DGLB_F32) { set_error("edge_softmax: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_dst >= 0 && nnz >= 0 && n_heads >= 0 && indptr, "edge_softmax_bwd: bad sizes / null indptr");
  DGLB_CHECK_ARG(nnz == 0 || (out && grad_out && grad_logits), "edge_softmax_bwd: null data");
  if (nnz == 0) return DGLB_OK;
  return edge_softmax_f32(true, n_dst, nnz, n_heads, indptr, eids, static_cast<const float*>(out),
                          static_cast<const float*>(grad_out), static_cast<float*>(grad_logits), hub,
                          static_cast<cudaStream_t>(stream));
}

static void gat_zero(GatParams& p) { memset(&p, 0, sizeof(p)); }

size_t dglb_gat_hub_workspace_bytes(int64_t n_seg, int64_t n_heads, int64_t head_dim) {
  return gat_hub_workspace_bytes(n_seg, n_heads, head_dim);
}

int dglb_gat_fused_fwd(int dtype, int64_t n_dst, int64_t n_src, int64_t nnz, int64_t n_heads, int64_t head_dim,
                       float negative_slope, float dropout_p, uint64_t seed, const int32_t* indptr,
                       const int32_t* indices, const int32_t* eids, const void* ft, const void* el, const void* er,
                       void* rst, float* row_max, float* row_sum, void* edge_scores, const dglb_hub_t* hub, void* stream) {
  (void)n_src;
  if (dtype != DGLB_F32) { set_error("gat_fused: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_dst >= 0 && nnz >= 0 && indptr && (indices || nnz == 0), "gat_fused_fwd: bad graph");
  DGLB_CHECK_ARG((ft && el) || nnz == 0, "gat_fused_fwd: null ft/el");
  DGLB_CHECK_ARG(n_dst == 0 || (er && rst && row_max && row_sum), "gat_fused_fwd: null er/rst/row stats");
  GatParams p;
  gat_zero(p);
  p.indptr = indptr; p.indices = indices; p.eids = eids;
  p.ft = static_cast<const float*>(ft); p.el = static_cast<const float*>(el); p.er = static_cast<const float*>(er);
  p.out_feat = static_cast<float*>(rst); p.out_h0 = row_max; p.out_h1 = row_sum;
  p.row_max = row_max; p.row_sum = row_sum;
  p.edge_scores = static_cast<float*>(edge_scores); p.n_rows = n_dst; p.slope = negative_slope;
  return gat_fused_f32(0, p, n_heads, head_dim, dropout_p, seed, hub, static_cast<cudaStream_t>(stream));
}

int dglb_gat_fused_bwd_dst(int dtype, int64_t n_dst, int64_t n_src, int64_t nnz, int64_t n_heads, int64_t head_dim,
                           float negative_slope, float dropout_p, uint64_t seed, const int32_t* indptr,
                           const int32_t* indices, const int32_t* eids, const void* ft, const void* el,
                           const void* er, const float* row_max, const float* row_sum, const void* grad_rst,
                           float* row_pack, void* grad_er, const dglb_hub_t* hub, void* stream) {
  (void)n_src;
  if (dtype != DGLB_F32) { set_error("gat_fused: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_dst >= 0 && nnz >= 0 && indptr && (indices || nnz == 0), "gat_fused_bwd_dst: bad graph");
  DGLB_CHECK_ARG(n_dst == 0 || (er && row_max && row_sum && grad_rst && row_pack && grad_er), "gat_fused_bwd_dst: null data");
  DGLB_CHECK_ARG((ft && el) || nnz == 0, "gat_fused_bwd_dst: null ft/el");
  DGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(row_pack) & 15) == 0, "gat_fused_bwd_dst: row_pack must be 16-byte aligned");
  GatParams p;
  gat_zero(p);
  p.indptr = indptr; p.indices = indices; p.eids = eids;
  p.ft = static_cast<const float*>(ft); p.el = static_cast<const float*>(el); p.er = static_cast<const float*>(er);
  p.row_max = row_max; p.row_sum = row_sum; p.dZ = static_cast<const float*>(grad_rst);
  p.out_pack = reinterpret_cast<float4*>(row_pack); p.out_h0 = static_cast<float*>(grad_er);
  p.n_rows = n_dst; p.slope = negative_slope;
  return gat_fused_f32(1, p, n_heads, head_dim, dropout_p, seed, hub, static_cast<cudaStream_t>(stream));
}

int dglb_gat_fused_bwd_src(int dtype, int64_t n_src, int64_t n_dst, int64_t nnz, int64_t n_heads, int64_t head_dim,
                           float negative_slope, float dropout_p, uint64_t seed, const int32_t* indptr_csr,
                           const int32_t* indices_csr, const int32_t* eids_csr, const void* ft, const void* el,
                           const float* row_pack, const void* grad_rst, void* grad_ft, void* grad_el,
                           const dglb_hub_t* hub, void* stream) {
  (void)n_dst;
  if (dtype != DGLB_F32) { set_error("gat_fused: only f32 is implemented"); return DGLB_E_UNSUPPORTED; }
  DGLB_CHECK_ARG(n_src >= 0 && nnz >= 0 && indptr_csr && (indices_csr || nnz == 0), "gat_fused_bwd_src: bad graph");
  DGLB_CHECK_ARG(n_src == 0 || (ft && el && grad_ft && grad_el), "gat_fused_bwd_src: null src data");
  DGLB_CHECK_ARG(nnz == 0 || (row_pack && grad_rst), "gat_fused_bwd_src: null dst data");
  DGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(row_pack) & 15) == 0, "gat_fused_bwd_src: row_pack must be 16-byte aligned");
  GatParams p;
  gat_zero(p);
  p.indptr = indptr_csr; p.indices = indices_csr; p.eids = eids_csr;
  p.ft = static_cast<const float*>(ft); p.el = static_cast<const float*>(el);
  p.pack = reinterpret_cast<const float4*>(row_pack); p.dZ = static_cast<const float*>(grad_rst);
  p.out_feat = static_cast<float*>(grad_ft); p.out_h0 = static_cast<float*>(grad_el);
  p.n_rows = n_src; p.slope = negative_slope;
  return gat_fused_f32(2, p, n_heads, head_dim, dropout_p, seed, hub, static_cast<cudaStream_t>(stream));
}

int dglb_gcn_msg_sum_fwd(int64_t n_dst, int64_t n_src, int64_t nnz, int64_t feat_len, const int32_t* indptr,
                         const int32_t* indices, const int32_t* eids, const float* x, const float* w, const float* c_src,
                         const float* c_dst, float* out, void* stream) {
  DGLB_CHECK_ARG(n_dst >= 0 && n_src >= 0 && nnz >= 0 && feat_len >= 0, "gcn_msg_sum_fwd: negative size");
  DGLB_CHECK_ARG(n_dst == 0 || (indptr && c_dst && (out || feat_len == 0)), "gcn_msg_sum_fwd: null dst data");
  DGLB_CHECK_ARG(nnz == 0 || (indices && x && w && c_src), "gcn_msg_sum_fwd: null operand with nnz > 0");
  return gcn_msg_sum_fwd(n_dst, feat_len, indptr, indices, eids, x, w, c_src, c_dst, out, static_cast<cudaStream_t>(stream));
}

int dglb_gcn_msg_sum_bwd(int64_t n_src, int64_t n_dst, int64_t nnz, int64_t feat_len, const int32_t* indptr_csr,
                         const int32_t* indices_csr, const int32_t* eids_csr, const float* x, const float* w,
                         const float* c_src, const float* c_dst, const float* grad_out, float* grad_x, float* grad_w,
                         void* stream) {
  DGLB_CHECK_ARG(n_dst >= 0 && n_src >= 0 && nnz >= 0 && feat_len >= 0, "gcn_msg_sum_bwd: negative size");
  DGLB_CHECK_ARG(n_src == 0 || (indptr_csr && c_src && x && (grad_x || feat_len == 0)), "gcn_msg_sum_bwd: null src data");
  DGLB_CHECK_ARG(nnz == 0 || (indices_csr && w && c_dst && grad_out && grad_w), "gcn_msg_sum_bwd: null operand with nnz > 0");
  return gcn_msg_sum_bwd(n_src, feat_len, indptr_csr, indices_csr, eids_csr, x, w, c_src, c_dst, grad_out, grad_x, grad_w,
                         static_cast<cudaStream_t>(stream));
}

int dglb_cat_embed_sum_fwd(int64_t n_rows, int64_t n_columns, int64_t feat_len, const int64_t* x, const int32_t* offsets_host,
                           const float* table, float* out, void* stream) {
  DGLB_CHECK_ARG(n_rows >= 0 && n_columns >= 0 && feat_len >= 0 && offsets_host, "cat_embed_sum_fwd: bad sizes / offsets");
  DGLB_CHECK_ARG(n_rows == 0 || feat_len == 0 || (out && (n_columns == 0 || (x && table))), "cat_embed_sum_fwd: null array");
  return cat_embed_sum(false, n_rows, n_columns, feat_len, x, offsets_host, table, nullptr, out, nullptr, 0,
                       static_cast<cudaStream_t>(stream));
}

size_t dglb_cat_embed_sum_bwd_workspace_bytes(int64_t n_rows, int64_t n_table_rows, int64_t feat_len) {
  return cat_embed_bwd_workspace_bytes(n_rows, n_table_rows, feat_len);
}

int dglb_cat_embed_sum_bwd(int64_t n_rows, int64_t n_columns, int64_t feat_len, const int64_t* x, const int32_t* offsets_host,
                           const float* grad_out, float* grad_table, void* workspace, size_t workspace_bytes, void* stream) {
  DGLB_CHECK_ARG(n_rows >= 0 && n_columns >= 0 && feat_len >= 0 && offsets_host, "cat_embed_sum_bwd: bad sizes / offsets");
  DGLB_CHECK_ARG(n_columns == 0 || feat_len == 0 || offsets_host[n_columns] == 0 || (grad_table && (n_rows == 0 || (x && grad_out))),
                 "cat_embed_sum_bwd: null array");
  return cat_embed_sum(true, n_rows, n_columns, feat_len, x, offsets_host, nullptr, grad_out, grad_table, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream));
}

int dglb_batch_offsets(int64_t n_sel, const int32_t* graph_ids, const int32_t* node_ptr, const int32_t* edge_ptr,
                       int32_t* out_node_ptr, int32_t* out_edge_ptr, int64_t n_nodes_pad, int64_t n_edges_pad,
                       int32_t* status, void* stream) {
  DGLB_CHECK_ARG(n_sel >= 0 && n_sel < (1LL << 31) && n_nodes_pad >= 0 && n_nodes_pad < (1LL << 31) && n_edges_pad >= 0 &&
                     n_edges_pad < (1LL << 31), "batch_offsets: sizes must fit int32");
  DGLB_CHECK_ARG(out_node_ptr && out_edge_ptr && node_ptr && edge_ptr && (graph_ids || n_sel == 0),
                 "batch_offsets: null array");
  return batch_offsets(n_sel, graph_ids, node_ptr, edge_ptr, out_node_ptr, out_edge_ptr, n_nodes_pad, n_edges_pad, status,
                       static_cast<cudaStream_t>(stream));
}

int dglb_batch_gather(const dglb_batch_io_t* io, void* stream) {
  DGLB_CHECK_ARG(io != nullptr, "batch_gather: io is null");
  DGLB_CHECK_ARG(io->n_sel >= 0 && io->n_nodes_pad >= 0 && io->n_edges_pad >= 0, "batch_gather: negative size");
  DGLB_CHECK_ARG(io->node_ptr && io->edge_ptr && io->out_node_ptr && io->out_edge_ptr && (io->graph_ids || io->n_sel == 0),
                 "batch_gather: null selection / offset array");
  DGLB_CHECK_ARG(!(io->src || io->dst) || (io->u_src && io->u_dst), "batch_gather: COO outputs need the union COO");
  DGLB_CHECK_ARG(!(io->csc_indptr || io->csc_indices) || (io->u_csc_indptr && io->u_csc_indices),
                 "batch_gather: CSC outputs need the union CSC");
  DGLB_CHECK_ARG(!(io->csr_indptr || io->csr_indices) || (io->u_csr_indptr && io->u_csr_indices),
                 "batch_gather: CSR outputs need the union CSR");
  DGLB_CHECK_ARG((io->csc_indices == nullptr) == (io->csc_eids == nullptr) &&
                     (io->csr_indices == nullptr) == (io->csr_eids == nullptr),
                 "batch_gather: indices and eids outputs come in pairs");
  return batch_gather(*io, static_cast<cudaStream_t>(stream));
}

int dglb_copy_rows_indexed(int64_t n_idx, const int32_t* idx, int64_t row_bytes, const void* src, void* dst, void* stream) {
  DGLB_CHECK_ARG(n_idx >= 0 && row_bytes >= 0, "copy_rows_indexed: negative size");
  DGLB_CHECK_ARG(n_idx == 0 || row_bytes == 0 || (idx && src && dst), "copy_rows_indexed: null array");
  return copy_rows_indexed(n_idx, idx, row_bytes, src, dst, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
