// kernels.cuh -- parameter blocks and entry points shared between capi.cu and the kernel
// translation units.
#pragma once
#include "common.cuh"

namespace dglb {

struct GenericSddmmParams {
  // COO form: src/dst by edge id.  CSR form: indptr/indices/eids (row found by binary search).
  const int32_t* src;
  const int32_t* dst;
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eids;
  const float* L;
  const float* R;
  float* out;
  int64_t nnz, n_rows;
  int op, lhs_target, rhs_target;
  int64_t reduce_size;
  BcastShape b;
};

struct GatParams {
  const int32_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;  // neighbour ids (src for CSC, dst for CSR)
  const int32_t* __restrict__ eids;     // edge ids (dropout counter, edge_scores); null = position
  const float* __restrict__ ft;         // (n_src, H, F)
  const float* __restrict__ el;         // (n_src, H)
  const float* __restrict__ er;         // (n_dst, H)
  const float* __restrict__ row_max;    // (n_dst, H)
  const float* __restrict__ row_sum;    // (n_dst, H)
  const float4* __restrict__ pack;      // (n_dst, H) {er, max, sum, s1}   (bwd_src: read)
  const float* __restrict__ dZ;         // (n_dst, H, F) grad of rst
  float* __restrict__ out_feat;         // fwd: rst (n_dst,H,F); bwd_src: grad_ft (n_src,H,F)
  float* __restrict__ out_h0;           // rowstats: row_max; bwd_dst: grad_er; bwd_src: grad_el
  float* __restrict__ out_h1;           // rowstats: row_sum
  float4* __restrict__ out_pack;        // bwd_dst: (n_dst, H) {er, max, sum, s1}
  float* __restrict__ edge_scores;      // fwd only, may be null
  const int32_t* __restrict__ hub_rows;
  const int32_t* __restrict__ seg_ptr;  // [n_hub+1] segmented hub path (null: one CTA per hub row)
  const int32_t* __restrict__ seg_hub;  // [n_seg]
  float* __restrict__ ws_feat;          // [n_seg][D]    per-segment partial feature rows
  float* __restrict__ ws_tot;           // [n_seg][H][4] per-segment {max,sum} (fwd) / {S1,S2,S3} partials (bwd)
  int seg_len, n_seg, n_hub;
  int64_t n_rows;
  int H, F, D;  // D = H*F
  int ncols;    // D / VEC
  int G, log2G;
  int HP, log2HP;  // rowstats: heads padded to a power of two (lanes per edge)
  int hub_threshold;
  float slope;
  float drop_p, drop_scale;  // drop_scale = 1/(1-p)
  uint32_t seed_lo, seed_hi;
  int prefetch;              // fwd: request the first gather batch before the softmax statistics
};

// graph_build.cu
size_t coo_to_csr_workspace_bytes(int64_t n_rows, int64_t nnz);
int coo_to_csr(int64_t n_rows, int64_t nnz, const int32_t* row, const int32_t* col, int32_t* indptr,
               int32_t* indices, int32_t* data, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int csr_degrees(int64_t n_rows, const int32_t* indptr, int32_t* deg, cudaStream_t stream);
int is_identity_perm(int64_t n, const int32_t* data, int32_t* flag, cudaStream_t stream);
int csr_find_hub_rows(int64_t n_rows, const int32_t* indptr, int32_t threshold, int32_t* hub_rows,
                      int64_t cap, int32_t* n_hub, cudaStream_t stream);
// spmm.cu
int spmm_csr_f32(int op, int reduce, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* indptr,
                 const int32_t* indices, const int32_t* eids, const float* X, const float* W,
                 const BcastShape& b, float* out, int32_t* arg_u, int32_t* arg_e, const float* row_scale,
                 int accumulate, const dglb_hub_t* hub, cudaStream_t stream);
int spmm_csr_bf16(int op, int reduce, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* indptr,
                  const int32_t* indices, const void* X, int64_t D, void* out, const float* row_scale,
                  int accumulate, const dglb_hub_t* hub, cudaStream_t stream);
// sddmm.cu
int sddmm_csr_fast_f32(int op, int64_t n_dst, int64_t n_src, int64_t nnz, const int32_t* indptr,
                       const int32_t* indices, const int32_t* eids, const float* Uf, const float* Vf,
                       const BcastShape& b, int64_t reduce_size, float* out, const dglb_hub_t* hub,
                       cudaStream_t stream, int dtype = DGLB_F32);
// ring.cu: the non-hub rows of gspmm(copy_lhs, sum) (dot = false) or gsddmm(u_dot_v) (dot = true) through the
// bulk-copy ring kernel; DGLB_E_UNSUPPORTED (no error set) when the shape is outside its range
int ring_rows(bool dot, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* indptr,
              const int32_t* indices, const int32_t* eids, const void* X, const void* V, int64_t D, void* out,
              const float* row_scale, int accumulate, int hub_threshold, const int32_t* light_indptr,
              cudaStream_t stream);
int sddmm_generic_f32(const GenericSddmmParams& g, cudaStream_t stream);
int sddmm_coo_narrow_f32(int op, int lhs_target, int rhs_target, int64_t nnz, const int32_t* src, const int32_t* dst,
                         const float* L, const float* R, const BcastShape& b, int64_t reduce_size, float* out,
                         cudaStream_t stream);
// edge_stage.cu
size_t edge_stage_plan_workspace_bytes(int64_t nnz, int log2_bucket);
int edge_stage_plan(int64_t nnz, const int32_t* eids, int log2_bucket, int32_t* stage_pos, int32_t* slot,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);
int edge_stage_move(int to_staged, int64_t nnz, int64_t row_words, const int32_t* stage_pos, const void* src, void* dst,
                    cudaStream_t stream);
// edge_softmax.cu
int edge_softmax_f32(bool bwd, int64_t n_dst, int64_t nnz, int64_t n_heads, const int32_t* indptr, const int32_t* eids,
                     const float* a, const float* b, float* out, const dglb_hub_t* hub, cudaStream_t stream);
size_t edge_softmax_workspace_bytes(int64_t n_seg, int64_t n_hub, int64_t n_heads);
// gat_fused.cu  (which: 0 fwd, 1 bwd_dst, 2 bwd_src)
int gat_fused_f32(int which, GatParams& p, int64_t H, int64_t F, float dropout_p, uint64_t seed,
                  const dglb_hub_t* hub, cudaStream_t stream);
size_t gat_hub_workspace_bytes(int64_t n_seg, int64_t H, int64_t F);

// small_graph.cu
int gcn_msg_sum_fwd(int64_t n_dst, int64_t D, const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                    const float* x, const float* w, const float* c_src, const float* c_dst, float* out,
                    cudaStream_t stream);
int gcn_msg_sum_bwd(int64_t n_src, int64_t D, const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                    const float* x, const float* w, const float* c_src, const float* c_dst, const float* gout, float* gx,
                    float* gw, cudaStream_t stream);
int batch_offsets(int64_t n_sel, const int32_t* graph_ids, const int32_t* node_ptr, const int32_t* edge_ptr,
                  int32_t* out_node_ptr, int32_t* out_edge_ptr, int64_t n_nodes_pad, int64_t n_edges_pad, int32_t* status,
                  cudaStream_t stream);
int batch_gather(const dglb_batch_io_t& io, cudaStream_t stream);
int cat_embed_sum(bool bwd, int64_t n_rows, int64_t K, int64_t D, const int64_t* x, const int32_t* offsets_host,
                  const float* T, const float* g, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t cat_embed_bwd_workspace_bytes(int64_t n_rows, int64_t n_table_rows, int64_t D);

// row_copy.cu
int copy_rows_indexed(int64_t n_idx, const int32_t* idx, int64_t row_bytes, const void* src, void* dst, cudaStream_t stream);

}  // namespace dglb
