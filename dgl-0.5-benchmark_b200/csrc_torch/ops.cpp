// ops.cpp -- PyTorch C++ extension (TORCH_LIBRARY "dglb200") over the C-ABI of include/dglb200.h.
//
// This is the layer upstream DGL v0.6.1 implements as python/dgl/sparse.py::_gspmm/_gsddmm (allocate the outputs,
// hand zero-copy NDArrays to the packed functions `_CAPI_DGLKernelSpMM` / `_CAPI_DGLKernelSDDMM`) plus the argument
// checks of src/array/kernel.cc (CheckCtx / CheckContiguous): every op below
//   * checks device / dtype / contiguity of what it is given (TORCH_CHECK -> RuntimeError, re-raised as DGLError by
//     the Python layer),
//   * allocates outputs and workspaces through torch's caching allocator (so a step can be captured in a CUDA graph
//     with torch-owned memory),
//   * fetches torch's CURRENT stream of the tensors' device under a device guard,
//   * calls the plain-C entry point and turns a non-zero status into an error carrying dglb_last_error().
// No arithmetic happens here; the kernels live in lib/libdglb200.so (csrc/*.cu).  The Python side
// (dgl/sparse.py) keeps only what upstream's Python keeps: op-name validation and broadcast-shape inference.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/dglb200.h"

namespace {

using at::Tensor;
using OptTensor = std::optional<Tensor>;

void check_status(int rc, const char* what) {
  TORCH_CHECK(rc == DGLB_OK, what, " failed (status ", rc, "): ", dglb_last_error());
}

const Tensor& need_cuda(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), "dgl-b200 sparse kernels are CUDA-only (sm_100a); ", name, " is on ", t.device(),
              ". There is no CPU fallback.");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  return t;
}

const int32_t* i32(const Tensor& t, const char* name) {
  need_cuda(t, name);
  TORCH_CHECK(t.scalar_type() == at::kInt, name, " must be int32");
  return t.data_ptr<int32_t>();
}

const int32_t* i32_opt(const OptTensor& t, const char* name) { return t.has_value() ? i32(*t, name) : nullptr; }

const void* feat_opt(const OptTensor& t, const char* name, c10::DeviceIndex dev) {
  if (!t.has_value()) return nullptr;
  need_cuda(*t, name);
  TORCH_CHECK(t->get_device() == dev, "expected all operands on cuda:", (int)dev, ", ", name, " is on ", t->device());
  return t->data_ptr();
}

const Tensor& cuda_anchor(const Tensor& t) {
  TORCH_CHECK(t.is_cuda(), "dgl-b200 sparse kernels are CUDA-only (sm_100a); got a tensor on ", t.device(),
              ". There is no CPU fallback.");
  return t;
}

struct Entered {
  c10::cuda::CUDAGuard guard;
  void* stream;
  explicit Entered(const Tensor& anchor) : guard(cuda_anchor(anchor).device()) {
    check_status(dglb_set_device(anchor.get_device()), "dglb_set_device");
    stream = at::cuda::getCurrentCUDAStream(anchor.get_device()).stream();
  }
};

// hub_meta = {n_hub, n_seg, seg_len, threshold}; returns false when no hub rows were passed
bool make_hub(dglb_hub_t* h, const OptTensor& rows, const OptTensor& seg_ptr, const OptTensor& seg_hub,
              const OptTensor& light, at::IntArrayRef meta, const OptTensor& order = std::nullopt) {
  h->row_order = i32_opt(order, "hub row_order");
  if (!rows.has_value() || meta.size() != 4 || meta[0] <= 0) {
    if (!order.has_value()) return false;
    // no hub rows, but a degree-ordered row hand-out for the row kernels
    h->rows = h->seg_ptr = h->seg_hub = h->light_indptr = nullptr;
    h->n_hub = h->n_seg = h->seg_len = 0; h->threshold = INT32_MAX;
    h->workspace = nullptr; h->workspace_bytes = 0;
    return true;
  }
  h->rows = i32(*rows, "hub rows");
  h->seg_ptr = i32_opt(seg_ptr, "hub seg_ptr");
  h->seg_hub = i32_opt(seg_hub, "hub seg_hub");
  h->light_indptr = i32_opt(light, "hub light_indptr");
  h->n_hub = (int32_t)meta[0]; h->n_seg = (int32_t)meta[1]; h->seg_len = (int32_t)meta[2]; h->threshold = (int32_t)meta[3];
  h->workspace = nullptr; h->workspace_bytes = 0;
  return true;
}

int dtype_code(const Tensor& t) {
  if (t.scalar_type() == at::kFloat) return DGLB_F32;
  if (t.scalar_type() == at::kBFloat16) return DGLB_BF16;
  TORCH_CHECK(false, "dgl-b200 kernels take float32 (or bfloat16 storage where supported); got ", t.scalar_type());
}

std::vector<int64_t> with_rows(int64_t rows, at::IntArrayRef feat) {
  std::vector<int64_t> s{rows};
  s.insert(s.end(), feat.begin(), feat.end());
  return s;
}

int64_t prod(at::IntArrayRef v) { int64_t p = 1; for (auto x : v) p *= x; return p; }

// ------------------------------------------------------------------ graph build
std::tuple<Tensor, Tensor, Tensor> coo_to_csr(const Tensor& row, const Tensor& col, int64_t n_rows) {
  const int64_t nnz = row.numel();
  i32(row, "row"); i32(col, "col");
  auto opt = row.options();
  Tensor indptr = at::empty({n_rows + 1}, opt), indices = at::empty({nnz}, opt), data = at::empty({nnz}, opt);
  const size_t ws_bytes = dglb_coo_to_csr_workspace_bytes(n_rows, nnz);
  Tensor ws = at::empty({(int64_t)std::max<size_t>(ws_bytes, 1)}, opt.dtype(at::kByte));
  Entered en(row);
  check_status(dglb_coo_to_csr(n_rows, nnz, row.data_ptr<int32_t>(), col.data_ptr<int32_t>(), indptr.data_ptr<int32_t>(),
                               indices.data_ptr<int32_t>(), data.data_ptr<int32_t>(), ws.data_ptr(), ws_bytes, en.stream),
               "dglb_coo_to_csr");
  return {indptr, indices, data};
}

Tensor csr_degrees(const Tensor& indptr) {
  const int64_t n = indptr.numel() - 1;
  Tensor deg = at::empty({n}, indptr.options());
  if (n > 0) {
    Entered en(indptr);
    check_status(dglb_csr_degrees(n, i32(indptr, "indptr"), deg.data_ptr<int32_t>(), en.stream), "dglb_csr_degrees");
  }
  return deg;
}

Tensor is_identity_perm(const Tensor& data) {
  Tensor flag = at::empty({1}, data.options());
  Entered en(data);
  check_status(dglb_is_identity_perm(data.numel(), i32(data, "data"), flag.data_ptr<int32_t>(), en.stream),
               "dglb_is_identity_perm");
  return flag;
}

std::tuple<Tensor, Tensor> find_hub_rows(const Tensor& indptr, int64_t threshold, int64_t cap) {
  Tensor n_hub = at::zeros({1}, indptr.options()), rows = at::empty({std::max<int64_t>(cap, 1)}, indptr.options());
  Entered en(indptr);
  check_status(dglb_csr_find_hub_rows(indptr.numel() - 1, i32(indptr, "indptr"), (int32_t)threshold,
                                      rows.data_ptr<int32_t>(), cap, n_hub.data_ptr<int32_t>(), en.stream),
               "dglb_csr_find_hub_rows");
  return {rows, n_hub};
}

std::tuple<Tensor, Tensor> edge_stage_plan(const Tensor& eids, int64_t log2_bucket) {
  const int64_t nnz = eids.numel();
  Tensor stage_pos = at::empty({nnz}, eids.options()), slot = at::empty({nnz}, eids.options());
  const size_t ws_bytes = dglb_edge_stage_plan_workspace_bytes(nnz, (int)log2_bucket);
  Tensor ws = at::empty({(int64_t)std::max<size_t>(ws_bytes, 1)}, eids.options().dtype(at::kByte));
  Entered en(eids);
  check_status(dglb_edge_stage_plan(nnz, i32(eids, "eids"), (int)log2_bucket, stage_pos.data_ptr<int32_t>(),
                                    slot.data_ptr<int32_t>(), ws.data_ptr(), ws_bytes, en.stream),
               "dglb_edge_stage_plan");
  return {stage_pos, slot};
}

Tensor edge_stage(const Tensor& stage_pos, const Tensor& t, bool to_staged) {
  need_cuda(t, "per-edge tensor");
  Tensor out = at::empty_like(t);
  const int64_t n = t.size(0);
  if (n == 0) return out;
  Entered en(t);
  check_status(dglb_edge_stage(to_staged ? 1 : 0, n, (t.numel() / n) * t.element_size(), i32(stage_pos, "stage_pos"),
                               t.data_ptr(), out.data_ptr(), en.stream), "dglb_edge_stage");
  return out;
}

// ------------------------------------------------------------------ gspmm
// out_feat: broadcast feature shape (Python: infer_broadcast_shape); lhs_shape / rhs_shape: right-aligned trailing
// shapes handed to the C-ABI.  Returns (out, arg_u, arg_e) (args: empty tensors unless reduce is max / min).
std::tuple<Tensor, Tensor, Tensor> gspmm(
    const Tensor& indptr, const Tensor& indices, const OptTensor& eids, int64_t n_cols, int64_t op, int64_t reduce,
    const OptTensor& u, const OptTensor& e, at::IntArrayRef out_feat, at::IntArrayRef lhs_shape, at::IntArrayRef rhs_shape,
    const OptTensor& row_scale, const OptTensor& out_in, int64_t flags, const OptTensor& hub_rows,
    const OptTensor& hub_seg_ptr, const OptTensor& hub_seg_hub, const OptTensor& hub_light, const OptTensor& hub_order,
    at::IntArrayRef hub_meta) {
  const Tensor& ref = u.has_value() ? *u : *e;
  const auto dev = indptr.get_device();
  const int64_t n_rows = indptr.numel() - 1, nnz = indices.numel();
  const int dtype = dtype_code(ref);
  const bool cmp = reduce != DGLB_REDUCE_SUM;
  Tensor out = out_in.has_value() ? *out_in : at::empty(with_rows(n_rows, out_feat), ref.options());
  Tensor arg_u, arg_e;
  if (cmp) {
    if (u.has_value()) arg_u = at::empty(with_rows(n_rows, out_feat), indptr.options());
    if (e.has_value()) arg_e = at::empty(with_rows(n_rows, out_feat), indptr.options());
  }
  dglb_hub_t hub;
  Tensor ws;
  const bool has_hub = make_hub(&hub, hub_rows, hub_seg_ptr, hub_seg_hub, hub_light, hub_meta, hub_order);
  if (has_hub && hub.n_hub > 0) {
    const size_t need = dglb_hub_workspace_bytes(hub.n_seg, prod(out_feat), cmp ? 1 : 0);
    ws = at::empty({(int64_t)std::max<size_t>(need, 4)}, ref.options().dtype(at::kByte));
    hub.workspace = ws.data_ptr(); hub.workspace_bytes = need;
  }
  TORCH_CHECK(lhs_shape.size() == rhs_shape.size() && !lhs_shape.empty(), "gspmm: bad trailing shapes");
  Entered en(indptr);
  check_status(dglb_gspmm_csr((int)op, (int)reduce, dtype, n_rows, n_cols, nnz, i32(indptr, "indptr"), i32(indices, "indices"),
                              i32_opt(eids, "eids"), feat_opt(u, "node data", dev), feat_opt(e, "edge data", dev),
                              (int)lhs_shape.size(), lhs_shape.data(), rhs_shape.data(), need_cuda(out, "out").data_ptr(),
                              arg_u.defined() ? arg_u.data_ptr<int32_t>() : nullptr,
                              arg_e.defined() ? arg_e.data_ptr<int32_t>() : nullptr,
                              row_scale.has_value() ? need_cuda(*row_scale, "row_scale").data_ptr<float>() : nullptr,
                              (int)flags, has_hub ? &hub : nullptr, en.stream),
               "dglb_gspmm_csr");
  return {out, arg_u.defined() ? arg_u : at::empty({0}, indptr.options()),
          arg_e.defined() ? arg_e : at::empty({0}, indptr.options())};
}

// ------------------------------------------------------------------ gsddmm
Tensor gsddmm_csr(const Tensor& indptr, const Tensor& indices, const OptTensor& eids, int64_t n_src, int64_t op,
                  int64_t lhs_target, int64_t rhs_target, const OptTensor& lhs, const OptTensor& rhs,
                  at::IntArrayRef out_feat, at::IntArrayRef lhs_shape, at::IntArrayRef rhs_shape, const OptTensor& hub_rows,
                  const OptTensor& hub_seg_ptr, const OptTensor& hub_seg_hub, const OptTensor& hub_light,
                  const OptTensor& hub_order, at::IntArrayRef hub_meta) {
  const Tensor& ref = lhs.has_value() ? *lhs : *rhs;
  const auto dev = indptr.get_device();
  const int64_t n_dst = indptr.numel() - 1, nnz = indices.numel();
  Tensor out = at::empty(with_rows(nnz, out_feat), ref.options());
  if (nnz == 0 || out.numel() == 0) return out;
  dglb_hub_t hub;
  const bool has_hub = make_hub(&hub, hub_rows, hub_seg_ptr, hub_seg_hub, hub_light, hub_meta, hub_order);
  Entered en(indptr);
  check_status(dglb_gsddmm_csr((int)op, dtype_code(ref), (int)lhs_target, (int)rhs_target, n_dst, n_src, nnz,
                               i32(indptr, "indptr"), i32(indices, "indices"), i32_opt(eids, "eids"),
                               feat_opt(lhs, "lhs", dev), feat_opt(rhs, "rhs", dev), (int)lhs_shape.size(), lhs_shape.data(),
                               rhs_shape.data(), out.data_ptr(), has_hub ? &hub : nullptr, en.stream),
               "dglb_gsddmm_csr");
  return out;
}

Tensor gsddmm_coo(const Tensor& src, const Tensor& dst, int64_t n_src, int64_t n_dst, int64_t op, int64_t lhs_target,
                  int64_t rhs_target, const OptTensor& lhs, const OptTensor& rhs, at::IntArrayRef out_feat,
                  at::IntArrayRef lhs_shape, at::IntArrayRef rhs_shape) {
  const Tensor& ref = lhs.has_value() ? *lhs : *rhs;
  const auto dev = src.get_device();
  const int64_t nnz = src.numel();
  Tensor out = at::empty(with_rows(nnz, out_feat), ref.options());
  if (nnz == 0 || out.numel() == 0) return out;
  Entered en(src);
  check_status(dglb_gsddmm_coo((int)op, dtype_code(ref), (int)lhs_target, (int)rhs_target, n_src, n_dst, nnz, i32(src, "src"),
                               i32(dst, "dst"), feat_opt(lhs, "lhs", dev), feat_opt(rhs, "rhs", dev), (int)lhs_shape.size(),
                               lhs_shape.data(), rhs_shape.data(), out.data_ptr(), en.stream),
               "dglb_gsddmm_coo");
  return out;
}

// ------------------------------------------------------------------ edge_softmax
Tensor edge_softmax(bool bwd, const Tensor& indptr, const OptTensor& eids, const Tensor& a, const OptTensor& b, int64_t heads,
                    const OptTensor& hub_rows, const OptTensor& hub_seg_ptr, const OptTensor& hub_seg_hub,
                    at::IntArrayRef hub_meta) {
  need_cuda(a, "edge data");
  TORCH_CHECK(a.scalar_type() == at::kFloat, "dgl-b200 kernels compute in float32; got ", a.scalar_type());
  Tensor out = at::empty_like(a);
  const int64_t nnz = a.size(0), n_dst = indptr.numel() - 1;
  if (nnz == 0) return out;
  dglb_hub_t hub;
  Tensor ws;
  const bool has_hub = heads <= 32 && make_hub(&hub, hub_rows, hub_seg_ptr, hub_seg_hub, std::nullopt, hub_meta);
  if (has_hub) {
    const size_t need = dglb_edge_softmax_workspace_bytes(hub.n_seg, hub.n_hub, heads);
    ws = at::empty({(int64_t)std::max<size_t>(need, 4)}, a.options().dtype(at::kByte));
    hub.workspace = ws.data_ptr(); hub.workspace_bytes = need;
  }
  Entered en(a);
  if (!bwd) {
    check_status(dglb_edge_softmax_fwd(DGLB_F32, n_dst, nnz, heads, i32(indptr, "indptr"), i32_opt(eids, "eids"), a.data_ptr(),
                                       out.data_ptr(), has_hub ? &hub : nullptr, en.stream), "dglb_edge_softmax_fwd");
  } else {
    TORCH_CHECK(b.has_value(), "edge_softmax_bwd needs grad_out");
    check_status(dglb_edge_softmax_bwd(DGLB_F32, n_dst, nnz, heads, i32(indptr, "indptr"), i32_opt(eids, "eids"), a.data_ptr(),
                                       need_cuda(*b, "grad_out").data_ptr(), out.data_ptr(), has_hub ? &hub : nullptr,
                                       en.stream), "dglb_edge_softmax_bwd");
  }
  return out;
}

Tensor edge_softmax_fwd(const Tensor& indptr, const OptTensor& eids, const Tensor& logits, int64_t heads,
                        const OptTensor& hr, const OptTensor& hs, const OptTensor& hh, at::IntArrayRef hm) {
  return edge_softmax(false, indptr, eids, logits, std::nullopt, heads, hr, hs, hh, hm);
}

Tensor edge_softmax_bwd(const Tensor& indptr, const OptTensor& eids, const Tensor& out, const Tensor& grad_out, int64_t heads,
                        const OptTensor& hr, const OptTensor& hs, const OptTensor& hh, at::IntArrayRef hm) {
  return edge_softmax(true, indptr, eids, out, grad_out, heads, hr, hs, hh, hm);
}

// ------------------------------------------------------------------ fused GAT
struct GatHub {
  dglb_hub_t hub;
  Tensor ws;
  bool on = false;
  GatHub(const Tensor& like, int64_t H, int64_t F, bool segments, const OptTensor& hr, const OptTensor& hs,
         const OptTensor& hh, at::IntArrayRef hm) {
    on = make_hub(&hub, hr, hs, hh, std::nullopt, hm);
    if (on && segments) {
      const size_t need = dglb_gat_hub_workspace_bytes(hub.n_seg, H, F);
      ws = at::empty({(int64_t)std::max<size_t>(need, 16)}, like.options().dtype(at::kByte));
      hub.workspace = ws.data_ptr(); hub.workspace_bytes = need;
    }
  }
  const dglb_hub_t* ptr() const { return on ? &hub : nullptr; }
};

std::tuple<Tensor, Tensor, Tensor, Tensor> gat_fwd(const Tensor& indptr, const Tensor& indices, const OptTensor& eids,
                                                   const Tensor& ft, const Tensor& el, const Tensor& er, double slope,
                                                   double dropout_p, int64_t seed, bool want_scores, bool hub_segments,
                                                   const OptTensor& hr, const OptTensor& hs, const OptTensor& hh,
                                                   at::IntArrayRef hm) {
  need_cuda(ft, "ft"); need_cuda(el, "el"); need_cuda(er, "er");
  TORCH_CHECK(ft.scalar_type() == at::kFloat && el.scalar_type() == at::kFloat && er.scalar_type() == at::kFloat,
              "dgl-b200 kernels compute in float32");
  const int64_t n_dst = indptr.numel() - 1, nnz = indices.numel(), H = ft.size(1), F = ft.size(2);
  Tensor rst = at::empty({n_dst, H, F}, ft.options()), row_max = at::empty({n_dst, H}, ft.options()),
         row_sum = at::empty({n_dst, H}, ft.options());
  Tensor scores = want_scores ? at::empty({nnz, H}, ft.options()) : at::empty({0}, ft.options());
  if (n_dst == 0) return {rst, row_max, row_sum, scores};
  GatHub gh(ft, H, F, hub_segments, hr, hs, hh, hm);
  Entered en(ft);
  check_status(dglb_gat_fused_fwd(DGLB_F32, n_dst, ft.size(0), nnz, H, F, (float)slope, (float)dropout_p, (uint64_t)seed,
                                  i32(indptr, "indptr"), i32(indices, "indices"), i32_opt(eids, "eids"), ft.data_ptr(),
                                  el.data_ptr(), er.data_ptr(), rst.data_ptr(), row_max.data_ptr<float>(),
                                  row_sum.data_ptr<float>(), want_scores ? scores.data_ptr() : nullptr, gh.ptr(), en.stream),
               "dglb_gat_fused_fwd");
  return {rst, row_max, row_sum, scores};
}

std::tuple<Tensor, Tensor> gat_bwd_dst(const Tensor& indptr, const Tensor& indices, const OptTensor& eids, const Tensor& ft,
                                       const Tensor& el, const Tensor& er, const Tensor& row_max, const Tensor& row_sum,
                                       const Tensor& grad_rst, double slope, double dropout_p, int64_t seed,
                                       bool hub_segments, const OptTensor& hr, const OptTensor& hs, const OptTensor& hh,
                                       at::IntArrayRef hm) {
  const int64_t n_dst = indptr.numel() - 1, nnz = indices.numel(), H = ft.size(1), F = ft.size(2);
  Tensor row_pack = at::empty({n_dst, H, 4}, ft.options()), grad_er = at::empty({n_dst, H}, ft.options());
  if (n_dst == 0) return {row_pack, grad_er};
  GatHub gh(ft, H, F, hub_segments, hr, hs, hh, hm);
  Entered en(ft);
  check_status(dglb_gat_fused_bwd_dst(DGLB_F32, n_dst, ft.size(0), nnz, H, F, (float)slope, (float)dropout_p, (uint64_t)seed,
                                      i32(indptr, "indptr"), i32(indices, "indices"), i32_opt(eids, "eids"),
                                      need_cuda(ft, "ft").data_ptr(), need_cuda(el, "el").data_ptr(),
                                      need_cuda(er, "er").data_ptr(), need_cuda(row_max, "row_max").data_ptr<float>(),
                                      need_cuda(row_sum, "row_sum").data_ptr<float>(), need_cuda(grad_rst, "grad_rst").data_ptr(),
                                      row_pack.data_ptr<float>(), grad_er.data_ptr(), gh.ptr(), en.stream),
               "dglb_gat_fused_bwd_dst");
  return {row_pack, grad_er};
}

std::tuple<Tensor, Tensor> gat_bwd_src(const Tensor& indptr, const Tensor& indices, const OptTensor& eids, int64_t n_dst,
                                       const Tensor& ft, const Tensor& el, const Tensor& row_pack, const Tensor& grad_rst,
                                       double slope, double dropout_p, int64_t seed, bool hub_segments, const OptTensor& hr,
                                       const OptTensor& hs, const OptTensor& hh, at::IntArrayRef hm) {
  const int64_t n_src = indptr.numel() - 1, nnz = indices.numel(), H = ft.size(1), F = ft.size(2);
  Tensor grad_ft = at::empty_like(ft), grad_el = at::empty({n_src, H}, ft.options());
  if (n_src == 0) return {grad_ft, grad_el};
  GatHub gh(ft, H, F, hub_segments, hr, hs, hh, hm);
  Entered en(ft);
  check_status(dglb_gat_fused_bwd_src(DGLB_F32, n_src, n_dst, nnz, H, F, (float)slope, (float)dropout_p, (uint64_t)seed,
                                      i32(indptr, "indptr"), i32(indices, "indices"), i32_opt(eids, "eids"),
                                      need_cuda(ft, "ft").data_ptr(), need_cuda(el, "el").data_ptr(),
                                      need_cuda(row_pack, "row_pack").data_ptr<float>(), need_cuda(grad_rst, "grad_rst").data_ptr(),
                                      grad_ft.data_ptr(), grad_el.data_ptr(), gh.ptr(), en.stream),
               "dglb_gat_fused_bwd_src");
  return {grad_ft, grad_el};
}

// ------------------------------------------------------------------ batched small graphs
Tensor gcn_msg_sum_fwd(const Tensor& indptr, const Tensor& indices, const OptTensor& eids, const Tensor& x, const Tensor& w,
                       const Tensor& c_src, const Tensor& c_dst) {
  need_cuda(x, "x"); need_cuda(w, "w"); need_cuda(c_src, "c_src"); need_cuda(c_dst, "c_dst");
  TORCH_CHECK(x.scalar_type() == at::kFloat && w.scalar_type() == at::kFloat && c_src.scalar_type() == at::kFloat &&
                  c_dst.scalar_type() == at::kFloat, "dgl-b200 kernels compute in float32");
  const int64_t n_dst = indptr.numel() - 1, nnz = indices.numel(), n_src = x.size(0);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && w.size(0) == nnz && w.size(1) == x.size(1),
              "gcn_msg_sum: expect x (n_src, D) and w (n_edges, D); got ", x.sizes(), " and ", w.sizes());
  TORCH_CHECK(c_src.numel() == n_src && c_dst.numel() == n_dst, "gcn_msg_sum: norm vectors must have one entry per node");
  Tensor out = at::empty({n_dst, x.size(1)}, x.options());
  if (out.numel() == 0) return out;
  Entered en(x);
  check_status(dglb_gcn_msg_sum_fwd(n_dst, n_src, nnz, x.size(1), i32(indptr, "indptr"), i32(indices, "indices"),
                                    i32_opt(eids, "eids"), x.data_ptr<float>(), w.data_ptr<float>(), c_src.data_ptr<float>(),
                                    c_dst.data_ptr<float>(), out.data_ptr<float>(), en.stream), "dglb_gcn_msg_sum_fwd");
  return out;
}

// CSR over the source nodes; zero_grad_w: the graph has edge slots that no CSR row covers (fixed-size padded batch)
std::tuple<Tensor, Tensor> gcn_msg_sum_bwd(const Tensor& indptr_csr, const Tensor& indices_csr, const OptTensor& eids_csr,
                                           const Tensor& x, const Tensor& w, const Tensor& c_src, const Tensor& c_dst,
                                           const Tensor& grad_out, bool zero_grad_w) {
  need_cuda(x, "x"); need_cuda(w, "w"); need_cuda(c_src, "c_src"); need_cuda(c_dst, "c_dst"); need_cuda(grad_out, "grad_out");
  TORCH_CHECK(grad_out.scalar_type() == at::kFloat && x.scalar_type() == at::kFloat && w.scalar_type() == at::kFloat,
              "dgl-b200 kernels compute in float32");
  const int64_t n_src = indptr_csr.numel() - 1, nnz = indices_csr.numel(), n_dst = grad_out.size(0);
  TORCH_CHECK(x.size(0) == n_src && w.size(0) == nnz && grad_out.size(1) == x.size(1) && c_dst.numel() == n_dst,
              "gcn_msg_sum_bwd: shape mismatch");
  Tensor gx = at::empty_like(x), gw = zero_grad_w ? at::zeros_like(w) : at::empty_like(w);
  if (gx.numel() == 0) return {gx, gw};
  Entered en(x);
  check_status(dglb_gcn_msg_sum_bwd(n_src, n_dst, nnz, x.size(1), i32(indptr_csr, "indptr"), i32(indices_csr, "indices"),
                                    i32_opt(eids_csr, "eids"), x.data_ptr<float>(), w.data_ptr<float>(),
                                    c_src.data_ptr<float>(), c_dst.data_ptr<float>(), grad_out.data_ptr<float>(),
                                    gx.data_ptr<float>(), gw.data_ptr<float>(), en.stream), "dglb_gcn_msg_sum_bwd");
  return {gx, gw};
}

// x (n, K) int64 codes; table (R, D); offsets: K + 1 table starts (host ints)
Tensor cat_embed_sum_fwd(const Tensor& x, const Tensor& table, at::IntArrayRef offsets) {
  need_cuda(x, "x"); need_cuda(table, "table");
  TORCH_CHECK(x.scalar_type() == at::kLong && x.dim() == 2 && table.scalar_type() == at::kFloat && table.dim() == 2,
              "cat_embed_sum: expect int64 codes (n, K) and a float32 table (R, D)");
  TORCH_CHECK((int64_t)offsets.size() == x.size(1) + 1 && offsets.back() == table.size(0), "cat_embed_sum: offsets must list K + 1 table starts ending at R");
  std::vector<int32_t> off(offsets.begin(), offsets.end());
  Tensor out = at::empty({x.size(0), table.size(1)}, table.options());
  if (out.numel() == 0) return out;
  Entered en(table);
  check_status(dglb_cat_embed_sum_fwd(x.size(0), x.size(1), table.size(1), x.data_ptr<int64_t>(), off.data(),
                                      table.data_ptr<float>(), out.data_ptr<float>(), en.stream), "dglb_cat_embed_sum_fwd");
  return out;
}

Tensor cat_embed_sum_bwd(const Tensor& x, const Tensor& grad_out, at::IntArrayRef offsets) {
  need_cuda(x, "x"); need_cuda(grad_out, "grad_out");
  TORCH_CHECK(x.scalar_type() == at::kLong && x.dim() == 2 && grad_out.scalar_type() == at::kFloat && grad_out.dim() == 2 &&
                  grad_out.size(0) == x.size(0) && (int64_t)offsets.size() == x.size(1) + 1, "cat_embed_sum_bwd: shape mismatch");
  std::vector<int32_t> off(offsets.begin(), offsets.end());
  Tensor gt = at::empty({offsets.back(), grad_out.size(1)}, grad_out.options());
  if (gt.numel() == 0) return gt;
  const size_t ws_bytes = dglb_cat_embed_sum_bwd_workspace_bytes(x.size(0), offsets.back(), grad_out.size(1));
  Tensor ws = at::empty({(int64_t)std::max<size_t>(ws_bytes, 4)}, grad_out.options().dtype(at::kByte));
  Entered en(grad_out);
  check_status(dglb_cat_embed_sum_bwd(x.size(0), x.size(1), grad_out.size(1), x.data_ptr<int64_t>(), off.data(),
                                      grad_out.data_ptr<float>(), gt.data_ptr<float>(), ws.data_ptr(), ws_bytes, en.stream),
               "dglb_cat_embed_sum_bwd");
  return gt;
}

int32_t* i32_mut(const OptTensor& t, const char* name, int64_t min_numel) {
  if (!t.has_value()) return nullptr;
  i32(*t, name);
  TORCH_CHECK(t->numel() >= min_numel, name, " holds ", t->numel(), " entries, need ", min_numel);
  return t->data_ptr<int32_t>();
}

// Fills the caller's fixed-size batch buffers IN PLACE (two launches, no allocation, no sync): capturable in a CUDA graph.
// store  = [node_ptr, edge_ptr, u_src, u_dst, csc_indptr, csc_indices, csc_eids?, csr_indptr, csr_indices, csr_eids?]
// batch  = [out_node_ptr, out_edge_ptr, status, src, dst, csc_indptr, csc_indices, csc_eids, csr_indptr, csr_indices,
//           csr_eids, node_graph, node_map, edge_map]   (entries from `src` on may be None)
void batch_build(const Tensor& graph_ids, const c10::List<OptTensor>& store_l, const c10::List<OptTensor>& batch_l,
                 int64_t n_nodes_pad, int64_t n_edges_pad) {
  std::vector<OptTensor> store, batch;
  for (size_t i = 0; i < store_l.size(); ++i) store.push_back(store_l.get(i));
  for (size_t i = 0; i < batch_l.size(); ++i) batch.push_back(batch_l.get(i));
  TORCH_CHECK(store.size() == 10 && batch.size() == 14, "batch_build: expect 10 store tensors and 14 batch tensors");
  const int64_t n_sel = graph_ids.numel();
  dglb_batch_io_t io{};
  io.n_sel = (int32_t)n_sel; io.n_nodes_pad = (int32_t)n_nodes_pad; io.n_edges_pad = (int32_t)n_edges_pad;
  io.graph_ids = i32(graph_ids, "graph_ids");
  TORCH_CHECK(store[0].has_value() && store[1].has_value() && batch[0].has_value() && batch[1].has_value(),
              "batch_build: node_ptr / edge_ptr / out_node_ptr / out_edge_ptr are required");
  io.node_ptr = i32(*store[0], "node_ptr"); io.edge_ptr = i32(*store[1], "edge_ptr");
  io.u_src = i32_opt(store[2], "u_src"); io.u_dst = i32_opt(store[3], "u_dst");
  io.u_csc_indptr = i32_opt(store[4], "u_csc_indptr"); io.u_csc_indices = i32_opt(store[5], "u_csc_indices");
  io.u_csc_eids = i32_opt(store[6], "u_csc_eids");
  io.u_csr_indptr = i32_opt(store[7], "u_csr_indptr"); io.u_csr_indices = i32_opt(store[8], "u_csr_indices");
  io.u_csr_eids = i32_opt(store[9], "u_csr_eids");
  int32_t* out_node_ptr = i32_mut(batch[0], "out_node_ptr", n_sel + 2);
  int32_t* out_edge_ptr = i32_mut(batch[1], "out_edge_ptr", n_sel + 2);
  int32_t* status = i32_mut(batch[2], "status", 1);
  io.out_node_ptr = out_node_ptr; io.out_edge_ptr = out_edge_ptr;
  io.src = i32_mut(batch[3], "src", n_edges_pad); io.dst = i32_mut(batch[4], "dst", n_edges_pad);
  io.csc_indptr = i32_mut(batch[5], "csc_indptr", n_nodes_pad + 1);
  io.csc_indices = i32_mut(batch[6], "csc_indices", n_edges_pad); io.csc_eids = i32_mut(batch[7], "csc_eids", n_edges_pad);
  io.csr_indptr = i32_mut(batch[8], "csr_indptr", n_nodes_pad + 1);
  io.csr_indices = i32_mut(batch[9], "csr_indices", n_edges_pad); io.csr_eids = i32_mut(batch[10], "csr_eids", n_edges_pad);
  io.node_graph = i32_mut(batch[11], "node_graph", n_nodes_pad);
  io.node_map = i32_mut(batch[12], "node_map", n_nodes_pad);
  io.edge_map = i32_mut(batch[13], "edge_map", n_edges_pad);
  Entered en(graph_ids);
  check_status(dglb_batch_offsets(n_sel, io.graph_ids, io.node_ptr, io.edge_ptr, out_node_ptr, out_edge_ptr, n_nodes_pad,
                                  n_edges_pad, status, en.stream), "dglb_batch_offsets");
  check_status(dglb_batch_gather(&io, en.stream), "dglb_batch_gather");
}

// ------------------------------------------------------------------ halo exchange
// dst[idx] = src[idx] row-wise, in place (src may be a peer GPU's symmetric-memory view); runs on the CURRENT stream of
// dst's device
void copy_rows_indexed(const Tensor& src, Tensor dst, const Tensor& idx) {
  need_cuda(src, "src"); need_cuda(dst, "dst");
  TORCH_CHECK(src.scalar_type() == dst.scalar_type() && src.dim() >= 1 && dst.dim() == src.dim(),
              "copy_rows_indexed: src and dst must have the same dtype and rank");
  const int64_t n = idx.numel();
  if (n == 0) return;
  const int64_t row_bytes = (dst.numel() / std::max<int64_t>(dst.size(0), 1)) * dst.element_size();
  TORCH_CHECK(src.size(0) > 0 && (src.numel() / src.size(0)) * src.element_size() == row_bytes,
              "copy_rows_indexed: row sizes differ");
  Entered en(dst);
  check_status(dglb_copy_rows_indexed(n, i32(idx, "idx"), row_bytes, src.data_ptr(), dst.data_ptr(), en.stream),
               "dglb_copy_rows_indexed");
}

int64_t abi_version() { return dglb_abi_version(); }
int64_t default_hub_threshold(int64_t which, int64_t arg) {
  return which == 0 ? dglb_default_hub_threshold(arg) : (which == 1 ? dglb_default_row_hub_threshold(arg)
                                                                    : dglb_default_softmax_hub_threshold(arg));
}

}  // namespace

TORCH_LIBRARY(dglb200, m) {
  m.def("abi_version() -> int", &abi_version);
  m.def("default_hub_threshold(int which, int arg) -> int", &default_hub_threshold);
  m.def("coo_to_csr(Tensor row, Tensor col, int n_rows) -> (Tensor, Tensor, Tensor)", &coo_to_csr);
  m.def("csr_degrees(Tensor indptr) -> Tensor", &csr_degrees);
  m.def("is_identity_perm(Tensor data) -> Tensor", &is_identity_perm);
  m.def("find_hub_rows(Tensor indptr, int threshold, int cap) -> (Tensor, Tensor)", &find_hub_rows);
  m.def("edge_stage_plan(Tensor eids, int log2_bucket) -> (Tensor, Tensor)", &edge_stage_plan);
  m.def("edge_stage(Tensor stage_pos, Tensor t, bool to_staged) -> Tensor", &edge_stage);
  m.def("gspmm(Tensor indptr, Tensor indices, Tensor? eids, int n_cols, int op, int reduce, Tensor? u, Tensor? e, "
        "int[] out_feat, int[] lhs_shape, int[] rhs_shape, Tensor? row_scale, Tensor? out, int flags, Tensor? hub_rows, "
        "Tensor? hub_seg_ptr, Tensor? hub_seg_hub, Tensor? hub_light, Tensor? hub_order, int[] hub_meta) -> (Tensor, Tensor, Tensor)", &gspmm);
  m.def("gsddmm_csr(Tensor indptr, Tensor indices, Tensor? eids, int n_src, int op, int lhs_target, int rhs_target, "
        "Tensor? lhs, Tensor? rhs, int[] out_feat, int[] lhs_shape, int[] rhs_shape, Tensor? hub_rows, Tensor? hub_seg_ptr, "
        "Tensor? hub_seg_hub, Tensor? hub_light, Tensor? hub_order, int[] hub_meta) -> Tensor", &gsddmm_csr);
  m.def("gsddmm_coo(Tensor src, Tensor dst, int n_src, int n_dst, int op, int lhs_target, int rhs_target, Tensor? lhs, "
        "Tensor? rhs, int[] out_feat, int[] lhs_shape, int[] rhs_shape) -> Tensor", &gsddmm_coo);
  m.def("edge_softmax_fwd(Tensor indptr, Tensor? eids, Tensor logits, int heads, Tensor? hub_rows, Tensor? hub_seg_ptr, "
        "Tensor? hub_seg_hub, int[] hub_meta) -> Tensor", &edge_softmax_fwd);
  m.def("edge_softmax_bwd(Tensor indptr, Tensor? eids, Tensor out, Tensor grad_out, int heads, Tensor? hub_rows, "
        "Tensor? hub_seg_ptr, Tensor? hub_seg_hub, int[] hub_meta) -> Tensor", &edge_softmax_bwd);
  m.def("gat_fwd(Tensor indptr, Tensor indices, Tensor? eids, Tensor ft, Tensor el, Tensor er, float slope, float dropout_p, "
        "int seed, bool want_scores, bool hub_segments, Tensor? hub_rows, Tensor? hub_seg_ptr, Tensor? hub_seg_hub, "
        "int[] hub_meta) -> (Tensor, Tensor, Tensor, Tensor)", &gat_fwd);
  m.def("gat_bwd_dst(Tensor indptr, Tensor indices, Tensor? eids, Tensor ft, Tensor el, Tensor er, Tensor row_max, "
        "Tensor row_sum, Tensor grad_rst, float slope, float dropout_p, int seed, bool hub_segments, Tensor? hub_rows, "
        "Tensor? hub_seg_ptr, Tensor? hub_seg_hub, int[] hub_meta) -> (Tensor, Tensor)", &gat_bwd_dst);
  m.def("gcn_msg_sum_fwd(Tensor indptr, Tensor indices, Tensor? eids, Tensor x, Tensor w, Tensor c_src, Tensor c_dst) -> Tensor",
        &gcn_msg_sum_fwd);
  m.def("gcn_msg_sum_bwd(Tensor indptr_csr, Tensor indices_csr, Tensor? eids_csr, Tensor x, Tensor w, Tensor c_src, "
        "Tensor c_dst, Tensor grad_out, bool zero_grad_w) -> (Tensor, Tensor)", &gcn_msg_sum_bwd);
  m.def("cat_embed_sum_fwd(Tensor x, Tensor table, int[] offsets) -> Tensor", &cat_embed_sum_fwd);
  m.def("cat_embed_sum_bwd(Tensor x, Tensor grad_out, int[] offsets) -> Tensor", &cat_embed_sum_bwd);
  m.def("batch_build(Tensor graph_ids, Tensor?[] store, Tensor?[] batch, int n_nodes_pad, int n_edges_pad) -> ()",
        &batch_build);
  m.def("copy_rows_indexed(Tensor src, Tensor(a!) dst, Tensor idx) -> ()", &copy_rows_indexed);
  m.def("gat_bwd_src(Tensor indptr, Tensor indices, Tensor? eids, int n_dst, Tensor ft, Tensor el, Tensor row_pack, "
        "Tensor grad_rst, float slope, float dropout_p, int seed, bool hub_segments, Tensor? hub_rows, Tensor? hub_seg_ptr, "
        "Tensor? hub_seg_hub, int[] hub_meta) -> (Tensor, Tensor)", &gat_bwd_src);
}
