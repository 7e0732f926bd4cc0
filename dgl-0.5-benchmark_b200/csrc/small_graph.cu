// small_graph.cu -- the batched small-graph pipeline (BASELINE config 5): ogbg-molhiv-shaped
// batches of ~25-node graphs, N ~ 1.6 K / E ~ 3.5 K at batch 64, where the reference loop
// (end_to_end/full_graph/graph_classification/main_dgl_molhiv_gcn.py:95-115) is bound by per-batch graph construction and
// by dozens of tiny launches per layer, not by bytes.
//
//  1. Fused GCN message + sum (main_dgl_molhiv_gcn.py:46,50-52):
//         out[v,:] = sum_{e=(u->v)} (c[u] * c[v]) * relu(x[u,:] + w[e,:])
//     Upstream runs the UDF `message` with torch ops on (E, D) tensors (two index_selects of the norm, one of x, add, relu,
//     mul) and then a copy_e-sum SpMM; its backward is the mirror image plus an index_add with atomics.  Here: one
//     forward kernel over the CSC (a thread per (row, 4-float column), edges of a row in CSC order, products and sums
//     rounded exactly like the composite: no FMA contraction -> bit-identical to it) and one backward kernel over the CSR
//     that produces grad_x (sum over a node's out-edges in CSR order, deterministic) and grad_w (one row per edge) in a
//     single pass, recomputing the ReLU mask from x + w instead of saving the (E, D) message.
//
//  1b. Sum of categorical embeddings (ogb.graphproppred.mol_encoder AtomEncoder / BondEncoder, main_dgl_molhiv_gcn.py:28,72:
//     out[i,:] = sum_k table_k[x[i,k],:], 9 atom / 3 bond columns).  torch runs one embedding lookup + one add per
//     column forward and, backward, a radix sort + segmented reduction per column: 15 encoders x ~40 launches were
//     half of a captured training iteration (profiles/r02_molhiv_graph_profile.txt).  Here the tables are one
//     concatenated matrix: one forward kernel (same order of additions as the column loop: bit-identical) and one
//     DETERMINISTIC backward kernel (a CTA per table row scans the rows of x in order; the tables have a few hundred
//     rows in total, so that is O(N) work per CTA for a batch of small graphs).
//
//  2. Device-side dgl.batch (upstream python/dgl/batch.py::batch + the COO->CSC/CSR conversions it triggers,
//     main_dgl_molhiv_gcn.py:101,163): the dataset lives on the device as ONE union graph (all member graphs side by
//     side) with its CSC / CSR built once.  Because the node ranges of member graphs are disjoint and increasing, the CSC of
//     any batch is the concatenation of the members' CSC slices with shifted ids -- no sort.  batch_offsets scans the
//     selected graphs' node / edge counts (one CTA); batch_gather writes COO, CSC, CSR, node -> member-graph ids and the
//     node / edge gather maps for the batch into caller-owned buffers of FIXED (padded) size, so a whole training step
//     -- batch construction included -- replays as one CUDA graph.  Padding nodes are isolated and belong to an extra
//     member graph; padding edge slots lie beyond indptr[n_nodes_pad].
#include <cub/block/block_scan.cuh>

#include "kernels.cuh"

namespace dglb {

// ------------------------------------------------------------------------------------------------ fused GCN message
template <int VEC>
__global__ void __launch_bounds__(kBlockThreads)
gcn_msg_sum_fwd_kernel(int64_t n_rows, int D, int ncol, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const int32_t* __restrict__ eids, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ c_src, const float* __restrict__ c_dst, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int64_t row = idx / ncol;
  if (row >= n_rows) return;
  const int col = (int)(idx - row * ncol) * VEC;
  const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
  const float cv = __ldg(c_dst + row);
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  for (int j = s; j < e; ++j) {
    const int u = __ldg(indices + j);
    const int64_t eid = eids ? (int64_t)__ldg(eids + j) : (int64_t)j;
    const float nrm = __fmul_rn(__ldg(c_src + u), cv);
    const FVec<VEC> xv = ldg_vec<VEC>(x + (int64_t)u * D + col);
    const FVec<VEC> wv = ldg_vec<VEC>(w + eid * D + col);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float t = __fadd_rn(xv.v[v], wv.v[v]);
      const float r = t < 0.f ? 0.f : t;                       // relu; NaN propagates like torch's
      acc[v] = __fadd_rn(acc[v], __fmul_rn(nrm, r));
    }
  }
  FVec<VEC> o;
#pragma unroll
  for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
  st_vec<VEC>(out + row * D + col, o);
}

// CSR over the source nodes: row = u, indices = destination ids
template <int VEC>
__global__ void __launch_bounds__(kBlockThreads)
gcn_msg_sum_bwd_kernel(int64_t n_rows, int D, int ncol, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const int32_t* __restrict__ eids, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ c_src, const float* __restrict__ c_dst, const float* __restrict__ gout,
                       float* __restrict__ gx, float* __restrict__ gw) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int64_t row = idx / ncol;
  if (row >= n_rows) return;
  const int col = (int)(idx - row * ncol) * VEC;
  const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
  const float cu = __ldg(c_src + row);
  const FVec<VEC> xu = ldg_vec<VEC>(x + row * D + col);
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  for (int j = s; j < e; ++j) {
    const int dv = __ldg(indices + j);
    const int64_t eid = eids ? (int64_t)__ldg(eids + j) : (int64_t)j;
    const float nrm = __fmul_rn(cu, __ldg(c_dst + dv));
    const FVec<VEC> g = ldg_vec<VEC>(gout + (int64_t)dv * D + col);
    const FVec<VEC> wv = ldg_vec<VEC>(w + eid * D + col);
    FVec<VEC> d;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float t = __fadd_rn(xu.v[v], wv.v[v]);
      d.v[v] = t > 0.f ? __fmul_rn(g.v[v], nrm) : 0.f;         // d relu = grad * (out > 0), out = relu(x + w)
      acc[v] = __fadd_rn(acc[v], d.v[v]);
    }
    st_vec<VEC>(gw + eid * D + col, d);
  }
  FVec<VEC> o;
#pragma unroll
  for (int v = 0; v < VEC; ++v) o.v[v] = acc[v];
  st_vec<VEC>(gx + row * D + col, o);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gcn_msg_sum_fwd(int64_t n_dst, int64_t D, const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                    const float* x, const float* w, const float* c_src, const float* c_dst, float* out,
                    cudaStream_t stream) {
  if (n_dst == 0 || D == 0) return DGLB_OK;
  const bool v4 = D % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(out);
  const int ncol = (int)(v4 ? D / 4 : D);
  const int64_t blocks = (n_dst * ncol + kBlockThreads - 1) / kBlockThreads;
  if (v4)
    gcn_msg_sum_fwd_kernel<4><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(n_dst, (int)D, ncol, indptr, indices, eids, x, w,
                                                                                 c_src, c_dst, out);
  else
    gcn_msg_sum_fwd_kernel<1><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(n_dst, (int)D, ncol, indptr, indices, eids, x, w,
                                                                                 c_src, c_dst, out);
  DGLB_LAUNCH_CHECK("gcn_msg_sum_fwd_kernel");
  return DGLB_OK;
}

int gcn_msg_sum_bwd(int64_t n_src, int64_t D, const int32_t* indptr, const int32_t* indices, const int32_t* eids,
                    const float* x, const float* w, const float* c_src, const float* c_dst, const float* gout, float* gx,
                    float* gw, cudaStream_t stream) {
  if (n_src == 0 || D == 0) return DGLB_OK;
  const bool v4 = D % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(gout) && aligned16(gx) && aligned16(gw);
  const int ncol = (int)(v4 ? D / 4 : D);
  const int64_t blocks = (n_src * ncol + kBlockThreads - 1) / kBlockThreads;
  if (v4)
    gcn_msg_sum_bwd_kernel<4><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(n_src, (int)D, ncol, indptr, indices, eids, x, w,
                                                                                 c_src, c_dst, gout, gx, gw);
  else
    gcn_msg_sum_bwd_kernel<1><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(n_src, (int)D, ncol, indptr, indices, eids, x, w,
                                                                                 c_src, c_dst, gout, gx, gw);
  DGLB_LAUNCH_CHECK("gcn_msg_sum_bwd_kernel");
  return DGLB_OK;
}

// ------------------------------------------------------------------------------------------------ categorical embeddings
struct CatEmbedParams {
  const int64_t* __restrict__ x;   // (n_rows, K) categorical codes (int64, as OGB stores them)
  const float* __restrict__ T;     // (R, D) concatenated tables
  const float* __restrict__ g;     // bwd: (n_rows, D)
  float* __restrict__ out;         // fwd: (n_rows, D); bwd: (R, D)
  int64_t n_rows;
  int K, D, ncol, R;
  int off[DGLB_MAX_CAT_COLUMNS + 1];   // first row of every column's table; off[K] = R
};

template <int VEC>
__global__ void __launch_bounds__(kBlockThreads) cat_embed_sum_fwd_kernel(const CatEmbedParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int64_t row = idx / p.ncol;
  if (row >= p.n_rows) return;
  const int col = (int)(idx - row * p.ncol) * VEC;
  float acc[VEC];
  for (int k = 0; k < p.K; ++k) {
    const int64_t r = (int64_t)p.off[k] + __ldg(p.x + row * p.K + k);
    const FVec<VEC> t = ldg_vec<VEC>(p.T + r * p.D + col);
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = k == 0 ? t.v[v] : __fadd_rn(acc[v], t.v[v]);
  }
  FVec<VEC> o;
#pragma unroll
  for (int v = 0; v < VEC; ++v) o.v[v] = p.K > 0 ? acc[v] : 0.f;
  st_vec<VEC>(p.out + row * p.D + col, o);
}

// Grid (table row r, chunk s): the CTA sums grad_out over the rows of x in chunk s = [s * kCatChunk, (s+1) * kCatChunk)
// whose code in column k(r) is r - off[k].  Inside the chunk warp w owns a contiguous eighth: it compacts the matching
// rows into its shared-memory list (ballot + popcount, order preserved, coalesced scan), then sums them with several rows
// in flight, each lane holding its columns in registers; the 8 per-warp sums are combined in warp order and written to
// partial[s][r][:].  A second kernel adds the chunks in order: deterministic, no atomics.  (Two earlier versions -- a
// per-row "load code, compare, branch" loop, then one CTA per table row -- took ~150 us per call: the tables have few
// rows, binary columns match half of x, and a single CTA was summing ~1 800 rows.)
constexpr int kCatChunk = 512;
constexpr int kCatWarps = kBlockThreads / 32;
constexpr int kCatMaxCols = 8;   // float4 columns per lane: D <= 32 * 4 * 8 = 1024

template <bool V4>
__global__ void __launch_bounds__(kBlockThreads) cat_embed_sum_bwd_kernel(const CatEmbedParams p, float* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char cat_smem[];
  constexpr int kPer = kCatChunk / kCatWarps;
  int* lists = reinterpret_cast<int*>(cat_smem);                                  // [kCatWarps][kPer]
  float* sums = reinterpret_cast<float*>(cat_smem + sizeof(int) * kCatChunk);      // [kCatWarps][D]
  const int r = blockIdx.x;
  int k = 0;
  while (k + 1 < p.K && p.off[k + 1] <= r) ++k;
  const int64_t code = r - p.off[k];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* my = lists + warp * kPer;
  const int D = p.D;
  const int nvec = (D + 127) / 128;           // float4 columns per lane (V4: D % 4 == 0, 16-byte aligned rows)
  float acc[kCatMaxCols][4];
#pragma unroll
  for (int c = 0; c < kCatMaxCols; ++c)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[c][v] = 0.f;

  const int64_t lo = (int64_t)blockIdx.y * kCatChunk + (int64_t)warp * kPer;
  const int64_t hi = min(p.n_rows, lo + kPer);
  int n = 0;
  for (int64_t i0 = lo; i0 < hi; i0 += 32) {
    const int64_t i = i0 + lane;
    const bool m = i < hi && __ldg(p.x + i * p.K + k) == code;
    const unsigned b = __ballot_sync(FULL_MASK, m);
    if (m) my[n + __popc(b & ((1u << lane) - 1u))] = (int)(i - lo);
    n += __popc(b);
  }
  __syncwarp();
  if constexpr (V4) {
#pragma unroll 4
    for (int j = 0; j < n; ++j) {          // rows in order; the loads of up to 4 rows are issued before their adds
      const float* g = p.g + (lo + my[j]) * D;
#pragma unroll
      for (int c = 0; c < kCatMaxCols; ++c) {
        const int col = (c * 32 + lane) * 4;
        if (c < nvec && col < D) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(g + col));
          acc[c][0] = __fadd_rn(acc[c][0], t.x); acc[c][1] = __fadd_rn(acc[c][1], t.y);
          acc[c][2] = __fadd_rn(acc[c][2], t.z); acc[c][3] = __fadd_rn(acc[c][3], t.w);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kCatMaxCols; ++c) {
      const int col = (c * 32 + lane) * 4;
      if (c < nvec && col < D) {
#pragma unroll
        for (int v = 0; v < 4; ++v) sums[warp * D + col + v] = acc[c][v];
      }
    }
  } else {
    for (int j = 0; j < n; ++j) {
      const float* g = p.g + (lo + my[j]) * D;
#pragma unroll
      for (int c = 0; c < kCatMaxCols * 4; ++c) {
        const int col = c * 32 + lane;
        if (col < D) acc[c >> 2][c & 3] = __fadd_rn(acc[c >> 2][c & 3], __ldg(g + col));
      }
    }
#pragma unroll
    for (int c = 0; c < kCatMaxCols * 4; ++c) {
      const int col = c * 32 + lane;
      if (col < D) sums[warp * D + col] = acc[c >> 2][c & 3];
    }
  }
  __syncthreads();
  float* dst = partial + ((int64_t)blockIdx.y * p.R + r) * D;
  for (int col = threadIdx.x; col < D; col += kBlockThreads) {
    float a = sums[col];
    for (int w = 1; w < kCatWarps; ++w) a = __fadd_rn(a, sums[w * D + col]);
    dst[col] = a;
  }
}

// out[r,:] = partial[0][r,:] + partial[1][r,:] + ... (chunk order)
__global__ void __launch_bounds__(kBlockThreads)
cat_embed_combine_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t rd, int n_chunks) {
  const int64_t i = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (i >= rd) return;
  float a = partial[i];
  for (int s = 1; s < n_chunks; ++s) a = __fadd_rn(a, partial[(int64_t)s * rd + i]);
  out[i] = a;
}

size_t cat_embed_bwd_workspace_bytes(int64_t n_rows, int64_t n_table_rows, int64_t D) {
  const int64_t chunks = (n_rows + kCatChunk - 1) / kCatChunk;
  return chunks > 1 ? (size_t)chunks * (size_t)n_table_rows * (size_t)D * sizeof(float) : 0;
}

int cat_embed_sum(bool bwd, int64_t n_rows, int64_t K, int64_t D, const int64_t* x, const int32_t* offsets_host,
                  const float* T, const float* g, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (K > DGLB_MAX_CAT_COLUMNS) {
    set_error("cat_embed_sum: at most %d categorical columns (got %lld)", DGLB_MAX_CAT_COLUMNS, (long long)K);
    return DGLB_E_UNSUPPORTED;
  }
  CatEmbedParams p;
  p.x = x; p.T = T; p.g = g; p.out = out; p.n_rows = n_rows; p.K = (int)K; p.D = (int)D;
  for (int k = 0; k <= K; ++k) p.off[k] = offsets_host[k];
  p.R = offsets_host[K];
  const bool v4 = D % 4 == 0 && aligned16(T) && aligned16(out) && (!bwd || aligned16(g));
  p.ncol = (int)(v4 ? D / 4 : D);
  if (!bwd) {
    if (n_rows == 0 || D == 0) return DGLB_OK;
    const int64_t blocks = (n_rows * p.ncol + kBlockThreads - 1) / kBlockThreads;
    if (v4) cat_embed_sum_fwd_kernel<4><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    else cat_embed_sum_fwd_kernel<1><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("cat_embed_sum_fwd_kernel");
  } else {
    if (p.R == 0 || D == 0) return DGLB_OK;
    if (D > 32 * 4 * kCatMaxCols) {
      set_error("cat_embed_sum_bwd: feature width %lld exceeds %d", (long long)D, 32 * 4 * kCatMaxCols);
      return DGLB_E_UNSUPPORTED;
    }
    const int64_t chunks = n_rows > 0 ? (n_rows + kCatChunk - 1) / kCatChunk : 1;
    const size_t need = cat_embed_bwd_workspace_bytes(n_rows, p.R, D);
    if (need > 0 && (!workspace || workspace_bytes < need)) {
      set_error("cat_embed_sum_bwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
      return DGLB_E_WORKSPACE;
    }
    if (chunks > 65535) {
      set_error("cat_embed_sum_bwd: %lld rows exceed the supported batch size", (long long)n_rows);
      return DGLB_E_UNSUPPORTED;
    }
    float* partial = chunks > 1 ? static_cast<float*>(workspace) : out;   // a single chunk writes the result directly
    const size_t smem = sizeof(int) * kCatChunk + (size_t)kCatWarps * (size_t)D * sizeof(float);
    const dim3 grid((unsigned)p.R, (unsigned)chunks);
    if (v4) cat_embed_sum_bwd_kernel<true><<<grid, kBlockThreads, smem, stream>>>(p, partial);
    else cat_embed_sum_bwd_kernel<false><<<grid, kBlockThreads, smem, stream>>>(p, partial);
    if (chunks > 1) {
      const int64_t rd = (int64_t)p.R * D;
      cat_embed_combine_kernel<<<(unsigned)((rd + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, stream>>>(
          partial, out, rd, (int)chunks);
    }
    DGLB_LAUNCH_CHECK("cat_embed_sum_bwd_kernel");
  }
  return DGLB_OK;
}

// ------------------------------------------------------------------------------------------------ device-side dgl.batch
constexpr int kScanThreads = 256;

// out_node_ptr / out_edge_ptr [n_sel + 2]: exclusive prefix sums of the selected graphs' node / edge counts, then the
// padded totals (entry n_sel + 1), so that [n_sel, n_sel + 1) is the padding member graph.  status[0] |= 1 when the batch
// does not fit the padded sizes (the gather kernel then clips; the caller checks the flag or sizes the pads from the
// host-side counts it already has).
__global__ void __launch_bounds__(kScanThreads)
batch_offsets_kernel(int n_sel, const int32_t* __restrict__ graph_ids, const int32_t* __restrict__ node_ptr,
                     const int32_t* __restrict__ edge_ptr, int32_t* __restrict__ out_node_ptr,
                     int32_t* __restrict__ out_edge_ptr, int32_t n_nodes_pad, int32_t n_edges_pad, int32_t* status) {
  using Scan = cub::BlockScan<int2, kScanThreads>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int2 carry_s;
  struct Add {
    __device__ int2 operator()(const int2& a, const int2& b) const { return make_int2(a.x + b.x, a.y + b.y); }
  };
  int2 carry = make_int2(0, 0);
  for (int base = 0; base < n_sel; base += kScanThreads) {
    const int i = base + threadIdx.x;
    int2 c = make_int2(0, 0);
    if (i < n_sel) {
      const int g = __ldg(graph_ids + i);
      c.x = __ldg(node_ptr + g + 1) - __ldg(node_ptr + g);
      c.y = __ldg(edge_ptr + g + 1) - __ldg(edge_ptr + g);
    }
    int2 excl, total;
    Scan(tmp).ExclusiveScan(c, excl, make_int2(0, 0), Add(), total);
    if (i < n_sel) {
      out_node_ptr[i] = carry.x + excl.x;
      out_edge_ptr[i] = carry.y + excl.y;
    }
    if (threadIdx.x == 0) carry_s = make_int2(carry.x + total.x, carry.y + total.y);
    __syncthreads();
    carry = carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_node_ptr[n_sel] = carry.x;
    out_edge_ptr[n_sel] = carry.y;
    out_node_ptr[n_sel + 1] = n_nodes_pad;
    out_edge_ptr[n_sel + 1] = n_edges_pad;
    if (status) status[0] = (carry.x > n_nodes_pad || carry.y > n_edges_pad) ? 1 : 0;
  }
}

// member graph of batch position `i`: last b with ptr[b] <= i  (ptr has n + 1 entries, ptr[0] = 0; i < ptr[n])
__device__ __forceinline__ int member_of(const int32_t* __restrict__ ptr, int n, int i) {
  int lo = 0, hi = n;   // invariant: ptr[lo] <= i < ptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBlockThreads)
batch_gather_kernel(const dglb_batch_io_t io) {
  const int64_t tid = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int B = io.n_sel;
  const int N = min(__ldg(io.out_node_ptr + B), io.n_nodes_pad);   // real nodes / edges of this batch (clipped)
  const int E = min(__ldg(io.out_edge_ptr + B), io.n_edges_pad);
  // ---- nodes (one extra thread writes the closing indptr entries)
  if (tid <= io.n_nodes_pad) {
    const int i = (int)tid;
    if (i < N) {
      const int b = member_of(io.out_node_ptr, B, i);
      const int g = __ldg(io.graph_ids + b);
      const int n0 = __ldg(io.node_ptr + g), e0 = __ldg(io.edge_ptr + g);
      const int o_n0 = __ldg(io.out_node_ptr + b), o_e0 = __ldg(io.out_edge_ptr + b);
      const int sn = n0 + (i - o_n0);                               // node of the union graph
      if (io.csc_indptr) io.csc_indptr[i] = min(o_e0 + (__ldg(io.u_csc_indptr + sn) - e0), E);
      if (io.csr_indptr) io.csr_indptr[i] = min(o_e0 + (__ldg(io.u_csr_indptr + sn) - e0), E);
      if (io.node_graph) io.node_graph[i] = b;
      if (io.node_map) io.node_map[i] = sn;
    } else {
      if (io.csc_indptr) io.csc_indptr[i] = E;
      if (io.csr_indptr) io.csr_indptr[i] = E;
      if (i < io.n_nodes_pad) {
        if (io.node_graph) io.node_graph[i] = B;                    // the padding member graph
        if (io.node_map) io.node_map[i] = 0;                        // any valid node: its features feed isolated rows only
      }
    }
  }
  // ---- edges: position j is at once an edge id (COO, edge_map), a CSC position and a CSR position of the batch
  if (tid < io.n_edges_pad) {
    const int j = (int)tid;
    if (j < E) {
      const int b = member_of(io.out_edge_ptr, B, j);
      const int g = __ldg(io.graph_ids + b);
      const int n0 = __ldg(io.node_ptr + g), e0 = __ldg(io.edge_ptr + g);
      const int dn = __ldg(io.out_node_ptr + b) - n0, de = __ldg(io.out_edge_ptr + b) - e0;
      const int se = j - de;                                        // edge id / CSC position / CSR position in the union graph
      if (io.src) io.src[j] = __ldg(io.u_src + se) + dn;
      if (io.dst) io.dst[j] = __ldg(io.u_dst + se) + dn;
      if (io.edge_map) io.edge_map[j] = se;
      if (io.csc_indices) {
        io.csc_indices[j] = __ldg(io.u_csc_indices + se) + dn;
        io.csc_eids[j] = (io.u_csc_eids ? __ldg(io.u_csc_eids + se) : se) + de;
      }
      if (io.csr_indices) {
        io.csr_indices[j] = __ldg(io.u_csr_indices + se) + dn;
        io.csr_eids[j] = (io.u_csr_eids ? __ldg(io.u_csr_eids + se) : se) + de;
      }
    } else {
      const int pad_node = io.n_nodes_pad - 1;
      if (io.src) io.src[j] = pad_node;
      if (io.dst) io.dst[j] = pad_node;
      if (io.edge_map) io.edge_map[j] = 0;
      if (io.csc_indices) { io.csc_indices[j] = pad_node; io.csc_eids[j] = j; }
      if (io.csr_indices) { io.csr_indices[j] = pad_node; io.csr_eids[j] = j; }
    }
  }
}

int batch_offsets(int64_t n_sel, const int32_t* graph_ids, const int32_t* node_ptr, const int32_t* edge_ptr,
                  int32_t* out_node_ptr, int32_t* out_edge_ptr, int64_t n_nodes_pad, int64_t n_edges_pad, int32_t* status,
                  cudaStream_t stream) {
  batch_offsets_kernel<<<1, kScanThreads, 0, stream>>>((int)n_sel, graph_ids, node_ptr, edge_ptr, out_node_ptr, out_edge_ptr,
                                                       (int32_t)n_nodes_pad, (int32_t)n_edges_pad, status);
  DGLB_LAUNCH_CHECK("batch_offsets_kernel");
  return DGLB_OK;
}

int batch_gather(const dglb_batch_io_t& io, cudaStream_t stream) {
  const int64_t work = (int64_t)(io.n_nodes_pad + 1 > io.n_edges_pad ? io.n_nodes_pad + 1 : io.n_edges_pad);
  const int64_t blocks = (work + kBlockThreads - 1) / kBlockThreads;
  batch_gather_kernel<<<(unsigned)blocks, kBlockThreads, 0, stream>>>(io);
  DGLB_LAUNCH_CHECK("batch_gather_kernel");
  return DGLB_OK;
}

}  // namespace dglb
