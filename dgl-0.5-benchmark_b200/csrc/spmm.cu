// spmm.cu -- generalized SpMM on a CSR-like structure (the CSC of the graph for forward
// aggregation, its CSR for the reverse graph), hand-written for sm_100a.
//
// Replaces upstream DGL v0.6.1 src/array/cuda/spmm.cu(h) (SpMMCsrKernel: one thread per
// (row, feature), scalar loads, no balancing; CusparseCsrmm2 for copy_lhs/sum) behind
// `_CAPI_DGLKernelSpMM`; reached from kernel/dgl-new.py:20 and every update_all() in
// end_to_end/full_graph (e.g. main_dgl_citation_sage.py:75-77).
//
// Design (HBM-bound gather; no tensor cores -- ~0.25 flop/byte):
//   * row-per-group: a power-of-two group of G<=32 lanes owns one destination row; each lane
//     owns CH vector columns (VEC = 4/2/1 floats => LDG.128/64/32) spaced G apart, so a
//     neighbour row is fetched by fully coalesced 16*G-byte requests.  G shrinks with the
//     feature width so narrow features pack several rows into one warp.
//   * the group's lanes fetch G column indices with ONE coalesced load and broadcast them with
//     warp shuffles; U neighbour rows x CH chunks are issued back to back before the first
//     add (U*CH = 8 independent 128-bit loads in flight per lane).
//   * per output element the neighbours are accumulated strictly in CSR order with
//     non-contracted fp32 ops, which reproduces the sequential order of DGL's CPU kernel
//     (SpMMSumCsr) bit for bit, and makes the max/min arg tie-break (first wins) exact.
//   * hub rows (nnz > threshold, listed by dglb_csr_find_hub_rows) are skipped by the row
//     kernel and cut into segments of <= seg_len entries: one CTA per SEGMENT (a 20 000-edge hub
//     is spread over ~30 SMs instead of crawling through one CTA), its groups take contiguous
//     slices, partials meet in shared memory and are combined in slice order, the segment's
//     partial row goes to a workspace, and a small second kernel folds the segments of each hub
//     row in segment order (deterministic, no atomics; ties still resolve to the first CSR entry).
//   * wide rows are processed in feature tiles of G*CH*VEC floats (indices re-read from L1).
//   * anything that does not fit the vector paths (exotic broadcasts, add/sub/div with
//     max/min) goes to a generic thread-per-(row,feature) kernel -- still CUDA, never CPU.
#include "kernels.cuh"

namespace dglb {

// how the edge operand W is indexed: not at all / same width as the output / one value per head
// (column k reads W[e, k / inner]) / ONE scalar per edge (rhs_len == 1: GCN / RGCN edge weights).
enum : int { RMODE_NONE = 0, RMODE_FULL = 1, RMODE_HEAD = 2, RMODE_SCALAR = 3 };

struct SpmmParams {
  const int32_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ eids;  // null => edge id == CSR position
  const float* __restrict__ X;       // lhs rows (D floats each)
  const float* __restrict__ W;       // rhs rows (rhs_len floats each)
  float* __restrict__ out;
  int32_t* __restrict__ arg_u;
  int32_t* __restrict__ arg_e;
  const float* __restrict__ row_scale;
  const int32_t* __restrict__ row_order; // [n_rows] rows by non-increasing nnz (null: natural order)
  const int32_t* __restrict__ hub_rows;
  const int32_t* __restrict__ seg_ptr;   // [n_hub+1]
  const int32_t* __restrict__ seg_hub;   // [n_seg]
  float* __restrict__ ws_val;            // [n_seg][D] per-segment partial rows
  int32_t* __restrict__ ws_au;           // max/min only
  int32_t* __restrict__ ws_ae;
  int seg_len;
  int64_t n_rows;
  int D;        // out_len
  int xlen;     // relation-broadcast kernel: floats per lhs row (out_len = rhs_len * xlen)
  int rhs_len;  // floats per rhs row
  int inner;    // RMODE_HEAD: rhs column = k / inner
  int ncols;    // D / VEC
  int G, log2G;
  int hub_threshold;
  int skip_rows;   // the ordinary rows were already processed by the ring kernel (ring.cu): hub kernels only
  int accumulate;  // sum only: add the result to the existing contents of `out`
  int zero_inf;    // max/min only: store 0 where the result is +-inf (empty rows), as upstream's Python does
};

template <int VEC, int CH, int RED>
struct Acc {
  float a[CH][VEC];
  int32_t au[RED == DGLB_REDUCE_SUM ? 1 : CH][RED == DGLB_REDUCE_SUM ? 1 : VEC];
  int32_t ae[RED == DGLB_REDUCE_SUM ? 1 : CH][RED == DGLB_REDUCE_SUM ? 1 : VEC];
  __device__ __forceinline__ void init() {
    const float z = RED == DGLB_REDUCE_SUM ? 0.f : (RED == DGLB_REDUCE_MAX ? -INFINITY : INFINITY);
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        a[c][v] = z;
        if constexpr (RED != DGLB_REDUCE_SUM) { au[c][v] = 0; ae[c][v] = 0; }
      }
  }
};

// TU / TE: track the source-node / edge-id argument.  Upstream records arg_u only when the op reads
// the node operand and arg_e only when it reads the edge operand; an untracked argument stays 0 and
// costs no registers (copy_u_max: 98 -> 80 registers per thread).
template <int RED, bool TU = true, bool TE = true>
__device__ __forceinline__ void combine(float& acc, int32_t& au, int32_t& ae, float val, int32_t c,
                                        int32_t e) {
  if constexpr (RED == DGLB_REDUCE_SUM) {
    acc = __fadd_rn(acc, val);
  } else {
    const bool better = RED == DGLB_REDUCE_MAX ? (acc < val) : (acc > val);
    if (better) {
      acc = val;
      if constexpr (TU) au = c;
      if constexpr (TE) ae = e;
    }
  }
}

// Accumulate CSR positions [j0, j0+n) into `acc` for the feature tile starting at vector column
// `tile0`.  n is uniform within the group, nmax is the warp-wide maximum of n so every lane of
// the warp runs the same trip count (shuffles stay converged; extra trips are predicated off).
template <int VEC, int CH, int OP, int RED, int RMODE, typename T = float>
__device__ __forceinline__ void accumulate_range(const SpmmParams& p, int64_t j0, int n, int nmax,
                                                 int lg, int tile0, Acc<VEC, CH, RED>& acc) {
  static_assert(sizeof(T) == 4 || RMODE == RMODE_NONE, "bf16 storage is implemented for copy_lhs");
  // neighbour rows per load batch: 8 chunks in flight per lane; two-operand gathers (mul with a
  // per-column or per-head W) stage twice as much per row, so they batch half as many rows
  constexpr int U = (OP == DGLB_OP_MUL && (RMODE == RMODE_FULL || RMODE == RMODE_HEAD)) ? (CH >= 4 ? 1 : 4 / CH) : 8 / CH;
  constexpr bool USE_L = OP != DGLB_OP_COPY_RHS;
  constexpr bool USE_R = OP != DGLB_OP_COPY_LHS;
  constexpr bool TRACK_U = RED != DGLB_REDUCE_SUM && USE_L;
  constexpr bool TRACK_E = RED != DGLB_REDUCE_SUM && USE_R;
  constexpr bool NEED_E = USE_R;
  const int G = p.G;
  bool colv[CH];
  int k[CH], hk[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int vc = tile0 + c * G + lg;
    colv[c] = vc < p.ncols;
    k[c] = vc * VEC;
    hk[c] = (RMODE == RMODE_HEAD) ? k[c] / p.inner : 0;
  }
  for (int off = 0; off < nmax; off += G) {
    const int m = min(max(n - off, 0), G);
    int my_c = 0, my_e = 0;
    float my_w = 0.f;
    if (lg < m) {
      const int64_t j = j0 + off + lg;
      if (USE_L) my_c = __ldg(p.indices + j);
      if (NEED_E) my_e = p.eids ? __ldg(p.eids + j) : (int)j;
      // scalar edge weights ride along with the indices: one load per lane per G neighbours, then a
      // shuffle broadcast, instead of one (uniform-address) load per lane per neighbour
      if constexpr (RMODE == RMODE_SCALAR) my_w = __ldg(p.W + my_e);
    }
    const int mmax = min(G, nmax - off);
    for (int t = 0; t < mmax; t += U) {
      int cc[U], ee[U];
      float ww[RMODE == RMODE_SCALAR ? U : 1];
      StageVec<T, VEC> xv[U][CH];  // bf16 rows stay packed until consumed
      FVec<(VEC > 4 ? 4 : VEC)> wv[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cc[u] = USE_L ? __shfl_sync(FULL_MASK, my_c, t + u, G) : 0;
        ee[u] = (NEED_E && (RMODE != RMODE_SCALAR || TRACK_E)) ? __shfl_sync(FULL_MASK, my_e, t + u, G) : 0;
        // scalar weights: every shuffle of the batch happens BEFORE the gathers are issued -- with the weight
        // shuffles in the consume phase ptxas split the 8 gathers into 3 + 5 around them (SASS)
        if constexpr (RMODE == RMODE_SCALAR) ww[u] = __shfl_sync(FULL_MASK, my_w, t + u, G);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool valid = (t + u) < m;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if (valid && colv[c]) {
            if constexpr (USE_L)
              xv[u][c] = ldg_stage<T, VEC>(reinterpret_cast<const T*>(p.X) + (int64_t)cc[u] * p.D + k[c]);
            if constexpr (RMODE == RMODE_FULL)
              wv[u][c] = ldg_vec_t<float, (VEC > 4 ? 4 : VEC)>(p.W + (int64_t)ee[u] * p.D + k[c]);
            if constexpr (RMODE == RMODE_HEAD)
              wv[u][c].v[0] = __ldg(p.W + (int64_t)ee[u] * p.rhs_len + hk[c]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool valid = (t + u) < m;
        float ws = 0.f;
        if constexpr (RMODE == RMODE_SCALAR) ws = ww[u];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if (valid && colv[c]) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              float val;
              const float w = (RMODE == RMODE_SCALAR) ? ws
                              : (RMODE == RMODE_HEAD) ? wv[u][c].v[0] : (USE_R ? wv[u][c].v[v] : 0.f);
              if constexpr (OP == DGLB_OP_COPY_LHS) val = xv[u][c].at(v);
              else if constexpr (OP == DGLB_OP_COPY_RHS) val = w;
              else if constexpr (OP == DGLB_OP_MUL) val = __fmul_rn(xv[u][c].at(v), w);
              else val = __fadd_rn(xv[u][c].at(v), w);
              if constexpr (RED == DGLB_REDUCE_SUM) {
                acc.a[c][v] = __fadd_rn(acc.a[c][v], val);
              } else {
                combine<RED, TRACK_U, TRACK_E>(acc.a[c][v], acc.au[c][v], acc.ae[c][v], val, cc[u], ee[u]);
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ row-per-group kernel
// Residency matters more than anything else here: the same copy_u_sum kernel ran 2.72 ms at 70
// registers (3 CTAs / SM) and 4.78 ms at 84 registers (2 CTAs / SM) on the products graph, and
// copy_u_max went from 4.04 ms (98 registers) to 2.88 ms (80).  The max/min and scalar-weight variants
// are therefore held to 80 registers (except the widest tiles, which would spill in the gather loop).
// The copy/sum variants already fit and keep ptxas' own allocation (0 = no minimum): forcing the bound
// on them made ptxas start consuming the first gathered row before the last one was issued.
template <int OP, int RED>
constexpr int spmm_min_ctas(int rmode, int ch, int vec) {
  if (RED != DGLB_REDUCE_SUM) return ch * vec >= 16 ? 0 : 3;
  if (rmode == RMODE_SCALAR) return vec == 4 ? 3 : 0;  // what keeps the 8 gathers of a batch together (SASS-checked)
  return 0;
}

// ORDERED: rows are handed out through p.row_order.  A separate instantiation, so that the natural-order kernels keep
// exactly the code (and the gather batching ptxas gives it, tests/test_sass_load_batching.py) they had before: with a
// runtime `if (p.row_order)` in the one kernel ptxas split the 8-gather batch of the scalar-weight variant into 3 + 5.
template <int VEC, int CH, int OP, int RED, int RMODE, typename T = float, bool ORDERED = false>
__global__ void __launch_bounds__(kBlockThreads, spmm_min_ctas<OP, RED>(RMODE, CH, VEC))
spmm_rows_kernel(const SpmmParams p) {
  const int G = p.G;
  const int lg = threadIdx.x & (G - 1);
  int64_t row = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> p.log2G;
  int row_start = 0, deg = 0;
  bool active = row < p.n_rows;
  if (active) {
    // degree-ordered hand-out: the groups of a warp get rows of (nearly) equal length, longest rows first
    if constexpr (ORDERED) row = __ldg(p.row_order + row);
    row_start = __ldg(p.indptr + row);
    deg = __ldg(p.indptr + row + 1) - row_start;
    if (deg > p.hub_threshold) { active = false; deg = 0; }  // left to the hub kernel
  }
  float scale = 1.f;
  if (p.row_scale && active) scale = __ldg(p.row_scale + row);
  const int nmax = __reduce_max_sync(FULL_MASK, deg);
  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    Acc<VEC, CH, RED> acc;
    acc.init();
    accumulate_range<VEC, CH, OP, RED, RMODE, T>(p, row_start, deg, nmax, lg, tile0, acc);
    if (active) {
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int vc = tile0 + c * G + lg;
        if (vc < p.ncols) {
          FVec<VEC> o;
          const int64_t off = row * (int64_t)p.D + (int64_t)vc * VEC;
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            o.v[v] = p.row_scale ? __fdiv_rn(acc.a[c][v], scale) : acc.a[c][v];
          if constexpr (RED == DGLB_REDUCE_SUM) {
            if (p.accumulate) {
              const FVec<VEC> prev = ldg_vec_t<T, VEC>(reinterpret_cast<const T*>(p.out) + off);
#pragma unroll
              for (int v = 0; v < VEC; ++v) o.v[v] = __fadd_rn(prev.v[v], o.v[v]);
            }
          } else {
            if (p.zero_inf) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) o.v[v] = isinf(o.v[v]) ? 0.f : o.v[v];
            }
          }
          st_vec_t<T, VEC>(reinterpret_cast<T*>(p.out) + off, o);
          if constexpr (RED != DGLB_REDUCE_SUM) {
            if (p.arg_u) st_vec_i32<VEC>(p.arg_u + off, acc.au[c]);
            if (p.arg_e) st_vec_i32<VEC>(p.arg_e + off, acc.ae[c]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ hub rows: one CTA per segment
template <int VEC, int CH, int OP, int RED, int RMODE, typename T = float>
__global__ void __launch_bounds__(kBlockThreads)
spmm_hub_kernel(const SpmmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = p.G;
  const int tile_elems = G * CH * VEC;
  const int n_groups = kBlockThreads >> p.log2G;
  float* s_val = reinterpret_cast<float*>(smem_raw);                   // [n_groups][tile_elems]
  int32_t* s_au = reinterpret_cast<int32_t*>(s_val + n_groups * tile_elems);
  int32_t* s_ae = s_au + n_groups * tile_elems;

  const int seg = blockIdx.x;
  const int hub = __ldg(p.seg_hub + seg);
  const int64_t row = __ldg(p.hub_rows + hub);
  const int k = seg - __ldg(p.seg_ptr + hub);                          // segment number inside the row
  const int lg = threadIdx.x & (G - 1);
  const int gidx = threadIdx.x >> p.log2G;
  const int row_start = __ldg(p.indptr + row);
  const int row_deg = __ldg(p.indptr + row + 1) - row_start;
  const int seg_begin = k * p.seg_len;
  const int deg = min(p.seg_len, row_deg - seg_begin);                 // entries of this segment
  const int per = (deg + n_groups - 1) / n_groups;
  const int my_begin = min(gidx * per, deg);
  const int my_n = min(per, deg - my_begin);
  const int nmax = __reduce_max_sync(FULL_MASK, my_n);

  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    Acc<VEC, CH, RED> acc;
    acc.init();
    accumulate_range<VEC, CH, OP, RED, RMODE, T>(p, (int64_t)row_start + seg_begin + my_begin, my_n, nmax, lg,
                                                 tile0, acc);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int el = (c * G + lg) * VEC;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        s_val[gidx * tile_elems + el + v] = acc.a[c][v];
        if constexpr (RED != DGLB_REDUCE_SUM) {
          s_au[gidx * tile_elems + el + v] = acc.au[c][v];
          s_ae[gidx * tile_elems + el + v] = acc.ae[c][v];
        }
      }
    }
    __syncthreads();
    for (int el = threadIdx.x; el < tile_elems; el += kBlockThreads) {
      const int kk = tile0 * VEC + el;
      if (kk < p.D) {
        float a = s_val[el];
        int32_t au = 0, ae = 0;
        if constexpr (RED != DGLB_REDUCE_SUM) { au = s_au[el]; ae = s_ae[el]; }
        for (int g = 1; g < n_groups; ++g) {  // slice order == CSR order: first still wins ties
          const float b = s_val[g * tile_elems + el];
          if constexpr (RED == DGLB_REDUCE_SUM) {
            a = __fadd_rn(a, b);
          } else {
            combine<RED>(a, au, ae, b, s_au[g * tile_elems + el], s_ae[g * tile_elems + el]);
          }
        }
        const int64_t off = (int64_t)seg * p.D + kk;
        p.ws_val[off] = a;
        if constexpr (RED != DGLB_REDUCE_SUM) { p.ws_au[off] = au; p.ws_ae[off] = ae; }
      }
    }
    __syncthreads();
  }
}

// fold the per-segment partial rows of every hub row, in segment (= CSR) order
template <int RED, typename T = float>
__global__ void __launch_bounds__(kBlockThreads) spmm_hub_combine_kernel(const SpmmParams p, int n_hub) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (idx >= (int64_t)n_hub * p.D) return;
  const int hub = (int)(idx / p.D);
  const int kk = (int)(idx - (int64_t)hub * p.D);
  const int s0 = __ldg(p.seg_ptr + hub), s1 = __ldg(p.seg_ptr + hub + 1);
  const int64_t row = __ldg(p.hub_rows + hub);
  float a = p.ws_val[(int64_t)s0 * p.D + kk];
  int32_t au = 0, ae = 0;
  if constexpr (RED != DGLB_REDUCE_SUM) { au = p.ws_au[(int64_t)s0 * p.D + kk]; ae = p.ws_ae[(int64_t)s0 * p.D + kk]; }
  if constexpr (RED == DGLB_REDUCE_SUM) {
    // fetch 8 partial rows at a time (independent loads), add them in segment order
    for (int sg = s0 + 1; sg < s1; sg += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (sg + i < s1) ? p.ws_val[(int64_t)(sg + i) * p.D + kk] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (sg + i < s1) a = __fadd_rn(a, v[i]);
    }
  } else {
    for (int sg = s0 + 1; sg < s1; ++sg) {
      const int64_t o = (int64_t)sg * p.D + kk;
      combine<RED>(a, au, ae, p.ws_val[o], p.ws_au[o], p.ws_ae[o]);
    }
  }
  const int64_t off = row * (int64_t)p.D + kk;
  if (p.row_scale) a = __fdiv_rn(a, __ldg(p.row_scale + row));
  if constexpr (RED == DGLB_REDUCE_SUM) {
    if (p.accumulate) a = __fadd_rn(load_scalar_t<T>(reinterpret_cast<const T*>(p.out) + off), a);
  } else {
    if (p.zero_inf && isinf(a)) a = 0.f;
  }
  store_scalar_t<T>(reinterpret_cast<T*>(p.out) + off, a);
  if constexpr (RED != DGLB_REDUCE_SUM) {
    if (p.arg_u) p.arg_u[off] = au;
    if (p.arg_e) p.arg_e[off] = ae;
  }
}

// ------------------------------------------------------------------ relation-broadcast rows (RGCN)
// out[v, r, :] = sum_{(u->v)} W[eid, r] * X[u, :]  for r < RR: lhs (N, 1, D) x rhs (E, RR, 1) -> (N, RR, D), the batched form
// of main_dgl_proteins_rgcn_for.py:50-53, which runs one update_all(u_mul_e, mean) per relation on the SAME graph and the
// SAME node features: here a neighbour row is gathered ONCE for all RR relations (168 instead of 8 x 140 bytes per edge at
// D = 32, RR = 8).  Row-per-group like spmm_rows_kernel, one vector column per lane (D <= 32 * VEC), accumulators
// acc[RR][VEC] in registers; per output element the terms are multiplied and added in CSR order with non-contracted ops,
// so every relation's slice is bit-identical to the per-relation u_mul_e_sum.  No hub path: meant for graphs without
// extreme rows (ogbn-proteins: in-degree ~600 everywhere).
template <int VEC, int RR>
__global__ void __launch_bounds__(kBlockThreads, 2) spmm_rel_rows_kernel(const SpmmParams p) {
  constexpr int U = 16 / RR >= 8 ? 8 : (16 / RR < 2 ? 2 : 16 / RR);
  const int G = p.G;
  const int lg = threadIdx.x & (G - 1);
  const int64_t row = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> p.log2G;
  int row_start = 0, deg = 0;
  const bool active = row < p.n_rows;
  if (active) {
    row_start = __ldg(p.indptr + row);
    deg = __ldg(p.indptr + row + 1) - row_start;
  }
  const int nmax = __reduce_max_sync(FULL_MASK, deg);
  const bool colv = lg < p.ncols;
  const int k = lg * VEC;
  float acc[RR][VEC];
#pragma unroll
  for (int r = 0; r < RR; ++r)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;
  for (int off = 0; off < nmax; off += G) {
    const int m = min(max(deg - off, 0), G);
    int my_c = 0, my_e = 0;
    if (lg < m) {
      const int64_t j = (int64_t)row_start + off + lg;
      my_c = __ldg(p.indices + j);
      my_e = p.eids ? __ldg(p.eids + j) : (int)j;
    }
    const int mmax = min(G, nmax - off);
    for (int t = 0; t < mmax; t += U) {
      int cc[U], ee[U];
      FVec<VEC> xv[U];
      float w[U][RR];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
        ee[u] = __shfl_sync(FULL_MASK, my_e, t + u, G);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if ((t + u) < m && colv) {
          xv[u] = ldg_vec<VEC>(p.X + (int64_t)cc[u] * p.xlen + k);
          const float* wp = p.W + (int64_t)ee[u] * RR;
          if constexpr (RR % 4 == 0) {
#pragma unroll
            for (int r = 0; r < RR; r += 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(wp + r));
              w[u][r] = q.x; w[u][r + 1] = q.y; w[u][r + 2] = q.z; w[u][r + 3] = q.w;
            }
          } else {
            const float2 q = __ldg(reinterpret_cast<const float2*>(wp));
            w[u][0] = q.x; w[u][1] = q.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if ((t + u) < m && colv) {
#pragma unroll
          for (int r = 0; r < RR; ++r)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[r][v] = __fadd_rn(acc[r][v], __fmul_rn(xv[u].v[v], w[u][r]));
        }
      }
    }
  }
  if (active && colv) {
    const float scale = p.row_scale ? __ldg(p.row_scale + row) : 1.f;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      FVec<VEC> o;
#pragma unroll
      for (int v = 0; v < VEC; ++v) o.v[v] = p.row_scale ? __fdiv_rn(acc[r][v], scale) : acc[r][v];
      st_vec<VEC>(p.out + row * (int64_t)p.D + (int64_t)r * p.xlen + k, o);
    }
  }
}

template <int VEC>
static int launch_rel(const SpmmParams& p, int rr, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks <= 0) return DGLB_OK;
  if (rr == 8) spmm_rel_rows_kernel<VEC, 8><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  else if (rr == 4) spmm_rel_rows_kernel<VEC, 4><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  else spmm_rel_rows_kernel<VEC, 2><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("spmm_rel_rows_kernel");
  return DGLB_OK;
}

// ------------------------------------------------------------------ generic fallback
struct GenericSpmmParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* eids;
  const float* X;
  const float* W;
  float* out;
  int32_t* arg_u;
  int32_t* arg_e;
  const float* row_scale;
  int64_t n_rows;
  int op, red;
  int accumulate, zero_inf;
  BcastShape b;
};

__device__ __forceinline__ float generic_binop(int op, float l, float r) {
  switch (op) {
    case DGLB_OP_ADD: return __fadd_rn(l, r);
    case DGLB_OP_SUB: return __fsub_rn(l, r);
    case DGLB_OP_MUL: return __fmul_rn(l, r);
    case DGLB_OP_DIV: return __fdiv_rn(l, r);
    case DGLB_OP_COPY_LHS: return l;
    default: return r;
  }
}

__global__ void __launch_bounds__(kBlockThreads) spmm_generic_kernel(const GenericSpmmParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int64_t D = p.b.out_len;
  if (idx >= p.n_rows * D) return;
  const int64_t row = idx / D;
  int64_t rem = idx - row * D;
  int64_t lk = 0, rk = 0, sl = 1, sr = 1;
  for (int d = p.b.ndim - 1; d >= 0; --d) {
    const int64_t i = rem % p.b.out[d];
    rem /= p.b.out[d];
    lk += (p.b.lhs[d] == 1 ? 0 : i) * sl;
    rk += (p.b.rhs[d] == 1 ? 0 : i) * sr;
    sl *= p.b.lhs[d];
    sr *= p.b.rhs[d];
  }
  const bool use_l = p.op != DGLB_OP_COPY_RHS, use_r = p.op != DGLB_OP_COPY_LHS;
  const int start = p.indptr[row], end = p.indptr[row + 1];
  float acc = p.red == DGLB_REDUCE_SUM ? 0.f : (p.red == DGLB_REDUCE_MAX ? -INFINITY : INFINITY);
  int32_t au = 0, ae = 0;
  for (int j = start; j < end; ++j) {
    const int32_t c = p.indices[j];
    const int32_t e = p.eids ? p.eids[j] : j;
    const float l = use_l ? __ldg(p.X + (int64_t)c * p.b.lhs_len + lk) : 0.f;
    const float r = use_r ? __ldg(p.W + (int64_t)e * p.b.rhs_len + rk) : 0.f;
    const float val = generic_binop(p.op, l, r);
    if (p.red == DGLB_REDUCE_SUM) acc = __fadd_rn(acc, val);
    else if (p.red == DGLB_REDUCE_MAX) { if (acc < val) { acc = val; au = c; ae = e; } }
    else { if (acc > val) { acc = val; au = c; ae = e; } }
  }
  if (p.row_scale) acc = __fdiv_rn(acc, p.row_scale[row]);
  if (p.accumulate) acc = __fadd_rn(p.out[idx], acc);
  if (p.zero_inf && p.red != DGLB_REDUCE_SUM && isinf(acc)) acc = 0.f;
  p.out[idx] = acc;
  if (p.red != DGLB_REDUCE_SUM) {
    if (p.arg_u) p.arg_u[idx] = au;
    if (p.arg_e) p.arg_e[idx] = ae;
  }
}

// ------------------------------------------------------------------ dispatch
template <int VEC, int CH, int OP, int RED, int RMODE, typename T = float>
static int launch_fast(const SpmmParams& p, int n_hub, int n_seg, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks > 0 && !p.skip_rows) {
    if (p.row_order) spmm_rows_kernel<VEC, CH, OP, RED, RMODE, T, true><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    else spmm_rows_kernel<VEC, CH, OP, RED, RMODE, T, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("spmm_rows_kernel");
  }
  if (n_hub > 0 && n_seg > 0) {
    const size_t smem = (size_t)kBlockThreads * CH * VEC * 4 * (RED == DGLB_REDUCE_SUM ? 1 : 3);
    if (smem > 48 * 1024)  // per device and cheap: set on every launch that needs it (no process-wide latch)
      DGLB_CUDA(cudaFuncSetAttribute(spmm_hub_kernel<VEC, CH, OP, RED, RMODE, T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spmm_hub_kernel<VEC, CH, OP, RED, RMODE, T><<<n_seg, kBlockThreads, smem, stream>>>(p);
    DGLB_LAUNCH_CHECK("spmm_hub_kernel");
    const int64_t cblocks = ((int64_t)n_hub * p.D + kBlockThreads - 1) / kBlockThreads;
    spmm_hub_combine_kernel<RED, T><<<(unsigned)cblocks, kBlockThreads, 0, stream>>>(p, n_hub);
    DGLB_LAUNCH_CHECK("spmm_hub_combine_kernel");
  }
  return DGLB_OK;
}

template <int OP, int RED, int RMODE>
static int dispatch_vec_ch(SpmmParams& p, int vec, int n_hub, int n_seg, cudaStream_t stream) {
  p.ncols = p.D / vec;
  p.G = group_lanes(p.ncols);
  p.log2G = 0;
  while ((1 << p.log2G) < p.G) ++p.log2G;
  const int per_lane = (p.ncols + p.G - 1) / p.G;
  const int ch = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
#define DGLB_CASE(V, C) \
  if (vec == V && ch == C) return launch_fast<V, C, OP, RED, RMODE>(p, n_hub, n_seg, stream);
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  set_error("spmm: no kernel for vec=%d ch=%d", vec, ch);
  return DGLB_E_UNSUPPORTED;
}

template <int OP, int RMODE>
static int dispatch_red(SpmmParams& p, int red, int vec, int n_hub, int n_seg, cudaStream_t stream) {
  switch (red) {
    case DGLB_REDUCE_SUM: return dispatch_vec_ch<OP, DGLB_REDUCE_SUM, RMODE>(p, vec, n_hub, n_seg, stream);
    case DGLB_REDUCE_MAX: return dispatch_vec_ch<OP, DGLB_REDUCE_MAX, RMODE>(p, vec, n_hub, n_seg, stream);
    default: return dispatch_vec_ch<OP, DGLB_REDUCE_MIN, RMODE>(p, vec, n_hub, n_seg, stream);
  }
}

static int and_vec(int a, int b) { return a < b ? a : b; }

int spmm_csr_f32(int op, int reduce, int64_t n_rows, int64_t n_cols, int64_t nnz,
                 const int32_t* indptr, const int32_t* indices, const int32_t* eids, const float* X,
                 const float* W, const BcastShape& b, float* out, int32_t* arg_u, int32_t* arg_e,
                 const float* row_scale, int flags, const dglb_hub_t* hub, cudaStream_t stream) {
  if (n_rows == 0 || b.out_len == 0) return DGLB_OK;
  const bool use_l = op != DGLB_OP_COPY_RHS, use_r = op != DGLB_OP_COPY_LHS;
  // ---- classify the broadcast pattern
  int rmode = -1;
  int64_t inner = 1;
  if (op == DGLB_OP_COPY_LHS) {
    rmode = RMODE_NONE;
  } else if (op == DGLB_OP_COPY_RHS) {
    rmode = RMODE_FULL;
  } else if (b.lhs_len == b.out_len) {
    if (b.rhs_len == b.out_len) {
      rmode = RMODE_FULL;
    } else {
      // rhs equals lhs on the leading dims and is 1 on the trailing ones => column = k / inner
      int t = b.ndim;
      while (t > 0 && b.rhs[t - 1] == 1) --t;
      bool ok = true;
      for (int d = 0; d < t; ++d) ok = ok && (b.rhs[d] == b.lhs[d]);
      if (ok) {
        for (int d = t; d < b.ndim; ++d) inner *= b.lhs[d];
        rmode = RMODE_HEAD;
      }
    }
  }
  // relation broadcast: lhs (1, D) x rhs (R, 1) -> (R, D), mul / sum, no accumulate, no hub rows handed in
  if (op == DGLB_OP_MUL && reduce == DGLB_REDUCE_SUM && b.ndim == 2 && b.lhs[0] == 1 && b.rhs[1] == 1 && b.lhs[1] > 1 &&
      (b.rhs[0] == 2 || b.rhs[0] == 4 || b.rhs[0] == 8) && !(flags & DGLB_SPMM_ACCUMULATE) &&
      !(hub && hub->n_hub > 0) && b.lhs[1] <= 128) {
    const int64_t xl = b.lhs[1], rr = b.rhs[0];
    int vec = and_vec(pick_vec(xl, X), pick_vec(xl, out));
    if (rr % 4 ? (reinterpret_cast<uintptr_t>(W) % 8 != 0) : (reinterpret_cast<uintptr_t>(W) % 16 != 0)) vec = 0;
    if (vec > 0 && xl / vec <= 32) {
      SpmmParams p;
      p.row_order = nullptr;
      memset(&p, 0, sizeof(p));
      p.indptr = indptr; p.indices = indices; p.eids = eids; p.X = X; p.W = W; p.out = out; p.row_scale = row_scale;
      p.n_rows = n_rows; p.D = (int)b.out_len; p.xlen = (int)xl; p.rhs_len = (int)rr;
      p.ncols = (int)(xl / vec);
      p.G = group_lanes(p.ncols);
      while ((1 << p.log2G) < p.G) ++p.log2G;
      p.hub_threshold = INT32_MAX;
      if (vec == 4) return launch_rel<4>(p, (int)rr, stream);
      if (vec == 2) return launch_rel<2>(p, (int)rr, stream);
      return launch_rel<1>(p, (int)rr, stream);
    }
  }
  const bool fast_op = (op == DGLB_OP_COPY_LHS || op == DGLB_OP_COPY_RHS ||
                        (op == DGLB_OP_MUL && reduce == DGLB_REDUCE_SUM));
  if (rmode >= 0 && fast_op && b.out_len < (1 << 30)) {
    SpmmParams p;
    p.row_order = nullptr;
    p.indptr = indptr; p.indices = indices; p.eids = eids;
    p.X = X; p.W = W; p.out = out; p.arg_u = arg_u; p.arg_e = arg_e;
    p.row_scale = row_scale;
    const bool with_args = reduce != DGLB_REDUCE_SUM;
    const bool use_hub = hub && hub->n_hub > 0 && hub->n_seg > 0 && hub->rows && hub->seg_ptr && hub->seg_hub &&
                         hub->seg_len > 0;
    int n_hub = use_hub ? hub->n_hub : 0;
    const int n_seg = use_hub ? hub->n_seg : 0;
    if (use_hub) {
      const size_t need = (size_t)n_seg * (size_t)b.out_len * 4 * (with_args ? 3 : 1);
      if (!hub->workspace || hub->workspace_bytes < need) {
        set_error("gspmm: hub workspace too small (%zu < %zu bytes)", hub->workspace_bytes, need);
        return DGLB_E_WORKSPACE;
      }
      p.hub_rows = hub->rows; p.seg_ptr = hub->seg_ptr; p.seg_hub = hub->seg_hub; p.seg_len = hub->seg_len;
      p.ws_val = static_cast<float*>(hub->workspace);
      p.ws_au = with_args ? reinterpret_cast<int32_t*>(p.ws_val + (size_t)n_seg * b.out_len) : nullptr;
      p.ws_ae = with_args ? p.ws_au + (size_t)n_seg * b.out_len : nullptr;
    } else {
      p.hub_rows = nullptr; p.seg_ptr = nullptr; p.seg_hub = nullptr; p.seg_len = 0;
      p.ws_val = nullptr; p.ws_au = nullptr; p.ws_ae = nullptr;
    }
    p.n_rows = n_rows; p.D = (int)b.out_len; p.rhs_len = (int)b.rhs_len; p.inner = (int)inner;
    p.row_order = hub ? hub->row_order : nullptr;
    p.hub_threshold = use_hub ? hub->threshold : INT32_MAX;
    p.accumulate = (flags & DGLB_SPMM_ACCUMULATE) ? 1 : 0;
    p.zero_inf = (flags & DGLB_SPMM_ZERO_INF) ? 1 : 0;
    p.skip_rows = 0;
    if (op == DGLB_OP_COPY_LHS && reduce == DGLB_REDUCE_SUM) {
      // wide rows: whole-row bulk copies into a shared-memory ring (ring.cu); hub rows stay on the segmented path
      const int rc = ring_rows(false, DGLB_F32, n_rows, n_cols, nnz, indptr, indices, nullptr, X, nullptr, b.out_len,
                               out, row_scale, p.accumulate, p.hub_threshold, use_hub ? hub->light_indptr : nullptr, stream);
      if (rc == DGLB_OK) {
        if (!use_hub) return DGLB_OK;
        p.skip_rows = 1;
      } else if (rc != DGLB_E_UNSUPPORTED) {
        return rc;
      }
    }
    int vec = pick_vec(b.out_len, out);
    if (use_l) vec = and_vec(vec, pick_vec(b.out_len, X));
    if (rmode == RMODE_FULL) vec = and_vec(vec, pick_vec(b.out_len, W));
    if (rmode == RMODE_HEAD && b.rhs_len == 1) rmode = RMODE_SCALAR;
    if (rmode == RMODE_HEAD) { while (inner % vec) vec >>= 1; }
    if (reduce != DGLB_REDUCE_SUM) {
      if (arg_u) vec = and_vec(vec, pick_vec(b.out_len, arg_u));
      if (arg_e) vec = and_vec(vec, pick_vec(b.out_len, arg_e));
    }
    if (op == DGLB_OP_COPY_LHS) return dispatch_red<DGLB_OP_COPY_LHS, RMODE_NONE>(p, reduce, vec, n_hub, n_seg, stream);
    if (op == DGLB_OP_COPY_RHS) return dispatch_red<DGLB_OP_COPY_RHS, RMODE_FULL>(p, reduce, vec, n_hub, n_seg, stream);
    if (rmode == RMODE_FULL) return dispatch_vec_ch<DGLB_OP_MUL, DGLB_REDUCE_SUM, RMODE_FULL>(p, vec, n_hub, n_seg, stream);
    if (rmode == RMODE_SCALAR) return dispatch_vec_ch<DGLB_OP_MUL, DGLB_REDUCE_SUM, RMODE_SCALAR>(p, vec, n_hub, n_seg, stream);
    return dispatch_vec_ch<DGLB_OP_MUL, DGLB_REDUCE_SUM, RMODE_HEAD>(p, vec, n_hub, n_seg, stream);
  }
  // ---- generic path
  GenericSpmmParams g;
  g.indptr = indptr; g.indices = indices; g.eids = eids; g.X = X; g.W = W; g.out = out;
  g.arg_u = arg_u; g.arg_e = arg_e; g.row_scale = row_scale; g.n_rows = n_rows;
  g.op = op; g.red = reduce; g.b = b;
  g.accumulate = (flags & DGLB_SPMM_ACCUMULATE) ? 1 : 0;
  g.zero_inf = (flags & DGLB_SPMM_ZERO_INF) ? 1 : 0;
  (void)use_r;
  const int64_t total = n_rows * b.out_len;
  const int64_t blocks = (total + kBlockThreads - 1) / kBlockThreads;
  if (blocks > 0x7fffffffLL) { set_error("spmm generic: problem too large"); return DGLB_E_UNSUPPORTED; }
  spmm_generic_kernel<<<(unsigned)blocks, kBlockThreads, 0, stream>>>(g);
  DGLB_LAUNCH_CHECK("spmm_generic_kernel");
  return DGLB_OK;
}

// bf16 storage / fp32 accumulate: copy_lhs x sum (the SAGE aggregation and the micro-benchmark op)
int spmm_csr_bf16(int op, int reduce, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* indptr,
                  const int32_t* indices, const void* X, int64_t D, void* out, const float* row_scale,
                  int accumulate, const dglb_hub_t* hub, cudaStream_t stream) {
  if (op != DGLB_OP_COPY_LHS || reduce != DGLB_REDUCE_SUM) {
    set_error("gspmm: bf16 storage is implemented for copy_lhs with reducer sum (mean through row_scale)");
    return DGLB_E_UNSUPPORTED;
  }
  if (n_rows == 0 || D == 0) return DGLB_OK;
  if (D >= (1 << 30)) return DGLB_E_UNSUPPORTED;
  SpmmParams p;
  p.row_order = nullptr;
  p.indptr = indptr; p.indices = indices; p.eids = nullptr;
  p.X = static_cast<const float*>(X); p.W = nullptr; p.out = static_cast<float*>(out);
  p.arg_u = nullptr; p.arg_e = nullptr; p.row_scale = row_scale;
  p.n_rows = n_rows; p.D = (int)D; p.rhs_len = 0; p.inner = 1;
  p.accumulate = (accumulate & DGLB_SPMM_ACCUMULATE) ? 1 : 0; p.zero_inf = 0;
  const bool use_hub = hub && hub->n_hub > 0 && hub->n_seg > 0 && hub->rows && hub->seg_ptr && hub->seg_hub &&
                       hub->seg_len > 0;
  const int n_hub = use_hub ? hub->n_hub : 0, n_seg = use_hub ? hub->n_seg : 0;
  if (use_hub) {
    const size_t need = (size_t)n_seg * (size_t)D * 4;
    if (!hub->workspace || hub->workspace_bytes < need) { set_error("gspmm: hub workspace too small"); return DGLB_E_WORKSPACE; }
    p.hub_rows = hub->rows; p.seg_ptr = hub->seg_ptr; p.seg_hub = hub->seg_hub; p.seg_len = hub->seg_len;
    p.ws_val = static_cast<float*>(hub->workspace);
  } else {
    p.hub_rows = nullptr; p.seg_ptr = nullptr; p.seg_hub = nullptr; p.seg_len = 0; p.ws_val = nullptr;
  }
  p.ws_au = nullptr; p.ws_ae = nullptr;
  p.row_order = hub ? hub->row_order : nullptr;
  p.hub_threshold = use_hub ? hub->threshold : INT32_MAX;
  p.skip_rows = 0;
  {
    const int rc = ring_rows(false, DGLB_BF16, n_rows, n_cols, nnz, indptr, indices, nullptr, X, nullptr, D, out,
                             row_scale, p.accumulate, p.hub_threshold, use_hub ? hub->light_indptr : nullptr, stream);
    if (rc == DGLB_OK) {
      if (!use_hub) return DGLB_OK;
      p.skip_rows = 1;
    } else if (rc != DGLB_E_UNSUPPORTED) {
      return rc;
    }
  }
  const int vec = min_int(pick_vec_bf16(D, X), pick_vec_bf16(D, out));
  p.ncols = (int)(D / vec);
  p.G = group_lanes(p.ncols);
  p.log2G = 0;
  while ((1 << p.log2G) < p.G) ++p.log2G;
  const int per_lane = (p.ncols + p.G - 1) / p.G;
  const int ch = per_lane >= 3 ? 4 : (per_lane >= 2 ? 2 : 1);
#define DGLB_CASE(V, C) \
  if (vec == V && ch == C) \
    return launch_fast<V, C, DGLB_OP_COPY_LHS, DGLB_REDUCE_SUM, RMODE_NONE, __nv_bfloat16>(p, n_hub, n_seg, stream);
  DGLB_CASE(8, 1) DGLB_CASE(8, 2) DGLB_CASE(8, 4)
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  return DGLB_E_UNSUPPORTED;
}

}  // namespace dglb
