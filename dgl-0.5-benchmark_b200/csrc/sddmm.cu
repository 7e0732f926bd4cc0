// sddmm.cu -- generalized SDDMM (per-edge op of a source-node and a destination-node row),
// hand-written for sm_100a.
//
// Replaces upstream DGL v0.6.1 src/array/cuda/sddmm.cuh (SDDMMCooKernel: a thread-row per edge,
// TWO random row gathers per edge and a sequential per-thread dot; SDDMMCsrKernel: the same plus a
// per-edge binary search) behind `_CAPI_DGLKernelSDDMM`; reached from kernel/dgl-new.py:39 and
// from GATConv (u_add_v forward, u_dot_v in the backward of u_mul_e_sum).
//
// Design: destination-major traversal of the CSC.  A group of G lanes owns one destination row,
// keeps that row's V tile in registers and streams the row's in-edges: per edge only the source
// row is gathered (half the gather traffic of the edge-parallel form).  `dot` partial products
// are reduced across the group with warp shuffles (segmented by head for (N,H,F) operands) and
// written to out[eid]; elementwise ops store the combined row at out[eid,:].  Hub rows get one
// CTA each whose groups split the row's edges (no combine step is needed).  Shapes the vector
// path does not cover use a generic thread-per-(edge,feature) kernel (COO or CSR form).
#include "kernels.cuh"

namespace dglb {

struct SddmmParams {
  const int32_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ eids;
  const float* __restrict__ U;  // lhs rows gathered by source id
  const float* __restrict__ V;  // rhs rows, one per destination row
  float* __restrict__ out;
  const int32_t* __restrict__ row_order;  // [n_rows] rows by non-increasing nnz (null: natural order)
  const int32_t* __restrict__ hub_rows;
  const int32_t* __restrict__ seg_ptr;  // [n_hub+1] first segment of each hub row
  const int32_t* __restrict__ seg_hub;  // [n_seg]   hub index of each segment
  int seg_len;
  int64_t n_rows;
  int D;      // floats per node row
  int ncols;  // D / VEC
  int G, log2G;
  int H;      // dot: number of output columns (heads)
  int seg;    // dot: lanes per head
  int hub_threshold;
  int skip_rows;  // dot: the ordinary rows were already processed by the ring kernel (ring.cu)
};

template <int OP>
__device__ __forceinline__ float ew_op(float l, float r) {
  if constexpr (OP == DGLB_OP_ADD) return __fadd_rn(l, r);
  else if constexpr (OP == DGLB_OP_SUB) return __fsub_rn(l, r);
  else if constexpr (OP == DGLB_OP_MUL) return __fmul_rn(l, r);
  else if constexpr (OP == DGLB_OP_DIV) return __fdiv_rn(l, r);
  else if constexpr (OP == DGLB_OP_COPY_LHS) return l;
  else return r;
}

// (row, first CSR position, count) of the calling group; HUB: one CTA per hub-row SEGMENT (<= seg_len
// edges), its groups take contiguous slices -- edges are independent, so nothing is combined
template <bool HUB>
__device__ __forceinline__ void group_work(const SddmmParams& p, int64_t& row, int64_t& j0, int& n) {
  if constexpr (!HUB) {
    row = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> p.log2G;
    j0 = 0; n = 0;
    if (row < p.n_rows) {
      if (p.row_order) row = __ldg(p.row_order + row);   // degree-ordered hand-out (see dglb_hub_t)
      const int s = __ldg(p.indptr + row);
      const int d = __ldg(p.indptr + row + 1) - s;
      if (d <= p.hub_threshold) { j0 = s; n = d; }
    } else {
      row = 0;
    }
  } else {
    const int seg = blockIdx.x;
    const int hub = __ldg(p.seg_hub + seg);
    row = __ldg(p.hub_rows + hub);
    const int k = seg - __ldg(p.seg_ptr + hub);
    const int n_groups = kBlockThreads >> p.log2G;
    const int gidx = threadIdx.x >> p.log2G;
    const int s = __ldg(p.indptr + row) + k * p.seg_len;
    const int d = min(p.seg_len, __ldg(p.indptr + row + 1) - s);
    const int per = (d + n_groups - 1) / n_groups;
    const int b = min(gidx * per, d);
    j0 = (int64_t)s + b;
    n = min(per, d - b);
  }
}

// ------------------------------------------------------------------ dot
// Reduce U per-lane partial sums across a group of G = 2^LOGG lanes with a "transposing" butterfly:
// each exchange step halves the number of live values per lane (lanes whose `mask` bit is set keep
// the upper half), so U values cost (U-1) + log2(G/U) shuffles instead of U*log2(G).  Afterwards
// lane lg holds, in v[0..max(1,U/G)), the totals of edges ((lg*U) >> LOGG) + i.
template <int U, int LOGG>
__device__ __forceinline__ void group_transpose_reduce(float (&v)[U], int lg) {
  constexpr int G = 1 << LOGG;
  int cnt = U;
#pragma unroll
  for (int mask = G >> 1; mask >= 1; mask >>= 1) {
    if (cnt > 1) {
      const int half = cnt >> 1;
      const bool up = (lg & mask) != 0;
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        if (i < half) {
          const float send = up ? v[i] : v[i + half];
          const float keep = up ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(FULL_MASK, send, mask);
        }
      }
      cnt = half;
    } else {
      v[0] += __shfl_xor_sync(FULL_MASK, v[0], mask);
    }
  }
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// LOGG >= 0: single output column, G = 2^LOGG known at compile time (transposing reduction).
// LOGG == -1: (N,H,F) operands, runtime G and `seg` lanes per head (segmented butterfly).
// SINGLE: the whole row fits one feature tile, so the destination row lives in registers.
template <int VEC, int CH, int LOGG, bool SINGLE, bool HUB, typename T = float>
__global__ void __launch_bounds__(kBlockThreads) sddmm_dot_kernel(const SddmmParams p) {
  constexpr int U = 8 / CH;
  const int G = LOGG >= 0 ? (1 << LOGG) : p.G;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  int n;
  group_work<HUB>(p, row, j0, n);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const int tile_cols = G * CH;
  const T* __restrict__ vrow = reinterpret_cast<const T*>(p.V) + row * (int64_t)p.D;
  const T* __restrict__ ubase = reinterpret_cast<const T*>(p.U);
  T* __restrict__ obase = reinterpret_cast<T*>(p.out);
  StageVec<T, VEC> vreg[CH];
  if constexpr (SINGLE) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = c * G + lg;
      if (n > 0 && vc < p.ncols) vreg[c] = ldg_stage<T, VEC>(vrow + vc * VEC);
    }
  }
  for (int off = 0; off < nmax; off += G) {
    const int m = min(max(n - off, 0), G);
    int my_c = 0, my_e = 0;
    if (lg < m) {
      const int64_t j = j0 + off + lg;
      my_c = __ldg(p.indices + j);
      my_e = p.eids ? __ldg(p.eids + j) : (int)j;
    }
    const int mmax = min(G, nmax - off);
    for (int t = 0; t < mmax; t += U) {
      int cc[U];
      float part[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
        part[u] = 0.f;
      }
      for (int tile0 = 0; tile0 < (SINGLE ? 1 : p.ncols); tile0 += tile_cols) {
        StageVec<T, VEC> xv[U][CH];
        StageVec<T, VEC> vv[CH];
        bool colv[CH];
        // phase 1: issue every load of this (edge batch, tile) before the first use
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int vc = tile0 + c * G + lg;
          colv[c] = vc < p.ncols;
          if constexpr (!SINGLE) {
            if (colv[c] && m > 0) vv[c] = ldg_stage<T, VEC>(vrow + vc * VEC);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c])
              xv[u][c] = ldg_stage<T, VEC>(ubase + (int64_t)cc[u] * p.D + (tile0 + c * G + lg) * VEC);
          }
        }
          // phase 2: multiply-accumulate
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
#pragma unroll
              for (int v = 0; v < VEC; ++v)
                part[u] = fmaf(xv[u][c].at(v), SINGLE ? vreg[c].at(v) : vv[c].at(v), part[u]);
            }
          }
        }
      }
      if constexpr (LOGG >= 0) {
        group_transpose_reduce<U, LOGG>(part, lg);
        constexpr int GG = 1 << LOGG;
        if constexpr (GG >= U) {
          const int u_l = lg >> (LOGG - ilog2(U));
          const int e = __shfl_sync(FULL_MASK, my_e, t + u_l, GG);
          if ((lg & (GG / U - 1)) == 0 && (t + u_l) < m) store_scalar_t<T>(obase + e, part[0]);
        } else {
          constexpr int CF = U / GG;
#pragma unroll
          for (int i = 0; i < CF; ++i) {
            const int u_l = lg * CF + i;
            const int e = __shfl_sync(FULL_MASK, my_e, t + u_l, GG);
            if ((t + u_l) < m) store_scalar_t<T>(obase + e, part[i]);
          }
        }
      } else {
        // segmented butterfly: seg lanes per head
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = __shfl_sync(FULL_MASK, my_e, t + u, G);
          for (int s = p.seg >> 1; s > 0; s >>= 1) part[u] += __shfl_xor_sync(FULL_MASK, part[u], s);
          if ((t + u) < m && (lg & (p.seg - 1)) == 0 && (lg / p.seg) < p.H)
            store_scalar_t<T>(obase + (int64_t)e * p.H + (lg / p.seg), part[u]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ elementwise, full shapes
template <int VEC, int CH, int OP, bool HUB>
__global__ void __launch_bounds__(kBlockThreads) sddmm_ew_kernel(const SddmmParams p) {
  constexpr int U = 8 / CH;
  constexpr bool USE_L = OP != DGLB_OP_COPY_RHS;
  constexpr bool USE_R = OP != DGLB_OP_COPY_LHS;
  const int G = p.G;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  int n;
  group_work<HUB>(p, row, j0, n);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const float* __restrict__ vrow = p.V + row * (int64_t)p.D;
  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    FVec<VEC> vreg[CH];
    bool colv[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      if (USE_R && colv[c] && n > 0) vreg[c] = ldg_vec<VEC>(vrow + vc * VEC);
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      int my_c = 0, my_e = 0;
      if (lg < m) {
        const int64_t j = j0 + off + lg;
        if (USE_L) my_c = __ldg(p.indices + j);
        my_e = p.eids ? __ldg(p.eids + j) : (int)j;
      }
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U], ee[U];
        FVec<VEC> xv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = USE_L ? __shfl_sync(FULL_MASK, my_c, t + u, G) : 0;
          ee[u] = __shfl_sync(FULL_MASK, my_e, t + u, G);
        }
        if constexpr (USE_L) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool valid = (t + u) < m;
#pragma unroll
            for (int c = 0; c < CH; ++c)
              if (valid && colv[c])
                xv[u][c] = ldg_vec<VEC>(p.U + (int64_t)cc[u] * p.D + (tile0 + c * G + lg) * VEC);
          }
        }
  #pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool valid = (t + u) < m;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if (valid && colv[c]) {
              FVec<VEC> o;
#pragma unroll
              for (int v = 0; v < VEC; ++v)
                o.v[v] = ew_op<OP>(USE_L ? xv[u][c].v[v] : 0.f, USE_R ? vreg[c].v[v] : 0.f);
              st_vec<VEC>(p.out + (int64_t)ee[u] * p.D + (tile0 + c * G + lg) * VEC, o);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ generic (any target / broadcast)

__global__ void __launch_bounds__(kBlockThreads) sddmm_generic_kernel(const GenericSddmmParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  const int64_t OL = p.b.out_len;
  if (idx >= p.nnz * OL) return;
  const int64_t pos = idx / OL;
  int64_t rem = idx - pos * OL;
  int64_t s, d, e;
  if (p.src) {
    e = pos; s = p.src[pos]; d = p.dst[pos];
  } else {
    // upper_bound(indptr, pos) - 1
    int64_t lo = 0, hi = p.n_rows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.indptr[mid + 1] <= pos) lo = mid + 1; else hi = mid;
    }
    d = lo; s = p.indices[pos]; e = p.eids ? p.eids[pos] : pos;
  }
  int64_t lk = 0, rk = 0, sl = 1, sr = 1;
  for (int dd = p.b.ndim - 1; dd >= 0; --dd) {
    const bool red_axis = (p.op == DGLB_OP_DOT && dd == p.b.ndim - 1);
    const int64_t od = p.b.out[dd];
    const int64_t i = rem % od;
    rem /= od;
    if (!red_axis) {
      lk += (p.b.lhs[dd] == 1 ? 0 : i) * sl;
      rk += (p.b.rhs[dd] == 1 ? 0 : i) * sr;
    }
    sl *= p.b.lhs[dd];
    sr *= p.b.rhs[dd];
  }
  const int64_t lid = p.lhs_target == DGLB_TARGET_U ? s : (p.lhs_target == DGLB_TARGET_E ? e : d);
  const int64_t rid = p.rhs_target == DGLB_TARGET_U ? s : (p.rhs_target == DGLB_TARGET_E ? e : d);
  const float* l = p.L ? p.L + lid * p.b.lhs_len + lk : nullptr;
  const float* r = p.R ? p.R + rid * p.b.rhs_len + rk : nullptr;
  float val;
  if (p.op == DGLB_OP_DOT) {
    val = 0.f;
    for (int64_t i = 0; i < p.reduce_size; ++i) val = __fadd_rn(val, __fmul_rn(__ldg(l + i), __ldg(r + i)));
  } else {
    const float lv = (p.op != DGLB_OP_COPY_RHS) ? __ldg(l) : 0.f;
    const float rv = (p.op != DGLB_OP_COPY_LHS) ? __ldg(r) : 0.f;
    switch (p.op) {
      case DGLB_OP_ADD: val = __fadd_rn(lv, rv); break;
      case DGLB_OP_SUB: val = __fsub_rn(lv, rv); break;
      case DGLB_OP_MUL: val = __fmul_rn(lv, rv); break;
      case DGLB_OP_DIV: val = __fdiv_rn(lv, rv); break;
      case DGLB_OP_COPY_LHS: val = lv; break;
      default: val = rv; break;
    }
  }
  p.out[e * OL + (idx - pos * OL)] = val;
}

int sddmm_generic_f32(const GenericSddmmParams& g, cudaStream_t stream) {
  const int64_t total = g.nnz * g.b.out_len;
  const int64_t blocks = (total + kBlockThreads - 1) / kBlockThreads;
  if (blocks == 0) return DGLB_OK;
  if (blocks > 0x7fffffffLL) { set_error("sddmm generic: problem too large"); return DGLB_E_UNSUPPORTED; }
  sddmm_generic_kernel<<<(unsigned)blocks, kBlockThreads, 0, stream>>>(g);
  DGLB_LAUNCH_CHECK("sddmm_generic_kernel");
  return DGLB_OK;
}

// ------------------------------------------------------------------ narrow rows, edge-parallel over the COO
// Rows of <= 8 floats (attention logits el[u] + er[v] with 1-8 heads, (N,1) scores, ...): one thread per EDGE.  src[e] / dst[e]
// and the (E, W) result are read / written fully coalesced in edge-id order -- no edge-id permutation at all, whatever the
// order the edges were created in -- and the two W-float gathers hit L2 (N * W * 4 bytes is a few MB).  The destination-
// major CSC kernel spends a lane group per row on these shapes, reads its indices 1-2 lanes wide and scatters its
// result through eids[]: reddit (N,1,1) u_add_v 0.19-0.26 ms there vs the ~0.04 ms its 140 MB take at the HBM rate.
template <int OP, int VEC>
__global__ void __launch_bounds__(kBlockThreads)
sddmm_coo_narrow_kernel(int64_t nnz, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                        const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ out, int W, int nvec,
                        int lhs_target, int rhs_target) {
  const int64_t e = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (e >= nnz) return;
  const int64_t s = __ldg(src + e), d = __ldg(dst + e);
  const int64_t lid = lhs_target == DGLB_TARGET_U ? s : (lhs_target == DGLB_TARGET_E ? e : d);
  const int64_t rid = rhs_target == DGLB_TARGET_U ? s : (rhs_target == DGLB_TARGET_E ? e : d);
  const float* l = OP != DGLB_OP_COPY_RHS ? L + lid * W : nullptr;
  const float* r = OP != DGLB_OP_COPY_LHS ? R + rid * W : nullptr;
  if constexpr (OP == DGLB_OP_DOT) {
    float acc = 0.f;
    for (int v = 0; v < nvec; ++v) {
      const FVec<VEC> lv = ldg_vec<VEC>(l + v * VEC), rv = ldg_vec<VEC>(r + v * VEC);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc = __fadd_rn(acc, __fmul_rn(lv.v[i], rv.v[i]));
    }
    out[e] = acc;
  } else {
    for (int v = 0; v < nvec; ++v) {
      FVec<VEC> lv, rv, o;
      if constexpr (OP != DGLB_OP_COPY_RHS) lv = ldg_vec<VEC>(l + v * VEC);
      if constexpr (OP != DGLB_OP_COPY_LHS) rv = ldg_vec<VEC>(r + v * VEC);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if constexpr (OP == DGLB_OP_COPY_LHS) o.v[i] = lv.v[i];
        else if constexpr (OP == DGLB_OP_COPY_RHS) o.v[i] = rv.v[i];
        else o.v[i] = ew_op<OP>(lv.v[i], rv.v[i]);
      }
      st_vec<VEC>(out + e * W + v * VEC, o);
    }
  }
}

template <int OP>
static int launch_coo_narrow(int64_t nnz, const int32_t* src, const int32_t* dst, const float* L, const float* R, float* out,
                             int W, int lt, int rt, cudaStream_t stream) {
  const uintptr_t al = (L ? reinterpret_cast<uintptr_t>(L) : 0) | (R ? reinterpret_cast<uintptr_t>(R) : 0) |
                       (OP == DGLB_OP_DOT ? 0 : reinterpret_cast<uintptr_t>(out));
  const int vec = (W % 4 == 0 && (al & 15) == 0) ? 4 : ((W % 2 == 0 && (al & 7) == 0) ? 2 : 1);
  const unsigned blocks = (unsigned)((nnz + kBlockThreads - 1) / kBlockThreads);
  if (vec == 4) sddmm_coo_narrow_kernel<OP, 4><<<blocks, kBlockThreads, 0, stream>>>(nnz, src, dst, L, R, out, W, W / 4, lt, rt);
  else if (vec == 2) sddmm_coo_narrow_kernel<OP, 2><<<blocks, kBlockThreads, 0, stream>>>(nnz, src, dst, L, R, out, W, W / 2, lt, rt);
  else sddmm_coo_narrow_kernel<OP, 1><<<blocks, kBlockThreads, 0, stream>>>(nnz, src, dst, L, R, out, W, W, lt, rt);
  DGLB_LAUNCH_CHECK("sddmm_coo_narrow_kernel");
  return DGLB_OK;
}

// DGLB_E_UNSUPPORTED (no error set) when the shapes are not "same trailing shape on both sides, <= 8 floats per row"
int sddmm_coo_narrow_f32(int op, int lhs_target, int rhs_target, int64_t nnz, const int32_t* src, const int32_t* dst,
                         const float* L, const float* R, const BcastShape& b, int64_t reduce_size, float* out,
                         cudaStream_t stream) {
  int64_t W;
  if (op == DGLB_OP_DOT) {
    if (b.ndim != 1 || b.out_len != 1) return DGLB_E_UNSUPPORTED;
    W = reduce_size;
  } else {
    if (op != DGLB_OP_COPY_LHS && op != DGLB_OP_COPY_RHS)
      for (int d = 0; d < b.ndim; ++d)
        if (b.lhs[d] != b.rhs[d]) return DGLB_E_UNSUPPORTED;
    W = b.out_len;
  }
  if (W < 1 || W > 8 || nnz >= (1LL << 31) * (int64_t)kBlockThreads) return DGLB_E_UNSUPPORTED;
  switch (op) {
    case DGLB_OP_ADD: return launch_coo_narrow<DGLB_OP_ADD>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_SUB: return launch_coo_narrow<DGLB_OP_SUB>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_MUL: return launch_coo_narrow<DGLB_OP_MUL>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_DIV: return launch_coo_narrow<DGLB_OP_DIV>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_COPY_LHS: return launch_coo_narrow<DGLB_OP_COPY_LHS>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_COPY_RHS: return launch_coo_narrow<DGLB_OP_COPY_RHS>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    case DGLB_OP_DOT: return launch_coo_narrow<DGLB_OP_DOT>(nnz, src, dst, L, R, out, (int)W, lhs_target, rhs_target, stream);
    default: return DGLB_E_UNSUPPORTED;
  }
}

// ------------------------------------------------------------------ dispatch
template <int VEC, int CH, int LOGG, bool SINGLE, typename T>
static int launch_dot(const SddmmParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks > 0 && !p.skip_rows) {
    sddmm_dot_kernel<VEC, CH, LOGG, SINGLE, false, T><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("sddmm_dot_kernel");
  }
  if (n_hub > 0) {
    sddmm_dot_kernel<VEC, CH, LOGG, SINGLE, true, T><<<n_hub, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("sddmm_dot_kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, typename T = float>
static int dispatch_dot(const SddmmParams& p, int ch, int n_hub, cudaStream_t stream) {
  const bool single = p.ncols <= p.G * ch;
  if (p.H > 1) return launch_dot<VEC, 1, -1, true, T>(p, n_hub, stream);  // caller guarantees ch == 1
  if (ch == 1) {
    switch (p.log2G) {
      case 0: return launch_dot<VEC, 1, 0, true, T>(p, n_hub, stream);
      case 1: return launch_dot<VEC, 1, 1, true, T>(p, n_hub, stream);
      case 2: return launch_dot<VEC, 1, 2, true, T>(p, n_hub, stream);
      case 3: return launch_dot<VEC, 1, 3, true, T>(p, n_hub, stream);
      case 4: return launch_dot<VEC, 1, 4, true, T>(p, n_hub, stream);
      default: return launch_dot<VEC, 1, 5, true, T>(p, n_hub, stream);
    }
  }
  if (ch == 2)
    return single ? launch_dot<VEC, 2, 5, true, T>(p, n_hub, stream) : launch_dot<VEC, 2, 5, false, T>(p, n_hub, stream);
  return single ? launch_dot<VEC, 4, 5, true, T>(p, n_hub, stream) : launch_dot<VEC, 4, 5, false, T>(p, n_hub, stream);
}

template <int VEC, int CH, int OP>
static int launch_ew(const SddmmParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  if (blocks > 0) {
    sddmm_ew_kernel<VEC, CH, OP, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("sddmm_ew_kernel");
  }
  if (n_hub > 0) {
    sddmm_ew_kernel<VEC, CH, OP, true><<<n_hub, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("sddmm_ew_kernel(hub)");
  }
  return DGLB_OK;
}

template <int OP>
static int dispatch_ew(const SddmmParams& p, int vec, int ch, int n_hub, cudaStream_t stream) {
#define DGLB_CASE(V, C) if (vec == V && ch == C) return launch_ew<V, C, OP>(p, n_hub, stream);
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  return DGLB_E_UNSUPPORTED;
}

// returns DGLB_E_UNSUPPORTED (without setting an error) when the vector path does not apply
int sddmm_csr_fast_f32(int op, int64_t n_dst, int64_t n_src, int64_t nnz, const int32_t* indptr,
                       const int32_t* indices, const int32_t* eids, const float* Uf, const float* Vf,
                       const BcastShape& b, int64_t reduce_size, float* out, const dglb_hub_t* hub,
                       cudaStream_t stream, int dtype) {
  if (dtype == DGLB_BF16 && op != DGLB_OP_DOT) return DGLB_E_UNSUPPORTED;
  if (b.lhs_len != b.rhs_len) return DGLB_E_UNSUPPORTED;
  for (int d = 0; d < b.ndim; ++d)
    if (b.lhs[d] != b.rhs[d]) return DGLB_E_UNSUPPORTED;
  const int64_t D = b.lhs_len;
  if (D <= 0 || D >= (1 << 30)) return DGLB_E_UNSUPPORTED;
  SddmmParams p;
  p.indptr = indptr; p.indices = indices; p.eids = eids; p.U = Uf; p.V = Vf; p.out = out;
  const bool use_hub = hub && hub->n_hub > 0 && hub->n_seg > 0 && hub->rows && hub->seg_ptr && hub->seg_hub &&
                       hub->seg_len > 0;
  p.row_order = hub ? hub->row_order : nullptr;
  p.hub_rows = use_hub ? hub->rows : nullptr;
  p.seg_ptr = use_hub ? hub->seg_ptr : nullptr;
  p.seg_hub = use_hub ? hub->seg_hub : nullptr;
  p.seg_len = use_hub ? hub->seg_len : 0;
  p.n_rows = n_dst; p.D = (int)D;
  p.hub_threshold = use_hub ? hub->threshold : INT32_MAX;
  p.skip_rows = 0;
  const int n_hub = use_hub ? hub->n_seg : 0;  // hub launches are sized by SEGMENTS
  if (op == DGLB_OP_DOT && b.out_len == 1) {
    // wide rows: whole-row bulk copies into a shared-memory ring (ring.cu); hub rows stay on the segmented path
    const int rc = ring_rows(true, dtype, n_dst, n_src, nnz, indptr, indices, eids, Uf, Vf, D, out, nullptr, 0,
                             p.hub_threshold, use_hub ? hub->light_indptr : nullptr, stream);
    if (rc == DGLB_OK) {
      if (!use_hub) return DGLB_OK;
      p.skip_rows = 1;
    } else if (rc != DGLB_E_UNSUPPORTED) {
      return rc;
    }
  }
  int vec = dtype == DGLB_BF16 ? 8 : 4;
  if (dtype == DGLB_BF16) {
    vec = min_int(pick_vec_bf16(D, Uf), pick_vec_bf16(D, Vf));
  } else {
    if (op != DGLB_OP_COPY_RHS) vec = min_int(vec, pick_vec(D, Uf));
    if (op != DGLB_OP_COPY_LHS) vec = min_int(vec, pick_vec(D, Vf));
  }
  if (op == DGLB_OP_DOT) {
    const int64_t H = b.out_len;
    if (H > 1) {
      // each head must be a power-of-two lane segment inside a single chunk
      while (vec > 1 && (reduce_size % vec)) vec >>= 1;
      const int64_t seg = reduce_size / vec;
      if ((seg & (seg - 1)) != 0 || seg * H > 32) return DGLB_E_UNSUPPORTED;
      int64_t cols = seg * H;
      if ((cols & (cols - 1)) != 0 && group_lanes(cols) < cols) return DGLB_E_UNSUPPORTED;
      p.seg = (int)seg;
    }
    p.H = (int)H;
  } else {
    vec = min_int(vec, pick_vec(D, out));
  }
  p.ncols = (int)(D / vec);
  p.G = group_lanes(p.ncols);
  p.log2G = 0;
  while ((1 << p.log2G) < p.G) ++p.log2G;
  if (op == DGLB_OP_DOT && p.H == 1) p.seg = p.G;
  const int per_lane = (p.ncols + p.G - 1) / p.G;
  const int ch = (op == DGLB_OP_DOT && dtype == DGLB_BF16) ? (per_lane >= 3 ? 4 : (per_lane >= 2 ? 2 : 1))
                                                           : (per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1));
  if (op == DGLB_OP_DOT) {
    if (p.H > 1 && ch != 1) return DGLB_E_UNSUPPORTED;
    if (dtype == DGLB_BF16) {
      if (vec == 8) return dispatch_dot<8, __nv_bfloat16>(p, ch, n_hub, stream);
      if (vec == 4) return dispatch_dot<4, __nv_bfloat16>(p, ch, n_hub, stream);
      if (vec == 2) return dispatch_dot<2, __nv_bfloat16>(p, ch, n_hub, stream);
      return dispatch_dot<1, __nv_bfloat16>(p, ch, n_hub, stream);
    }
    if (vec == 4) return dispatch_dot<4>(p, ch, n_hub, stream);
    if (vec == 2) return dispatch_dot<2>(p, ch, n_hub, stream);
    return dispatch_dot<1>(p, ch, n_hub, stream);
  }
  switch (op) {
    case DGLB_OP_ADD: return dispatch_ew<DGLB_OP_ADD>(p, vec, ch, n_hub, stream);
    case DGLB_OP_SUB: return dispatch_ew<DGLB_OP_SUB>(p, vec, ch, n_hub, stream);
    case DGLB_OP_MUL: return dispatch_ew<DGLB_OP_MUL>(p, vec, ch, n_hub, stream);
    case DGLB_OP_DIV: return dispatch_ew<DGLB_OP_DIV>(p, vec, ch, n_hub, stream);
    case DGLB_OP_COPY_LHS: return dispatch_ew<DGLB_OP_COPY_LHS>(p, vec, ch, n_hub, stream);
    case DGLB_OP_COPY_RHS: return dispatch_ew<DGLB_OP_COPY_RHS>(p, vec, ch, n_hub, stream);
    default: return DGLB_E_UNSUPPORTED;
  }
}

}  // namespace dglb
