// gat_fused.cu -- fused SDDMM -> leaky_relu -> edge_softmax -> (attention dropout) -> SpMM for
// GAT attention, forward and backward, hand-written for sm_100a.
//
// Replaces, for dgl.nn.pytorch.GATConv (main_dgl_arxiv_gat.py:9; written-out twin
// main_pyg_arxiv_gat.py:98-111), upstream DGL v0.6.1's chain of ~7 sparse + 3 elementwise
// launches forward and ~6 + 5 backward and its per-edge (E,H) intermediates.
//
// All three kernels share one scheme (a group of G lanes per row, as in spmm.cu):
//   * "owner" phase, once per batch of G edges: lane j loads the neighbour id of edge j and the H
//     per-head scalars it needs (el[src,:] forward / dst pass, the 16-byte destination record in the
//     src pass), computes the attention weight(s) of ITS edge -- one exp and one divide per
//     (edge, head), no redundancy -- and parks them in shared memory;
//   * gather phase: every lane streams its 128-bit chunks of the neighbour rows (U*CH independent
//     loads in flight) and picks the weight of its own head with ONE shared-memory load per chunk
//     (a broadcast read: no shuffles, no select chains, no per-lane exp).
//   Round-1 history (profiles/r01_notes.md): v1 broadcast the weights with H shuffles + CH*H selects
//   per edge, v2 recomputed exp/div in every lane; both were instruction-bound at 2x the time of a
//   plain u_mul_e_sum.  This version issues ~3x fewer instructions per edge.
// Forward (CSC): per-head max and sum of exp are reduced inside the kernel (group shuffles; rows of
//   <= G edges keep their logits in registers), following upstream's edge_softmax formula
//   exp(e - max) / sum; row_max / row_sum (N,H) are saved; nothing per-edge is written.
// Backward, scores recomputed from (el, er, row_max, row_sum):
//   dst pass (CSC): per-lane partials of S1 = sum_j a_j dd_j and S2 = sum_j a_j g_j dd_j
//      (dd_j = drop_j <ft[src_j,h,:], dZ[v,h,:]>, g = lrelu') are accumulated over the row's edges
//      and reduced across lanes ONCE per row (the sums are linear in the per-lane partial dots);
//      S3 = sum_j a_j g_j comes from the owner lanes;  grad_er = S2 - S1*S3;  writes the 16-byte
//      record row_pack[v,h] = {er, max, sum, S1}.
//   src pass (CSR): grad_ft[u] = sum a*drop*dZ[v];  grad_el[u] = sum a g (dd - S1[v]).
// Hub rows (more than hub_threshold edges), when the caller hands over the segment lists and a workspace
//   (dglb_hub_t): every SEGMENT of <= seg_len edges is processed by a group exactly like an ordinary row
//   (same kernels, MODE = GAT_SEG) and writes PARTIAL results -- all three passes are linear in the edges once
//   the row's max / sum are known -- which small combine kernels fold in segment order (deterministic).
//   The forward statistics of hub rows come from two small kernels first (per-segment max / sum, rescaled
//   and combined per row).  Without a workspace: one CTA per hub row (MODE = GAT_HUB_CTA), the groups split
//   the edges and meet in shared memory.
#include <cstdlib>

#include "kernels.cuh"

namespace dglb {

constexpr int kMaxHeads = 8;

enum : int { GAT_ROWS = 0, GAT_HUB_CTA = 1, GAT_SEG = 2 };

// counter-based dropout: keep iff hash(seed, edge*H + h) maps to u >= p
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float drop_factor(const GatParams& p, int64_t e, int h) {
  const uint64_t ctr = (uint64_t)e * (uint64_t)p.H + (uint64_t)h;
  uint32_t x = mix32((uint32_t)ctr ^ p.seed_lo);
  x = mix32(x + (uint32_t)(ctr >> 32) * 0x9e3779b9u + p.seed_hi);
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);  // [0,1)
  return u < p.drop_p ? 0.f : p.drop_scale;
}

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

template <int HT>
__device__ __forceinline__ void group_allreduce_sum(float (&v)[HT], int G) {
#pragma unroll
  for (int h = 0; h < HT; ++h)
    for (int s = G >> 1; s > 0; s >>= 1) v[h] += __shfl_xor_sync(FULL_MASK, v[h], s);
}

// Combine per-group values (uniform within a group) across the CTA's groups, fixed order.
template <int HT>
__device__ __forceinline__ void cta_allreduce_sum(float (&v)[HT], float* s_buf, int gidx, int lg, int n_groups) {
  __syncthreads();
  if (lg == 0) {
#pragma unroll
    for (int h = 0; h < HT; ++h) s_buf[gidx * HT + h] = v[h];
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    float r = s_buf[h];
    for (int g = 1; g < n_groups; ++g) r += s_buf[g * HT + h];
    v[h] = r;
  }
}

// row / slice / segment of the calling group (same scheme as spmm.cu / sddmm.cu)
template <int MODE>
__device__ __forceinline__ void gat_group_work(const GatParams& p, int64_t& row, bool& active, int64_t& j0,
                                               int& n, int& gidx, int& n_groups, int& seg) {
  n_groups = kBlockThreads >> p.log2G;
  gidx = threadIdx.x >> p.log2G;
  seg = 0;
  if constexpr (MODE == GAT_ROWS) {
    row = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> p.log2G;
    active = row < p.n_rows;
    j0 = 0; n = 0;
    if (active) {
      const int s = __ldg(p.indptr + row);
      const int d = __ldg(p.indptr + row + 1) - s;
      if (d > p.hub_threshold) active = false;
      else { j0 = s; n = d; }
    }
    if (!active) row = 0;
  } else if constexpr (MODE == GAT_SEG) {
    const int64_t sg = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> p.log2G;
    active = sg < p.n_seg;
    row = 0; j0 = 0; n = 0;
    if (active) {
      seg = (int)sg;
      const int hub = __ldg(p.seg_hub + seg);
      row = __ldg(p.hub_rows + hub);
      const int k = seg - __ldg(p.seg_ptr + hub);
      const int s = __ldg(p.indptr + row) + k * p.seg_len;
      j0 = s;
      n = min(p.seg_len, __ldg(p.indptr + row + 1) - s);
    }
  } else {
    row = p.hub_rows[blockIdx.x];
    active = true;
    const int s = __ldg(p.indptr + row);
    const int d = __ldg(p.indptr + row + 1) - s;
    const int per = (d + n_groups - 1) / n_groups;
    const int b = min(gidx * per, d);
    j0 = (int64_t)s + b;
    n = min(per, d - b);
  }
}

// ------------------------------------------------------------------ shared helpers
template <int HT>
__device__ __forceinline__ void group_allreduce_max(float (&v)[HT], int G) {
#pragma unroll
  for (int h = 0; h < HT; ++h)
    for (int s = G >> 1; s > 0; s >>= 1) v[h] = fmaxf(v[h], __shfl_xor_sync(FULL_MASK, v[h], s));
}
template <int HT>
__device__ __forceinline__ void cta_allreduce_max(float (&v)[HT], float* s_buf, int gidx, int lg, int n_groups) {
  __syncthreads();
  if (lg == 0) {
#pragma unroll
    for (int h = 0; h < HT; ++h) s_buf[gidx * HT + h] = v[h];
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    float r = s_buf[h];
    for (int g = 1; g < n_groups; ++g) r = fmaxf(r, s_buf[g * HT + h]);
    v[h] = r;
  }
}

// stride (in floats) between the weight slots of consecutive groups: the +HT skews the groups of a
// warp onto different banks
template <int HT, int NW>
__device__ __forceinline__ int group_stride(int G) { return (G * HT + (G >= 8 ? HT : 0)) * NW; }

// shared epilogue: write a feature tile to `out_row` (the row of the output, or the segment's row of the
// workspace) or combine the CTA's groups (CTA-per-hub-row kernel)
template <int VEC, int CH, int MODE>
__device__ __forceinline__ void store_feat_tile(const GatParams& p, float (&acc)[CH][VEC], const bool (&colv)[CH],
                                                const int (&k)[CH], float* out_row, int64_t row, bool active,
                                                int tile0, int lg, int gidx, int n_groups, float* s_val) {
  if constexpr (MODE != GAT_HUB_CTA) {
    if (active) {
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (colv[c]) {
          FVec<VEC> o;
#pragma unroll
          for (int v = 0; v < VEC; ++v) o.v[v] = acc[c][v];
          st_vec<VEC>(out_row + k[c], o);
        }
    }
  } else {
    const int G = p.G;
    const int tile_elems = G * CH * VEC;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_val[gidx * tile_elems + (c * G + lg) * VEC + v] = acc[c][v];
    __syncthreads();
    for (int el = threadIdx.x; el < tile_elems; el += kBlockThreads) {
      const int kk = tile0 * VEC + el;
      if (kk < p.D) {
        float a = s_val[el];
        for (int g = 1; g < n_groups; ++g) a = __fadd_rn(a, s_val[g * tile_elems + el]);
        p.out_feat[row * (int64_t)p.D + kk] = a;
      }
    }
  }
}

// ------------------------------------------------------------------ forward
template <int VEC, int CH, int HT, int UT, int MODE>
__global__ void __launch_bounds__(kBlockThreads, UT == 4 ? 3 : 2) gat_fwd_kernel(const GatParams p) {
  constexpr int U = UT / CH > 0 ? UT / CH : 1;
  constexpr bool HUB = MODE == GAT_HUB_CTA;
  constexpr bool SEG = MODE == GAT_SEG;
  __shared__ __align__(16) float s_w[(kBlockThreads + 32) * HT];  // [group][edge slot][head] weights
  extern __shared__ __align__(16) unsigned char smem_raw[];       // hub rows only
  float* s_buf = reinterpret_cast<float*>(smem_raw);
  const int G = p.G, H = p.H;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  bool active;
  int n, gidx, n_groups, seg;
  gat_group_work<MODE>(p, row, active, j0, n, gidx, n_groups, seg);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const bool single = !SEG && nmax <= G;  // every row of this warp fits one batch: logits stay in registers
  const bool live = active || HUB;
  const bool need_e = p.drop_p > 0.f || p.edge_scores != nullptr;
  const bool prefetch = !SEG && p.prefetch;
  float* out_row = SEG ? p.ws_feat + (int64_t)seg * p.D : p.out_feat + row * (int64_t)p.D;
  float* my_w = s_w + gidx * group_stride<HT, 1>(G);

  float er_h[HT], mx[HT], sm[HT], e_reg[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    er_h[h] = (h < H && live) ? __ldg(p.er + row * H + h) : 0.f;
    mx[h] = -INFINITY; sm[h] = 0.f; e_reg[h] = -INFINITY;
  }
  // staging registers of the gather loop live at function scope: the first batch of source rows is
  // requested BEFORE the softmax statistics (it only needs the neighbour ids), so its latency
  // overlaps the idx -> el -> reduce chain below
  FVec<VEC> xv[U][CH];
  const int m_first = min(n, G);
  if constexpr (SEG) {
    // segment of a hub row: the row's statistics were combined from the per-segment ones beforehand
    if (active) {
#pragma unroll
      for (int h = 0; h < HT; ++h)
        if (h < H) { mx[h] = __ldg(p.row_max + row * H + h); sm[h] = __ldg(p.row_sum + row * H + h); }
    }
  } else {
  // ---- statistics 1: per-head max of the logits
  for (int off = 0; off < nmax; off += G) {
    const bool valid = off + lg < n;
    const int c = valid ? __ldg(p.indices + j0 + off + lg) : 0;
    if (off == 0 && prefetch) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cu = __shfl_sync(FULL_MASK, c, u, G);
#pragma unroll
        for (int cch = 0; cch < CH; ++cch) {
          const int vc = cch * G + lg;
          if (u < m_first && vc < p.ncols) xv[u][cch] = ldg_vec<VEC>(p.ft + (int64_t)cu * p.D + vc * VEC);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const float e = (valid && h < H) ? lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er_h[h]), p.slope)
                                       : -INFINITY;
      e_reg[h] = e;
      mx[h] = fmaxf(mx[h], e);
    }
  }
  group_allreduce_max<HT>(mx, G);
  if constexpr (HUB) cta_allreduce_max<HT>(mx, s_buf, gidx, lg, n_groups);
  // ---- statistics 2: per-head sum of exp(e - max)
  for (int off = 0; off < nmax; off += G) {
    const bool valid = off + lg < n;
    const int c = (valid && !single) ? __ldg(p.indices + j0 + off + lg) : 0;
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      if (valid && h < H) {
        const float e = single ? e_reg[h]
                               : lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er_h[h]), p.slope);
        const float ex = expf(__fsub_rn(e, mx[h]));
        sm[h] += ex;
        if (single) e_reg[h] = ex;  // single-batch rows keep exp(e - max) for the weights
      }
    }
  }
  group_allreduce_sum<HT>(sm, G);
  if constexpr (HUB) cta_allreduce_sum<HT>(sm, s_buf, gidx, lg, n_groups);
  if (!SEG && active && lg < H && (!HUB || gidx == 0)) {
#pragma unroll
    for (int h = 0; h < HT; ++h)
      if (h == lg) { p.out_h0[row * H + h] = mx[h]; p.out_h1[row * H + h] = sm[h]; }
  }

  }

  // ---- weighted gather of the source rows
  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    bool colv[CH];
    int k[CH], hk[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = colv[c] ? k[c] / p.F : 0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      // owner phase: lane lg computes the weights of edge off+lg and parks them in shared memory
      int my_c = 0;
      if (lg < m) {
        my_c = __ldg(p.indices + j0 + off + lg);
        const int64_t my_e = need_e ? (p.eids ? (int64_t)__ldg(p.eids + j0 + off + lg) : (j0 + off + lg)) : 0;
#pragma unroll
        for (int h = 0; h < HT; ++h) {
          if (h < H) {
            const float ex = single ? e_reg[h]
                                    : expf(__fsub_rn(lrelu(__fadd_rn(__ldg(p.el + (int64_t)my_c * H + h), er_h[h]),
                                                           p.slope), mx[h]));
            float a = __fdiv_rn(ex, sm[h]);
            if (p.edge_scores != nullptr && tile0 == 0) p.edge_scores[my_e * H + h] = a;
            if (p.drop_p > 0.f) a *= drop_factor(p, my_e, h);
            my_w[lg * HT + h] = a;
          }
        }
      }
      __syncwarp();
      const int mmax = min(G, nmax - off);
      int t_begin = 0;
      if (tile0 == 0 && off == 0 && prefetch) {
        // first batch of the row: its source rows were requested before the statistics
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if (u < m && colv[c]) {
              const float w0 = my_w[u * HT + hk[c]];
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], w0));
            }
          }
        }
        t_begin = U;
      }
      for (int t = t_begin; t < mmax; t += U) {
        int cc[U];
        float w[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              xv[u][c] = ldg_vec<VEC>(p.ft + (int64_t)cc[u] * p.D + k[c]);
              w[u][c] = my_w[(t + u) * HT + hk[c]];
            }
          }
        }
  #pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], w[u][c]));
            }
          }
        }
      }
      __syncwarp();  // the next batch overwrites the weight slots
    }
    store_feat_tile<VEC, CH, MODE>(p, acc, colv, k, out_row, row, active, tile0, lg, gidx, n_groups, s_buf + n_groups * HT);
  }
}

// ------------------------------------------------------------------ backward
// SRC_PASS = false: CSC over dst rows v.  neighbour = src u: gathers ft[u] (owner: el[u,:]); own row: dZ[v].
//            outputs row_pack[v,h] = {er, max, sum, S1}, grad_er[v,h].
// SRC_PASS = true : CSR over src rows u.  neighbour = dst v: gathers dZ[v] (owner: row_pack[v,:]); own row: ft[u].
//            outputs grad_ft[u,:], grad_el[u,h].
template <int VEC, int CH, int HT, int UT, bool SRC_PASS, int MODE>
__global__ void __launch_bounds__(kBlockThreads, UT == 4 ? 3 : 2) gat_bwd_kernel(const GatParams p) {
  constexpr int U = UT / CH > 0 ? UT / CH : 1;
  constexpr bool HUB = MODE == GAT_HUB_CTA;
  constexpr bool SEG = MODE == GAT_SEG;
  __shared__ __align__(16) float s_w[(kBlockThreads + 32) * HT * 2];  // [group][edge slot][head]{a*drop, a*drop*g}
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_buf = reinterpret_cast<float*>(smem_raw);
  const int G = p.G, H = p.H;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  bool active;
  int n, gidx, n_groups, seg;
  gat_group_work<MODE>(p, row, active, j0, n, gidx, n_groups, seg);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const bool live = active || HUB;
  const bool need_e = p.drop_p > 0.f;
  float* out_row = SEG ? p.ws_feat + (int64_t)seg * p.D : p.out_feat + row * (int64_t)p.D;
  float2* my_w = reinterpret_cast<float2*>(s_w + gidx * group_stride<HT, 2>(G));

  const float* __restrict__ own_feat = (SRC_PASS ? p.ft : p.dZ) + row * (int64_t)p.D;
  const float* __restrict__ nb_feat = SRC_PASS ? p.dZ : p.ft;

  // per-head constants of the own row (owner phase): dst pass er/max/sum, src pass el
  float own0[HT], own1[HT], own2[HT];
  // per-head totals: tot1/tot2 per lane until the once-per-row reduction; tot3 from the owner lanes
  float tot1[HT], tot2[HT], tot3[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    tot1[h] = tot2[h] = tot3[h] = 0.f;
    own0[h] = own1[h] = 0.f; own2[h] = 1.f;
    if (live && h < H) {
      if constexpr (!SRC_PASS) {
        own0[h] = __ldg(p.er + row * H + h);
        own1[h] = __ldg(p.row_max + row * H + h);
        own2[h] = __ldg(p.row_sum + row * H + h);
      } else {
        own0[h] = __ldg(p.el + row * H + h);
      }
    }
  }

  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    FVec<VEC> ownv[CH];
    bool colv[CH];
    int k[CH], hk[CH];
    float p1[CH], p2[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = colv[c] ? k[c] / p.F : 0;
      p1[c] = p2[c] = 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { acc[c][v] = 0.f; ownv[c].v[v] = 0.f; }
      if (live && colv[c]) ownv[c] = ldg_vec<VEC>(own_feat + k[c]);
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      int my_c = 0;
      if (lg < m) {
        my_c = __ldg(p.indices + j0 + off + lg);
        const int64_t my_e = need_e ? (p.eids ? (int64_t)__ldg(p.eids + j0 + off + lg) : (j0 + off + lg)) : 0;
#pragma unroll
        for (int h = 0; h < HT; ++h) {
          if (h < H) {
            float x, mxv, smv, s1v = 0.f;
            if constexpr (!SRC_PASS) {
              x = __fadd_rn(__ldg(p.el + (int64_t)my_c * H + h), own0[h]);
              mxv = own1[h]; smv = own2[h];
            } else {
              const float4 pk = __ldg(p.pack + (int64_t)my_c * H + h);  // {er, max, sum, S1} of the destination
              x = __fadd_rn(own0[h], pk.x);
              mxv = pk.y; smv = pk.z; s1v = pk.w;
            }
            const float a = __fdiv_rn(expf(__fsub_rn(lrelu(x, p.slope), mxv)), smv);
            const float g = x > 0.f ? 1.f : p.slope;
            if (tile0 == 0) tot3[h] = fmaf(a * g, SRC_PASS ? s1v : 1.f, tot3[h]);  // S3 = sum a g | T = sum a g S1
            const float ad = need_e ? a * drop_factor(p, my_e, h) : a;
            my_w[lg * HT + h] = make_float2(ad, ad * g);
          }
        }
      }
      __syncwarp();
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U];
        FVec<VEC> xv[U][CH];
        float2 w[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              xv[u][c] = ldg_vec<VEC>(nb_feat + (int64_t)cc[u] * p.D + k[c]);
              w[u][c] = my_w[(t + u) * HT + hk[c]];
            }
          }
        }
  #pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              float dot = 0.f;
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                dot = fmaf(xv[u][c].v[v], ownv[c].v[v], dot);
                if constexpr (SRC_PASS) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], w[u][c].x));
              }
              p1[c] = fmaf(w[u][c].x, dot, p1[c]);
              p2[c] = fmaf(w[u][c].y, dot, p2[c]);
            }
          }
        }
      }
      __syncwarp();
    }
    // fold this tile's per-chunk partials into per-head totals (still per lane)
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int h = 0; h < HT; ++h)
        if (colv[c] && hk[c] == h) { tot1[h] += p1[c]; tot2[h] += p2[c]; }
    if constexpr (SRC_PASS) {
      store_feat_tile<VEC, CH, MODE>(p, acc, colv, k, out_row, row, active, tile0, lg, gidx, n_groups, s_buf + n_groups * HT);
    }
  }
  // ---- once-per-row reductions
  group_allreduce_sum<HT>(tot1, G);
  group_allreduce_sum<HT>(tot2, G);
  group_allreduce_sum<HT>(tot3, G);
  if constexpr (HUB) {
    cta_allreduce_sum<HT>(tot1, s_buf, gidx, lg, n_groups);
    cta_allreduce_sum<HT>(tot2, s_buf, gidx, lg, n_groups);
    cta_allreduce_sum<HT>(tot3, s_buf, gidx, lg, n_groups);
  }
  if constexpr (SEG) {
    // partial sums of this segment; gat_seg_tot_combine_kernel folds a row's segments in order
    if (active && lg < H) {
#pragma unroll
      for (int h = 0; h < HT; ++h)
        if (h == lg) {
          float* w = p.ws_tot + ((int64_t)seg * H + h) * 4;
          w[0] = tot1[h]; w[1] = tot2[h]; w[2] = tot3[h];
        }
    }
    return;
  }
  if (active && lg < H && (!HUB || gidx == 0)) {
#pragma unroll
    for (int h = 0; h < HT; ++h)
      if (h == lg) {
        if constexpr (!SRC_PASS) {
          p.out_pack[row * H + h] = make_float4(own0[h], own1[h], own2[h], tot1[h]);
          p.out_h0[row * H + h] = __fsub_rn(tot2[h], tot1[h] * tot3[h]);  // grad_er = S2 - S1*S3
        } else {
          p.out_h0[row * H + h] = __fsub_rn(tot2[h], tot3[h]);            // grad_el = sum a g dd - sum a g S1
        }
      }
  }
}

// ------------------------------------------------------------------ hub rows by segments: small kernels
// forward statistics of one segment: a warp per segment, lanes = (edge slot, head) as in edge_softmax.cu;
// ws_tot[seg, h] = {max_s, sum_s = sum exp(e - max_s)}
__global__ void __launch_bounds__(kBlockThreads) gat_seg_stats_kernel(const GatParams p) {
  const int lane = threadIdx.x & 31;
  const int seg = (int)(((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> 5);
  if (seg >= p.n_seg) return;
  const int H = p.H;
  const int h = lane & (p.HP - 1);
  const bool hv = h < H;
  const int slot = lane >> p.log2HP, nslots = 32 >> p.log2HP;
  const int hub = __ldg(p.seg_hub + seg);
  const int64_t row = __ldg(p.hub_rows + hub);
  const int begin = __ldg(p.indptr + row) + (seg - __ldg(p.seg_ptr + hub)) * p.seg_len;
  const int n = min(p.seg_len, __ldg(p.indptr + row + 1) - begin);
  const float er = hv ? __ldg(p.er + row * H + h) : 0.f;
  float mx = -INFINITY;
#pragma unroll 4
  for (int i = slot; i < n; i += nslots) {
    const int c = __ldg(p.indices + begin + i);
    if (hv) mx = fmaxf(mx, lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er), p.slope));
  }
  for (int s = 16; s >= p.HP; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, s));
  float sum = 0.f;
#pragma unroll 4
  for (int i = slot; i < n; i += nslots) {
    const int c = __ldg(p.indices + begin + i);
    if (hv) sum += expf(__fsub_rn(lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er), p.slope), mx));
  }
  for (int s = 16; s >= p.HP; s >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, s);
  if (slot == 0 && hv) {
    float* w = p.ws_tot + ((int64_t)seg * H + h) * 4;
    w[0] = mx; w[1] = sum;
  }
}

// per hub row and head: max = max_s max_s, sum = sum_s sum_s * exp(max_s - max), segment order
__global__ void __launch_bounds__(kBlockThreads) gat_seg_stats_combine_kernel(const GatParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (idx >= (int64_t)p.n_hub * p.H) return;
  const int hub = (int)(idx / p.H), h = (int)(idx - (int64_t)hub * p.H);
  const int s0 = __ldg(p.seg_ptr + hub), s1 = __ldg(p.seg_ptr + hub + 1);
  const int64_t row = __ldg(p.hub_rows + hub);
  float mx = -INFINITY;
  for (int sg = s0; sg < s1; ++sg) mx = fmaxf(mx, p.ws_tot[((int64_t)sg * p.H + h) * 4]);
  float sum = 0.f;
  for (int sg = s0; sg < s1; ++sg) {
    const float* w = p.ws_tot + ((int64_t)sg * p.H + h) * 4;
    sum += w[1] * expf(__fsub_rn(w[0], mx));
  }
  p.out_h0[row * p.H + h] = mx;
  p.out_h1[row * p.H + h] = sum;
}

// out_feat[row, k] = sum over the row's segments of ws_feat[seg, k], segment order
__global__ void __launch_bounds__(kBlockThreads) gat_seg_feat_combine_kernel(const GatParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (idx >= (int64_t)p.n_hub * p.D) return;
  const int hub = (int)(idx / p.D), k = (int)(idx - (int64_t)hub * p.D);
  const int s0 = __ldg(p.seg_ptr + hub), s1 = __ldg(p.seg_ptr + hub + 1);
  // a 20 000-edge row has ~160 segments: fetch 8 partials at a time (independent loads), add them in order
  float a = 0.f;
  for (int sg = s0; sg < s1; sg += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (sg + i < s1) ? p.ws_feat[(int64_t)(sg + i) * p.D + k] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (sg + i < s1) a = (sg + i == s0) ? v[i] : __fadd_rn(a, v[i]);
  }
  p.out_feat[(int64_t)__ldg(p.hub_rows + hub) * p.D + k] = a;
}

// backward totals of a hub row from its segments' partials (same closing formulas as gat_bwd_kernel)
template <bool SRC_PASS>
__global__ void __launch_bounds__(kBlockThreads) gat_seg_tot_combine_kernel(const GatParams p) {
  const int64_t idx = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
  if (idx >= (int64_t)p.n_hub * p.H) return;
  const int hub = (int)(idx / p.H), h = (int)(idx - (int64_t)hub * p.H);
  const int s0 = __ldg(p.seg_ptr + hub), s1 = __ldg(p.seg_ptr + hub + 1);
  const int64_t row = __ldg(p.hub_rows + hub);
  float t1 = 0.f, t2 = 0.f, t3 = 0.f;
  for (int sg = s0; sg < s1; sg += 4) {
    float v[4][3];  // (ws_tot is only 4-byte aligned when H*F is odd: scalar loads)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* w = p.ws_tot + ((int64_t)(sg + i) * p.H + h) * 4;
      const bool ok = sg + i < s1;
      v[i][0] = ok ? w[0] : 0.f; v[i][1] = ok ? w[1] : 0.f; v[i][2] = ok ? w[2] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { t1 += v[i][0]; t2 += v[i][1]; t3 += v[i][2]; }
  }
  if constexpr (!SRC_PASS) {
    p.out_pack[row * p.H + h] = make_float4(__ldg(p.er + row * p.H + h), __ldg(p.row_max + row * p.H + h),
                                            __ldg(p.row_sum + row * p.H + h), t1);
    p.out_h0[row * p.H + h] = __fsub_rn(t2, t1 * t3);
  } else {
    p.out_h0[row * p.H + h] = __fsub_rn(t2, t3);
  }
}

static unsigned blocks_for(int64_t items) { return (unsigned)((items + kBlockThreads - 1) / kBlockThreads); }

// ------------------------------------------------------------------ dispatch
static int gat_geometry(GatParams& p, int64_t H, int64_t F, const void* a0, const void* a1, const void* a2,
                        int* vec_out, int* ch_out, int* ht_out) {
  if (H < 1 || H > kMaxHeads || F < 1 || H * F >= (1 << 30)) return DGLB_E_UNSUPPORTED;
  p.H = (int)H; p.F = (int)F; p.D = (int)(H * F);
  int vec = 4;
  const void* ptrs[3] = {a0, a1, a2};
  for (const void* q : ptrs)
    if (q) vec = min_int(vec, pick_vec(p.D, q));
  while (vec > 1 && (F % vec)) vec >>= 1;
  p.ncols = p.D / vec;
  p.G = group_lanes(p.ncols);
  if (p.G < H) p.G = group_lanes(H);  // lanes 0..H-1 write the per-head outputs
  p.log2G = 0;
  while ((1 << p.log2G) < p.G) ++p.log2G;
  const int per_lane = (p.ncols + p.G - 1) / p.G;
  *ch_out = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
  *vec_out = vec;
  int ht = 1;
  while (ht < H) ht <<= 1;
  *ht_out = ht;
  p.HP = ht; p.log2HP = 0;
  while ((1 << p.log2HP) < p.HP) ++p.log2HP;
  return DGLB_OK;
}

static size_t gat_hub_smem(const GatParams& p, int vec, int ch, int ht) {
  const int n_groups = kBlockThreads / p.G;
  return sizeof(float) * ((size_t)n_groups * ht + (size_t)kBlockThreads * ch * vec);
}

// gathers in flight per lane (U*CH): 4 (<= 80 registers, 3 CTAs/SM).  The 8-deep variant (2 CTAs/SM) was
// measured slower on every GAT shape once the kernels were held to 3 CTAs/SM and is no longer built.
constexpr int kGatUT = 4;

template <int VEC, int CH, int HT, int UT>
static int launch_gat_fwd_ut(const GatParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  gat_fwd_kernel<VEC, CH, HT, UT, GAT_ROWS><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("gat_fwd_kernel");
  if (n_hub > 0 && p.n_seg > 0) {
    gat_seg_stats_kernel<<<blocks_for((int64_t)p.n_seg * 32), kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_seg_stats_kernel");
    gat_seg_stats_combine_kernel<<<blocks_for((int64_t)p.n_hub * p.H), kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_seg_stats_combine_kernel");
    const int64_t sblocks = (p.n_seg + rows_per_block - 1) / rows_per_block;
    gat_fwd_kernel<VEC, CH, HT, UT, GAT_SEG><<<(unsigned)sblocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_fwd_kernel(seg)");
    gat_seg_feat_combine_kernel<<<blocks_for((int64_t)p.n_hub * p.D), kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_seg_feat_combine_kernel");
  } else if (n_hub > 0) {
    gat_fwd_kernel<VEC, CH, HT, UT, GAT_HUB_CTA><<<n_hub, kBlockThreads, gat_hub_smem(p, VEC, CH, HT), stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_fwd_kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, int CH, int HT>
static int launch_gat_fwd(const GatParams& p, int n_hub, cudaStream_t stream) {
  return launch_gat_fwd_ut<VEC, CH, HT, kGatUT>(p, n_hub, stream);
}

template <int VEC, int CH, int HT, int UT>
static int launch_gat_bwd_ut(int which, const GatParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = gat_hub_smem(p, VEC, CH, HT);
  if (which == 1) gat_bwd_kernel<VEC, CH, HT, UT, false, GAT_ROWS><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  else gat_bwd_kernel<VEC, CH, HT, UT, true, GAT_ROWS><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("gat_bwd_kernel");
  if (n_hub > 0 && p.n_seg > 0) {
    const int64_t sblocks = (p.n_seg + rows_per_block - 1) / rows_per_block;
    if (which == 1) {
      gat_bwd_kernel<VEC, CH, HT, UT, false, GAT_SEG><<<(unsigned)sblocks, kBlockThreads, 0, stream>>>(p);
      DGLB_LAUNCH_CHECK("gat_bwd_kernel(seg)");
      gat_seg_tot_combine_kernel<false><<<blocks_for((int64_t)p.n_hub * p.H), kBlockThreads, 0, stream>>>(p);
    } else {
      gat_bwd_kernel<VEC, CH, HT, UT, true, GAT_SEG><<<(unsigned)sblocks, kBlockThreads, 0, stream>>>(p);
      DGLB_LAUNCH_CHECK("gat_bwd_kernel(seg)");
      gat_seg_feat_combine_kernel<<<blocks_for((int64_t)p.n_hub * p.D), kBlockThreads, 0, stream>>>(p);
      DGLB_LAUNCH_CHECK("gat_seg_feat_combine_kernel");
      gat_seg_tot_combine_kernel<true><<<blocks_for((int64_t)p.n_hub * p.H), kBlockThreads, 0, stream>>>(p);
    }
    DGLB_LAUNCH_CHECK("gat_seg_tot_combine_kernel");
  } else if (n_hub > 0) {
    if (which == 1) gat_bwd_kernel<VEC, CH, HT, UT, false, GAT_HUB_CTA><<<n_hub, kBlockThreads, smem, stream>>>(p);
    else gat_bwd_kernel<VEC, CH, HT, UT, true, GAT_HUB_CTA><<<n_hub, kBlockThreads, smem, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_bwd_kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, int CH, int HT>
static int launch_gat_bwd(int which, const GatParams& p, int n_hub, cudaStream_t stream) {
  return launch_gat_bwd_ut<VEC, CH, HT, kGatUT>(which, p, n_hub, stream);
}

template <int VEC, int CH>
static int dispatch_gat(int which, const GatParams& p, int ht, int n_hub, cudaStream_t stream) {
  if (which == 0) {
    switch (ht) {
      case 1: return launch_gat_fwd<VEC, CH, 1>(p, n_hub, stream);
      case 2: return launch_gat_fwd<VEC, CH, 2>(p, n_hub, stream);
      case 4: return launch_gat_fwd<VEC, CH, 4>(p, n_hub, stream);
      default: return launch_gat_fwd<VEC, CH, 8>(p, n_hub, stream);
    }
  }
  switch (ht) {
    case 1: return launch_gat_bwd<VEC, CH, 1>(which, p, n_hub, stream);
    case 2: return launch_gat_bwd<VEC, CH, 2>(which, p, n_hub, stream);
    case 4: return launch_gat_bwd<VEC, CH, 4>(which, p, n_hub, stream);
    default: return launch_gat_bwd<VEC, CH, 8>(which, p, n_hub, stream);
  }
}

size_t gat_hub_workspace_bytes(int64_t n_seg, int64_t H, int64_t F) {
  return (size_t)n_seg * (size_t)(H * F + 4 * H) * sizeof(float);
}

// which: 0 fwd, 1 bwd_dst, 2 bwd_src
int gat_fused_f32(int which, GatParams& p, int64_t H, int64_t F, float dropout_p, uint64_t seed,
                  const dglb_hub_t* hub, cudaStream_t stream) {
  if (p.n_rows == 0) return DGLB_OK;
  int vec, ch, ht;
  const void* a1 = which == 0 ? (const void*)p.out_feat : (const void*)p.dZ;
  const void* a2 = which == 2 ? (const void*)p.out_feat : nullptr;
  if (gat_geometry(p, H, F, p.ft, a1, a2, &vec, &ch, &ht) != DGLB_OK) {
    set_error("gat_fused: unsupported shape H=%lld F=%lld (need 1<=H<=8)", (long long)H, (long long)F);
    return DGLB_E_UNSUPPORTED;
  }
  const bool use_hub = hub && hub->n_hub > 0 && hub->rows;
  int n_hub = use_hub ? hub->n_hub : 0;
  p.hub_rows = use_hub ? hub->rows : nullptr;
  p.hub_threshold = use_hub ? hub->threshold : INT32_MAX;
  p.n_hub = n_hub; p.n_seg = 0; p.seg_len = 0;
  p.seg_ptr = nullptr; p.seg_hub = nullptr; p.ws_feat = nullptr; p.ws_tot = nullptr;
  // segmented hub path when the caller provides the segment lists and a workspace; otherwise one CTA per hub row
  if (use_hub && hub->seg_ptr && hub->seg_hub && hub->n_seg > 0 && hub->seg_len > 0 && hub->workspace) {
    const size_t need = gat_hub_workspace_bytes(hub->n_seg, H, F);
    if (hub->workspace_bytes < need) {
      set_error("gat_fused: hub workspace too small (%zu < %zu bytes)", hub->workspace_bytes, need);
      return DGLB_E_WORKSPACE;
    }
    DGLB_CHECK_ARG((reinterpret_cast<uintptr_t>(hub->workspace) & 15) == 0, "gat_fused: hub workspace must be 16-byte aligned");
    p.seg_ptr = hub->seg_ptr; p.seg_hub = hub->seg_hub; p.n_seg = hub->n_seg; p.seg_len = hub->seg_len;
    p.ws_feat = static_cast<float*>(hub->workspace);
    p.ws_tot = p.ws_feat + (size_t)hub->n_seg * (size_t)(H * F);
  }
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) { set_error("gat_fused: dropout_p must be in [0,1)"); return DGLB_E_INVALID; }
  p.drop_p = dropout_p;
  p.drop_scale = 1.f / (1.f - dropout_p);
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
  static const int prefetch = [] { const char* e = getenv("DGLB_GAT_PREFETCH"); return e ? atoi(e) : 1; }();
  p.prefetch = prefetch;
#define DGLB_CASE(V, C) if (vec == V && ch == C) return dispatch_gat<V, C>(which, p, ht, n_hub, stream);
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  return DGLB_E_UNSUPPORTED;
}

}  // namespace dglb
