// gat_fused.cu -- fused SDDMM -> leaky_relu -> edge_softmax -> (attention dropout) -> SpMM for
// GAT attention, forward and backward, hand-written for sm_100a.
//
// Replaces, for dgl.nn.pytorch.GATConv (main_dgl_arxiv_gat.py:9; written-out twin
// main_pyg_arxiv_gat.py:98-111), upstream DGL v0.6.1's chain of ~7 sparse + 3 elementwise
// launches forward and ~6 + 5 backward and its per-edge (E,H) intermediates.
//
// Forward (CSC, one group of G lanes per destination row v, lanes over the H*F feature columns):
//   A. lanes walk the in-edges (one edge per lane): e = lrelu(el[src] + er[v]); group-reduce the
//      per-head max, then the per-head sum of exp(e - max)  (same formula as upstream's
//      edge_softmax: exp(x - max), sum, divide).  Rows of <= G edges keep e in registers.
//   B. the lane that owns edge j computes a_j[h] = exp(e-max)/sum (times the dropout factor)
//      and the group broadcasts (src_j, a_j[*]) with shuffles while every lane gathers its
//      128-bit chunks of ft[src_j] -- exactly the SpMM inner loop with a head-broadcast weight.
//   Nothing per-edge is written; row_max / row_sum (N,H) are saved for the backward.
// Backward: two passes that RECOMPUTE a_j from (el, er, row_max, row_sum):
//   dst pass (CSC): per-lane partials of  S1 = sum_j a_j dd_j,  S2 = sum_j a_j g_j dd_j  with
//      dd_j = drop_j * <ft[src_j,h,:], dZ[v,h,:]> are accumulated over the row's edges and reduced
//      across lanes ONCE per row (the sums are linear in the per-lane partial dots), plus
//      S3 = sum_j a_j g_j;  grad_er = S2 - S1*S3.
//   src pass (CSR): grad_ft[u] = sum a*drop*dZ[v];  grad_el[u] = sum a g (dd - S1[v]) with the
//      same once-per-row reduction.
//   The sign bit of the broadcast weight carries lrelu' (a >= 0), so one shuffle per head per
//   edge moves both.
// Hub rows: one CTA per row; the groups split the edges and meet in shared memory.
#include "kernels.cuh"

namespace dglb {

constexpr int kMaxHeads = 8;


// counter-based dropout: keep iff hash(seed, edge*H + h) >= p * 2^32
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float drop_factor(const GatParams& p, int64_t e, int h) {
  if (p.drop_p <= 0.f) return 1.f;
  const uint64_t ctr = (uint64_t)e * (uint64_t)p.H + (uint64_t)h;
  uint32_t x = mix32((uint32_t)ctr ^ p.seed_lo);
  x = mix32(x + (uint32_t)(ctr >> 32) * 0x9e3779b9u + p.seed_hi);
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);  // [0,1)
  return u < p.drop_p ? 0.f : p.drop_scale;
}

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }

template <int HT>
__device__ __forceinline__ void group_allreduce_max(float (&v)[HT], int G) {
#pragma unroll
  for (int h = 0; h < HT; ++h)
    for (int s = G >> 1; s > 0; s >>= 1) v[h] = fmaxf(v[h], __shfl_xor_sync(FULL_MASK, v[h], s));
}
template <int HT>
__device__ __forceinline__ void group_allreduce_sum(float (&v)[HT], int G) {
#pragma unroll
  for (int h = 0; h < HT; ++h)
    for (int s = G >> 1; s > 0; s >>= 1) v[h] += __shfl_xor_sync(FULL_MASK, v[h], s);
}

// Combine per-group values (uniform within a group) across the CTA's groups, fixed order.
template <int HT, bool IS_MAX>
__device__ __forceinline__ void cta_allreduce(float (&v)[HT], float* s_buf, int gidx, int lg, int n_groups) {
  __syncthreads();
  if (lg == 0) {
#pragma unroll
    for (int h = 0; h < HT; ++h) s_buf[gidx * HT + h] = v[h];
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    float r = s_buf[h];
    for (int g = 1; g < n_groups; ++g) r = IS_MAX ? fmaxf(r, s_buf[g * HT + h]) : r + s_buf[g * HT + h];
    v[h] = r;
  }
}

// row / slice of the calling group (same scheme as spmm.cu / sddmm.cu)
template <bool HUB>
__device__ __forceinline__ void gat_group_work(const GatParams& p, int64_t& row, bool& active, int64_t& j0,
                                               int& n, int& gidx, int& n_groups) {
  n_groups = kBlockThreads >> p.log2G;
  gidx = threadIdx.x >> p.log2G;
  if constexpr (!HUB) {
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

// ------------------------------------------------------------------ forward
template <int VEC, int CH, int HT, bool HUB>
__global__ void __launch_bounds__(kBlockThreads) gat_fwd_kernel(const GatParams p) {
  constexpr int U = 8 / CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_buf = reinterpret_cast<float*>(smem_raw);
  const int G = p.G, H = p.H;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  bool active;
  int n, gidx, n_groups;
  gat_group_work<HUB>(p, row, active, j0, n, gidx, n_groups);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const bool single = nmax <= G;  // every row of this warp fits one batch: keep e in registers

  float er_h[HT], mx[HT], sm[HT], e_reg[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    er_h[h] = (h < H && (active || HUB)) ? __ldg(p.er + row * H + h) : 0.f;
    mx[h] = -INFINITY; sm[h] = 0.f; e_reg[h] = -INFINITY;
  }
  // ---- A1: per-head max
  for (int off = 0; off < nmax; off += G) {
    const bool valid = off + lg < n;
    const int c = valid ? __ldg(p.indices + j0 + off + lg) : 0;
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      const float e = (valid && h < H) ? lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er_h[h]), p.slope)
                                       : -INFINITY;
      e_reg[h] = e;
      mx[h] = fmaxf(mx[h], e);
    }
  }
  group_allreduce_max<HT>(mx, G);
  if constexpr (HUB) cta_allreduce<HT, true>(mx, s_buf, gidx, lg, n_groups);
  // ---- A2: per-head sum of exp(e - max)
  for (int off = 0; off < nmax; off += G) {
    const bool valid = off + lg < n;
    const int c = (valid && !single) ? __ldg(p.indices + j0 + off + lg) : 0;
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      if (valid && h < H) {
        const float e = single ? e_reg[h]
                               : lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * H + h), er_h[h]), p.slope);
        sm[h] += expf(__fsub_rn(e, mx[h]));
      }
    }
  }
  group_allreduce_sum<HT>(sm, G);
  if constexpr (HUB) cta_allreduce<HT, false>(sm, s_buf, gidx, lg, n_groups);
  if (active && lg < H && (!HUB || gidx == 0)) {
    // lane h writes head h (register arrays are indexed statically: select by loop)
#pragma unroll
    for (int h = 0; h < HT; ++h)
      if (h == lg) { p.out_h0[row * H + h] = mx[h]; p.out_h1[row * H + h] = sm[h]; }
  }

  // ---- B: weighted gather of ft rows
  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    bool colv[CH];
    int k[CH], hk[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = k[c] / p.F;
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      const bool valid = lg < m;
      int my_c = 0;
      float a_my[HT];
      int64_t my_e = 0;
      if (valid) {
        my_c = __ldg(p.indices + j0 + off + lg);
        my_e = p.eids ? (int64_t)__ldg(p.eids + j0 + off + lg) : (j0 + off + lg);
      }
#pragma unroll
      for (int h = 0; h < HT; ++h) {
        a_my[h] = 0.f;
        if (valid && h < H) {
          const float e = single ? e_reg[h]
                                 : lrelu(__fadd_rn(__ldg(p.el + (int64_t)my_c * H + h), er_h[h]), p.slope);
          const float a = __fdiv_rn(expf(__fsub_rn(e, mx[h])), sm[h]);
          if (p.edge_scores && tile0 == 0) p.edge_scores[my_e * H + h] = a;
          a_my[h] = a * drop_factor(p, my_e, h);
        }
      }
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U];
        float w[U][CH];
        FVec<VEC> xv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
#pragma unroll
          for (int c = 0; c < CH; ++c) w[u][c] = 0.f;
#pragma unroll
          for (int h = 0; h < HT; ++h) {
            if (h < H) {
              const float tmp = __shfl_sync(FULL_MASK, a_my[h], t + u, G);
#pragma unroll
              for (int c = 0; c < CH; ++c)
                if (hk[c] == h) w[u][c] = tmp;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ev = (t + u) < m;
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (ev && colv[c]) xv[u][c] = ldg_vec<VEC>(p.ft + (int64_t)cc[u] * p.D + k[c]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ev = (t + u) < m;
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (ev && colv[c]) {
#pragma unroll
              for (int v = 0; v < VEC; ++v)
                acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], w[u][c]));
            }
        }
      }
    }
    if constexpr (!HUB) {
      if (active) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (colv[c]) {
            FVec<VEC> o;
#pragma unroll
            for (int v = 0; v < VEC; ++v) o.v[v] = acc[c][v];
            st_vec<VEC>(p.out_feat + row * (int64_t)p.D + k[c], o);
          }
      }
    } else {
      const int tile_elems = G * CH * VEC;
      float* s_val = s_buf + n_groups * HT;  // after the head scratch
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
}

// ------------------------------------------------------------------ backward
// SRC_PASS = false: CSC over dst rows v.  neighbour = src u: gathers ft[u]; own row: dZ[v].
//            outputs s1[v,h], grad_er[v,h].
// SRC_PASS = true : CSR over src rows u.  neighbour = dst v: gathers dZ[v]; own row: ft[u].
//            outputs grad_ft[u,:], grad_el[u,h].
template <int VEC, int CH, int HT, bool SRC_PASS, bool HUB>
__global__ void __launch_bounds__(kBlockThreads) gat_bwd_kernel(const GatParams p) {
  constexpr int U = 8 / CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_buf = reinterpret_cast<float*>(smem_raw);
  const int G = p.G, H = p.H;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  bool active;
  int n, gidx, n_groups;
  gat_group_work<HUB>(p, row, active, j0, n, gidx, n_groups);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const bool live = active || HUB;

  // per-head constants of the own row
  float own0[HT], own1[HT], own2[HT];  // dst pass: er, max, sum ; src pass: el
#pragma unroll
  for (int h = 0; h < HT; ++h) {
    own0[h] = own1[h] = own2[h] = 0.f;
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
  // head-level accumulators owned by edge-owner lanes: dst pass S3 = sum a g ; src pass T = sum a g s1[v]
  float hacc[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) hacc[h] = 0.f;
  // per-lane per-chunk partial sums of  a*dd  and  a*g*dd  (reduced once per row)
  float p1[CH], p2[CH];
  // feature accumulators (src pass: grad_ft)
  const float* __restrict__ own_feat = (SRC_PASS ? p.ft : p.dZ) + row * (int64_t)p.D;
  const float* __restrict__ nb_feat = SRC_PASS ? p.dZ : p.ft;

  // per-head totals over feature tiles
  float tot1[HT], tot2[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) tot1[h] = tot2[h] = 0.f;

  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    FVec<VEC> ownv[CH];
    bool colv[CH];
    int k[CH], hk[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = k[c] / p.F;
      p1[c] = p2[c] = 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { acc[c][v] = 0.f; ownv[c].v[v] = 0.f; }
      if (live && colv[c]) ownv[c] = ldg_vec<VEC>(own_feat + k[c]);
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      const bool valid = lg < m;
      int my_c = 0;
      float w_my[HT];   // |w| = a (no dropout), sign bit set when lrelu' == slope
      float d_my[HT];   // dropout factor
      if (valid) my_c = __ldg(p.indices + j0 + off + lg);
      int64_t my_e = 0;
      if (valid && p.drop_p > 0.f)
        my_e = p.eids ? (int64_t)__ldg(p.eids + j0 + off + lg) : (j0 + off + lg);
#pragma unroll
      for (int h = 0; h < HT; ++h) {
        w_my[h] = 0.f; d_my[h] = 1.f;
        if (valid && h < H) {
          float x, mxv, smv;
          if constexpr (!SRC_PASS) {
            x = __fadd_rn(__ldg(p.el + (int64_t)my_c * H + h), own0[h]);
            mxv = own1[h]; smv = own2[h];
          } else {
            x = __fadd_rn(own0[h], __ldg(p.er + (int64_t)my_c * H + h));
            mxv = __ldg(p.row_max + (int64_t)my_c * H + h);
            smv = __ldg(p.row_sum + (int64_t)my_c * H + h);
          }
          const float a = __fdiv_rn(expf(__fsub_rn(lrelu(x, p.slope), mxv)), smv);
          const float g = x > 0.f ? 1.f : p.slope;
          if (tile0 == 0) {
            if constexpr (!SRC_PASS) hacc[h] += a * g;
            else hacc[h] += a * g * __ldg(p.s1 + (int64_t)my_c * H + h);
          }
          w_my[h] = x > 0.f ? a : -a;
          d_my[h] = drop_factor(p, my_e, h);
        }
      }
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U];
        float wa[U][CH], wd[U][CH];
        FVec<VEC> xv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
#pragma unroll
          for (int c = 0; c < CH; ++c) { wa[u][c] = 0.f; wd[u][c] = 1.f; }
#pragma unroll
          for (int h = 0; h < HT; ++h) {
            if (h < H) {
              const float tw = __shfl_sync(FULL_MASK, w_my[h], t + u, G);
              float td = 1.f;
              if (p.drop_p > 0.f) td = __shfl_sync(FULL_MASK, d_my[h], t + u, G);
#pragma unroll
              for (int c = 0; c < CH; ++c)
                if (hk[c] == h) { wa[u][c] = tw; wd[u][c] = td; }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ev = (t + u) < m;
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (ev && colv[c]) xv[u][c] = ldg_vec<VEC>(nb_feat + (int64_t)cc[u] * p.D + k[c]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ev = (t + u) < m;
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (ev && colv[c]) {
              const float a = fabsf(wa[u][c]);
              const float g = (__float_as_int(wa[u][c]) < 0) ? p.slope : 1.f;
              const float ad = a * wd[u][c];  // a * drop
              float dot = 0.f;
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                dot = fmaf(xv[u][c].v[v], ownv[c].v[v], dot);
                if constexpr (SRC_PASS) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], ad));
              }
              p1[c] = fmaf(ad, dot, p1[c]);
              p2[c] = fmaf(ad * g, dot, p2[c]);
            }
        }
      }
    }
    // fold this tile's per-chunk partials into per-head totals (still per lane)
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int h = 0; h < HT; ++h)
        if (colv[c] && hk[c] == h) { tot1[h] += p1[c]; tot2[h] += p2[c]; }
    if constexpr (SRC_PASS) {
      if constexpr (!HUB) {
        if (active) {
#pragma unroll
          for (int c = 0; c < CH; ++c)
            if (colv[c]) {
              FVec<VEC> o;
#pragma unroll
              for (int v = 0; v < VEC; ++v) o.v[v] = acc[c][v];
              st_vec<VEC>(p.out_feat + row * (int64_t)p.D + k[c], o);
            }
        }
      } else {
        const int tile_elems = G * CH * VEC;
        float* s_val = s_buf + n_groups * HT;
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
  }
  // ---- once-per-row reductions
  group_allreduce_sum<HT>(tot1, G);
  group_allreduce_sum<HT>(tot2, G);
  group_allreduce_sum<HT>(hacc, G);
  if constexpr (HUB) {
    cta_allreduce<HT, false>(tot1, s_buf, gidx, lg, n_groups);
    cta_allreduce<HT, false>(tot2, s_buf, gidx, lg, n_groups);
    cta_allreduce<HT, false>(hacc, s_buf, gidx, lg, n_groups);
  }
  if (active && lg < H && (!HUB || gidx == 0)) {
#pragma unroll
    for (int h = 0; h < HT; ++h)
      if (h == lg) {
        if constexpr (!SRC_PASS) {
          p.out_h0[row * H + h] = tot1[h];                                // S1
          p.out_h1[row * H + h] = __fsub_rn(tot2[h], tot1[h] * hacc[h]);  // grad_er = S2 - S1*S3
        } else {
          p.out_h0[row * H + h] = __fsub_rn(tot2[h], hacc[h]);            // grad_el = sum a g dd - sum a g s1
        }
      }
  }
}

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
  return DGLB_OK;
}

static size_t gat_hub_smem(const GatParams& p, int vec, int ch, int ht) {
  const int n_groups = kBlockThreads / p.G;
  return sizeof(float) * ((size_t)n_groups * ht + (size_t)kBlockThreads * ch * vec);
}

template <int VEC, int CH, int HT>
static int launch_gat(int which, const GatParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = gat_hub_smem(p, VEC, CH, HT);
  if (blocks > 0) {
    if (which == 0) gat_fwd_kernel<VEC, CH, HT, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    else if (which == 1) gat_bwd_kernel<VEC, CH, HT, false, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    else gat_bwd_kernel<VEC, CH, HT, true, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat kernel");
  }
  if (n_hub > 0) {
    if (which == 0) gat_fwd_kernel<VEC, CH, HT, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    else if (which == 1) gat_bwd_kernel<VEC, CH, HT, false, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    else gat_bwd_kernel<VEC, CH, HT, true, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, int CH>
static int dispatch_ht(int which, const GatParams& p, int ht, int n_hub, cudaStream_t stream) {
  switch (ht) {
    case 1: return launch_gat<VEC, CH, 1>(which, p, n_hub, stream);
    case 2: return launch_gat<VEC, CH, 2>(which, p, n_hub, stream);
    case 4: return launch_gat<VEC, CH, 4>(which, p, n_hub, stream);
    default: return launch_gat<VEC, CH, 8>(which, p, n_hub, stream);
  }
}

// which: 0 fwd, 1 bwd_dst, 2 bwd_src
int gat_fused_f32(int which, GatParams& p, int64_t H, int64_t F, float dropout_p, uint64_t seed,
                  int32_t n_hub, int32_t hub_threshold, cudaStream_t stream) {
  if (p.n_rows == 0) return DGLB_OK;
  int vec, ch, ht;
  const void* a1 = which == 0 ? (const void*)p.out_feat : (const void*)p.dZ;
  const void* a2 = which == 2 ? (const void*)p.out_feat : nullptr;
  if (gat_geometry(p, H, F, p.ft, a1, a2, &vec, &ch, &ht) != DGLB_OK) {
    set_error("gat_fused: unsupported shape H=%lld F=%lld (need 1<=H<=8)", (long long)H, (long long)F);
    return DGLB_E_UNSUPPORTED;
  }
  const bool hub = n_hub > 0 && p.hub_rows;
  p.hub_threshold = hub ? hub_threshold : INT32_MAX;
  if (!hub) n_hub = 0;
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) { set_error("gat_fused: dropout_p must be in [0,1)"); return DGLB_E_INVALID; }
  p.drop_p = dropout_p;
  p.drop_scale = 1.f / (1.f - dropout_p);
  p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
#define DGLB_CASE(V, C) if (vec == V && ch == C) return dispatch_ht<V, C>(which, p, ht, n_hub, stream);
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  return DGLB_E_UNSUPPORTED;
}

}  // namespace dglb
