// gat_fused.cu -- fused SDDMM -> leaky_relu -> edge_softmax -> (attention dropout) -> SpMM for
// GAT attention, forward and backward, hand-written for sm_100a.
//
// Replaces, for dgl.nn.pytorch.GATConv (main_dgl_arxiv_gat.py:9; written-out twin
// main_pyg_arxiv_gat.py:98-111), upstream DGL v0.6.1's chain of ~7 sparse + 3 elementwise
// launches forward and ~6 + 5 backward and its per-edge (E,H) intermediates.
//
// Forward = two launches, nothing per-edge is ever written:
//   1. gat_rowstats_kernel (CSC; a warp per destination row, lanes = (edge slot, head)):
//      e = lrelu(el[src] + er[v]); row_max[v,h] = max e; row_sum[v,h] = sum exp(e - max)
//      -- the same formula as upstream's edge_softmax.  Traffic: 4 + 4H bytes per edge.
//   2. gat_fwd_kernel: the SpMM row-per-group gather in which EVERY LANE recomputes the weight
//      of its own head,  a = exp(lrelu(el[src,h] + er[v,h]) - max[v,h]) / sum[v,h]  (x dropout),
//      from one extra 4-byte load per chunk.  The redundant exp per lane is free on a gather-bound
//      kernel, and it removes every cross-lane dependency from the inner loop: the source row and
//      its el value are fetched by independent loads issued back to back (U*CH in flight), exactly
//      like gspmm u_mul_e.  (Round-1 profile: the first version, whose owner lane computed the
//      weights and broadcast them with H shuffles per edge after in-kernel reductions, ran 2x
//      slower than plain u_mul_e_sum -- profiles/r01_notes.md.)
// Backward = two launches that RECOMPUTE a_j per lane the same way:
//   gat_bwd_kernel<SRC_PASS=false> (CSC): per-lane partials of  S1 = sum_j a_j dd_j  and
//      S2 = sum_j a_j g_j dd_j  (dd_j = drop_j <ft[src_j,h,:], dZ[v,h,:]>, g = lrelu') are
//      accumulated over the row's edges and reduced across lanes ONCE per row (the sums are linear
//      in the per-lane partial dots); S3 = sum_j a_j g_j;  grad_er = S2 - S1*S3;  writes the 16-byte
//      record row_pack[v,h] = {er, max, sum, S1}.
//   gat_bwd_kernel<SRC_PASS=true> (CSR): grad_ft[u] = sum a*drop*dZ[v];
//      grad_el[u] = sum a g (dd - S1[v]); the destination's record is one LDG.128 per edge.
// Hub rows: one CTA per row; the groups split the edges and meet in shared memory.
#include "kernels.cuh"

namespace dglb {

constexpr int kMaxHeads = 8;

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

// ------------------------------------------------------------------ forward 1/2: row statistics
__device__ __forceinline__ float slot_max(float v, int HP) {
  for (int s = 16; s >= HP; s >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, s));
  return v;
}
__device__ __forceinline__ float slot_sum(float v, int HP) {
  for (int s = 16; s >= HP; s >>= 1) v += __shfl_xor_sync(FULL_MASK, v, s);
  return v;
}
template <bool IS_MAX>
__device__ __forceinline__ float cta_slot_reduce(float v, float* s_buf /* [8][32] */) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  s_buf[w * 32 + lane] = v;
  __syncthreads();
  float r = s_buf[lane];
#pragma unroll
  for (int i = 1; i < kBlockThreads / 32; ++i) {
    const float o = s_buf[i * 32 + lane];
    r = IS_MAX ? fmaxf(r, o) : r + o;
  }
  return r;
}

template <bool HUB>
__global__ void __launch_bounds__(kBlockThreads) gat_rowstats_kernel(const GatParams p) {
  __shared__ float s_buf[HUB ? kBlockThreads : 1];
  const int lane = threadIdx.x & 31;
  const int h = lane & (p.HP - 1);
  const bool hv = h < p.H;
  int64_t row;
  int start = 0, deg = 0, slot, nslots;
  bool write;
  if constexpr (!HUB) {
    row = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
    write = row < p.n_rows;
    if (write) {
      start = __ldg(p.indptr + row);
      deg = __ldg(p.indptr + row + 1) - start;
      if (deg > p.hub_threshold) { deg = 0; write = false; }
    } else {
      row = 0;
    }
    slot = lane >> p.log2HP;
    nslots = 32 >> p.log2HP;
  } else {
    row = p.hub_rows[blockIdx.x];
    write = true;
    start = __ldg(p.indptr + row);
    deg = __ldg(p.indptr + row + 1) - start;
    slot = threadIdx.x >> p.log2HP;
    nslots = kBlockThreads >> p.log2HP;
  }
  const float er = (hv && write) ? __ldg(p.er + row * p.H + h) : 0.f;
  float mx = -INFINITY;
  for (int i = slot; i < deg; i += nslots) {
    const int c = __ldg(p.indices + start + i);
    if (hv) mx = fmaxf(mx, lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * p.H + h), er), p.slope));
  }
  mx = slot_max(mx, p.HP);
  if constexpr (HUB) mx = cta_slot_reduce<true>(mx, s_buf);
  float sum = 0.f;
  for (int i = slot; i < deg; i += nslots) {
    const int c = __ldg(p.indices + start + i);
    if (hv) sum += expf(__fsub_rn(lrelu(__fadd_rn(__ldg(p.el + (int64_t)c * p.H + h), er), p.slope), mx));
  }
  sum = slot_sum(sum, p.HP);
  if constexpr (HUB) sum = cta_slot_reduce<false>(sum, s_buf);
  if (write && hv && slot == 0) {
    p.out_h0[row * p.H + h] = mx;
    p.out_h1[row * p.H + h] = sum;
  }
}

// shared epilogue: write a feature tile (row kernel) or combine the CTA's groups (hub kernel)
template <int VEC, int CH, bool HUB>
__device__ __forceinline__ void store_feat_tile(const GatParams& p, float (&acc)[CH][VEC], const bool (&colv)[CH],
                                                const int (&k)[CH], int64_t row, bool active, int tile0, int lg,
                                                int gidx, int n_groups, float* s_val) {
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

// ------------------------------------------------------------------ forward 2/2: weighted gather
template <int VEC, int CH, bool HUB>
__global__ void __launch_bounds__(kBlockThreads, 2) gat_fwd_kernel(const GatParams p) {
  constexpr int U = 8 / CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_val = reinterpret_cast<float*>(smem_raw);
  const int G = p.G, H = p.H;
  const int lg = threadIdx.x & (G - 1);
  int64_t row, j0;
  bool active;
  int n, gidx, n_groups;
  gat_group_work<HUB>(p, row, active, j0, n, gidx, n_groups);
  const int nmax = __reduce_max_sync(FULL_MASK, n);
  const bool live = active || HUB;
  const bool need_e = p.drop_p > 0.f || p.edge_scores != nullptr;

  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    bool colv[CH];
    int k[CH], hk[CH];
    float er_c[CH], mx_c[CH], sm_c[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = colv[c] ? k[c] / p.F : 0;
      er_c[c] = 0.f; mx_c[c] = 0.f; sm_c[c] = 1.f;
      if (live && colv[c]) {
        er_c[c] = __ldg(p.er + row * H + hk[c]);
        mx_c[c] = __ldg(p.row_max + row * H + hk[c]);
        sm_c[c] = __ldg(p.row_sum + row * H + hk[c]);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      int my_c = 0, my_e = 0;
      if (lg < m) {
        my_c = __ldg(p.indices + j0 + off + lg);
        if (need_e) my_e = p.eids ? __ldg(p.eids + j0 + off + lg) : (int)(j0 + off + lg);
      }
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U], ee[U];
        FVec<VEC> xv[U][CH];
        float elv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
          ee[u] = need_e ? __shfl_sync(FULL_MASK, my_e, t + u, G) : 0;
        }
        // load phase: source rows and their el values, all independent
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              xv[u][c] = ldg_vec<VEC>(p.ft + (int64_t)cc[u] * p.D + k[c]);
              elv[u][c] = __ldg(p.el + (int64_t)cc[u] * H + hk[c]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              const float e = lrelu(__fadd_rn(elv[u][c], er_c[c]), p.slope);
              float a = __fdiv_rn(expf(__fsub_rn(e, mx_c[c])), sm_c[c]);
              if (p.edge_scores != nullptr && k[c] == hk[c] * p.F)
                p.edge_scores[(int64_t)ee[u] * H + hk[c]] = a;  // first lane of the head writes the score
              if (p.drop_p > 0.f) a *= drop_factor(p, ee[u], hk[c]);
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], a));
            }
          }
        }
      }
    }
    store_feat_tile<VEC, CH, HUB>(p, acc, colv, k, row, active, tile0, lg, gidx, n_groups, s_val);
  }
}

// ------------------------------------------------------------------ backward
// SRC_PASS = false: CSC over dst rows v.  neighbour = src u: gathers ft[u], el[u]; own row: dZ[v].
//            outputs row_pack[v,h] = {er, max, sum, S1}, grad_er[v,h].
// SRC_PASS = true : CSR over src rows u.  neighbour = dst v: gathers dZ[v], row_pack[v]; own row: ft[u].
//            outputs grad_ft[u,:], grad_el[u,h].
template <int VEC, int CH, int HT, bool SRC_PASS, bool HUB>
__global__ void __launch_bounds__(kBlockThreads, 2) gat_bwd_kernel(const GatParams p) {
  // the src pass stages a 16-byte record next to every gathered chunk: halve the batch to keep
  // the kernel at <= 128 registers (2 CTAs / SM)
  constexpr int U = SRC_PASS ? (CH >= 4 ? 1 : 4 / CH) : 8 / CH;
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
  const bool need_e = p.drop_p > 0.f;
  float* s_val = s_buf + n_groups * HT;

  const float* __restrict__ own_feat = (SRC_PASS ? p.ft : p.dZ) + row * (int64_t)p.D;
  const float* __restrict__ nb_feat = SRC_PASS ? p.dZ : p.ft;

  // per-head totals (per lane until the once-per-row reduction)
  float tot1[HT], tot2[HT], tot3[HT];
#pragma unroll
  for (int h = 0; h < HT; ++h) tot1[h] = tot2[h] = tot3[h] = 0.f;

  for (int tile0 = 0; tile0 < p.ncols; tile0 += G * CH) {
    float acc[CH][VEC];
    FVec<VEC> ownv[CH];
    bool colv[CH], lead[CH];
    int k[CH], hk[CH];
    float own0[CH], own1[CH], own2[CH];  // dst pass: er, max, sum of the own row;  src pass: el
    float p1[CH], p2[CH], p3[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int vc = tile0 + c * G + lg;
      colv[c] = vc < p.ncols;
      k[c] = vc * VEC;
      hk[c] = colv[c] ? k[c] / p.F : 0;
      lead[c] = colv[c] && (k[c] == hk[c] * p.F);  // first lane of the head: owns the per-(edge,head) terms
      p1[c] = p2[c] = p3[c] = 0.f;
      own0[c] = 0.f; own1[c] = 0.f; own2[c] = 1.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { acc[c][v] = 0.f; ownv[c].v[v] = 0.f; }
      if (live && colv[c]) {
        ownv[c] = ldg_vec<VEC>(own_feat + k[c]);
        if constexpr (!SRC_PASS) {
          own0[c] = __ldg(p.er + row * H + hk[c]);
          own1[c] = __ldg(p.row_max + row * H + hk[c]);
          own2[c] = __ldg(p.row_sum + row * H + hk[c]);
        } else {
          own0[c] = __ldg(p.el + row * H + hk[c]);
        }
      }
    }
    for (int off = 0; off < nmax; off += G) {
      const int m = min(max(n - off, 0), G);
      int my_c = 0, my_e = 0;
      if (lg < m) {
        my_c = __ldg(p.indices + j0 + off + lg);
        if (need_e) my_e = p.eids ? __ldg(p.eids + j0 + off + lg) : (int)(j0 + off + lg);
      }
      const int mmax = min(G, nmax - off);
      for (int t = 0; t < mmax; t += U) {
        int cc[U], ee[U];
        FVec<VEC> xv[U][CH];
        float4 nb[U][CH];  // dst pass: .x = el[u,h];  src pass: {er, max, sum, s1} of the destination
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = __shfl_sync(FULL_MASK, my_c, t + u, G);
          ee[u] = need_e ? __shfl_sync(FULL_MASK, my_e, t + u, G) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              xv[u][c] = ldg_vec<VEC>(nb_feat + (int64_t)cc[u] * p.D + k[c]);
              if constexpr (!SRC_PASS) nb[u][c].x = __ldg(p.el + (int64_t)cc[u] * H + hk[c]);
              else nb[u][c] = __ldg(p.pack + (int64_t)cc[u] * H + hk[c]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if ((t + u) < m && colv[c]) {
              float x, mxv, smv;
              if constexpr (!SRC_PASS) { x = __fadd_rn(nb[u][c].x, own0[c]); mxv = own1[c]; smv = own2[c]; }
              else { x = __fadd_rn(own0[c], nb[u][c].x); mxv = nb[u][c].y; smv = nb[u][c].z; }
              const float a = __fdiv_rn(expf(__fsub_rn(lrelu(x, p.slope), mxv)), smv);
              const float g = x > 0.f ? 1.f : p.slope;
              const float ad = need_e ? a * drop_factor(p, ee[u], hk[c]) : a;  // a * drop
              float dot = 0.f;
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                dot = fmaf(xv[u][c].v[v], ownv[c].v[v], dot);
                if constexpr (SRC_PASS) acc[c][v] = __fadd_rn(acc[c][v], __fmul_rn(xv[u][c].v[v], ad));
              }
              p1[c] = fmaf(ad, dot, p1[c]);
              p2[c] = fmaf(ad * g, dot, p2[c]);
              if (lead[c]) {
                if constexpr (!SRC_PASS) p3[c] = fmaf(a, g, p3[c]);                // S3 = sum a g
                else p3[c] = fmaf(a * g, nb[u][c].w, p3[c]);                       // T  = sum a g S1[v]
              }
            }
          }
        }
      }
    }
    // fold this tile's per-chunk partials into per-head totals (still per lane)
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int h = 0; h < HT; ++h)
        if (colv[c] && hk[c] == h) { tot1[h] += p1[c]; tot2[h] += p2[c]; tot3[h] += p3[c]; }
    if constexpr (SRC_PASS) {
      store_feat_tile<VEC, CH, HUB>(p, acc, colv, k, row, active, tile0, lg, gidx, n_groups, s_val);
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
  if (active && lg < H && (!HUB || gidx == 0)) {
#pragma unroll
    for (int h = 0; h < HT; ++h)
      if (h == lg) {
        if constexpr (!SRC_PASS) {
          p.out_pack[row * H + h] = make_float4(__ldg(p.er + row * H + h), __ldg(p.row_max + row * H + h),
                                                __ldg(p.row_sum + row * H + h), tot1[h]);
          p.out_h0[row * H + h] = __fsub_rn(tot2[h], tot1[h] * tot3[h]);  // grad_er = S2 - S1*S3
        } else {
          p.out_h0[row * H + h] = __fsub_rn(tot2[h], tot3[h]);            // grad_el = sum a g dd - sum a g S1
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
  p.HP = ht; p.log2HP = 0;
  while ((1 << p.log2HP) < p.HP) ++p.log2HP;
  return DGLB_OK;
}

static size_t gat_hub_smem(const GatParams& p, int vec, int ch, int ht) {
  const int n_groups = kBlockThreads / p.G;
  return sizeof(float) * ((size_t)n_groups * ht + (size_t)kBlockThreads * ch * vec);
}

template <int VEC, int CH>
static int launch_gat_fwd(const GatParams& p, int n_hub, cudaStream_t stream) {
  // 1. row statistics (warp per row)
  const int64_t sblocks = (p.n_rows + (kBlockThreads / 32) - 1) / (kBlockThreads / 32);
  gat_rowstats_kernel<false><<<(unsigned)sblocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("gat_rowstats_kernel");
  if (n_hub > 0) {
    gat_rowstats_kernel<true><<<n_hub, kBlockThreads, 0, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_rowstats_kernel(hub)");
  }
  // 2. weighted gather
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  gat_fwd_kernel<VEC, CH, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("gat_fwd_kernel");
  if (n_hub > 0) {
    const size_t smem = sizeof(float) * (size_t)kBlockThreads * CH * VEC;
    gat_fwd_kernel<VEC, CH, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_fwd_kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, int CH, int HT>
static int launch_gat_bwd(int which, const GatParams& p, int n_hub, cudaStream_t stream) {
  const int rows_per_block = kBlockThreads / p.G;
  const int64_t blocks = (p.n_rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = gat_hub_smem(p, VEC, CH, HT);
  if (which == 1) gat_bwd_kernel<VEC, CH, HT, false, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  else gat_bwd_kernel<VEC, CH, HT, true, false><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(p);
  DGLB_LAUNCH_CHECK("gat_bwd_kernel");
  if (n_hub > 0) {
    if (which == 1) gat_bwd_kernel<VEC, CH, HT, false, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    else gat_bwd_kernel<VEC, CH, HT, true, true><<<n_hub, kBlockThreads, smem, stream>>>(p);
    DGLB_LAUNCH_CHECK("gat_bwd_kernel(hub)");
  }
  return DGLB_OK;
}

template <int VEC, int CH>
static int dispatch_gat(int which, const GatParams& p, int ht, int n_hub, cudaStream_t stream) {
  if (which == 0) return launch_gat_fwd<VEC, CH>(p, n_hub, stream);
  switch (ht) {
    case 1: return launch_gat_bwd<VEC, CH, 1>(which, p, n_hub, stream);
    case 2: return launch_gat_bwd<VEC, CH, 2>(which, p, n_hub, stream);
    case 4: return launch_gat_bwd<VEC, CH, 4>(which, p, n_hub, stream);
    default: return launch_gat_bwd<VEC, CH, 8>(which, p, n_hub, stream);
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
#define DGLB_CASE(V, C) if (vec == V && ch == C) return dispatch_gat<V, C>(which, p, ht, n_hub, stream);
  DGLB_CASE(4, 1) DGLB_CASE(4, 2) DGLB_CASE(4, 4)
  DGLB_CASE(2, 1) DGLB_CASE(2, 2) DGLB_CASE(2, 4)
  DGLB_CASE(1, 1) DGLB_CASE(1, 2) DGLB_CASE(1, 4)
#undef DGLB_CASE
  return DGLB_E_UNSUPPORTED;
}

}  // namespace dglb
