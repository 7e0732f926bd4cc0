// ring.cu -- wide-row gathers through a per-warp ring of whole-row bulk copies (sm_100a).
//
// The row-per-group kernels of spmm.cu / sddmm.cu keep a gather batch in REGISTERS (8 x 64/128-bit loads
// per lane), so the bytes a warp has in flight are bounded by its register budget: at D = 602 (2 408-byte
// rows, 8-byte aligned only) that is 2 KB per warp, 64 KB per SM, and ncu showed the kernel waiting on the
// long scoreboard at 42 % active warps and 0.82 of the measured HBM peak.  Here the in-flight depth is paid
// in SHARED MEMORY instead:
//
//   * persistent grid (resident CTAs x SMs); every warp owns a contiguous range of destination rows that holds
//     an equal share of the edges (binary search over indptr: nnz-balanced, rows are never split), and walks
//     the range's edges as ONE stream across row boundaries;
//   * one elected lane requests each neighbour's whole feature row with a single `cp.async.bulk`
//     (global -> shared, completion on an mbarrier) into a ring of S slots, S rows ahead of the consumer --
//     at D = 602 that is 5 x 2.4 KB in flight per warp, 16 warps per SM, ~190 KB per SM, no registers;
//     the copy starts at the row's 16-byte-aligned floor and covers the row rounded up to 16 bytes, so rows
//     that are only 4- or 8-byte aligned (D = 602 fp32, D = 602 bf16) need no padded copy of X;
//   * all 32 lanes then read the landed row from shared memory with conflict-free vector loads and add it to
//     register accumulators strictly in CSR order (bit-identical to the CPU kernel's sequential order, like
//     the row-per-group kernel), or -- DOT mode, gsddmm u_dot_v -- multiply it with the destination row held
//     in registers and reduce across the warp;
//   * rows above the hub threshold are skipped (both by the producer and the consumer cursor) and left to the
//     segmented hub kernels of spmm.cu / sddmm.cu, exactly like the row-per-group kernels do.
//
// Replaces, for wide rows, the same upstream kernels as spmm.cu / sddmm.cu (dmlc/dgl@0.6.1
// src/array/cuda/spmm.cuh::SpMMCsrKernel / CusparseCsrmm2, sddmm.cuh::SDDMMCooKernel); reached from
// kernel/dgl-new.py:20,39 at the widths BASELINE.json names (256, 602) and from every SAGE aggregation wider
// than 128 floats.
#include <cstdlib>

#include "bulk.cuh"
#include "kernels.cuh"

namespace dglb {

constexpr int kRingThreads = 256;
constexpr int kRingWarps = kRingThreads / 32;

struct RingParams {
  const int32_t* __restrict__ indptr;
  const int32_t* __restrict__ balance;  // [n_rows+1] prefix sum the warps' row ranges are cut on (indptr, or the
                                        // prefix sum of the non-hub rows' nnz when hub rows are left to other kernels)
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ eids;     // DOT: edge id of each CSR position (null = identity)
  const unsigned char* __restrict__ X;  // gathered rows: n_cols rows of row_bytes
  const unsigned char* __restrict__ V;  // DOT: one row per destination row
  unsigned char* __restrict__ out;      // SPMM: (n_rows, D); DOT: (nnz) in edge-id order
  const float* __restrict__ row_scale;
  int64_t n_rows, n_cols, nnz;
  int D, ncols, row_bytes, slot_bytes, S;
  int hub_threshold, accumulate;
};

// first row r in [0, n_rows] with indptr[r] >= target (warp-uniform; every lane runs the same search)
__device__ __forceinline__ int64_t row_lower_bound(const int32_t* __restrict__ indptr, int64_t n_rows, int64_t target) {
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(indptr + mid) >= target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// position in the edge stream of a row range: skips empty rows and rows left to the hub kernels
struct EdgeCursor {
  int64_t row, row_end;
  int pos, end;
  bool valid;
  __device__ __forceinline__ void settle(const RingParams& p) {
    while (pos >= end) {
      ++row;
      if (row >= row_end) { valid = false; return; }
      const int s = __ldg(p.indptr + row), e = __ldg(p.indptr + row + 1);
      if (e - s > p.hub_threshold) { pos = end = e; } else { pos = s; end = e; }
    }
    valid = true;
  }
};

template <typename T, int VEC>
__device__ __forceinline__ void lds_vec(const unsigned char* src, float (&v)[VEC]) {
  constexpr int BYTES = (int)sizeof(T) * VEC;
  if constexpr (sizeof(T) == 4) {
    if constexpr (BYTES == 16) {
      const float4 t = *reinterpret_cast<const float4*>(src);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (BYTES == 8) {
      const float2 t = *reinterpret_cast<const float2*>(src);
      v[0] = t.x; v[1] = t.y;
    } else {
      v[0] = *reinterpret_cast<const float*>(src);
    }
  } else {
    if constexpr (BYTES == 16) {
      const uint4 t = *reinterpret_cast<const uint4*>(src);
      v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
      v[4] = bf16lo(t.z); v[5] = bf16hi(t.z); v[6] = bf16lo(t.w); v[7] = bf16hi(t.w);
    } else if constexpr (BYTES == 8) {
      const uint2 t = *reinterpret_cast<const uint2*>(src);
      v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
    } else {
      const uint32_t t = *reinterpret_cast<const uint32_t*>(src);
      v[0] = bf16lo(t); v[1] = bf16hi(t);
    }
  }
}

// NCH = vector columns per lane (lane l owns columns l, l+32, ...); DOT = gsddmm u_dot_v instead of gspmm sum.
// The per-edge path is kept short on purpose (ncu, first version: 200 warp instructions per edge, issue slots 54 %
// busy at 16 resident warps, 92 registers): slot indices and phases are carried as wrapping counters instead of
// `% S` and `/ S`, only the last vector column of a lane is range-checked, shared-memory offsets are immediates.
template <typename T, int VEC, int NCH, bool DOT>
__global__ void __launch_bounds__(kRingThreads, 3) ring_kernel(const RingParams p) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int slot_bytes = p.slot_bytes;
  unsigned char* slots = ring_smem + (size_t)warp * S * slot_bytes;
  unsigned char* tail = ring_smem + (size_t)kRingWarps * S * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail) + warp * S;
  int* meta = reinterpret_cast<int*>(tail + (size_t)kRingWarps * S * 8) + warp * S;
  const uint32_t slots_u32 = smem_u32(slots), bars_u32 = smem_u32(bars);

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars_u32 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // ---- this warp's rows: an equal share of the (non-hub) edges, cut at row boundaries
  const int64_t gw = (int64_t)blockIdx.x * kRingWarps + warp, nw = (int64_t)gridDim.x * kRingWarps;
  const int64_t work = __ldg(p.balance + p.n_rows);
  const int64_t r0 = gw == 0 ? 0 : row_lower_bound(p.balance, p.n_rows, work * gw / nw);
  const int64_t r1 = gw + 1 == nw ? p.n_rows : row_lower_bound(p.balance, p.n_rows, work * (gw + 1) / nw);
  if (r0 >= r1) return;

  EdgeCursor pc;  // producer cursor
  pc.row = r0 - 1; pc.row_end = r1; pc.pos = pc.end = 0;
  pc.settle(p);
  // the source id of the cursor's edge is fetched one step ahead (ncu: the producer's index load was the kernel's top
  // stall -- long scoreboard 10.8 per issue -- because every copy request waited for its own index)
  int pc_col = pc.valid ? __ldg(p.indices + pc.pos) : 0;
  int in_flight = 0;            // copies requested and not yet consumed
  int pslot = 0, cslot = 0;     // next slot to fill / to consume
  uint32_t cphase = 0;          // parity the consumer waits for on cslot
  const int row_bytes = p.row_bytes;
  const int last_col = (int)p.n_cols - 1;
  // the highest row of X: would a copy rounded up to 16 bytes read past the end of the tensor?
  const bool last_unsafe = ((reinterpret_cast<uintptr_t>(p.X) + (uint64_t)p.n_cols * row_bytes) & 15) != 0;
  const int lane_off = lane * VEC * (int)sizeof(T);
  constexpr int kColStride = 32 * VEC * (int)sizeof(T);

  auto fill = [&]() {
    while (pc.valid && in_flight < S) {
      const int c = pc_col;
      const unsigned char* g = p.X + (int64_t)c * row_bytes;
      const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
      if (c == last_col && last_unsafe) {   // copy it by hand (warp-uniform branch)
        unsigned char* d = slots + pslot * slot_bytes + shift;
        for (int i = lane * 4; i < row_bytes; i += 128)
          *reinterpret_cast<uint32_t*>(d + i) = __ldg(reinterpret_cast<const uint32_t*>(g + i));
        __syncwarp();
        if (lane == 0) { meta[pslot] = (int)shift; mbar_arrive(bars_u32 + 8 * pslot); }
      } else if (lane == 0) {
        const uint32_t bytes = (shift + row_bytes + 15u) & ~15u;
        meta[pslot] = (int)shift;
        mbar_expect_tx(bars_u32 + 8 * pslot, bytes);
        bulk_g2s(slots_u32 + pslot * slot_bytes, g - shift, bytes, bars_u32 + 8 * pslot);
      }
      ++in_flight;
      if (++pslot == S) pslot = 0;
      if (++pc.pos >= pc.end) pc.settle(p);
      if (pc.valid) pc_col = __ldg(p.indices + pc.pos);
    }
  };

  const int D = p.D;
  const int ncols = p.ncols;
  for (int64_t row = r0; row < r1; ++row) {
    const int s = __ldg(p.indptr + row), e = __ldg(p.indptr + row + 1);
    if (e - s > p.hub_threshold) continue;  // the hub kernels write this row / these edges
    float acc[NCH][VEC];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    if constexpr (DOT) {
      if (e > s) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int vc = lane + 32 * c;
          if (c < NCH - 1 || vc < ncols) {
            const FVec<VEC> t = ldg_vec_t<T, VEC>(reinterpret_cast<const T*>(p.V) + row * (int64_t)D + (int64_t)vc * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] = t.v[v];
          }
        }
      }
    }
    float pending = 0.f;  // DOT: lane k holds the result of the k-th edge of the current batch of 32
    for (int j = s; j < e; ++j) {
      fill();
      mbar_wait(bars_u32 + 8 * cslot, cphase);
      const unsigned char* src = slots + cslot * slot_bytes + meta[cslot] + lane_off;
      if constexpr (!DOT) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < NCH - 1 || lane + 32 * c < ncols) {
            float x[VEC];
            lds_vec<T, VEC>(src + c * kColStride, x);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] = __fadd_rn(acc[c][v], x[v]);
          }
        }
      } else {
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < NCH - 1 || lane + 32 * c < ncols) {
            float x[VEC];
            lds_vec<T, VEC>(src + c * kColStride, x);
#pragma unroll
            for (int v = 0; v < VEC; ++v) part = fmaf(acc[c][v], x[v], part);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL_MASK, part, o);
        const int k = (j - s) & 31;
        if (lane == k) pending = part;
        if (k == 31 || j + 1 == e) {  // flush up to 32 results with one (coalesced when eids is null) store
          const int jj = j - k + lane;
          if (lane <= k) {
            const int64_t eid = p.eids ? __ldg(p.eids + jj) : jj;
            store_scalar_t<T>(reinterpret_cast<T*>(p.out) + eid, pending);
          }
        }
      }
      __syncwarp();  // every lane is done with the slot before the producer lane refills it
      --in_flight;
      if (++cslot == S) { cslot = 0; cphase ^= 1u; }
    }
    if constexpr (!DOT) {
      const float scale = p.row_scale ? __ldg(p.row_scale + row) : 1.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int vc = lane + 32 * c;
        if (c < NCH - 1 || vc < ncols) {
          FVec<VEC> o;
          const int64_t off = row * (int64_t)D + (int64_t)vc * VEC;
#pragma unroll
          for (int v = 0; v < VEC; ++v) o.v[v] = p.row_scale ? __fdiv_rn(acc[c][v], scale) : acc[c][v];
          if (p.accumulate) {
            const FVec<VEC> prev = ldg_vec_t<T, VEC>(reinterpret_cast<const T*>(p.out) + off);
#pragma unroll
            for (int v = 0; v < VEC; ++v) o.v[v] = __fadd_rn(prev.v[v], o.v[v]);
          }
          st_vec_t<T, VEC>(reinterpret_cast<T*>(p.out) + off, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ host side
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <typename T, int VEC, int NCH, bool DOT>
static int launch_ring(const RingParams& p, size_t smem, cudaStream_t stream) {
  auto kern = ring_kernel<T, VEC, NCH, DOT>;
  DGLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  DGLB_CUDA(cudaGetDevice(&dev));
  DGLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DGLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRingThreads, smem));
  if (per_sm < 1) { set_error("ring kernel does not fit on an SM (%zu bytes of shared memory)", smem); return DGLB_E_UNSUPPORTED; }
  kern<<<sms * per_sm, kRingThreads, smem, stream>>>(p);
  DGLB_LAUNCH_CHECK("ring_kernel");
  return DGLB_OK;
}

template <typename T, int VEC, bool DOT>
static int dispatch_nch(const RingParams& p, size_t smem, cudaStream_t stream) {
  const int nch = (p.ncols + 31) / 32;
#define DGLB_RING_CASE(N) \
  if constexpr (VEC * N <= 64) { if (nch <= N) return launch_ring<T, VEC, N, DOT>(p, smem, stream); }
  DGLB_RING_CASE(1) DGLB_RING_CASE(2) DGLB_RING_CASE(3) DGLB_RING_CASE(4) DGLB_RING_CASE(6) DGLB_RING_CASE(8)
  DGLB_RING_CASE(10) DGLB_RING_CASE(12) DGLB_RING_CASE(16)
#undef DGLB_RING_CASE
  return DGLB_E_UNSUPPORTED;
}

// Rows of the gspmm(copy_lhs, sum) / gsddmm(dot) that are not hub rows, through the ring kernel.
// Returns DGLB_E_UNSUPPORTED (without setting an error) when the shape is outside the ring kernel's range, in which
// case the caller runs the row-per-group kernel instead.
int ring_rows(bool dot, int dtype, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* indptr,
              const int32_t* indices, const int32_t* eids, const void* X, const void* V, int64_t D, void* out,
              const float* row_scale, int accumulate, int hub_threshold, const int32_t* light_indptr,
              cudaStream_t stream) {
  // Where the ring pays (profiles/r02_notes.md, reddit / products shapes): the copy engine retires ~5.3 G bulk copies
  // per second chip-wide whatever their size, so a row must be >= ~2 KB for the copies not to be the bound; and rows
  // that are not 16-byte aligned (D = 602 fp32 / bf16), which the register path can only fetch with 8- or 4-byte
  // loads, win from 1 KB up.  Everything else stays on the row-per-group kernels.  The knobs are read per call so a
  // single process can A/B them (examples/ring_tune.py, tests/test_gpu_ring.py).
  const int esz = dtype == DGLB_BF16 ? 2 : 4;
  const int64_t row_bytes = D * esz;
  // (measured with the short per-edge path: bf16 D = 602, 1 204-byte rows: gspmm 2.08 ms vs 4.43 ms on the padded register
  //  path, u_dot_v 3.25 vs 4.08 ms; bf16 D = 1000, 2 000-byte aligned rows: 3.11 vs 3.92 ms, 3.37 vs 4.16 ms)
  const int min_bytes = env_int("DGLB_RING_MIN_BYTES", (row_bytes % 16) ? 1024 : 1792);
  const int min_nnz = env_int("DGLB_RING_MIN_NNZ", 1 << 18);
  const int force_s = env_int("DGLB_RING_STAGES", 0);
  // Depth: throughput peaks at ~115-130 KB of rows in flight per SM and DROPS beyond it (D = 602: 3.9 ms at 48 rows in
  // flight, 4.1 ms at 80, 5.5 ms at 72 rows over 24 warps), so the ring is 2 deep and the CTA count per SM does the rest:
  // 3 CTAs (24 warps) for rows up to 3 KB, 2 CTAs (forced through the shared-memory request) beyond.
  const int smem_budget = env_int("DGLB_RING_SMEM", 0);
  if (row_bytes < min_bytes || nnz < min_nnz || n_rows < 1 || n_cols < 1) return DGLB_E_UNSUPPORTED;
  if (row_bytes % 4 || (reinterpret_cast<uintptr_t>(X) & 15) || nnz >= (1LL << 31)) return DGLB_E_UNSUPPORTED;
  int vec;
  if (dtype == DGLB_BF16) vec = D % 8 == 0 ? 8 : (D % 4 == 0 ? 4 : 2);
  else vec = D % 4 == 0 ? 4 : (D % 2 == 0 ? 2 : 1);
  const uintptr_t oa = reinterpret_cast<uintptr_t>(dot ? V : out);
  while (vec > 1 && (oa % (vec * esz))) vec >>= 1;   // the destination-side rows are accessed from global memory
  if (dtype == DGLB_BF16 && vec < 2) return DGLB_E_UNSUPPORTED;
  const int64_t ncols = D / vec;
  if (ncols > 16 * 32) return DGLB_E_UNSUPPORTED;
  RingParams p;
  p.indptr = indptr; p.indices = indices; p.eids = eids;
  p.balance = light_indptr ? light_indptr : indptr;
  p.X = static_cast<const unsigned char*>(X); p.V = static_cast<const unsigned char*>(V);
  p.out = static_cast<unsigned char*>(out); p.row_scale = row_scale;
  p.n_rows = n_rows; p.n_cols = n_cols; p.nnz = nnz;
  p.D = (int)D; p.ncols = (int)ncols; p.row_bytes = (int)row_bytes;
  p.slot_bytes = (int)((row_bytes + 15) / 16 * 16 + 16);
  int S = force_s > 0 ? force_s : (smem_budget > 0 ? (smem_budget / kRingWarps) / (p.slot_bytes + 12) : 2);
  if (S > 16) S = 16;
  if (S < 2) return DGLB_E_UNSUPPORTED;
  p.S = S;
  p.hub_threshold = hub_threshold; p.accumulate = accumulate;
  size_t smem = (size_t)kRingWarps * S * (p.slot_bytes + 8 + 4);
  if (smem > 110 * 1024) return DGLB_E_UNSUPPORTED;                       // rows beyond ~6.8 KB: register path
  if (force_s <= 0 && smem_budget <= 0 && row_bytes > 3072 && smem < 76 * 1024) smem = 76 * 1024;   // 2 CTAs per SM
#define DGLB_RING_VEC(TT, VV) \
  if (vec == VV) return dot ? dispatch_nch<TT, VV, true>(p, smem, stream) : dispatch_nch<TT, VV, false>(p, smem, stream);
  if (dtype == DGLB_BF16) {
    DGLB_RING_VEC(__nv_bfloat16, 8) DGLB_RING_VEC(__nv_bfloat16, 4) DGLB_RING_VEC(__nv_bfloat16, 2)
  } else {
    DGLB_RING_VEC(float, 4) DGLB_RING_VEC(float, 2) DGLB_RING_VEC(float, 1)
  }
#undef DGLB_RING_VEC
  return DGLB_E_UNSUPPORTED;
}

}  // namespace dglb
