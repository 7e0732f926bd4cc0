// edge_stage.cu -- "staged" edge order: a per-graph permutation of the edge ids that makes narrow per-edge
// tensors (edge weights (E,1), attention scores (E,H), u_dot_v / edge_softmax results) cheap to address from a
// kernel that walks the CSC / CSR of a graph whose edges were NOT created in that order.
//
// The problem (profiles/r01_notes.md section 10, ncu): a 4..16-byte access W[eid[j]] at a random edge id pulls a
// 128-byte line from DRAM -- products, u_mul_e with (E,1) weights: 8 GB of DRAM reads to fetch 0.25 GB of
// weights; edge_softmax H = 4: 5.3x its algorithmic bytes; u_dot_v: a 32-byte read-modify-write per 4-byte
// store.  Upstream has the same access (cuda/spmm.cuh: `eid = data[j]`, sddmm.cuh: out[eid]).
//
// The fix: cut the CSR positions into buckets of 2^log2_bucket consecutive positions (32 K) and, inside every
// bucket, order the slots by EDGE ID.  That defines, for every edge e, a staged slot stage_pos[e] in the same
// bucket as its CSR position, and for every CSR position j the slot slot[j] = stage_pos[eid[j]].  Then
//   * moving a tensor between edge-id order and staged order (dglb_edge_stage) is ONE pass whose edge-id side
//     is fully coalesced and whose staged side walks ~E/32K bucket cursors forward: every 128-byte line on
//     the staged side is completed within a short window and stays in L2 meanwhile (~12 B of DRAM traffic per
//     4-byte element instead of ~128 B);
//   * the compute kernels are unchanged: they are handed `slot` in place of the edge-id array and the staged
//     tensor in place of the edge-id-ordered one, so W_staged[slot[j]] touches a 128 KB window around j.
// One-off per (graph, format), like the CSC itself; caller-owned arrays, no library state.
#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace dglb {

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

__global__ void stage_keys_kernel(const int32_t* __restrict__ eids, int32_t* __restrict__ key_by_edge,
                                  int32_t* __restrict__ iota, int64_t n, int log2_bucket) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  key_by_edge[__ldg(eids + j)] = (int32_t)(j >> log2_bucket);   // bucket of the edge's CSR position, by edge id
  iota[j] = (int32_t)j;
}

__global__ void stage_invert_kernel(const int32_t* __restrict__ order, int32_t* __restrict__ stage_pos, int64_t n) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) stage_pos[__ldg(order + s)] = (int32_t)s;
}

__global__ void stage_slot_kernel(const int32_t* __restrict__ eids, const int32_t* __restrict__ stage_pos,
                                  int32_t* __restrict__ slot, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) slot[j] = __ldg(stage_pos + __ldg(eids + j));
}

static int stage_key_bits(int64_t nnz, int log2_bucket) {
  const int64_t buckets = ((nnz - 1) >> log2_bucket) + 1;
  int bits = 1;
  while (bits < 32 && (1LL << bits) < buckets) ++bits;
  return bits;
}

size_t edge_stage_plan_workspace_bytes(int64_t nnz, int log2_bucket) {
  if (nnz <= 0) return 0;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs<int32_t, int32_t>(nullptr, cub_bytes, nullptr, nullptr, nullptr, nullptr, (int)nnz, 0,
                                                    stage_key_bits(nnz, log2_bucket));
  return 4 * align256((size_t)nnz * 4) + align256(cub_bytes) + 256;
}

int edge_stage_plan(int64_t nnz, const int32_t* eids, int log2_bucket, int32_t* stage_pos, int32_t* slot,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (nnz == 0) return DGLB_OK;
  if (nnz >= (1LL << 31)) { set_error("edge_stage_plan: nnz must be < 2^31"); return DGLB_E_UNSUPPORTED; }
  if (!workspace || workspace_bytes < edge_stage_plan_workspace_bytes(nnz, log2_bucket)) {
    set_error("edge_stage_plan: workspace too small");
    return DGLB_E_WORKSPACE;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const size_t a = align256((size_t)nnz * 4);
  int32_t* keys_in = reinterpret_cast<int32_t*>(ws);
  int32_t* keys_out = reinterpret_cast<int32_t*>(ws + a);
  int32_t* iota = reinterpret_cast<int32_t*>(ws + 2 * a);
  int32_t* order = reinterpret_cast<int32_t*>(ws + 3 * a);   // staged slot -> edge id
  void* cub_ws = ws + 4 * a;
  size_t cub_bytes = workspace_bytes - 4 * a;
  const unsigned blocks = (unsigned)((nnz + 255) / 256);
  stage_keys_kernel<<<blocks, 256, 0, stream>>>(eids, keys_in, iota, nnz, log2_bucket);
  DGLB_LAUNCH_CHECK("stage_keys_kernel");
  // stable sort of the edge ids by bucket: inside a bucket the slots end up in increasing edge id
  DGLB_CUDA((cub::DeviceRadixSort::SortPairs<int32_t, int32_t>(cub_ws, cub_bytes, keys_in, keys_out, iota, order, (int)nnz,
                                                               0, stage_key_bits(nnz, log2_bucket), stream)));
  stage_invert_kernel<<<blocks, 256, 0, stream>>>(order, stage_pos, nnz);
  DGLB_LAUNCH_CHECK("stage_invert_kernel");
  stage_slot_kernel<<<blocks, 256, 0, stream>>>(eids, stage_pos, slot, nnz);
  DGLB_LAUNCH_CHECK("stage_slot_kernel");
  return DGLB_OK;
}

// rows of W 4-byte words; VW words per access.  TO_STAGED: dst[stage_pos[e]] = src[e], else dst[e] = src[stage_pos[e]]
template <int VW, bool TO_STAGED>
__global__ void __launch_bounds__(256) edge_stage_kernel(const int32_t* __restrict__ stage_pos,
                                                         const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                         int64_t nnz, int chunks /* W / VW */) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nnz * chunks) return;
  const int64_t e = idx / chunks;
  const int c = (int)(idx - e * chunks);
  const int64_t s = __ldg(stage_pos + e);
  const int64_t from = ((TO_STAGED ? e : s) * chunks + c) * VW;
  const int64_t to = ((TO_STAGED ? s : e) * chunks + c) * VW;
  if constexpr (VW == 4) {
    *reinterpret_cast<uint4*>(dst + to) = __ldg(reinterpret_cast<const uint4*>(src + from));
  } else if constexpr (VW == 2) {
    *reinterpret_cast<uint2*>(dst + to) = __ldg(reinterpret_cast<const uint2*>(src + from));
  } else {
    dst[to] = __ldg(src + from);
  }
}

int edge_stage_move(int to_staged, int64_t nnz, int64_t row_words, const int32_t* stage_pos, const void* src, void* dst,
                    cudaStream_t stream) {
  if (nnz == 0 || row_words == 0) return DGLB_OK;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  int vw = 1;
  if (row_words % 4 == 0 && a % 16 == 0) vw = 4;
  else if (row_words % 2 == 0 && a % 8 == 0) vw = 2;
  const int64_t chunks = row_words / vw;
  const int64_t total = nnz * chunks;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL || chunks > (1 << 20)) { set_error("edge_stage: problem too large"); return DGLB_E_UNSUPPORTED; }
  const uint32_t* s = static_cast<const uint32_t*>(src);
  uint32_t* d = static_cast<uint32_t*>(dst);
#define DGLB_STAGE(V) \
  if (vw == V) { \
    if (to_staged) edge_stage_kernel<V, true><<<(unsigned)blocks, 256, 0, stream>>>(stage_pos, s, d, nnz, (int)chunks); \
    else edge_stage_kernel<V, false><<<(unsigned)blocks, 256, 0, stream>>>(stage_pos, s, d, nnz, (int)chunks); \
  }
  DGLB_STAGE(4) DGLB_STAGE(2) DGLB_STAGE(1)
#undef DGLB_STAGE
  DGLB_LAUNCH_CHECK("edge_stage_kernel");
  return DGLB_OK;
}

}  // namespace dglb
