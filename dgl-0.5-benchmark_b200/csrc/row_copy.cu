// row_copy.cu -- indexed row copy for the halo exchange of the 1-D row partition ("with
// locality-preserving orderings switch to halo send/recv of only referenced rows").
//
//   dst[idx[i], :] = src[idx[i], :]      i in [0, n_idx)
//
// `src` is a PEER's symmetric-memory buffer mapped into this process (the loads travel over NVLink), `dst` the local
// gather buffer at the peer's slot; `idx` lists -- sorted, unique -- the rows of that peer which this rank's CSC (forward)
// or CSR (backward) slice references.  Rows keep their place, so the column ids of the rank's blocks need no remapping
// and the aggregation result is bit-identical to the one after a full exchange; rows nobody references are simply never
// written (and never read).  A warp moves one row at a time with 16-byte accesses (row_bytes % 16 == 0 and 16-byte
// aligned bases), else 4-byte ones; 8 rows per warp keep several remote requests in flight per lane.
#include "kernels.cuh"

namespace dglb {

template <typename V>
__global__ void __launch_bounds__(kBlockThreads)
copy_rows_indexed_kernel(int64_t n_idx, const int32_t* __restrict__ idx, int64_t row_vecs, const V* __restrict__ src,
                         V* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kBlockThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kBlockThreads) >> 5;
  for (int64_t i = warp; i < n_idx; i += n_warps) {
    const int64_t base = (int64_t)__ldg(idx + i) * row_vecs;
    for (int64_t c = lane; c < row_vecs; c += 32) dst[base + c] = src[base + c];
  }
}

int copy_rows_indexed(int64_t n_idx, const int32_t* idx, int64_t row_bytes, const void* src, void* dst,
                      cudaStream_t stream) {
  if (n_idx == 0 || row_bytes == 0) return DGLB_OK;
  const bool v16 = row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (!v16 && (row_bytes % 4 != 0 || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) != 0)) {
    set_error("copy_rows_indexed: rows must be multiples of 4 bytes and 4-byte aligned (got %lld bytes)", (long long)row_bytes);
    return DGLB_E_INVALID;
  }
  const int64_t warps_per_cta = kBlockThreads / 32;
  int64_t blocks = (n_idx + warps_per_cta - 1) / warps_per_cta;
  const int64_t cap = 148 * 8;   // a few CTAs per SM; rows beyond that are walked by the grid-stride loop
  if (blocks > cap) blocks = cap;
  if (v16)
    copy_rows_indexed_kernel<uint4><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(
        n_idx, idx, row_bytes / 16, static_cast<const uint4*>(src), static_cast<uint4*>(dst));
  else
    copy_rows_indexed_kernel<uint32_t><<<(unsigned)blocks, kBlockThreads, 0, stream>>>(
        n_idx, idx, row_bytes / 4, static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst));
  DGLB_LAUNCH_CHECK("copy_rows_indexed_kernel");
  return DGLB_OK;
}

}  // namespace dglb
