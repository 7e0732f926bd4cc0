// common.cuh -- shared helpers for the sm_100a sparse message-passing kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "../../include/dglb200.h"

namespace dglb {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DGLB_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      ::dglb::set_error(__VA_ARGS__);  \
      return DGLB_E_INVALID;           \
    }                                  \
  } while (0)

#define DGLB_CUDA(call)                                            \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return ::dglb::cuda_fail(e__, #call);  \
  } while (0)

#define DGLB_LAUNCH_CHECK(name)                                       \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return ::dglb::cuda_fail(e__, name);      \
  } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

constexpr int kBlockThreads = 256;

// ------------------------------------------------------------------ small vectors of floats
template <int VEC>
struct FVec {
  float v[VEC];
  __device__ __forceinline__ float at(int i) const { return v[i]; }
};

// read-only (non-coherent) vector load of VEC consecutive floats; p must be VEC*4-byte aligned
template <int VEC>
__device__ __forceinline__ FVec<VEC> ldg_vec(const float* p) {
  FVec<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}



template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const FVec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

template <int VEC>
__device__ __forceinline__ void st_vec_i32(int32_t* p, const int32_t (&r)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<int4*>(p) = make_int4(r[0], r[1], r[2], r[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<int2*>(p) = make_int2(r[0], r[1]);
  } else {
    *p = r[0];
  }
}

// ------------------------------------------------------------------ typed rows: fp32 or bf16 storage
// bf16 storage, fp32 arithmetic: VEC elements per access = 2*VEC bytes (VEC = 8 -> one 128-bit load).
__device__ __forceinline__ float bf16lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf16hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);  // round-to-nearest-even, .x = low half
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <typename T, int VEC>
__device__ __forceinline__ FVec<VEC> ldg_vec_t(const T* p) {
  if constexpr (sizeof(T) == 4) {
    static_assert(VEC <= 4, "fp32 rows use at most 128-bit accesses");
    return ldg_vec<VEC>(reinterpret_cast<const float*>(p));
  } else {
    FVec<VEC> r;
    if constexpr (VEC == 8) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
      r.v[0] = bf16lo(t.x); r.v[1] = bf16hi(t.x); r.v[2] = bf16lo(t.y); r.v[3] = bf16hi(t.y);
      r.v[4] = bf16lo(t.z); r.v[5] = bf16hi(t.z); r.v[6] = bf16lo(t.w); r.v[7] = bf16hi(t.w);
    } else if constexpr (VEC == 4) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
      r.v[0] = bf16lo(t.x); r.v[1] = bf16hi(t.x); r.v[2] = bf16lo(t.y); r.v[3] = bf16hi(t.y);
    } else if constexpr (VEC == 2) {
      const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p));
      r.v[0] = bf16lo(t); r.v[1] = bf16hi(t);
    } else {
      r.v[0] = bf16lo((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)));
    }
    return r;
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void st_vec_t(T* p, const FVec<VEC>& r) {
  if constexpr (sizeof(T) == 4) {
    st_vec<VEC>(reinterpret_cast<float*>(p), r);
  } else {
    if constexpr (VEC == 8) {
      *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(r.v[0], r.v[1]), pack_bf16x2(r.v[2], r.v[3]),
                                               pack_bf16x2(r.v[4], r.v[5]), pack_bf16x2(r.v[6], r.v[7]));
    } else if constexpr (VEC == 4) {
      *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(r.v[0], r.v[1]), pack_bf16x2(r.v[2], r.v[3]));
    } else if constexpr (VEC == 2) {
      *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(r.v[0], r.v[1]);
    } else {
      *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(r.v[0]);
    }
  }
}

// Raw staging registers: gathered rows stay PACKED until they are consumed (bf16: VEC/2 registers
// instead of VEC converted floats -- at VEC = 8 that halves the staging footprint of a load batch).
template <typename T, int VEC>
struct RawVec {
  static constexpr int NW = sizeof(T) == 4 ? VEC : (VEC + 1) / 2;
  uint32_t w[NW];
  template <int I>
  __device__ __forceinline__ float get() const {
    if constexpr (sizeof(T) == 4) return __uint_as_float(w[I]);
    else return (I & 1) ? bf16hi(w[I / 2]) : bf16lo(w[I / 2]);
  }
  __device__ __forceinline__ float at(int i) const {  // i is a compile-time constant after unrolling
    if constexpr (sizeof(T) == 4) return __uint_as_float(w[i]);
    else return (i & 1) ? bf16hi(w[i >> 1]) : bf16lo(w[i >> 1]);
  }
};

template <typename T, int VEC>
__device__ __forceinline__ RawVec<T, VEC> ldg_raw(const T* p) {
  RawVec<T, VEC> r;
  constexpr int BYTES = (int)sizeof(T) * VEC;
  if constexpr (BYTES == 16) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    r.w[0] = t.x; r.w[1] = t.y; r.w[2] = t.z; r.w[3] = t.w;
  } else if constexpr (BYTES == 8) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    r.w[0] = t.x; r.w[1] = t.y;
  } else if constexpr (BYTES == 4) {
    r.w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else {
    r.w[0] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p));
  }
  return r;
}

// Staging type of a gathered chunk: fp32 rows keep the plain float vector (this is the code shape the
// fp32 kernels were tuned with -- the packed form made ptxas interleave loads and FMAs in the narrow
// SDDMM variants, 1.5x slower); bf16 rows stay packed.
template <typename T, int VEC>
struct StageSel { using type = RawVec<T, VEC>; };
template <int VEC>
struct StageSel<float, VEC> { using type = FVec<VEC>; };
template <typename T, int VEC>
using StageVec = typename StageSel<T, VEC>::type;

template <typename T, int VEC>
__device__ __forceinline__ StageVec<T, VEC> ldg_stage(const T* p) {
  if constexpr (sizeof(T) == 4) return ldg_vec<VEC>(reinterpret_cast<const float*>(p));
  else return ldg_raw<T, VEC>(p);
}

template <typename T>
__device__ __forceinline__ float load_scalar_t(const T* p) {
  if constexpr (sizeof(T) == 4) return *reinterpret_cast<const float*>(p);
  else return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
}
template <typename T>
__device__ __forceinline__ void store_scalar_t(T* p, float v) {
  if constexpr (sizeof(T) == 4) *reinterpret_cast<float*>(p) = v;
  else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------ host-side launch geometry
// Lanes per row ("group"): the smallest power of two >= ncols (vector columns), capped at 32.
inline int group_lanes(int64_t ncols) {
  int g = 1;
  while (g < 32 && g < ncols) g <<= 1;
  return g;
}

// widest vector width (floats) usable for rows of `len` floats starting at `ptr`
inline int pick_vec(int64_t len, const void* ptr) {
  uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
  if (len % 4 == 0 && a % 16 == 0) return 4;
  if (len % 2 == 0 && a % 8 == 0) return 2;
  return 1;
}

// bf16 rows: widest element count per access out of {8,4,2,1} (16/8/4/2 bytes)
inline int pick_vec_bf16(int64_t len, const void* ptr) {
  uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
  if (len % 8 == 0 && a % 16 == 0) return 8;
  if (len % 4 == 0 && a % 8 == 0) return 4;
  if (len % 2 == 0 && a % 4 == 0) return 2;
  return 1;
}

inline int min_int(int a, int b) { return a < b ? a : b; }

// right-aligned broadcast description of two trailing feature shapes
struct BcastShape {
  int ndim;
  int64_t lhs[DGLB_MAX_BCAST_NDIM], rhs[DGLB_MAX_BCAST_NDIM], out[DGLB_MAX_BCAST_NDIM];
  int64_t lhs_len, rhs_len, out_len;
};

// returns 0 on success; fills b.  For op DOT the last axis is the reduction axis: out's last
// dim is 1 and *reduce_size receives its length.
int make_bcast(int op, int ndim, const int64_t* lhs_shape, const int64_t* rhs_shape, BcastShape* b,
               int64_t* reduce_size);

}  // namespace dglb
