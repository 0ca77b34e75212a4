// The "operand split" format of the tensor-core projection GEMMs (proj_tc.cu), shared with the kernels that produce
// operands directly in it (edge_bwd.cu's finish pass).  A split blob holds one fp32 tensor T [rows, cols] as
//   +0    float    inv_scale = 1/s       +4  uint32  bit pattern of (an upper bound of) max|T|
//   +256  __half   hi[rows, ldp]         hi = fp16(T*s)
//   +256 + plane_bytes   __half lo[rows, ldp]         lo = fp16(T*s - hi)            ldp = round_up(cols, 8)
// with s the power of two that puts the bound in [2^14, 2^15) (fp16 max is 65504).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace b200gat {

constexpr size_t BLOB_HEADER = 256;
inline int64_t pad8(int64_t v) { return (v + 7) / 8 * 8; }
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
inline size_t plane_bytes(int64_t rows, int64_t cols) { return up256(size_t(rows) * size_t(pad8(cols)) * 2); }
inline size_t blob_bytes(int64_t rows, int64_t cols) { return BLOB_HEADER + 2 * plane_bytes(rows, cols); }

struct Blob {
  uint8_t* base; int64_t rows, cols, ldp;
  float* inv_scale() const { return reinterpret_cast<float*>(base); }
  uint32_t* amax_bits() const { return reinterpret_cast<uint32_t*>(base) + 1; }
  __half* hi() const { return reinterpret_cast<__half*>(base + BLOB_HEADER); }
  __half* lo() const { return reinterpret_cast<__half*>(base + BLOB_HEADER + plane_bytes(rows, cols)); }
};
inline Blob make_blob(void* base, int64_t rows, int64_t cols) {
  return Blob{static_cast<uint8_t*>(base), rows, cols, pad8(cols)};
}

// power-of-two scale putting bound * s in [2^14, 2^15); 1 for a zero or non-finite bound
__device__ __forceinline__ float scale_from_amax(uint32_t bits) {
  const float a = __uint_as_float(bits);
  if (!(a > 0.f) || !(a <= 3.4028234e38f)) return 1.f;
  int e;
  frexpf(a, &e);                 // a = m * 2^e, m in [0.5, 1)
  int sh = 15 - e;
  if (sh > 126) sh = 126;
  return ldexpf(1.f, sh);
}

__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// two values at once: one packed convert per pair and plane (F2FP.F16.F32.PACK_AB) instead of one F2F per value
__device__ __forceinline__ void split_half2(float x0, float x1, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
}

// activations fused into operand loads / gradient passes (GATNet.py:63-75: F.elu between the layers)
enum { ACT_NONE = 0, ACT_ELU = 1 };
__device__ __forceinline__ float elu_fwd(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float elu_grad(float x) { return x > 0.f ? 1.f : expf(x); }   // d/dx ELU(x), alpha = 1

// warp-reduced atomic max of a non-negative float's bit pattern
__device__ __forceinline__ void warp_atomic_amax(uint32_t* slot, float v_nonneg) {
  uint32_t m = __float_as_uint(v_nonneg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(slot, m);
}

}  // namespace b200gat
