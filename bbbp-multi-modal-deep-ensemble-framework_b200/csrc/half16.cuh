// 16-bit operand formats of the tensor-core path.  BBBP_FMT_BF16: 8-bit mantissa, fp32 range.  BBBP_FMT_F16: 11-bit
// mantissa (the precision of a TF32 operand at twice its tensor rate and half its bytes), range +-65504 -- conversions
// saturate instead of producing inf.  "hi + lo" pairs carry an fp32 value as two 16-bit operands (hi = rn(x),
// lo = rn(x - hi)): the strict mode feeds both through the same weights (two MMAs per K step) so that the rounding of
// STRUCTURED activations -- a depiction is mostly one background value, so every background pixel carries the same
// rounding error -- no longer adds up coherently through conv1 -> conv2 -> Linear(65536,128) -> head.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/bbbp_b200.h"

namespace bbbp {

// fp16 conversions saturate at +-65504 in the conversion instruction itself (F2FP.SATFINITE: no extra clamp instructions
// in the latency-bound epilogues)
__device__ __forceinline__ uint32_t f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
__device__ __forceinline__ uint16_t f16_sat(float v) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return r;
}

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (FMT == BBBP_FMT_F16) {
    return f16x2_sat(a, b);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <int FMT>
__device__ __forceinline__ float round16(float a) {
  if constexpr (FMT == BBBP_FMT_F16) {
    const uint16_t h = f16_sat(a);
    return __half2float(*reinterpret_cast<const __half*>(&h));
  } else {
    return __bfloat162float(__float2bfloat16(a));
  }
}
// hi = rn(a, b), lo = rn(a - hi_a, b - hi_b)
template <int FMT>
__device__ __forceinline__ void split16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack16<FMT>(a, b);
  lo = pack16<FMT>(a - round16<FMT>(a), b - round16<FMT>(b));
}
template <int FMT>
__device__ __forceinline__ uint16_t cvt16(float a) {
  if constexpr (FMT == BBBP_FMT_F16) {
    return f16_sat(a);
  } else {
    __nv_bfloat16 h = __float2bfloat16(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
}
// run-time format (kernels whose only format-dependent step is a conversion in a memory-bound loop)
__device__ __forceinline__ uint32_t pack16_rt(float a, float b, int fmt) {
  return fmt == BBBP_FMT_F16 ? pack16<BBBP_FMT_F16>(a, b) : pack16<BBBP_FMT_BF16>(a, b);
}
__device__ __forceinline__ uint16_t cvt16_rt(float a, int fmt) {
  return fmt == BBBP_FMT_F16 ? cvt16<BBBP_FMT_F16>(a) : cvt16<BBBP_FMT_BF16>(a);
}
__device__ __forceinline__ float round16_rt(float a, int fmt) {
  return fmt == BBBP_FMT_F16 ? round16<BBBP_FMT_F16>(a) : round16<BBBP_FMT_BF16>(a);
}
template <int FMT>
__device__ __forceinline__ float2 unpack16(uint32_t v) {
  if constexpr (FMT == BBBP_FMT_F16) return __half22float2(*reinterpret_cast<__half2*>(&v));
  else return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

}  // namespace bbbp
