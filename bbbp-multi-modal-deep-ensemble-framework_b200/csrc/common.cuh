// Shared helpers for the bbbp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/bbbp_b200.h"

namespace bbbp {

// thread-local error string behind bbbp_last_error()
void set_error(const char* fmt, ...);
// cudaGetLastError() after a launch -> BBBP_OK / BBBP_ECUDA (message recorded)
int launch_status(const char* what);
// adds n to the launch counter behind bbbp_launch_count() (launch_status itself counts one)
void note_launches(int n);

inline cudaStream_t as_stream(bbbp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kWarp = 32;

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 4-byte asynchronous global -> shared copy (zero fill when !valid): lets a thread keep dozens of scalar loads in
// flight without holding them in registers.  Pair with cp_async_wait_all() + a barrier.
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == BBBP_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == BBBP_ACT_TANH) return tanhf(v);
  return v;
}

}  // namespace bbbp

#define BBBP_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      bbbp::set_error(__VA_ARGS__);          \
      return BBBP_EINVAL;                    \
    }                                        \
  } while (0)
