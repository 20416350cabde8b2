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

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property: a process driving several GPUs must set it
// once on each of them.  `static PerDeviceOnce once; if (once.first()) { ... }` at the launch site.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};
// SM count of the CURRENT device (cached per device ordinal)
int current_sm_count();

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

// Per-site dropout seed (host, baked into a CUDA graph) combined with the per-step seed that lives in device memory.
// NOT additive: with seed_site = a + i*C and seed_step = a + s*C a plain sum makes (site i, step s) and
// (site i+1, step s-1) share one Philox key, i.e. the same mask travels down the layer stack on successive steps.
// splitmix64 of the step value, xor-ed into the site seed and mixed again, keys every (site, step) pair differently.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t mix_seed(uint64_t seed, const uint64_t* __restrict__ seed_dev) {
  return seed_dev ? splitmix64(seed ^ splitmix64(*seed_dev)) : seed;
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
