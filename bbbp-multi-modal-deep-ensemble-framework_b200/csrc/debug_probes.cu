// Micro-benchmarks behind the roofline statements in DESIGN.md (not on any product path).
//   bbbp_debug_tmem_read_probe: bytes per clock one SM moves TMEM -> registers with tcgen05.ld, for 4 or 8 reading warps.
//   The epilogue of the pooled convolutions must read FOUR pre-pool fp32 accumulators per output value (the 2x2 window
//   members live in four TMEM column blocks), so this rate -- not HBM, not the MMA rate -- is the first layer's floor.
#include "common.cuh"
#include "umma.cuh"

namespace bbbp {
using namespace sm100;

__global__ void __launch_bounds__(288) tmem_read_probe_kernel(unsigned long long* __restrict__ out, int warps, int reps, int cols) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 8) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = slot;
  unsigned acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    named_bar_sync(1, warps * 32);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int c = 0; c < cols; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= v[k];
      }
    }
    named_bar_sync(1, warps * 32);
    t1 = clock64();
  }
  if (threadIdx.x == 0) {
    out[0] = (unsigned long long)(t1 - t0);
    out[1] = (unsigned long long)warps * reps * cols * 32 * 4;      // bytes read by the CTA
    out[2] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(base, 512);
  }
}
}  // namespace bbbp

// out: DEVICE array of 3 uint64 = {cycles, bytes, checksum} for ONE CTA (one SM) with `warps` (4 or 8) reading warps
extern "C" int bbbp_debug_tmem_read_probe(void* out, int warps, int reps, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(out && (warps == 4 || warps == 8) && reps > 0, "tmem_read_probe: warps must be 4 or 8");
  tmem_read_probe_kernel<<<1, 288, 0, as_stream(stream)>>>(static_cast<unsigned long long*>(out), warps, reps, 256);
  return launch_status("tmem_read_probe");
}
