// Lossless sparse encoding of the 2-D depictions for host-fed screening (SURVEY cfg4 / 8f N2; extension with no reference
// code: the reference keeps depictions as PNG files, Descriptors/convert_smiles_2_img.py, and decodes them on the CPU).
// An RDKit depiction is a white canvas with strokes: ~93 % of the 128x128 pixels of the shipped B3DB depictions are
// (255, 255, 255).  A molecule is stored as
//   mask    16 384 bits (2 048 bytes, little-endian bit order): pixel p differs from white in some channel
//   values  the (R, G, B) triples of the marked pixels in scan order (3 bytes each)
//   offsets running pixel count, so molecule m's triples are values[3*offsets[m] : 3*offsets[m+1]]
// i.e. ~5.5 KB instead of 49 152 bytes: the host -> device copy, which bounds the dense uint8 pipeline at 55 GB/s per GPU
// (and at 23 GB/s per GPU with eight GPUs on one host), shrinks 9x.  This kernel rebuilds the exact uint8 CHW image on the
// device (one block per molecule: popcount per mask word, block-wide exclusive scan, 16 pixels per 128-bit store); the
// first layer then normalises it in its producers as before.
#include "common.cuh"

namespace bbbp {

constexpr int SD_PIX = 128 * 128, SD_WORDS = SD_PIX / 32, SD_THREADS = 256;   // 512 mask words, 2 per thread

__global__ void __launch_bounds__(SD_THREADS) decode_sparse_depictions_kernel(const uint32_t* __restrict__ mask,
                                                                              const uint8_t* __restrict__ values,
                                                                              const int64_t* __restrict__ offsets,
                                                                              uint8_t* __restrict__ out) {
  __shared__ uint32_t warp_tot[SD_THREADS / 32];
  const int m = blockIdx.x, t = threadIdx.x, lane = t % 32, warp = t / 32;
  const uint32_t* mw = mask + (size_t)m * SD_WORDS;
  const uint32_t w0 = mw[2 * t], w1 = mw[2 * t + 1];
  const uint32_t mine = __popc(w0) + __popc(w1);
  uint32_t inc = mine;                                   // inclusive scan over the warp, then over the 8 warp totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t before = inc - mine;
  for (int w = 0; w < warp; ++w) before += warp_tot[w];
  const uint8_t* val = values + 3 * ((size_t)(offsets[m] - offsets[0]) + before);
  uint8_t* o = out + (size_t)m * 3 * SD_PIX + (size_t)t * 64;      // this thread's 64 consecutive pixels, per plane
  // 64 pixels x 3 planes, 4 pixels per word; fully unrolled so that every register index is static (no local memory)
  uint32_t rgb[3][16];
  uint32_t idx = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t r = 0xffffffffu, g = 0xffffffffu, bl = 0xffffffffu;
    const uint32_t nib = ((k < 8 ? w0 : w1) >> ((k & 7) * 4)) & 15u;       // the mask bits of these four pixels
    if (nib) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((nib >> j) & 1u) {
          const uint8_t* v = val + 3 * idx++;
          const uint32_t keep = ~(0xffu << (8 * j));
          r = (r & keep) | ((uint32_t)v[0] << (8 * j));
          g = (g & keep) | ((uint32_t)v[1] << (8 * j));
          bl = (bl & keep) | ((uint32_t)v[2] << (8 * j));
        }
      }
    }
    rgb[0][k] = r, rgb[1][k] = g, rgb[2][k] = bl;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint4* dst = reinterpret_cast<uint4*>(o + (size_t)c * SD_PIX);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(rgb[c][4 * q], rgb[c][4 * q + 1], rgb[c][4 * q + 2], rgb[c][4 * q + 3]);
  }
}

}  // namespace bbbp

// mask: [n][2048] bytes, values: the triples of these n molecules (molecule 0's first), offsets: [n + 1] running pixel
// counts (only differences to offsets[0] are used, so a slice of a longer table works), out: [n][3][128][128] uint8
extern "C" int bbbp_decode_sparse_depictions_u8(const uint8_t* mask, const uint8_t* values, const int64_t* offsets, uint8_t* out,
                                                int n, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(mask && values && offsets && out && n >= 0, "decode_sparse_depictions: bad argument");
  BBBP_CHECK_ARG(((uintptr_t)mask % 4) == 0 && ((uintptr_t)out % 16) == 0, "decode_sparse_depictions: mask must be 4-byte, out 16-byte aligned");
  if (n == 0) return BBBP_OK;
  decode_sparse_depictions_kernel<<<n, SD_THREADS, 0, as_stream(stream)>>>(reinterpret_cast<const uint32_t*>(mask), values, offsets, out);
  return launch_status("decode_sparse_depictions_u8");
}
