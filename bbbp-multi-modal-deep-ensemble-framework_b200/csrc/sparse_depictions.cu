// Lossless sparse encoding of the 2-D depictions for host-fed screening (SURVEY cfg4 / 8f N2; extension with no reference
// code: the reference keeps depictions as PNG files, Descriptors/convert_smiles_2_img.py, and decodes them on the CPU).
// An RDKit depiction is a white canvas with strokes: ~93 % of the 128x128 pixels of the shipped B3DB depictions are
// (255, 255, 255).  A molecule is stored as
//   mask    16 384 bits (2 048 bytes, little-endian bit order): pixel p differs from white in some channel
//   values  the (R, G, B) triples of the marked pixels in scan order (3 bytes each)
//   offsets running pixel count, so molecule m's triples are values[3*offsets[m] : 3*offsets[m+1]]
// i.e. ~5.5 KB instead of 49 152 bytes: the host -> device copy, which bounds the dense uint8 pipeline at 55 GB/s per GPU
// (and at 23 GB/s per GPU with eight GPUs on one host), shrinks 9x.  This kernel rebuilds the exact uint8 CHW image on the
// device (one block per molecule: popcount per 16-bit mask word, block-wide exclusive scans, 16 pixels per 128-bit store); the
// first layer then normalises it in its producers as before.
#include "common.cuh"

namespace bbbp {

constexpr int SD_PIX = 128 * 128, SD_GROUPS = SD_PIX / 16, SD_THREADS = 256, SD_Q = SD_GROUPS / SD_THREADS;   // 1 024 groups of 16 pixels

// Thread t owns the pixel groups G = t + 256 * q (q = 0..3; a group = 16 consecutive pixels = one 16-bit mask word = one
// 128-bit store per plane), so a warp's stores are 512 contiguous bytes.  The triples of group G start after those of all
// earlier groups: per-segment (q) block scans of the group popcounts, two 16-bit counters packed per 32-bit word.
__global__ void __launch_bounds__(SD_THREADS) decode_sparse_depictions_kernel(const uint16_t* __restrict__ mask,
                                                                              const uint8_t* __restrict__ values,
                                                                              const int64_t* __restrict__ offsets,
                                                                              uint8_t* __restrict__ out) {
  __shared__ uint32_t warp_tot[2][SD_THREADS / 32];
  const int m = blockIdx.x, t = threadIdx.x, lane = t % 32, warp = t / 32;
  const uint16_t* mw = mask + (size_t)m * SD_GROUPS;
  uint32_t w[SD_Q];
#pragma unroll
  for (int q = 0; q < SD_Q; ++q) w[q] = mw[t + SD_THREADS * q];
  // counts of segments (0, 1) in A, (2, 3) in B; a segment holds at most 4 096 marked pixels: 16 bits are enough
  const uint32_t mineA = __popc(w[0]) | (__popc(w[1]) << 16), mineB = __popc(w[2]) | (__popc(w[3]) << 16);
  uint32_t incA = mineA, incB = mineB;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, incA, o), b = __shfl_up_sync(0xffffffffu, incB, o);
    if (lane >= o) incA += a, incB += b;
  }
  if (lane == 31) warp_tot[0][warp] = incA, warp_tot[1][warp] = incB;
  __syncthreads();
  uint32_t befA = incA - mineA, befB = incB - mineB, totA = 0, totB = 0;
#pragma unroll
  for (int k = 0; k < SD_THREADS / 32; ++k) {
    const uint32_t a = warp_tot[0][k], b = warp_tot[1][k];
    if (k < warp) befA += a, befB += b;
    totA += a, totB += b;
  }
  const uint32_t seg_tot[3] = {totA & 0xffffu, totA >> 16, totB & 0xffffu};
  uint32_t before[SD_Q] = {befA & 0xffffu, befA >> 16, befB & 0xffffu, befB >> 16};
  before[1] += seg_tot[0];
  before[2] += seg_tot[0] + seg_tot[1];
  before[3] += seg_tot[0] + seg_tot[1] + seg_tot[2];
  const uint8_t* val0 = values + 3 * (size_t)(offsets[m] - offsets[0]);
  uint8_t* obase = out + (size_t)m * 3 * SD_PIX;
#pragma unroll
  for (int q = 0; q < SD_Q; ++q) {
    const uint8_t* val = val0 + 3 * (size_t)before[q];
    uint32_t rgb[3][4];
    uint32_t idx = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                          // 4 pixels per output word
      uint32_t r = 0xffffffffu, g = 0xffffffffu, bl = 0xffffffffu;
      const uint32_t nib = (w[q] >> (4 * k)) & 15u;
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
    const size_t at = (size_t)(t + SD_THREADS * q) * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<uint4*>(obase + (size_t)c * SD_PIX + at) = make_uint4(rgb[c][0], rgb[c][1], rgb[c][2], rgb[c][3]);
  }
}

}  // namespace bbbp

// mask: [n][2048] bytes, values: the triples of these n molecules (molecule 0's first), offsets: [n + 1] running pixel
// counts (only differences to offsets[0] are used, so a slice of a longer table works), out: [n][3][128][128] uint8
extern "C" int bbbp_decode_sparse_depictions_u8(const uint8_t* mask, const uint8_t* values, const int64_t* offsets, uint8_t* out,
                                                int n, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(mask && values && offsets && out && n >= 0, "decode_sparse_depictions: bad argument");
  BBBP_CHECK_ARG(((uintptr_t)mask % 2) == 0 && ((uintptr_t)out % 16) == 0, "decode_sparse_depictions: mask must be 2-byte, out 16-byte aligned");
  if (n == 0) return BBBP_OK;
  decode_sparse_depictions_kernel<<<n, SD_THREADS, 0, as_stream(stream)>>>(reinterpret_cast<const uint16_t*>(mask), values, offsets, out);
  return launch_status("decode_sparse_depictions_u8");
}
