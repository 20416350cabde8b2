// Helper kernels of the MIXED-PRECISION TRAINING path of the image branch (nn.Conv2d + ReLU + MaxPool2d twice, then
// Flatten + Linear, 20250113.py:85-93; forward + loss.backward() of :188-190).  Every contraction of that path -- both
// convolutions forward, their data and weight gradients, the Linear(65536,128) forward and both of its gradients -- runs on
// the tcgen05 GEMM (gemm_umma.cu): convolutions as im2col rows x weights, the weight gradients as
//   dW[co][(tap, ci)] = sum_pixels dpre[pixel][co] * cols[pixel][(tap, ci)]
// with BOTH operands read in place in MN-major storage (bbbp_gemm16_tn), the data gradient as im2col(dpre) x flipped weights.
// What is left for this file is the memory-bound glue around those GEMMs, all NHWC, 16-bit activations (bf16 or fp16):
//   * 2x2 max-pool WITH the arg-max the backward pass needs,
//   * its backward fused with the ReLU mask: pooled fp32 gradient -> pre-pool 16-bit gradient (the GEMM operand) and the
//     masked pooled gradient (whose column sum is the bias gradient),
//   * weight-layout conversions between nn.Conv2d's (Cout, Cin, 3, 3) and the GEMM's (Cout, 9 * Cpad) rows, flipped /
//     transposed weights for the data gradient, (H, W, C) <-> (C, H, W) ordering of the Linear weight gradient,
//   * planar fp32 image -> NHWC with 8 channels per pixel in either 16-bit format.
#include "common.cuh"
#include "half16.cuh"

namespace bbbp {
namespace ctrain {

// y = 2x2 max-pool of x (NHWC, C % 8 == 0), arg = index 2*i + j of the FIRST maximum in torch's scan order
template <int FMT>
__global__ void __launch_bounds__(256) maxpool_argmax_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                             uint2* __restrict__ arg, size_t total, int H, int W, int C8) {
  const int OW = W / 2, OH = H / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % OW), oy = (int)((i / ((size_t)C8 * OW)) % OH);
    const size_t n = i / ((size_t)C8 * OW * OH);
    const uint4* p = x + ((n * H + 2 * oy) * W + 2 * ox) * C8 + c;
    const uint4 q[4] = {p[0], p[C8], p[(size_t)W * C8], p[(size_t)W * C8 + C8]};
    float best[8];
    uint32_t idx[8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const uint32_t w4[4] = {q[m].x, q[m].y, q[m].z, q[m].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 v = unpack16<FMT>(w4[k]);
        if (m == 0) {
          best[2 * k] = v.x, best[2 * k + 1] = v.y, idx[2 * k] = idx[2 * k + 1] = 0;
        } else {
          if (v.x > best[2 * k]) best[2 * k] = v.x, idx[2 * k] = m;
          if (v.y > best[2 * k + 1]) best[2 * k + 1] = v.y, idx[2 * k + 1] = m;
        }
      }
    }
    y[i] = make_uint4(pack16<FMT>(best[0], best[1]), pack16<FMT>(best[2], best[3]), pack16<FMT>(best[4], best[5]),
                      pack16<FMT>(best[6], best[7]));
    arg[i] = make_uint2(idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24),
                        idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24));
  }
}

// pooled gradient dy (fp32 NHWC, pitch = C) -> pre-pool gradient dpre (16-bit NHWC): the value goes to the arg-max member of
// each window where the pooled activation is positive (ReLU), zeros elsewhere; dym (optional) = the masked pooled gradient
template <int FMT>
__global__ void __launch_bounds__(256) unpool_relu_kernel(const float4* __restrict__ dy, const uint4* __restrict__ y,
                                                          const uint2* __restrict__ arg, uint4* __restrict__ dpre,
                                                          float4* __restrict__ dym, size_t total, int H, int W, int C8) {
  const int OW = W / 2, OH = H / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % OW), oy = (int)((i / ((size_t)C8 * OW)) % OH);
    const size_t n = i / ((size_t)C8 * OW * OH);
    const float4 g0 = dy[2 * i], g1 = dy[2 * i + 1];
    float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const uint4 yv = y[i];
    const uint32_t y4[4] = {yv.x, yv.y, yv.z, yv.w};
    const uint2 av = arg[i];
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v = unpack16<FMT>(y4[k]);
      if (!(v.x > 0.0f)) g[2 * k] = 0.0f;
      if (!(v.y > 0.0f)) g[2 * k + 1] = 0.0f;
      a[k] = (av.x >> (8 * k)) & 255u;
      a[4 + k] = (av.y >> (8 * k)) & 255u;
    }
    if (dym) {
      dym[2 * i] = make_float4(g[0], g[1], g[2], g[3]);
      dym[2 * i + 1] = make_float4(g[4], g[5], g[6], g[7]);
    }
    uint4* o = dpre + ((n * H + 2 * oy) * W + 2 * ox) * C8 + c;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = a[k] == (uint32_t)m ? g[k] : 0.0f;
      o[(m >> 1) * (size_t)W * C8 + (m & 1) * C8] =
          make_uint4(pack16<FMT>(v[0], v[1]), pack16<FMT>(v[2], v[3]), pack16<FMT>(v[4], v[5]), pack16<FMT>(v[6], v[7]));
    }
  }
}

// fp32 planar (N, C <= 8, H, W) -> NHWC with 8 channels per pixel, 16-bit format fmt
__global__ void __launch_bounds__(256) image_to_nhwc8_16_kernel(const float* __restrict__ img, uint4* __restrict__ out, int C,
                                                                int HW, size_t total_pixels, int fmt) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= total_pixels) return;
  const size_t n = p / HW, hw = p % HW;
  const float* s = img + n * (size_t)C * HW + hw;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? s[(size_t)c * HW] : 0.0f;
  out[p] = make_uint4(pack16_rt(v[0], v[1], fmt), pack16_rt(v[2], v[3], fmt), pack16_rt(v[4], v[5], fmt), pack16_rt(v[6], v[7], fmt));
}

// GEMM-row weights out[co][tap*Cpad + c] = w[co][c][tap] (zero for c >= Cin), 16-bit format fmt
__global__ void weight_to_im2col16_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cin, int Cpad, int total,
                                          int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % Cpad, tap = (i / Cpad) % 9, co = i / (Cpad * 9);
  out[i] = cvt16_rt(c < Cin ? w[((size_t)co * Cin + c) * 9 + tap] : 0.0f, fmt);
}
// ... and back for the gradient: dw[co][c][tap] = g[co][tap*Cpad + c] (fp32), c < Cin
__global__ void wgrad_from_im2col_kernel(const float* __restrict__ g, float* __restrict__ dw, int Cin, int Cpad, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = i % 9, c = (i / 9) % Cin, co = i / (9 * Cin);
  dw[i] = g[((size_t)co * 9 + tap) * Cpad + c];
}
// data-gradient weights: out[ci][tap*Cout + co] = w[co][ci][8 - tap] (the transposed convolution = correlation with the
// spatially flipped, channel-transposed filter), rows ci < CinPad (zero rows above Cin)
__global__ void weight_to_dgrad16_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int Cin, int Cout, int total,
                                         int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = i % Cout, tap = (i / Cout) % 9, ci = i / (Cout * 9);
  out[i] = cvt16_rt(ci < Cin ? w[((size_t)co * Cin + ci) * 9 + (8 - tap)] : 0.0f, fmt);
}
// Linear weight gradient over the (H, W, C) flattening -> nn.Flatten's (C, H, W) order: dw[o][c*HW + hw] = g[o][hw*C + c]
__global__ void __launch_bounds__(256) fc_grad_hwc_to_chw_kernel(const float* __restrict__ g, float* __restrict__ dw, int C,
                                                                 int HW, size_t total) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t K = (size_t)C * HW;
  const size_t o = i / K, k = i % K;
  const size_t c = k / HW, hw = k % HW;
  dw[i] = g[o * K + hw * C + c];
}

}  // namespace ctrain
}  // namespace bbbp

using namespace bbbp;

static unsigned grid_for(size_t total) {
  const size_t b = ceil_div(total, (size_t)256), cap = (size_t)current_sm_count() * 32;
  return (unsigned)(b < cap ? b : cap);
}
#define BBBP_FMT_OK(who) BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, who ": bad fmt %d", fmt)

extern "C" int bbbp_maxpool2x2_argmax_nhwc16(int fmt, const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C,
                                             bbbp_stream_t stream) {
  BBBP_FMT_OK("maxpool2x2_argmax");
  BBBP_CHECK_ARG(x && y && argmax && N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                 "maxpool2x2_argmax: even H, W and C %% 8 == 0 required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return BBBP_OK;
  if (fmt == BBBP_FMT_F16)
    ctrain::maxpool_argmax_kernel<BBBP_FMT_F16><<<grid_for(total), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), reinterpret_cast<uint2*>(argmax), total, H, W, C / 8);
  else
    ctrain::maxpool_argmax_kernel<BBBP_FMT_BF16><<<grid_for(total), 256, 0, as_stream(stream)>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), reinterpret_cast<uint2*>(argmax), total, H, W, C / 8);
  return launch_status("maxpool2x2_argmax_nhwc16");
}

extern "C" int bbbp_unpool_relu_nhwc16(int fmt, const float* dy, const void* y, const uint8_t* argmax, void* dpre, float* dy_masked,
                                       int N, int H, int W, int C, bbbp_stream_t stream) {
  BBBP_FMT_OK("unpool_relu");
  BBBP_CHECK_ARG(dy && y && argmax && dpre && N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                 "unpool_relu: even H, W and C %% 8 == 0 required");
  BBBP_CHECK_ARG(((uintptr_t)dy % 16) == 0 && ((uintptr_t)dy_masked % 16) == 0, "unpool_relu: gradients must be 16-byte aligned");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return BBBP_OK;
  if (fmt == BBBP_FMT_F16)
    ctrain::unpool_relu_kernel<BBBP_FMT_F16><<<grid_for(total), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(dy), static_cast<const uint4*>(y), reinterpret_cast<const uint2*>(argmax),
        static_cast<uint4*>(dpre), reinterpret_cast<float4*>(dy_masked), total, H, W, C / 8);
  else
    ctrain::unpool_relu_kernel<BBBP_FMT_BF16><<<grid_for(total), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(dy), static_cast<const uint4*>(y), reinterpret_cast<const uint2*>(argmax),
        static_cast<uint4*>(dpre), reinterpret_cast<float4*>(dy_masked), total, H, W, C / 8);
  return launch_status("unpool_relu_nhwc16");
}

extern "C" int bbbp_image_to_nhwc8_16(int fmt, const float* img_nchw, void* out_nhwc8, int N, int C, int H, int W,
                                      bbbp_stream_t stream) {
  BBBP_FMT_OK("image_to_nhwc8_16");
  BBBP_CHECK_ARG(img_nchw && out_nhwc8 && N >= 0 && C >= 1 && C <= 8 && H > 0 && W > 0, "image_to_nhwc8_16: bad argument");
  const size_t total = (size_t)N * H * W;
  if (total == 0) return BBBP_OK;
  ctrain::image_to_nhwc8_16_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(
      img_nchw, static_cast<uint4*>(out_nhwc8), C, H * W, total, fmt);
  return launch_status("image_to_nhwc8_16");
}

extern "C" int bbbp_conv3x3_weight_im2col16(int fmt, const float* w, void* out16, int Cin, int Cpad, int Cout, bbbp_stream_t stream) {
  BBBP_FMT_OK("conv3x3_weight_im2col16");
  BBBP_CHECK_ARG(w && out16 && Cin > 0 && Cpad >= Cin && Cpad % 8 == 0 && Cout > 0, "conv3x3_weight_im2col16: bad argument");
  const int total = Cout * 9 * Cpad;
  ctrain::weight_to_im2col16_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, static_cast<uint16_t*>(out16), Cin, Cpad,
                                                                                         total, fmt);
  return launch_status("conv3x3_weight_im2col16");
}

extern "C" int bbbp_conv3x3_wgrad_from_im2col_f32(const float* g, float* dw, int Cin, int Cpad, int Cout, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(g && dw && Cin > 0 && Cpad >= Cin && Cout > 0, "conv3x3_wgrad_from_im2col: bad argument");
  const int total = Cout * Cin * 9;
  ctrain::wgrad_from_im2col_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(g, dw, Cin, Cpad, total);
  return launch_status("conv3x3_wgrad_from_im2col");
}

extern "C" int bbbp_conv3x3_weight_dgrad16(int fmt, const float* w, void* out16, int Cin, int CinPad, int Cout, bbbp_stream_t stream) {
  BBBP_FMT_OK("conv3x3_weight_dgrad16");
  BBBP_CHECK_ARG(w && out16 && Cin > 0 && CinPad >= Cin && Cout > 0 && Cout % 8 == 0, "conv3x3_weight_dgrad16: bad argument");
  const int total = CinPad * 9 * Cout;
  ctrain::weight_to_dgrad16_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, static_cast<uint16_t*>(out16), Cin, Cout,
                                                                                        total, fmt);
  return launch_status("conv3x3_weight_dgrad16");
}

extern "C" int bbbp_fc_grad_hwc_to_chw_f32(const float* g, float* dw, int rows, int C, int HW, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(g && dw && rows > 0 && C > 0 && HW > 0, "fc_grad_hwc_to_chw: bad argument");
  const size_t total = (size_t)rows * C * HW;
  ctrain::fc_grad_hwc_to_chw_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(g, dw, C, HW, total);
  return launch_status("fc_grad_hwc_to_chw");
}
