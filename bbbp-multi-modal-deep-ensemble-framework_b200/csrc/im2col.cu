// Generic tensor-core route for 3x3 convolutions whose channel counts the implicit-GEMM kernel (conv_umma.cu) is not
// instantiated for -- the 64 / 128 / 256-channel stack of the big variant (20250107_network.py:133-141): an explicit
// bf16 im2col feeds the tcgen05 GEMM (bias + ReLU in its epilogue), then a 2x2 max-pool over the NHWC result
// (ReLU and max commute, so conv -> +bias -> ReLU -> pool is preserved).  Memory-bound helpers: 16-byte accesses,
// one 8-channel chunk per thread.
#include "common.cuh"

namespace bbbp {

// out[(n*H + y)*W + x][tap*C + c] = x[n][y+dy][x+dx][c] (zero outside the image), tap = 3*(dy+1) + (dx+1)
__global__ void __launch_bounds__(256) im2col3x3_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, size_t total,
                                                             int H, int W, int C8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int tap = (int)((i / C8) % 9);
    const size_t pix = i / ((size_t)C8 * 9);
    const int px = (int)(pix % W), py = (int)((pix / W) % H);
    const size_t n = pix / ((size_t)W * H);
    const int sy = py + tap / 3 - 1, sx = px + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = x[((n * H + sy) * W + sx) * C8 + c];
    out[i] = v;
  }
}

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  return make_uint4(max_bf16x2(a.x, b.x), max_bf16x2(a.y, b.y), max_bf16x2(a.z, b.z), max_bf16x2(a.w, b.w));
}

// y[n][oy][ox][c] = max over the 2x2 window of x[n][2oy+i][2ox+j][c], NHWC bf16
__global__ void __launch_bounds__(256) maxpool2x2_nhwc_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                                   size_t total, int H, int W, int C8) {
  const int OW = W / 2, OH = H / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const int ox = (int)((i / C8) % OW), oy = (int)((i / ((size_t)C8 * OW)) % OH);
    const size_t n = i / ((size_t)C8 * OW * OH);
    const uint4* p = x + ((n * H + 2 * oy) * W + 2 * ox) * C8 + c;
    y[i] = max_bf16x8(max_bf16x8(p[0], p[C8]), max_bf16x8(p[(size_t)W * C8], p[(size_t)W * C8 + C8]));
  }
}

// out[co][tap*Cpad + c] = bf16(w[co][c][tap]) (zero for c >= Cin): the GEMM's W operand for the im2col row order
__global__ void conv_weight_im2col_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cin, int Cpad,
                                          int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % Cpad, tap = (i / Cpad) % 9, co = i / (Cpad * 9);
  out[i] = __float2bfloat16(c < Cin ? w[((size_t)co * Cin + c) * 9 + tap] : 0.0f);
}

}  // namespace bbbp

extern "C" int bbbp_im2col3x3_bf16(const void* x_nhwc, void* out, int N, int H, int W, int C, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x_nhwc && out && N >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "im2col3x3: C=%d must be a multiple of 8", C);
  const size_t total = (size_t)N * H * W * 9 * (C / 8);
  if (total == 0) return BBBP_OK;
  const unsigned blocks = (unsigned)(ceil_div(total, (size_t)256) < 148u * 64 ? ceil_div(total, (size_t)256) : 148u * 64);
  im2col3x3_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(x_nhwc), static_cast<uint4*>(out), total,
                                                                H, W, C / 8);
  return launch_status("im2col3x3_bf16");
}

extern "C" int bbbp_maxpool2x2_nhwc_bf16(const void* x_nhwc, void* y_nhwc, int N, int H, int W, int C, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x_nhwc && y_nhwc && N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                 "maxpool2x2_nhwc: even H, W and C %% 8 == 0 required");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return BBBP_OK;
  const unsigned blocks = (unsigned)(ceil_div(total, (size_t)256) < 148u * 64 ? ceil_div(total, (size_t)256) : 148u * 64);
  maxpool2x2_nhwc_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(x_nhwc), static_cast<uint4*>(y_nhwc),
                                                                      total, H, W, C / 8);
  return launch_status("maxpool2x2_nhwc_bf16");
}

extern "C" int bbbp_conv3x3_weight_im2col_bf16(const float* w, void* out_bf16, int Cin, int Cpad, int Cout, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(w && out_bf16 && Cin > 0 && Cpad >= Cin && Cpad % 8 == 0 && Cout > 0, "conv3x3_weight_im2col: bad argument");
  const int total = Cout * 9 * Cpad;
  conv_weight_im2col_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, static_cast<__nv_bfloat16*>(out_bf16), Cin,
                                                                                 Cpad, total);
  return launch_status("conv3x3_weight_im2col");
}
