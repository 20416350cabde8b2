// fp32 CUDA-core 3x3 convolution family (validation / training path).
// Forward: conv(k3,s1,p1) + bias + ReLU + MaxPool2d(2) fused -- the pre-pool activation (2 MB per molecule
// for conv1) never reaches HBM.  Reference call sites: 20250113.py:85-90, 20250107_network.py:133-141.
#include "common.cuh"

namespace bbbp {

constexpr int CI_CHUNK = 8;

// block: 256 threads = 64 pooling windows (16x16 pre-pool tile) x 4 channel groups of CPT channels.
// The input-channel chunks are double buffered: chunk c+1 (halo tile + weight slice) is fetched with cp.async while
// chunk c is multiplied, so the FMA pipe no longer idles through a global-load round trip per chunk (ncu before:
// FMA pipe 48 % active, top stall long_scoreboard at 16 resident warps).
template <int CPT>
struct ConvF32Smem {
  static constexpr int CB = 4 * CPT;
  static constexpr int XS = CI_CHUNK * 18 * 18, WS = CI_CHUNK * 9 * CB;
  static constexpr int STAGE = XS + WS;                     // floats per buffer
  static constexpr size_t BYTES = 2 * (size_t)STAGE * sizeof(float);
};

template <int CPT>
__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y,
                                                          uint8_t* __restrict__ argmax, int Cin, int Cout, int H, int W,
                                                          int pool) {
  using S = ConvF32Smem<CPT>;
  constexpr int CB = S::CB;  // channels per block
  extern __shared__ __align__(16) float conv_smem[];
  const int tid = threadIdx.x;
  const int tiles_x = W / 16;
  const int tx0 = (blockIdx.x % tiles_x) * 16, ty0 = (blockIdx.x / tiles_x) * 16;
  const int co0 = blockIdx.y * CB;
  const int n = blockIdx.z;
  const int wq = tid % 64, cg = tid / 64;
  const int wx = wq % 8, wy = wq / 8;
  float acc[CPT][4];
#pragma unroll
  for (int c = 0; c < CPT; ++c)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[c][p] = 0.0f;

  const float* xn = x + (size_t)n * Cin * H * W;
  auto fetch = [&](int ci0, float* buf) {
    float* xs = buf;
    float* ws = buf + S::XS;
    const int nci = min(CI_CHUNK, Cin - ci0);
    for (int i = tid; i < nci * 324; i += 256) {
      const int ci = i / 324, rc = i % 324, r = rc / 18, c = rc % 18;
      const int gy = ty0 + r - 1, gx = tx0 + c - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async_f32(xs + i, ok ? xn + ((size_t)(ci0 + ci) * H + gy) * W + gx : xn, ok);
    }
    for (int i = tid; i < nci * 9 * CB; i += 256) {
      const int co = i % CB, tap = (i / CB) % 9, ci = i / (CB * 9);
      cp_async_f32(ws + i, w + ((size_t)(co0 + co) * Cin + ci0 + ci) * 9 + tap, true);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  const int chunks = ceil_div(Cin, CI_CHUNK);
  fetch(0, conv_smem);
  for (int ch = 0; ch < chunks; ++ch) {
    float* buf = conv_smem + (ch & 1) * S::STAGE;
    if (ch + 1 < chunks) {
      fetch((ch + 1) * CI_CHUNK, conv_smem + ((ch + 1) & 1) * S::STAGE);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // chunk ch has landed, chunk ch+1 may still fly
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    const float* xs = buf;
    const float* ws = buf + S::XS;
    const int nci = min(CI_CHUNK, Cin - ch * CI_CHUNK);
    for (int ci = 0; ci < nci; ++ci) {
      float p[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[i][j] = xs[ci * 324 + (2 * wy + i) * 18 + 2 * wx + j];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            float wv = ws[(ci * 9 + kh * 3 + kw) * CB + cg * CPT + c];
            acc[c][0] = fmaf(p[kh][kw], wv, acc[c][0]);
            acc[c][1] = fmaf(p[kh][kw + 1], wv, acc[c][1]);
            acc[c][2] = fmaf(p[kh + 1][kw], wv, acc[c][2]);
            acc[c][3] = fmaf(p[kh + 1][kw + 1], wv, acc[c][3]);
          }
    }
    __syncthreads();        // everyone is done with buf before the fetch of chunk ch+2 overwrites it
  }

  const int gy = ty0 + 2 * wy, gx = tx0 + 2 * wx;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int co = co0 + cg * CPT + c;
    const float b = bias ? bias[co] : 0.0f;
    if (pool) {
      float best = fmaxf(acc[c][0] + b, 0.0f);
      int arg = 0;
#pragma unroll
      for (int p = 1; p < 4; ++p) {
        float v = fmaxf(acc[c][p] + b, 0.0f);
        if (v > best) { best = v; arg = p; }
      }
      size_t o = (((size_t)n * Cout + co) * (H / 2) + gy / 2) * (W / 2) + gx / 2;
      y[o] = best;
      if (argmax) argmax[o] = (uint8_t)arg;
    } else {
      float* yo = y + (((size_t)n * Cout + co) * H + gy) * W + gx;
      yo[0] = acc[c][0] + b;
      yo[1] = acc[c][1] + b;
      yo[W] = acc[c][2] + b;
      yo[W + 1] = acc[c][3] + b;
    }
  }
}

__global__ void relu_pool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                     const uint8_t* __restrict__ argmax, float* __restrict__ dpre, size_t total, int H,
                                     int W) {
  // one thread per pooled element; writes its 2x2 window
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int PW = W / 2, PH = H / 2;
  int pw = i % PW, ph = (i / PW) % PH;
  size_t plane = i / ((size_t)PW * PH);
  float g = y[i] > 0.0f ? dy[i] : 0.0f;
  int a = argmax[i];
  float* o = dpre + (plane * H + 2 * ph) * W + 2 * pw;
  o[0] = a == 0 ? g : 0.0f;
  o[1] = a == 1 ? g : 0.0f;
  o[W] = a == 2 ? g : 0.0f;
  o[W + 1] = a == 3 ? g : 0.0f;
}

// Partial weight gradient, register tiled.  A CTA owns one 8-row strip of the image (band), a tile of COT*16 output
// channels and a tile of CI_T input channels, and walks images z, z+gridDim.z, ...  Thread (cog, lane): lane -> (ci,
// row slice), cog -> COT consecutive output channels; it keeps COT x 9 accumulators and slides a 3x3 input window along
// each row, so a pixel costs 3 + COT shared-memory loads for 9*COT FMAs (the previous kernel: 10 loads for 9 FMAs, and
// 13 of its 16 ci lanes idle on the 3-channel first layer, which the SLICES split over rows now fills).
// Partials [z][band][slice][Cout][Cin][9] are summed by sum_over_images_kernel in a fixed order (deterministic).
constexpr int WG_ROWS = 8, WG_COLS = 16;
template <int CI_T, int COT>
struct WgradSmem {
  static constexpr int CO_TILE = 16 * COT, XS_PITCH = (WG_ROWS + 2) * (WG_COLS + 2) + 1;   // odd: conflict-free across ci
  static constexpr int DS = CO_TILE * WG_ROWS * WG_COLS, XS = CI_T * XS_PITCH;
  static constexpr int STAGE = DS + XS;                                                     // floats per buffer
  static constexpr size_t BYTES = 2 * (size_t)STAGE * sizeof(float);
};

// (image, x tile) pairs are double buffered: tile t+1 is fetched with cp.async while tile t is accumulated.
template <int CI_T, int SLICES, int COT>
__global__ void __launch_bounds__(256) conv3x3_wgrad_tiled_kernel(const float* __restrict__ dpre, const float* __restrict__ x,
                                                                  float* __restrict__ part, int N, int Cin, int Cout, int H,
                                                                  int W) {
  static_assert(CI_T * SLICES == 16 && WG_ROWS % SLICES == 0, "lane split");
  using S = WgradSmem<CI_T, COT>;
  constexpr int CO_TILE = S::CO_TILE, XS_PITCH = S::XS_PITCH;
  constexpr int ROWS_PER_SLICE = WG_ROWS / SLICES;
  extern __shared__ __align__(16) float wg_smem[];
  const int tid = threadIdx.x;
  const int bands = H / WG_ROWS;
  const int band = blockIdx.x % bands, ci0 = (blockIdx.x / bands) * CI_T, co0 = blockIdx.y * CO_TILE;
  const int lane16 = tid % 16, cog = tid / 16;
  const int ci_l = lane16 % CI_T, slice = lane16 / CI_T;
  const int nci = min(CI_T, Cin - ci0);
  const int ty0 = band * WG_ROWS;
  float acc[COT][9];
#pragma unroll
  for (int j = 0; j < COT; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.0f;

  const int tiles_x = W / WG_COLS;
  const int my_images = (N - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;
  const int total = my_images * tiles_x;
  auto fetch = [&](int t, float* buf) {
    const int n = blockIdx.z + (t / tiles_x) * gridDim.z, tx0 = (t % tiles_x) * WG_COLS;
    const float* xn = x + (size_t)n * Cin * H * W;
    const float* dn = dpre + (size_t)n * Cout * H * W;
    float* ds = buf;
    float* xs = buf + S::DS;
    for (int i = tid; i < CO_TILE * WG_ROWS * WG_COLS; i += 256) {
      const int co = i / (WG_ROWS * WG_COLS), p = i % (WG_ROWS * WG_COLS);
      cp_async_f32(ds + i, dn + ((size_t)(co0 + co) * H + ty0 + p / WG_COLS) * W + tx0 + p % WG_COLS, true);
    }
    for (int i = tid; i < CI_T * (WG_ROWS + 2) * (WG_COLS + 2); i += 256) {
      const int ci = i / ((WG_ROWS + 2) * (WG_COLS + 2)), rc = i % ((WG_ROWS + 2) * (WG_COLS + 2));
      const int gy = ty0 + rc / (WG_COLS + 2) - 1, gx = tx0 + rc % (WG_COLS + 2) - 1;
      const bool ok = ci < nci && gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async_f32(xs + ci * XS_PITCH + rc, ok ? xn + ((size_t)(ci0 + ci) * H + gy) * W + gx : xn, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  if (total > 0) fetch(0, wg_smem);
  for (int t = 0; t < total; ++t) {
    const float* buf = wg_smem + (t & 1) * S::STAGE;
    if (t + 1 < total) {
      fetch(t + 1, wg_smem + ((t + 1) & 1) * S::STAGE);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    const float* ds = buf;
    const float* xc = buf + S::DS + ci_l * XS_PITCH;
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_SLICE; ++rr) {
      const int py = slice * ROWS_PER_SLICE + rr;
      const float* r0 = xc + py * (WG_COLS + 2);           // input rows py-1, py, py+1 in tile coordinates (+1 halo)
      const float* r1 = r0 + (WG_COLS + 2);
      const float* r2 = r1 + (WG_COLS + 2);
      float w00 = r0[0], w01 = r0[1], w10 = r1[0], w11 = r1[1], w20 = r2[0], w21 = r2[1];
#pragma unroll
      for (int px = 0; px < WG_COLS; ++px) {
        const float w02 = r0[px + 2], w12 = r1[px + 2], w22 = r2[px + 2];
#pragma unroll
        for (int j = 0; j < COT; ++j) {
          const float d = ds[(cog * COT + j) * (WG_ROWS * WG_COLS) + py * WG_COLS + px];
          acc[j][0] = fmaf(d, w00, acc[j][0]);
          acc[j][1] = fmaf(d, w01, acc[j][1]);
          acc[j][2] = fmaf(d, w02, acc[j][2]);
          acc[j][3] = fmaf(d, w10, acc[j][3]);
          acc[j][4] = fmaf(d, w11, acc[j][4]);
          acc[j][5] = fmaf(d, w12, acc[j][5]);
          acc[j][6] = fmaf(d, w20, acc[j][6]);
          acc[j][7] = fmaf(d, w21, acc[j][7]);
          acc[j][8] = fmaf(d, w22, acc[j][8]);
        }
        w00 = w01; w01 = w02; w10 = w11; w11 = w12; w20 = w21; w21 = w22;
      }
    }
    __syncthreads();        // done with buf before the fetch of tile t+2 overwrites it
  }
  if (ci_l < nci) {
    const size_t per_w = (size_t)Cout * Cin * 9;
    float* base = part + (((size_t)blockIdx.z * bands + band) * SLICES + slice) * per_w;
#pragma unroll
    for (int j = 0; j < COT; ++j) {
      float* o = base + ((size_t)(co0 + cog * COT + j) * Cin + ci0 + ci_l) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) o[t] = acc[j][t];
    }
  }
}

struct WgradPlan { int ci_t, slices, cot, zn, bands; };
static WgradPlan wgrad_plan(int N, int Cin, int Cout, int H) {
  WgradPlan p;
  p.ci_t = Cin <= 4 ? 4 : Cin <= 8 ? 8 : 16;
  p.slices = 16 / p.ci_t;
  p.cot = Cout % 64 == 0 ? 4 : Cout % 32 == 0 ? 2 : 1;
  p.zn = N < 32 ? N : 32;
  p.bands = H / WG_ROWS;
  return p;
}

// bias-gradient partial: block = (co, n) -> part_b[n][co] = sum over the plane, fixed reduction tree
__global__ void __launch_bounds__(256) conv_bgrad_partial_kernel(const float* __restrict__ dpre, float* __restrict__ part_b,
                                                                 int Cout, int HW) {
  __shared__ float red[8];
  const int co = blockIdx.x, n = blockIdx.y;
  const float* d = dpre + ((size_t)n * Cout + co) * HW;
  float s = 0.0f;
  for (int i = threadIdx.x; i < HW; i += 256) s += d[i];
  s = warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    part_b[(size_t)n * Cout + co] = t;
  }
}

// out[i] = sum_n part[n][i]: block = 32 outputs x 8 partial lanes (lane l sums n = l, l+8, ...), then a fixed-order
// combine of the 8 lanes -- deterministic, coalesced, and 8x the parallelism of one thread per output
__global__ void __launch_bounds__(256) sum_over_images_kernel(const float* __restrict__ part, float* __restrict__ out, int N,
                                                              size_t per_image) {
  __shared__ float red[8][33];
  const size_t i = blockIdx.x * (size_t)32 + threadIdx.x;
  float s = 0.0f;
  if (i < per_image)
    for (int n = threadIdx.y; n < N; n += 8) s += part[(size_t)n * per_image + i];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < per_image) {
    float t = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += red[l][threadIdx.x];
    out[i] = t;
  }
}

__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cin, int Cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin * Cout * 9) return;
  int tap = i % 9, co = (i / 9) % Cout, ci = i / (9 * Cout);  // wt[ci][co][tap]
  wt[i] = w[((size_t)co * Cin + ci) * 9 + (8 - tap)];
}

}  // namespace bbbp

extern "C" int bbbp_conv3x3_f32(const float* x, const float* w, const float* b, float* y, uint8_t* argmax, int N,
                                int Cin, int Cout, int H, int W, int pool, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x && w && y, "conv3x3_f32: null operand");
  BBBP_CHECK_ARG(N >= 0 && Cin > 0 && Cout > 0 && Cout % 32 == 0, "conv3x3_f32: Cout=%d must be a multiple of 32", Cout);
  BBBP_CHECK_ARG(H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0, "conv3x3_f32: H=%d W=%d must be multiples of 16", H, W);
  if (N == 0) return BBBP_OK;
  BBBP_CHECK_ARG(N <= 65535, "conv3x3_f32: N=%d exceeds 65535 images per launch", N);
  cudaStream_t s = as_stream(stream);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(conv3x3_f32_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvF32Smem<16>::BYTES);
    cudaFuncSetAttribute(conv3x3_f32_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvF32Smem<8>::BYTES);
  }
  if (Cout % 64 == 0) {
    dim3 grid((H / 16) * (W / 16), Cout / 64, N);
    conv3x3_f32_kernel<16><<<grid, 256, ConvF32Smem<16>::BYTES, s>>>(x, w, b, y, argmax, Cin, Cout, H, W, pool);
  } else {
    dim3 grid((H / 16) * (W / 16), Cout / 32, N);
    conv3x3_f32_kernel<8><<<grid, 256, ConvF32Smem<8>::BYTES, s>>>(x, w, b, y, argmax, Cin, Cout, H, W, pool);
  }
  return launch_status("conv3x3_f32");
}

extern "C" int bbbp_relu_pool_bwd_f32(const float* dy, const float* y, const uint8_t* argmax, float* dpre, int N, int C,
                                      int H, int W, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dy && y && argmax && dpre, "relu_pool_bwd: null operand");
  BBBP_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "relu_pool_bwd: odd plane");
  size_t total = (size_t)N * C * (H / 2) * (W / 2);
  if (total == 0) return BBBP_OK;
  relu_pool_bwd_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(dy, y, argmax, dpre, total,
                                                                                               H, W);
  return launch_status("relu_pool_bwd");
}

extern "C" size_t bbbp_conv3x3_wgrad_workspace(int N, int Cin, int Cout, int H, int W) {
  using namespace bbbp;
  (void)W;
  if (N <= 0 || Cin <= 0 || Cout <= 0 || H <= 0) return 0;
  const WgradPlan p = wgrad_plan(N, Cin, Cout, H);
  return ((size_t)p.zn * p.bands * p.slices * Cout * Cin * 9 + (size_t)N * Cout) * sizeof(float);
}

extern "C" int bbbp_conv3x3_wgrad_f32(const float* dpre, const float* x, float* dw, float* db, int N, int Cin, int Cout,
                                      int H, int W, float* workspace, size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dpre && x && dw, "conv3x3_wgrad: null operand");
  BBBP_CHECK_ARG(Cout % 16 == 0 && H % 16 == 0 && W % 16 == 0, "conv3x3_wgrad: Cout/H/W must be multiples of 16");
  BBBP_CHECK_ARG(N > 0 && N <= 65535, "conv3x3_wgrad: N=%d out of range", N);
  const WgradPlan p = wgrad_plan(N, Cin, Cout, H);
  const size_t per_w = (size_t)Cout * Cin * 9, per_b = Cout;
  const size_t n_part = (size_t)p.zn * p.bands * p.slices;
  const size_t need = bbbp_conv3x3_wgrad_workspace(N, Cin, Cout, H, W);
  if (!workspace || workspace_bytes < need) {
    set_error("conv3x3_wgrad: needs %zu workspace bytes, got %zu", need, workspace_bytes);
    return BBBP_EWORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  float* part_w = workspace;
  float* part_b = workspace + n_part * per_w;
  const dim3 grid(ceil_div(Cin, p.ci_t) * p.bands, Cout / (16 * p.cot), p.zn);
#define BBBP_WGRAD(CI, SL, CT)                                                                                              \
  do {                                                                                                                     \
    cudaFuncSetAttribute(conv3x3_wgrad_tiled_kernel<CI, SL, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                         (int)WgradSmem<CI, CT>::BYTES);                                                                   \
    conv3x3_wgrad_tiled_kernel<CI, SL, CT><<<grid, 256, WgradSmem<CI, CT>::BYTES, s>>>(dpre, x, part_w, N, Cin, Cout, H, W); \
  } while (0)
  if (p.cot == 4) {
    if (p.ci_t == 4) BBBP_WGRAD(4, 4, 4); else if (p.ci_t == 8) BBBP_WGRAD(8, 2, 4); else BBBP_WGRAD(16, 1, 4);
  } else if (p.cot == 2) {
    if (p.ci_t == 4) BBBP_WGRAD(4, 4, 2); else if (p.ci_t == 8) BBBP_WGRAD(8, 2, 2); else BBBP_WGRAD(16, 1, 2);
  } else {
    if (p.ci_t == 4) BBBP_WGRAD(4, 4, 1); else if (p.ci_t == 8) BBBP_WGRAD(8, 2, 1); else BBBP_WGRAD(16, 1, 1);
  }
#undef BBBP_WGRAD
  int st = launch_status("conv3x3_wgrad partial");
  if (st != BBBP_OK) return st;
  sum_over_images_kernel<<<(unsigned)ceil_div(per_w, (size_t)32), dim3(32, 8), 0, s>>>(part_w, dw, (int)n_part, per_w);
  st = launch_status("conv3x3_wgrad reduce");
  if (st != BBBP_OK || !db) return st;
  conv_bgrad_partial_kernel<<<dim3(Cout, N), 256, 0, s>>>(dpre, part_b, Cout, H * W);
  sum_over_images_kernel<<<(unsigned)ceil_div(per_b, (size_t)32), dim3(32, 8), 0, s>>>(part_b, db, N, per_b);
  note_launches(1);
  return launch_status("conv3x3 bias grad");
}

extern "C" int bbbp_conv3x3_flip_weights_f32(const float* w, float* w_t, int Cin, int Cout, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(w && w_t && Cin > 0 && Cout > 0, "flip_weights: bad argument");
  int total = Cin * Cout * 9;
  flip_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, w_t, Cin, Cout);
  return launch_status("flip_weights");
}
