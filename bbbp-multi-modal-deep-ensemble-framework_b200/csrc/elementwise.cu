// Memory-bound kernels of the hot path: fusion-block mixing, activation gradients, column reductions,
// dropout, losses, fused multi-tensor AdamW, and the packed-input contracts.  All single pass, coalesced,
// warp-shuffle reductions, deterministic (no floating-point atomics anywhere).
#include <math.h>
#include <algorithm>
#include <string.h>
#include "common.cuh"
#include "half16.cuh"

namespace bbbp {

// ---- fusion: softmax over n head scores, out = sum_h w_h * c (20250113.py:60-65) -------------------------------
__global__ void __launch_bounds__(128) fusion_softmax_mix_fwd_kernel(const float* __restrict__ scores,
                                                                     const float* __restrict__ c, float* __restrict__ out,
                                                                     float* __restrict__ w_out, int rows, int n, int dim) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float sc = lane < n ? scores[(size_t)row * n + lane] : -INFINITY;
  const float mx = warp_max(sc);
  const float e = lane < n ? expf(sc - mx) : 0.0f;
  const float w = e / warp_sum(e);
  if (w_out && lane < n) w_out[(size_t)row * n + lane] = w;
  for (int i0 = 0; i0 < dim; i0 += 32) {  // uniform trip count: the shuffles need the whole warp
    const int i = i0 + lane;
    const float cv = i < dim ? c[(size_t)row * dim + i] : 0.0f;
    float acc = 0.0f;
    for (int h = 0; h < n; ++h) acc += __shfl_sync(0xffffffffu, w, h) * cv;
    if (i < dim) out[(size_t)row * dim + i] = acc;
  }
}

__global__ void __launch_bounds__(128) fusion_softmax_mix_bwd_kernel(const float* __restrict__ w, const float* __restrict__ c,
                                                                     const float* __restrict__ dout, float* __restrict__ dc,
                                                                     float* __restrict__ dscores, int rows, int n, int dim) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float wl = lane < n ? w[(size_t)row * n + lane] : 0.0f;
  const float wsum = warp_sum(wl);
  float t = 0.0f;
  for (int i0 = 0; i0 < dim; i0 += 32) {
    const int i = i0 + lane;
    const float g = i < dim ? dout[(size_t)row * dim + i] : 0.0f;
    if (i < dim) t = fmaf(g, c[(size_t)row * dim + i], t);
    float acc = 0.0f;
    for (int h = 0; h < n; ++h) acc += __shfl_sync(0xffffffffu, wl, h) * g;
    if (i < dim) dc[(size_t)row * dim + i] = acc;
  }
  t = warp_sum(t);  // d out / d w_h is the same inner product for every head
  if (lane < n) dscores[(size_t)row * n + lane] = wl * (t - wsum * t);
}

// ---- plain row softmax over n <= 32 scores (big variant's 2-way modality weights, 20250107_network.py:83) ----------
__global__ void __launch_bounds__(128) softmax_rows_fwd_kernel(const float* __restrict__ scores, float* __restrict__ w,
                                                               int rows, int n) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float sc = lane < n ? scores[(size_t)row * n + lane] : -INFINITY;
  const float mx = warp_max(sc);
  const float e = lane < n ? expf(sc - mx) : 0.0f;
  const float s = warp_sum(e);
  if (lane < n) w[(size_t)row * n + lane] = e / s;
}
__global__ void __launch_bounds__(128) softmax_rows_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                               float* __restrict__ dscores, int rows, int n) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float wl = lane < n ? w[(size_t)row * n + lane] : 0.0f;
  const float gl = lane < n ? dw[(size_t)row * n + lane] : 0.0f;
  const float dot = warp_sum(wl * gl);
  if (lane < n) dscores[(size_t)row * n + lane] = wl * (gl - dot);
}

// ---- scaled column mean (big variant, 20250107_network.py:85-96) -------------------------------------------------
__device__ __forceinline__ float strip_reduce(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) scaled_colmean_fwd_kernel(const float* __restrict__ x, int ldx,
                                                                 const float* __restrict__ scale, int ld_scale,
                                                                 float* __restrict__ out, int ld_out,
                                                                 float* __restrict__ colmean_out, int rows, int cols) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < cols;
  float s = 0.0f;
  if (ok)
    for (int r = threadIdx.y; r < rows; r += 8) s += x[(size_t)r * ldx + c];
  const float mean = strip_reduce(s, red) / rows;
  if (!ok) return;
  if (colmean_out && threadIdx.y == 0) colmean_out[c] = mean;
  for (int r = threadIdx.y; r < rows; r += 8) out[(size_t)r * ld_out + c] = scale[(size_t)r * ld_scale] * mean;
}

__global__ void __launch_bounds__(256) scaled_colmean_bwd_dx_kernel(const float* __restrict__ dout, int ld_dout,
                                                                    const float* __restrict__ scale, int ld_scale,
                                                                    float* __restrict__ dx, int ldx, int rows, int cols) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < cols;
  float s = 0.0f;
  if (ok)
    for (int r = threadIdx.y; r < rows; r += 8) s = fmaf(scale[(size_t)r * ld_scale], dout[(size_t)r * ld_dout + c], s);
  const float t = strip_reduce(s, red) / rows;
  if (!ok) return;
  for (int r = threadIdx.y; r < rows; r += 8) dx[(size_t)r * ldx + c] = t;
}

__global__ void __launch_bounds__(128) rowdot_vec_kernel(const float* __restrict__ a, int lda, const float* __restrict__ v,
                                                         float* __restrict__ out, int rows, int cols) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  float t = 0.0f;
  for (int i = lane; i < cols; i += 32) t = fmaf(a[(size_t)row * lda + i], v[i], t);
  t = warp_sum(t);
  if (lane == 0) out[row] = t;
}

// ---- elementwise ------------------------------------------------------------------------------------------------
__global__ void act_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y,
                               float* __restrict__ dx, int ld_dx, int rows, int cols, int act) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  int r = i / cols, c = i % cols;
  float g = dy[(size_t)r * ld_dy + c], o = y[(size_t)r * ld_y + c];
  float d = act == BBBP_ACT_RELU ? (o > 0.0f ? g : 0.0f) : act == BBBP_ACT_TANH ? g * (1.0f - o * o) : g;
  dx[(size_t)r * ld_dx + c] = d;
}

// contiguous operands: 128-bit loads / stores, two vectors per thread in flight (the scalar kernel sat at 0.53 of HBM)
__global__ void __launch_bounds__(256) act_bwd_flat4_kernel(const float4* __restrict__ dy, const float4* __restrict__ y,
                                                            float4* __restrict__ dx, size_t n4, int act) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const size_t j = i + stride;
    const bool two = j < n4;
    const float4 g0 = dy[i], o0 = y[i];
    const float4 g1 = two ? dy[j] : make_float4(0.f, 0.f, 0.f, 0.f), o1 = two ? y[j] : g0;
    auto f = [act](float g, float o) {
      return act == BBBP_ACT_RELU ? (o > 0.0f ? g : 0.0f) : act == BBBP_ACT_TANH ? g * (1.0f - o * o) : g;
    };
    dx[i] = make_float4(f(g0.x, o0.x), f(g0.y, o0.y), f(g0.z, o0.z), f(g0.w, o0.w));
    if (two) dx[j] = make_float4(f(g1.x, o1.x), f(g1.y, o1.y), f(g1.z, o1.z), f(g1.w, o1.w));
  }
}

__global__ void scale_by_device_scalar_kernel(const float* __restrict__ x, const float* __restrict__ scalar,
                                              float* __restrict__ y, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * scalar[0];
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int rows,
                                                     int cols) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.0f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) s += x[(size_t)r * ldx + c];
  s = strip_reduce(s, red);
  if (c < cols && threadIdx.y == 0) out[c] = s;
}

// many rows: blockIdx.y owns a contiguous chunk of rows and writes one partial row; a second colsum_kernel pass over the
// [chunks][cols] partials finishes in a fixed order (deterministic, and the first pass fills the GPU)
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int ldx, float* __restrict__ part,
                                                             int rows, int cols, int rows_per_chunk) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s = 0.0f;
  if (c < cols) {
    // four independent row streams per thread: four loads in flight instead of one (0.61 -> of the HBM roofline before)
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      s0 += x[(size_t)r * ldx + c];
      s1 += x[(size_t)(r + 8) * ldx + c];
      s2 += x[(size_t)(r + 16) * ldx + c];
      s3 += x[(size_t)(r + 24) * ldx + c];
    }
    for (; r < r1; r += 8) s0 += x[(size_t)r * ldx + c];
    s = (s0 + s1) + (s2 + s3);
  }
  s = strip_reduce(s, red);
  if (c < cols && threadIdx.y == 0) part[(size_t)blockIdx.y * cols + c] = s;
}

__global__ void copy2d_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst, int rows,
                              int cols) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  int r = i / cols, c = i % cols;
  dst[(size_t)r * ld_dst + c] = src[(size_t)r * ld_src + c];
}

// fp32 rows -> 16-bit rows (bf16 or fp16), optional lo part (hi + lo carries ~2x the mantissa), zero fill of the pad columns.
// One thread = 8 output columns = one 16-byte store per output (the scalar version sat at 0.28 of the HBM roofline).
template <int FMT>
__global__ void __launch_bounds__(256) cast16_kernel(const float* __restrict__ src, int ld_src, const float* __restrict__ col_sub,
                                                     uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int ld_dst, int rows,
                                                     int cols, int cols_pad) {
  const int groups = cols_pad / 8;            // cols_pad is a multiple of 8 on this path
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * groups) return;
  const int r = (int)(i / groups), c0 = (int)(i % groups) * 8;
  const float* s = src + (size_t)r * ld_src + c0;
  float v[8];
  if (c0 + 8 <= cols && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
    const float4 a = reinterpret_cast<const float4*>(s)[0], b = reinterpret_cast<const float4*>(s)[1];
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = c0 + k < cols ? s[k] : 0.0f;
  }
  if (col_sub) {       // fused centring (PCA.transform: X - mean_), in fp32 before the split
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = c0 + k < cols ? v[k] - col_sub[c0 + k] : 0.0f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (lo) split16<FMT>(v[2 * k], v[2 * k + 1], h[k], l[k]);
    else h[k] = pack16<FMT>(v[2 * k], v[2 * k + 1]);
  }
  const size_t at = (size_t)r * ld_dst + c0;
  *reinterpret_cast<uint4*>(hi + at) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + at) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                 int rows, int cols, int cols_pad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols_pad) return;
  int r = i / cols_pad, c = i % cols_pad;
  dst[(size_t)r * ld_dst + c] = __float2bfloat16(c < cols ? src[(size_t)r * ld_src + c] : 0.0f);
}

// ---- row gather: dst[r, :] = src[idx[r], :] (device-resident batch feeder, replaces MixedDataset.__getitem__ + collate) ----
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                          float* __restrict__ dst, int rows, size_t cols) {
  const int r = blockIdx.y;
  const float* s = src + (size_t)idx[r] * cols;
  float* d = dst + (size_t)r * cols;
  const size_t n4 = cols / 4;
  const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
  if (vec) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
      reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cols; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cols; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
  }
}

// ---- Philox-4x32-10 dropout -------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, float p, float inv_keep,
                               uint64_t seed, uint64_t offset, const uint64_t* __restrict__ seed_dev) {
  size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // one Philox block = 4 elements
  if (q * 4 >= n) return;
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  uint64_t ctr = q + offset;
  uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t bits[4] = {r.x, r.y, r.z, r.w};
  if (q * 4 + 4 <= n && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {      // one 128-bit load and store per Philox block
    const float4 v = reinterpret_cast<const float4*>(x)[q];
    const float in[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = bits[k] * 2.3283064365386963e-10f >= p ? in[k] * inv_keep : 0.0f;
    reinterpret_cast<float4*>(y)[q] = make_float4(o[0], o[1], o[2], o[3]);
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    size_t i = q * 4 + k;
    if (i < n) {
      float u = bits[k] * 2.3283064365386963e-10f;  // [0,1)
      y[i] = u >= p ? x[i] * inv_keep : 0.0f;
    }
  }
}

// ---- losses (single block: n is a batch size) -------------------------------------------------------------------
template <int KIND>  // 0 = MSE, 1 = BCE with logits
__global__ void __launch_bounds__(1024) loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                    float* __restrict__ loss, float* __restrict__ dpred, int n,
                                                    float grad_scale) {
  __shared__ float red[32];
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float z = pred[i], t = target[i];
    if (KIND == 0) {
      const float d = z - t;
      s = fmaf(d, d, s);
      if (dpred) dpred[i] = 2.0f * d / n * grad_scale;
    } else {
      // max(z,0) - z t + log1p(exp(-|z|)): torch's stable binary_cross_entropy_with_logits
      s += fmaxf(z, 0.0f) - z * t + log1pf(expf(-fabsf(z)));
      if (dpred) dpred[i] = (1.0f / (1.0f + expf(-z)) - t) / n * grad_scale;
    }
  }
  s = warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(red[threadIdx.x]);
    if (threadIdx.x == 0) loss[0] = t / n;
  }
}

// ---- fused multi-tensor AdamW (torch.optim.AdamW single-tensor semantics, 20250113.py:172,191) -------------------
constexpr int ADAMW_CHUNK = 16384;   // bbbp_adamw_chunk(): ~820 CTAs for the 13.5 M-parameter net (5.5 per SM)
struct AdamwHyper { float v[8]; };  // bbbp_adamw_hyper() layout

// ONE kernel behind both entry points, so an eager step and a graph replay round identically: the eight per-step
// scalars come either by value (bbbp_adamw_f32) or from device memory when the kernel runs (bbbp_adamw_dev_f32).
__global__ void __launch_bounds__(256) adamw_kernel(void* const* __restrict__ ptrs, const int64_t* __restrict__ sizes,
                                                    const int32_t* __restrict__ chunk_tensor,
                                                    const int64_t* __restrict__ chunk_offset, int ntensors,
                                                    AdamwHyper by_value, const float* __restrict__ hyper_dev) {
  const float one_minus_beta1 = hyper_dev ? hyper_dev[0] : by_value.v[0], beta2 = hyper_dev ? hyper_dev[1] : by_value.v[1],
              one_minus_beta2 = hyper_dev ? hyper_dev[2] : by_value.v[2], eps = hyper_dev ? hyper_dev[3] : by_value.v[3],
              decay_mul = hyper_dev ? hyper_dev[4] : by_value.v[4], step_size = hyper_dev ? hyper_dev[5] : by_value.v[5],
              bc2_sqrt = hyper_dev ? hyper_dev[6] : by_value.v[6], grad_scale = hyper_dev ? hyper_dev[7] : by_value.v[7];
  const int t = chunk_tensor[blockIdx.x];
  const int64_t off = chunk_offset[blockIdx.x];
  float* __restrict__ p = static_cast<float*>(ptrs[t]) + off;
  const float* __restrict__ g = static_cast<const float*>(ptrs[ntensors + t]) + off;
  float* __restrict__ m = static_cast<float*>(ptrs[2 * ntensors + t]) + off;
  float* __restrict__ v = static_cast<float*>(ptrs[3 * ntensors + t]) + off;
  const int64_t left = sizes[t] - off;
  const int n = left < ADAMW_CHUNK ? (int)left : ADAMW_CHUNK;
  auto update = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    pp *= decay_mul;
    mm = mm + (gg - mm) * one_minus_beta1;            // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + one_minus_beta2 * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = n / 4;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      update(pp.x, gg.x, mm.x, vv.x);
      update(pp.y, gg.y, mm.y, vv.y);
      update(pp.z, gg.z, mm.z, vv.z);
      update(pp.w, gg.w, mm.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += 256) update(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < n; i += 256) update(p[i], g[i], m[i], v[i]);
  }
}

// ---- input contracts --------------------------------------------------------------------------------------------
// one warp per molecule: popcount -> mean/std in double (as sklearn StandardScaler on the 0/1 column) -> two values
__global__ void __launch_bounds__(128) unpack_zscore_kernel(const uint8_t* __restrict__ packed, int bytes_per_row,
                                                            float* __restrict__ out, int ld_out, int rows, int n_bits) {
  const int row = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= rows) return;
  const uint8_t* pr = packed + (size_t)row * bytes_per_row;
  int pop = 0;
  for (int b = lane; b < bytes_per_row; b += 32) {
    uint32_t byte = pr[b];
    int valid = n_bits - b * 8;
    if (valid < 8) byte &= (1u << (valid > 0 ? valid : 0)) - 1u;
    pop += __popc(byte);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pop += __shfl_xor_sync(0xffffffffu, pop, o);
  const double mean = (double)pop / n_bits;
  // population variance of a 0/1 vector, written as the mean of squared deviations
  const double var = ((double)pop * (1.0 - mean) * (1.0 - mean) + (double)(n_bits - pop) * mean * mean) / n_bits;
  double sd = sqrt(var);
  if (sd == 0.0) sd = 1.0;
  const float v0 = (float)((0.0 - mean) / sd), v1 = (float)((1.0 - mean) / sd);
  for (int i = lane; i < n_bits; i += 32) out[(size_t)row * ld_out + i] = ((pr[i >> 3] >> (i & 7)) & 1) ? v1 : v0;
}

// Contiguous output (ld_out == n_bits, 16-byte aligned base): a block owns 128 consecutive rows, whose 128 * n_bits floats
// start on a 16-byte boundary whatever n_bits is, and writes them as 128-bit stores (the warp-per-row kernel above writes
// rows of 167 floats with scalar stores: 0.27 of the HBM roofline).  The packed rows are staged in shared memory with one
// coalesced pass, the statistics take one thread per row (same float64 formula, same two values per row), and the store
// loop carries its (row, bit) position incrementally: no division, two shared-memory bytes and one 128-bit store per four
// outputs.
constexpr int UZ_ROWS = 128;
__global__ void __launch_bounds__(256) unpack_zscore_vec_kernel(const uint8_t* __restrict__ packed, int bytes_per_row,
                                                                float* __restrict__ out, int rows, int n_bits) {
  extern __shared__ __align__(16) uint8_t sm_bits[];   // UZ_ROWS rows of packed bytes (+ 4 bytes of slack for the 2-byte reads)
  __shared__ float2 sval[UZ_ROWS];
  const int row0 = blockIdx.x * UZ_ROWS;
  const int nrows = min(UZ_ROWS, rows - row0);
  const int nbytes = nrows * bytes_per_row;
  const uint8_t* src = packed + (size_t)row0 * bytes_per_row;
  if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    for (int i = threadIdx.x; i < nbytes / 4; i += 256) reinterpret_cast<uint32_t*>(sm_bits)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    for (int i = (nbytes / 4) * 4 + threadIdx.x; i < nbytes; i += 256) sm_bits[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += 256) sm_bits[i] = src[i];
  }
  if (threadIdx.x < 4) sm_bits[nbytes + threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < nrows) {
    const uint8_t* pr = sm_bits + threadIdx.x * bytes_per_row;
    int pop = 0;
    for (int b = 0; b < bytes_per_row; ++b) {
      uint32_t byte = pr[b];
      const int valid = n_bits - b * 8;
      if (valid < 8) byte &= (1u << (valid > 0 ? valid : 0)) - 1u;
      pop += __popc(byte);
    }
    const double mean = (double)pop / n_bits;
    const double var = ((double)pop * (1.0 - mean) * (1.0 - mean) + (double)(n_bits - pop) * mean * mean) / n_bits;
    double sd = sqrt(var);
    if (sd == 0.0) sd = 1.0;
    sval[threadIdx.x] = make_float2((float)((0.0 - mean) / sd), (float)((1.0 - mean) / sd));
  }
  __syncthreads();
  const int total = nrows * n_bits, total4 = total / 4;                   // total % 4 != 0 only for the last, ragged block
  float* obase = out + (size_t)row0 * n_bits;
  // element 4 * g = (row r, bit i); a step of 256 threads advances it by 1 024 elements = (dr rows, di bits)
  const int dr = 1024 / n_bits, di = 1024 - dr * n_bits;
  int r = (4 * (int)threadIdx.x) / n_bits, i = 4 * (int)threadIdx.x - r * n_bits;
  for (int g = threadIdx.x; g < total4; g += 256) {
    float v[4];
    if (i + 4 <= n_bits) {                                               // the four outputs belong to one row
      const uint8_t* pb = sm_bits + r * bytes_per_row + (i >> 3);
      const uint32_t bits = (((uint32_t)pb[0] | ((uint32_t)pb[1] << 8)) >> (i & 7));
      const float2 two = sval[r];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ((bits >> k) & 1u) ? two.y : two.x;
    } else {
      int rr = r, ii = i;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 two = sval[rr];
        v[k] = ((sm_bits[rr * bytes_per_row + (ii >> 3)] >> (ii & 7)) & 1) ? two.y : two.x;
        if (++ii == n_bits) ii = 0, ++rr;
      }
    }
    reinterpret_cast<float4*>(obase)[g] = make_float4(v[0], v[1], v[2], v[3]);
    i += di, r += dr;
    if (i >= n_bits) i -= n_bits, ++r;
  }
  for (int e = total4 * 4 + threadIdx.x; e < total; e += 256) {
    const int rr = e / n_bits, ii = e - rr * n_bits;
    obase[e] = ((sm_bits[rr * bytes_per_row + (ii >> 3)] >> (ii & 7)) & 1) ? sval[rr].y : sval[rr].x;
  }
}

__device__ __forceinline__ double block_sum_double(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)blockDim.x / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) u8_zscore_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int n) {
  __shared__ double red[8];
  const uint8_t* src = img + (size_t)blockIdx.x * n;
  float* dst = out + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)((float)src[i] / 255.0f);
  const double mean = block_sum_double(s, red) / n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    double d = (double)((float)src[i] / 255.0f) - mean;
    q += d * d;
  }
  double sd = sqrt(block_sum_double(q, red) / n);
  if (sd == 0.0) sd = 1.0;
  for (int i = threadIdx.x; i < n; i += 256) dst[i] = (float)(((double)((float)src[i] / 255.0f) - mean) / sd);
}

// Vector form for rows that are whole 16-byte groups (n % 16 == 0, aligned base): 128-bit loads, exact INTEGER sums for
// the statistics (sum u < 2^32, sum u^2 < 2^40), and -- since a depiction has only 256 distinct input values -- a
// 256-entry table per image holding ((float)(v / 255.f) - mean) / sd evaluated in double exactly as the scalar kernel
// evaluates it per pixel, so the write pass is table look-ups + 128-bit stores (6 B per pixel of traffic instead of 7,
// and no double division per pixel).
__global__ void __launch_bounds__(256) u8_zscore_vec_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int n) {
  __shared__ double red[8];
  __shared__ float lut[256];
  const uint4* src = reinterpret_cast<const uint4*>(img + (size_t)blockIdx.x * n);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)blockIdx.x * n);
  const int n16 = n / 16;
  unsigned long long su = 0, sq = 0;
  for (int i = threadIdx.x; i < n16; i += 256) {
    const uint4 w = src[i];
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned u = (ws[k] >> (8 * b)) & 255u;
        su += u;
        sq += u * u;
      }
  }
  const double s1 = block_sum_double((double)su, red), s2 = block_sum_double((double)sq, red);
  const double mean = s1 / (255.0 * n);
  double var = s2 / (255.0 * 255.0 * n) - mean * mean;
  double sd = sqrt(var > 0.0 ? var : 0.0);
  if (s2 * n == s1 * s1) sd = 0.0;             // constant image: exactly zero spread (integers, no cancellation noise)
  if (sd == 0.0) sd = 1.0;
  lut[threadIdx.x] = (float)(((double)((float)threadIdx.x / 255.0f) - mean) / sd);
  __syncthreads();
  for (int i = threadIdx.x; i < n16; i += 256) {
    const uint4 w = src[i];
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      dst[4 * i + k] = make_float4(lut[ws[k] & 255u], lut[(ws[k] >> 8) & 255u], lut[(ws[k] >> 16) & 255u], lut[ws[k] >> 24]);
  }
}

}  // namespace bbbp

using namespace bbbp;

extern "C" int bbbp_fusion_softmax_mix_fwd_f32(const float* scores, const float* c, float* out, float* w_out, int rows,
                                               int n, int dim, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(scores && c && out && rows >= 0 && n > 0 && n <= 32 && dim > 0, "fusion_softmax_mix_fwd: bad argument");
  if (rows == 0) return BBBP_OK;
  fusion_softmax_mix_fwd_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(scores, c, out, w_out, rows, n, dim);
  return launch_status("fusion_softmax_mix_fwd");
}

extern "C" int bbbp_fusion_softmax_mix_bwd_f32(const float* w, const float* c, const float* dout, float* dc,
                                               float* dscores, int rows, int n, int dim, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(w && c && dout && dc && dscores && rows >= 0 && n > 0 && n <= 32 && dim > 0,
                 "fusion_softmax_mix_bwd: bad argument");
  if (rows == 0) return BBBP_OK;
  fusion_softmax_mix_bwd_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(w, c, dout, dc, dscores, rows, n, dim);
  return launch_status("fusion_softmax_mix_bwd");
}

extern "C" int bbbp_scaled_colmean_fwd_f32(const float* x, int ldx, const float* scale, int ld_scale, float* out,
                                           int ld_out, float* colmean_out, int rows, int cols, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x && scale && out && rows > 0 && cols > 0, "scaled_colmean_fwd: bad argument");
  scaled_colmean_fwd_kernel<<<ceil_div(cols, 32), dim3(32, 8), 0, as_stream(stream)>>>(x, ldx, scale, ld_scale, out, ld_out,
                                                                                     colmean_out, rows, cols);
  return launch_status("scaled_colmean_fwd");
}

extern "C" int bbbp_scaled_colmean_bwd_f32(const float* dout, int ld_dout, const float* scale, int ld_scale,
                                           const float* colmean, float* dx, int ldx, float* dscale, int rows, int cols,
                                           bbbp_stream_t stream) {
  BBBP_CHECK_ARG(dout && scale && colmean && dx && dscale && rows > 0 && cols > 0, "scaled_colmean_bwd: bad argument");
  scaled_colmean_bwd_dx_kernel<<<ceil_div(cols, 32), dim3(32, 8), 0, as_stream(stream)>>>(dout, ld_dout, scale, ld_scale, dx,
                                                                                        ldx, rows, cols);
  rowdot_vec_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(dout, ld_dout, colmean, dscale, rows, cols);
  note_launches(1);
  return launch_status("scaled_colmean_bwd");
}

extern "C" int bbbp_act_bwd_f32(const float* dy, int ld_dy, const float* y, int ld_y, float* dx, int ld_dx, int rows,
                                int cols, int act, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(dy && y && dx && rows >= 0 && cols >= 0, "act_bwd: bad argument");
  size_t total = (size_t)rows * cols;
  if (total == 0) return BBBP_OK;
  const bool flat = ld_dy == cols && ld_y == cols && ld_dx == cols && total % 4 == 0 &&
                    (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dx) & 15) == 0;
  if (flat && total >= (1u << 16)) {
    const size_t n4 = total / 4;
    const unsigned blocks = (unsigned)std::min<size_t>(ceil_div(n4, (size_t)512), (size_t)current_sm_count() * 16);
    act_bwd_flat4_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y),
                                                                reinterpret_cast<float4*>(dx), n4, act);
    return launch_status("act_bwd (flat)");
  }
  act_bwd_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(dy, ld_dy, y, ld_y, dx, ld_dx, rows,
                                                                                         cols, act);
  return launch_status("act_bwd");
}

extern "C" int bbbp_scale_by_device_scalar_f32(const float* x, const float* scalar, float* y, size_t n,
                                               bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x && scalar && y, "scale_by_device_scalar: null operand");
  if (n == 0) return BBBP_OK;
  scale_by_device_scalar_kernel<<<(unsigned)ceil_div(n, (size_t)256), 256, 0, as_stream(stream)>>>(x, scalar, y, n);
  return launch_status("scale_by_device_scalar");
}

static int colsum_chunks(int rows) { return rows <= 2048 ? 1 : (rows / 512 < 1024 ? rows / 512 : 1024); }

extern "C" size_t bbbp_colsum_workspace(int rows, int cols) {
  const int chunks = colsum_chunks(rows);
  return chunks > 1 && cols > 0 ? (size_t)chunks * cols * sizeof(float) : 0;
}

extern "C" int bbbp_colsum_f32(const float* x, int ldx, float* out, int rows, int cols, float* workspace, size_t workspace_bytes,
                               bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x && out && rows >= 0 && cols > 0, "colsum: bad argument");
  const int chunks = colsum_chunks(rows);
  if (chunks == 1) {
    colsum_kernel<<<ceil_div(cols, 32), dim3(32, 8), 0, as_stream(stream)>>>(x, ldx, out, rows, cols);
    return launch_status("colsum");
  }
  const size_t need = bbbp_colsum_workspace(rows, cols);
  if (!workspace || workspace_bytes < need) {
    set_error("colsum: %d rows need %zu workspace bytes, got %zu", rows, need, workspace_bytes);
    return BBBP_EWORKSPACE;
  }
  const int rows_per_chunk = ceil_div(rows, chunks);
  colsum_partial_kernel<<<dim3(ceil_div(cols, 32), chunks), dim3(32, 8), 0, as_stream(stream)>>>(x, ldx, workspace, rows, cols,
                                                                                                 rows_per_chunk);
  int st = launch_status("colsum partial");
  if (st != BBBP_OK) return st;
  colsum_kernel<<<ceil_div(cols, 32), dim3(32, 8), 0, as_stream(stream)>>>(workspace, cols, out, chunks, cols);
  return launch_status("colsum");
}

extern "C" int bbbp_copy2d_f32(const float* src, int ld_src, float* dst, int ld_dst, int rows, int cols,
                               bbbp_stream_t stream) {
  BBBP_CHECK_ARG(src && dst && rows >= 0 && cols >= 0, "copy2d: bad argument");
  size_t total = (size_t)rows * cols;
  if (total == 0) return BBBP_OK;
  copy2d_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, rows, cols);
  return launch_status("copy2d");
}

extern "C" int bbbp_cast_bf16(const float* src, int ld_src, void* dst_bf16, int ld_dst, int rows, int cols, int cols_pad,
                              bbbp_stream_t stream) {
  BBBP_CHECK_ARG(src && dst_bf16 && rows >= 0 && cols >= 0 && cols_pad >= cols && ld_dst >= cols_pad, "cast_bf16: bad argument");
  size_t total = (size_t)rows * cols_pad;
  if (total == 0) return BBBP_OK;
  cast_bf16_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(
      src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), ld_dst, rows, cols, cols_pad);
  return launch_status("cast_bf16");
}

namespace bbbp {
template <typename T>
__global__ void __launch_bounds__(256) fill_zero_kernel(T* __restrict__ dst, size_t rows, size_t row_elems, size_t pitch_elems) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= rows * row_elems) return;
  dst[(i / row_elems) * pitch_elems + i % row_elems] = T{};
}
}  // namespace bbbp

extern "C" int bbbp_fill_zero(void* dst, long long rows, long long row_bytes, long long pitch_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dst && rows >= 0 && row_bytes >= 0 && pitch_bytes >= row_bytes, "fill_zero: bad argument");
  if (rows == 0 || row_bytes == 0) return BBBP_OK;
  const uintptr_t mix = (uintptr_t)dst | (uintptr_t)row_bytes | (uintptr_t)pitch_bytes;
  cudaStream_t s = as_stream(stream);
#define BBBP_FILL(T)                                                                                                \
  do {                                                                                                              \
    const size_t re = (size_t)row_bytes / sizeof(T), total = (size_t)rows * re;                                     \
    fill_zero_kernel<T><<<(unsigned)ceil_div(total, (size_t)256), 256, 0, s>>>(static_cast<T*>(dst), (size_t)rows, re, \
                                                                               (size_t)pitch_bytes / sizeof(T));    \
  } while (0)
  if (mix % 16 == 0) BBBP_FILL(uint4);
  else if (mix % 4 == 0) BBBP_FILL(uint32_t);
  else if (mix % 2 == 0) BBBP_FILL(uint16_t);
  else BBBP_FILL(uint8_t);
#undef BBBP_FILL
  return launch_status("fill_zero");
}

extern "C" int bbbp_cast16(int fmt, const float* src, int ld_src, const float* col_sub, void* dst_hi, void* dst_lo, int ld_dst,
                           int rows, int cols, int cols_pad, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "cast16: bad fmt %d", fmt);
  BBBP_CHECK_ARG(src && dst_hi && rows >= 0 && cols >= 0 && cols_pad >= cols && ld_dst >= cols_pad, "cast16: bad argument");
  BBBP_CHECK_ARG(cols_pad % 8 == 0 && ld_dst % 8 == 0 && ((uintptr_t)dst_hi % 16) == 0 && ((uintptr_t)dst_lo % 16) == 0,
                 "cast16: cols_pad and ld_dst must be multiples of 8, destinations 16-byte aligned");
  const size_t total = (size_t)rows * (cols_pad / 8);
  if (total == 0) return BBBP_OK;
  const unsigned blocks = (unsigned)ceil_div(total, (size_t)256);
  if (fmt == BBBP_FMT_F16)
    cast16_kernel<BBBP_FMT_F16><<<blocks, 256, 0, as_stream(stream)>>>(src, ld_src, col_sub, static_cast<uint16_t*>(dst_hi),
                                                                       static_cast<uint16_t*>(dst_lo), ld_dst, rows, cols, cols_pad);
  else
    cast16_kernel<BBBP_FMT_BF16><<<blocks, 256, 0, as_stream(stream)>>>(src, ld_src, col_sub, static_cast<uint16_t*>(dst_hi),
                                                                        static_cast<uint16_t*>(dst_lo), ld_dst, rows, cols, cols_pad);
  return launch_status("cast16");
}

extern "C" int bbbp_dropout_f32(const float* x, float* y, size_t n, float p, uint64_t seed, uint64_t offset,
                                const uint64_t* seed_dev, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x && y && p >= 0.0f && p < 1.0f, "dropout: p must be in [0,1)");
  if (n == 0) return BBBP_OK;
  size_t quads = ceil_div(n, (size_t)4);
  dropout_kernel<<<(unsigned)ceil_div(quads, (size_t)256), 256, 0, as_stream(stream)>>>(x, y, n, p, 1.0f / (1.0f - p), seed,
                                                                                         offset, seed_dev);
  return launch_status("dropout");
}

extern "C" int bbbp_mse_loss_f32(const float* pred, const float* target, float* loss, float* dpred, int n,
                                 float grad_scale, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(pred && target && loss && n > 0, "mse_loss: bad argument");
  loss_kernel<0><<<1, 1024, 0, as_stream(stream)>>>(pred, target, loss, dpred, n, grad_scale);
  return launch_status("mse_loss");
}

extern "C" int bbbp_bce_logits_loss_f32(const float* logit, const float* target, float* loss, float* dlogit, int n,
                                        float grad_scale, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(logit && target && loss && n > 0, "bce_logits_loss: bad argument");
  loss_kernel<1><<<1, 1024, 0, as_stream(stream)>>>(logit, target, loss, dlogit, n, grad_scale);
  return launch_status("bce_logits_loss");
}

namespace bbbp {
struct SmallBlob { uint32_t w[16]; };
__global__ void store_small_kernel(SmallBlob blob, uint8_t* __restrict__ dst, int n_bytes) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(blob.w);
  if ((int)threadIdx.x < n_bytes) dst[threadIdx.x] = src[threadIdx.x];
}
}  // namespace bbbp

extern "C" int bbbp_store_small(const void* host_src, int n_bytes, void* dst_dev, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(host_src && dst_dev && n_bytes > 0 && n_bytes <= 64, "store_small: 1..64 bytes");
  SmallBlob blob = {};
  memcpy(blob.w, host_src, (size_t)n_bytes);
  store_small_kernel<<<1, 64, 0, as_stream(stream)>>>(blob, static_cast<uint8_t*>(dst_dev), n_bytes);
  return launch_status("store_small");
}

extern "C" int bbbp_adamw_chunk(void) { return bbbp::ADAMW_CHUNK; }

extern "C" int bbbp_adamw_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                                float grad_scale, float* h) {
  BBBP_CHECK_ARG(h && step >= 1, "adamw_hyper: bad argument");
  // hyper-parameters arrive as doubles (Python floats) so every derived scalar is formed exactly as torch forms it
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  // torch forms 1 - beta in double before narrowing; 1.0f - (float)beta would be off by ~1e-5 relative
  h[0] = (float)(1.0 - beta1);
  h[1] = (float)beta2;
  h[2] = (float)(1.0 - beta2);
  h[3] = (float)eps;
  h[4] = (float)(1.0 - lr * weight_decay);
  h[5] = (float)(lr / bc1);
  h[6] = (float)sqrt(bc2);
  h[7] = grad_scale;
  return BBBP_OK;
}

extern "C" int bbbp_adamw_f32(void* const* ptrs, const int64_t* sizes, const int32_t* chunk_tensor,
                              const int64_t* chunk_offset, int ntensors, int nchunks, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int step, float grad_scale,
                              bbbp_stream_t stream) {
  BBBP_CHECK_ARG(ptrs && sizes && chunk_tensor && chunk_offset && ntensors > 0 && nchunks > 0 && step >= 1,
                 "adamw: bad argument");
  AdamwHyper h;
  bbbp_adamw_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale, h.v);
  adamw_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(ptrs, sizes, chunk_tensor, chunk_offset, ntensors, h, nullptr);
  return launch_status("adamw");
}

extern "C" int bbbp_adamw_dev_f32(void* const* ptrs, const int64_t* sizes, const int32_t* chunk_tensor,
                                  const int64_t* chunk_offset, int ntensors, int nchunks, const float* hyper_dev,
                                  bbbp_stream_t stream) {
  BBBP_CHECK_ARG(ptrs && sizes && chunk_tensor && chunk_offset && ntensors > 0 && nchunks > 0 && hyper_dev,
                 "adamw_dev: bad argument");
  adamw_kernel<<<nchunks, 256, 0, as_stream(stream)>>>(ptrs, sizes, chunk_tensor, chunk_offset, ntensors, AdamwHyper{}, hyper_dev);
  return launch_status("adamw_dev");
}

extern "C" int bbbp_unpack_zscore_f32(const uint8_t* packed, int bytes_per_row, float* out, int ld_out, int rows,
                                      int n_bits, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(packed && out && rows >= 0 && n_bits > 0 && bytes_per_row * 8 >= n_bits && ld_out >= n_bits,
                 "unpack_zscore: bad argument");
  if (rows == 0) return BBBP_OK;
  if (ld_out == n_bits && ((uintptr_t)out & 15) == 0 && bytes_per_row <= 256)
    unpack_zscore_vec_kernel<<<ceil_div(rows, UZ_ROWS), 256, UZ_ROWS * bytes_per_row + 4, as_stream(stream)>>>(packed, bytes_per_row,
                                                                                                               out, rows, n_bits);
  else
    unpack_zscore_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(packed, bytes_per_row, out, ld_out, rows, n_bits);
  return launch_status("unpack_zscore");
}

extern "C" int bbbp_u8_zscore_f32(const uint8_t* img, float* out, int rows, int n, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img && out && rows >= 0 && n > 0, "u8_zscore: bad argument");
  if (rows == 0) return BBBP_OK;
  if (n % 16 == 0 && ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    u8_zscore_vec_kernel<<<rows, 256, 0, as_stream(stream)>>>(img, out, n);
  else
    u8_zscore_kernel<<<rows, 256, 0, as_stream(stream)>>>(img, out, n);
  return launch_status("u8_zscore");
}

extern "C" int bbbp_softmax_rows_fwd_f32(const float* scores, float* w, int rows, int n, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(scores && w && rows >= 0 && n > 0 && n <= 32, "softmax_rows_fwd: bad argument");
  if (rows == 0) return BBBP_OK;
  softmax_rows_fwd_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(scores, w, rows, n);
  return launch_status("softmax_rows_fwd");
}

extern "C" int bbbp_softmax_rows_bwd_f32(const float* w, const float* dw, float* dscores, int rows, int n,
                                         bbbp_stream_t stream) {
  BBBP_CHECK_ARG(w && dw && dscores && rows >= 0 && n > 0 && n <= 32, "softmax_rows_bwd: bad argument");
  if (rows == 0) return BBBP_OK;
  softmax_rows_bwd_kernel<<<ceil_div(rows, 4), 128, 0, as_stream(stream)>>>(w, dw, dscores, rows, n);
  return launch_status("softmax_rows_bwd");
}

namespace bbbp {
__global__ void scatter_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst,
                               size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}
}  // namespace bbbp

extern "C" int bbbp_scatter_f32(const float* src, const int64_t* idx, float* dst, size_t n, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(src && idx && dst, "scatter: null operand");
  if (n == 0) return BBBP_OK;
  bbbp::scatter_kernel<<<(unsigned)ceil_div(n, (size_t)256), 256, 0, as_stream(stream)>>>(src, idx, dst, n);
  return launch_status("scatter");
}

extern "C" int bbbp_gather_rows_f32(const float* src, const int64_t* idx, float* dst, int rows, long long cols,
                                    bbbp_stream_t stream) {
  BBBP_CHECK_ARG(src && idx && dst && rows >= 0 && cols > 0, "gather_rows: bad argument");
  if (rows == 0) return BBBP_OK;
  BBBP_CHECK_ARG(rows <= 65535, "gather_rows: at most 65535 rows per call");
  const unsigned bx = (unsigned)ceil_div((size_t)cols, (size_t)(256 * 16));
  gather_rows_kernel<<<dim3(bx < 1 ? 1 : bx, rows), 256, 0, as_stream(stream)>>>(src, idx, dst, rows, (size_t)cols);
  return launch_status("gather_rows");
}
