// LayerNorm (post-norm residual blocks of nn.TransformerEncoderLayer, eps 1e-5, via 20250113.py:75-78)
// and BatchNorm1d (20250113.py:101).  Memory-bound single-purpose kernels: one warp per row for LN
// (warp-shuffle statistics over the TRUE width, never a padded one), column-strip blocks for BN.
#include "common.cuh"
#include "half16.cuh"

namespace bbbp {

constexpr int LN_WARPS = 4;

__global__ void __launch_bounds__(LN_WARPS * 32) add_layernorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ sum_out, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, uint16_t* __restrict__ y16, int ld16, int rows, int dim, float eps, int ld_x,
    int ld_res, int ld_y, int fmt) {
  const int row = blockIdx.x * LN_WARPS + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ld_x;
  const float* rr = res ? res + (size_t)row * ld_res : nullptr;
  float s = 0.0f;
  for (int i = lane; i < dim; i += 32) s += xr[i] + (rr ? rr[i] : 0.0f);
  const float mean = warp_sum(s) / dim;
  float v = 0.0f;
  for (int i = lane; i < dim; i += 32) {
    float t = xr[i] + (rr ? rr[i] : 0.0f) - mean;
    v = fmaf(t, t, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / dim + eps);
  for (int i = lane; i < dim; i += 32) {
    float t = xr[i] + (rr ? rr[i] : 0.0f);
    float o = (t - mean) * rstd * gamma[i] + beta[i];
    y[(size_t)row * ld_y + i] = o;
    if (sum_out) sum_out[(size_t)row * dim + i] = t;
    if (y16) y16[(size_t)row * ld16 + i] = cvt16_rt(o, fmt);
  }
  if (y16)
    for (int i = dim + lane; i < ld16; i += 32) y16[(size_t)row * ld16 + i] = 0;
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// The same kernel with the row held in registers (VPL values per lane, dim <= 32 * VPL): x and res are read ONCE instead of
// three times and every lane has VPL (x2) independent loads in flight -- the three-pass version sat at 0.59 of the HBM
// roofline.  Per-lane summation order is unchanged (k ascending = i ascending), so results are bit-identical to it.
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) add_layernorm_fwd_reg_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ sum_out, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, uint16_t* __restrict__ y16, int ld16, int rows, int dim, float eps, int ld_x,
    int ld_res, int ld_y, int fmt) {
  const int row = blockIdx.x * LN_WARPS + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ld_x;
  const float* rr = res ? res + (size_t)row * ld_res : nullptr;
  float v[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < dim ? xr[i] : 0.0f;
  }
  if (rr) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      if (i < dim) v[k] += rr[i];
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) s += v[k];            // lanes beyond dim hold zeros
  const float mean = warp_sum(s) / dim;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const float t = v[k] - mean;
    q = (lane + 32 * k) < dim ? fmaf(t, t, q) : q;
  }
  const float rstd = rsqrtf(warp_sum(q) / dim + eps);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < dim) {
      const float o = (v[k] - mean) * rstd * gamma[i] + beta[i];
      y[(size_t)row * ld_y + i] = o;
      if (sum_out) sum_out[(size_t)row * dim + i] = v[k];
      if (y16) y16[(size_t)row * ld16 + i] = cvt16_rt(o, fmt);
    }
  }
  if (y16)
    for (int i = dim + lane; i < ld16; i += 32) y16[(size_t)row * ld16 + i] = 0;
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

template <typename... Args>
static void launch_add_layernorm_fwd(int rows, int dim, cudaStream_t stream, Args... args) {
  const dim3 grid(ceil_div(rows, LN_WARPS)), block(LN_WARPS * 32);
  if (dim <= 192) add_layernorm_fwd_reg_kernel<6><<<grid, block, 0, stream>>>(args...);
  else if (dim <= 512) add_layernorm_fwd_reg_kernel<16><<<grid, block, 0, stream>>>(args...);
  else if (dim <= 2048) add_layernorm_fwd_reg_kernel<64><<<grid, block, 0, stream>>>(args...);
  else add_layernorm_fwd_kernel<<<grid, block, 0, stream>>>(args...);
}

// dx with dy * gamma and the normalised row held in registers (same per-lane summation order as the two-pass kernel below)
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_dx_reg_kernel(const float* __restrict__ dy,
                                                                             const float* __restrict__ s,
                                                                             const float* __restrict__ mean,
                                                                             const float* __restrict__ rstd,
                                                                             const float* __restrict__ gamma,
                                                                             float* __restrict__ dx, int rows, int dim) {
  const int row = blockIdx.x * LN_WARPS + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  const float* dyr = dy + (size_t)row * dim;
  const float* sr = s + (size_t)row * dim;
  float g[VPL], xh[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    g[k] = i < dim ? dyr[i] : 0.0f;
    xh[k] = i < dim ? sr[i] : mu;
  }
  float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < dim) {
      g[k] *= gamma[i];
      xh[k] = (xh[k] - mu) * rs;
      c1 += g[k];
      c2 = fmaf(g[k], xh[k], c2);
    }
  }
  c1 = warp_sum(c1) / dim;
  c2 = warp_sum(c2) / dim;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < dim) dx[(size_t)row * dim + i] = rs * (g[k] - c1 - xh[k] * c2);
  }
}

// Many rows: dx AND the parameter-gradient partials in ONE pass over dy and s (the separate partial kernel re-read both:
// 5 array passes for 3 algorithmic ones).  A warp is a row class (rows c, c + classes, ...): it keeps dy * gamma and the
// normalised row in registers for dx and accumulates its class's dgamma / dbeta per lane; partial row c of the workspace
// is summed over the classes in a fixed order by layernorm_bwd_param_final_kernel (deterministic).
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_fused_kernel(const float* __restrict__ dy,
                                                                            const float* __restrict__ s,
                                                                            const float* __restrict__ mean,
                                                                            const float* __restrict__ rstd,
                                                                            const float* __restrict__ gamma,
                                                                            float* __restrict__ dx, float* __restrict__ part,
                                                                            int rows, int dim, int classes) {
  const int cls = blockIdx.x * LN_WARPS + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (cls >= classes) return;
  float gam[VPL], dg[VPL], db[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    gam[k] = i < dim ? gamma[i] : 0.0f;
    dg[k] = db[k] = 0.0f;
  }
  for (int row = cls; row < rows; row += classes) {
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (size_t)row * dim;
    const float* sr = s + (size_t)row * dim;
    float d[VPL], xh[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      d[k] = i < dim ? dyr[i] : 0.0f;
      xh[k] = i < dim ? sr[i] : mu;
    }
    float c1 = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      xh[k] = (xh[k] - mu) * rs;
      dg[k] = fmaf(d[k], xh[k], dg[k]);
      db[k] += d[k];
      d[k] *= gam[k];
      c1 += d[k];
      c2 = fmaf(d[k], xh[k], c2);
    }
    c1 = warp_sum(c1) / dim;
    c2 = warp_sum(c2) / dim;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int i = lane + 32 * k;
      if (i < dim) dx[(size_t)row * dim + i] = rs * (d[k] - c1 - xh[k] * c2);
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < dim) {
      part[(size_t)cls * dim + i] = dg[k];
      part[(size_t)(classes + cls) * dim + i] = db[k];
    }
  }
}

__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_dx_kernel(const float* __restrict__ dy,
                                                                         const float* __restrict__ s,
                                                                         const float* __restrict__ mean,
                                                                         const float* __restrict__ rstd,
                                                                         const float* __restrict__ gamma,
                                                                         float* __restrict__ dx, int rows, int dim) {
  const int row = blockIdx.x * LN_WARPS + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  const float* dyr = dy + (size_t)row * dim;
  const float* sr = s + (size_t)row * dim;
  float c1 = 0.0f, c2 = 0.0f;
  for (int i = lane; i < dim; i += 32) {
    float g = dyr[i] * gamma[i];
    c1 += g;
    c2 = fmaf(g, (sr[i] - mu) * rs, c2);
  }
  c1 = warp_sum(c1) / dim;
  c2 = warp_sum(c2) / dim;
  for (int i = lane; i < dim; i += 32) {
    float g = dyr[i] * gamma[i];
    dx[(size_t)row * dim + i] = rs * (g - c1 - (sr[i] - mu) * rs * c2);
  }
}

// parts = gridDim.x row classes (rows r = part, part + parts, ...): enough of them to fill the GPU at any row count
__host__ __device__ inline int ln_parts(int rows) { return rows / 8 < 64 ? 64 : (rows / 8 < 4736 ? rows / 8 : 4736); }
__global__ void layernorm_bwd_param_partial_kernel(const float* __restrict__ dy, const float* __restrict__ s,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   float* __restrict__ part, int rows, int dim) {
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col >= dim) return;
  const int parts = gridDim.x;
  float dg = 0.0f, db = 0.0f;
  for (int r = blockIdx.x; r < rows; r += parts) {
    float d = dy[(size_t)r * dim + col];
    dg = fmaf(d, (s[(size_t)r * dim + col] - mean[r]) * rstd[r], dg);
    db += d;
  }
  part[(size_t)blockIdx.x * dim + col] = dg;
  part[(size_t)(parts + blockIdx.x) * dim + col] = db;
}
__global__ void layernorm_bwd_param_final_kernel(const float* __restrict__ part, float* __restrict__ dgamma,
                                                 float* __restrict__ dbeta, int dim, int parts) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= dim) return;
  float dg = 0.0f, db = 0.0f;
  for (int p = 0; p < parts; ++p) {
    dg += part[(size_t)p * dim + col];
    db += part[(size_t)(parts + p) * dim + col];
  }
  dgamma[col] = dg;
  dbeta[col] = db;
}

// Few rows (a reference training batch is 32 molecules): ONE launch.  Blocks [0, row_blocks) compute dx exactly as
// layernorm_bwd_dx_kernel does; the remaining blocks own 32 columns each and reduce over the rows for dgamma / dbeta.
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_small_kernel(const float* __restrict__ dy,
                                                                            const float* __restrict__ s,
                                                                            const float* __restrict__ mean,
                                                                            const float* __restrict__ rstd,
                                                                            const float* __restrict__ gamma,
                                                                            float* __restrict__ dx, float* __restrict__ dgamma,
                                                                            float* __restrict__ dbeta, int rows, int dim,
                                                                            int row_blocks) {
  if ((int)blockIdx.x < row_blocks) {
    const int row = blockIdx.x * LN_WARPS + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (size_t)row * dim;
    const float* sr = s + (size_t)row * dim;
    float c1 = 0.0f, c2 = 0.0f;
    for (int i = lane; i < dim; i += 32) {
      float g = dyr[i] * gamma[i];
      c1 += g;
      c2 = fmaf(g, (sr[i] - mu) * rs, c2);
    }
    c1 = warp_sum(c1) / dim;
    c2 = warp_sum(c2) / dim;
    for (int i = lane; i < dim; i += 32) {
      float g = dyr[i] * gamma[i];
      dx[(size_t)row * dim + i] = rs * (g - c1 - (sr[i] - mu) * rs * c2);
    }
    return;
  }
  // parameter gradients: 32 columns x 4 row lanes per block (lane l walks rows l, l+4, ...; eight independent loads in
  // flight per thread), then a fixed-order combine of the four lanes -- deterministic, and a quarter of the dependent
  // load round trips of one thread per column
  __shared__ float red[2][LN_WARPS][33];
  const int cx = threadIdx.x % 32, ry = threadIdx.x / 32;
  const int col = (blockIdx.x - row_blocks) * 32 + cx;
  float dg = 0.0f, db = 0.0f;
  if (col < dim) {
#pragma unroll 4
    for (int r = ry; r < rows; r += LN_WARPS) {
      const float d = dy[(size_t)r * dim + col];
      dg = fmaf(d, (s[(size_t)r * dim + col] - mean[r]) * rstd[r], dg);
      db += d;
    }
  }
  red[0][ry][cx] = dg;
  red[1][ry][cx] = db;
  __syncthreads();
  if (ry == 0 && col < dim) {
    float tg = 0.0f, tb = 0.0f;
#pragma unroll
    for (int l = 0; l < LN_WARPS; ++l) {
      tg += red[0][l][cx];
      tb += red[1][l][cx];
    }
    dgamma[col] = tg;
    dbeta[col] = tb;
  }
}

// ---- BatchNorm1d: block = 32 channels x 8 row lanes -------------------------------------------------------------
__device__ __forceinline__ float bn_col_reduce(float v, float (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) batchnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ rmean,
                                                            float* __restrict__ rvar, float* __restrict__ y,
                                                            float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                            int rows, int C, int training, float momentum, float eps) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  float mean, rstd;
  if (training) {
    float s = 0.0f;
    if (ok)
      for (int r = threadIdx.y; r < rows; r += 8) s += x[(size_t)r * C + c];
    mean = bn_col_reduce(s, red) / rows;
    float v = 0.0f;
    if (ok)
      for (int r = threadIdx.y; r < rows; r += 8) {
        float t = x[(size_t)r * C + c] - mean;
        v = fmaf(t, t, v);
      }
    const float var = bn_col_reduce(v, red) / rows;  // biased: used to normalise
    rstd = rsqrtf(var + eps);
    if (ok && threadIdx.y == 0) {
      save_mean[c] = mean;
      save_rstd[c] = rstd;
      const float unbiased = rows > 1 ? var * rows / (rows - 1) : var;
      rmean[c] = (1.0f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.0f - momentum) * rvar[c] + momentum * unbiased;
    }
  } else {
    mean = ok ? rmean[c] : 0.0f;
    rstd = ok ? rsqrtf(rvar[c] + eps) : 0.0f;
  }
  if (!ok) return;
  const float g = gamma[c], b = beta[c];
  for (int r = threadIdx.y; r < rows; r += 8) y[(size_t)r * C + c] = (x[(size_t)r * C + c] - mean) * rstd * g + b;
}

// eval-mode forward over many rows: purely elementwise, so parallelise over rows as well (the strip kernel above has
// only channels/32 blocks, far too few when a screening pass normalises thousands of molecules)
__global__ void __launch_bounds__(256) batchnorm_eval_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta,
                                                                 const float* __restrict__ rmean,
                                                                 const float* __restrict__ rvar, float* __restrict__ y,
                                                                 size_t total, int C, float eps) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % C;
  y[i] = (x[i] - rmean[c]) * rsqrtf(rvar[c] + eps) * gamma[c] + beta[c];
}

// training != 0: batch-statistics backward; else running statistics are constants
__global__ void __launch_bounds__(256) batchnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, const float* __restrict__ rvar,
                                                            float* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int rows, int C, int training,
                                                            float eps) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  const float mean = ok ? mean_in[c] : 0.0f;
  const float rstd = ok ? (training ? rstd_in[c] : rsqrtf(rvar[c] + eps)) : 0.0f;
  float sg = 0.0f, sb = 0.0f;
  if (ok)
    for (int r = threadIdx.y; r < rows; r += 8) {
      float d = dy[(size_t)r * C + c];
      sg = fmaf(d, (x[(size_t)r * C + c] - mean) * rstd, sg);
      sb += d;
    }
  sg = bn_col_reduce(sg, red);
  sb = bn_col_reduce(sb, red);
  if (!ok) return;
  if (threadIdx.y == 0) {
    dgamma[c] = sg;
    dbeta[c] = sb;
  }
  const float g = gamma[c];
  for (int r = threadIdx.y; r < rows; r += 8) {
    float d = dy[(size_t)r * C + c];
    if (training) {
      float xh = (x[(size_t)r * C + c] - mean) * rstd;
      dx[(size_t)r * C + c] = g * rstd * (d - sb / rows - xh * sg / rows);
    } else {
      dx[(size_t)r * C + c] = d * g * rstd;
    }
  }
}

// Per-FEATURE standardisation inside chunks of ``chunk`` rows: the preprocessing that produces the ``lso_fixed_1`` pickle the
// canonical script trains on (Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101:
// ``StandardScaler().fit_transform`` on every block of 100 molecules, MACCS and pixel columns side by side).  One thread =
// one column of one chunk; adjacent threads read adjacent columns (coalesced), the 2nd / 3rd pass over the chunk's slab
// (100 x 49 319 floats = 19.7 MB) comes from L2.  The arithmetic restates sklearn's: float64 accumulators, the corrected
// two-pass variance of _incremental_mean_and_var, the near-constant test of _is_constant_feature (scale -> 1), and the
// float32 transform X -= float32(mean); X /= float32(scale).
__global__ void __launch_bounds__(256) standardize_chunks_kernel(const float* __restrict__ x, size_t ld_x,
                                                                 float* __restrict__ out, size_t ld_out, long long rows,
                                                                 int cols, int chunk) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= cols) return;
  const long long r0 = (long long)blockIdx.y * chunk;
  const int n = (int)min((long long)chunk, rows - r0);
  const float* xc = x + (size_t)r0 * ld_x + col;
  double sum = 0.0;
  for (int r = 0; r < n; ++r) sum += (double)xc[(size_t)r * ld_x];
  const double mean = sum / n;
  double corr = 0.0, sq = 0.0;
  for (int r = 0; r < n; ++r) {
    const double d = (double)xc[(size_t)r * ld_x] - mean;
    corr += d;
    sq += d * d;
  }
  const double var = (sq - corr * corr / n) / n;
  const double eps = 2.220446049250313e-16;
  const double t = (double)n * mean * eps;
  const double scale = var <= (double)n * eps * var + t * t ? 1.0 : sqrt(var);
  float* oc = out + (size_t)r0 * ld_out + col;
  const float mean32 = (float)mean, scale32 = (float)scale;     // sklearn 1.9: statistics cast to X's dtype, float32 arithmetic
  for (int r = 0; r < n; ++r) oc[(size_t)r * ld_out] = __fdiv_rn(__fsub_rn(xc[(size_t)r * ld_x], mean32), scale32);
}

}  // namespace bbbp

extern "C" int bbbp_standardize_chunks_f32(const float* x, long long ld_x, float* out, long long ld_out, long long rows, int cols,
                                           int chunk_rows, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x && out && rows >= 0 && cols > 0 && chunk_rows > 0 && ld_x >= cols && ld_out >= cols,
                 "standardize_chunks: bad argument");
  if (rows == 0) return BBBP_OK;
  const long long chunks = (rows + chunk_rows - 1) / chunk_rows;
  BBBP_CHECK_ARG(chunks <= 65535, "standardize_chunks: %lld chunks exceed 65535 per launch", chunks);
  standardize_chunks_kernel<<<dim3(ceil_div(cols, 256), (unsigned)chunks), 256, 0, as_stream(stream)>>>(
      x, (size_t)ld_x, out, (size_t)ld_out, rows, cols, chunk_rows);
  return launch_status("standardize_chunks");
}

extern "C" int bbbp_add_layernorm_fwd_f32(const float* x, const float* res, const float* gamma, const float* beta,
                                          float* y, float* sum_out, float* mean, float* rstd, void* y_bf16, int ld_bf16,
                                          int rows, int dim, float eps, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x && gamma && beta && y && rows >= 0 && dim > 0, "add_layernorm_fwd: bad argument");
  BBBP_CHECK_ARG(!y_bf16 || ld_bf16 >= dim, "add_layernorm_fwd: ld_bf16 < dim");
  if (rows == 0) return BBBP_OK;
  launch_add_layernorm_fwd(rows, dim, as_stream(stream), x, res, gamma, beta, y, sum_out, mean, rstd,
                           reinterpret_cast<uint16_t*>(y_bf16), ld_bf16, rows, dim, eps, dim, dim, dim, (int)BBBP_FMT_BF16);
  return launch_status("add_layernorm_fwd");
}

extern "C" int bbbp_add_layernorm_fwd_pitched_f32(const float* x, int ld_x, const float* res, int ld_res, const float* gamma,
                                                  const float* beta, float* y, int ld_y, void* y_bf16, int ld_bf16, int rows,
                                                  int dim, float eps, bbbp_stream_t stream) {
  return bbbp_add_layernorm_fwd_pitched16(BBBP_FMT_BF16, x, ld_x, res, ld_res, gamma, beta, y, ld_y, y_bf16, ld_bf16, rows, dim,
                                          eps, stream);
}

extern "C" int bbbp_add_layernorm_fwd_pitched16(int fmt, const float* x, int ld_x, const float* res, int ld_res,
                                                const float* gamma, const float* beta, float* y, int ld_y, void* y_bf16,
                                                int ld_bf16, int rows, int dim, float eps, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "add_layernorm_fwd_pitched: bad fmt %d", fmt);
  BBBP_CHECK_ARG(x && gamma && beta && y && rows >= 0 && dim > 0 && ld_x >= dim && ld_y >= dim && (!res || ld_res >= dim),
                 "add_layernorm_fwd_pitched: bad argument");
  BBBP_CHECK_ARG(!y_bf16 || ld_bf16 >= dim, "add_layernorm_fwd_pitched: ld_bf16 < dim");
  if (rows == 0) return BBBP_OK;
  launch_add_layernorm_fwd(rows, dim, as_stream(stream), x, res, gamma, beta, y, (float*)nullptr, (float*)nullptr,
                           (float*)nullptr, reinterpret_cast<uint16_t*>(y_bf16), ld_bf16, rows, dim, eps, ld_x, ld_res, ld_y, fmt);
  return launch_status("add_layernorm_fwd_pitched");
}

extern "C" size_t bbbp_layernorm_bwd_workspace(int rows, int dim) {
  return rows > 0 && dim > 0 ? (size_t)2 * bbbp::ln_parts(rows) * dim * sizeof(float) : 0;
}

extern "C" int bbbp_layernorm_bwd_f32(const float* dy, const float* s, const float* mean, const float* rstd,
                                      const float* gamma, float* dx, float* dgamma, float* dbeta, int rows, int dim,
                                      float* workspace, size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dy && s && mean && rstd && gamma && dx && dgamma && dbeta && rows > 0 && dim > 0,
                 "layernorm_bwd: bad argument");
  const int parts = ln_parts(rows);
  size_t need = bbbp_layernorm_bwd_workspace(rows, dim);
  if (!workspace || workspace_bytes < need) {
    set_error("layernorm_bwd: needs %zu workspace bytes, got %zu", need, workspace_bytes);
    return BBBP_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (rows <= 256) {
    const int row_blocks = ceil_div(rows, LN_WARPS);
    layernorm_bwd_small_kernel<<<row_blocks + ceil_div(dim, 32), LN_WARPS * 32, 0, st>>>(
        dy, s, mean, rstd, gamma, dx, dgamma, dbeta, rows, dim, row_blocks);
    return launch_status("layernorm_bwd (small)");
  }
  if (dim <= 512) {
    const dim3 grid(ceil_div(parts, LN_WARPS)), block(LN_WARPS * 32);
    if (dim <= 192) layernorm_bwd_fused_kernel<6><<<grid, block, 0, st>>>(dy, s, mean, rstd, gamma, dx, workspace, rows, dim, parts);
    else layernorm_bwd_fused_kernel<16><<<grid, block, 0, st>>>(dy, s, mean, rstd, gamma, dx, workspace, rows, dim, parts);
    layernorm_bwd_param_final_kernel<<<ceil_div(dim, 128), 128, 0, st>>>(workspace, dgamma, dbeta, dim, parts);
    note_launches(1);
    return launch_status("layernorm_bwd (fused)");
  }
  layernorm_bwd_dx_kernel<<<ceil_div(rows, LN_WARPS), LN_WARPS * 32, 0, st>>>(dy, s, mean, rstd, gamma, dx, rows, dim);
  layernorm_bwd_param_partial_kernel<<<dim3(parts, ceil_div(dim, 128)), 128, 0, st>>>(dy, s, mean, rstd, workspace, rows, dim);
  layernorm_bwd_param_final_kernel<<<ceil_div(dim, 128), 128, 0, st>>>(workspace, dgamma, dbeta, dim, parts);
  note_launches(2);
  return launch_status("layernorm_bwd");
}

extern "C" int bbbp_batchnorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, float* y, float* save_mean, float* save_rstd, int rows,
                                      int channels, int training, float momentum, float eps, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(x && gamma && beta && running_mean && running_var && y && rows > 0 && channels > 0,
                 "batchnorm_fwd: bad argument");
  BBBP_CHECK_ARG(!training || (save_mean && save_rstd), "batchnorm_fwd: training needs save_mean/save_rstd");
  BBBP_CHECK_ARG(!training || rows > 1, "batchnorm_fwd: Expected more than 1 value per channel when training");
  if (!training && rows >= 64) {
    const size_t total = (size_t)rows * channels;
    batchnorm_eval_fwd_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(
        x, gamma, beta, running_mean, running_var, y, total, channels, eps);
    return launch_status("batchnorm_fwd (eval)");
  }
  batchnorm_fwd_kernel<<<ceil_div(channels, 32), dim3(32, 8), 0, as_stream(stream)>>>(
      x, gamma, beta, running_mean, running_var, y, save_mean, save_rstd, rows, channels, training, momentum, eps);
  return launch_status("batchnorm_fwd");
}

extern "C" int bbbp_batchnorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* save_mean,
                                      const float* save_rstd, float* dx, float* dgamma, float* dbeta, int rows,
                                      int channels, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dy && x && gamma && save_mean && save_rstd && dx && dgamma && dbeta && rows > 0 && channels > 0,
                 "batchnorm_bwd: bad argument");
  batchnorm_bwd_kernel<<<ceil_div(channels, 32), dim3(32, 8), 0, as_stream(stream)>>>(
      dy, x, gamma, save_mean, save_rstd, nullptr, dx, dgamma, dbeta, rows, channels, 1, 0.0f);
  return launch_status("batchnorm_bwd");
}

extern "C" int bbbp_batchnorm_eval_bwd_f32(const float* dy, const float* x, const float* gamma,
                                           const float* running_mean, const float* running_var, float* dx,
                                           float* dgamma, float* dbeta, int rows, int channels, float eps,
                                           bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(dy && x && gamma && running_mean && running_var && dx && dgamma && dbeta && rows > 0 && channels > 0,
                 "batchnorm_eval_bwd: bad argument");
  batchnorm_bwd_kernel<<<ceil_div(channels, 32), dim3(32, 8), 0, as_stream(stream)>>>(
      dy, x, gamma, running_mean, nullptr, running_var, dx, dgamma, dbeta, rows, channels, 0, eps);
  return launch_status("batchnorm_eval_bwd");
}
