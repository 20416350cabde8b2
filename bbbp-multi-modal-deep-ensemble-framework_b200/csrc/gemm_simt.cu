// fp32 CUDA-core GEMM with fused bias/activation epilogue.  Validation-grade arithmetic (plain
// fmaf chain in k order, deterministic) for BBBP_PREC_FP32 and for the training path's dgrad/wgrad.
// Replaces the ATen addmm calls behind nn.Linear at 20250113.py:80,92,53-55,99-106 and the encoder
// projections (torch.nn.TransformerEncoderLayer via 20250113.py:75-78).
#include "common.cuh"

namespace bbbp {

constexpr int TM = 64, TN = 64, TK = 16;

// element (m,k) of A at A[m*sAm + k*sAk]; element (k,n) of B at B[k*sBk + n*sBn]
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, long sAm,
                                                       long sAk, const float* __restrict__ B, long sBk, long sBn,
                                                       float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                                       int act, int accumulate, int k_per_split,
                                                       float* __restrict__ partial) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};

  const bool a_k_fast = (sAk == 1);
  const bool b_n_fast = (sBn == 1);

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (a_k_fast) { kk = idx % TK; mm = idx / TK; } else { mm = idx % TM; kk = idx / TM; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? A[gm * sAm + gk * sAk] : 0.0f;
      int nn, kb;
      if (b_n_fast) { nn = idx % TN; kb = idx / TN; } else { kb = idx % TK; nn = idx / TK; }
      int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < kend) ? B[gkb * sBk + gn * sBn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (partial) {
        partial[((size_t)blockIdx.z * M + gm) * N + gn] = acc[i][j];
      } else {
        float v = acc[i][j] + (bias ? bias[gn] : 0.0f);
        v = apply_act(v, act);
        float* dst = C + (size_t)gm * ldc + gn;
        *dst = accumulate ? (*dst + v) : v;
      }
    }
  }
}

__global__ void splitk_finish_kernel(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C,
                                     int ldc, const float* __restrict__ bias, int act, int accumulate) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  int m = i / N, n = i % N;
  float v = 0.0f;
  for (int s = 0; s < splits; ++s) v += partial[(size_t)s * M * N + i];
  v += bias ? bias[n] : 0.0f;
  v = apply_act(v, act);
  float* dst = C + (size_t)m * ldc + n;
  *dst = accumulate ? (*dst + v) : v;
}

// ---- few output rows (M <= 32: one reference training batch) -------------------------------------------------------
// The tiled kernel above walks K in 16-wide steps with two block barriers each; at M = 32 its grid is a handful of CTAs
// and every step is a full global-load round trip, so a 32 x 167 x 2048 product is ~130 dependent round trips.  Here a
// CTA owns a 32 x 32 output tile and a K slab of up to 256: it issues ALL of the slab's loads at once (one round trip),
// then multiplies out of shared memory.  K slabs are spread over blockIdx.y and summed by splitk_finish_kernel.
constexpr int SK_M = 32, SK_TN = 32, SK_KC = 256, SK_PITCH = 34;

__global__ void __launch_bounds__(256) gemm_f32_skinny_kernel(int M, int N, int K, const float* __restrict__ A, long sAm,
                                                              long sAk, const float* __restrict__ B, long sBk, long sBn,
                                                              float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                                              int act, int accumulate, float* __restrict__ partial) {
  extern __shared__ float sk_smem[];
  float* As = sk_smem;                       // [kc][SK_PITCH], m fast
  float* Bs = sk_smem + SK_KC * SK_PITCH;    // [kc][SK_PITCH], n fast
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * SK_TN;
  const int kbeg = blockIdx.y * SK_KC;
  const int kc = min(SK_KC, K - kbeg);
  // every load of the slab is issued before anything waits (cp.async, zero fill outside the matrix).  The fast global
  // axis runs along the lanes and the slow one along the warps, so the index arithmetic is additions only (the first
  // version divided by the runtime slab width per element and spent more issue slots on addresses than on FMAs);
  // src pointers of masked elements are clamped to the matrix base so no out-of-range address is ever formed
  const int lane = tid % 32, warp = tid / 32;
  if (sAk == 1) {          // k along lanes, rows along warps
    for (int m = warp; m < SK_M; m += 8) {
      const bool ok = m < M;
      const float* src = ok ? A + m * sAm + kbeg : A;
      for (int k = lane; k < kc; k += 32) cp_async_f32(As + k * SK_PITCH + m, ok ? src + k : src, ok);
    }
  } else {                 // rows along lanes (SK_M == 32), k along warps
    const bool ok = lane < M;
    const float* src = ok ? A + lane * sAm + kbeg * sAk : A;
    for (int k = warp; k < kc; k += 8) cp_async_f32(As + k * SK_PITCH + lane, ok ? src + k * sAk : src, ok);
  }
  if (sBk == 1) {
    for (int n = warp; n < SK_TN; n += 8) {
      const bool ok = n0 + n < N;
      const float* src = ok ? B + (n0 + n) * sBn + kbeg : B;
      for (int k = lane; k < kc; k += 32) cp_async_f32(Bs + k * SK_PITCH + n, ok ? src + k : src, ok);
    }
  } else {
    const bool ok = n0 + lane < N;
    const float* src = ok ? B + kbeg * sBk + (n0 + lane) * sBn : B;
    for (int k = warp; k < kc; k += 8) cp_async_f32(Bs + k * SK_PITCH + lane, ok ? src + k * sBk : src, ok);
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = tid % 16, ty = tid / 16;    // 2 x 2 outputs per thread: rows 2ty.., cols 2tx..
  float a00 = 0.0f, a01 = 0.0f, a10 = 0.0f, a11 = 0.0f;
#pragma unroll 8
  for (int k = 0; k < kc; ++k) {
    const float2 a = *reinterpret_cast<const float2*>(As + k * SK_PITCH + 2 * ty);
    const float2 b = *reinterpret_cast<const float2*>(Bs + k * SK_PITCH + 2 * tx);
    a00 = fmaf(a.x, b.x, a00);
    a01 = fmaf(a.x, b.y, a01);
    a10 = fmaf(a.y, b.x, a10);
    a11 = fmaf(a.y, b.y, a11);
  }
  const float acc[2][2] = {{a00, a01}, {a10, a11}};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gm = 2 * ty + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gn = n0 + 2 * tx + j;
      if (gn >= N) continue;
      if (partial) {
        partial[((size_t)blockIdx.y * M + gm) * N + gn] = acc[i][j];
      } else {
        float v = apply_act(acc[i][j] + (bias ? bias[gn] : 0.0f), act);
        float* dst = C + (size_t)gm * ldc + gn;
        *dst = accumulate ? (*dst + v) : v;
      }
    }
  }
}

}  // namespace bbbp

extern "C" size_t bbbp_gemm_f32_auto_workspace(int M, int N, int K) {
  using namespace bbbp;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M <= SK_M) {
    const int slabs = ceil_div(K, SK_KC);
    return slabs > 1 ? (size_t)slabs * M * N * sizeof(float) : 0;
  }
  // the tiled kernel: split K until ~2 CTAs per SM exist
  const long tiles = (long)ceil_div(M, TM) * ceil_div(N, TN);
  long split = ceil_div(2L * 148, tiles);
  if (split > K / 128) split = K / 128;
  if (split < 1) split = 1;
  return split > 1 ? (size_t)split * M * N * sizeof(float) : 0;
}

extern "C" int bbbp_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
                             int ldb, float* C, int ldc, const float* bias, int act, int accumulate, int split_k,
                             float* workspace, size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "gemm_f32: negative dimension");
  BBBP_CHECK_ARG(A && B && C, "gemm_f32: null operand");
  if (M == 0 || N == 0) return BBBP_OK;
  if (split_k == 0 && K > 0) {
    // latency mode (training path): the library picks the kernel and the K partition from (M, N, K)
    const size_t need = bbbp_gemm_f32_auto_workspace(M, N, K);
    if (need && (!workspace || workspace_bytes < need)) {
      set_error("gemm_f32: auto mode needs %zu workspace bytes, got %zu", need, workspace_bytes);
      return BBBP_EWORKSPACE;
    }
    if (M <= SK_M) {
      const long sAm = transA ? 1 : lda, sAk = transA ? lda : 1;
      const long sBk = transB ? 1 : ldb, sBn = transB ? ldb : 1;
      const int slabs = ceil_div(K, SK_KC);
      float* partial = slabs > 1 ? workspace : nullptr;
      const size_t smem = (size_t)2 * SK_KC * SK_PITCH * sizeof(float);
      static PerDeviceOnce attr_once;
      if (attr_once.first())
        cudaFuncSetAttribute(gemm_f32_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      gemm_f32_skinny_kernel<<<dim3(ceil_div(N, SK_TN), slabs), 256, smem, as_stream(stream)>>>(
          M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, bias, act, accumulate, partial);
      int st = launch_status("gemm_f32 (skinny)");
      if (st != BBBP_OK || !partial) return st;
      const size_t total = (size_t)M * N;
      splitk_finish_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(partial, slabs, M, N, C, ldc,
                                                                                                 bias, act, accumulate);
      return launch_status("gemm_f32 split-k finish");
    }
    split_k = need ? (int)(need / ((size_t)M * N * sizeof(float))) : 1;
  }
  if (split_k < 1) split_k = 1;
  if (split_k > K) split_k = K > 0 ? K : 1;
  long sAm = transA ? 1 : lda, sAk = transA ? lda : 1;
  long sBk = transB ? 1 : ldb, sBn = transB ? ldb : 1;
  int k_per_split = ceil_div(ceil_div(K, split_k), TK) * TK;
  if (k_per_split == 0) k_per_split = TK;
  split_k = K > 0 ? ceil_div(K, k_per_split) : 1;
  float* partial = nullptr;
  if (split_k > 1) {
    size_t need = (size_t)split_k * M * N * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("gemm_f32: split_k=%d needs %zu workspace bytes, got %zu", split_k, need, workspace_bytes);
      return BBBP_EWORKSPACE;
    }
    partial = workspace;
  }
  dim3 grid(ceil_div(N, TN), ceil_div(M, TM), split_k);
  gemm_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, bias, act,
                                                        accumulate, k_per_split, partial);
  int st = launch_status("gemm_f32");
  if (st != BBBP_OK || !partial) return st;
  size_t total = (size_t)M * N;
  splitk_finish_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(partial, split_k, M, N, C,
                                                                                             ldc, bias, act, accumulate);
  return launch_status("gemm_f32 split-k finish");
}
