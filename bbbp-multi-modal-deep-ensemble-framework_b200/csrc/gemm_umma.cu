// tcgen05 / TMEM / TMA GEMM for every nn.Linear on the path (bf16 operands, fp32 accumulation in TMEM):
//   out = act(A[M,K] * W[N,K]^T + bias) (+ residual)
// Call sites replaced: encoder in_proj / out_proj / linear1 / linear2 (torch.nn.TransformerEncoderLayer via
// 20250113.py:75-78), fingerprint_fc :80, image_cnn Linear(65536,128) :92, fusion heads :53-55, head fc :99-106,
// PCA transform (_opt.py:30-33).
//
// One CTA computes one 128 x BN output tile over a K range (split-K over blockIdx.z):
//   warp 0   TMA producer   cp.async.bulk.tensor 2-D, 128B-swizzled 128x64 / BNx64 bf16 boxes, STAGES-deep mbarrier ring
//   warp 1   MMA issuer     one elected thread: tcgen05.mma.cta_group::1.kind::f16 M=128 N=BN K=16, accumulators in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      tcgen05.ld 32 lanes x 32 columns -> bias / activation / residual -> global (fp32 and/or bf16)
// K tails and M/N tails are covered by TMA out-of-bounds zero fill; nothing has to be padded in HBM except the row
// pitch (multiple of 8 elements).
#include "common.cuh"
#include "umma.cuh"

namespace bbbp {

tensormap_encode_fn get_tensormap_encoder() {
  static tensormap_encode_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return nullptr;
  }
  fn = reinterpret_cast<tensormap_encode_fn>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  tensormap_encode_fn enc = get_tensormap_encoder();
  if (!enc) return BBBP_ECUDA;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return BBBP_ECUDA;
  }
  return BBBP_OK;
}

namespace gemm {
using namespace sm100;

constexpr int BM = 128, BK = 64;
constexpr int THREADS = 256;

template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 128 ? 3 : 4;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // tiles | full[STAGES] empty[STAGES] accum | tmem slot ; +1024 for manual alignment of the tile area
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16;
};

template <int BN>
__global__ void __launch_bounds__(THREADS) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, int M, int N,
                                                            int total_kb, int kb_per_split,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ residual, int ld_res,
                                                            float* __restrict__ out, int ld_out,
                                                            __nv_bfloat16* __restrict__ out16, int ld_out16, int act,
                                                            float* __restrict__ partial, int M_pad, int N_pad) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = tiles;
  uint8_t* sB = tiles + C::STAGES * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* accum = empty + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int num_kb = min(total_kb, kb_begin + kb_per_split) - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, BN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (i / C::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
        tma_load_2d(&tmA, &full[s], sA + s * C::A_BYTES, (kb_begin + i) * BK, m0);
        tma_load_2d(&tmB, &full[s], sB + s * C::B_BYTES, (kb_begin + i) * BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::STAGES;
        const uint32_t ph = (i / C::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sA + s * C::A_BYTES), b_addr = smem_u32(sB + s * C::B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major, 128B swizzle: 8-row atoms 1024 B apart (SBO); one atom along K, advance 32 B per UMMA_K
          const uint64_t ad = make_smem_desc(a_addr + k * 32, 0, 1024, kLayoutSw128);
          const uint64_t bd = make_smem_desc(b_addr + k * 32, 0, 1024, kLayoutSw128);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        umma_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(accum);  // accumulator complete
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp - 4;  // TMEM lane quadrant == warp_id % 4
    mbar_wait(accum, 0);
    tc_fence_after_sync();
    const int row = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (partial) {
        float4* dst = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.z * M_pad + row) * N_pad + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
      } else if (row < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < N) {
            float v = __uint_as_float(r[j]) + (bias ? __ldg(bias + col) : 0.0f);
            v = apply_act(v, act);
            if (residual) v += residual[(size_t)row * ld_res + col];
            if (out) out[(size_t)row * ld_out + col] = v;
            if (out16) out16[(size_t)row * ld_out16 + col] = __float2bfloat16(v);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BN);
  }
}

__global__ void __launch_bounds__(256) splitk_finish_bf16_kernel(const float* __restrict__ partial, int splits, int M,
                                                                 int N, int M_pad, int N_pad,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ residual, int ld_res,
                                                                 float* __restrict__ out, int ld_out,
                                                                 __nv_bfloat16* __restrict__ out16, int ld_out16,
                                                                 int act) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int m = i / N, n = i % N;
  float v = 0.0f;
  for (int s = 0; s < splits; ++s) v += partial[((size_t)s * M_pad + m) * N_pad + n];
  v += bias ? bias[n] : 0.0f;
  v = apply_act(v, act);
  if (residual) v += residual[(size_t)m * ld_res + n];
  if (out) out[(size_t)m * ld_out + n] = v;
  if (out16) out16[(size_t)m * ld_out16 + n] = __float2bfloat16(v);
}

template <int BN>
int launch(int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias, const float* residual,
           int ld_res, float* out, int ld_out, void* out16, int ld_out16, int act, int split_k, float* partial,
           cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB;
  int st = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  st = make_tmap_bf16_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  const int total_kb = ceil_div(K, BK);
  int kb_per_split = ceil_div(total_kb, split_k);
  split_k = ceil_div(total_kb, kb_per_split);
  const int M_pad = ceil_div(M, 128) * 128, N_pad = ceil_div(N, 128) * 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    attr_set = true;
  }
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), split_k);
  gemm_bf16_kernel<BN><<<grid, THREADS, C::SMEM_BYTES, stream>>>(
      tmA, tmB, M, N, total_kb, kb_per_split, bias, residual, ld_res, out, ld_out,
      reinterpret_cast<__nv_bfloat16*>(out16), ld_out16, act, split_k > 1 ? partial : nullptr, M_pad, N_pad);
  st = launch_status("gemm_bf16");
  if (st != BBBP_OK || split_k <= 1) return st;
  const size_t total = (size_t)M * N;
  splitk_finish_bf16_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, stream>>>(
      partial, split_k, M, N, M_pad, N_pad, bias, residual, ld_res, out, ld_out, reinterpret_cast<__nv_bfloat16*>(out16),
      ld_out16, act);
  return launch_status("gemm_bf16 split-k finish");
}

}  // namespace gemm
}  // namespace bbbp

extern "C" size_t bbbp_gemm_bf16_workspace(int M, int N, int split_k) {
  if (split_k <= 1 || M <= 0 || N <= 0) return 0;
  const size_t M_pad = (size_t)bbbp::ceil_div(M, 128) * 128, N_pad = (size_t)bbbp::ceil_div(N, 128) * 128;
  return (size_t)split_k * M_pad * N_pad * sizeof(float);
}

extern "C" int bbbp_gemm_bf16(int M, int N, int K, const void* A_bf16, int lda, const void* W_bf16, int ldw,
                              const float* bias, const float* residual, int ld_res, float* out_f32, int ld_out,
                              void* out_bf16, int ld_out16, int act, int split_k, void* workspace,
                              size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "gemm_bf16: bad dimension M=%d N=%d K=%d", M, N, K);
  BBBP_CHECK_ARG(A_bf16 && W_bf16, "gemm_bf16: null operand");
  BBBP_CHECK_ARG(out_f32 || out_bf16, "gemm_bf16: no output given");
  BBBP_CHECK_ARG(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0, "gemm_bf16: lda=%d ldw=%d must be >= K=%d and multiples of 8", lda, ldw, K);
  BBBP_CHECK_ARG(((uintptr_t)A_bf16 % 16) == 0 && ((uintptr_t)W_bf16 % 16) == 0, "gemm_bf16: operands must be 16-byte aligned");
  BBBP_CHECK_ARG(!residual || ld_res >= N, "gemm_bf16: ld_res < N");
  BBBP_CHECK_ARG((!out_f32 || ld_out >= N) && (!out_bf16 || ld_out16 >= N), "gemm_bf16: output pitch < N");
  if (M == 0 || N == 0) return BBBP_OK;
  if (split_k < 1) split_k = 1;
  const int total_kb = ceil_div(K, gemm::BK);
  if (split_k > total_kb) split_k = total_kb;
  if (split_k > 1) {
    // the effective split count may shrink inside launch(); size for the requested one
    const size_t need = bbbp_gemm_bf16_workspace(M, N, split_k);
    if (!workspace || workspace_bytes < need) {
      set_error("gemm_bf16: split_k=%d needs %zu workspace bytes, got %zu", split_k, need, workspace_bytes);
      return BBBP_EWORKSPACE;
    }
  }
  cudaStream_t s = as_stream(stream);
  if (N <= 64)
    return gemm::launch<64>(M, N, K, A_bf16, lda, W_bf16, ldw, bias, residual, ld_res, out_f32, ld_out, out_bf16, ld_out16,
                            act, split_k, static_cast<float*>(workspace), s);
  return gemm::launch<128>(M, N, K, A_bf16, lda, W_bf16, ldw, bias, residual, ld_res, out_f32, ld_out, out_bf16, ld_out16,
                           act, split_k, static_cast<float*>(workspace), s);
}
