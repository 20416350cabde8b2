// tcgen05 / TMEM / TMA GEMM for every nn.Linear on the path and for the two attention contractions
// (bf16 operands, fp32 accumulation in TMEM):
//   out[b] = act(A[b][M,K] * W[b][N,K]^T + bias) (+ residual[b])                      (linear epilogue)
//   P[b]   = softmax_rows(scale * A[b] * W[b]^T)  as bf16                               (attention-score epilogue)
// Call sites replaced: encoder in_proj / out_proj / linear1 / linear2 and the softmax(QK^T/sqrt(d)) V products inside
// nn.MultiheadAttention (torch.nn.TransformerEncoderLayer via 20250113.py:75-78, 110-111), fingerprint_fc :80,
// image_cnn Linear(65536,128) :92, fusion heads :53-55, head fc :99-106, PCA transform (_opt.py:30-33).
//
// One CTA computes one 128 x BN output tile of one batch entry over a K range (blockIdx.z = batch * splits + split):
//   warp 0   TMA producer   cp.async.bulk.tensor 3-D {K, rows, batch}, 128B-swizzled 128x64 / BNx64 bf16 boxes, mbarrier ring
//   warp 1   MMA issuer     one elected thread: tcgen05.mma.cta_group::1.kind::f16 M=128 N=BN K=16, accumulators in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue      tcgen05.ld 32 lanes x 32 columns -> bias / activation / residual (or row softmax) -> global
// K tails and M/N tails are covered by TMA out-of-bounds zero fill; nothing has to be padded in HBM except the row
// pitch (multiple of 8 elements).
//
// Operand formats and split passes (half16.cuh): the 16-bit operands are bf16 or fp16 (Params::fmt); an optional second A
// tile ("lo" part of hi + lo activations) and an optional second W tile add one MMA each per K step into the SAME
// accumulator:  D += A_hi W_hi^T (+ A_lo W_hi^T) (+ A_hi W_lo^T).  The epilogue can emit its 16-bit output as a hi + lo
// pair for the next split GEMM.
#include "common.cuh"
#include "umma.cuh"
#include "half16.cuh"

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

// bf16 [batches][rows][cols] with row pitch ld and batch pitch batch_stride (elements); box = box_rows x box_cols x 1
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t batches,
                      uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  tensormap_encode_fn enc = get_tensormap_encoder();
  if (!enc) return BBBP_ECUDA;
  cuuint64_t dims[3] = {cols, rows, batches};
  cuuint64_t strides[2] = {ld * 2, (batches > 1 ? batch_stride : rows * ld) * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu batches=%llu stride=%llu box=%ux%u",
              (int)r, base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld,
              (unsigned long long)batches, (unsigned long long)batch_stride, box_rows, box_cols);
    return BBBP_ECUDA;
  }
  return BBBP_OK;
}
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  return make_tmap_bf16_3d(map, base, rows, cols, ld, 1, 0, box_rows, box_cols, swizzle);
}

namespace gemm {
using namespace sm100;

constexpr int BM = 128, BK = 64;
// warps: 0 TMA producer, 1 MMA issue, 2 TMEM allocation, 3 idle, 4-11 epilogue.  TWO epilogue warps per TMEM lane quadrant
// (quadrant = warp % 4) take alternate 32-column chunks: the epilogue is a serial, latency-bound walk over the tile's
// columns, and for the encoder's K = 167 products it -- not the MMAs -- is the tile's critical path.
constexpr int THREADS = 384;
enum { EPI_LINEAR = 0, EPI_SOFTMAX = 1 };

constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 220 * 1024;   // tile ring budget (227 KB per CTA minus barriers / bias / alignment slack)
template <int BN>
struct Cfg {
  static constexpr int STAGES = BN == 64 ? 4 : 3;          // ring depth of the plain (one A tile, one W tile) product
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int stage_bytes(int n_a, int n_w) { return n_a * A_BYTES + n_w * B_BYTES; }
  static constexpr int stages(int n_a, int n_w) {
    const int fit = SMEM_BUDGET / stage_bytes(n_a, n_w);
    return fit < STAGES ? (fit < 2 ? 2 : fit) : STAGES;
  }
  // tiles | full[MAX_STAGES] empty[MAX_STAGES] accum | tmem slot | bias ; +1024 for manual alignment of the tile area
  static constexpr int smem_bytes(int n_a, int n_w) {
    return 1024 + stages(n_a, n_w) * stage_bytes(n_a, n_w) + (2 * MAX_STAGES + 1) * 8 + 16 + 16 + BN * 4;
  }
};

struct Params {
  int M, N, total_kb, kb_per_split, splits;
  const float* bias;
  const float* residual;
  int ld_res;
  long long res_bs;
  float* out;
  int ld_out;
  long long out_bs;
  uint16_t* out16;
  int ld_out16;
  long long out16_bs;
  int act;
  float* partial;
  int M_pad, N_pad;
  float scale;  // EPI_SOFTMAX: logits are scale * (A W^T)
  int fmt;      // BBBP_FMT_BF16 | BBBP_FMT_F16: format of A, W and of the 16-bit outputs
  int n_a, n_w; // 1, or 2 when a "lo" tile of A / W rides along (split passes)
  int stages;   // ring depth chosen by the host for this (n_a, n_w)
  uint16_t* out16_lo;   // optional lo part of the 16-bit output (same pitch / batch stride as out16)
  // MN-major operands (the backward products read activations / weights in place, no transposed copies):
  //   a_mn: A is stored [K][M] (M contiguous), b_mn: W is stored [K][N] (N contiguous).  The TMA box is then 64 elements of
  //   M (N) x 64 rows of K, one box per 64-wide block of the tile; in shared memory a block is 64 K-rows of 128 bytes
  //   (128B swizzle), the canonical MN-major layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) with SBO = 1024 (8 K-rows) and
  //   LBO = 8192 (next 64-wide block); a K = 16 MMA step advances the start address by 16 rows = 2048 bytes.
  int a_mn, b_mn;
  // optional fp32 [M][N] addend applied BEFORE the activation (a per-row bias: the background-referenced strict mode adds
  // what the constant part of the activation contributes, bbbp_gemm16_pre)
  const float* pre_add;
  int ld_pre;
  // implicit 3x3 convolution (bbbp_conv3x3_gemm16): the A operand is an NHWC activation read through a 4-D tensor map.
  // A GEMM row is a pixel, an M tile = 128 consecutive pixels = conv_rows whole image rows; K block kb = (tap, 64-channel
  // block) is ONE shifted box {64 channels, W, conv_rows, 1} at (x, y) offset (kw - 1, kh - 1): the im2col matrix never
  // exists, and the zero padding is the TMA's out-of-bounds fill.  conv_cblocks = C / 64 (0: plain GEMM).
  int conv_cblocks, conv_rows, conv_tiles_per_img;
};

// 32 fp32 values -> 32 16-bit values as four 16-byte stores (and the matching lo parts when lo != nullptr)
template <int FMT>
__device__ __forceinline__ void store16_chunk(const float (&v)[32], uint16_t* o, uint16_t* lo) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t hi4[4], lo4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (lo) split16<FMT>(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], hi4[j], lo4[j]);
      else hi4[j] = pack16<FMT>(v[8 * g + 2 * j], v[8 * g + 2 * j + 1]);
    }
    reinterpret_cast<uint4*>(o)[g] = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    if (lo) reinterpret_cast<uint4*>(lo)[g] = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
  }
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int BN, int EPI>
__global__ void __launch_bounds__(THREADS) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmA2,
                                                            const __grid_constant__ CUtensorMap tmB2,
                                                            const __grid_constant__ Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // one ring slot = [A hi][A lo?][W hi][W lo?], every tile 1024-byte aligned (128B-swizzle atoms)
  const int STAGES = p.stages;
  const int stage_bytes = p.n_a * C::A_BYTES + p.n_w * C::B_BYTES;
  const int b_off = p.n_a * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accum = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  float* sBias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));  // bias of this CTA's columns

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  for (int i = threadIdx.x; i < BN; i += THREADS) sBias[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.0f;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int kb_begin = split * p.kb_per_split;
  const int num_kb = min(p.total_kb, kb_begin + p.kb_per_split) - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
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
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        uint8_t* st = tiles + s * stage_bytes;
        const int kc = (kb_begin + i) * BK;
        if (p.conv_cblocks) {
          const int kbg = kb_begin + i, tap = kbg / p.conv_cblocks, cb = kbg - tap * p.conv_cblocks;
          const int img = (int)blockIdx.y / p.conv_tiles_per_img, y0 = ((int)blockIdx.y - img * p.conv_tiles_per_img) * p.conv_rows;
          tma_load_4d(&tmA, &full[s], st, cb * 64, tap % 3 - 1, y0 + tap / 3 - 1, img);
        } else if (p.a_mn) {
          for (int mb = 0; mb < BM / 64; ++mb) tma_load_3d(&tmA, &full[s], st + mb * 8192, m0 + mb * 64, kc, batch);
        } else {
          tma_load_3d(&tmA, &full[s], st, kc, m0, batch);
          if (p.n_a == 2) tma_load_3d(&tmA2, &full[s], st + C::A_BYTES, kc, m0, batch);
        }
        if (p.b_mn) {
          for (int nb = 0; nb < BN / 64; ++nb) tma_load_3d(&tmB, &full[s], st + b_off + nb * 8192, n0 + nb * 64, kc, batch);
        } else {
          tma_load_3d(&tmB, &full[s], st + b_off, kc, n0, batch);
          if (p.n_w == 2) tma_load_3d(&tmB2, &full[s], st + b_off + C::B_BYTES, kc, n0, batch);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(BM, BN, p.fmt) | (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(tiles + s * stage_bytes), b_addr = a_addr + b_off;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major, 128B swizzle: 8-row atoms 1024 B apart (SBO); one atom along K, advance 32 B per UMMA_K
          const uint64_t ad = p.a_mn ? make_smem_desc(a_addr + k * 2048, 8192, 1024, kLayoutSw128)
                                     : make_smem_desc(a_addr + k * 32, 0, 1024, kLayoutSw128);
          const uint64_t bd = p.b_mn ? make_smem_desc(b_addr + k * 2048, 8192, 1024, kLayoutSw128)
                                     : make_smem_desc(b_addr + k * 32, 0, 1024, kLayoutSw128);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0);
          if (p.n_a == 2)      // + A_lo W_hi^T
            umma_bf16(tmem_base, make_smem_desc(a_addr + C::A_BYTES + k * 32, 0, 1024, kLayoutSw128), bd, idesc, true);
          if (p.n_w == 2)      // + A_hi W_lo^T
            umma_bf16(tmem_base, ad, make_smem_desc(b_addr + C::B_BYTES + k * 32, 0, 1024, kLayoutSw128), idesc, true);
        }
        umma_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(accum);  // accumulator complete
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quadrant == warp_id % 4
    const int half = (warp - 4) >> 2;   // which of the quadrant's two epilogue warps
    if (EPI == EPI_SOFTMAX && half != 0) goto done;   // a softmax row is owned by ONE thread (three passes over its lane)
    mbar_wait(accum, 0);
    tc_fence_after_sync();
    const int row = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (EPI == EPI_SOFTMAX) {
      // one thread owns one full row of logits (N <= BN): three passes over its TMEM lane
      const int N = p.N;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      const float sl2 = p.scale * 1.4426950408889634f;  // exp(scale*(x-mx)) = exp2(sl2*(x-mx))
      const float off = mx * sl2;
      float sum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < N) sum += exp2f(fmaf(__uint_as_float(r[j]), sl2, -off));
      }
      const float inv = 1.0f / sum;
      uint16_t* dst = p.out16 + (size_t)batch * p.out16_bs + (size_t)row * p.ld_out16;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (c * 32 >= p.ld_out16) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);  // .sync.aligned: the whole warp loads, only valid rows store
        tmem_ld_wait();
        if (row < p.M) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float v0 = c * 32 + j < N ? exp2f(fmaf(__uint_as_float(r[j]), sl2, -off)) * inv : 0.0f;
            const float v1 = c * 32 + j + 1 < N ? exp2f(fmaf(__uint_as_float(r[j + 1]), sl2, -off)) * inv : 0.0f;
            pk[j / 2] = pack16_rt(v0, v1, p.fmt);
          }
          // ld_out16 is a multiple of 8: write whole 16-byte groups, zero in the pad columns
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (c * 32 + g * 8 < p.ld_out16)
              reinterpret_cast<uint4*>(dst + c * 32)[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
    } else {
      // Every runtime switch (activation, residual, which outputs, vector eligibility) is resolved ONCE PER 32-column
      // chunk, never per element: with one epilogue warp per scheduler the per-element instruction count is what the
      // epilogue costs.  Thread = output row; it owns 32 consecutive columns per chunk (128 B fp32 / 64 B bf16).
      const bool al32 = p.out && p.ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && p.out_bs % 4 == 0;
      const bool al16 = p.out16 && p.ld_out16 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out16) & 15) == 0 && p.out16_bs % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(p.out16_lo) & 15) == 0;
      const bool alres = p.residual && p.ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0 && p.res_bs % 4 == 0;
      const float4* bias4 = reinterpret_cast<const float4*>(sBias);
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int col0 = n0 + c * 32;
        if (!p.partial && col0 >= max(p.N, p.out16 ? p.ld_out16 : 0)) break;
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (p.partial) {
          float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.z * p.M_pad + row) * p.N_pad + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
          continue;
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = bias4[c * 8 + j];  // zero where there is no bias or the column is >= N
          v[4 * j] = __uint_as_float(r[4 * j]) + b.x;
          v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
          v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
        }
        if (p.pre_add && row < p.M) {
          const float* pa = p.pre_add + (size_t)row * p.ld_pre + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += pa[j];
        }
        if (p.act == BBBP_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (p.act == BBBP_ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
        }
        if (row >= p.M) continue;
        const int nvalid = p.N - col0;  // columns of this chunk that belong to the result (may be <= 0 in the bf16 pad)
        if (p.residual) {
          const float* res = p.residual + (size_t)batch * p.res_bs + (size_t)row * p.ld_res + col0;
          if (alres && col0 + 32 <= p.ld_res) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = reinterpret_cast<const float4*>(res)[j];
              v[4 * j] += t.x, v[4 * j + 1] += t.y, v[4 * j + 2] += t.z, v[4 * j + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += res[j];
          }
        }
        if (nvalid < 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j >= nvalid) v[j] = 0.0f;  // pad columns carry zeros
        }
        if (p.out) {
          float* o = p.out + (size_t)batch * p.out_bs + (size_t)row * p.ld_out + col0;
          if (al32 && col0 + 32 <= p.ld_out) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) o[j] = v[j];
          }
        }
        if (p.out16) {
          const size_t at = (size_t)batch * p.out16_bs + (size_t)row * p.ld_out16 + col0;
          uint16_t* o = p.out16 + at;
          uint16_t* olo = p.out16_lo ? p.out16_lo + at : nullptr;
          if (al16 && col0 + 32 <= p.ld_out16) {
            if (p.fmt == BBBP_FMT_F16) store16_chunk<BBBP_FMT_F16>(v, o, olo);
            else store16_chunk<BBBP_FMT_BF16>(v, o, olo);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.ld_out16) {
                o[j] = cvt16_rt(v[j], p.fmt);
                if (olo) olo[j] = cvt16_rt(v[j] - round16_rt(v[j], p.fmt), p.fmt);
              }
          }
        }
      }
    }
  }
done:
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
                                                                 uint16_t* __restrict__ out16,
                                                                 uint16_t* __restrict__ out16_lo, int ld_out16,
                                                                 int act, int fmt, const float* __restrict__ pre_add,
                                                                 int ld_pre) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  const int m = i / N, n = i % N;
  float v = 0.0f;
  for (int s = 0; s < splits; ++s) v += partial[((size_t)s * M_pad + m) * N_pad + n];
  v += bias ? bias[n] : 0.0f;
  if (pre_add) v += pre_add[(size_t)m * ld_pre + n];
  v = apply_act(v, act);
  if (residual) v += residual[(size_t)m * ld_res + n];
  if (out) out[(size_t)m * ld_out + n] = v;
  if (out16) out16[(size_t)m * ld_out16 + n] = cvt16_rt(v, fmt);
  if (out16_lo) out16_lo[(size_t)m * ld_out16 + n] = cvt16_rt(v - round16_rt(v, fmt), fmt);
}

struct Problem {
  int M, N, K, batches;
  const void* A;
  int lda;
  long long a_bs;
  const void* W;
  int ldw;
  long long w_bs;
  Params p;
  const void* A_lo = nullptr;   // optional lo parts (same pitch / batch stride as the hi parts)
  const void* W_lo = nullptr;
  bool a_mn = false, b_mn = false;   // A stored [K][M] / W stored [K][N]
  int conv_n = 0, conv_h = 0, conv_w = 0, conv_c = 0;   // implicit 3x3 convolution over an NHWC activation (conv_c != 0)
};

template <int BN, int EPI>
int launch(const Problem& pr, int split_k, cudaStream_t stream) {
  using C = Cfg<BN>;
  // (fp16 tiles go through the same 2-byte tensor maps: the element type only selects the out-of-bounds fill, zero in both)
  CUtensorMap tmA, tmB, tmA2, tmB2;
  // K-major: rows = M (N), cols = K, box = tile rows x 64.  MN-major: rows = K, cols = M (N), box = 64 K-rows x 64 columns.
  int st = BBBP_OK;
  if (pr.conv_c) {
    // NHWC activation as {C, W, H, N}; one box = 64 channels of conv_rows whole image rows = the 128 x 64 A tile of a K block
    tensormap_encode_fn enc = get_tensormap_encoder();
    if (!enc) return BBBP_ECUDA;
    const cuuint64_t C = (cuuint64_t)pr.conv_c, Wd = (cuuint64_t)pr.conv_w, Hd = (cuuint64_t)pr.conv_h;
    cuuint64_t dims[4] = {C, Wd, Hd, (cuuint64_t)pr.conv_n};
    cuuint64_t strides[3] = {C * 2, Wd * C * 2, Hd * Wd * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)pr.conv_w, (cuuint32_t)(BM / pr.conv_w), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(pr.A), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv3x3_gemm16: cuTensorMapEncodeTiled failed (%d) for %d x %d x %d x %d", (int)r, pr.conv_n, pr.conv_h, pr.conv_w, pr.conv_c);
      return BBBP_ECUDA;
    }
  } else {
    st = pr.a_mn ? make_tmap_bf16_3d(&tmA, pr.A, (uint64_t)pr.K, (uint64_t)pr.M, (uint64_t)pr.lda, (uint64_t)pr.batches,
                                     (uint64_t)pr.a_bs, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B)
                 : make_tmap_bf16_3d(&tmA, pr.A, (uint64_t)pr.M, (uint64_t)pr.K, (uint64_t)pr.lda, (uint64_t)pr.batches,
                                     (uint64_t)pr.a_bs, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (st != BBBP_OK) return st;
  st = pr.b_mn ? make_tmap_bf16_3d(&tmB, pr.W, (uint64_t)pr.K, (uint64_t)pr.N, (uint64_t)pr.ldw, (uint64_t)pr.batches,
                                   (uint64_t)pr.w_bs, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B)
               : make_tmap_bf16_3d(&tmB, pr.W, (uint64_t)pr.N, (uint64_t)pr.K, (uint64_t)pr.ldw, (uint64_t)pr.batches,
                                   (uint64_t)pr.w_bs, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  tmA2 = tmA, tmB2 = tmB;
  if (pr.A_lo) {
    st = make_tmap_bf16_3d(&tmA2, pr.A_lo, (uint64_t)pr.M, (uint64_t)pr.K, (uint64_t)pr.lda, (uint64_t)pr.batches,
                           (uint64_t)pr.a_bs, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != BBBP_OK) return st;
  }
  if (pr.W_lo) {
    st = make_tmap_bf16_3d(&tmB2, pr.W_lo, (uint64_t)pr.N, (uint64_t)pr.K, (uint64_t)pr.ldw, (uint64_t)pr.batches,
                           (uint64_t)pr.w_bs, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != BBBP_OK) return st;
  }
  Params p = pr.p;
  p.n_a = pr.A_lo ? 2 : 1;
  p.n_w = pr.W_lo ? 2 : 1;
  p.a_mn = pr.a_mn, p.b_mn = pr.b_mn;
  p.conv_cblocks = pr.conv_c / 64;
  p.conv_rows = pr.conv_c ? BM / pr.conv_w : 0;
  p.conv_tiles_per_img = pr.conv_c ? pr.conv_h * pr.conv_w / BM : 0;
  p.stages = C::stages(p.n_a, p.n_w);
  const int smem_bytes = C::smem_bytes(p.n_a, p.n_w);
  p.M = pr.M;
  p.N = pr.N;
  p.total_kb = ceil_div(pr.K, BK);
  p.kb_per_split = ceil_div(p.total_kb, split_k);
  p.splits = ceil_div(p.total_kb, p.kb_per_split);
  p.M_pad = ceil_div(pr.M, 128) * 128;
  p.N_pad = ceil_div(pr.N, 128) * 128;
  if (p.splits <= 1) p.partial = nullptr;
  static PerDeviceOnce attr_once;
  if (attr_once.first())    // the largest configuration this instantiation can be asked for
    cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         max(max(C::smem_bytes(1, 1), C::smem_bytes(2, 1)), max(C::smem_bytes(1, 2), C::smem_bytes(2, 2))));
  dim3 grid(ceil_div(pr.N, BN), ceil_div(pr.M, BM), pr.batches * p.splits);
  gemm_bf16_kernel<BN, EPI><<<grid, THREADS, smem_bytes, stream>>>(tmA, tmB, tmA2, tmB2, p);
  st = launch_status("gemm_bf16");
  if (st != BBBP_OK || p.splits <= 1) return st;
  const size_t total = (size_t)pr.M * pr.N;
  splitk_finish_bf16_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, stream>>>(
      p.partial, p.splits, pr.M, pr.N, p.M_pad, p.N_pad, p.bias, p.residual, p.ld_res, p.out, p.ld_out, p.out16, p.out16_lo,
      p.ld_out16, p.act, p.fmt, p.pre_add, p.ld_pre);
  return launch_status("gemm_bf16 split-k finish");
}

// batched [batches][rows][cols] bf16 -> [batches][cols][rows] bf16 (V -> V^T for the P V product), 32x32 smem tiles
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int ld_src,
                                                             long long src_bs, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                                             long long dst_bs, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const __nv_bfloat16* s = src + (size_t)b * src_bs;
  __nv_bfloat16* d = dst + (size_t)b * dst_bs;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? s[(size_t)r * ld_src + c] : __float2bfloat16(0.0f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ld_dst) d[(size_t)c * ld_dst + r] = tile[threadIdx.x][i];  // rows >= `rows` carry zeros
  }
}

// Row softmax of fp32 logits -> bf16 probabilities for attention scopes wider than one CTA tile (S > 256):
// p[r, c] = softmax_c(scale * s[r, c]); one block per row, three passes (max, sum, write) of 128-bit loads.
__global__ void __launch_bounds__(256) softmax_rows_scaled_bf16_kernel(const float* __restrict__ s, size_t ld_s,
                                                                       uint16_t* __restrict__ p, size_t ld_p, int cols,
                                                                       float scale, int fmt) {
  __shared__ float red[8];
  const float* row = s + (size_t)blockIdx.x * ld_s;
  uint16_t* out = p + (size_t)blockIdx.x * ld_p;
  const int n4 = cols / 4;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int i = n4 * 4 + threadIdx.x; i < cols; i += 256) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  const float sl2 = scale * 1.4426950408889634f, off = mx * sl2;
  float sum = 0.0f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    sum += exp2f(fmaf(v.x, sl2, -off)) + exp2f(fmaf(v.y, sl2, -off)) + exp2f(fmaf(v.z, sl2, -off)) + exp2f(fmaf(v.w, sl2, -off));
  }
  for (int i = n4 * 4 + threadIdx.x; i < cols; i += 256) sum += exp2f(fmaf(row[i], sl2, -off));
  sum = warp_sum(sum);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < (int)ld_p / 2; i += 256) {   // ld_p is a multiple of 8: bf16 pairs, zeros in the pad
    const int c = 2 * i;
    const float a = c < cols ? exp2f(fmaf(row[c], sl2, -off)) * inv : 0.0f;
    const float b = c + 1 < cols ? exp2f(fmaf(row[c + 1], sl2, -off)) * inv : 0.0f;
    reinterpret_cast<uint32_t*>(out)[i] = pack16_rt(a, b, fmt);
  }
}

}  // namespace gemm
}  // namespace bbbp

extern "C" size_t bbbp_gemm_bf16_workspace(int M, int N, int split_k) {
  if (split_k <= 1 || M <= 0 || N <= 0) return 0;
  const size_t M_pad = (size_t)bbbp::ceil_div(M, 128) * 128, N_pad = (size_t)bbbp::ceil_div(N, 128) * 128;
  return (size_t)split_k * M_pad * N_pad * sizeof(float);
}

static int check_operands(const char* who, int M, int N, int K, const void* A, int lda, const void* W, int ldw) {
  using namespace bbbp;
  if (M < 0 || N < 0 || K <= 0) {
    set_error("%s: bad dimension M=%d N=%d K=%d", who, M, N, K);
    return BBBP_EINVAL;
  }
  if (!A || !W) {
    set_error("%s: null operand", who);
    return BBBP_EINVAL;
  }
  if (lda < K || ldw < K || lda % 8 || ldw % 8) {
    set_error("%s: lda=%d ldw=%d must be >= K=%d and multiples of 8", who, lda, ldw, K);
    return BBBP_EINVAL;
  }
  if (((uintptr_t)A % 16) || ((uintptr_t)W % 16)) {
    set_error("%s: operands must be 16-byte aligned", who);
    return BBBP_EINVAL;
  }
  return BBBP_OK;
}

static int check_fmt(const char* who, int fmt) {
  if (fmt != BBBP_FMT_BF16 && fmt != BBBP_FMT_F16) {
    bbbp::set_error("%s: fmt=%d (BBBP_FMT_BF16 or BBBP_FMT_F16)", who, fmt);
    return BBBP_EINVAL;
  }
  return BBBP_OK;
}

static int gemm16_impl(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                       const void* W_lo, int ldw, const float* bias, const float* pre_add, int ld_pre, const float* residual,
                       int ld_res, float* out_f32, int ld_out, void* out16_hi, void* out16_lo, int ld_out16, int act,
                       int split_k, void* workspace, size_t workspace_bytes, bbbp_stream_t stream);

extern "C" int bbbp_gemm16(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                           const void* W_lo, int ldw, const float* bias, const float* residual, int ld_res, float* out_f32,
                           int ld_out, void* out16_hi, void* out16_lo, int ld_out16, int act, int split_k, void* workspace,
                           size_t workspace_bytes, bbbp_stream_t stream) {
  return gemm16_impl(fmt, M, N, K, A_hi, A_lo, lda, W_hi, W_lo, ldw, bias, nullptr, 0, residual, ld_res, out_f32, ld_out,
                     out16_hi, out16_lo, ld_out16, act, split_k, workspace, workspace_bytes, stream);
}

extern "C" int bbbp_gemm16_pre(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                               const void* W_lo, int ldw, const float* bias, const float* pre_add, int ld_pre, float* out_f32,
                               int ld_out, void* out16_hi, void* out16_lo, int ld_out16, int act, int split_k, void* workspace,
                               size_t workspace_bytes, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(!pre_add || ld_pre >= N, "gemm16_pre: ld_pre < N");
  return gemm16_impl(fmt, M, N, K, A_hi, A_lo, lda, W_hi, W_lo, ldw, bias, pre_add, ld_pre, nullptr, 0, out_f32, ld_out,
                     out16_hi, out16_lo, ld_out16, act, split_k, workspace, workspace_bytes, stream);
}

static int gemm16_impl(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi,
                       const void* W_lo, int ldw, const float* bias, const float* pre_add, int ld_pre, const float* residual,
                       int ld_res, float* out_f32, int ld_out, void* out16_hi, void* out16_lo, int ld_out16, int act,
                       int split_k, void* workspace, size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  int st = check_fmt("gemm16", fmt);
  if (st != BBBP_OK) return st;
  st = check_operands("gemm16", M, N, K, A_hi, lda, W_hi, ldw);
  if (st != BBBP_OK) return st;
  BBBP_CHECK_ARG(((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)W_lo % 16) == 0, "gemm16: lo operands must be 16-byte aligned");
  BBBP_CHECK_ARG(out_f32 || out16_hi, "gemm16: no output given");
  BBBP_CHECK_ARG(!out16_lo || out16_hi, "gemm16: out16_lo without out16_hi");
  BBBP_CHECK_ARG(!residual || ld_res >= N, "gemm16: ld_res < N");
  BBBP_CHECK_ARG((!out_f32 || ld_out >= N) && (!out16_hi || ld_out16 >= N), "gemm16: output pitch < N");
  if (M == 0 || N == 0) return BBBP_OK;
  if (split_k < 1) split_k = 1;
  const int total_kb = ceil_div(K, gemm::BK);
  if (split_k > total_kb) split_k = total_kb;
  if (split_k > 1) {
    // the effective split count may shrink inside launch(); size for the requested one
    const size_t need = bbbp_gemm_bf16_workspace(M, N, split_k);
    if (!workspace || workspace_bytes < need) {
      set_error("gemm16: split_k=%d needs %zu workspace bytes, got %zu", split_k, need, workspace_bytes);
      return BBBP_EWORKSPACE;
    }
  }
  gemm::Problem pr{};
  pr.M = M, pr.N = N, pr.K = K, pr.batches = 1;
  pr.A = A_hi, pr.A_lo = A_lo, pr.lda = lda, pr.W = W_hi, pr.W_lo = W_lo, pr.ldw = ldw;
  pr.p.bias = bias, pr.p.residual = residual, pr.p.ld_res = ld_res;
  pr.p.pre_add = pre_add, pr.p.ld_pre = ld_pre;
  pr.p.out = out_f32, pr.p.ld_out = ld_out;
  pr.p.out16 = static_cast<uint16_t*>(out16_hi), pr.p.out16_lo = static_cast<uint16_t*>(out16_lo), pr.p.ld_out16 = ld_out16;
  pr.p.act = act, pr.p.partial = static_cast<float*>(workspace), pr.p.fmt = fmt;
  cudaStream_t s = as_stream(stream);
  if (N <= 64) return gemm::launch<64, gemm::EPI_LINEAR>(pr, split_k, s);
  if (N >= 512 && split_k == 1) return gemm::launch<256, gemm::EPI_LINEAR>(pr, split_k, s);
  return gemm::launch<128, gemm::EPI_LINEAR>(pr, split_k, s);
}

// 3x3 convolution (stride 1, padding 1) + bias + activation as an IMPLICIT GEMM over an NHWC 16-bit activation: rows = pixels,
// K = 9 * C in (tap, channel) order, the A tiles fetched as shifted TMA boxes (see Params::conv_cblocks).
extern "C" int bbbp_conv3x3_gemm16(int fmt, const void* x_nhwc, int N, int H, int W, int C, const void* w_taps, int Cout,
                                   const float* bias, int act, void* y_nhwc, bbbp_stream_t stream) {
  using namespace bbbp;
  int st = check_fmt("conv3x3_gemm16", fmt);
  if (st != BBBP_OK) return st;
  BBBP_CHECK_ARG(x_nhwc && w_taps && y_nhwc && N >= 0, "conv3x3_gemm16: null operand");
  BBBP_CHECK_ARG(C >= 64 && C % 64 == 0 && Cout >= 8 && Cout % 8 == 0, "conv3x3_gemm16: C=%d must be a multiple of 64, Cout=%d of 8", C, Cout);
  BBBP_CHECK_ARG(W >= 8 && W <= gemm::BM && gemm::BM % W == 0 && H >= 1 && (H * W) % gemm::BM == 0,
                 "conv3x3_gemm16: W=%d must divide %d and H*W=%d be a multiple of it", W, gemm::BM, H * W);
  BBBP_CHECK_ARG(((uintptr_t)x_nhwc % 16) == 0 && ((uintptr_t)w_taps % 16) == 0 && ((uintptr_t)y_nhwc % 16) == 0,
                 "conv3x3_gemm16: operands must be 16-byte aligned");
  BBBP_CHECK_ARG((long long)N * H * W / gemm::BM <= 65535, "conv3x3_gemm16: %d x %d x %d pixels exceed 65 535 row tiles per launch", N, H, W);
  if (N == 0) return BBBP_OK;
  gemm::Problem pr{};
  pr.M = N * H * W, pr.N = Cout, pr.K = 9 * C, pr.batches = 1;
  pr.A = x_nhwc, pr.lda = 9 * C, pr.W = w_taps, pr.ldw = 9 * C;
  pr.conv_n = N, pr.conv_h = H, pr.conv_w = W, pr.conv_c = C;
  pr.p.bias = bias;
  pr.p.out16 = static_cast<uint16_t*>(y_nhwc), pr.p.ld_out16 = Cout;
  pr.p.act = act, pr.p.fmt = fmt;
  cudaStream_t s = as_stream(stream);
  if (Cout <= 64) return gemm::launch<64, gemm::EPI_LINEAR>(pr, 1, s);
  if (Cout >= 512) return gemm::launch<256, gemm::EPI_LINEAR>(pr, 1, s);     // (256-wide tiles at Cout = 256: measured, no gain)
  return gemm::launch<128, gemm::EPI_LINEAR>(pr, 1, s);
}

// out[M,N] = opA(A) opW(W)^T with either operand read in place in its MN-major storage (see Params::a_mn).
extern "C" int bbbp_gemm16_tn(int fmt, int trans_a, int trans_w, int M, int N, int K, const void* A, int lda, const void* W,
                              int ldw, const float* bias, float* out_f32, int ld_out, void* out16, int ld_out16, int act,
                              int split_k, void* workspace, size_t workspace_bytes, bbbp_stream_t stream) {
  using namespace bbbp;
  int st = check_fmt("gemm16_tn", fmt);
  if (st != BBBP_OK) return st;
  BBBP_CHECK_ARG(M >= 0 && N >= 0 && K > 0 && A && W, "gemm16_tn: bad argument");
  BBBP_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= (trans_a ? M : K) && ldw >= (trans_w ? N : K),
                 "gemm16_tn: pitches must be multiples of 8 and cover the contiguous dimension");
  BBBP_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm16_tn: operands must be 16-byte aligned");
  BBBP_CHECK_ARG(out_f32 || out16, "gemm16_tn: no output given");
  BBBP_CHECK_ARG((!out_f32 || ld_out >= N) && (!out16 || ld_out16 >= N), "gemm16_tn: output pitch < N");
  if (M == 0 || N == 0) return BBBP_OK;
  if (split_k < 1) split_k = 1;
  const int total_kb = ceil_div(K, gemm::BK);
  if (split_k > total_kb) split_k = total_kb;
  if (split_k > 1) {
    const size_t need = bbbp_gemm_bf16_workspace(M, N, split_k);
    if (!workspace || workspace_bytes < need) {
      set_error("gemm16_tn: split_k=%d needs %zu workspace bytes, got %zu", split_k, need, workspace_bytes);
      return BBBP_EWORKSPACE;
    }
  }
  gemm::Problem pr{};
  pr.M = M, pr.N = N, pr.K = K, pr.batches = 1;
  pr.A = A, pr.lda = lda, pr.W = W, pr.ldw = ldw;
  pr.a_mn = trans_a != 0, pr.b_mn = trans_w != 0;
  pr.p.bias = bias;
  pr.p.out = out_f32, pr.p.ld_out = ld_out;
  pr.p.out16 = static_cast<uint16_t*>(out16), pr.p.ld_out16 = ld_out16;
  pr.p.act = act, pr.p.partial = static_cast<float*>(workspace), pr.p.fmt = fmt;
  cudaStream_t s = as_stream(stream);
  if (N <= 64) return gemm::launch<64, gemm::EPI_LINEAR>(pr, split_k, s);
  if (N >= 512 && split_k == 1) return gemm::launch<256, gemm::EPI_LINEAR>(pr, split_k, s);
  return gemm::launch<128, gemm::EPI_LINEAR>(pr, split_k, s);
}

extern "C" int bbbp_gemm_bf16(int M, int N, int K, const void* A_bf16, int lda, const void* W_bf16, int ldw,
                              const float* bias, const float* residual, int ld_res, float* out_f32, int ld_out,
                              void* out_bf16, int ld_out16, int act, int split_k, void* workspace,
                              size_t workspace_bytes, bbbp_stream_t stream) {
  return bbbp_gemm16(BBBP_FMT_BF16, M, N, K, A_bf16, nullptr, lda, W_bf16, nullptr, ldw, bias, residual, ld_res, out_f32,
                     ld_out, out_bf16, nullptr, ld_out16, act, split_k, workspace, workspace_bytes, stream);
}

extern "C" int bbbp_gemm16_batched(int fmt, int batches, int M, int N, int K, const void* A, int lda, long long a_batch_stride,
                                   const void* W, int ldw, long long w_batch_stride, float* out_f32, int ld_out,
                                   long long out_batch_stride, void* out16, int ld_out16, long long out16_batch_stride,
                                   bbbp_stream_t stream) {
  using namespace bbbp;
  int st = check_fmt("gemm16_batched", fmt);
  if (st != BBBP_OK) return st;
  st = check_operands("gemm16_batched", M, N, K, A, lda, W, ldw);
  if (st != BBBP_OK) return st;
  BBBP_CHECK_ARG(batches >= 0 && batches <= 65535, "gemm16_batched: batches=%d out of range", batches);
  BBBP_CHECK_ARG(a_batch_stride % 8 == 0 && w_batch_stride % 8 == 0, "gemm16_batched: batch strides must be multiples of 8");
  BBBP_CHECK_ARG(out_f32 || out16, "gemm16_batched: no output given");
  BBBP_CHECK_ARG((!out_f32 || ld_out >= N) && (!out16 || ld_out16 >= N), "gemm16_batched: output pitch < N");
  if (M == 0 || N == 0 || batches == 0) return BBBP_OK;
  gemm::Problem pr{};
  pr.M = M, pr.N = N, pr.K = K, pr.batches = batches;
  pr.A = A, pr.lda = lda, pr.a_bs = a_batch_stride, pr.W = W, pr.ldw = ldw, pr.w_bs = w_batch_stride;
  pr.p.out = out_f32, pr.p.ld_out = ld_out, pr.p.out_bs = out_batch_stride;
  pr.p.out16 = static_cast<uint16_t*>(out16), pr.p.ld_out16 = ld_out16, pr.p.out16_bs = out16_batch_stride;
  pr.p.fmt = fmt;
  cudaStream_t s = as_stream(stream);
  if (N <= 64) return gemm::launch<64, gemm::EPI_LINEAR>(pr, 1, s);
  return gemm::launch<128, gemm::EPI_LINEAR>(pr, 1, s);
}

extern "C" int bbbp_gemm_bf16_batched(int batches, int M, int N, int K, const void* A_bf16, int lda, long long a_batch_stride,
                                      const void* W_bf16, int ldw, long long w_batch_stride, float* out_f32, int ld_out,
                                      long long out_batch_stride, void* out_bf16, int ld_out16,
                                      long long out16_batch_stride, bbbp_stream_t stream) {
  return bbbp_gemm16_batched(BBBP_FMT_BF16, batches, M, N, K, A_bf16, lda, a_batch_stride, W_bf16, ldw, w_batch_stride, out_f32,
                             ld_out, out_batch_stride, out_bf16, ld_out16, out16_batch_stride, stream);
}

extern "C" int bbbp_attention_scores_softmax16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq,
                                               const void* k, int ldk, long long group_stride, float scale, void* p_out,
                                               int ldp, bbbp_stream_t stream) {
  using namespace bbbp;
  int st = check_fmt("attention_scores_softmax", fmt);
  if (st != BBBP_OK) return st;
  st = check_operands("attention_scores_softmax", seq, seq, head_dim, q, ldq, k, ldk);
  if (st != BBBP_OK) return st;
  BBBP_CHECK_ARG(seq <= 256, "attention_scores_softmax: seq=%d > 256 (one CTA must own a full row of scores)", seq);
  BBBP_CHECK_ARG(groups >= 0 && groups <= 65535 && group_stride % 8 == 0, "attention_scores_softmax: bad groups/stride");
  BBBP_CHECK_ARG(p_out && ldp >= seq && ldp % 8 == 0 && ldp <= 256, "attention_scores_softmax: ldp=%d must be a multiple of 8 in [seq, 256]", ldp);
  if (groups == 0 || seq == 0) return BBBP_OK;
  gemm::Problem pr{};
  pr.M = seq, pr.N = seq, pr.K = head_dim, pr.batches = groups;
  pr.A = q, pr.lda = ldq, pr.a_bs = group_stride, pr.W = k, pr.ldw = ldk, pr.w_bs = group_stride;
  pr.p.out16 = static_cast<uint16_t*>(p_out), pr.p.ld_out16 = ldp, pr.p.out16_bs = (long long)seq * ldp;
  pr.p.scale = scale, pr.p.fmt = fmt;
  return gemm::launch<256, gemm::EPI_SOFTMAX>(pr, 1, as_stream(stream));
}

extern "C" int bbbp_attention_scores_softmax_bf16(int groups, int seq, int head_dim, const void* q_bf16, int ldq,
                                                  const void* k_bf16, int ldk, long long group_stride, float scale,
                                                  void* p_bf16, int ldp, bbbp_stream_t stream) {
  return bbbp_attention_scores_softmax16(BBBP_FMT_BF16, groups, seq, head_dim, q_bf16, ldq, k_bf16, ldk, group_stride, scale,
                                         p_bf16, ldp, stream);
}

extern "C" int bbbp_transpose_bf16(int batches, int rows, int cols, const void* src, int ld_src, long long src_batch_stride,
                                   void* dst, int ld_dst, long long dst_batch_stride, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(src && dst && batches >= 0 && batches <= 65535 && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= rows,
                 "transpose_bf16: bad argument");
  if (batches == 0) return BBBP_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(ld_dst, 32), batches);
  gemm::transpose_bf16_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), ld_src, src_batch_stride, static_cast<__nv_bfloat16*>(dst), ld_dst,
      dst_batch_stride, rows, cols);
  return launch_status("transpose_bf16");
}

extern "C" int bbbp_softmax_rows_scaled_bf16(const float* scores, long long ld_scores, void* p_bf16, long long ld_p,
                                             long long rows, int cols, float scale, bbbp_stream_t stream) {
  return bbbp_softmax_rows_scaled16(BBBP_FMT_BF16, scores, ld_scores, p_bf16, ld_p, rows, cols, scale, stream);
}

extern "C" int bbbp_softmax_rows_scaled16(int fmt, const float* scores, long long ld_scores, void* p_bf16, long long ld_p,
                                          long long rows, int cols, float scale, bbbp_stream_t stream) {
  using namespace bbbp;
  if (check_fmt("softmax_rows_scaled", fmt) != BBBP_OK) return BBBP_EINVAL;
  BBBP_CHECK_ARG(scores && p_bf16 && rows >= 0 && cols > 0, "softmax_rows_scaled: bad argument");
  BBBP_CHECK_ARG(ld_scores >= cols && ld_scores % 4 == 0 && ((uintptr_t)scores % 16) == 0,
                 "softmax_rows_scaled: ld_scores must be >= cols and a multiple of 4, base 16-byte aligned");
  BBBP_CHECK_ARG(ld_p >= cols && ld_p % 8 == 0, "softmax_rows_scaled: ld_p must be >= cols and a multiple of 8");
  BBBP_CHECK_ARG(rows <= 0x7fffffffLL, "softmax_rows_scaled: too many rows");
  if (rows == 0) return BBBP_OK;
  gemm::softmax_rows_scaled_bf16_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(
      scores, (size_t)ld_scores, static_cast<uint16_t*>(p_bf16), (size_t)ld_p, cols, scale, fmt);
  return launch_status("softmax_rows_scaled_bf16");
}
