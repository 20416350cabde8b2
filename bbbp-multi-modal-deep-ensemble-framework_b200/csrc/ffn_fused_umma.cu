// The feed-forward half of a post-norm nn.TransformerEncoderLayer (20250113.py:75-78) in ONE tcgen05 kernel:
//   y = LayerNorm(x + relu(x W1^T + b1) W2^T + b2) * gamma + beta
// As two library GEMMs + a LayerNorm kernel this block wrote the (rows x 2048) hidden activation to HBM and read it back,
// and both GEMMs were epilogue-bound (K = 167 for the first, N = 167 for the second: 32.5 + 21.3 + 8 us per 8 192 rows, ~7 %
// of the tensor peak).  Here the hidden activation never leaves the chip -- the same streaming structure as the attention
// kernel (attention_flash_umma.cu), with "keys" = hidden units:
//   S_j = X W1_j^T           tcgen05.mma  M=128 N=128 K=d      -> TMEM (double-buffered, 2 x 128 columns)
//   P_j = relu(S_j + b1_j)   4 activation warps, thread = row, tcgen05.ld 128 columns, 16-bit P into the 128B-swizzled
//                            K-major tile the next MMA reads as its A operand
//   O  += P_j W2_j^T         tcgen05.mma  M=128 N=ceil16(d) K=128 -> TMEM (columns 256..), B operand = a W2 column block (TMA)
// and the epilogue owns whole rows (d <= 192 fits one accumulator tile), so bias + residual + LayerNorm run straight out
// of TMEM: pass 1 forms s = O + b2 + x and stores it back to TMEM, pass 2 the variance around the mean, pass 3 normalises
// and writes the fp32 row (the next layer's residual) and its 16-bit copy (the next GEMM's A operand).
//
// Warp roles (192 threads): warp 0 TMA producer (X once; W1 blocks two ahead; W2 block one), warp 1 TMEM allocation + MMA
// issue (order S_0, S_1, O_0, S_2, O_1, ...), warps 2-5 activation + epilogue (TMEM lane quadrant = warp % 4).
#include "common.cuh"
#include "umma.cuh"
#include "half16.cuh"

namespace bbbp {
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t batches,
                      uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

namespace ffn {
using namespace sm100;

constexpr int BM = 128, BH = 128, KB = 64;             // rows per CTA, hidden units per block, columns per K-block tile
constexpr int THREADS = 192;
constexpr int TILE_B = 128 * 128;                       // one 128-row x 64-column 16-bit K-block tile: 16 KB
constexpr int MAX_DKB = 3, MAX_D = 176;                 // model width up to 176 (three K-blocks of 64, the last one partial)

struct Params {
  int rows, d, d_kb, dn, nb, fmt;                       // d_kb = ceil(d / 64), dn = ceil16(d), nb = hidden / 128
  const float *b1, *b2, *res, *gamma, *beta;
  int ld_res;
  float eps;
  float* y32;
  int ld_y;
  uint16_t* y16;
  int ld_y16;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive floats of a row vector / matrix row starting at column c0; columns >= d read as zero.  128-bit loads
// where a whole group of four lies below d (bases are 16-byte aligned and c0 is a multiple of 16).
__device__ __forceinline__ void load16(const float* __restrict__ src, int c0, int d, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = c0 + 4 * q;
    if (c + 4 <= d) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src + c));
      v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[4 * q + e] = (c + e) < d ? __ldg(src + c + e) : 0.0f;
    }
  }
}

// shared-memory plan (bytes from the 1024-aligned base):
//   X   d_kb tiles | W1  2 stages x d_kb tiles | W2  2 hidden K-blocks x (dn rows x 128 B) | P  2 hidden K-blocks x tile | barriers
//   | vectors: b2, gamma, beta (dn floats each, zero beyond d) and two blocks of b1 (double-buffered, one block ahead)
struct Smem {
  int x, w1, w2, p, bars, vec, total;
};
__host__ __device__ inline Smem plan(int d_kb, int dn) {
  Smem s;
  s.x = 0;
  s.w1 = s.x + d_kb * TILE_B;
  s.w2 = s.w1 + 2 * d_kb * TILE_B;
  s.p = s.w2 + ((2 * dn * 128 + 1023) / 1024) * 1024;
  s.bars = s.p + 2 * TILE_B;
  s.vec = s.bars + 16 * 8 + 16;
  s.total = s.vec + (3 * dn + 2 * BH) * 4 + 1024;
  return s;
}

__global__ void __launch_bounds__(THREADS, 1) ffn_layernorm_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmW1,
                                                                   const __grid_constant__ CUtensorMap tmW2,
                                                                   const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Smem sm = plan(p.d_kb, p.dn);
  uint8_t* sX = base + sm.x;
  uint8_t* sW1 = base + sm.w1;
  uint8_t* sW2 = base + sm.w2;
  uint8_t* sP = base + sm.p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + sm.bars);
  uint64_t* x_full = bars;             // 1
  uint64_t* w1_full = bars + 1;        // [2]
  uint64_t* w1_empty = bars + 3;       // [2]
  uint64_t* w2_full = bars + 5;        // 1
  uint64_t* w2_empty = bars + 6;       // 1
  uint64_t* s_full = bars + 7;         // [2]
  uint64_t* s_empty = bars + 9;        // [2]   (128 arrivals)
  // the 16-bit tile P is handed over in its two K-blocks (64 hidden units each): the second product of K-block 0 runs while
  // the activation warps still write K-block 1, and the next block's K-block 0 may be written as soon as ITS reader has finished
  uint64_t* p_full = bars + 11;        // [2]   (128 arrivals each)
  uint64_t* o_done = bars + 13;        // [2]   second product has read K-block kk of P
  uint64_t* o_full = bars + 15;        // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* sB2 = reinterpret_cast<float*>(base + sm.vec);      // every per-column vector the activation / epilogue threads need
  float* sGamma = sB2 + p.dn;                                //   lives in shared memory: read from global memory at their point
  float* sBeta = sGamma + p.dn;                              //   of use, each one cost an exposed L2 round trip per 16 columns
  float* sB1 = sBeta + p.dn;                                 //   (ncu: 2/3 of all stall samples of the first version)

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * BM;
  const int w1_stage = p.d_kb * TILE_B, w2_kb = p.dn * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&w1_full[s], 1);
      mbar_init(&w1_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
    }
    mbar_init(w2_full, 1);
    mbar_init(w2_empty, 1);
    for (int kk = 0; kk < 2; ++kk) {
      mbar_init(&p_full[kk], 128);
      mbar_init(&o_done[kk], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===== TMA producer =================================================================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, p.d_kb * TILE_B);
      for (int kb = 0; kb < p.d_kb; ++kb) tma_load_3d(&tmX, x_full, sX + kb * TILE_B, kb * KB, m0, 0);
      auto load_w1 = [&](int j) {
        const int s = j & 1;
        mbar_wait(&w1_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&w1_full[s], p.d_kb * TILE_B);
        for (int kb = 0; kb < p.d_kb; ++kb) tma_load_3d(&tmW1, &w1_full[s], sW1 + s * w1_stage + kb * TILE_B, kb * KB, j * BH, 0);
      };
      load_w1(0);
      for (int j = 0; j < p.nb; ++j) {
        if (j + 1 < p.nb) load_w1(j + 1);
        mbar_wait(w2_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(w2_full, 2 * w2_kb);
        for (int kk = 0; kk < 2; ++kk) tma_load_3d(&tmW2, w2_full, sW2 + kk * w2_kb, j * BH + kk * KB, 0, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer ===================================================================================================
    const uint32_t idesc_s = make_idesc_16(BM, BH, p.fmt);
    const uint32_t idesc_o = make_idesc_16(BM, (uint32_t)p.dn, p.fmt);
    const int last_steps = (p.d - (p.d_kb - 1) * KB + 15) / 16;       // K steps of 16 in the last K-block of the model width
    auto second = [&](int i) {
      mbar_wait(w2_full, i & 1);
      for (int kk = 0; kk < 2; ++kk) {
        mbar_wait(&p_full[kk], i & 1);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t a0 = smem_u32(sP), b0 = smem_u32(sW2);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem_o, make_smem_desc(a0 + kk * TILE_B + k4 * 32, 0, 1024, kLayoutSw128),
                      make_smem_desc(b0 + kk * w2_kb + k4 * 32, 0, 1024, kLayoutSw128), idesc_o, (i | kk | k4) != 0);
          umma_commit(&o_done[kk]);
          if (kk == 1) umma_commit(w2_empty);
        }
        __syncwarp();
      }
    };
    mbar_wait(x_full, 0);
    for (int j = 0; j < p.nb; ++j) {
      const int sb = j & 1;
      mbar_wait(&w1_full[sb], (j >> 1) & 1);
      mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t a0 = smem_u32(sX), b0 = smem_u32(sW1 + sb * w1_stage);
        for (int kb = 0; kb < p.d_kb; ++kb) {
          const int steps = kb == p.d_kb - 1 ? last_steps : 4;
          for (int k4 = 0; k4 < steps; ++k4)
            umma_bf16(tmem_base + sb * BH, make_smem_desc(a0 + kb * TILE_B + k4 * 32, 0, 1024, kLayoutSw128),
                      make_smem_desc(b0 + kb * TILE_B + k4 * 32, 0, 1024, kLayoutSw128), idesc_s, (kb | k4) != 0);
        }
        umma_commit(&w1_empty[sb]);
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
      if (j >= 1) second(j - 1);
    }
    second(p.nb - 1);
    if (elect_one_sync()) umma_commit(o_full);
    __syncwarp();
  } else {
    // ===== activation + epilogue: thread = row ==========================================================================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int swz = r & 7;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    // stage b2 / gamma / beta and the first block of b1 (coalesced loads, hidden behind the first TMA round trip)
    for (int c = r; c < p.dn; c += 128) {
      sB2[c] = c < p.d ? __ldg(p.b2 + c) : 0.0f;
      sGamma[c] = c < p.d ? __ldg(p.gamma + c) : 0.0f;
      sBeta[c] = c < p.d ? __ldg(p.beta + c) : 0.0f;
    }
    sB1[r] = __ldg(p.b1 + r);
    named_bar_sync(1, 128);
    for (int j = 0; j < p.nb; ++j) {
      const int sb = j & 1;
      const float b1_next = (j + 1 < p.nb) ? __ldg(p.b1 + (j + 1) * BH + r) : 0.0f;     // in flight during this block
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after_sync();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      const uint32_t ts = tmem_base + lane_off + sb * BH;
      tmem_ld_32x32(ts, s0);
      tmem_ld_32x32(ts + 32, s1);
      tmem_ld_32x32(ts + 64, s2);
      tmem_ld_32x32(ts + 96, s3);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&s_empty[sb]);                          // the pre-activations are in registers: S[sb] may be overwritten
      // P = relu(S + b1), 16-bit, into the swizzled A-operand tile: 16-byte chunk c of row r at position c ^ (r % 8)
      const uint32_t bias = smem_u32(sB1 + sb * BH);                            // the same for every thread: broadcast loads
#define BBBP_ACT_STORE(ARR, KK, CH0, COL0)                                                                 \
  _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                          \
    const float4 ba = ld_shared_f4(bias + ((COL0) + q * 8) * 4), bb = ld_shared_f4(bias + ((COL0) + q * 8 + 4) * 4); \
    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};                                  \
    uint32_t pk[4];                                                                                        \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
      const float a = fmaxf(__uint_as_float(ARR[q * 8 + 2 * e]) + bv[2 * e], 0.0f);                        \
      const float b = fmaxf(__uint_as_float(ARR[q * 8 + 2 * e + 1]) + bv[2 * e + 1], 0.0f);                \
      pk[e] = pack16_rt(a, b, p.fmt);                                                                      \
    }                                                                                                      \
    *reinterpret_cast<uint4*>(prow + (KK) * TILE_B + ((((CH0) + q) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]); \
  }
      if (j > 0) mbar_wait(&o_done[0], (j - 1) & 1);      // the previous block's second product has read K-block 0 of P
      BBBP_ACT_STORE(s0, 0, 0, 0)
      BBBP_ACT_STORE(s1, 0, 4, 32)
      fence_proxy_async_smem();                           // generic-proxy writes of P -> visible to tcgen05.mma
      tc_fence_before_sync();
      mbar_arrive(&p_full[0]);
      if (j > 0) mbar_wait(&o_done[1], (j - 1) & 1);
      BBBP_ACT_STORE(s2, 1, 0, 64)
      BBBP_ACT_STORE(s3, 1, 4, 96)
#undef BBBP_ACT_STORE
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&p_full[1]);
      sB1[(sb ^ 1) * BH + r] = b1_next;                   // (that buffer was last read in block j - 1: every thread is past it)
      named_bar_sync(1, 128);
    }
    // epilogue: s = O + b2 + residual; LayerNorm over the d true columns of the row this thread owns
    mbar_wait(o_full, 0);
    tc_fence_after_sync();
    const int row = m0 + r;
    const bool live = row < p.rows;
    const float* rrow = p.res + (size_t)(live ? row : 0) * p.ld_res;
    const uint32_t to = tmem_o + lane_off;
    float sum = 0.0f;
    const int dres = live ? p.d : 0;
    // the residual row comes from global memory (L2): the loads of the NEXT 32 columns are issued before the current 32 are
    // consumed, so the pass pays one L2 round trip instead of one per step
    float ra[16], rb[16], na[16], nb_[16];
    load16(rrow, 0, dres, ra);
    load16(rrow, 16, dres, rb);
    for (int c0 = 0; c0 < p.dn; c0 += 32) {
      load16(rrow, c0 + 32, dres, na);                    // (columns >= d read as zero without touching memory)
      load16(rrow, c0 + 48, dres, nb_);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = c0 + 16 * h;
        if (cc < p.dn) {
          uint32_t o[16];
          tmem_ld_32x16(to + cc, o);                      // .sync.aligned: every lane takes part, dead rows carry zeros
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = ld_shared_f4(smem_u32(sB2 + cc + 4 * q));
            const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * q + e;
              const float v = (cc + c) < p.d ? __uint_as_float(o[c]) + bv[e] + (h ? rb[c] : ra[c]) : 0.0f;
              sum += v;
              o[c] = __float_as_uint(v);
            }
          }
          tmem_st_32x16(to + cc, o);
        }
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) ra[c] = na[c], rb[c] = nb_[c];
    }
    tmem_st_wait();
    const float mean = sum / (float)p.d;
    float var = 0.0f;
    for (int c0 = 0; c0 < p.dn; c0 += 16) {
      uint32_t o[16];
      tmem_ld_32x16(to + c0, o);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float t = __uint_as_float(o[c]) - mean;
        var = (c0 + c) < p.d ? fmaf(t, t, var) : var;
      }
    }
    const float rstd = rsqrtf(var / (float)p.d + p.eps);
    float* yrow = p.y32 + (size_t)row * p.ld_y;
    uint16_t* hrow = p.y16 ? p.y16 + (size_t)row * p.ld_y16 : nullptr;
    for (int c0 = 0; c0 < p.dn; c0 += 16) {
      uint32_t o[16];
      float y[16];
      tmem_ld_32x16(to + c0, o);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 gq = ld_shared_f4(smem_u32(sGamma + c0 + 4 * q)), eq = ld_shared_f4(smem_u32(sBeta + c0 + 4 * q));
        const float gv[4] = {gq.x, gq.y, gq.z, gq.w}, ev[4] = {eq.x, eq.y, eq.z, eq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * q + e;
          y[c] = (c0 + c) < p.d ? (__uint_as_float(o[c]) - mean) * rstd * gv[e] + ev[e] : 0.0f;
        }
      }
      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + 4 * q + 4 <= p.ld_y)
            *reinterpret_cast<float4*>(yrow + c0 + 4 * q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        if (hrow) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (c0 + 8 * h + 8 <= p.ld_y16)
              *reinterpret_cast<uint4*>(hrow + c0 + 8 * h) =
                  make_uint4(pack16_rt(y[8 * h], y[8 * h + 1], p.fmt), pack16_rt(y[8 * h + 2], y[8 * h + 3], p.fmt),
                             pack16_rt(y[8 * h + 4], y[8 * h + 5], p.fmt), pack16_rt(y[8 * h + 6], y[8 * h + 7], p.fmt));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ffn
}  // namespace bbbp

extern "C" int bbbp_ffn_layernorm16(int fmt, int rows, int d, int hidden, const void* x16, int ldx, const void* w1_16, int ldw1,
                                    const float* b1, const void* w2_16, int ldw2, const float* b2, const float* residual,
                                    int ld_res, const float* gamma, const float* beta, float eps, float* y32, int ld_y, void* y16,
                                    int ld_y16, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "ffn_layernorm: bad fmt %d", fmt);
  BBBP_CHECK_ARG(x16 && w1_16 && b1 && w2_16 && b2 && residual && gamma && beta && y32 && rows >= 0, "ffn_layernorm: null operand");
  BBBP_CHECK_ARG(d >= 1 && d <= ffn::MAX_D, "ffn_layernorm: model width %d (1..%d)", d, ffn::MAX_D);
  BBBP_CHECK_ARG(hidden >= ffn::BH && hidden % ffn::BH == 0, "ffn_layernorm: hidden width %d must be a multiple of %d", hidden, ffn::BH);
  BBBP_CHECK_ARG(ldx % 8 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0 && ldx >= d && ldw1 >= d && ldw2 >= hidden,
                 "ffn_layernorm: 16-bit pitches must be multiples of 8 elements and cover their rows");
  BBBP_CHECK_ARG(ld_res >= d && ld_res % 4 == 0 && ld_y >= d && ld_y % 4 == 0 &&
                     (!y16 || (ld_y16 >= d && ld_y16 % 8 == 0 && ld_y16 <= ceil_div(d, 16) * 16)),
                 "ffn_layernorm: fp32 pitches must be multiples of 4 (16-bit output: 8, at most ceil16(d)) and cover the row");
  // (the bias / LayerNorm vectors are staged with scalar loads: parameters packed back to back need no alignment)
  BBBP_CHECK_ARG((((uintptr_t)x16 | (uintptr_t)w1_16 | (uintptr_t)w2_16 | (uintptr_t)residual | (uintptr_t)y32 | (uintptr_t)y16) % 16) == 0,
                 "ffn_layernorm: matrix operands must be 16-byte aligned");
  if (rows == 0) return BBBP_OK;
  ffn::Params p{};
  p.rows = rows, p.d = d, p.d_kb = ceil_div(d, ffn::KB), p.dn = ceil_div(d, 16) * 16, p.nb = hidden / ffn::BH, p.fmt = fmt;
  p.b1 = b1, p.b2 = b2, p.res = residual, p.gamma = gamma, p.beta = beta, p.ld_res = ld_res, p.eps = eps;
  p.y32 = y32, p.ld_y = ld_y, p.y16 = static_cast<uint16_t*>(y16), p.ld_y16 = ld_y16;
  CUtensorMap tmX, tmW1, tmW2;
  int st = make_tmap_bf16_3d(&tmX, x16, (uint64_t)rows, (uint64_t)d, (uint64_t)ldx, 1, (uint64_t)rows * ldx, ffn::BM, ffn::KB,
                             CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  st = make_tmap_bf16_3d(&tmW1, w1_16, (uint64_t)hidden, (uint64_t)d, (uint64_t)ldw1, 1, (uint64_t)hidden * ldw1, ffn::BH, ffn::KB,
                         CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  // W2 (d x hidden): a box of dn rows -- rows >= d are out of bounds and read as zero, so columns >= d of O stay zero
  st = make_tmap_bf16_3d(&tmW2, w2_16, (uint64_t)d, (uint64_t)hidden, (uint64_t)ldw2, 1, (uint64_t)d * ldw2, (uint32_t)p.dn, ffn::KB,
                         CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  const ffn::Smem sm = ffn::plan(p.d_kb, p.dn);
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    cudaFuncSetAttribute(ffn::ffn_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         ffn::plan(ffn::MAX_DKB, ffn::MAX_D).total);
  ffn::ffn_layernorm_kernel<<<ceil_div(rows, ffn::BM), ffn::THREADS, sm.total, as_stream(stream)>>>(tmX, tmW1, tmW2, p);
  return launch_status("ffn_layernorm16");
}
