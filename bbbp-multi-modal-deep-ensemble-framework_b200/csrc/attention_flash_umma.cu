// Streaming-softmax self-attention on tcgen05 / TMEM / TMA for ONE head of any width up to 192 and ANY scope length:
//   O = softmax(scale * Q K^T) V    per group (= reference batch; the encoder attends ACROSS the molecules of a batch,
//   nn.TransformerEncoder over the (B,1,F) tensor of 20250113.py:110-111 read as seq_len = B).
// The S x S logits never leave the chip: a CTA owns 128 queries and walks the keys in blocks of 128,
//   S_j = Q K_j^T            tcgen05.mma  M=128 N=128 K=d     -> TMEM (double-buffered, 2 x 128 columns)
//   P_j = exp2(S_j - m)       4 softmax warps, thread = query row, tcgen05.ld 128 columns into registers, 16-bit P written
//                             to shared memory in the 128B-swizzled K-major layout the next MMA reads as its A operand
//   O  += P_j V_j             tcgen05.mma  M=128 N=ceil16(d) K=128 -> TMEM (columns 256..), B operand = V^T tile (TMA)
// with a LAZY running maximum: the reference m only moves when a block's maximum exceeds it by more than 8 (in log2
// units, i.e. P may reach 256 -- exact in fp16/bf16 and far from overflow), so the O accumulator is rescaled in TMEM
// (tcgen05.ld / multiply / tcgen05.st) a handful of times per row instead of once per block.  l = sum P is kept in fp32
// registers; the epilogue writes O / l as 16-bit rows.
//
// Warp roles (192 threads): warp 0 TMA producer (Q once; K two blocks ahead; V^T one block), warp 1 TMEM allocation + MMA
// issue (order QK_0, QK_1, PV_0, QK_2, PV_1, ...: the next block's logits are computed while the softmax warps work),
// warps 2-5 softmax + correction + epilogue (TMEM lane quadrant = warp % 4).
//
// Fused tail (Params::fuse, bbbp_attention_flash_proj_ln16): the rest of the attention half of a post-norm encoder layer --
// out_proj, + bias, + residual, norm1 (nn.TransformerEncoderLayer via 20250113.py:75-78) -- runs in the same CTA.  After the
// last P V product the softmax threads write O / l as the 16-bit A operand into the (now dead) Q tiles, the producer fetches
// W_out into the (now dead) K stages, one more product (M=128, N=ceil16(d), K=d) lands in the (now dead) S columns of TMEM,
// and the epilogue normalises whole rows out of TMEM exactly as the feed-forward kernel does (ffn_fused_umma.cu).  The
// attention output never reaches HBM and three launches become one.
#include "common.cuh"
#include "umma.cuh"
#include "half16.cuh"

namespace bbbp {
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint64_t batches,
                      uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

namespace flash {
using namespace sm100;

constexpr int BM = 128, BN = 128, KB = 64;
constexpr int THREADS = 192;
constexpr int TILE_B = 128 * 128;                       // one 128-row x 64-column 16-bit K-block tile: 16 KB
constexpr int MAX_DKB = 3;                              // head width up to 192
constexpr int MAX_D_FUSED = 176;                        // ... up to 176 with the fused out_proj + LayerNorm tail (shared memory)
constexpr float RESCALE_THRESHOLD = 8.0f;               // log2 units

struct Params {
  int seq, d, d_kb, dn, nb;      // d_kb = ceil(d / 64), dn = ceil16(d) = columns of O, nb = key blocks
  int fmt;
  float scale_log2e;
  uint16_t* out;
  int ld_out;
  long long out_gs;              // group stride of out (elements)
  // fused tail (fuse != 0): y = LayerNorm(residual + (O / l) W_out^T + b_out) * gamma + beta, rows indexed g * seq + row
  int fuse;
  const float *b_out, *res, *gamma, *beta;
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// registers -> TMEM: 32 lanes x 32 / 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
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
// 16 consecutive floats of a row starting at column c0; columns >= d read as zero (128-bit loads where a whole group of four
// lies below d: the base is 16-byte aligned and c0 a multiple of 16)
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
//   Q   d_kb tiles | K  2 stages x d_kb tiles | V^T  2 key K-blocks x (dn rows x 128 B) | P  2 key K-blocks x tile | barriers
//   | (fused tail) b_out, gamma, beta: dn floats each
struct Smem {
  int q, k, v, p, bars, vec, total;
};
__host__ __device__ inline Smem plan(int d_kb, int dn, int fuse = 0) {
  Smem s;
  s.q = 0;
  s.k = s.q + d_kb * TILE_B;
  s.v = s.k + 2 * d_kb * TILE_B;
  s.p = s.v + ((2 * dn * 128 + 1023) / 1024) * 1024;
  s.bars = s.p + 2 * TILE_B;
  s.vec = s.bars + 20 * 8 + 16;
  s.total = s.vec + (fuse ? 3 * dn * 4 : 0) + 1024;
  return s;
}

__global__ void __launch_bounds__(THREADS, 1) attention_flash_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Smem sm = plan(p.d_kb, p.dn, p.fuse);
  uint8_t* sQ = base + sm.q;
  uint8_t* sK = base + sm.k;
  uint8_t* sV = base + sm.v;
  uint8_t* sP = base + sm.p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + sm.bars);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // 1
  uint64_t* v_empty = bars + 6;       // 1
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* s_empty = bars + 9;       // [2]   (128 arrivals)
  uint64_t* p_full = bars + 11;       // 1     (128 arrivals)
  uint64_t* pv_done = bars + 12;      // 1     P buffer free AND O stable
  uint64_t* o_full = bars + 13;       // 1
  uint64_t* w_full = bars + 14;       // 1     fused tail: W_out has landed in the K stages
  uint64_t* a_full = bars + 15;       // 1     fused tail: O / l written as the A operand (128 arrivals)
  uint64_t* y_full = bars + 16;       // 1     fused tail: projection complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  [[maybe_unused]] float* sBo = reinterpret_cast<float*>(base + sm.vec);
  [[maybe_unused]] float* sGamma = sBo + p.dn;
  [[maybe_unused]] float* sBeta = sGamma + p.dn;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = blockIdx.y, m0 = blockIdx.x * BM;
  const int k_stage = p.d_kb * TILE_B, v_kb = p.dn * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 128);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    mbar_init(w_full, 1);
    mbar_init(a_full, 128);
    mbar_init(y_full, 1);
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
      mbar_arrive_expect_tx(q_full, p.d_kb * TILE_B);
      for (int kb = 0; kb < p.d_kb; ++kb) tma_load_3d(&tmQ, q_full, sQ + kb * TILE_B, kb * KB, m0, g);
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], p.d_kb * TILE_B);
        for (int kb = 0; kb < p.d_kb; ++kb) tma_load_3d(&tmK, &k_full[s], sK + s * k_stage + kb * TILE_B, kb * KB, j * BN, g);
      };
      load_k(0);
      for (int j = 0; j < p.nb; ++j) {
        if (j + 1 < p.nb) load_k(j + 1);                // keys stay two blocks ahead of the values
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, 2 * v_kb);
        for (int kk = 0; kk < 2; ++kk) tma_load_3d(&tmV, v_full, sV + kk * v_kb, j * BN + kk * KB, 0, g);
      }
      if (p.fuse) {
        // both K stages are dead once the logits of the last two key blocks have been computed: W_out (dn rows x d) goes there
        for (int j = p.nb; j < p.nb + 2; ++j) mbar_wait(&k_empty[j & 1], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(w_full, p.d_kb * v_kb);
        for (int kb = 0; kb < p.d_kb; ++kb) tma_load_3d(&tmW, w_full, sK + kb * v_kb, kb * KB, 0, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer ===================================================================================================
    const uint32_t idesc_s = make_idesc_16(BM, BN, p.fmt);
    const uint32_t idesc_o = make_idesc_16(BM, (uint32_t)p.dn, p.fmt);
    const int last_steps = (p.d - (p.d_kb - 1) * KB + 15) / 16;       // K steps of 16 in the last K-block of the head width
    auto do_pv = [&](int i) {
      mbar_wait(p_full, i & 1);
      mbar_wait(v_full, i & 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t a0 = smem_u32(sP), b0 = smem_u32(sV);
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem_o, make_smem_desc(a0 + kk * TILE_B + k4 * 32, 0, 1024, kLayoutSw128),
                      make_smem_desc(b0 + kk * v_kb + k4 * 32, 0, 1024, kLayoutSw128), idesc_o, (i | kk | k4) != 0);
        umma_commit(v_empty);
        umma_commit(pv_done);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < p.nb; ++j) {
      const int sb = j & 1;
      mbar_wait(&k_full[sb], (j >> 1) & 1);
      mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK + sb * k_stage);
        for (int kb = 0; kb < p.d_kb; ++kb) {
          const int steps = kb == p.d_kb - 1 ? last_steps : 4;
          for (int k4 = 0; k4 < steps; ++k4)
            umma_bf16(tmem_base + sb * BN, make_smem_desc(a0 + kb * TILE_B + k4 * 32, 0, 1024, kLayoutSw128),
                      make_smem_desc(b0 + kb * TILE_B + k4 * 32, 0, 1024, kLayoutSw128), idesc_s, (kb | k4) != 0);
        }
        umma_commit(&k_empty[sb]);
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
      if (j >= 1) do_pv(j - 1);
    }
    do_pv(p.nb - 1);
    if (elect_one_sync()) umma_commit(o_full);
    __syncwarp();
    if (p.fuse) {
      mbar_wait(w_full, 0);
      mbar_wait(a_full, 0);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK);
        for (int kb = 0; kb < p.d_kb; ++kb) {
          const int steps = kb == p.d_kb - 1 ? last_steps : 4;
          for (int k4 = 0; k4 < steps; ++k4)
            umma_bf16(tmem_base, make_smem_desc(a0 + kb * TILE_B + k4 * 32, 0, 1024, kLayoutSw128),
                      make_smem_desc(b0 + kb * v_kb + k4 * 32, 0, 1024, kLayoutSw128), idesc_o, (kb | k4) != 0);
        }
        umma_commit(y_full);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / correction / epilogue: thread = query row ===========================================================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_used = -INFINITY, l = 0.0f;
    const int swz = r & 7;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    for (int j = 0; j < p.nb; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after_sync();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      const uint32_t ts = tmem_base + lane_off + sb * BN;
      tmem_ld_32x32(ts, s0);
      tmem_ld_32x32(ts + 32, s1);
      tmem_ld_32x32(ts + 64, s2);
      tmem_ld_32x32(ts + 96, s3);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&s_empty[sb]);                          // the logits are in registers: the MMA warp may overwrite S[sb]
      const int valid = p.seq - j * BN;                   // keys of this block that exist (>= 1)
      float bmax = -INFINITY;
#define BBBP_SCALE_MASK(ARR, OFF)                                                        \
  _Pragma("unroll") for (int c = 0; c < 32; ++c) {                                       \
    float t = __uint_as_float(ARR[c]) * p.scale_log2e;                                   \
    t = (OFF + c) < valid ? t : -INFINITY;                                               \
    ARR[c] = __float_as_uint(t);                                                         \
    bmax = fmaxf(bmax, t);                                                               \
  }
      BBBP_SCALE_MASK(s0, 0)
      BBBP_SCALE_MASK(s1, 32)
      BBBP_SCALE_MASK(s2, 64)
      BBBP_SCALE_MASK(s3, 96)
#undef BBBP_SCALE_MASK
      const bool need = bmax > m_used + RESCALE_THRESHOLD;         // always true for j == 0 (m_used = -inf)
      const bool any = __any_sync(0xffffffffu, need);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                           // P V of the previous block: P buffer free, O stable
        if (any) {
          // rescale this warp's 32 rows of O by alpha = 2^(m_old - m_new) (1 for the rows whose reference does not move)
          const float alpha = need ? ex2(m_used - bmax) : 1.0f;
          l *= alpha;
          tc_fence_after_sync();
          const uint32_t to = tmem_o + lane_off;
          for (int c0 = 0; c0 + 32 <= p.dn; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32(to + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x32(to + c0, o);
          }
          if (p.dn % 32) {
            uint32_t o[16];
            const int c0 = p.dn & ~31;
            tmem_ld_32x16(to + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x16(to + c0, o);
          }
          tmem_st_wait();
        }
      }
      if (need) m_used = bmax;
      // P = 2^(t - m_used), 16-bit, into the swizzled A-operand tile: 16-byte chunk c of row r at position c ^ (r % 8)
      float lsum = 0.0f;
#define BBBP_EXP_STORE(ARR, KK, CH0)                                                                       \
  _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                          \
    uint32_t pk[4];                                                                                        \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
      const float a = ex2(__uint_as_float(ARR[q * 8 + 2 * e]) - m_used);                                   \
      const float b = ex2(__uint_as_float(ARR[q * 8 + 2 * e + 1]) - m_used);                               \
      lsum += a + b;                                                                                       \
      pk[e] = pack16_rt(a, b, p.fmt);                                                                      \
    }                                                                                                      \
    *reinterpret_cast<uint4*>(prow + (KK) * TILE_B + ((((CH0) + q) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]); \
  }
      BBBP_EXP_STORE(s0, 0, 0)
      BBBP_EXP_STORE(s1, 0, 4)
      BBBP_EXP_STORE(s2, 1, 0)
      BBBP_EXP_STORE(s3, 1, 4)
#undef BBBP_EXP_STORE
      l += lsum;
      fence_proxy_async_smem();                           // generic-proxy writes of P -> visible to tcgen05.mma
      tc_fence_before_sync();                             // and the TMEM stores of the correction, ordered before the arrive
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> 16-bit rows
    mbar_wait(o_full, 0);
    tc_fence_after_sync();
    const float inv = 1.0f / l;
    const int row = m0 + r;
    if (p.fuse) {
      // ---- fused tail: O / l as the A operand of the projection (the Q tiles are dead: every product that read them has
      //      completed, o_full tracks them all), then LayerNorm(residual + proj + b_out) out of TMEM
      for (int c = r; c < p.dn; c += 128) {
        sBo[c] = c < p.d ? __ldg(p.b_out + c) : 0.0f;
        sGamma[c] = c < p.d ? __ldg(p.gamma + c) : 0.0f;
        sBeta[c] = c < p.d ? __ldg(p.beta + c) : 0.0f;
      }
      const uint32_t to = tmem_o + lane_off;
      uint8_t* arow = sQ + (r >> 3) * 1024 + (r & 7) * 128;
      for (int c0 = 0; c0 < p.dn; c0 += 16) {
        uint32_t o[16];
        tmem_ld_32x16(to + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = pack16_rt(__uint_as_float(o[h * 8 + 2 * e]) * inv, __uint_as_float(o[h * 8 + 2 * e + 1]) * inv, p.fmt);
          const int col = c0 + 8 * h;
          *reinterpret_cast<uint4*>(arow + (col >> 6) * TILE_B + ((((col & 63) >> 3) ^ swz) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(a_full);
      named_bar_sync(1, 128);                             // the staged vectors are visible to all four warps
      const size_t grow = (size_t)g * p.seq + row;
      const bool live = row < p.seq;
      const int dres = live ? p.d : 0;
      const float* rrow = p.res + (live ? grow : 0) * p.ld_res;
      float ra[16], rb[16], na[16], nb_[16];
      load16(rrow, 0, dres, ra);                          // (in flight while the projection runs)
      load16(rrow, 16, dres, rb);
      mbar_wait(y_full, 0);
      tc_fence_after_sync();
      const uint32_t ty = tmem_base + lane_off;           // the projection landed in the S columns
      float sum = 0.0f;
      for (int c0 = 0; c0 < p.dn; c0 += 32) {
        load16(rrow, c0 + 32, dres, na);
        load16(rrow, c0 + 48, dres, nb_);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = c0 + 16 * h;
          if (cc < p.dn) {
            uint32_t o[16];
            tmem_ld_32x16(ty + cc, o);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bq = ld_shared_f4(smem_u32(sBo + cc + 4 * q));
              const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = 4 * q + e;
                const float v = (cc + c) < p.d ? __uint_as_float(o[c]) + bv[e] + (h ? rb[c] : ra[c]) : 0.0f;
                sum += v;
                o[c] = __float_as_uint(v);
              }
            }
            tmem_st_32x16(ty + cc, o);
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
        tmem_ld_32x16(ty + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float t = __uint_as_float(o[c]) - mean;
          var = (c0 + c) < p.d ? fmaf(t, t, var) : var;
        }
      }
      const float rstd = rsqrtf(var / (float)p.d + p.eps);
      float* yrow = p.y32 + grow * p.ld_y;
      uint16_t* hrow = p.y16 ? p.y16 + grow * p.ld_y16 : nullptr;
      for (int c0 = 0; c0 < p.dn; c0 += 16) {
        uint32_t o[16];
        float y[16];
        tmem_ld_32x16(ty + c0, o);
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
    } else {
    uint16_t* orow = p.out + (size_t)g * p.out_gs + (size_t)row * p.ld_out;
    const uint32_t to = tmem_o + lane_off;
    for (int c0 = 0; c0 < p.dn; c0 += 16) {
      uint32_t o[16];
      tmem_ld_32x16(to + c0, o);                          // .sync.aligned: every lane loads, only valid rows store
      tmem_ld_wait();
      if (row < p.seq) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c0 + h * 8 < p.ld_out) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              pk[e] = pack16_rt(__uint_as_float(o[h * 8 + 2 * e]) * inv, __uint_as_float(o[h * 8 + 2 * e + 1]) * inv, p.fmt);
            *reinterpret_cast<uint4*>(orow + c0 + h * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
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

}  // namespace flash
}  // namespace bbbp

namespace bbbp {
namespace flash {
// shared host side of the two entry points (fused: the tail arguments are set in p, w_out / ldw give the projection weight)
static int launch(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk, long long group_stride,
                  const void* v_t, int ld_vt, long long vt_group_stride, float scale, Params p, const void* w_out, int ldw,
                  cudaStream_t stream) {
  p.seq = seq, p.d = head_dim, p.d_kb = ceil_div(head_dim, KB), p.dn = ceil_div(head_dim, 16) * 16;
  p.nb = ceil_div(seq, BN);
  p.fmt = fmt;
  p.scale_log2e = scale * 1.4426950408889634f;
  CUtensorMap tmQ, tmK, tmV, tmW;
  int st = make_tmap_bf16_3d(&tmQ, q, (uint64_t)seq, (uint64_t)head_dim, (uint64_t)ldq, (uint64_t)groups, (uint64_t)group_stride,
                             BM, KB, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  st = make_tmap_bf16_3d(&tmK, k, (uint64_t)seq, (uint64_t)head_dim, (uint64_t)ldk, (uint64_t)groups, (uint64_t)group_stride, BN, KB,
                         CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  // V^T: rows = head dimension (box of dn rows: rows >= head_dim are out of bounds -> zero), columns = keys
  st = make_tmap_bf16_3d(&tmV, v_t, (uint64_t)head_dim, (uint64_t)seq, (uint64_t)ld_vt, (uint64_t)groups, (uint64_t)vt_group_stride,
                         (uint32_t)p.dn, KB, CU_TENSOR_MAP_SWIZZLE_128B);
  if (st != BBBP_OK) return st;
  tmW = tmQ;
  if (p.fuse) {     // W_out (d x d, rows = output features): a box of dn rows x 64 columns per K block
    st = make_tmap_bf16_3d(&tmW, w_out, (uint64_t)head_dim, (uint64_t)head_dim, (uint64_t)ldw, 1, 0, (uint32_t)p.dn, KB,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    if (st != BBBP_OK) return st;
  }
  const Smem sm = plan(p.d_kb, p.dn, p.fuse);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    const int a = plan(MAX_DKB, 64 * MAX_DKB).total, b = plan(MAX_DKB, MAX_D_FUSED, 1).total;
    cudaFuncSetAttribute(attention_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a > b ? a : b);
  }
  dim3 grid(ceil_div(seq, BM), groups);
  attention_flash_kernel<<<grid, THREADS, sm.total, stream>>>(tmQ, tmK, tmV, tmW, p);
  return launch_status(p.fuse ? "attention_flash_proj_ln16" : "attention_flash16");
}
}  // namespace flash
}  // namespace bbbp

static int flash_check(const char* who, int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                       long long group_stride, const void* v_t, int ld_vt, long long vt_group_stride) {
  using namespace bbbp;
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "%s: bad fmt %d", who, fmt);
  BBBP_CHECK_ARG(q && k && v_t && groups >= 0 && groups <= 65535 && seq > 0, "%s: bad argument", who);
  BBBP_CHECK_ARG(head_dim >= 1 && head_dim <= 64 * flash::MAX_DKB, "%s: head_dim %d (1..%d)", who, head_dim, 64 * flash::MAX_DKB);
  BBBP_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ld_vt % 8 == 0 && ld_vt >= seq && group_stride % 8 == 0 && vt_group_stride % 8 == 0,
                 "%s: pitches and group strides must be multiples of 8 elements (ld_vt >= seq)", who);
  BBBP_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v_t % 16) == 0, "%s: operands must be 16-byte aligned", who);
  return BBBP_OK;
}

extern "C" int bbbp_attention_flash16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                                      long long group_stride, const void* v_t, int ld_vt, long long vt_group_stride, float scale,
                                      void* out, int ld_out, long long out_group_stride, bbbp_stream_t stream) {
  using namespace bbbp;
  if (int rc = flash_check("attention_flash", fmt, groups, seq, head_dim, q, ldq, k, ldk, group_stride, v_t, ld_vt, vt_group_stride)) return rc;
  BBBP_CHECK_ARG(out && ld_out % 8 == 0 && out_group_stride % 8 == 0 && ((uintptr_t)out % 16) == 0 && ld_out >= head_dim,
                 "attention_flash: out must be 16-byte aligned, ld_out a multiple of 8 >= head_dim");
  if (groups == 0) return BBBP_OK;
  flash::Params p{};
  p.out = static_cast<uint16_t*>(out), p.ld_out = ld_out, p.out_gs = out_group_stride;
  return flash::launch(fmt, groups, seq, head_dim, q, ldq, k, ldk, group_stride, v_t, ld_vt, vt_group_stride, scale, p, nullptr, 0,
                       as_stream(stream));
}

extern "C" int bbbp_attention_flash_proj_ln16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                                              long long group_stride, const void* v_t, int ld_vt, long long vt_group_stride,
                                              float scale, const void* w_out16, int ldw, const float* b_out, const float* residual,
                                              int ld_res, const float* gamma, const float* beta, float eps, float* y32, int ld_y,
                                              void* y16, int ld_y16, bbbp_stream_t stream) {
  using namespace bbbp;
  if (int rc = flash_check("attention_flash_proj_ln", fmt, groups, seq, head_dim, q, ldq, k, ldk, group_stride, v_t, ld_vt, vt_group_stride)) return rc;
  BBBP_CHECK_ARG(head_dim <= flash::MAX_D_FUSED, "attention_flash_proj_ln: width %d (1..%d)", head_dim, flash::MAX_D_FUSED);
  BBBP_CHECK_ARG(w_out16 && b_out && residual && gamma && beta && y32, "attention_flash_proj_ln: null operand");
  BBBP_CHECK_ARG(ldw % 8 == 0 && ldw >= head_dim && ((uintptr_t)w_out16 % 16) == 0, "attention_flash_proj_ln: W_out pitch / alignment");
  BBBP_CHECK_ARG(ld_res >= head_dim && ld_res % 4 == 0 && ld_y >= head_dim && ld_y % 4 == 0 &&
                     (!y16 || (ld_y16 >= head_dim && ld_y16 % 8 == 0 && ld_y16 <= ceil_div(head_dim, 16) * 16)),
                 "attention_flash_proj_ln: fp32 pitches must be multiples of 4 (16-bit output: 8, at most ceil16(d)) and cover the row");
  BBBP_CHECK_ARG((((uintptr_t)residual | (uintptr_t)y32 | (uintptr_t)y16) % 16) == 0, "attention_flash_proj_ln: residual / outputs must be 16-byte aligned");
  if (groups == 0) return BBBP_OK;
  flash::Params p{};
  p.fuse = 1, p.b_out = b_out, p.res = residual, p.ld_res = ld_res, p.gamma = gamma, p.beta = beta, p.eps = eps;
  p.y32 = y32, p.ld_y = ld_y, p.y16 = static_cast<uint16_t*>(y16), p.ld_y16 = ld_y16;
  return flash::launch(fmt, groups, seq, head_dim, q, ldq, k, ldk, group_stride, v_t, ld_vt, vt_group_stride, scale, p, w_out16, ldw,
                       as_stream(stream));
}
