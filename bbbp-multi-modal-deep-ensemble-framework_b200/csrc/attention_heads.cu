// Many small heads on the bf16 inference path: the 2048-bit fingerprint variants run nn.MultiheadAttention with 256
// heads of dimension 8 (20250113.py:71-73).  A tcgen05 tile (M = 128, K >= 16 per instruction, one CTA-wide issue) is the
// wrong shape for a 256 x 256 x 8 product per head, so this kernel uses warp-level mma.sync m16n8k16 (bf16 in, fp32
// accumulate): a warp owns 16 queries of one head and streams the keys in tiles of 64 with an online softmax -- the
// score fragments are re-used in registers as the A operand of the P V product (flash-attention dataflow), nothing
// but q, k, v (bf16, read in place from the packed in_proj output) and the bf16 result touches memory.
#include <math.h>
#include "common.cuh"
#include "half16.cuh"

namespace bbbp {

constexpr int AH_WARPS = 4, AH_KT = 64;      // 64 queries per CTA, 64 keys per tile

template <int FMT>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (FMT == BBBP_FMT_F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// qkv: bf16 rows of pitch ld with q at column 0, k at column koff, v at column voff; head h owns columns [h*D, (h+1)*D)
template <int D, int FMT>
__global__ void __launch_bounds__(AH_WARPS * 32) attention_heads_bf16_kernel(const uint16_t* __restrict__ qkv, int ld,
                                                                             int koff, int voff, uint16_t* __restrict__ out,
                                                                             int ld_out, int seq, float scale_log2e) {
  static_assert(D == 8 || D == 16, "head dimension 8 or 16");
  constexpr int ND = D / 8;                   // n-tiles of the output
  constexpr int VT_PITCH = AH_KT + 2;         // bf16 elements; 33 words: conflict-free fragment reads
  __shared__ __align__(16) uint16_t Ks[AH_KT * D];
  __shared__ __align__(16) uint16_t Vt[D * VT_PITCH];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int quad = lane / 4, tq = lane % 4;
  const int h = blockIdx.y, g = blockIdx.z;
  const size_t row0 = (size_t)g * seq;
  const int q0 = blockIdx.x * (AH_WARPS * 16) + warp * 16;
  const uint16_t* base = qkv + row0 * ld + h * D;

  // Q fragment (16 queries x 16): rows quad / quad + 8, columns 2*tq..+1 (and +8 when D == 16)
  uint32_t qa[4] = {0u, 0u, 0u, 0u};
  {
    const int r0 = q0 + quad, r1 = r0 + 8;
    if (r0 < seq) {
      qa[0] = *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + 2 * tq);
      if (D == 16) qa[2] = *reinterpret_cast<const uint32_t*>(base + (size_t)r0 * ld + 8 + 2 * tq);
    }
    if (r1 < seq) {
      qa[1] = *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + 2 * tq);
      if (D == 16) qa[3] = *reinterpret_cast<const uint32_t*>(base + (size_t)r1 * ld + 8 + 2 * tq);
    }
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  float o[ND][4];
#pragma unroll
  for (int n = 0; n < ND; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.0f;

  for (int k0 = 0; k0 < seq; k0 += AH_KT) {
    __syncthreads();                          // previous tile fully consumed
    for (int i = threadIdx.x; i < AH_KT * (D / 8); i += AH_WARPS * 32) {
      const int key = i / (D / 8), part = i % (D / 8);
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
      if (k0 + key < seq) {
        const uint16_t* src = base + (size_t)(k0 + key) * ld + 8 * part;
        kv = *reinterpret_cast<const uint4*>(src + koff);
        vv = *reinterpret_cast<const uint4*>(src + voff);
      }
      *reinterpret_cast<uint4*>(Ks + key * D + 8 * part) = kv;
      const uint16_t* ve = reinterpret_cast<const uint16_t*>(&vv);
#pragma unroll
      for (int e = 0; e < 8; ++e) Vt[(8 * part + e) * VT_PITCH + key] = ve[e];
    }
    __syncthreads();
    // S = Q K^T for the tile: 8 n-tiles of 8 keys
    float s[AH_KT / 8][4];
#pragma unroll
    for (int j = 0; j < AH_KT / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
      const uint16_t* kr = Ks + (j * 8 + quad) * D + 2 * tq;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
      const uint32_t b1 = D == 16 ? *reinterpret_cast<const uint32_t*>(kr + 8) : 0u;
      mma_16816<FMT>(s[j], qa, b0, b1);
    }
    // online softmax in the exp2 domain; thread holds rows quad (e = 0, 1) and quad + 8 (e = 2, 3), keys j*8 + 2*tq + {0, 1}
    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < AH_KT / 8; ++j) {
      const int key = k0 + j * 8 + 2 * tq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = (key + (e & 1) < seq) ? s[j][e] * scale_log2e : -INFINITY;
        s[j][e] = v;
        if (e < 2) t0 = fmaxf(t0, v); else t1 = fmaxf(t1, v);
      }
    }
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
    t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
    t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
    const float mn0 = fmaxf(m0, t0), mn1 = fmaxf(m1, t1);     // finite: every tile holds at least one valid key
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);   // m == -inf on the first tile -> 0
    m0 = mn0, m1 = mn1;
    l0 *= a0, l1 *= a1;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      o[n][0] *= a0, o[n][1] *= a0, o[n][2] *= a1, o[n][3] *= a1;
    }
#pragma unroll
    for (int j = 0; j < AH_KT / 8; ++j) {
      s[j][0] = exp2f(s[j][0] - m0), s[j][1] = exp2f(s[j][1] - m0);
      s[j][2] = exp2f(s[j][2] - m1), s[j][3] = exp2f(s[j][3] - m1);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    // O += P V: k-tile t = keys 16t .. 16t+15 = score n-tiles 2t, 2t+1 re-used as the A fragment
#pragma unroll
    for (int t = 0; t < AH_KT / 16; ++t) {
      const uint32_t pa[4] = {pack16<FMT>(s[2 * t][0], s[2 * t][1]), pack16<FMT>(s[2 * t][2], s[2 * t][3]),
                              pack16<FMT>(s[2 * t + 1][0], s[2 * t + 1][1]), pack16<FMT>(s[2 * t + 1][2], s[2 * t + 1][3])};
#pragma unroll
      for (int n = 0; n < ND; ++n) {
        const uint16_t* vr = Vt + (n * 8 + quad) * VT_PITCH + t * 16 + 2 * tq;
        mma_16816<FMT>(o[n], pa, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + quad, r1 = r0 + 8;
#pragma unroll
  for (int n = 0; n < ND; ++n) {
    if (r0 < seq)
      *reinterpret_cast<uint32_t*>(out + (row0 + r0) * ld_out + h * D + n * 8 + 2 * tq) = pack16<FMT>(o[n][0] * i0, o[n][1] * i0);
    if (r1 < seq)
      *reinterpret_cast<uint32_t*>(out + (row0 + r1) * ld_out + h * D + n * 8 + 2 * tq) = pack16<FMT>(o[n][2] * i1, o[n][3] * i1);
  }
}

}  // namespace bbbp

extern "C" int bbbp_attention_heads16(int fmt, const void* qkv, int ld, int k_offset, int v_offset, void* out, int ld_out,
                                      int groups, int seq, int heads, int head_dim, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "attention_heads: bad fmt %d", fmt);
  BBBP_CHECK_ARG(qkv && out && groups >= 0 && seq > 0 && heads > 0, "attention_heads: bad argument");
  BBBP_CHECK_ARG(head_dim == 8 || head_dim == 16, "attention_heads: head_dim %d (8 or 16 supported)", head_dim);
  BBBP_CHECK_ARG(ld % 8 == 0 && k_offset % 8 == 0 && v_offset % 8 == 0 && ld_out % 2 == 0,
                 "attention_heads: 16-byte aligned q/k/v rows required");
  BBBP_CHECK_ARG(heads <= 65535 && groups <= 65535, "attention_heads: heads/groups exceed 65535");
  if (groups == 0) return BBBP_OK;
  const float scale_log2e = rsqrtf((float)head_dim) * 1.4426950408889634f;
  const dim3 grid(ceil_div(seq, AH_WARPS * 16), heads, groups);
  auto q = static_cast<const uint16_t*>(qkv);
  auto o = static_cast<uint16_t*>(out);
  cudaStream_t s = as_stream(stream);
#define BBBP_AH(D, F) attention_heads_bf16_kernel<D, F><<<grid, AH_WARPS * 32, 0, s>>>(q, ld, k_offset, v_offset, o, ld_out, seq, scale_log2e)
  if (head_dim == 8) {
    if (fmt == BBBP_FMT_F16) BBBP_AH(8, BBBP_FMT_F16); else BBBP_AH(8, BBBP_FMT_BF16);
  } else {
    if (fmt == BBBP_FMT_F16) BBBP_AH(16, BBBP_FMT_F16); else BBBP_AH(16, BBBP_FMT_BF16);
  }
#undef BBBP_AH
  return launch_status("attention_heads16");
}

extern "C" int bbbp_attention_heads_bf16(const void* qkv_bf16, int ld, int k_offset, int v_offset, void* out_bf16, int ld_out,
                                         int groups, int seq, int heads, int head_dim, bbbp_stream_t stream) {
  return bbbp_attention_heads16(BBBP_FMT_BF16, qkv_bf16, ld, k_offset, v_offset, out_bf16, ld_out, groups, seq, heads, head_dim,
                                stream);
}
