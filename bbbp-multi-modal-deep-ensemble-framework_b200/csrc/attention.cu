// Cross-molecule self-attention of the fingerprint encoder: nn.TransformerEncoder is built
// with batch_first=False at 20250113.py:75-78 and fed (B,1,F) at :110-111, so the sequence axis IS the
// reference mini-batch.  One launch covers `groups` independent reference batches of `seq` molecules.
// Streaming (online) softmax, fp32 CUDA cores: S x S scores are never materialised.
#include <math.h>
#include "common.cuh"

namespace bbbp {

constexpr int KT = 32;      // keys (or queries) per shared-memory tile
constexpr int MAX_DT = 8;   // head_dim <= 256
constexpr int ATT_WARPS = 4;

// Keep mask of attention-probability dropout: one Philox-4x32-10 block per (row, key, head) triple.
__device__ __forceinline__ uint32_t philox_first_word(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr.x;
}
// multiplier applied to probability (query row, key row) of head h: 0 or 1/(1-p); 1 when p == 0
__device__ __forceinline__ float keep_scale(float p, float inv_keep, uint64_t seed, size_t qrow, size_t krow, int h) {
  if (p <= 0.0f) return 1.0f;
  uint32_t r = philox_first_word(make_uint4((uint32_t)qrow, (uint32_t)krow, (uint32_t)h, (uint32_t)(qrow >> 32)),
                                 make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return (r * 2.3283064365386963e-10f) >= p ? inv_keep : 0.0f;
}

// dynamic smem layout (floats): Ks[KT][ld] | Vs[KT][ld] | qs[ATT_WARPS][ld]
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_fwd_kernel(const float* __restrict__ qkv,
                                                                       float* __restrict__ out, float* __restrict__ lse,
                                                                       int seq, int heads, int d, int q_per_block,
                                                                       float drop_p, uint64_t seed,
                                                                       const uint64_t* __restrict__ seed_dev) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  extern __shared__ float smem[];
  const int ld = d | 1;  // odd pitch: conflict-free row-per-lane reads
  float* Ks = smem;
  float* Vs = Ks + KT * ld;
  float* qs = Vs + KT * ld;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int h = blockIdx.y, g = blockIdx.z;
  const int E = heads * d, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  const float scale = rsqrtf((float)d);
  const float inv_keep = 1.0f / (1.0f - drop_p);
  const int qbase = blockIdx.x * q_per_block;
  const int nq = min(q_per_block, seq - qbase);
  const int q_iters = ceil_div(nq, ATT_WARPS);

  for (int it = 0; it < q_iters; ++it) {
    const int qi = qbase + it * ATT_WARPS + warp;
    const bool active = qi < qbase + nq;
    if (active)
      for (int dd = lane; dd < d; dd += 32) qs[warp * ld + dd] = qkv[(row0 + qi) * ldq + h * d + dd] * scale;
    float m = -INFINITY, l = 0.0f, o[MAX_DT];
#pragma unroll
    for (int t = 0; t < MAX_DT; ++t) o[t] = 0.0f;

    for (int k0 = 0; k0 < seq; k0 += KT) {
      __syncthreads();  // previous tile fully consumed (and qs visible)
      for (int i = threadIdx.x; i < KT * d; i += ATT_WARPS * 32) {
        int j = i / d, dd = i % d;
        bool ok = k0 + j < seq;
        const float* src = qkv + (row0 + k0 + j) * ldq + h * d + dd;
        Ks[j * ld + dd] = ok ? src[E] : 0.0f;
        Vs[j * ld + dd] = ok ? src[2 * E] : 0.0f;
      }
      __syncthreads();
      if (!active) continue;
      float s = 0.0f;
      for (int dd = 0; dd < d; ++dd) s = fmaf(qs[warp * ld + dd], Ks[lane * ld + dd], s);
      if (k0 + lane >= seq) s = -INFINITY;
      const float m_new = fmaxf(m, warp_max(s));
      const float alpha = __expf(m - m_new);  // m == -inf on the first tile -> 0
      const float p = __expf(s - m_new);
      l = l * alpha + warp_sum(p);
      m = m_new;
      // dropout acts on the normalised probabilities: the normaliser l keeps every key
      const float pd = p * keep_scale(drop_p, inv_keep, seed, row0 + qi, row0 + k0 + lane, h);
#pragma unroll
      for (int t = 0; t < MAX_DT; ++t) o[t] *= alpha;
      for (int j = 0; j < KT; ++j) {
        const float pj = __shfl_sync(0xffffffffu, pd, j);
#pragma unroll
        for (int t = 0; t < MAX_DT; ++t) {
          int dd = lane + 32 * t;
          if (dd < d) o[t] = fmaf(pj, Vs[j * ld + dd], o[t]);
        }
      }
    }
    if (active) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int t = 0; t < MAX_DT; ++t) {
        int dd = lane + 32 * t;
        if (dd < d) out[(row0 + qi) * E + h * d + dd] = o[t] * inv;
      }
      if (lse && lane == 0) lse[(row0 + qi) * heads + h] = m + __logf(l);
    }
  }
}

// dQ: one warp per query row, streaming over key tiles.
// smem: Ks | Vs | qs[W][ld] | dos[W][ld]
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_bwd_dq_kernel(const float* __restrict__ qkv,
                                                                          const float* __restrict__ out,
                                                                          const float* __restrict__ lse,
                                                                          const float* __restrict__ dout,
                                                                          float* __restrict__ dqkv, int seq, int heads,
                                                                          int d, int q_per_block, float drop_p,
                                                                          uint64_t seed,
                                                                          const uint64_t* __restrict__ seed_dev) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  const float inv_keep = 1.0f / (1.0f - drop_p);
  extern __shared__ float smem[];
  const int ld = d | 1;
  float* Ks = smem;
  float* Vs = Ks + KT * ld;
  float* qs = Vs + KT * ld;
  float* dos = qs + ATT_WARPS * ld;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int h = blockIdx.y, g = blockIdx.z;
  const int E = heads * d, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  const float scale = rsqrtf((float)d);
  const int qbase = blockIdx.x * q_per_block;
  const int nq = min(q_per_block, seq - qbase);
  const int q_iters = ceil_div(nq, ATT_WARPS);

  for (int it = 0; it < q_iters; ++it) {
    const int qi = qbase + it * ATT_WARPS + warp;
    const bool active = qi < qbase + nq;
    float D = 0.0f, L = 0.0f;
    if (active) {
      for (int dd = lane; dd < d; dd += 32) {
        qs[warp * ld + dd] = qkv[(row0 + qi) * ldq + h * d + dd] * scale;
        float go = dout[(row0 + qi) * E + h * d + dd];
        dos[warp * ld + dd] = go;
        D = fmaf(go, out[(row0 + qi) * E + h * d + dd], D);
      }
      D = warp_sum(D);
      L = lse[(row0 + qi) * heads + h];
    }
    float dq[MAX_DT];
#pragma unroll
    for (int t = 0; t < MAX_DT; ++t) dq[t] = 0.0f;
    for (int k0 = 0; k0 < seq; k0 += KT) {
      __syncthreads();
      for (int i = threadIdx.x; i < KT * d; i += ATT_WARPS * 32) {
        int j = i / d, dd = i % d;
        bool ok = k0 + j < seq;
        const float* src = qkv + (row0 + k0 + j) * ldq + h * d + dd;
        Ks[j * ld + dd] = ok ? src[E] : 0.0f;
        Vs[j * ld + dd] = ok ? src[2 * E] : 0.0f;
      }
      __syncthreads();
      if (!active) continue;
      float s = 0.0f, dp = 0.0f;
      for (int dd = 0; dd < d; ++dd) {
        s = fmaf(qs[warp * ld + dd], Ks[lane * ld + dd], s);
        dp = fmaf(dos[warp * ld + dd], Vs[lane * ld + dd], dp);
      }
      const float p = (k0 + lane < seq) ? __expf(s - L) : 0.0f;
      const float ds = p * (dp * keep_scale(drop_p, inv_keep, seed, row0 + qi, row0 + k0 + lane, h) - D);
      for (int j = 0; j < KT; ++j) {
        const float dsj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
        for (int t = 0; t < MAX_DT; ++t) {
          int dd = lane + 32 * t;
          if (dd < d) dq[t] = fmaf(dsj, Ks[j * ld + dd], dq[t]);
        }
      }
    }
    if (active) {
#pragma unroll
      for (int t = 0; t < MAX_DT; ++t) {
        int dd = lane + 32 * t;
        if (dd < d) dqkv[(row0 + qi) * ldq + h * d + dd] = dq[t] * scale;
      }
    }
  }
}

// dK, dV: one warp per key row, streaming over query tiles.
// smem: Qs[KT][ld] (pre-scaled) | dOs[KT][ld] | ks[W][ld] | vs[W][ld] | Ls[KT] | Ds[KT]
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_bwd_dkv_kernel(const float* __restrict__ qkv,
                                                                           const float* __restrict__ out,
                                                                           const float* __restrict__ lse,
                                                                           const float* __restrict__ dout,
                                                                           float* __restrict__ dqkv, int seq, int heads,
                                                                           int d, int k_per_block, float drop_p,
                                                                           uint64_t seed,
                                                                           const uint64_t* __restrict__ seed_dev) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  const float inv_keep = 1.0f / (1.0f - drop_p);
  extern __shared__ float smem[];
  const int ld = d | 1;
  float* Qs = smem;
  float* dOs = Qs + KT * ld;
  float* ks = dOs + KT * ld;
  float* vs = ks + ATT_WARPS * ld;
  float* Ls = vs + ATT_WARPS * ld;
  float* Ds = Ls + KT;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int h = blockIdx.y, g = blockIdx.z;
  const int E = heads * d, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  const float scale = rsqrtf((float)d);
  const int kbase = blockIdx.x * k_per_block;
  const int nk = min(k_per_block, seq - kbase);
  const int k_iters = ceil_div(nk, ATT_WARPS);

  for (int it = 0; it < k_iters; ++it) {
    const int kj = kbase + it * ATT_WARPS + warp;
    const bool active = kj < kbase + nk;
    if (active)
      for (int dd = lane; dd < d; dd += 32) {
        ks[warp * ld + dd] = qkv[(row0 + kj) * ldq + E + h * d + dd];
        vs[warp * ld + dd] = qkv[(row0 + kj) * ldq + 2 * E + h * d + dd];
      }
    float dk[MAX_DT], dv[MAX_DT];
#pragma unroll
    for (int t = 0; t < MAX_DT; ++t) dk[t] = dv[t] = 0.0f;
    for (int q0 = 0; q0 < seq; q0 += KT) {
      __syncthreads();
      for (int i = threadIdx.x; i < KT * d; i += ATT_WARPS * 32) {
        int j = i / d, dd = i % d;
        bool ok = q0 + j < seq;
        Qs[j * ld + dd] = ok ? qkv[(row0 + q0 + j) * ldq + h * d + dd] * scale : 0.0f;
        dOs[j * ld + dd] = ok ? dout[(row0 + q0 + j) * E + h * d + dd] : 0.0f;
      }
      __syncthreads();
      // D_i = <dO_i, O_i> and lse_i for the tile: warp w covers queries w*8 .. w*8+7
      for (int jj = 0; jj < KT / ATT_WARPS; ++jj) {
        int j = warp * (KT / ATT_WARPS) + jj;
        float acc = 0.0f;
        if (q0 + j < seq)
          for (int dd = lane; dd < d; dd += 32) acc = fmaf(dOs[j * ld + dd], out[(row0 + q0 + j) * E + h * d + dd], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
          Ds[j] = acc;
          Ls[j] = (q0 + j < seq) ? lse[(row0 + q0 + j) * heads + h] : 0.0f;
        }
      }
      __syncthreads();
      if (!active) continue;
      float s = 0.0f, dp = 0.0f;
      for (int dd = 0; dd < d; ++dd) {
        s = fmaf(Qs[lane * ld + dd], ks[warp * ld + dd], s);
        dp = fmaf(dOs[lane * ld + dd], vs[warp * ld + dd], dp);
      }
      const float p = (q0 + lane < seq) ? __expf(s - Ls[lane]) : 0.0f;
      const float keep = keep_scale(drop_p, inv_keep, seed, row0 + q0 + lane, row0 + kj, h);
      const float ds = p * (dp * keep - Ds[lane]);
      const float pk = p * keep;
      for (int i = 0; i < KT; ++i) {
        const float pi = __shfl_sync(0xffffffffu, pk, i);
        const float dsi = __shfl_sync(0xffffffffu, ds, i);
#pragma unroll
        for (int t = 0; t < MAX_DT; ++t) {
          int dd = lane + 32 * t;
          if (dd < d) {
            dv[t] = fmaf(pi, dOs[i * ld + dd], dv[t]);
            dk[t] = fmaf(dsi, Qs[i * ld + dd], dk[t]);  // Qs already carries the 1/sqrt(d) factor
          }
        }
      }
    }
    if (active) {
#pragma unroll
      for (int t = 0; t < MAX_DT; ++t) {
        int dd = lane + 32 * t;
        if (dd < d) {
          dqkv[(row0 + kj) * ldq + E + h * d + dd] = dk[t];
          dqkv[(row0 + kj) * ldq + 2 * E + h * d + dd] = dv[t];
        }
      }
    }
  }
}

// Many small heads (Morgan / RDKit-2048: 256 heads of dimension 8, 20250113.py:71-73): one block per (group, head),
// one thread per query.  K and V of the head (seq x D each) live in shared memory and every thread walks all keys with
// its query row in registers: score reads are warp broadcasts, no cross-lane traffic, exact two-pass softmax.
template <int D>
__global__ void __launch_bounds__(256) attention_small_head_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                       float* __restrict__ lse, int seq, int heads) {
  extern __shared__ float smem[];
  float* Ks = smem;
  float* Vs = smem + (size_t)seq * D;
  const int h = blockIdx.x, g = blockIdx.y;
  const int E = heads * D, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  for (int i = threadIdx.x; i < seq * D; i += blockDim.x) {
    const int j = i / D, dd = i % D;
    const float* src = qkv + (row0 + j) * ldq + h * D + dd;
    Ks[i] = src[E];
    Vs[i] = src[2 * E];
  }
  __syncthreads();
  const float scale = rsqrtf((float)D);
  for (int qi = threadIdx.x; qi < seq; qi += blockDim.x) {
    float q[D], o[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      q[dd] = qkv[(row0 + qi) * ldq + h * D + dd] * scale;
      o[dd] = 0.0f;
    }
    float m = -INFINITY;
    for (int j = 0; j < seq; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int dd = 0; dd < D; ++dd) s = fmaf(q[dd], Ks[j * D + dd], s);
      m = fmaxf(m, s);
    }
    float l = 0.0f;
    for (int j = 0; j < seq; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int dd = 0; dd < D; ++dd) s = fmaf(q[dd], Ks[j * D + dd], s);
      const float p = __expf(s - m);
      l += p;
#pragma unroll
      for (int dd = 0; dd < D; ++dd) o[dd] = fmaf(p, Vs[j * D + dd], o[dd]);
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) out[(row0 + qi) * E + h * D + dd] = o[dd] * inv;
    if (lse) lse[(row0 + qi) * heads + h] = m + __logf(l);
  }
}


// ---- short attention scopes (seq <= 32: one reference training batch) ---------------------------------------------
// One CTA (32 warps) per (group, head) holds Q, K, V (and dO) of the whole scope in shared memory: forward is scores ->
// softmax -> PV in one launch, backward produces dQ, dK and dV in one launch.  The streaming kernels above need 2 + 8 +
// 8 CTAs of 4 warps for this shape and re-stage K/V per query block.  A single SM is instruction-issue bound on this
// much work, so every product runs on 128-bit shared-memory accesses:
//   scores   warp = query, lane = key:        Q row broadcast, K rows at a pitch of ld = 4*odd floats (conflict free)
//   outputs  warp = 4-column block, lane = row: the V / K / Q / dO operand is one broadcast float4 per j, the P / dS
//            operand a contiguous row of the (transposed where needed) 32 x 32 matrix
constexpr int SS = 32;          // maximum scope = warps per CTA
constexpr int SP = 33;          // pitch of the 32 x 32 matrices
__host__ __device__ inline int short_ld(int d) { return 4 * (((d + 3) / 4) | 1); }

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void axpy4(float a, float4 x, float4& y) {
  y.x = fmaf(a, x.x, y.x);
  y.y = fmaf(a, x.y, y.y);
  y.z = fmaf(a, x.z, y.z);
  y.w = fmaf(a, x.w, y.w);
}
// rows [0, SS) x columns [0, 4*ceil(d/4)) of a (seq, d) block with row pitch ld_src -> smem rows of pitch ld, zero outside
__device__ __forceinline__ void short_stage(float* dst, int ld, const float* src, size_t ld_src, int seq, int d) {
  const int row = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int d4 = (d + 3) / 4 * 4;
  for (int dd = lane; dd < d4; dd += 32) {
    const bool ok = row < seq && dd < d;
    cp_async_f32(dst + row * ld + dd, ok ? src + row * ld_src + dd : src, ok);
  }
}

// smem: Q[SS][ld] | K[SS][ld] | V[SS][ld] | PT[SS][SP] (PT[key][query])
__global__ void __launch_bounds__(SS * 32) attention_short_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                      float* __restrict__ lse, int seq, int heads, int d,
                                                                      float drop_p, uint64_t seed,
                                                                      const uint64_t* __restrict__ seed_dev) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  extern __shared__ __align__(16) float smem[];
  const int ld = short_ld(d), n4 = (d + 3) / 4;
  float* Qs = smem;
  float* Ks = Qs + SS * ld;
  float* Vs = Ks + SS * ld;
  float* PT = Vs + SS * ld;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int h = blockIdx.x, g = blockIdx.y;
  const int E = heads * d, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  const float scale = rsqrtf((float)d);
  const float inv_keep = 1.0f / (1.0f - drop_p);
  const float* base = qkv + row0 * ldq + h * d;
  short_stage(Qs, ld, base, ldq, seq, d);
  short_stage(Ks, ld, base + E, ldq, seq, d);
  short_stage(Vs, ld, base + 2 * E, ldq, seq, d);
  cp_async_wait_all();
  __syncthreads();
  {  // scores and softmax of query `warp`
    const int qi = warp;
    float s = 0.0f;
    const float4* q4 = reinterpret_cast<const float4*>(Qs + qi * ld);
    const float4* k4 = reinterpret_cast<const float4*>(Ks + lane * ld);
    for (int c = 0; c < n4; ++c) s = dot4(q4[c], k4[c], s);
    s = (qi < seq && lane < seq) ? s * scale : -INFINITY;
    float p = 0.0f;
    if (qi < seq) {                                 // warp-uniform
      const float m = warp_max(s);
      p = __expf(s - m);
      const float l = warp_sum(p);
      // dropout acts on the normalised probabilities: the normaliser l keeps every key
      p = p * keep_scale(drop_p, inv_keep, seed, row0 + qi, row0 + lane, h) / l;
      if (lse && lane == 0) lse[(row0 + qi) * heads + h] = m + __logf(l);
    }
    PT[lane * SP + qi] = p;
  }
  __syncthreads();
  // O = P V: warp = 4-column block c, lane = query.  The result is staged in Qs (no longer read: every warp passed the
  // barrier above after its last use) so that the global store below is one contiguous row per warp -- storing straight
  // from this (lane = row) mapping would touch 32 different rows per instruction.
  for (int c = warp; c < n4; c += SS) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < seq; ++j) axpy4(PT[j * SP + lane], *reinterpret_cast<const float4*>(Vs + j * ld + 4 * c), o);
    *reinterpret_cast<float4*>(Qs + lane * ld + 4 * c) = o;
  }
  __syncthreads();
  if (warp < seq) {
    float* dst = out + (row0 + warp) * E + h * d;
    for (int dd = lane; dd < d; dd += 32) dst[dd] = Qs[warp * ld + dd];
  }
}

// smem: Q | K | V | dO (each [SS][ld]) | P[SS][SP] (kept probabilities, [query][key]) | dS[SS][SP] | dST[SS][SP]
__global__ void __launch_bounds__(SS * 32) attention_short_bwd_kernel(const float* __restrict__ qkv,
                                                                      const float* __restrict__ out,
                                                                      const float* __restrict__ lse,
                                                                      const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                      int seq, int heads, int d, float drop_p, uint64_t seed,
                                                                      const uint64_t* __restrict__ seed_dev) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  extern __shared__ __align__(16) float smem[];
  const int ld = short_ld(d), n4 = (d + 3) / 4;
  float* Qs = smem;
  float* Ks = Qs + SS * ld;
  float* Vs = Ks + SS * ld;
  float* dOs = Vs + SS * ld;
  float* Ps = dOs + SS * ld;
  float* dSs = Ps + SS * SP;
  float* dST = dSs + SS * SP;
  float* Gs = dST + SS * SP;        // dQ | dK | dV staging, [3][SS][ld]: coalesced row stores at the end
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int h = blockIdx.x, g = blockIdx.y;
  const int E = heads * d, ldq = 3 * E;
  const size_t row0 = (size_t)g * seq;
  const float scale = rsqrtf((float)d);
  const float inv_keep = 1.0f / (1.0f - drop_p);
  const float* base = qkv + row0 * ldq + h * d;
  short_stage(Qs, ld, base, ldq, seq, d);
  short_stage(Ks, ld, base + E, ldq, seq, d);
  short_stage(Vs, ld, base + 2 * E, ldq, seq, d);
  short_stage(dOs, ld, dout + row0 * E + h * d, E, seq, d);
  // D_i = <dO_i, O_i> straight from global memory while the copies fly
  float D = 0.0f, L = 0.0f;
  if (warp < seq) {
    for (int dd = lane; dd < d; dd += 32)
      D = fmaf(dout[(row0 + warp) * E + h * d + dd], out[(row0 + warp) * E + h * d + dd], D);
    D = warp_sum(D);
    L = lse[(row0 + warp) * heads + h];
  }
  cp_async_wait_all();
  __syncthreads();
  {  // P and dS of query `warp` (lane = key)
    const int qi = warp;
    float s = 0.0f, dp = 0.0f;
    const float4* q4 = reinterpret_cast<const float4*>(Qs + qi * ld);
    const float4* g4 = reinterpret_cast<const float4*>(dOs + qi * ld);
    const float4* k4 = reinterpret_cast<const float4*>(Ks + lane * ld);
    const float4* v4 = reinterpret_cast<const float4*>(Vs + lane * ld);
    for (int c = 0; c < n4; ++c) {
      s = dot4(q4[c], k4[c], s);
      dp = dot4(g4[c], v4[c], dp);
    }
    const float p = (qi < seq && lane < seq) ? __expf(s * scale - L) : 0.0f;
    const float keep = keep_scale(drop_p, inv_keep, seed, row0 + qi, row0 + lane, h);
    const float ds = p * (dp * keep - D);
    Ps[qi * SP + lane] = p * keep;
    dSs[qi * SP + lane] = ds;
    dST[lane * SP + qi] = ds;
  }
  __syncthreads();
  // warp = 4-column block c, lane = row r:
  //   dQ_r = scale * sum_j dS[r][j] K_j ; dK_r = scale * sum_j dS[j][r] Q_j ; dV_r = sum_j P[j][r] dO_j
  for (int c = warp; c < n4; c += SS) {
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dk = dq, dv = dq;
    for (int j = 0; j < seq; ++j) {
      axpy4(dST[j * SP + lane], *reinterpret_cast<const float4*>(Ks + j * ld + 4 * c), dq);
      axpy4(dSs[j * SP + lane], *reinterpret_cast<const float4*>(Qs + j * ld + 4 * c), dk);
      axpy4(Ps[j * SP + lane], *reinterpret_cast<const float4*>(dOs + j * ld + 4 * c), dv);
    }
    dq.x *= scale, dq.y *= scale, dq.z *= scale, dq.w *= scale;
    dk.x *= scale, dk.y *= scale, dk.z *= scale, dk.w *= scale;
    *reinterpret_cast<float4*>(Gs + lane * ld + 4 * c) = dq;
    *reinterpret_cast<float4*>(Gs + (SS + lane) * ld + 4 * c) = dk;
    *reinterpret_cast<float4*>(Gs + (2 * SS + lane) * ld + 4 * c) = dv;
  }
  __syncthreads();
  if (warp < seq) {           // one contiguous row of dq, dk and dv per warp
    float* dst = dqkv + (row0 + warp) * ldq + h * d;
    for (int dd = lane; dd < d; dd += 32) {
      dst[dd] = Gs[warp * ld + dd];
      dst[E + dd] = Gs[(SS + warp) * ld + dd];
      dst[2 * E + dd] = Gs[(2 * SS + warp) * ld + dd];
    }
  }
}

static size_t short_fwd_smem(int d) { return ((size_t)3 * SS * short_ld(d) + SS * SP) * sizeof(float); }
static size_t short_bwd_smem(int d) { return ((size_t)7 * SS * short_ld(d) + 3 * SS * SP) * sizeof(float); }

// ---- mid-size single-head scopes on the training path (32 < seq, e.g. batch 256): GEMM route ------------------------
// scores = Q K^T and O = P V (and the four backward products) run on the tiled fp32 GEMM; these two kernels are the
// row softmax in between.  One block per query row; the keep mask is the same Philox function as everywhere else.
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                            // red may still be read from a previous reduction
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = is_max ? fmaxf(t, red[i]) : t + red[i];
  return t;
}

// p[r, :] = softmax(scale * s[r, :]);  pd = p * keep (dropout on the normalised probabilities); pd may alias p when drop_p == 0
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(const float* __restrict__ s, int ld_s, float* __restrict__ p,
                                                               float* __restrict__ pd, int ld_p, int cols, float scale,
                                                               float drop_p, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                                                               size_t row_base) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  __shared__ float red[8];
  const size_t r = blockIdx.x;
  const float* sr = s + r * ld_s;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) mx = fmaxf(mx, sr[c] * scale);
  mx = block_reduce(mx, red, true);
  float sum = 0.0f;
  for (int c = threadIdx.x; c < cols; c += 256) sum += __expf(sr[c] * scale - mx);
  sum = block_reduce(sum, red, false);
  const float inv = 1.0f / sum, inv_keep = 1.0f / (1.0f - drop_p);
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float v = __expf(sr[c] * scale - mx) * inv;
    p[r * ld_p + c] = v;
    if (drop_p > 0.0f) pd[r * ld_p + c] = v * keep_scale(drop_p, inv_keep, seed, row_base + r, row_base + c, 0);
  }
}

// ds[r, :] = scale * p * (dpd * keep - D_r),  D_r = sum_c dpd * keep * p
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dpd,
                                                               float* __restrict__ ds, int ld, int cols, float scale,
                                                               float drop_p, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                                                               size_t row_base) {
  seed = mix_seed(seed, seed_dev);  // per-site seed x per-step device seed (common.cuh)
  __shared__ float red[8];
  const size_t r = blockIdx.x;
  const float inv_keep = 1.0f / (1.0f - drop_p);
  float D = 0.0f;
  for (int c = threadIdx.x; c < cols; c += 256)
    D = fmaf(dpd[r * ld + c] * keep_scale(drop_p, inv_keep, seed, row_base + r, row_base + c, 0), p[r * ld + c], D);
  D = block_reduce(D, red, false);
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float dp = dpd[r * ld + c] * keep_scale(drop_p, inv_keep, seed, row_base + r, row_base + c, 0);
    ds[r * ld + c] = scale * p[r * ld + c] * (dp - D);
  }
}

// Query (or key) rows per CTA: 16 (four passes of the CTA's four warps over one staged K/V tile stream) when that
// already fills the GPU, otherwise 4 (one pass) -- a reference training batch (seq 32, one head) is 2 CTAs at 16 rows
// per CTA and the kernel is pure latency.
static int rows_per_block(int groups, int seq, int heads) {
  return (long long)ceil_div(seq, 16) * heads * groups >= 2 * 148 ? 16 : ATT_WARPS;
}

static int attention_args_ok(const char* who, int groups, int seq, int heads, int d) {
  if (groups < 0 || seq <= 0 || heads <= 0 || d <= 0 || d > 32 * MAX_DT) {
    set_error("%s: groups=%d seq=%d heads=%d head_dim=%d unsupported (head_dim <= %d)", who, groups, seq, heads, d,
              32 * MAX_DT);
    return 0;
  }
  if (heads > 65535 || groups > 65535) {
    set_error("%s: heads/groups exceed 65535", who);
    return 0;
  }
  return 1;
}

}  // namespace bbbp

extern "C" int bbbp_attention_fwd_f32(const float* qkv, float* out, float* lse, int groups, int seq, int heads,
                                      int head_dim, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                                      bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(qkv && out, "attention_fwd: null operand");
  BBBP_CHECK_ARG(dropout_p >= 0.0f && dropout_p < 1.0f, "attention_fwd: dropout_p must be in [0,1)");
  if (!attention_args_ok("attention_fwd", groups, seq, heads, head_dim)) return BBBP_EINVAL;
  if (groups == 0) return BBBP_OK;
  if (dropout_p == 0.0f && (head_dim == 8 || head_dim == 16) && (size_t)seq * head_dim * 8 <= 160 * 1024) {
    const size_t sm = (size_t)seq * head_dim * 2 * sizeof(float);
    const int threads = seq >= 256 ? 256 : (seq + 31) / 32 * 32;
    dim3 grid(heads, groups);
    if (head_dim == 8) {
      cudaFuncSetAttribute(attention_small_head_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      attention_small_head_fwd_kernel<8><<<grid, threads, sm, as_stream(stream)>>>(qkv, out, lse, seq, heads);
    } else {
      cudaFuncSetAttribute(attention_small_head_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      attention_small_head_fwd_kernel<16><<<grid, threads, sm, as_stream(stream)>>>(qkv, out, lse, seq, heads);
    }
    return launch_status("attention_fwd (small heads)");
  }
  if (seq <= SS && short_fwd_smem(head_dim) <= 200 * 1024) {   // a function of the scope only: bit-identical for any grouping
    const size_t sm = short_fwd_smem(head_dim);
    cudaFuncSetAttribute(attention_short_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attention_short_fwd_kernel<<<dim3(heads, groups), SS * 32, sm, as_stream(stream)>>>(qkv, out, lse, seq, heads, head_dim,
                                                                                             dropout_p, seed, seed_dev);
    return launch_status("attention_fwd (short scope)");
  }
  const int ld = head_dim | 1;
  const int qpb = rows_per_block(groups, seq, heads);
  size_t smem = (size_t)(2 * KT + ATT_WARPS) * ld * sizeof(float);
  cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ceil_div(seq, qpb), heads, groups);
  attention_fwd_kernel<<<grid, ATT_WARPS * 32, smem, as_stream(stream)>>>(qkv, out, lse, seq, heads, head_dim, qpb,
                                                                          dropout_p, seed, seed_dev);
  return launch_status("attention_fwd");
}

extern "C" int bbbp_attn_softmax_fwd_f32(const float* scores, int ld_scores, float* p, float* p_dropped, int ld_p, int rows,
                                         int cols, float scale, float dropout_p, uint64_t seed, const uint64_t* seed_dev,
                                         long long row_base, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(scores && p && rows >= 0 && cols > 0 && dropout_p >= 0.0f && dropout_p < 1.0f, "attn_softmax_fwd: bad argument");
  BBBP_CHECK_ARG(dropout_p == 0.0f || p_dropped, "attn_softmax_fwd: dropout needs the p_dropped output");
  if (rows == 0) return BBBP_OK;
  attn_softmax_fwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(scores, ld_scores, p, p_dropped, ld_p, cols, scale, dropout_p, seed,
                                                               seed_dev, (size_t)row_base);
  return launch_status("attn_softmax_fwd");
}

extern "C" int bbbp_attn_softmax_bwd_f32(const float* p, const float* dp_dropped, float* dscores, int ld, int rows, int cols,
                                         float scale, float dropout_p, uint64_t seed, const uint64_t* seed_dev, long long row_base,
                                         bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(p && dp_dropped && dscores && rows >= 0 && cols > 0 && dropout_p >= 0.0f && dropout_p < 1.0f,
                 "attn_softmax_bwd: bad argument");
  if (rows == 0) return BBBP_OK;
  attn_softmax_bwd_kernel<<<rows, 256, 0, as_stream(stream)>>>(p, dp_dropped, dscores, ld, cols, scale, dropout_p, seed, seed_dev,
                                                               (size_t)row_base);
  return launch_status("attn_softmax_bwd");
}

extern "C" int bbbp_attention_bwd_f32(const float* qkv, const float* out, const float* lse, const float* dout,
                                      float* dqkv, int groups, int seq, int heads, int head_dim, float dropout_p,
                                      uint64_t seed, const uint64_t* seed_dev, bbbp_stream_t stream) {
  using namespace bbbp;
  BBBP_CHECK_ARG(qkv && out && lse && dout && dqkv, "attention_bwd: null operand");
  if (!attention_args_ok("attention_bwd", groups, seq, heads, head_dim)) return BBBP_EINVAL;
  if (groups == 0) return BBBP_OK;
  if (seq <= SS && short_bwd_smem(head_dim) <= 200 * 1024) {
    const size_t sm = short_bwd_smem(head_dim);
    cudaFuncSetAttribute(attention_short_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attention_short_bwd_kernel<<<dim3(heads, groups), SS * 32, sm, as_stream(stream)>>>(
        qkv, out, lse, dout, dqkv, seq, heads, head_dim, dropout_p, seed, seed_dev);
    return launch_status("attention_bwd (short scope)");
  }
  const int ld = head_dim | 1;
  const int per_block = rows_per_block(groups, seq, heads);
  dim3 grid(ceil_div(seq, per_block), heads, groups);
  size_t smem_q = (size_t)(2 * KT + 2 * ATT_WARPS) * ld * sizeof(float);
  cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
  attention_bwd_dq_kernel<<<grid, ATT_WARPS * 32, smem_q, as_stream(stream)>>>(qkv, out, lse, dout, dqkv, seq, heads,
                                                                               head_dim, per_block, dropout_p, seed, seed_dev);
  int st = launch_status("attention_bwd dq");
  if (st != BBBP_OK) return st;
  size_t smem_kv = ((size_t)(2 * KT + 2 * ATT_WARPS) * ld + 2 * KT) * sizeof(float);
  cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv);
  attention_bwd_dkv_kernel<<<grid, ATT_WARPS * 32, smem_kv, as_stream(stream)>>>(qkv, out, lse, dout, dqkv, seq, heads,
                                                                                 head_dim, per_block, dropout_p, seed, seed_dev);
  return launch_status("attention_bwd dkv");
}
