// Whole-model inference entry points: bbbp_model_prepare / bbbp_workspace_bytes / bbbp_fwd.
//
// Host code only.  MixedInputModel.forward of Models/multi_input_data_regression_opt_transformer_cnn_20250113.py:109-119
// (eval mode) is a fixed sequence of this library's own entry points; this file runs that sequence for a host that is not
// Python.  It mirrors bbbp_b200/model.py launch for launch (TransformerCnnModel._encoder_tensor_core,
// _image_branch_tensor_core, MultiHeadAttentionFusion.forward, _head), with the same operand pitches and split-K factors,
// so the scores are bit-identical to the Python host's (tests/test_model_gpu.py::test_c_host_forward_*).
// The caller owns every byte: parameters, the prepared (derived-weight) arena and the workspace are carved by the two
// bump allocators below, which also answer the *_bytes() queries by running without a base pointer.
#include <cmath>
#include <cstring>
#include <cstdio>
#include "common.cuh"

namespace {

constexpr int kLayers = 6;        // nn.TransformerEncoder(num_layers=6), 20250113.py:75-78
constexpr int kFF = 2048;         // nn.TransformerEncoderLayer default dim_feedforward
constexpr int kBranch = 128;      // fingerprint_fc / image_cnn output width, 20250113.py:80,92
constexpr int kFused = 2 * kBranch;
constexpr int kFusionHeads = 4;   // MultiHeadAttentionFusion(256, num_heads=4), 20250113.py:95
constexpr int kFusionHidden = 128;
constexpr int kSide = 128;        // image.view(-1, 3, 128, 128), 20250113.py:114
constexpr int kImage = 3 * kSide * kSide;
constexpr int kFlat = 64 * 32 * 32;   // nn.Flatten() after two conv + pool blocks
constexpr float kLnEps = 1e-5f, kBnEps = 1e-5f, kBnMomentum = 0.1f;

// ---- the reference's state_dict order ------------------------------------------------------------------------------------
enum { L_IN_W, L_IN_B, L_OUT_W, L_OUT_B, L_L1_W, L_L1_B, L_L2_W, L_L2_B, L_N1_W, L_N1_B, L_N2_W, L_N2_B, L_COUNT };
enum {
  P_FPFC_W = kLayers * L_COUNT, P_FPFC_B, P_C1_W, P_C1_B, P_C2_W, P_C2_B, P_IFC_W, P_IFC_B,
  P_FUSION,                                      // + 4*head + {0: first weight, 1: first bias, 2: second weight, 3: second bias}
  P_H0_W = P_FUSION + 4 * kFusionHeads, P_H0_B, P_BN_W, P_BN_B, P_BN_MEAN, P_BN_VAR, P_BN_COUNT, P_H3_W, P_H3_B, P_H5_W,
  P_H5_B, P_H7_W, P_H7_B, P_TOTAL
};
static_assert(P_TOTAL == 109, "MixedInputModel.state_dict() has 109 entries");

inline int ceil8(int v) { return (v + 7) / 8 * 8; }
inline int ceil16(int v) { return (v + 15) / 16 * 16; }

// 20250113.py:71-73: nhead = max(1, F // 8), lowered until it divides F
int encoder_heads(int F) {
  int h = F / 8 > 1 ? F / 8 : 1;
  while (F % h != 0 && h > 1) --h;
  return h;
}

struct Dims {
  int F, Fq, heads, head_dim, groups, seq, fmt;
  long long R;
  bool split, u8, flash;
};

int parse(const bbbp_model_desc* d, Dims* m, bool need_batch) {
  BBBP_CHECK_ARG(d != nullptr, "bbbp_model: desc is NULL");
  BBBP_CHECK_ARG(d->abi_version == BBBP_ABI_VERSION, "bbbp_model: desc.abi_version %d, library is %d", d->abi_version,
                 BBBP_ABI_VERSION);
  if (d->variant != BBBP_MODEL_TCNN_20250113) {
    bbbp::set_error("bbbp_model: variant %d is not built into the whole-model entry points", d->variant);
    return BBBP_EUNSUPPORTED;
  }
  BBBP_CHECK_ARG(d->fingerprint_size >= 1, "bbbp_model: fingerprint_size %d", d->fingerprint_size);
  m->F = d->fingerprint_size;
  m->Fq = ceil8(m->F);
  m->heads = encoder_heads(m->F);
  m->head_dim = m->F / m->heads;
  switch (d->precision) {
    case BBBP_PREC_BF16: m->fmt = BBBP_FMT_BF16; m->split = false; break;
    case BBBP_PREC_F16: m->fmt = BBBP_FMT_F16; m->split = false; break;
    case BBBP_PREC_STRICT: m->fmt = BBBP_FMT_F16; m->split = true; break;
    default:
      bbbp::set_error("bbbp_model: precision %d: the whole-model entry points are built for BF16, F16 and STRICT (the fp32 "
                      "validation mode runs through the per-kernel entry points)", d->precision);
      return BBBP_EUNSUPPORTED;
  }
  if (!(m->heads == 1 || m->head_dim == 8 || m->head_dim == 16)) {
    bbbp::set_error("bbbp_model: fingerprint_size %d gives %d heads of dimension %d; built: one head, or heads of 8 / 16",
                    m->F, m->heads, m->head_dim);
    return BBBP_EUNSUPPORTED;
  }
  m->u8 = d->image_is_u8 != 0;
  m->groups = d->groups;
  m->seq = d->seq;
  m->R = (long long)d->groups * d->seq;
  m->flash = false;
  if (need_batch) {
    BBBP_CHECK_ARG(d->groups >= 1 && d->seq >= 1, "bbbp_model: groups %d, seq %d", d->groups, d->seq);
    BBBP_CHECK_ARG(m->R <= (1ll << 22), "bbbp_model: %lld molecules per call (limit 4 194 304)", m->R);
    if (m->heads == 1 && m->seq > 256 && m->F > 192) {
      bbbp::set_error("bbbp_model: one head of dimension %d with %d molecules per batch is not built (head_dim <= 192)", m->F,
                      m->seq);
      return BBBP_EUNSUPPORTED;
    }
    // scopes of 129 molecules or more take the streaming-softmax kernel (model.py: flash_min_seq)
    m->flash = m->heads == 1 && m->F <= 192 && m->seq >= 129;
  }
  return BBBP_OK;
}

// ---- bump allocator over a caller-owned buffer (base == NULL: size query) --------------------------------------------------
struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* b) : base(static_cast<char*>(b)) {}
  template <typename T = void>
  T* take(size_t bytes) {
    const size_t at = (off + 255) & ~size_t(255);
    off = at + bytes;
    return base ? reinterpret_cast<T*>(base + at) : nullptr;
  }
};

struct Prepared {
  struct Layer {
    void *w_in, *w_out, *w_l1, *w_l2;
    float* b_in;
  } layer[kLayers];
  void *fpfc_hi, *fpfc_lo, *conv1, *conv2, *wfc;
  float *tap1, *tap2, *possum;                 // strict mode: conv weights summed over taps, fc weight over positions
  void *fus1_hi, *fus1_lo, *fus2_hi, *fus2_lo;
  float *fus_b1, *fus_b2;
  void *head_hi[4], *head_lo[4];
};
const int kHeadN[4] = {256, 128, 64, 1}, kHeadK[4] = {256, 256, 128, 64};
const int kHeadW[4] = {P_H0_W, P_H3_W, P_H5_W, P_H7_W};

void lay_prepared(const Dims& m, Arena& a, Prepared* p) {
  const size_t e = 2;     // bytes per 16-bit element
  for (auto& l : p->layer) {
    l.w_in = a.take(size_t(3) * m.Fq * m.Fq * e);
    l.b_in = a.take<float>(size_t(3) * m.Fq * 4);
    l.w_out = a.take(size_t(m.F) * m.Fq * e);
    l.w_l1 = a.take(size_t(kFF) * m.Fq * e);
    l.w_l2 = a.take(size_t(m.F) * kFF * e);
  }
  p->fpfc_hi = a.take(size_t(kBranch) * m.Fq * e);
  p->fpfc_lo = m.split ? a.take(size_t(kBranch) * m.Fq * e) : nullptr;
  p->conv1 = a.take(bbbp_conv3x3_prepared_bytes(3, 32));
  p->conv2 = a.take(bbbp_conv3x3_prepared_bytes(32, 64));
  p->wfc = a.take(size_t(kBranch) * kFlat * e);
  p->tap1 = m.split ? a.take<float>(32 * 3 * 4) : nullptr;
  p->tap2 = m.split ? a.take<float>(64 * 32 * 4) : nullptr;
  p->possum = m.split ? a.take<float>(kBranch * 64 * 4) : nullptr;
  const size_t n1 = size_t(kFusionHeads) * kFusionHidden * kFused, n2 = size_t(kFusionHeads) * kFusionHeads * kFusionHidden;
  p->fus1_hi = a.take(n1 * e);
  p->fus1_lo = m.split ? a.take(n1 * e) : nullptr;
  p->fus2_hi = a.take(n2 * e);
  p->fus2_lo = m.split ? a.take(n2 * e) : nullptr;
  p->fus_b1 = a.take<float>(kFusionHeads * kFusionHidden * 4);
  p->fus_b2 = a.take<float>(kFusionHeads * 4);
  for (int i = 0; i < 4; ++i) {
    const size_t n = size_t(kHeadN[i]) * ceil8(kHeadK[i]) * e;
    p->head_hi[i] = a.take(n);
    p->head_lo[i] = m.split ? a.take(n) : nullptr;
  }
}

struct Work {
  float *x32[2], *sum32, *both, *stats, *bg1, *tab1, *tab2, *fc_add, *scores, *fused, *h0, *hbn, *h3, *h5;
  void *x16[2], *qkv, *probs, *vt, *attn, *hidden, *fp_hi, *fp_lo, *neg2, *y1, *y2, *splitk, *cat_hi, *cat_lo, *fh_hi,
      *fh_lo, *t_hi, *t_lo;
  size_t splitk_bytes;
  int ldp;
};

void lay_workspace(const Dims& m, Arena& a, Work* w) {
  const size_t R = (size_t)m.R, e = 2;
  w->ldp = ceil8(m.seq);
  for (int i = 0; i < 2; ++i) {
    w->x32[i] = a.take<float>(R * m.Fq * 4);
    w->x16[i] = a.take(R * m.Fq * e);
  }
  w->qkv = a.take(R * 3 * m.Fq * e);
  w->probs = (m.heads == 1 && !m.flash) ? a.take(R * w->ldp * e) : nullptr;
  w->vt = m.heads == 1 ? a.take(size_t(m.groups) * m.F * w->ldp * e) : nullptr;
  w->attn = a.take(R * m.Fq * e);
  w->sum32 = a.take<float>(R * m.Fq * 4);
  w->hidden = m.F > 176 ? a.take(R * kFF * e) : nullptr;     // narrower models run the feed-forward block as one fused kernel
  w->fp_hi = m.split ? a.take(R * m.Fq * e) : nullptr;
  w->fp_lo = m.split ? a.take(R * m.Fq * e) : nullptr;
  w->both = a.take<float>(R * kFused * 4);
  w->stats = m.u8 ? a.take<float>(R * 2 * 4) : nullptr;
  w->bg1 = m.split ? a.take<float>(R * 4 * 4) : nullptr;
  w->tab1 = m.split ? a.take<float>(R * 2 * 32 * 4) : nullptr;
  w->neg2 = m.split ? a.take(R * 32 * e) : nullptr;
  w->tab2 = m.split ? a.take<float>(R * 2 * 64 * 4) : nullptr;
  w->fc_add = m.split ? a.take<float>(R * kBranch * 4) : nullptr;
  w->y1 = a.take(R * 64 * 64 * 32 * e);
  w->y2 = a.take(R * kFlat * e);
  const int sk = m.split ? 32 : 8;      // ops.fixed_split_k_strict(65536) / ops.fixed_split_k(65536)
  w->splitk_bytes = bbbp_gemm_bf16_workspace((int)m.R, kBranch, sk);
  w->splitk = a.take(w->splitk_bytes);
  w->cat_hi = a.take(R * kFused * e);
  w->cat_lo = m.split ? a.take(R * kFused * e) : nullptr;
  w->fh_hi = a.take(R * kFusionHeads * kFusionHidden * e);
  w->fh_lo = m.split ? a.take(R * kFusionHeads * kFusionHidden * e) : nullptr;
  w->scores = a.take<float>(R * kFusionHeads * 4);
  w->fused = a.take<float>(R * kFused * 4);
  w->t_hi = a.take(R * 256 * e);
  w->t_lo = m.split ? a.take(R * 256 * e) : nullptr;
  w->h0 = a.take<float>(R * 256 * 4);
  w->hbn = a.take<float>(R * 256 * 4);
  w->h3 = a.take<float>(R * 128 * 4);
  w->h5 = a.take<float>(R * 64 * 4);
}

#define RUN(call)                    \
  do {                               \
    const int rc_ = (call);          \
    if (rc_ != BBBP_OK) return rc_;  \
  } while (0)

inline const float* fparam(const void* const* params, int i) { return static_cast<const float*>(params[i]); }
inline char* at16(void* base, size_t elements) { return static_cast<char*>(base) + 2 * elements; }

// ops.gemm_bf16: act(A W^T + bias [+ pre_add]) [+ residual]; the split-K finish kernel writes only the N result columns of a
// 16-bit output, so its pad columns are zeroed first
int gemm(const Dims& m, long long M, int N, int K, const void* a_hi, const void* a_lo, int lda, const void* w_hi,
         const void* w_lo, int ldw, const float* bias, const float* residual, int ld_res, const float* pre_add, int ld_pre,
         float* out32, int ld_out, void* out16, void* out16_lo, int ld16, int act, int split_k, void* ws, size_t ws_bytes,
         bbbp_stream_t s) {
  const size_t need = bbbp_gemm_bf16_workspace((int)M, N, split_k);
  if (need > ws_bytes) {
    bbbp::set_error("bbbp_fwd: split-K workspace %zu > %zu", need, ws_bytes);
    return BBBP_EWORKSPACE;
  }
  if (out16 && need && ld16 > N) {
    RUN(bbbp_fill_zero(at16(out16, N), M, (ld16 - N) * 2ll, ld16 * 2ll, s));
    if (out16_lo) RUN(bbbp_fill_zero(at16(out16_lo, N), M, (ld16 - N) * 2ll, ld16 * 2ll, s));
  }
  if (pre_add)
    return bbbp_gemm16_pre(m.fmt, (int)M, N, K, a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, pre_add, ld_pre, out32, ld_out, out16,
                           out16_lo, ld16, act, split_k, need ? ws : nullptr, need, s);
  return bbbp_gemm16(m.fmt, (int)M, N, K, a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, residual, ld_res, out32, ld_out, out16,
                     out16_lo, ld16, act, split_k, need ? ws : nullptr, need, s);
}

// autograd.Linear.forward in a tensor-core mode: a generic nn.Linear call site (both operands split in the strict mode)
int linear(const Dims& m, const Work& w, long long M, int N, int K, const float* x, int ldx, const void* w_hi, const void* w_lo,
           const float* bias, float* out, int ld_out, int act, bbbp_stream_t s) {
  const int ld = ceil8(K);
  RUN(bbbp_cast16(m.fmt, x, ldx, nullptr, w.t_hi, m.split ? w.t_lo : nullptr, ld, (int)M, K, ld, s));
  return gemm(m, M, N, K, w.t_hi, m.split ? w.t_lo : nullptr, ld, w_hi, w_lo, ld, bias, nullptr, 0, nullptr, 0, out, ld_out,
              nullptr, nullptr, ceil8(N), act, 1, nullptr, 0, s);
}

}  // namespace

extern "C" int bbbp_model_param_count(const bbbp_model_desc* desc) {
  Dims m;
  RUN(parse(desc, &m, false));
  return P_TOTAL;
}

extern "C" const char* bbbp_model_param_name(const bbbp_model_desc* desc, int i) {
  static thread_local char name[96];
  Dims m;
  if (parse(desc, &m, false) != BBBP_OK || i < 0 || i >= P_TOTAL) return nullptr;
  static const char* kLayer[L_COUNT] = {"self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                                        "self_attn.out_proj.bias",  "linear1.weight",         "linear1.bias",
                                        "linear2.weight",           "linear2.bias",           "norm1.weight",
                                        "norm1.bias",               "norm2.weight",           "norm2.bias"};
  static const char* kRest[] = {"fingerprint_fc.0.weight", "fingerprint_fc.0.bias", "image_cnn.0.weight", "image_cnn.0.bias",
                                "image_cnn.3.weight",      "image_cnn.3.bias",      "image_cnn.7.weight", "image_cnn.7.bias"};
  static const char* kHead[] = {"fc.0.weight",       "fc.0.bias",        "fc.2.weight",           "fc.2.bias",  "fc.2.running_mean",
                                "fc.2.running_var",  "fc.2.num_batches_tracked", "fc.3.weight",   "fc.3.bias",  "fc.5.weight",
                                "fc.5.bias",         "fc.7.weight",      "fc.7.bias"};
  if (i < P_FPFC_W)
    snprintf(name, sizeof(name), "fingerprint_transformer.layers.%d.%s", i / L_COUNT, kLayer[i % L_COUNT]);
  else if (i < P_FUSION)
    snprintf(name, sizeof(name), "%s", kRest[i - P_FPFC_W]);
  else if (i < P_H0_W)
    snprintf(name, sizeof(name), "attention_fusion.attention_heads.%d.%d.%s", (i - P_FUSION) / 4, ((i - P_FUSION) % 4) / 2 * 2,
             (i - P_FUSION) % 2 ? "bias" : "weight");
  else
    snprintf(name, sizeof(name), "%s", kHead[i - P_H0_W]);
  return name;
}

extern "C" size_t bbbp_model_param_numel(const bbbp_model_desc* desc, int i) {
  Dims m;
  if (parse(desc, &m, false) != BBBP_OK || i < 0 || i >= P_TOTAL) return 0;
  const size_t F = (size_t)m.F;
  if (i < P_FPFC_W) {
    const size_t n[L_COUNT] = {3 * F * F, 3 * F, F * F, F, kFF * F, kFF, F * kFF, F, F, F, F, F};
    return n[i % L_COUNT];
  }
  if (i >= P_FUSION && i < P_H0_W) {
    const size_t n[4] = {size_t(kFusionHidden) * kFused, kFusionHidden, kFusionHidden, 1};
    return n[(i - P_FUSION) % 4];
  }
  switch (i) {
    case P_FPFC_W: return kBranch * F;
    case P_FPFC_B: return kBranch;
    case P_C1_W: return 32 * 3 * 9;
    case P_C1_B: return 32;
    case P_C2_W: return 64 * 32 * 9;
    case P_C2_B: return 64;
    case P_IFC_W: return size_t(kBranch) * kFlat;
    case P_IFC_B: return kBranch;
    case P_H0_W: return 256 * 256;
    case P_H0_B: case P_BN_W: case P_BN_B: case P_BN_MEAN: case P_BN_VAR: return 256;
    case P_BN_COUNT: return 1;
    case P_H3_W: return 128 * 256;
    case P_H3_B: return 128;
    case P_H5_W: return 64 * 128;
    case P_H5_B: return 64;
    case P_H7_W: return 64;
    case P_H7_B: return 1;
  }
  return 0;
}

extern "C" size_t bbbp_model_prepared_bytes(const bbbp_model_desc* desc) {
  Dims m;
  if (parse(desc, &m, false) != BBBP_OK) return 0;
  Arena a(nullptr);
  Prepared p;
  lay_prepared(m, a, &p);
  return a.off + 256;
}

extern "C" size_t bbbp_workspace_bytes(const bbbp_model_desc* desc) {
  Dims m;
  if (parse(desc, &m, true) != BBBP_OK) return 0;
  Arena a(nullptr);
  Work w;
  lay_workspace(m, a, &w);
  return a.off + 256;
}

extern "C" int bbbp_model_prepare(const bbbp_model_desc* desc, const void* const* params, void* prepared, size_t prepared_bytes,
                                  bbbp_stream_t s) {
  Dims m;
  RUN(parse(desc, &m, false));
  BBBP_CHECK_ARG(params && prepared, "bbbp_model_prepare: NULL argument");
  BBBP_CHECK_ARG((reinterpret_cast<uintptr_t>(prepared) & 255) == 0, "bbbp_model_prepare: prepared must be 256-byte aligned");
  for (int i = 0; i < P_TOTAL; ++i)
    BBBP_CHECK_ARG(params[i] != nullptr || i == P_BN_COUNT, "bbbp_model_prepare: params[%d] (%s) is NULL", i,
                   bbbp_model_param_name(desc, i));
  Arena a(prepared);
  Prepared p;
  lay_prepared(m, a, &p);
  if (a.off > prepared_bytes) {
    bbbp::set_error("bbbp_model_prepare: prepared buffer %zu bytes, need %zu", prepared_bytes, a.off);
    return BBBP_EWORKSPACE;
  }
  const int F = m.F, Fq = m.Fq, fmt = m.fmt;
  for (int l = 0; l < kLayers; ++l) {
    const void* const* lp = params + l * L_COUNT;
    const Prepared::Layer& d = p.layer[l];
    // in_proj (3F, F) -> (3*Fq, Fq): q | k | v each start on a 16-byte boundary (zero rows / columns in the pads)
    RUN(bbbp_fill_zero(d.w_in, 1, 3ll * Fq * Fq * 2, 3ll * Fq * Fq * 2, s));
    RUN(bbbp_fill_zero(d.b_in, 1, 3ll * Fq * 4, 3ll * Fq * 4, s));
    for (int part = 0; part < 3; ++part) {
      RUN(bbbp_cast16(fmt, fparam(lp, L_IN_W) + (size_t)part * F * F, F, nullptr, at16(d.w_in, (size_t)part * Fq * Fq), nullptr, Fq,
                      F, F, Fq, s));
      RUN(bbbp_copy2d_f32(fparam(lp, L_IN_B) + part * F, F, d.b_in + part * Fq, Fq, 1, F, s));
    }
    RUN(bbbp_cast16(fmt, fparam(lp, L_OUT_W), F, nullptr, d.w_out, nullptr, Fq, F, F, Fq, s));
    RUN(bbbp_cast16(fmt, fparam(lp, L_L1_W), F, nullptr, d.w_l1, nullptr, Fq, kFF, F, Fq, s));
    RUN(bbbp_cast16(fmt, fparam(lp, L_L2_W), kFF, nullptr, d.w_l2, nullptr, kFF, F, kFF, kFF, s));
  }
  RUN(bbbp_cast16(fmt, fparam(params, P_FPFC_W), F, nullptr, p.fpfc_hi, p.fpfc_lo, Fq, kBranch, F, Fq, s));
  RUN(bbbp_conv3x3_prepare16(fmt, fparam(params, P_C1_W), p.conv1, 3, 32, s));
  RUN(bbbp_conv3x3_prepare16(fmt, fparam(params, P_C2_W), p.conv2, 32, 64, s));
  RUN(bbbp_fc_weight_to_hwc16(fmt, fparam(params, P_IFC_W), p.wfc, kBranch, 64, 32 * 32, s));
  if (m.split) {
    RUN(bbbp_fc_weight_channel_sums(fparam(params, P_C1_W), p.tap1, 32, 3, 9, s));
    RUN(bbbp_fc_weight_channel_sums(fparam(params, P_C2_W), p.tap2, 64, 32, 9, s));
    RUN(bbbp_fc_weight_channel_sums(fparam(params, P_IFC_W), p.possum, kBranch, 64, 32 * 32, s));
  }
  // fusion heads: first layers stacked into one (4*128, 256) operand, second layers into one block-diagonal (4, 4*128) one
  const int stacked = kFusionHeads * kFusionHidden;
  RUN(bbbp_fill_zero(p.fus2_hi, 1, kFusionHeads * stacked * 2ll, kFusionHeads * stacked * 2ll, s));
  if (p.fus2_lo) RUN(bbbp_fill_zero(p.fus2_lo, 1, kFusionHeads * stacked * 2ll, kFusionHeads * stacked * 2ll, s));
  for (int h = 0; h < kFusionHeads; ++h) {
    const void* const* hp = params + P_FUSION + 4 * h;
    const size_t row0 = (size_t)h * kFusionHidden;
    RUN(bbbp_cast16(fmt, fparam(hp, 0), kFused, nullptr, at16(p.fus1_hi, row0 * kFused),
                    p.fus1_lo ? at16(p.fus1_lo, row0 * kFused) : nullptr, kFused, kFusionHidden, kFused, kFused, s));
    RUN(bbbp_copy2d_f32(fparam(hp, 1), kFusionHidden, p.fus_b1 + row0, kFusionHidden, 1, kFusionHidden, s));
    RUN(bbbp_cast16(fmt, fparam(hp, 2), kFusionHidden, nullptr, at16(p.fus2_hi, (size_t)h * stacked + row0),
                    p.fus2_lo ? at16(p.fus2_lo, (size_t)h * stacked + row0) : nullptr, stacked, 1, kFusionHidden, kFusionHidden, s));
    RUN(bbbp_copy2d_f32(fparam(hp, 3), 1, p.fus_b2 + h, 1, 1, 1, s));
  }
  for (int i = 0; i < 4; ++i)
    RUN(bbbp_cast16(fmt, fparam(params, kHeadW[i]), kHeadK[i], nullptr, p.head_hi[i], p.head_lo[i], ceil8(kHeadK[i]), kHeadN[i],
                    kHeadK[i], ceil8(kHeadK[i]), s));
  return BBBP_OK;
}

extern "C" int bbbp_fwd(const bbbp_model_desc* desc, const void* fingerprint, const void* image, const void* const* params,
                        const void* prepared, float* out, void* workspace, size_t workspace_bytes, bbbp_stream_t s) {
  Dims m;
  RUN(parse(desc, &m, true));
  BBBP_CHECK_ARG(fingerprint && image && params && prepared && out && workspace, "bbbp_fwd: NULL argument");
  BBBP_CHECK_ARG(((reinterpret_cast<uintptr_t>(prepared) | reinterpret_cast<uintptr_t>(workspace)) & 255) == 0,
                 "bbbp_fwd: prepared and workspace must be 256-byte aligned");
  Arena pa(const_cast<void*>(prepared));
  Prepared p;
  lay_prepared(m, pa, &p);
  Arena wa(workspace);
  Work w;
  lay_workspace(m, wa, &w);
  if (wa.off > workspace_bytes) {
    bbbp::set_error("bbbp_fwd: workspace %zu bytes, need %zu (bbbp_workspace_bytes)", workspace_bytes, wa.off);
    return BBBP_EWORKSPACE;
  }
  const int F = m.F, Fq = m.Fq, fmt = m.fmt, R = (int)m.R, groups = m.groups, seq = m.seq;
  const float* fp = static_cast<const float*>(fingerprint);

  // ---- image branch first (20250113.py:85-93,114-115): the two long convolution kernels, then Linear(65536,128) ----------
  if (m.u8) RUN(bbbp_u8_image_stats_f32(static_cast<const uint8_t*>(image), w.stats, R, kImage, s));
  float* image_features = w.both + kBranch;      // cat((fingerprint, image), dim=1): both branches write their half in place
  if (m.split) {
    RUN(bbbp_image_background(image, m.u8, w.stats, w.bg1, R, kSide, kSide, s));
    RUN(bbbp_bg_layer(p.tap1, fparam(params, P_C1_B), w.bg1, 4, 3, 32, w.tab1, 64, w.tab1 + 32, 64, w.neg2, fmt, R, s));
    RUN(bbbp_bg_layer(p.tap2, fparam(params, P_C2_B), w.tab1 + 32, 64, 32, 64, w.tab2, 128, w.tab2 + 64, 128, nullptr, -1, R, s));
    RUN(bbbp_bg_layer(p.possum, nullptr, w.tab2 + 64, 128, 64, kBranch, w.fc_add, kBranch, nullptr, 0, nullptr, -1, R, s));
    // raw uint8 depictions take the exact-integer one-pass form, standardised fp32 planes a (hi, lo) pair
    RUN(bbbp_conv1_from_image_bg16(fmt, m.u8 ? 1 : 2, image, m.u8, w.stats, p.conv1, w.bg1, w.tab1, w.y1, R, kSide, kSide, s));
    RUN(bbbp_conv3x3_relu_pool_bg16(fmt, w.y1, p.conv2, w.neg2, w.tab2, w.y2, R, 32, 64, 64, 64, s));
    RUN(gemm(m, R, kBranch, kFlat, w.y2, nullptr, kFlat, p.wfc, nullptr, kFlat, fparam(params, P_IFC_B), nullptr, 0, w.fc_add,
             kBranch, image_features, kFused, nullptr, nullptr, kBranch, BBBP_ACT_RELU, 32, w.splitk, w.splitk_bytes, s));
  } else {
    RUN(bbbp_conv1_from_image16(fmt, 1, image, m.u8, w.stats, p.conv1, fparam(params, P_C1_B), w.y1, nullptr, R, kSide, kSide, s));
    RUN(bbbp_conv3x3_relu_pool16(fmt, 1, w.y1, nullptr, p.conv2, fparam(params, P_C2_B), w.y2, nullptr, R, 32, 64, 64, 64, s));
    RUN(gemm(m, R, kBranch, kFlat, w.y2, nullptr, kFlat, p.wfc, nullptr, kFlat, fparam(params, P_IFC_B), nullptr, 0, nullptr, 0,
             image_features, kFused, nullptr, nullptr, kBranch, BBBP_ACT_RELU, 8, w.splitk, w.splitk_bytes, s));
  }

  // ---- encoder (20250113.py:75-78,110-111): attention ACROSS the molecules of each reference batch -----------------------
  const float* x32 = fp;
  int ldx = F;
  if (Fq != F) {       // fp32 activations keep pitch Fq so residual reads and LayerNorm stores are 128-bit although F is odd
    RUN(bbbp_copy2d_f32(fp, F, w.x32[0], Fq, R, F, s));
    x32 = w.x32[0];
    ldx = Fq;
  }
  RUN(bbbp_cast16(fmt, fp, F, nullptr, w.x16[0], nullptr, Fq, R, F, Fq, s));
  const void* x16 = w.x16[0];
  const float scale = (float)std::pow((double)F, -0.5);       // heads == 1: head_dim = F
  int cur = 0;
  for (int l = 0; l < kLayers; ++l) {
    const void* const* lp = params + l * L_COUNT;
    const Prepared::Layer& d = p.layer[l];
    RUN(gemm(m, R, 3 * Fq, F, x16, nullptr, Fq, d.w_in, nullptr, Fq, d.b_in, nullptr, 0, nullptr, 0, nullptr, 3 * Fq, w.qkv,
             nullptr, 3 * Fq, BBBP_ACT_NONE, 1, nullptr, 0, s));
    const void *q = w.qkv, *k = at16(w.qkv, Fq), *v = at16(w.qkv, 2 * Fq);
    bool tail_done = false;
    if (m.heads == 1) {
      RUN(bbbp_transpose_bf16(groups, seq, F, v, 3 * Fq, (long long)seq * 3 * Fq, w.vt, w.ldp, (long long)F * w.ldp, s));
      if (m.flash && F <= 176) {
        // streaming attention with out_proj + residual + norm1 fused into its tail (model.py: fused_attention_tail)
        const int mid = cur ^ 1;
        RUN(bbbp_attention_flash_proj_ln16(fmt, groups, seq, F, q, 3 * Fq, k, 3 * Fq, (long long)seq * 3 * Fq, w.vt, w.ldp,
                                           (long long)F * w.ldp, scale, d.w_out, Fq, fparam(lp, L_OUT_B), x32, ldx,
                                           fparam(lp, L_N1_W), fparam(lp, L_N1_B), kLnEps, w.x32[mid], Fq, w.x16[mid], Fq, s));
        tail_done = true;
      } else if (m.flash) {
        if (Fq > ceil16(F)) RUN(bbbp_fill_zero(at16(w.attn, ceil16(F)), R, (Fq - ceil16(F)) * 2ll, Fq * 2ll, s));
        RUN(bbbp_attention_flash16(fmt, groups, seq, F, q, 3 * Fq, k, 3 * Fq, (long long)seq * 3 * Fq, w.vt, w.ldp,
                                   (long long)F * w.ldp, scale, w.attn, Fq, (long long)seq * Fq, s));
      } else {
        RUN(bbbp_attention_scores_softmax16(fmt, groups, seq, F, q, 3 * Fq, k, 3 * Fq, (long long)seq * 3 * Fq, scale, w.probs,
                                            w.ldp, s));
        RUN(bbbp_gemm16_batched(fmt, groups, seq, F, seq, w.probs, w.ldp, (long long)seq * w.ldp, w.vt, w.ldp,
                                (long long)F * w.ldp, nullptr, F, (long long)seq * F, w.attn, Fq, (long long)seq * Fq, s));
      }
    } else {
      RUN(bbbp_attention_heads16(fmt, w.qkv, 3 * Fq, Fq, 2 * Fq, w.attn, Fq, groups, seq, m.heads, m.head_dim, s));
    }
    const int mid = cur ^ 1;
    if (!tail_done) {
      RUN(gemm(m, R, F, F, w.attn, nullptr, Fq, d.w_out, nullptr, Fq, fparam(lp, L_OUT_B), x32, ldx, nullptr, 0, w.sum32, Fq,
               nullptr, nullptr, Fq, BBBP_ACT_NONE, 1, nullptr, 0, s));
      RUN(bbbp_add_layernorm_fwd_pitched16(fmt, w.sum32, Fq, nullptr, 0, fparam(lp, L_N1_W), fparam(lp, L_N1_B), w.x32[mid], Fq,
                                           w.x16[mid], Fq, R, F, kLnEps, s));
    }
    if (F <= 176) {     // linear1 + ReLU + linear2 + residual + norm2 in one kernel (the hidden activation stays on the chip)
      RUN(bbbp_ffn_layernorm16(fmt, R, F, kFF, w.x16[mid], Fq, d.w_l1, Fq, fparam(lp, L_L1_B), d.w_l2, kFF, fparam(lp, L_L2_B),
                               w.x32[mid], Fq, fparam(lp, L_N2_W), fparam(lp, L_N2_B), kLnEps, w.x32[cur], Fq, w.x16[cur], Fq, s));
      x32 = w.x32[cur];
      ldx = Fq;
      x16 = w.x16[cur];
      continue;
    }
    RUN(gemm(m, R, kFF, F, w.x16[mid], nullptr, Fq, d.w_l1, nullptr, Fq, fparam(lp, L_L1_B), nullptr, 0, nullptr, 0, nullptr, kFF,
             w.hidden, nullptr, kFF, BBBP_ACT_RELU, 1, nullptr, 0, s));
    RUN(gemm(m, R, F, kFF, w.hidden, nullptr, kFF, d.w_l2, nullptr, kFF, fparam(lp, L_L2_B), w.x32[mid], Fq, nullptr, 0, w.sum32,
             Fq, nullptr, nullptr, Fq, BBBP_ACT_NONE, 1, nullptr, 0, s));
    RUN(bbbp_add_layernorm_fwd_pitched16(fmt, w.sum32, Fq, nullptr, 0, fparam(lp, L_N2_W), fparam(lp, L_N2_B), w.x32[cur], Fq,
                                         w.x16[cur], Fq, R, F, kLnEps, s));
    x32 = w.x32[cur];
    ldx = Fq;
    x16 = w.x16[cur];
  }
  // fingerprint_fc (20250113.py:79-82); the strict mode splits both operands (the encoder itself is insensitive)
  if (m.split) {
    RUN(bbbp_cast16(fmt, x32, ldx, nullptr, w.fp_hi, w.fp_lo, Fq, R, F, Fq, s));
    RUN(gemm(m, R, kBranch, F, w.fp_hi, w.fp_lo, Fq, p.fpfc_hi, p.fpfc_lo, Fq, fparam(params, P_FPFC_B), nullptr, 0, nullptr, 0,
             w.both, kFused, nullptr, nullptr, kBranch, BBBP_ACT_RELU, 1, nullptr, 0, s));
  } else {
    RUN(gemm(m, R, kBranch, F, x16, nullptr, Fq, p.fpfc_hi, nullptr, Fq, fparam(params, P_FPFC_B), nullptr, 0, nullptr, 0, w.both,
             kFused, nullptr, nullptr, kBranch, BBBP_ACT_RELU, 1, nullptr, 0, s));
  }

  // ---- MultiHeadAttentionFusion (20250113.py:48-65): two GEMMs for all four heads, softmax over the head axis ------------
  const int stacked = kFusionHeads * kFusionHidden;
  RUN(bbbp_cast16(fmt, w.both, kFused, nullptr, w.cat_hi, w.cat_lo, kFused, R, kFused, kFused, s));
  RUN(gemm(m, R, stacked, kFused, w.cat_hi, w.cat_lo, kFused, p.fus1_hi, p.fus1_lo, kFused, p.fus_b1, nullptr, 0, nullptr, 0,
           nullptr, stacked, w.fh_hi, w.fh_lo, stacked, BBBP_ACT_TANH, 1, nullptr, 0, s));
  RUN(gemm(m, R, kFusionHeads, stacked, w.fh_hi, w.fh_lo, stacked, p.fus2_hi, p.fus2_lo, stacked, p.fus_b2, nullptr, 0, nullptr, 0,
           w.scores, kFusionHeads, nullptr, nullptr, 8, BBBP_ACT_NONE, 1, nullptr, 0, s));
  RUN(bbbp_fusion_softmax_mix_fwd_f32(w.scores, w.both, w.fused, nullptr, R, kFusionHeads, kFused, s));

  // ---- head (20250113.py:98-107): Linear+ReLU, BatchNorm1d on its running statistics, three more Linear layers ----------
  RUN(linear(m, w, R, 256, 256, w.fused, kFused, p.head_hi[0], p.head_lo[0], fparam(params, P_H0_B), w.h0, 256, BBBP_ACT_RELU, s));
  RUN(bbbp_batchnorm_fwd_f32(w.h0, fparam(params, P_BN_W), fparam(params, P_BN_B), const_cast<float*>(fparam(params, P_BN_MEAN)),
                             const_cast<float*>(fparam(params, P_BN_VAR)), w.hbn, nullptr, nullptr, R, 256, 0, kBnMomentum, kBnEps, s));
  RUN(linear(m, w, R, 128, 256, w.hbn, 256, p.head_hi[1], p.head_lo[1], fparam(params, P_H3_B), w.h3, 128, BBBP_ACT_RELU, s));
  RUN(linear(m, w, R, 64, 128, w.h3, 128, p.head_hi[2], p.head_lo[2], fparam(params, P_H5_B), w.h5, 64, BBBP_ACT_RELU, s));
  RUN(linear(m, w, R, 1, 64, w.h5, 64, p.head_hi[3], p.head_lo[3], fparam(params, P_H7_B), out, 1, BBBP_ACT_NONE, s));
  return BBBP_OK;
}
