/*
 * bbbp_b200.h -- C ABI of the B200 (sm_100a) kernels behind the reference's
 * MixedInputModel hot path (forward / backward / batched inference / AdamW).
 *
 * Conventions:
 *   - extern "C", plain pointers and sizes, no torch types.  Every pointer is a DEVICE
 *     pointer on the current device unless a comment says "host".
 *   - The caller owns every buffer, including workspaces.  The library allocates nothing
 *     persistent and keeps no global state besides a thread-local error string.
 *   - All work is enqueued on the stream argument (a cudaStream_t passed as void*); no
 *     entry point synchronises.
 *   - Return 0 on success, a negative BBBP_E* code on failure; bbbp_last_error() gives
 *     the message.  Nothing throws or exits.
 *   - Matrices are row-major.  "ld*" is the row pitch in ELEMENTS.
 *
 * Reference = /root/reference (FengDushuo/BBBP-Multi-Modal-Deep-Ensemble-Framework).
 * "C:" below abbreviates Models/multi_input_data_regression_opt_transformer_cnn_20250113.py.
 */
#ifndef BBBP_B200_H
#define BBBP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBBP_ABI_VERSION 4

enum { BBBP_OK = 0, BBBP_EINVAL = -1, BBBP_ECUDA = -2, BBBP_EWORKSPACE = -3, BBBP_EUNSUPPORTED = -4 };

/* activation fused into a linear epilogue */
enum { BBBP_ACT_NONE = 0, BBBP_ACT_RELU = 1, BBBP_ACT_TANH = 2 };

/* arithmetic mode of the fused forward */
enum {
  BBBP_PREC_FP32 = 0, /* CUDA-core fp32 FMA everywhere (validation mode, 1e-4 class parity)       */
  BBBP_PREC_BF16 = 1, /* tcgen05 bf16 operands, fp32 TMEM accumulation, fp32 statistics (fastest)  */
  BBBP_PREC_F16 = 2,  /* tcgen05 fp16 operands, one pass (11-bit operand mantissa = a TF32 operand)   */
  BBBP_PREC_STRICT = 3 /* tcgen05 fp16 operands, background-referenced image branch, split small GEMMs:
                          |d logBB| <= 1e-3 vs the fp32 reference at trained output scale          */
};

/* 16-bit operand format of the tensor-core entry points (the *16 functions; the *_bf16 names are the fmt = BF16 forms) */
#define BBBP_FMT_BF16 0 /* 8-bit mantissa, fp32 range                                                              */
#define BBBP_FMT_F16 1  /* 11-bit mantissa = the precision of a TF32 operand; conversions saturate at +-65504       */

typedef void* bbbp_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------------- */
int bbbp_abi_version(void);
const char* bbbp_last_error(void);
/* 0 when the current device is compute capability 10.x, BBBP_EUNSUPPORTED otherwise. */
int bbbp_device_check(void);
/* Number of kernels this library has launched in this process (monotonic; bench.py reports the
 * difference over its timed region as "gpu_launches"). */
uint64_t bbbp_launch_count(void);

/* ---- dense layers: nn.Linear call sites C:80,92,53-55,99-106; encoder in_proj/out_proj/
 *      linear1/linear2 (torch.nn.TransformerEncoderLayer via C:75-78); PCA transform
 *      Models/..._transformer_cnn_opt.py:30-33 ------------------------------------------------ */

/* C[M,N] = act( opA(A)[M,K] * opB(B)[K,N] + bias[N] ) (+ C_in when accumulate != 0), fp32 CUDA cores.
 * transA == 0: A stored [M,K] (lda >= K); != 0: A stored [K,M] (lda >= M).  Same for B with [K,N]/[N,K].
 * nn.Linear forward is transA=0, transB=1 with B = weight[N,K].  bias may be NULL.
 * split_k > 1 needs workspace of split_k*M*N floats (deterministic two-pass reduction).
 * split_k == 0 is the latency mode of the training path: the library picks the kernel (a one-shot shared-memory kernel
 * when M <= 32, i.e. one reference training batch) and the K partition from (M, N, K); results are deterministic for a
 * given shape but the reduction order depends on M, so the inference path (bit-identical scores however batches are
 * grouped) passes an explicit split_k instead.  Workspace for split_k == 0: bbbp_gemm_f32_auto_workspace() bytes. */
size_t bbbp_gemm_f32_auto_workspace(int M, int N, int K);
int bbbp_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                  float* C, int ldc, const float* bias, int act, int accumulate, int split_k, float* workspace,
                  size_t workspace_bytes, bbbp_stream_t stream);

/* dst[r, 0:cols_pad) = bf16(src[r, 0:cols)) with zero fill of [cols, cols_pad); ld_dst in bf16 elements. */
int bbbp_cast_bf16(const float* src, int ld_src, void* dst_bf16, int ld_dst, int rows, int cols, int cols_pad,
                   bbbp_stream_t stream);

/* dst[r][0:row_bytes) = 0 for r < rows, rows pitch_bytes apart (any alignment; 128-bit stores when everything is 16-byte
 * aligned).  Pad columns of pitched operands, accumulators. */
int bbbp_fill_zero(void* dst, long long rows, long long row_bytes, long long pitch_bytes, bbbp_stream_t stream);
/* fp32 rows -> 16-bit rows in format fmt: dst_hi = rn(src), optional dst_lo = rn(src - dst_hi) (NULL to skip), zero fill of
 * [cols, cols_pad).  cols_pad and ld_dst multiples of 8, destinations 16-byte aligned.  A (hi, lo) pair carries ~2x the
 * mantissa of one 16-bit operand; the strict inference mode feeds both through the same weights. */
/* col_sub (may be NULL): a row vector subtracted from every row in fp32 before the split (PCA.transform's X - mean_). */
int bbbp_cast16(int fmt, const float* src, int ld_src, const float* col_sub, void* dst_hi, void* dst_lo, int ld_dst, int rows,
                int cols, int cols_pad, bbbp_stream_t stream);

/* tcgen05 / TMEM / TMA GEMM: out = act( A[M,K] * W[N,K]^T + bias ) (+ residual), bf16 operands, fp32
 * accumulate.  A and W are bf16 row-major with K contiguous; lda/ldw in elements, multiples of 8, 16-byte
 * aligned bases.  Columns >= K are never read (TMA zero fill), so K need not be padded.
 * out_f32 and/or out_bf16 may be NULL (at least one must be given).  residual (fp32, ld = ld_res) may be NULL.
 * split_k >= 1; split_k > 1 needs workspace of split_k*M_pad*N_pad floats where *_pad round up to 128.
 * bbbp_gemm_bf16_workspace() returns the bytes needed. */
size_t bbbp_gemm_bf16_workspace(int M, int N, int split_k);
int bbbp_gemm_bf16(int M, int N, int K, const void* A_bf16, int lda, const void* W_bf16, int ldw, const float* bias,
                   const float* residual, int ld_res, float* out_f32, int ld_out, void* out_bf16, int ld_out16,
                   int act, int split_k, void* workspace, size_t workspace_bytes, bbbp_stream_t stream);

/* General form: operands in format fmt; optional lo parts A_lo / W_lo (same pitch as the hi parts, NULL to skip) add one
 * MMA each per K step into the same TMEM accumulator: out = act((A_hi + A_lo) W_hi^T + A_hi W_lo^T + bias) (+ residual).
 * The 16-bit output can be emitted as a (hi, lo) pair for the next split GEMM (out16_lo may be NULL). */
int bbbp_gemm16(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi, const void* W_lo,
                int ldw, const float* bias, const float* residual, int ld_res, float* out_f32, int ld_out, void* out16_hi,
                void* out16_lo, int ld_out16, int act, int split_k, void* workspace, size_t workspace_bytes,
                bbbp_stream_t stream);
/* bbbp_gemm16 with an fp32 [M][N] addend applied BEFORE the activation (a per-row bias; pitch ld_pre >= N, NULL to skip):
 * out = act((A_hi + A_lo) W_hi^T + A_hi W_lo^T + bias + pre_add).  Used by the background-referenced strict mode, where
 * the Linear(65536,128) of the image branch (20250113.py:92) sees activations relative to a per-image background. */
int bbbp_gemm16_pre(int fmt, int M, int N, int K, const void* A_hi, const void* A_lo, int lda, const void* W_hi, const void* W_lo,
                    int ldw, const float* bias, const float* pre_add, int ld_pre, float* out_f32, int ld_out, void* out16_hi,
                    void* out16_lo, int ld_out16, int act, int split_k, void* workspace, size_t workspace_bytes,
                    bbbp_stream_t stream);
/* Operands read in place in their TRANSPOSED storage (the backward products of a Linear layer and the convolution weight
 * gradient over im2col rows need no transposed copies): trans_a != 0: A is stored [K][M] (M contiguous, lda >= M);
 * trans_w != 0: W is stored [K][N] (N contiguous, ldw >= N).  out = act(opA(A) opW(W)^T + bias).  UMMA MN-major operand
 * descriptors; split_k as bbbp_gemm_bf16. */
int bbbp_gemm16_tn(int fmt, int trans_a, int trans_w, int M, int N, int K, const void* A, int lda, const void* W, int ldw,
                   const float* bias, float* out_f32, int ld_out, void* out16, int ld_out16, int act, int split_k,
                   void* workspace, size_t workspace_bytes, bbbp_stream_t stream);
int bbbp_gemm16_batched(int fmt, int batches, int M, int N, int K, const void* A, int lda, long long a_batch_stride,
                        const void* W, int ldw, long long w_batch_stride, float* out_f32, int ld_out,
                        long long out_batch_stride, void* out16, int ld_out16, long long out16_batch_stride,
                        bbbp_stream_t stream);
int bbbp_attention_scores_softmax16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                                    long long group_stride, float scale, void* p_out, int ldp, bbbp_stream_t stream);
int bbbp_softmax_rows_scaled16(int fmt, const float* scores, long long ld_scores, void* p_out, long long ld_p, long long rows,
                               int cols, float scale, bbbp_stream_t stream);

/* Batched form (blockIdx.z = batch): out[b] = A[b][M,K] * W[b][N,K]^T, no bias/activation; batch strides in elements
 * (multiples of 8 for the operands).  Used for the attention P V product with W = V^T. */
int bbbp_gemm_bf16_batched(int batches, int M, int N, int K, const void* A_bf16, int lda, long long a_batch_stride,
                           const void* W_bf16, int ldw, long long w_batch_stride, float* out_f32, int ld_out,
                           long long out_batch_stride, void* out_bf16, int ld_out16, long long out16_batch_stride,
                           bbbp_stream_t stream);
/* Attention scores with the row softmax in the TMEM epilogue (single head, seq <= 256: one CTA owns whole rows):
 * P[g] = softmax_rows(scale * Q[g] K[g]^T) as bf16 [groups][seq][ldp], pad columns [seq, ldp) zero.
 * q/k: bf16, row pitches ldq/ldk, group g starts group_stride elements after group g-1 (nn.MultiheadAttention's
 * scaled_dot_product_attention over S = the reference batch). */
int bbbp_attention_scores_softmax_bf16(int groups, int seq, int head_dim, const void* q_bf16, int ldq, const void* k_bf16,
                                       int ldk, long long group_stride, float scale, void* p_bf16, int ldp,
                                       bbbp_stream_t stream);
/* Wide attention scopes (seq > 256): p[r, 0:cols) = softmax(scale * scores[r, 0:cols)) as bf16, pad columns zero.
 * scores come from bbbp_gemm_bf16_batched (fp32 out); one block per row. */
int bbbp_softmax_rows_scaled_bf16(const float* scores, long long ld_scores, void* p_bf16, long long ld_p, long long rows,
                                  int cols, float scale, bbbp_stream_t stream);
/* dst[b][c][r] = src[b][r][c] (bf16); rows r in [rows, ld_dst) of dst are zero filled */
int bbbp_transpose_bf16(int batches, int rows, int cols, const void* src, int ld_src, long long src_batch_stride, void* dst,
                        int ld_dst, long long dst_batch_stride, bbbp_stream_t stream);

/* ---- image branch: nn.Conv2d(k3,s1,p1) + ReLU + MaxPool2d(2) C:85-90 (and 20250107_network.py:133-141) */

/* y[N,Cout,H/2,W/2] = maxpool2(relu(conv3x3(x[N,Cin,H,W], w[Cout,Cin,3,3]) + b)), NCHW fp32, CUDA cores.
 * argmax (uint8 in 0..3 = 2*dh+dw of the first maximum in torch's scan order) may be NULL.
 * pool == 0 gives the plain convolution (+bias when b != NULL, no ReLU) at full resolution.
 * Cout must be a multiple of 32; H, W multiples of 16. */
int bbbp_conv3x3_f32(const float* x, const float* w, const float* b, float* y, uint8_t* argmax, int N, int Cin,
                     int Cout, int H, int W, int pool, bbbp_stream_t stream);
/* dpre[N,C,H,W] = scatter of dy[N,C,H/2,W/2] to the arg-max position, zero where y <= 0 (ReLU) */
int bbbp_relu_pool_bwd_f32(const float* dy, const float* y, const uint8_t* argmax, float* dpre, int N, int C, int H,
                           int W, bbbp_stream_t stream);
/* dw[Cout,Cin,3,3] and db[Cout] from dpre[N,Cout,H,W] and x[N,Cin,H,W]; deterministic two-pass (partials per image
 * group, 8-row strip and row slice, summed in a fixed order).  workspace: bbbp_conv3x3_wgrad_workspace() bytes. */
size_t bbbp_conv3x3_wgrad_workspace(int N, int Cin, int Cout, int H, int W);
int bbbp_conv3x3_wgrad_f32(const float* dpre, const float* x, float* dw, float* db, int N, int Cin, int Cout, int H,
                           int W, float* workspace, size_t workspace_bytes, bbbp_stream_t stream);
/* w_t[Cin,Cout,3,3] = spatially flipped, channel-transposed w[Cout,Cin,3,3] (weights of the data-gradient conv) */
int bbbp_conv3x3_flip_weights_f32(const float* w, float* w_t, int Cin, int Cout, bbbp_stream_t stream);

/* tcgen05 inference path of the same block: NHWC bf16 activations with 8*k channels per pixel, implicit GEMM with the
 * halo tile staged once in shared memory, bias + ReLU + 2x2 max-pool fused into the TMEM epilogue (conv_umma.cu).
 * Built for the reference's layers: (Cin 3 -> padded 8, Cout 32), (Cin 32, Cout 64) and -- planar fp32 input only -- (3, 64). */
size_t bbbp_conv3x3_prepared_bytes(int Cin, int Cout);
/* w[Cout,Cin,3,3] fp32 -> the kernel's bf16 shared-memory weight image (call again whenever w changes) */
int bbbp_conv3x3_prepare_bf16(const float* w, void* wprep, int Cin, int Cout, bbbp_stream_t stream);
/* y[N,H/2,W/2,Cout] = maxpool2(relu(conv3x3(x[N,H,W,Cin_pad]) + bias)), bf16 NHWC in and out; H % 32 == 0, W % 16 == 0 */
int bbbp_conv3x3_relu_pool_bf16(const void* x_nhwc, const void* wprep, const float* bias, void* y_nhwc, int N,
                                int Cin_pad, int Cout, int H, int W, bbbp_stream_t stream);
/* First layer straight from the reference's input contract: img_chw is (N, 3, H, W) planar, either fp32 (already
 * standardised, contract P2) or -- img_is_u8 != 0 -- raw uint8 depictions normalised on the fly as
 * (u/255 - mean) * rstd with stats[n] = {mean, rstd} from bbbp_u8_image_stats_f32 (ToTensor + per-molecule z-score).
 * The producers pack 3 channels to one bf16 chunk per pixel in registers; output as bbbp_conv3x3_relu_pool_bf16. */
int bbbp_conv1_from_image_bf16(const void* img_chw, int img_is_u8, const float* stats, const void* wprep,
                               const float* bias, void* y_nhwc, int N, int H, int W, bbbp_stream_t stream);
/* General forms.  fmt: operand / output format.  split = 2 is the strict mode: the input is a (hi, lo) pair -- x_lo, or
 * both parts formed by the first layer's producers from the fp32 / uint8 image -- each part staged as its own ring slot and
 * multiplied against the same once-rounded weights into the same accumulators, and the output is emitted as a (hi, lo)
 * pair (y_lo).  Built: (BF16, 1), (F16, 1), (F16, 2). */
int bbbp_conv3x3_prepare16(int fmt, const float* w, void* wprep, int Cin, int Cout, bbbp_stream_t stream);
int bbbp_conv3x3_relu_pool16(int fmt, int split, const void* x_nhwc, const void* x_lo, const void* wprep, const float* bias,
                             void* y_nhwc, void* y_lo, int N, int Cin_pad, int Cout, int H, int W, bbbp_stream_t stream);
int bbbp_conv1_from_image16(int fmt, int split, const void* img_chw, int img_is_u8, const float* stats, const void* wprep,
                            const float* bias, void* y_nhwc, void* y_lo, int N, int H, int W, bbbp_stream_t stream);
int bbbp_fc_weight_to_hwc16(int fmt, const float* w, void* out16, int rows, int C, int HW, bbbp_stream_t stream);
/* First block of the big variant (Conv2d(3, 64) + ReLU + MaxPool2d(2), 20250107_network.py:133-135) on the same fused kernel:
 * fp32 planar (N, 3, H, W) in, bf16 NHWC (N, H/2, W/2, 64) out; wprep from bbbp_conv3x3_prepare_bf16(w, wprep, 3, 64). */
int bbbp_conv1_from_image_c64_bf16(const float* img_chw, const void* wprep, const float* bias, void* y_nhwc, int N, int H, int W,
                                   bbbp_stream_t stream);
/* Background-referenced strict mode of the same two blocks (conv_umma.cu, BG = 1).  A depiction is
 * mostly one value per channel, so each layer works on activations RELATIVE to a per-image background (exactly 0 on the
 * canvas: rounding them to fp16 costs nothing there) and its epilogue adds back, in fp32 and from the fp32 weights, what the
 * constant part contributes:  conv(x)[p] = conv(x - bg)[p] + T[image][cout],  T = bias + sum over all taps of w . bg  (the
 * zero padding of x is staged as -bg, so this holds at the image border too).
 *   bbbp_image_background   bg[n][0..2] = background value of image n per channel (majority of 8 probe pixels); uint8 input:
 *                           normalised with stats exactly as the first layer's producers do, and bg[n][3] carries the raw
 *                           background bytes r | g << 8 | b << 16 as a bit pattern (fp32 input: 0)
 *   bbbp_fc_weight_channel_sums  out[o][c] = sum_j w[o][c*HW + j]: a Linear over a flattened (C, HW) activation, or (HW = 9)
 *                           a 3x3 convolution, applied to a per-channel constant
 *   bbbp_bg_layer           one step of the background chain: out0[n][co] = bias[co] + sum_ci wsum[co][ci] * in[n][ci]
 *                           (pitches ld_in / ld0; bias may be NULL); out1 (may be NULL) = relu(out0), rounded to the 16-bit
 *                           format fmt when fmt >= 0 (-1: no rounding) = the background of the layer's OUTPUT; neg16 (may be
 *                           NULL) = -out1 in that format, [n][Cout] = what the next layer's padding holds.  Cout 32..256
 *   bbbp_conv1_from_image_bg16 / bbbp_conv3x3_relu_pool_bg16: the layers.  tab[n] = {T[Cout], bg_out[Cout]} (out0 / out1 of
 *                           bbbp_bg_layer written with pitch 2*Cout); bg_in = bbbp_image_background's table; neg_bg_in = the
 *                           previous layer's neg16.  split = 2: the first layer's producers stage x - bg as a (hi, lo) pair.
 *                           split = 1 on a uint8 image is the EXACT-INTEGER form: the producers stage u - background byte
 *                           (an integer, exact in fp16; the padding's fraction rides in the pixel's spare channel) and the
 *                           epilogue scales the accumulator by rstd / 255 -- one pass, no activation rounding at all.
 *                           Output: ONE fp16 NHWC tensor holding maxpool(relu(conv(x))) - bg_out.  Built for fp16. */
int bbbp_image_background(const void* img_chw, int img_is_u8, const float* stats, float* bg, int N, int H, int W,
                          bbbp_stream_t stream);
int bbbp_fc_weight_channel_sums(const float* w, float* out, int rows, int C, int HW, bbbp_stream_t stream);
int bbbp_bg_layer(const float* wsum, const float* bias, const float* in, int ld_in, int Cin, int Cout, float* out0, int ld0,
                  float* out1, int ld1, void* neg16, int fmt, int N, bbbp_stream_t stream);
int bbbp_conv1_from_image_bg16(int fmt, int split, const void* img_chw, int img_is_u8, const float* stats, const void* wprep,
                               const float* bg_in, const float* tab, void* y_nhwc, int N, int H, int W, bbbp_stream_t stream);
int bbbp_conv3x3_relu_pool_bg16(int fmt, const void* x_nhwc, const void* wprep, const void* neg_bg_in, const float* tab,
                                void* y_nhwc, int N, int Cin_pad, int Cout, int H, int W, bbbp_stream_t stream);
/* Lossless sparse depictions (extension, sparse_depictions.cu): rebuilds out[n][3][128][128] uint8 from mask[n][2048]
 * (bit p, little-endian, set where pixel p is not white), values (the RGB triples of the marked pixels in scan order) and
 * offsets[n + 1] (running pixel counts; only differences to offsets[0] are used).  ~5.5 KB per molecule instead of 49 152. */
int bbbp_decode_sparse_depictions_u8(const uint8_t* mask, const uint8_t* values, const int64_t* offsets, uint8_t* out, int n,
                                     bbbp_stream_t stream);
/* stats[r] = {mean, 1/std} of img[r, 0:n] / 255 (population std, 0 -> 1), exact integer sums, fp64 finish */
int bbbp_u8_image_stats_f32(const uint8_t* img, float* stats, int rows, int n, bbbp_stream_t stream);
/* Diagnostics: probe = DEVICE array of 16 uint64 cycle counters that CTA 0 of the tcgen05 conv kernels accumulates
 * per pipeline role (see conv_umma.cu), or NULL to switch the probe off (default). */
int bbbp_debug_conv_probe(void* probe);
/* Diagnostics: TMEM -> register read rate of one SM (tcgen05.ld 32x32b.x32 in a loop, `warps` = 4 or 8 reading warps).
 * out: DEVICE uint64[3] = {cycles, bytes read, checksum}.  The pooled convolutions read four pre-pool accumulators per
 * output value, so this rate is the first layer's floor. */
int bbbp_debug_tmem_read_probe(void* out, int warps, int reps, bbbp_stream_t stream);
/* fp32 NCHW image with C <= 8 planes (the reference's (B,3*128*128) input viewed as (B,3,128,128), 20250113.py:114)
 * -> bf16 NHWC with 8 channels per pixel, channels >= C zero */
int bbbp_image_to_nhwc8_bf16(const float* img_nchw, void* out_nhwc8, int N, int C, int H, int W, bbbp_stream_t stream);
/* out[o][(hw)*C + c] = bf16(w[o][c*HW + hw]): nn.Linear weight over nn.Flatten's (C,H,W) order re-laid for an NHWC
 * activation (20250113.py:91-92) */
int bbbp_fc_weight_to_hwc_bf16(const float* w, void* out_bf16, int rows, int C, int HW, bbbp_stream_t stream);

/* Generic tensor-core route for channel counts the implicit-GEMM kernel is not instantiated for (the 64/128/256-channel
 * stack of 20250107_network.py:133-141): explicit bf16 im2col -> bbbp_gemm_bf16 (bias + ReLU epilogue) -> 2x2 max-pool.
 * out[(n*H + y)*W + x][tap*C + c] = x[n][y+dy][x+dx][c], zero outside the image, tap = 3*(dy+1) + (dx+1); C % 8 == 0. */
int bbbp_im2col3x3_bf16(const void* x_nhwc, void* out, int N, int H, int W, int C, bbbp_stream_t stream);
/* The same convolution WITHOUT the im2col matrix (channel counts that are multiples of 64: the 64 -> 128 and 128 -> 256 layers
 * of 20250107_network.py:136-141): y[N,H,W,Cout] = act(conv3x3(x[N,H,W,C]) + bias), 16-bit NHWC in and out, as an implicit GEMM
 * on the tcgen05 GEMM kernel -- a GEMM row is a pixel, a K block is one (tap, 64-channel block), and its A tile is fetched as
 * ONE shifted 4-D TMA box of the activation (the zero padding is the TMA's out-of-bounds fill), so the activation is read from
 * L2 nine times instead of a 9x larger matrix being written to and read from HBM.  w_taps: [Cout][9*C] in (tap, channel)
 * order (bbbp_conv3x3_weight_im2col16).  W must divide 128, H*W % 128 == 0, N*H*W/128 <= 65535. */
int bbbp_conv3x3_gemm16(int fmt, const void* x_nhwc, int N, int H, int W, int C, const void* w_taps, int Cout, const float* bias,
                        int act, void* y_nhwc, bbbp_stream_t stream);
/* y[N,H/2,W/2,C] = 2x2 max-pool of x[N,H,W,C], bf16 NHWC, C % 8 == 0 */
int bbbp_maxpool2x2_nhwc_bf16(const void* x_nhwc, void* y_nhwc, int N, int H, int W, int C, bbbp_stream_t stream);
/* out[Cout][9*Cpad] bf16 with out[co][tap*Cpad + c] = w[co][c][tap] (zero for c >= Cin): the W operand matching the
 * im2col row order */
int bbbp_conv3x3_weight_im2col_bf16(const float* w, void* out_bf16, int Cin, int Cpad, int Cout, bbbp_stream_t stream);

/* Mixed-precision TRAINING path of the image branch (conv_train.cu): every contraction runs on the tcgen05 GEMM -- the
 * convolutions as im2col rows x weights (bbbp_im2col3x3_bf16 is format-agnostic), their weight gradients as
 * dpre^T x im2col rows through bbbp_gemm16_tn (both operands MN-major, no transposes), their data gradient as
 * im2col(dpre) x flipped weights -- and these are the memory-bound pieces in between, NHWC, 16-bit format fmt. */
/* y[N,H/2,W/2,C] = 2x2 max-pool of x[N,H,W,C]; argmax[N,H/2,W/2,C] (uint8) = 2*i + j of the first maximum (torch order) */
int bbbp_maxpool2x2_argmax_nhwc16(int fmt, const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C,
                                  bbbp_stream_t stream);
/* dpre[N,H,W,C] (16-bit) = dy[N,H/2,W/2,C] (fp32) routed to the arg-max member where y > 0 (ReLU), zeros elsewhere;
 * dy_masked (fp32, may be NULL) = dy * (y > 0): its column sum is the bias gradient */
int bbbp_unpool_relu_nhwc16(int fmt, const float* dy, const void* y, const uint8_t* argmax, void* dpre, float* dy_masked, int N,
                            int H, int W, int C, bbbp_stream_t stream);
int bbbp_image_to_nhwc8_16(int fmt, const float* img_nchw, void* out_nhwc8, int N, int C, int H, int W, bbbp_stream_t stream);
/* out[Cout][9*Cpad] with out[co][tap*Cpad + c] = w[co][c][tap]; and its inverse for the fp32 gradient */
int bbbp_conv3x3_weight_im2col16(int fmt, const float* w, void* out16, int Cin, int Cpad, int Cout, bbbp_stream_t stream);
int bbbp_conv3x3_wgrad_from_im2col_f32(const float* g, float* dw, int Cin, int Cpad, int Cout, bbbp_stream_t stream);
/* out[CinPad][9*Cout] with out[ci][tap*Cout + co] = w[co][ci][8 - tap]: the W operand of the data-gradient GEMM */
int bbbp_conv3x3_weight_dgrad16(int fmt, const float* w, void* out16, int Cin, int CinPad, int Cout, bbbp_stream_t stream);
/* dw[o][c*HW + hw] = g[o][hw*C + c]: Linear weight gradient computed over the NHWC flattening -> nn.Flatten's order */
int bbbp_fc_grad_hwc_to_chw_f32(const float* g, float* dw, int rows, int C, int HW, bbbp_stream_t stream);

/* ---- encoder self-attention across the molecules of a reference batch: the (B,1,F)
 *      input of C:110-111 is read by nn.TransformerEncoder as seq_len = B, batch = 1.  ``groups``
 *      independent reference batches of ``seq`` molecules each are processed in one launch. ------ */

/* qkv[groups*seq, 3E] = [q | k | v], E = heads*head_dim (nn.MultiheadAttention in_proj layout).
 * out[groups*seq, E]; lse[groups*seq, heads] (log-sum-exp of the scaled scores, may be NULL). */
/* dropout_p > 0 drops attention probabilities after the softmax (nn.MultiheadAttention dropout, train mode);
 * the keep mask is Philox(seed; group, head, query, key), so backward regenerates it from the same seed. */
/* seed_dev (may be NULL): DEVICE uint64 added to ``seed`` when the kernel RUNS, so a training step captured in a CUDA
 * graph draws a fresh mask at every replay (the host refreshes *seed_dev before replaying; see bbbp_adamw_dev_f32). */
int bbbp_attention_fwd_f32(const float* qkv, float* out, float* lse, int groups, int seq, int heads, int head_dim,
                           float dropout_p, uint64_t seed, const uint64_t* seed_dev, bbbp_stream_t stream);
int bbbp_attention_bwd_f32(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv,
                           int groups, int seq, int heads, int head_dim, float dropout_p, uint64_t seed,
                           const uint64_t* seed_dev, bbbp_stream_t stream);

/* Row softmax of the GEMM route the training path takes for mid-size single-head scopes (seq > 32, e.g. batch 256: the six
 * products Q K^T, P V, dO V^T, P^T dO, dS K, dS^T Q run on bbbp_gemm_f32).  p = softmax(scale * scores) row-wise;
 * p_dropped = p * keep with the Philox keep mask of (seed (+ *seed_dev); row_base + r, row_base + c) -- only written
 * when dropout_p > 0.  Backward: dscores = scale * p * (dp_dropped * keep - sum_c dp_dropped * keep * p). */
int bbbp_attn_softmax_fwd_f32(const float* scores, int ld_scores, float* p, float* p_dropped, int ld_p, int rows, int cols,
                              float scale, float dropout_p, uint64_t seed, const uint64_t* seed_dev, long long row_base,
                              bbbp_stream_t stream);
int bbbp_attn_softmax_bwd_f32(const float* p, const float* dp_dropped, float* dscores, int ld, int rows, int cols, float scale,
                              float dropout_p, uint64_t seed, const uint64_t* seed_dev, long long row_base, bbbp_stream_t stream);

/* Many small heads on the bf16 inference path (the 2048-bit fingerprint variants: 256 heads of dimension 8, C:71-73):
 * out[r, h*D + c] = softmax(q_h k_h^T / sqrt(D)) v_h per group, warp-level mma.sync m16n8k16 with an online softmax.
 * qkv rows (bf16, pitch ld) hold q at column 0, k at column k_offset, v at column v_offset (all multiples of 8);
 * head_dim 8 or 16; any seq. */
int bbbp_attention_heads_bf16(const void* qkv_bf16, int ld, int k_offset, int v_offset, void* out_bf16, int ld_out, int groups,
                              int seq, int heads, int head_dim, bbbp_stream_t stream);

/* Streaming-softmax attention for ONE head (head_dim <= 192) and ANY scope length on tcgen05 / TMEM / TMA: a CTA owns 128
 * queries and walks the keys in blocks of 128 (S = Q K^T -> TMEM, P = exp2(S - m) -> shared memory, O += P V -> TMEM) with a
 * lazily moved running maximum; the seq x seq logits never reach HBM (at seq = 65 536 they would be 17 GB per layer).
 * q / k: 16-bit rows (pitches ldq / ldk, group g starts group_stride elements after g - 1); v_t: the TRANSPOSED values
 * [groups][head_dim][ld_vt] (bbbp_transpose_bf16), ld_vt >= seq; out: [groups*seq][ld_out] 16-bit, columns >= head_dim of a
 * row are written as zeros up to ld_out.  All pitches / strides multiples of 8 elements. */
int bbbp_attention_flash16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                           long long group_stride, const void* v_t, int ld_vt, long long vt_group_stride, float scale, void* out,
                           int ld_out, long long out_group_stride, bbbp_stream_t stream);
/* The same kernel with the REST of the attention half of a post-norm encoder layer fused into its tail (out_proj, + bias,
 * + residual, norm1: nn.TransformerEncoderLayer via 20250113.py:75-78):
 *   y[g*seq + r] = LayerNorm(residual[g*seq + r] + softmax(scale Q K^T) V  W_out^T + b_out) * gamma + beta
 * After the last P V product the CTA writes O / l as a 16-bit operand into the dead Q tiles, fetches W_out (head_dim x head_dim,
 * 16-bit, pitch ldw) into the dead K stages, runs one more product into the dead S columns of TMEM and normalises whole rows
 * out of TMEM; the attention output never reaches HBM and three launches become one.  residual / y32: fp32 rows (pitches
 * multiples of 4); y16 (may be NULL): 16-bit copy, columns [head_dim, ld_y16) zero, ld_y16 <= ceil16(head_dim).  head_dim <= 176. */
int bbbp_attention_flash_proj_ln16(int fmt, int groups, int seq, int head_dim, const void* q, int ldq, const void* k, int ldk,
                                   long long group_stride, const void* v_t, int ld_vt, long long vt_group_stride, float scale,
                                   const void* w_out16, int ldw, const float* b_out, const float* residual, int ld_res,
                                   const float* gamma, const float* beta, float eps, float* y32, int ld_y, void* y16, int ld_y16,
                                   bbbp_stream_t stream);
int bbbp_attention_heads16(int fmt, const void* qkv, int ld, int k_offset, int v_offset, void* out, int ld_out, int groups,
                           int seq, int heads, int head_dim, bbbp_stream_t stream);

/* The feed-forward half of a post-norm nn.TransformerEncoderLayer (linear1 -> ReLU -> linear2 -> + x -> norm2, via
 * 20250113.py:75-78) in ONE tcgen05 kernel (ffn_fused_umma.cu):
 *   y = LayerNorm(residual + relu(x16 W1^T + b1) W2^T + b2) * gamma + beta
 * The (rows x hidden) activation never reaches HBM: a CTA owns 128 rows, walks the hidden units in blocks of 128 (first
 * product -> TMEM, bias + ReLU -> 16-bit tile in shared memory, second product accumulates in TMEM) and normalises whole rows
 * straight out of TMEM.  x16 (rows x d, pitch ldx), W1 (hidden x d, pitch ldw1), W2 (d x hidden, pitch ldw2): 16-bit in format
 * fmt; residual / y32: fp32 rows (pitches multiples of 4); y16 (may be NULL): the 16-bit copy of y, columns [d, ld_y16) zero.
 * d <= 176, hidden % 128 == 0, ld_y16 <= ceil16(d). */
int bbbp_ffn_layernorm16(int fmt, int rows, int d, int hidden, const void* x16, int ldx, const void* w1_16, int ldw1,
                         const float* b1, const void* w2_16, int ldw2, const float* b2, const float* residual, int ld_res,
                         const float* gamma, const float* beta, float eps, float* y32, int ld_y, void* y16, int ld_y16,
                         bbbp_stream_t stream);

/* out[r, c] = (x[r, c] - mean_c) / std_c with the statistics of column c taken over the rows of r's CHUNK (chunk_rows
 * consecutive rows; the last chunk may be shorter): StandardScaler().fit_transform per block of 100 molecules,
 * Descriptors/multi_input_data_preprocess_maccs_opt_IsolationForest_fixed_1.py:86-101.  float64 statistics, sklearn's
 * near-constant rule (scale 1), float32 transform (x - float32(mean)) / float32(std) as sklearn 1.9.  In place allowed. */
int bbbp_standardize_chunks_f32(const float* x, long long ld_x, float* out, long long ld_out, long long rows, int cols,
                                int chunk_rows, bbbp_stream_t stream);

/* ---- normalisation: nn.LayerNorm (post-norm residual, eps 1e-5) and nn.BatchNorm1d C:101 ------- */

/* s = x + res (res may be NULL); y = LN(s)*gamma + beta.  Optional outputs: sum_out (= s), mean[rows],
 * rstd[rows], y_bf16 (pitch ld_bf16, zero filled up to it). */
int bbbp_add_layernorm_fwd_f32(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                               float* sum_out, float* mean, float* rstd, void* y_bf16, int ld_bf16, int rows,
                               int dim, float eps, bbbp_stream_t stream);
/* Inference form with independent row pitches (elements) for x, res and y, so activations can keep a 16-byte-aligned
 * pitch when dim is not a multiple of 4 (F = 167): y = LN(x + res)*gamma + beta, optional bf16 copy. */
int bbbp_add_layernorm_fwd_pitched_f32(const float* x, int ld_x, const float* res, int ld_res, const float* gamma,
                                       const float* beta, float* y, int ld_y, void* y_bf16, int ld_bf16, int rows, int dim,
                                       float eps, bbbp_stream_t stream);
int bbbp_add_layernorm_fwd_pitched16(int fmt, const float* x, int ld_x, const float* res, int ld_res, const float* gamma,
                                     const float* beta, float* y, int ld_y, void* y16, int ld16, int rows, int dim, float eps,
                                     bbbp_stream_t stream);
/* dx (= gradient wrt s, which is also the gradient of both x and res), dgamma[dim], dbeta[dim].
 * workspace: bbbp_layernorm_bwd_workspace() bytes (partials of dgamma / dbeta per row chunk, summed in a fixed order). */
size_t bbbp_layernorm_bwd_workspace(int rows, int dim);
int bbbp_layernorm_bwd_f32(const float* dy, const float* s, const float* mean, const float* rstd, const float* gamma,
                           float* dx, float* dgamma, float* dbeta, int rows, int dim, float* workspace,
                           size_t workspace_bytes, bbbp_stream_t stream);
/* training != 0: batch statistics, running stats updated in place (momentum, unbiased variance), save_mean /
 * save_rstd written.  training == 0: running statistics, save_* untouched (may be NULL). */
int bbbp_batchnorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float* y, float* save_mean, float* save_rstd, int rows, int channels,
                           int training, float momentum, float eps, bbbp_stream_t stream);
int bbbp_batchnorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* save_mean,
                           const float* save_rstd, float* dx, float* dgamma, float* dbeta, int rows, int channels,
                           bbbp_stream_t stream);
/* eval-mode backward (running statistics are constants): dx = dy * gamma * rsqrt(var + eps) */
int bbbp_batchnorm_eval_bwd_f32(const float* dy, const float* x, const float* gamma, const float* running_mean,
                                const float* running_var, float* dx, float* dgamma, float* dbeta, int rows,
                                int channels, float eps, bbbp_stream_t stream);

/* ---- fusion blocks C:48-65, _rdkit.py:53-66, 20250107_network.py:51-105 ------------------------ */

/* w = softmax over the last axis of scores[rows, n] (n <= 32); out[rows, dim] = sum_h w[r,h] * c[r, :]
 * accumulated in head order.  w_out[rows, n] optional. */
int bbbp_fusion_softmax_mix_fwd_f32(const float* scores, const float* c, float* out, float* w_out, int rows, int n,
                                    int dim, bbbp_stream_t stream);
/* dc[rows, dim] and dscores[rows, n] from dout */
int bbbp_fusion_softmax_mix_bwd_f32(const float* w, const float* c, const float* dout, float* dc, float* dscores,
                                    int rows, int n, int dim, bbbp_stream_t stream);
/* w[rows, n] = softmax over the last axis of scores[rows, n], n <= 32 (nn.Softmax(dim=1) of the fusion blocks) */
int bbbp_softmax_rows_fwd_f32(const float* scores, float* w, int rows, int n, bbbp_stream_t stream);
/* dscores = w * (dw - sum_h dw_h w_h) */
int bbbp_softmax_rows_bwd_f32(const float* w, const float* dw, float* dscores, int rows, int n, bbbp_stream_t stream);
/* out[r, c] = scale[r*ld_scale] * colmean(x)[c]  (the (B,1,1)*(B,D) -> mean(dim=1) broadcast of the big variant);
 * colmean_out[cols] optional. */
int bbbp_scaled_colmean_fwd_f32(const float* x, int ldx, const float* scale, int ld_scale, float* out, int ld_out,
                                float* colmean_out, int rows, int cols, bbbp_stream_t stream);
/* dscale[r] = <dout[r,:], colmean>;  dx[r', c] = (1/rows) * sum_r scale[r] * dout[r, c]  (same for every r') */
int bbbp_scaled_colmean_bwd_f32(const float* dout, int ld_dout, const float* scale, int ld_scale, const float* colmean,
                                float* dx, int ldx, float* dscale, int rows, int cols, bbbp_stream_t stream);

/* ---- elementwise / reductions ------------------------------------------------------------------ */
/* dx = dy * act'(y) computed from the activation OUTPUT y (relu: y > 0; tanh: 1 - y^2); pitches in elements */
int bbbp_act_bwd_f32(const float* dy, int ld_dy, const float* y, int ld_y, float* dx, int ld_dx, int rows, int cols,
                     int act, bbbp_stream_t stream);
/* y[i] = x[i] * scalar[0] with the scalar read from DEVICE memory (chain-rule factor of a loss, no host sync) */
int bbbp_scale_by_device_scalar_f32(const float* x, const float* scalar, float* y, size_t n, bbbp_stream_t stream);
/* out[c] = sum_r x[r, c] (bias gradients), deterministic.  More than 2048 rows are reduced in two passes over row chunks and
 * need bbbp_colsum_workspace() bytes of workspace (0 below that; workspace may then be NULL). */
size_t bbbp_colsum_workspace(int rows, int cols);
int bbbp_colsum_f32(const float* x, int ldx, float* out, int rows, int cols, float* workspace, size_t workspace_bytes,
                    bbbp_stream_t stream);
/* dst[r, 0:cols) = src[r, 0:cols) with independent pitches (concatenation / slicing) */
int bbbp_copy2d_f32(const float* src, int ld_src, float* dst, int ld_dst, int rows, int cols, bbbp_stream_t stream);
/* dst[r, 0:cols) = src[idx[r], 0:cols): the device-resident batch feeder that replaces MixedDataset.__getitem__ + the
 * DataLoader collate (C:31-45, 165-168); idx is a DEVICE int64 array of `rows` dataset indices. */
int bbbp_gather_rows_f32(const float* src, const int64_t* idx, float* dst, int rows, long long cols, bbbp_stream_t stream);
/* dst[idx[i]] = src[i], i < n (idx: DEVICE int64, unique): a cross-validation fold's scores written to their dataset
 * positions, the device form of ``nn_predictions[test_idx] = nn_fold_predictions`` (C:237). */
int bbbp_scatter_f32(const float* src, const int64_t* idx, float* dst, size_t n, bbbp_stream_t stream);
/* Philox-4x32-10 dropout: y = x * keep / (1-p); the mask is a function of (seed, element index).  Not
 * stream-compatible with torch's generator (SURVEY hard part f). */
int bbbp_dropout_f32(const float* x, float* y, size_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* seed_dev,
                     bbbp_stream_t stream);

/* ---- loss C:143,189 and optimiser C:172,191 ----------------------------------------------------- */
/* loss[0] = mean((pred - target)^2); dpred = 2 (pred - target) / n * grad_scale (dpred may be NULL) */
int bbbp_mse_loss_f32(const float* pred, const float* target, float* loss, float* dpred, int n, float grad_scale,
                      bbbp_stream_t stream);
/* extension (no reference code): mean BCE-with-logits, dlogit = (sigmoid(z) - t) / n * grad_scale */
int bbbp_bce_logits_loss_f32(const float* logit, const float* target, float* loss, float* dlogit, int n,
                             float grad_scale, bbbp_stream_t stream);
/* torch.optim.AdamW semantics, one launch for all tensors.  ptrs is a DEVICE array of 4*ntensors pointers
 * laid out [param | grad | exp_avg | exp_avg_sq] and sizes a DEVICE array of ntensors element counts;
 * chunk_tensor / chunk_offset (DEVICE, nchunks entries) enumerate chunks of bbbp_adamw_chunk() elements (one CTA each).
 * step is 1-based.
 * Hyper-parameters are doubles so that 1-beta, lr/bias_correction etc. round exactly as torch's do. */
int bbbp_adamw_chunk(void);
int bbbp_adamw_f32(void* const* ptrs, const int64_t* sizes, const int32_t* chunk_tensor, const int64_t* chunk_offset,
                   int ntensors, int nchunks, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int step, float grad_scale, bbbp_stream_t stream);
/* The same update with the per-step scalars read from DEVICE memory when the kernel runs, so that a whole training step
 * (forward + backward + this launch) can be captured once in a CUDA graph and replayed while the step count, the
 * learning rate (torch LR schedulers) and the dropout seed keep changing.  hyper_dev = 8 floats produced by
 * bbbp_adamw_hyper(); the host copies them to the device before each replay. */
int bbbp_adamw_dev_f32(void* const* ptrs, const int64_t* sizes, const int32_t* chunk_tensor, const int64_t* chunk_offset,
                       int ntensors, int nchunks, const float* hyper_dev, bbbp_stream_t stream);
/* dst_dev[0:n_bytes) = host_src[0:n_bytes), n_bytes <= 64, travelling as KERNEL PARAMETERS: host_src is read before the
 * call returns, so no pinned staging buffer has to outlive the call and the store is ordered on the stream like any
 * launch.  This is how the per-step scalars (hyper_dev, seed_dev) are refreshed between CUDA-graph replays. */
int bbbp_store_small(const void* host_src, int n_bytes, void* dst_dev, bbbp_stream_t stream);
/* HOST helper (no device work): hyper_out[8] (host) = {1-beta1, beta2, 1-beta2, eps, 1-lr*wd, lr/(1-beta1^step),
 * sqrt(1-beta2^step), grad_scale}, each formed in double exactly as bbbp_adamw_f32 forms it. */
int bbbp_adamw_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, int step, float grad_scale,
                     float* hyper_out);

/* ---- input contracts P1/P2 (extensions; oracle = oracle/preprocess.py) --------------------------- */
/* packed little-endian bit rows (bytes_per_row = ceil(n_bits/8)) -> per-molecule z-scored fp32 rows */
int bbbp_unpack_zscore_f32(const uint8_t* packed, int bytes_per_row, float* out, int ld_out, int rows, int n_bits,
                           bbbp_stream_t stream);
/* uint8 depictions (rows x n values) -> x/255 -> per-molecule z-score, fp32 */
int bbbp_u8_zscore_f32(const uint8_t* img, float* out, int rows, int n, bbbp_stream_t stream);

/* ---- whole-model inference (bbbp_fwd / bbbp_workspace_bytes) ------------------------------------------
 * MixedInputModel.forward of Models/multi_input_data_regression_opt_transformer_cnn_20250113.py:109-119 in eval mode as
 * ONE call for a host written in C / C++ / anything with an FFI: the same kernel sequence bbbp_b200.model runs from Python
 * (bit-identical scores), orchestrated by the library.  Nothing is allocated: the caller passes
 *   params    HOST array of bbbp_model_param_count() DEVICE pointers, fp32, in the order of the reference's
 *             model.state_dict() (bbbp_model_param_name(i) is the key, bbbp_model_param_numel(i) the element count;
 *             fc.2.num_batches_tracked is listed but never read and may be NULL),
 *   prepared  bbbp_model_prepared_bytes() bytes, filled by bbbp_model_prepare() from the parameters (16-bit copies, the conv
 *             kernels' shared-memory weight images, the (H,W,C) re-laid Linear(65536,128) weight, stacked fusion heads,
 *             tap / position sums for the strict mode); call it again whenever a parameter changes,
 *   workspace bbbp_workspace_bytes() bytes of scratch (activations; contents undefined between calls),
 * all 256-byte aligned.  A call scores desc->groups independent reference batches of desc->seq molecules each (attention
 * runs ACROSS the molecules of a batch) and writes out[groups*seq] fp32.  Everything is enqueued on `stream`
 * (graph-capturable: no synchronisation, no allocation, no host reads of device memory).
 * Built: variant BBBP_MODEL_TCNN_20250113, precisions BF16 / F16 / STRICT; fingerprint sizes whose encoder has one head
 * (F <= 192, e.g. MACCS-167; any seq) or heads of dimension 8 / 16 (e.g. Morgan-2048: 256 x 8).  The fp32 validation mode
 * and training stay behind the per-kernel entry points above (orchestrated by bbbp_b200.autograd). */
enum { BBBP_MODEL_TCNN_20250113 = 0 };
typedef struct bbbp_model_desc {
  int abi_version;      /* BBBP_ABI_VERSION the caller was compiled against */
  int variant;          /* BBBP_MODEL_TCNN_20250113 */
  int fingerprint_size; /* F (constructor argument fingerprint_size, 20250113.py:69) */
  int precision;        /* BBBP_PREC_BF16 | BBBP_PREC_F16 | BBBP_PREC_STRICT */
  int groups;           /* reference batches per call */
  int seq;              /* molecules per reference batch (the reference's batch_size) */
  int image_is_u8;      /* 0: fp32 standardised (B, 3*128*128) CHW rows (contract P2); 1: raw uint8 CHW depictions */
} bbbp_model_desc;
int bbbp_model_param_count(const bbbp_model_desc* desc);
const char* bbbp_model_param_name(const bbbp_model_desc* desc, int index);
size_t bbbp_model_param_numel(const bbbp_model_desc* desc, int index);
size_t bbbp_model_prepared_bytes(const bbbp_model_desc* desc);
int bbbp_model_prepare(const bbbp_model_desc* desc, const void* const* params, void* prepared, size_t prepared_bytes,
                       bbbp_stream_t stream);
size_t bbbp_workspace_bytes(const bbbp_model_desc* desc);
int bbbp_fwd(const bbbp_model_desc* desc, const void* fingerprint, const void* image, const void* const* params,
             const void* prepared, float* out, void* workspace, size_t workspace_bytes, bbbp_stream_t stream);

/* ---- multi-GPU: the two exchanges of the path over NCCL ------------------------------------------
 * comm is an ncclComm_t (one per process / GPU).  NCCL is resolved at first use with dlopen("libnccl.so.2") -- the copy
 * already loaded in the process (e.g. torch's) when there is one -- so the library has no link-time dependency on it;
 * BBBP_EUNSUPPORTED when it cannot be found.  id128 is a HOST buffer of 128 bytes (ncclUniqueId).
 *   bbbp_comm_gather_scores      screening: every rank contributes count fp32 scores, all ranks receive
 *                                recv[rank*count .. ) (ncclAllGather; the only collective of the screening path)
 *   bbbp_comm_average_gradients  data-parallel training: the flat fp32 gradient buffer averaged in place over the ranks
 *                                (ncclAllReduce, ncclAvg), 20250113.py:190-191 run as replicas */
typedef void* bbbp_comm_t;
int bbbp_comm_unique_id(void* id128);
int bbbp_comm_init_rank(bbbp_comm_t* comm, int nranks, const void* id128, int rank);
int bbbp_comm_destroy(bbbp_comm_t comm);
int bbbp_comm_gather_scores(bbbp_comm_t comm, const float* send, float* recv, size_t count, bbbp_stream_t stream);
int bbbp_comm_average_gradients(bbbp_comm_t comm, float* flat_grads, size_t count, bbbp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BBBP_B200_H */
