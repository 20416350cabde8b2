// tcgen05 implicit-GEMM 3x3 convolution with the bias + ReLU + 2x2 max-pool epilogue fused in, bf16 operands,
// fp32 accumulation in TMEM.  Replaces nn.Conv2d(3,32,3,p1)+ReLU+MaxPool2d(2) and nn.Conv2d(32,64,3,p1)+ReLU+
// MaxPool2d(2) of the reference's image branch (20250113.py:85-90) on the inference path; conv2 is 73 % of the
// network's FLOPs.
//
// Data layout: activations are NHWC bf16 with C = 8*KC (conv1: 3 real channels zero-padded to 8; conv2: 32).
// One CTA tile = 128 POOLED output pixels (8 wide x 16 high) x COUT channels; the four members of every 2x2 pooling
// window are four separate TMEM accumulators (same lane = same pooled pixel), so pooling is a per-thread max in the
// epilogue and the pre-pool activation never exists outside TMEM.
//
// A operand without im2col: the (16+2) x (32+2) input halo of the tile is staged ONCE in shared memory in the
// no-swizzle K-major core-matrix layout  [kc][x parity][y][x/2][8 channels = 16 B].  For window member (dy,dx) and
// filter tap (kh,kw), GEMM row m = (ph,pw) reads input pixel (2ph+dy+kh, 2pw+dx+kw) of the halo: consecutive pw are
// consecutive 16-byte core-matrix rows, consecutive ph are 2 halo rows apart (= the descriptor's stride byte offset),
// and a K=16 step pairs two 8-channel chunks through the leading byte offset.  So all 9 taps x 4 window members are
// just different START ADDRESSES into the same staged halo -- the tile is read from L2 once, not 9 times.
//
// Warp roles (416 threads): warps 0-7 epilogue (TMEM lane quadrant = warp % 4, channel half = warp / 4), warps 8-11
// producers (cp.async 16-byte chunks with zero fill = the conv's zero padding), warp 12 TMEM allocation +
// single-thread MMA issue.  Persistent
// over tiles: 3-stage halo ring (full/empty mbarriers) and a double-buffered accumulator (acc_full/acc_empty), so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand format and split passes (half16.cuh): FMT selects bf16 or fp16 operands (weights, staged halo, output).  SPLIT = 2
// is the strict mode: the input activation arrives as a hi + lo pair (two NHWC tensors, or both parts formed by the planar
// producers from the fp32 / uint8 image), each part is staged as its own ring slot and multiplied against the same
// once-rounded weights into the same TMEM accumulators (2x the MMAs), and the epilogue emits its output as a hi + lo pair
// for the next layer.
#include "common.cuh"
#include "umma.cuh"
#include "half16.cuh"
#include <utility>
#include <type_traits>

namespace bbbp {
namespace conv {
using namespace sm100;

#ifndef BBBP_CONV1_PF
#define BBBP_CONV1_PF 3
#endif
#ifndef BBBP_CONV1_PF_MERGED
#define BBBP_CONV1_PF_MERGED 2     // prefetch depth of the (hi, lo) first-layer kernels (their conversion code needs more registers)
#endif
#ifndef BBBP_CONV1_PROD_WARPS
#define BBBP_CONV1_PROD_WARPS 8
#endif
#ifndef BBBP_CONV1_CS
#define BBBP_CONV1_CS 8
#endif
constexpr int TILE_PW = 8, TILE_PH = 16;       // pooled tile
constexpr int HALO_W = 2 * TILE_PW + 2;        // 18
constexpr int HALO_H = 2 * TILE_PH + 2;        // 34
constexpr int XH = HALO_W / 2;                 // 9 halo columns per parity
constexpr int ROW_B = XH * 16;                 // 144 bytes per (parity, y) row
// plane pitches carry a few pad bytes so that the producers' 16-byte cp.async stores of one quarter-warp
// (x parity alternates, then the 8-channel chunk index) land in eight different 16-byte bank groups
constexpr int PAR_B = HALO_H * ROW_B + 32;     // 4928 = 64 (mod 128)
constexpr int KC_B = 2 * PAR_B + 16;           // 9872 = 16 (mod 128) bytes per 8-channel chunk plane

template <int KC, int COUT, int SRC = 0, int FMT = BBBP_FMT_BF16, int SPLIT = 1, int BG = 0>
struct Cfg {
  // conv2's epilogue (64 channels) gets two warps per TMEM lane quadrant; conv1's (32 channels) one, which also keeps
  // its CTA small enough for two CTAs per SM
  static constexpr bool PACK4 = SRC != 0;
  static constexpr int STAGES = 3;                      // halo ring depth
  static constexpr int EPI_WARPS = COUT >= 64 ? 8 : 4;
  // the producers are latency-bound (global loads / cp.async behind a shared-memory pipe the tensor core keeps busy):
  // eight warps halve the per-thread chunk count
  // (measured: conv2 1.21 -> 1.14 ms).  conv1 keeps four: with eight, two CTAs per SM need a 72-register cap that spills
  // in the epilogue and costs more than the producers gain.
  static constexpr int PROD_WARPS = (KC == 1 && !PACK4) ? 4 : (PACK4 ? BBBP_CONV1_PROD_WARPS : 8), PROD_THREADS = PROD_WARPS * 32;
  // + one MMA warp + one STORE warp.  The store warp exists because ISSUING the tile's bulk tensor store costs its thread
  // ~1 200 cycles (cycle probe: 1 205 on conv1, 1 362 on conv2 -- the 128 swizzled rows of the box are walked at issue),
  // and when epilogue thread 0 paid that, every epilogue warp paid it too at the next per-tile barrier: a third of
  // conv1's tile time.  Now the epilogue warps only ARRIVE on the "tile written" barrier and go on to the next tile.
  static constexpr int EPI_THREADS = EPI_WARPS * 32, THREADS = EPI_THREADS + PROD_THREADS + 64;
  static constexpr int PROD_WARP0 = EPI_WARPS, MMA_WARP = EPI_WARPS + PROD_WARPS, STORE_WARP = MMA_WARP + 1;
  // the first layer runs two small CTAs per SM (measured: one CTA with 8 + 8 + 1 warps and a 6-deep ring is slower,
  // 1.65 ms vs 1.42 ms per 8 192 images)
  // (64 output channels need all 512 TMEM columns for the double-buffered accumulators: one CTA per SM)
  static constexpr int MIN_CTAS = (KC == 1 && COUT <= 32) ? 2 : 1;
  static constexpr int A_BYTES = KC * KC_B;
  // planar first-layer sources in two parts: BOTH parts of a tile's halo share one ring slot (hi image, then lo image), so
  // the producers pay one wait / fence / arrive round per tile instead of two (they, not the MMAs, bound that kernel)
  static constexpr bool MERGED = PACK4 && SPLIT == 2;
  static constexpr int STAGE_BYTES = (MERGED ? 2 : 1) * A_BYTES;
  static constexpr int NMMA = KC == 1 ? 5 : 9 * (KC / 2);  // K=16 steps per window member
  // instructions per tile: conv1 issues one N=COUT MMA per (member, step); conv2 pairs the two members of a pooling
  // row that read the SAME halo view (dx=0 with tap kw+1, dx=1 with tap kw) into one N=2*COUT MMA: 24 per K-chunk pair
  // PACK4 (planar first-layer sources): 4 channels (8 bytes) per pixel, two pixels per 16-byte chunk, so one K=16 MMA
  // covers the four pixels X .. X+3 of a halo row with X = 2*pw.  Both window members of a pooling row need three of
  // them (dx = 0: X..X+2, dx = 1: X+1..X+3), so ONE N = 2*COUT MMA per (dy, kernel row) computes both -- their weights
  // sit side by side in the B image with a zero tap at the unused pixel: 6 MMAs per tile instead of 12 (an M = 128 MMA
  // costs the tensor pipe ~90 cycles whatever its N, and that issue cost is what bounded the first layer).
  static constexpr int NISSUE = PACK4 ? 6 : KC == 1 ? 4 * NMMA : 24 * (KC / 2);
  static constexpr int W8_BYTES = 2 * 5 * 2 * COUT * 16;                  // conv1 weight image for the NHWC8 source
  static constexpr int W_BYTES = PACK4 ? 3 * 2 * (2 * COUT) * 16 : KC == 1 ? W8_BYTES : 9 * KC * COUT * 16;
  // the prepared blob of the first layer: NHWC8 image | PACK4 hi | PACK4 lo | PACK4 with the channel sums in the spare channel
  static constexpr int W_OFFSET = PACK4 ? (BG == 2 ? W8_BYTES + 2 * (3 * 2 * (2 * COUT) * 16) : W8_BYTES) : 0;
  static constexpr int ACC_COLS = 4 * COUT;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int OUT_ROW_B = COUT * 2;            // bytes of one pooled pixel (= the TMA store's swizzle span)
  static constexpr int OUT_BYTES = 128 * OUT_ROW_B;     // one output tile: 128 pooled pixels x COUT bf16
  // BG (background-referenced activations, see the kernel comment): the input may still arrive as a (hi, lo) pair, the
  // output is ONE 16-bit tensor (values relative to the layer's per-image background response)
  static constexpr int OSPLIT = BG ? 1 : SPLIT;
  static constexpr int OUT_BUFS = 2 * OSPLIT;           // double-buffered hi (and lo) output tiles
  // strict mode, first layer: the WEIGHTS are split as well (w = hi + lo, the lo image follows the hi image in the prepared
  // blob and in shared memory) -- K is tiny there, so the third pass costs little, and the first layer's weight rounding
  // is the largest remaining term of the strict mode's error budget (tests/precision_study.py: 6e-4 of 8e-4)
  static constexpr int W_PARTS = (SPLIT == 2 && PACK4 && !BG) ? 2 : 1;
  // BG: the current and the next tile's per-image table rows (T and bg_out), double-buffered
  // (BG == 2: + the image's (mean, 1/std) pair, the epilogue's scale, in a 16-byte slot of its own)
  static constexpr int TAB_FLOATS = 2 * COUT, TAB_SLOT = TAB_FLOATS + (BG == 2 ? 4 : 0), TAB_BYTES = BG ? 2 * TAB_SLOT * 4 : 0;
  static constexpr int SMEM_BYTES = 1024 + OUT_BUFS * OUT_BYTES + STAGES * STAGE_BYTES + W_PARTS * W_BYTES + COUT * 4 + BAR_BYTES + TAB_BYTES;
};

__host__ __device__ constexpr int halo_offset(int dy, int dx, int tap) {
  const int kh = tap / 3, kw = tap % 3, s = dx + kw;
  return (s & 1) * PAR_B + (dy + kh) * ROW_B + (s >> 1) * 16;
}
// conv1 (one 8-channel chunk per pixel): a K=16 MMA step covers TWO taps, the second reached through the leading
// byte offset.  Pair i = taps (2i, 2i+1); the 9th tap is paired with tap 7 under zero weights.  The descriptor
// offset must be positive, so the pair is ordered by halo address, which depends on dx only.
struct TapPair {
  int first, second, zero_slot;  // zero_slot: which K chunk carries zero weights (-1 none)
};
__host__ __device__ constexpr TapPair conv1_pair(int dx, int i) {
  const bool pad = 2 * i + 1 >= 9;
  const int t0 = 2 * i, t1 = pad ? 7 : 2 * i + 1;
  if (halo_offset(0, dx, t1) > halo_offset(0, dx, t0)) return TapPair{t0, t1, pad ? 1 : -1};
  return TapPair{t1, t0, pad ? 0 : -1};
}

// t / d and t % d for a run-time d without the ~40-instruction integer-division sequence: the role loops below locate
// every tile (image, tile row, tile column) on latency-bound single warps.  q = umulhi(t, ceil(2^32 / d)) is exact for
// t < 2^32 / d (the launcher checks the tile count).
struct FastDiv {
  uint32_t d, m;
  __device__ explicit FastDiv(int dd) : d((uint32_t)dd), m(dd == 1 ? 0u : 0xFFFFFFFFu / (uint32_t)dd + 1u) {}
  __device__ __forceinline__ int div(int t) const { return m ? (int)__umulhi((uint32_t)t, m) : t; }
  __device__ __forceinline__ void divmod(int t, int& q, int& r) const {
    q = div(t);
    r = t - q * (int)d;
  }
};

struct MmaOp {
  uint32_t a_off, a_lbo, b_off, b_lbo, d_col, n, accumulate;
};
// Operands of the I-th MMA of a tile: byte offsets into the staged halo / the shared-memory weight image, the
// accumulator column, the instruction's N and its accumulate flag.  constexpr: the issue loop is fully unrolled and
// every descriptor is an immediate.
template <int KC, int COUT, bool PACK4>
__host__ __device__ constexpr MmaOp mma_op(int I) {
  if (PACK4) {
    // I = 2*kh + dy: kernel row outermost, the pooling row innermost -- consecutive MMAs accumulate into DIFFERENT TMEM
    // accumulators (columns [2*dy*COUT, +2*COUT) = members (dy, 0) | (dy, 1)) and pipeline in the tensor core.
    // A: halo row dy + kh, K chunk 0 = pixels (X, X+1), chunk 1 = (X+2, X+3) through the leading byte offset.
    const int dy = I % 2, kh = I / 2;
    return MmaOp{(uint32_t)((dy + kh) * ROW_B), 16u, (uint32_t)(kh * (2 * (2 * COUT) * 16)), (uint32_t)((2 * COUT) * 16),
                 (uint32_t)(dy * 2 * COUT), (uint32_t)(2 * COUT), (uint32_t)(kh != 0)};
  }
  if (KC == 1) {
    // conv1: member q = 2*dy + dx, step m covers the tap pair (2m, 2m+1) through the leading byte offset
    const int q = I % 4, m = I / 4, dy = q >> 1, dx = q & 1;   // members innermost (independent accumulators back to back)
    const TapPair p = conv1_pair(dx, m);
    const int oa = halo_offset(dy, dx, p.first), ob = halo_offset(dy, dx, p.second);
    return MmaOp{(uint32_t)oa, (uint32_t)(ob - oa), (uint32_t)((dx * 5 + m) * (2 * COUT * 16)), (uint32_t)(COUT * 16),
                 (uint32_t)(q * COUT), (uint32_t)COUT, (uint32_t)(m != 0)};
  }
  // conv2: K-chunk pair j, pooling row dy, then 12 (kh, sx) halo views; sx = dx + kw is the view's x shift.
  // sx = 1, 2: both members of the row use this view (dx=0 with kw=sx, dx=1 with kw=sx-1): one N = 2*COUT MMA whose
  // B rows are the two taps' weights, adjacent in the [kc][8 - tap][cout] weight image.  sx = 0 / 3: one member only.
  // The first view of every row is a paired one with accumulate = 0, so it initialises both accumulators.
  // issue order: K-chunk pair, view, then the pooling row innermost, so back-to-back MMAs alternate between the two
  // rows' accumulators (independent) instead of chaining on one accumulator
  const int j = I / 24, rem = I % 24, dy = rem % 2, e = rem / 2;
  const int kh = e < 4 ? 0 : (e - 4) / 4 + 1;
  const int order0[4] = {1, 2, 0, 3};
  const int sx = e < 4 ? order0[e] : (e - 4) % 4;
  const bool paired = sx == 1 || sx == 2;
  const int dx = sx == 3 ? 1 : 0;                     // member that owns columns d_col (the left one when paired)
  const int tap = kh * 3 + (sx == 3 ? 2 : sx);        // tap of that member: kw = sx - dx
  const int a_off = (2 * j) * KC_B + (sx & 1) * PAR_B + (dy + kh) * ROW_B + (sx >> 1) * 16;
  const int b_off = ((2 * j) * 9 + (8 - tap)) * (COUT * 16);
  return MmaOp{(uint32_t)a_off, (uint32_t)KC_B, (uint32_t)b_off, (uint32_t)(9 * COUT * 16), (uint32_t)((2 * dy + dx) * COUT),
               (uint32_t)(paired ? 2 * COUT : COUT), (uint32_t)!(j == 0 && e == 0)};
}

// descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version 1 (bit 46) | no swizzle
template <int KC, int COUT, bool PACK4, int FMT, bool FORCE_ACC, int I>
__device__ __forceinline__ void issue_one(uint32_t a_lo, uint32_t w_lo, uint32_t tmem_acc) {
  constexpr MmaOp op = mma_op<KC, COUT, PACK4>(I);
  constexpr uint32_t a_lo_c = (op.a_off >> 4) | ((op.a_lbo >> 4) << 16);
  constexpr uint32_t b_lo_c = (op.b_off >> 4) | ((op.b_lbo >> 4) << 16);
  constexpr uint64_t a_hi = (uint64_t)(((2 * ROW_B) >> 4) | (1u << 14)) << 32;
  constexpr uint64_t b_hi = (uint64_t)((128 >> 4) | (1u << 14)) << 32;
  constexpr uint32_t idesc = make_idesc_16(128, op.n, FMT);
  umma_bf16(tmem_acc + op.d_col, a_hi | (a_lo + a_lo_c), b_hi | (w_lo + b_lo_c), idesc, FORCE_ACC || op.accumulate != 0);
}
// FORCE_ACC: the lo part of a split tile adds into the accumulators the hi part has just initialised
template <int KC, int COUT, bool PACK4, int FMT, bool FORCE_ACC, int... I>
__device__ __forceinline__ void issue_tile(uint32_t a_lo, uint32_t w_lo, uint32_t tmem_acc,
                                           std::integer_sequence<int, I...>) {
  (issue_one<KC, COUT, PACK4, FMT, FORCE_ACC, I>(a_lo, w_lo, tmem_acc), ...);
}

enum { SRC_NHWC_BF16 = 0, SRC_CHW_F32 = 1, SRC_CHW_U8 = 2 };


// Optional cycle probe (bbbp_debug_conv_probe): when set, CTA 0 accumulates clock64() deltas of its role loops into
// probe[0..15]: MMA thread {wait acc_empty, wait full, issue}, epilogue thread 0 {wait acc_full, tmem+math+sts, barriers+
// store}, producer thread 0 {wait empty, fill}.  One predictable branch per tile when unset.
__device__ unsigned long long* g_conv_probe = nullptr;
#define PROBE_T0() const long long _t0 = probe ? clock64() : 0
#define PROBE_ADD(slot, t0) do { if (probe) { const long long _n = clock64(); probe[slot] += (unsigned long long)(_n - (t0)); (t0) = _n; } } while (0)

// SRC selects what the producers read: NHWC bf16 activations (cp.async), or -- first layer only -- the reference's
// own input contract, planar fp32 CHW (20250113.py:114), or raw uint8 CHW depictions normalised on the fly with a
// per-image (mean, 1/std) pair (ToTensor + per-molecule StandardScaler, Descriptors/..._preprocess_maccs_opt.py:52-67,
// 121-124).  In both planar cases the 3 channels are packed to one 16-byte bf16 chunk per pixel in registers, so the
// NHWC8 image never exists in HBM.
//
// BG = 1, background-referenced activations (the strict mode since round 2): a depiction is mostly ONE
// value per channel (the white canvas), and so is every activation map computed from it away from the strokes.  The layer
// therefore works on  x' = x - bg  (bg = this image's background per input channel: x' is exactly 0 on the canvas, so nothing
// is lost there when x' is rounded to 16 bits), and the epilogue adds back what the constant part contributes, in fp32 and
// from the fp32 weights:   conv(x)[p] = conv(x')[p] + T[co],   T[co] = bias[co] + sum over ALL nine taps of w[co][ci][tap] * bg[ci].
// For that to hold at the image border the zero padding of x must read as -bg in the shifted space: the first layer's
// producers stage it as such (hi + lo, exact); later layers receive a background that IS a 16-bit number (the previous
// layer subtracts bg_out = rn16(relu(T)), which leaves a residue of at most half an ulp of bg on the canvas -- itself
// rounded with negligible error), so their producers write the exact -bg into the out-of-image halo chunks instead of
// letting cp.async zero-fill them.  bg_tab[image] = {T[COUT], bg_out[COUT]} (bbbp_bg_layer); after ReLU + pooling the
// epilogue subtracts bg_out before rounding.  bg_in: first layer float[image][4]; later layers the NEGATED background of the
// input in the operand format, [image][8 * KC].
//
// BG = 2, raw uint8 depictions only: the EXACT-INTEGER form of the same idea, one pass.  With x = scale*u + shift (scale =
// rstd/255, shift = -mean*rstd) and r[c] the background's raw byte per channel, the producers stage d = u - r[c]: an integer
// of magnitude <= 255, exact in fp16, 0 on the canvas -- there is no activation rounding at all, so no (hi, lo) pair is needed.
// x = scale*d + bg[c] with bg[c] = scale*r[c] + shift, hence  conv(x) = scale * conv(d) + (T - bias)  and the epilogue forms
// relu(scale * max(acc) + T) - bg_out (scale > 0 commutes with the max).  The zero padding of x is d_pad = 255*mean - r[c]:
// its integer part floor(255*mean) - r[c] goes into the channel, the fraction in [0, 1) into the SPARE fourth channel of the
// pixel, whose weights are the filter's sum over its input channels (the third PACK4 weight image) -- the fraction is the same
// for all three channels.  bg_in[image][3] carries r as packed bytes (bbbp_image_background).
template <int KC, int COUT, int SRC, int FMT, int SPLIT, int BG>
__global__ void __launch_bounds__(Cfg<KC, COUT, SRC, FMT, SPLIT, BG>::THREADS, Cfg<KC, COUT, SRC, FMT, SPLIT, BG>::MIN_CTAS)
conv3x3_umma_kernel(const void* __restrict__ src_any, const void* __restrict__ src_lo_any, const float2* __restrict__ stats,
                    const uint4* __restrict__ wprep, const float* __restrict__ bias, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmOutLo, int n_img, int H, int W, const void* __restrict__ bg_in,
                    const float* __restrict__ bg_tab) {
  using C = Cfg<KC, COUT, SRC, FMT, SPLIT, BG>;
  static_assert(SRC == SRC_NHWC_BF16 || KC == 1, "planar sources feed the 3-channel first layer only");
  static_assert(BG != 2 || (SRC == SRC_CHW_U8 && SPLIT == 1), "the exact-integer form reads raw uint8 depictions in one pass");
    constexpr int THREADS = C::THREADS, EPI_THREADS = C::EPI_THREADS, PROD_WARP0 = C::PROD_WARP0, MMA_WARP = C::MMA_WARP;
  constexpr int PROD_THREADS = C::PROD_THREADS, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = base;                     // 2 (x2 with a lo part) swizzled output tiles (1024-byte aligned) for the TMA stores
  uint8_t* sA = sOut + C::OUT_BUFS * C::OUT_BYTES;
  uint8_t* sW = sA + STAGES * C::STAGE_BYTES;
  float* sBias = reinterpret_cast<float*>(sW + C::W_PARTS * C::W_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + COUT);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  [[maybe_unused]] float* sTab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + C::BAR_BYTES);   // 16-byte aligned

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int tiles_x = (W / 2) / TILE_PW, tiles_y = (H / 2) / TILE_PH;
  const int tiles_per_img = tiles_x * tiles_y;
  const int num_tiles = n_img * tiles_per_img;
  const FastDiv fd_img(tiles_per_img), fd_x(tiles_x);
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // ---- one-time setup: weights + bias to smem, barriers, TMEM -----------------------------------------------------
  for (int i = threadIdx.x; i < C::W_PARTS * C::W_BYTES / 16; i += THREADS)
    reinterpret_cast<uint4*>(sW)[i] = wprep[C::W_OFFSET / 16 + i];
  if constexpr (C::PACK4)   // pad bytes of the stage are never written by the producers: keep them finite
    for (int i = threadIdx.x; i < STAGES * C::STAGE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < COUT; i += THREADS) sBias[i] = BG ? 0.0f : bias[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], PROD_THREADS);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, C::TMEM_COLS);
  fence_proxy_async_smem();  // the weight stores above are generic-proxy writes read by tcgen05.mma (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long* probe = (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == PROD_WARP0 || warp == MMA_WARP)) ? g_conv_probe : nullptr;

  if (warp >= PROD_WARP0 && warp < MMA_WARP) {
    // ===== producers: stage the halo of each tile ====================================================================
    const int ptid = threadIdx.x - PROD_WARP0 * 32;
    constexpr int ROWC = HALO_W * KC;                  // 16-byte chunks per halo row
    constexpr int CHUNKS = HALO_H * ROWC;
    constexpr int STEP_Y = PROD_THREADS / ROWC, STEP_J = PROD_THREADS % ROWC;
    constexpr int PIX_B = KC * 16;
    const int Yi = ptid / ROWC, ji = ptid % ROWC;      // first chunk of this thread: the same for every tile
    if constexpr (SRC == SRC_NHWC_BF16) {
      // a ring slot holds one PART of one tile: unit u = tile * SPLIT + part (part 1 = the lo tensor of a split input)
      const int my_units = my_tiles * SPLIT;
      for (int i = 0; i < my_units; ++i) {
        const int t = blockIdx.x + (i / SPLIT) * gridDim.x;
        int n, r, ty, tx;
        fd_img.divmod(t, n, r);
        fd_x.divmod(r, ty, tx);
        const int y0 = 2 * ty * TILE_PH - 1, x0 = 2 * tx * TILE_PW - 1;
        const int s = i % STAGES;
        long long tp = probe ? clock64() : 0;
        mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
        PROBE_ADD(8, tp);
        const uint32_t stage = smem_u32(sA + s * C::STAGE_BYTES);
        const uint8_t* img = reinterpret_cast<const uint8_t*>((SPLIT == 2 && (i & 1)) ? src_lo_any : src_any) +
                             (size_t)n * H * W * PIX_B;
        // BG: what an out-of-image chunk must hold, -bg of this thread's 8-channel chunk (kc = ji % KC for every chunk of
        // the thread: the chunk stride PROD_THREADS and the row length are multiples of KC); copied like any other chunk
        [[maybe_unused]] const uint8_t* padsrc = nullptr;    // this image's -bg row, chunk kc
        if constexpr (BG) {
          padsrc = static_cast<const uint8_t*>(bg_in) + (size_t)n * PIX_B + (ji % KC) * 16;
          static_assert(PROD_THREADS % KC == 0 && ROWC % KC == 0, "a producer thread keeps one channel chunk");
        }
        int Y = Yi, j = ji;
  #pragma unroll 4
        for (int c = ptid; c < CHUNKS; c += PROD_THREADS) {
          const int X = j / KC, kc = j % KC;
          const int y = y0 + Y, x = x0 + X;
          const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
          const int goff = ok ? (y * W + x) * PIX_B + kc * 16 : 0;
          const uint32_t dst = stage + kc * KC_B + (X & 1) * PAR_B + Y * ROW_B + (X >> 1) * 16;
          if constexpr (BG) cp_async_16(dst, ok ? img + goff : padsrc, 16u);   // out of the image: -bg instead of zero fill
          else cp_async_16(dst, img + goff, ok ? 16u : 0u);
          j += STEP_J;
          Y += STEP_Y;
          if (j >= ROWC) {
            j -= ROWC;
            ++Y;
          }
        }
        cp_async_commit();
        PROBE_ADD(9, tp);
        if (i > 0) {
          cp_async_wait<1>();  // tile i-1 of this thread has landed
          fence_proxy_async_smem();
          mbar_arrive(&full[(i - 1) % STAGES]);
        }
        PROBE_ADD(10, tp);
      }
      if (my_units > 0) {
        cp_async_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(&full[(my_units - 1) % STAGES]);
      }
    } else {
      // planar source.  The halo row [x0, x0+18) with x0 = 16*tx - 1 is covered by six ALIGNED groups of four pixels
      // starting at 16*tx - 4: one 128-bit (fp32) or 32-bit (uint8) load per plane fetches a whole group, and because
      // W % 16 == 0 a group is either completely inside the image or completely outside (zero padding).  A task is
      // (halo row Y, group g): 204 tasks per tile, TPT per thread; the loads of tile i+1 are issued into a second
      // register set before tile i is converted and stored (loop unrolled by two, no register copies).
      constexpr int NTASK = HALO_H * 6, TPT = (NTASK + PROD_THREADS - 1) / PROD_THREADS, CH = 3;
      using Vec = typename std::conditional<SRC == SRC_CHW_F32, float4, uint32_t>::type;
      const int HW = H * W;
      int tY[TPT], tG[TPT];
#pragma unroll
      for (int k = 0; k < TPT; ++k) {
        const int task = ptid + k * PROD_THREADS;
        tY[k] = task / 6;
        tG[k] = task % 6;
      }
      auto tile_origin = [&](int i, int& n, int& y0, int& x0) {
        const int t = blockIdx.x + i * gridDim.x;
        int r, ty, tx;
        fd_img.divmod(t, n, r);
        fd_x.divmod(r, ty, tx);
        y0 = 2 * ty * TILE_PH - 1;
        x0 = 2 * tx * TILE_PW - 1;
      };
      auto load_tile = [&](int i, Vec (&v)[TPT][CH], uint32_t& okmask) {
        int n, y0, x0;
        tile_origin(i, n, y0, x0);
        // the per-image scalars store_tile() will need (PF - 1 tiles later): into L1 now, so that they cost a hit then
        // instead of an exposed L2 round trip per tile
        if constexpr (SRC == SRC_CHW_U8) prefetch_l1(stats + n);
        if constexpr (BG) prefetch_l1(static_cast<const float*>(bg_in) + 4 * (size_t)n);
        okmask = 0;
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
          const int y = y0 + tY[k], x = x0 - 3 + 4 * tG[k];
          const bool ok = (ptid + k * PROD_THREADS < NTASK) && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
          okmask |= (uint32_t)ok << k;
          const size_t off = (size_t)n * CH * HW + (ok ? y * W + x : 0);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if constexpr (SRC == SRC_CHW_F32) {
              v[k][c] = ok ? __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(src_any) + off + (size_t)c * HW))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              v[k][c] = ok ? __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(src_any) + off + (size_t)c * HW))
                           : 0u;
            }
          }
        }
      };
      auto store_tile = [&](int i, const Vec (&v)[TPT][CH], uint32_t okmask) {
        float scale = 1.0f, shift = 0.0f;   // (u/255 - mean) * rstd == u * scale + shift
        [[maybe_unused]] float pad_frac = 0.0f;   // BG == 2: fraction of the padding value 255 * mean (spare channel)
        [[maybe_unused]] float bgc[CH] = {0.0f, 0.0f, 0.0f};   // BG: this image's background value per channel
        if constexpr (BG == 2) {
          // exact integers: in the image u - r[c], outside floor(255 * mean) - r[c] (+ the fraction in the spare channel)
          const int n = fd_img.div(blockIdx.x + i * gridDim.x);
          const float m255 = 255.0f * __ldg(stats + n).x;
          shift = floorf(m255);
          pad_frac = m255 - shift;
          const uint32_t r = __float_as_uint(__ldg(static_cast<const float*>(bg_in) + 4 * (size_t)n + 3));
#pragma unroll
          for (int c = 0; c < CH; ++c) bgc[c] = (float)((r >> (8 * c)) & 255u);
        } else {
        if constexpr (SRC == SRC_CHW_U8) {
          const float2 st = __ldg(stats + fd_img.div(blockIdx.x + i * gridDim.x));
          scale = st.y * (1.0f / 255.0f), shift = -st.x * st.y;
        }
        if constexpr (BG) {
          const float* bp = static_cast<const float*>(bg_in) + 4 * (size_t)fd_img.div(blockIdx.x + i * gridDim.x);
#pragma unroll
          for (int c = 0; c < CH; ++c) bgc[c] = __ldg(bp + c);
        }
        }
        // one ring slot per tile; part 1 (the lo halves x - rn16(x) of the same pixels) follows part 0 inside the slot
        const int s = i % STAGES;
        long long tp = probe ? clock64() : 0;
        mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
        PROBE_ADD(8, tp);
#pragma unroll
        for (int part = 0; part < SPLIT; ++part) {
        uint8_t* stage = sA + s * C::STAGE_BYTES + part * C::A_BYTES;
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
          if (ptid + k * PROD_THREADS < NTASK) {
            const bool ok = (okmask >> k) & 1;
            float px[4][CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              if constexpr (SRC == SRC_CHW_F32) {
                px[0][c] = v[k][c].x, px[1][c] = v[k][c].y, px[2][c] = v[k][c].z, px[3][c] = v[k][c].w;
                if constexpr (BG) {           // the zero padding of x reads as -bg in the shifted space
#pragma unroll
                  for (int e = 0; e < 4; ++e) px[e][c] = (ok ? px[e][c] : 0.0f) - bgc[c];
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {   // zero padding applies to the NORMALISED image
                  if constexpr (BG == 2) px[e][c] = (ok ? (float)((v[k][c] >> (8 * e)) & 255u) : shift) - bgc[c];
                  else px[e][c] = (ok ? fmaf((float)((v[k][c] >> (8 * e)) & 255u), scale, shift) : 0.0f) - bgc[c];
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int X = 4 * tG[k] + e - 3;   // halo column of this pixel; groups 0 and 5 keep one pixel each
              if (X >= 0 && X < HALO_W) {
                uint2 pix;
                if (part == 0) {
                  pix = make_uint2(pack16<FMT>(px[e][0], px[e][1]), pack16<FMT>(px[e][2], (BG == 2 && !ok) ? pad_frac : 0.0f));
                } else {
                  pix = make_uint2(pack16<FMT>(px[e][0] - round16<FMT>(px[e][0]), px[e][1] - round16<FMT>(px[e][1])),
                                   pack16<FMT>(px[e][2] - round16<FMT>(px[e][2]), 0.0f));
                }
                uint8_t* row = stage + tY[k] * ROW_B;
                // chunk j = pixels (2j, 2j+1): one copy serves both window members of a pooling row (see mma_op)
                *reinterpret_cast<uint2*>(row + X * 8) = pix;
              }
            }
          }
        }
        }
        fence_proxy_async_smem();
        mbar_arrive(&full[s]);
        PROBE_ADD(10, tp);
      };
      // Register prefetch ring, PF tiles deep (loop unrolled by PF, no register copies): PF - 1 tile loads stay in flight
      // per CTA while tile i is converted and stored.  With the store warp in place the producers' global-load latency
      // is what the first layer waits for (cycle probe: producers 1 766 cycles per tile in "wait data").
      constexpr int PF = C::MERGED ? BBBP_CONV1_PF_MERGED : BBBP_CONV1_PF;
      Vec buf[PF][TPT][CH];
      uint32_t okm[PF];
#pragma unroll
      for (int q = 0; q < PF - 1; ++q)
        if (q < my_tiles) load_tile(q, buf[q], okm[q]);
      for (int i = 0; i < my_tiles; i += PF) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          if (i + q < my_tiles) {
            long long tl = probe ? clock64() : 0;
            if (i + q + PF - 1 < my_tiles) load_tile(i + q + PF - 1, buf[(q + PF - 1) % PF], okm[(q + PF - 1) % PF]);
            PROBE_ADD(9, tl);
            store_tile(i + q, buf[q], okm[q]);
          }
        }
      }
    }
  } else if (warp == C::STORE_WARP) {
    // ===== store warp: one bulk tensor store per tile, off the epilogue warps' critical path ============================
    // bar 1 ("output buffer b is free"): this warp arrives, the epilogue threads wait.  bar 2 ("tile written and fenced"):
    // the epilogue threads arrive, this warp waits.  Both count EPI_THREADS + 32.
    const int PH = H / 2;
    for (int i = 0; i < my_tiles; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      int n, r, ty, tx;
      fd_img.divmod(t, n, r);
      fd_x.divmod(r, ty, tx);
      if (lane == 0 && i >= 2) bulk_store_wait_read<1>();   // the store of tile i-2 has finished reading buffer i & 1
      __syncwarp();
      named_bar_arrive(1, EPI_THREADS + 32);
      named_bar_sync(2, EPI_THREADS + 32);
      if (lane == 0) {
        tma_store_3d(&tmOut, sOut + (i & 1) * C::OUT_BYTES, 0, tx * TILE_PW, n * PH + ty * TILE_PH);
        if constexpr (C::OSPLIT == 2)   // the lo tile travels in the same bulk group
          tma_store_3d(&tmOutLo, sOut + (2 + (i & 1)) * C::OUT_BYTES, 0, tx * TILE_PW, n * PH + ty * TILE_PH);
        bulk_store_commit();
      }
    }
    if (lane == 0) bulk_store_wait_all();   // shared memory must outlive the last store's reads
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer ==================================================================================================
    // The whole warp runs this loop (warp-uniform control flow, descriptor arithmetic on the uniform datapath); one
    // elected lane issues the MMAs and the commits.
    const uint32_t w_lo = smem_u32(sW) >> 4;
    for (int i = 0; i < my_tiles; ++i) {
      [[maybe_unused]] const int s = i % STAGES;
      const int b = i & 1;
      long long tp = probe ? clock64() : 0;
      mbar_wait(&acc_empty[b], ((i >> 1) & 1) ^ 1);
      PROBE_ADD(0, tp);
      if constexpr (SPLIT == 1 || C::MERGED) {
        mbar_wait(&full[s], (i / STAGES) & 1);
        PROBE_ADD(1, tp);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t a_lo = smem_u32(sA + s * C::STAGE_BYTES) >> 4;
          issue_tile<KC, COUT, C::PACK4, FMT, false>(a_lo, w_lo, tmem_base + b * C::ACC_COLS, std::make_integer_sequence<int, C::NISSUE>{});
          if constexpr (C::MERGED) {
            if constexpr (C::W_PARTS == 2)      // + x_hi * w_lo
              issue_tile<KC, COUT, C::PACK4, FMT, true>(a_lo, w_lo + (C::W_BYTES >> 4), tmem_base + b * C::ACC_COLS,
                                                        std::make_integer_sequence<int, C::NISSUE>{});
            issue_tile<KC, COUT, C::PACK4, FMT, true>(a_lo + (C::A_BYTES >> 4), w_lo, tmem_base + b * C::ACC_COLS,   // + x_lo * w_hi
                                                      std::make_integer_sequence<int, C::NISSUE>{});
          }
          umma_commit(&empty[s]);      // halo slot reusable once these MMAs have read it
          umma_commit(&acc_full[b]);   // accumulators of this tile complete
        }
      } else {
        // split input: ring units 2i (hi) and 2i + 1 (lo) feed the same accumulators
        const int u0 = 2 * i, s0 = u0 % STAGES, s1 = (u0 + 1) % STAGES;
        mbar_wait(&full[s0], (u0 / STAGES) & 1);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          issue_tile<KC, COUT, C::PACK4, FMT, false>(smem_u32(sA + s0 * C::STAGE_BYTES) >> 4, w_lo, tmem_base + b * C::ACC_COLS,
                                                     std::make_integer_sequence<int, C::NISSUE>{});
          if constexpr (C::W_PARTS == 2)      // + x_hi * w_lo
            issue_tile<KC, COUT, C::PACK4, FMT, true>(smem_u32(sA + s0 * C::STAGE_BYTES) >> 4, w_lo + (C::W_BYTES >> 4),
                                                      tmem_base + b * C::ACC_COLS, std::make_integer_sequence<int, C::NISSUE>{});
          umma_commit(&empty[s0]);
        }
        __syncwarp();
        mbar_wait(&full[s1], ((u0 + 1) / STAGES) & 1);
        PROBE_ADD(1, tp);
        tc_fence_after_sync();
        if (elect_one_sync()) {
          issue_tile<KC, COUT, C::PACK4, FMT, true>(smem_u32(sA + s1 * C::STAGE_BYTES) >> 4, w_lo, tmem_base + b * C::ACC_COLS,
                                                    std::make_integer_sequence<int, C::NISSUE>{});
          umma_commit(&empty[s1]);
          umma_commit(&acc_full[b]);
        }
      }
      __syncwarp();
      PROBE_ADD(2, tp);
      if (probe) probe[3] += 1;
    }
  } else {
    // ===== epilogue: max over the pooling window, + bias, ReLU, bf16, NHWC store ==========================================
    // 8 warps: warp % 4 is the TMEM lane quadrant it may read (32 pooled pixels), warp / 4 the half of the channels.
    // One warp per scheduler runs this mostly-serial code, so two warps per quadrant halve the epilogue's latency and
    // keep it hidden behind the next tile's MMAs.
    const int quad = warp & 3, half = warp >> 2;
    const int m = quad * 32 + lane;  // pooled pixel within the tile == TMEM lane == row of the output tile
    constexpr int CH = COUT / (C::EPI_WARPS / 4);  // channels per epilogue warp
    // The tile is written to shared memory in the TMA store's swizzled layout (row = pixel, 16-byte chunk j of row r
    // at position j ^ swz(r)): conflict-free 16-byte stores here, and ONE bulk tensor store per tile instead of 32
    // scattered 16-byte global stores per warp instruction (which cost the L1 data pipe 32 wavefronts each -- the
    // pipe this kernel is bound by).
    constexpr int SWZ_SHIFT = C::OUT_ROW_B == 128 ? 0 : 1, SWZ_MASK = C::OUT_ROW_B / 16 - 1;
    const int swz = (m >> SWZ_SHIFT) & SWZ_MASK;
    // BG: the table rows of a tile's image (T and bg_out = 2 * COUT floats) are copied to shared memory one tile
    // ahead by cp.async (thread e moves 16 bytes); the per-tile barrier below publishes them.  Reading them from global
    // memory in the channel loop instead cost the (latency-bound) epilogue ~2x its time.
    [[maybe_unused]] auto fetch_tab = [&](int it) {
      if constexpr (BG) {
        const int e = threadIdx.x;
        if (it < my_tiles && e < C::TAB_FLOATS / 4) {
          const int n = fd_img.div(blockIdx.x + it * gridDim.x);
          cp_async_16(smem_u32(sTab + (it & 1) * C::TAB_SLOT + 4 * e), bg_tab + (size_t)n * C::TAB_FLOATS + 4 * e, 16u);
        }
        if constexpr (BG == 2) {
          if (it < my_tiles && e == C::TAB_FLOATS / 4)      // the image's (mean, 1/std)
            cp_async_8(smem_u32(sTab + (it & 1) * C::TAB_SLOT + C::TAB_FLOATS), stats + fd_img.div(blockIdx.x + it * gridDim.x));
        }
        cp_async_commit();
      }
    };
    if constexpr (BG) {
      static_assert(C::TAB_FLOATS / 4 < EPI_THREADS, "one 16-byte copy per epilogue thread");
      fetch_tab(0);
      cp_async_wait<0>();
    }
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      uint8_t* otile = sOut + b * C::OUT_BYTES;
      long long tp = probe ? clock64() : 0;
      named_bar_sync(1, EPI_THREADS + 32);     // the store warp: tile i-2 has been read out of this buffer
      fetch_tab(i + 1);                        // (every epilogue thread is past tile i-1: its table buffer is free)
      PROBE_ADD(6, tp);
      mbar_wait(&acc_full[b], (i >> 1) & 1);
      PROBE_ADD(4, tp);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + b * C::ACC_COLS + half * CH;
      uint8_t* orow = otile + m * C::OUT_ROW_B;
      [[maybe_unused]] uint8_t* orow_lo = sOut + (2 + b) * C::OUT_BYTES + m * C::OUT_ROW_B;
      [[maybe_unused]] const uint32_t tb_s = BG ? smem_u32(sTab + b * C::TAB_SLOT + half * CH) : 0u;   // this tile's {T, bg_out}
      [[maybe_unused]] float acc_scale = 1.0f;   // BG == 2: rstd / 255, formed exactly as bbbp_image_background forms it
      if constexpr (BG == 2) acc_scale = ld_shared_f32(smem_u32(sTab + b * C::TAB_SLOT + C::TAB_FLOATS + 1)) * (1.0f / 255.0f);
      // CS channels per step (CS/8 output chunks): 4*CS live accumulator registers.  The first layer uses 8 so that its
      // variants fit the register cap of two CTAs per SM with eight producer warps; conv2 (no cap) uses 16.
      constexpr int CS = COUT >= 64 ? 16 : (C::PACK4 ? BBBP_CONV1_CS : 8);
#pragma unroll
      for (int c0 = 0; c0 < CH; c0 += CS) {
        uint32_t r0[CS], r1[CS], r2[CS], r3[CS];
        if constexpr (CS == 16) {
          tmem_ld_32x16(taddr + c0, r0);
          tmem_ld_32x16(taddr + COUT + c0, r1);
          tmem_ld_32x16(taddr + 2 * COUT + c0, r2);
          tmem_ld_32x16(taddr + 3 * COUT + c0, r3);
        } else {
          tmem_ld_32x8(taddr + c0, r0);
          tmem_ld_32x8(taddr + COUT + c0, r1);
          tmem_ld_32x8(taddr + 2 * COUT + c0, r2);
          tmem_ld_32x8(taddr + 3 * COUT + c0, r3);
        }
        tmem_ld_wait();
        uint32_t packed[CS / 2];
        [[maybe_unused]] uint32_t packed_lo[CS / 2];
#pragma unroll
        for (int j = 0; j < CS; j += 4) {
          if constexpr (BG) {
            const float4 t = ld_shared_f4(tb_s + (c0 + j) * 4), bo = ld_shared_f4(tb_s + (COUT + c0 + j) * 4);
            const float e[4] = {t.x, t.y, t.z, t.w}, o[4] = {bo.x, bo.y, bo.z, bo.w};
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float mx = fmaxf(fmaxf(__uint_as_float(r0[j + k]), __uint_as_float(r1[j + k])),
                                     fmaxf(__uint_as_float(r2[j + k]), __uint_as_float(r3[j + k])));
              v[k] = fmaxf(BG == 2 ? fmaf(mx, acc_scale, e[k]) : mx + e[k], 0.0f) - o[k];
            }
            packed[j / 2] = pack16<FMT>(v[0], v[1]);
            packed[j / 2 + 1] = pack16<FMT>(v[2], v[3]);
          } else {
          const float4 bv = *reinterpret_cast<const float4*>(sBias + half * CH + c0 + j);
          const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            float v0 = fmaxf(fmaxf(__uint_as_float(r0[j + k]), __uint_as_float(r1[j + k])),
                             fmaxf(__uint_as_float(r2[j + k]), __uint_as_float(r3[j + k])));
            float v1 = fmaxf(fmaxf(__uint_as_float(r0[j + k + 1]), __uint_as_float(r1[j + k + 1])),
                             fmaxf(__uint_as_float(r2[j + k + 1]), __uint_as_float(r3[j + k + 1])));
            v0 = fmaxf(v0 + bb[k], 0.0f);
            v1 = fmaxf(v1 + bb[k + 1], 0.0f);
            if constexpr (C::OSPLIT == 2) split16<FMT>(v0, v1, packed[(j + k) / 2], packed_lo[(j + k) / 2]);
            else packed[(j + k) / 2] = pack16<FMT>(v0, v1);
          }
          }
        }
        const int chunk = (half * CH + c0) / 8;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          *reinterpret_cast<uint4*>(orow + (((chunk + g) ^ swz) << 4)) =
              make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
          if constexpr (C::OSPLIT == 2)
            *reinterpret_cast<uint4*>(orow_lo + (((chunk + g) ^ swz) << 4)) =
                make_uint4(packed_lo[4 * g], packed_lo[4 * g + 1], packed_lo[4 * g + 2], packed_lo[4 * g + 3]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[b]);   // TMEM reads done: the MMA warp may start the tile after next
      if constexpr (BG) cp_async_wait<0>();   // the next tile's table has landed (published by the next per-tile barrier)
      PROBE_ADD(5, tp);
      fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_arrive(2, EPI_THREADS + 32);   // tile written: the store warp takes it from here
      PROBE_ADD(7, tp);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- weight / input re-layout (prepare-time and per-call helpers) --------------------------------------------------
// conv2-style (KC >= 2): wp[kc][8 - tap][n][8] = w[n][kc*8 + c][kh][kw]; taps are stored in REVERSE order so that the
// weights of taps (kh, kw) and (kh, kw-1) -- the two B halves of a paired MMA -- are adjacent 64-row blocks.
// One thread = one 3x3 filter (n, ci).  fp16 weights are rounded with ERROR DIFFUSION over the nine taps (tap order 0..8,
// float32 carry), so the filter's SUM -- its response to a locally constant input, which is what the background of a
// depiction is after the first layer -- keeps full precision: plain round-to-nearest leaves every filter a random DC error
// of ~0.9 ulp, and the same DC error at every background pixel adds up coherently in the Linear(65536, 128) that follows
// (tests/precision_study.py: the strict mode's largest remaining term, 7e-4 -> 2.6e-4).  bf16 weights keep plain
// round-to-nearest (the fast mode's results stay bit-identical to earlier releases).
__global__ void prep_weights_kc_kernel(const float* __restrict__ w, uint16_t* __restrict__ wp, int Cin, int Cout, int fmt) {
  const int KC = Cin / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KC * Cout * 8) return;
  const int c = i % 8, n = (i / 8) % Cout, kc = i / (8 * Cout);
  const float* f = w + ((size_t)n * Cin + kc * 8 + c) * 9;
  float carry = 0.0f;
  for (int tap = 0; tap < 9; ++tap) {
    const float target = f[tap] + carry;
    const uint16_t r = cvt16_rt(target, fmt);
    if (fmt == BBBP_FMT_F16) carry = target - round16<BBBP_FMT_F16>(target);
    wp[((size_t)(kc * 9 + (8 - tap)) * Cout + n) * 8 + c] = r;
  }
}
// conv1-style (Cin <= 8, one chunk): wp[dx][pair][chunk][n][8]
__global__ void prep_weights_c8_kernel(const float* __restrict__ w, uint16_t* __restrict__ wp, int Cin, int Cout, int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * 5 * 2 * Cout * 8) return;
  const int c = i % 8, n = (i / 8) % Cout, chunk = (i / (8 * Cout)) % 2, pair = (i / (16 * Cout)) % 5, dx = i / (80 * Cout);
  const TapPair tp = conv1_pair(dx, pair);
  const int tap = chunk == 0 ? tp.first : tp.second;
  const bool zero = chunk == tp.zero_slot || c >= Cin;
  wp[i] = cvt16_rt(zero ? 0.0f : w[((size_t)n * Cin + c) * 9 + tap], fmt);
}
// PACK4 image (planar sources): wp[kh][chunk][n2][8], n2 = dx*Cout + n.  K index k = chunk*8 + e = j*4 + c addresses pixel
// X + j of the halo row (X = 2*pw) and channel c; member dx uses tap kw = j - dx: zero where that is not in 0..2 or c >= Cin
// lo == 1: the low part rn(w - rn(w)) of the same image (strict mode, pair design); lo == 2: the image with the filter's sum
// over its input channels in the spare fourth channel (exact-integer uint8 form, BG == 2)
__global__ void prep_weights_pack4_kernel(const float* __restrict__ w, uint16_t* __restrict__ wp, int Cin, int Cout, int fmt,
                                          int lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 2 * (2 * Cout) * 8) return;
  const int e = i % 8, n2 = (i / 8) % (2 * Cout), chunk = (i / (16 * Cout)) % 2, kh = i / (32 * Cout);
  const int k = chunk * 8 + e, j = k / 4, c = k % 4, dx = n2 / Cout, n = n2 % Cout, kw = j - dx;
  float v = (kw >= 0 && kw < 3 && c < Cin) ? w[((size_t)n * Cin + c) * 9 + kh * 3 + kw] : 0.0f;
  if (lo == 2 && kw >= 0 && kw < 3 && c == Cin)     // spare channel: the tap summed over the input channels (BG == 2)
    for (int ci = 0; ci < Cin; ++ci) v += w[((size_t)n * Cin + ci) * 9 + kh * 3 + kw];
  wp[i] = cvt16_rt(lo == 1 ? v - round16_rt(v, fmt) : v, fmt);
}
// fp32 NCHW image (C <= 8 planes) -> bf16 NHWC with 8 channels per pixel (zero padded): one 16-byte store per pixel
__global__ void __launch_bounds__(256) image_to_nhwc8_kernel(const float* __restrict__ img, uint4* __restrict__ out,
                                                              int C, int HW, size_t total_pixels) {
  const size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (p >= total_pixels) return;
  const size_t n = p / HW, hw = p % HW;
  const float* s = img + n * (size_t)C * HW + hw;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? s[(size_t)c * HW] : 0.0f;
  __nv_bfloat162 h[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) h[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
  out[p] = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                      *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
}
// Linear weight over a flattened (C,H,W) activation -> the same weight over the (H,W,C) flattening, bf16
__global__ void fc_weight_to_hwc_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int C, int HW,
                                        size_t total, int fmt) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t K = (size_t)C * HW;
  const size_t o = i / K, k = i % K;
  const size_t hw = k / C, c = k % C;
  out[i] = cvt16_rt(w[o * K + c * HW + hw], fmt);
}

// per-image mean and 1/std of x/255 over all C*H*W values, fp64 accumulation (sklearn StandardScaler on one molecule's
// pixel column; std 0 -> 1), narrowed to fp32
__global__ void __launch_bounds__(256) u8_image_stats_kernel(const uint8_t* __restrict__ img, float2* __restrict__ stats,
                                                              int n) {
  __shared__ double red_s[8], red_q[8];
  const uint8_t* src = img + (size_t)blockIdx.x * n;
  // integer sums are exact: sum(u) < 2^32, sum(u^2) < 2^40 for n <= 2^24
  unsigned long long su = 0, sq = 0;
  for (int i = threadIdx.x * 16; i < n; i += 256 * 16) {
    if (i + 16 <= n && ((reinterpret_cast<uintptr_t>(src + i) & 15) == 0)) {
      const uint4 w = *reinterpret_cast<const uint4*>(src + i);
      const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const unsigned u = (ws[k] >> (8 * b)) & 255u;
          su += u;
          sq += u * u;
        }
    } else {
      for (int k = i; k < min(n, i + 16); ++k) {
        const unsigned u = src[k];
        su += u;
        sq += u * u;
      }
    }
  }
  double ds = (double)su, dq = (double)sq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  if (threadIdx.x % 32 == 0) red_s[threadIdx.x / 32] = ds, red_q[threadIdx.x / 32] = dq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0, Q = 0;
    for (int i = 0; i < 8; ++i) S += red_s[i], Q += red_q[i];
    const double mean = S / 255.0 / n;
    double var = Q / (255.0 * 255.0) / n - mean * mean;
    if (var < 0) var = 0;
    double sd = sqrt(var);
    if (sd == 0.0) sd = 1.0;
    stats[blockIdx.x] = make_float2((float)mean, (float)(1.0 / sd));
  }
}

// ---- background-referenced strict mode: per-image tables (see the kernel comment, BG = 1) ------------------------------
// Background value of an image per channel: the value most of the 8 probe pixels (4 corners, 4 edge midpoints) agree on,
// compared as raw (R, G, B) triples; computed with the producers' own arithmetic, so x - bg is EXACTLY 0 on the canvas.
// Any choice is mathematically exact (the tables below account for it); this one makes the background cost no precision.
__global__ void __launch_bounds__(128) image_background_kernel(const void* __restrict__ img, int is_u8,
                                                               const float2* __restrict__ stats, float* __restrict__ bg, int N,
                                                               int H, int W) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int HW = H * W;
  const int py[8] = {0, 0, H - 1, H - 1, 0, H / 2, H - 1, H / 2}, px[8] = {0, W - 1, 0, W - 1, W / 2, 0, W / 2, W - 1};
  uint32_t key[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t at = ((size_t)n * 3 + c) * HW + py[k] * W + px[k];
      key[k][c] = is_u8 ? (uint32_t) static_cast<const uint8_t*>(img)[at] : __float_as_uint(static_cast<const float*>(img)[at]);
    }
  int best = 0, best_votes = -1;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int votes = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) votes += (key[j][0] == key[k][0]) & (key[j][1] == key[k][1]) & (key[j][2] == key[k][2]);
    if (votes > best_votes) best_votes = votes, best = k;
  }
  float scale = 1.0f, shift = 0.0f;
  if (is_u8) {
    const float2 st = stats[n];
    scale = st.y * (1.0f / 255.0f), shift = -st.x * st.y;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t kv = key[0][c];
#pragma unroll
    for (int k = 1; k < 8; ++k) kv = best == k ? key[k][c] : kv;
    bg[4 * (size_t)n + c] = is_u8 ? fmaf((float)kv, scale, shift) : __uint_as_float(kv);
  }
  // column 3: the raw background bytes r | g << 8 | b << 16 of a uint8 image (read by the exact-integer first layer only)
  uint32_t raw = 0;
  if (is_u8) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t kv = key[0][c];
#pragma unroll
      for (int k = 1; k < 8; ++k) kv = best == k ? key[k][c] : kv;
      raw |= kv << (8 * c);
    }
  }
  bg[4 * (size_t)n + 3] = __uint_as_float(raw);
}

// One "layer" of the background chain:  T[n][co] = bias[co] + sum_ci wsum[co][ci] * in[n][ci]  (wsum = the layer's weights
// summed over the taps / positions a constant input reaches: bbbp_fc_weight_channel_sums), written to out0; optionally
// out1[n][co] = relu(T), rounded to the 16-bit operand format when fmt >= 0 (the next layer's background must be a 16-bit
// number, see the kernel comment), and neg16[n][co] = -out1 in that format (the next layer's padding value).  fp32 FMAs in
// a fixed order; weights transposed in shared memory, lane = output channel, 4 images per thread.
constexpr int BGL_IMGS = 4;
__global__ void __launch_bounds__(256) bg_layer_kernel(const float* __restrict__ wsum, const float* __restrict__ bias,
                                                       const float* __restrict__ in, int ld_in, int Cin, int Cout,
                                                       float* __restrict__ out0, int ld0, float* __restrict__ out1, int ld1,
                                                       uint16_t* __restrict__ neg16, int fmt, int N, int imgs_per_block) {
  extern __shared__ float sm_bgl[];
  const int pitch = Cout + 1;
  float* swT = sm_bgl;                         // [Cin][Cout + 1]
  float* sin = sm_bgl + Cin * pitch;           // [imgs_per_block][Cin]
  const int n0 = blockIdx.x * imgs_per_block, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < Cout * Cin; i += 256) swT[(i % Cin) * pitch + i / Cin] = wsum[i];
  for (int i = threadIdx.x; i < imgs_per_block * Cin; i += 256) {
    const int img = i / Cin, ci = i % Cin;
    sin[i] = n0 + img < N ? in[(size_t)(n0 + img) * ld_in + ci] : 0.0f;
  }
  __syncthreads();
  const int cogroups = Cout / 32, co = (warp % cogroups) * 32 + lane;
  const float bs = bias ? bias[co] : 0.0f;
  for (int q = warp / cogroups; q * BGL_IMGS < imgs_per_block; q += 8 / cogroups) {
    float S[BGL_IMGS];
#pragma unroll
    for (int g = 0; g < BGL_IMGS; ++g) S[g] = bs;
    for (int ci = 0; ci < Cin; ++ci) {
      const float wv = swT[ci * pitch + co];
#pragma unroll
      for (int g = 0; g < BGL_IMGS; ++g) S[g] = fmaf(wv, sin[(q * BGL_IMGS + g) * Cin + ci], S[g]);
    }
#pragma unroll
    for (int g = 0; g < BGL_IMGS; ++g) {
      const int n = n0 + q * BGL_IMGS + g;
      if (n >= N) continue;
      out0[(size_t)n * ld0 + co] = S[g];
      float r = fmaxf(S[g], 0.0f);
      if (fmt >= 0) r = round16_rt(r, fmt);
      if (out1) out1[(size_t)n * ld1 + co] = r;
      if (neg16) neg16[(size_t)n * Cout + co] = cvt16_rt(-r, fmt);
    }
  }
}

// out[o][c] = sum over hw of w[o][c*HW + hw]: what a Linear over the flattened (C, H, W) activation does to a per-channel
// constant.  One warp per (o, c), fixed order.
__global__ void __launch_bounds__(256) fc_weight_channel_sums_kernel(const float* __restrict__ w, float* __restrict__ out,
                                                                     int rows_x_C, int HW) {
  const int gw = (blockIdx.x * 256 + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (gw >= rows_x_C) return;
  const float* src = w + (size_t)gw * HW;
  float s = 0.0f;
  for (int i = lane; i < HW; i += 32) s += src[i];
  s = warp_sum(s);
  if (lane == 0) out[gw] = s;
}

struct BgArgs {
  const void* bg_in = nullptr;
  const float* tab = nullptr;
};
template <int KC, int COUT, int SRC, int FMT, int SPLIT, int BG = 0>
int launch(const void* x, const void* x_lo, const float* stats, const void* wprep, const float* bias, void* y, void* y_lo, int N,
           int H, int W, cudaStream_t stream, BgArgs bg = BgArgs{}) {
  using C = Cfg<KC, COUT, SRC, FMT, SPLIT, BG>;
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    cudaFuncSetAttribute(conv3x3_umma_kernel<KC, COUT, SRC, FMT, SPLIT, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  const int sms = current_sm_count();
  // NHWC output as a 3-D tensor {COUT, PW, N*PH}; one box = one tile (COUT x 8 x 16), swizzle span = one pixel row
  CUtensorMap tmOut, tmOutLo;
  {
    tensormap_encode_fn enc = get_tensormap_encoder();
    if (!enc) return BBBP_ECUDA;
    const int PH = H / 2, PW = W / 2;
    cuuint64_t dims[3] = {(cuuint64_t)COUT, (cuuint64_t)PW, (cuuint64_t)N * PH};
    cuuint64_t strides[2] = {(cuuint64_t)COUT * 2, (cuuint64_t)PW * COUT * 2};
    cuuint32_t box[3] = {(cuuint32_t)COUT, TILE_PW, TILE_PH};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int part = 0; part < C::OSPLIT; ++part) {
      CUresult r = enc(part ? &tmOutLo : &tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, part ? y_lo : y, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, COUT * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv3x3_relu_pool16: cuTensorMapEncodeTiled(output) failed (%d)", (int)r);
        return BBBP_ECUDA;
      }
    }
    if (C::OSPLIT == 1) tmOutLo = tmOut;
  }
  const int tiles = N * ((W / 2) / TILE_PW) * ((H / 2) / TILE_PH);
  if ((long long)tiles * (((W / 2) / TILE_PW) * ((H / 2) / TILE_PH)) >= (1ll << 32)) {
    set_error("conv3x3_relu_pool16: %d tiles of %dx%d images exceed the kernel's tile index range", tiles, H, W);
    return BBBP_EINVAL;
  }
  const int per_sm = C::MIN_CTAS;
  const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
  conv3x3_umma_kernel<KC, COUT, SRC, FMT, SPLIT, BG><<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(
      x, x_lo, reinterpret_cast<const float2*>(stats), static_cast<const uint4*>(wprep), bias, tmOut, tmOutLo, N, H, W, bg.bg_in,
      bg.tab);
  return launch_status("conv3x3_relu_pool16");
}

// (fmt, split) -> instantiation.  Built: bf16 one pass (fast mode), fp16 one pass, fp16 split (strict mode).
template <int KC, int COUT, int SRC>
int dispatch(int fmt, int split, const void* x, const void* x_lo, const float* stats, const void* wprep, const float* bias,
             void* y, void* y_lo, int N, int H, int W, cudaStream_t stream) {
  if (fmt == BBBP_FMT_BF16 && split == 1)
    return launch<KC, COUT, SRC, BBBP_FMT_BF16, 1>(x, nullptr, stats, wprep, bias, y, nullptr, N, H, W, stream);
  if (fmt == BBBP_FMT_F16 && split == 1)
    return launch<KC, COUT, SRC, BBBP_FMT_F16, 1>(x, nullptr, stats, wprep, bias, y, nullptr, N, H, W, stream);
  if (fmt == BBBP_FMT_F16 && split == 2)
    return launch<KC, COUT, SRC, BBBP_FMT_F16, 2>(x, x_lo, stats, wprep, bias, y, y_lo, N, H, W, stream);
  set_error("tcgen05 conv: (fmt %d, split %d) is not built (bf16 x1, fp16 x1, fp16 x2)", fmt, split);
  return BBBP_EUNSUPPORTED;
}

}  // namespace conv
}  // namespace bbbp

using namespace bbbp;

extern "C" size_t bbbp_conv3x3_prepared_bytes(int Cin, int Cout) {
  if (Cin <= 8) return (size_t)2 * 5 * 2 * Cout * 16 + (size_t)3 * 3 * 2 * (2 * Cout) * 16;   // NHWC8 | PACK4 hi | PACK4 lo | PACK4 + channel sums
  return (size_t)9 * (Cin / 8) * Cout * 16;
}

extern "C" int bbbp_conv3x3_prepare16(int fmt, const float* w, void* wprep, int Cin, int Cout, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "conv3x3_prepare: bad fmt %d", fmt);
  BBBP_CHECK_ARG(w && wprep, "conv3x3_prepare: null operand");
  BBBP_CHECK_ARG((Cin == 3 && (Cout == 32 || Cout == 64)) || (Cin == 32 && Cout == 64),
                 "conv3x3_prepare: only (3->32), (3->64) and (32->64) are built for the tcgen05 path, got %d->%d", Cin, Cout);
  const int total = (int)(bbbp_conv3x3_prepared_bytes(Cin, Cout) / 2);
  uint16_t* wp = static_cast<uint16_t*>(wprep);
  if (Cin <= 8) {
    const int n8 = 2 * 5 * 2 * Cout * 8, n4 = 3 * 2 * (2 * Cout) * 8;
    conv::prep_weights_c8_kernel<<<ceil_div(n8, 256), 256, 0, as_stream(stream)>>>(w, wp, Cin, Cout, fmt);
    conv::prep_weights_pack4_kernel<<<ceil_div(n4, 256), 256, 0, as_stream(stream)>>>(w, wp + n8, Cin, Cout, fmt, 0);
    conv::prep_weights_pack4_kernel<<<ceil_div(n4, 256), 256, 0, as_stream(stream)>>>(w, wp + n8 + n4, Cin, Cout, fmt, 1);
    conv::prep_weights_pack4_kernel<<<ceil_div(n4, 256), 256, 0, as_stream(stream)>>>(w, wp + n8 + 2 * n4, Cin, Cout, fmt, 2);
    note_launches(3);
  } else
    conv::prep_weights_kc_kernel<<<ceil_div(total / 9, 256), 256, 0, as_stream(stream)>>>(w, wp, Cin, Cout, fmt);
  return launch_status("conv3x3_prepare");
}
extern "C" int bbbp_conv3x3_prepare_bf16(const float* w, void* wprep, int Cin, int Cout, bbbp_stream_t stream) {
  return bbbp_conv3x3_prepare16(BBBP_FMT_BF16, w, wprep, Cin, Cout, stream);
}

extern "C" int bbbp_conv3x3_relu_pool16(int fmt, int split, const void* x_nhwc, const void* x_lo, const void* wprep,
                                        const float* bias, void* y_nhwc, void* y_lo, int N, int Cin_pad, int Cout, int H, int W,
                                        bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x_nhwc && wprep && bias && y_nhwc, "conv3x3_relu_pool16: null operand");
  BBBP_CHECK_ARG(split == 1 || (split == 2 && x_lo && y_lo), "conv3x3_relu_pool16: split must be 1, or 2 with both lo tensors");
  BBBP_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 32 == 0 && W % 16 == 0,
                 "conv3x3_relu_pool16: H=%d must be a multiple of 32 and W=%d of 16", H, W);
  BBBP_CHECK_ARG(((uintptr_t)x_nhwc % 16) == 0 && ((uintptr_t)y_nhwc % 16) == 0 && ((uintptr_t)wprep % 16) == 0 &&
                     ((uintptr_t)x_lo % 16) == 0 && ((uintptr_t)y_lo % 16) == 0,
                 "conv3x3_relu_pool16: operands must be 16-byte aligned");
  if (N == 0) return BBBP_OK;
  cudaStream_t s = as_stream(stream);
  if (Cin_pad == 8 && Cout == 32 && fmt == BBBP_FMT_BF16 && split == 1)
    return conv::launch<1, 32, conv::SRC_NHWC_BF16, BBBP_FMT_BF16, 1>(x_nhwc, nullptr, nullptr, wprep, bias, y_nhwc, nullptr, N, H, W, s);
  if (Cin_pad == 32 && Cout == 64)
    return conv::dispatch<4, 64, conv::SRC_NHWC_BF16>(fmt, split, x_nhwc, x_lo, nullptr, wprep, bias, y_nhwc, y_lo, N, H, W, s);
  set_error("conv3x3_relu_pool16: unsupported channels %d->%d (built: 8->32 [bf16], 32->64)", Cin_pad, Cout);
  return BBBP_EUNSUPPORTED;
}
extern "C" int bbbp_conv3x3_relu_pool_bf16(const void* x_nhwc, const void* wprep, const float* bias, void* y_nhwc, int N,
                                           int Cin_pad, int Cout, int H, int W, bbbp_stream_t stream) {
  return bbbp_conv3x3_relu_pool16(BBBP_FMT_BF16, 1, x_nhwc, nullptr, wprep, bias, y_nhwc, nullptr, N, Cin_pad, Cout, H, W, stream);
}

extern "C" int bbbp_image_to_nhwc8_bf16(const float* img_nchw, void* out_nhwc8, int N, int C, int H, int W,
                                        bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img_nchw && out_nhwc8 && N >= 0 && C >= 1 && C <= 8 && H > 0 && W > 0, "image_to_nhwc8: bad argument");
  const size_t total = (size_t)N * H * W;
  if (total == 0) return BBBP_OK;
  conv::image_to_nhwc8_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(
      img_nchw, static_cast<uint4*>(out_nhwc8), C, H * W, total);
  return launch_status("image_to_nhwc8");
}

extern "C" int bbbp_fc_weight_to_hwc16(int fmt, const float* w, void* out16, int rows, int C, int HW, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "fc_weight_to_hwc: bad fmt %d", fmt);
  BBBP_CHECK_ARG(w && out16 && rows > 0 && C > 0 && HW > 0, "fc_weight_to_hwc: bad argument");
  const size_t total = (size_t)rows * C * HW;
  conv::fc_weight_to_hwc_kernel<<<(unsigned)ceil_div(total, (size_t)256), 256, 0, as_stream(stream)>>>(
      w, static_cast<uint16_t*>(out16), C, HW, total, fmt);
  return launch_status("fc_weight_to_hwc");
}
extern "C" int bbbp_fc_weight_to_hwc_bf16(const float* w, void* out_bf16, int rows, int C, int HW, bbbp_stream_t stream) {
  return bbbp_fc_weight_to_hwc16(BBBP_FMT_BF16, w, out_bf16, rows, C, HW, stream);
}

extern "C" int bbbp_u8_image_stats_f32(const uint8_t* img, float* stats, int rows, int n, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img && stats && rows >= 0 && n > 0 && n <= (1 << 24), "u8_image_stats: bad argument");
  if (rows == 0) return BBBP_OK;
  conv::u8_image_stats_kernel<<<rows, 256, 0, as_stream(stream)>>>(img, reinterpret_cast<float2*>(stats), n);
  return launch_status("u8_image_stats");
}

extern "C" int bbbp_conv1_from_image16(int fmt, int split, const void* img_chw, int img_is_u8, const float* stats,
                                       const void* wprep, const float* bias, void* y_nhwc, void* y_lo, int N, int H, int W,
                                       bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img_chw && wprep && bias && y_nhwc, "conv1_from_image: null operand");
  BBBP_CHECK_ARG(split == 1 || (split == 2 && y_lo), "conv1_from_image: split must be 1, or 2 with y_lo");
  BBBP_CHECK_ARG(!img_is_u8 || stats, "conv1_from_image: uint8 input needs the per-image (mean, 1/std) table");
  BBBP_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 32 == 0 && W % 16 == 0,
                 "conv1_from_image: H=%d must be a multiple of 32 and W=%d of 16", H, W);
  if (N == 0) return BBBP_OK;
  cudaStream_t s = as_stream(stream);
  if (img_is_u8)
    return conv::dispatch<1, 32, conv::SRC_CHW_U8>(fmt, split, img_chw, nullptr, stats, wprep, bias, y_nhwc, y_lo, N, H, W, s);
  return conv::dispatch<1, 32, conv::SRC_CHW_F32>(fmt, split, img_chw, nullptr, nullptr, wprep, bias, y_nhwc, y_lo, N, H, W, s);
}
// First layer of the big variant (nn.Conv2d(3, 64, 3, padding=1) + ReLU + MaxPool2d(2), 20250107_network.py:133-135): the same
// fused kernel with 64 output channels (N = 128 MMAs, all 512 TMEM columns, one CTA per SM), bf16, fp32 planar input.
extern "C" int bbbp_conv1_from_image_c64_bf16(const float* img_chw, const void* wprep, const float* bias, void* y_nhwc, int N,
                                              int H, int W, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img_chw && wprep && bias && y_nhwc, "conv1_from_image_c64: null operand");
  BBBP_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 32 == 0 && W % 16 == 0,
                 "conv1_from_image_c64: H=%d must be a multiple of 32 and W=%d of 16", H, W);
  if (N == 0) return BBBP_OK;
  return conv::launch<1, 64, conv::SRC_CHW_F32, BBBP_FMT_BF16, 1>(img_chw, nullptr, nullptr, wprep, bias, y_nhwc, nullptr, N, H, W,
                                                                  as_stream(stream));
}
extern "C" int bbbp_conv1_from_image_bf16(const void* img_chw, int img_is_u8, const float* stats, const void* wprep,
                                          const float* bias, void* y_nhwc, int N, int H, int W, bbbp_stream_t stream) {
  return bbbp_conv1_from_image16(BBBP_FMT_BF16, 1, img_chw, img_is_u8, stats, wprep, bias, y_nhwc, nullptr, N, H, W, stream);
}

// ---- background-referenced strict mode (BG = 1) -------------------------------------------------------------------
extern "C" int bbbp_image_background(const void* img_chw, int img_is_u8, const float* stats, float* bg, int N, int H, int W,
                                     bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img_chw && bg && N >= 0 && H >= 2 && W >= 2, "image_background: bad argument");
  BBBP_CHECK_ARG(!img_is_u8 || stats, "image_background: uint8 input needs the per-image (mean, 1/std) table");
  if (N == 0) return BBBP_OK;
  conv::image_background_kernel<<<ceil_div(N, 128), 128, 0, as_stream(stream)>>>(img_chw, img_is_u8,
                                                                                reinterpret_cast<const float2*>(stats), bg, N, H, W);
  return launch_status("image_background");
}

extern "C" int bbbp_bg_layer(const float* wsum, const float* bias, const float* in, int ld_in, int Cin, int Cout, float* out0,
                             int ld0, float* out1, int ld1, void* neg16, int fmt, int N, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(wsum && in && out0 && N >= 0, "bg_layer: null operand");
  BBBP_CHECK_ARG(Cin >= 1 && Cin <= 256 && ld_in >= Cin && Cout >= 32 && Cout <= 256 && Cout % 32 == 0 && (256 / 32) % (Cout / 32) == 0,
                 "bg_layer: Cin=%d (1..256), ld_in=%d, Cout=%d (32, 64, 128 or 256)", Cin, ld_in, Cout);
  BBBP_CHECK_ARG(ld0 >= Cout && (!out1 || ld1 >= Cout), "bg_layer: output pitch < Cout");
  BBBP_CHECK_ARG(fmt == -1 || fmt == BBBP_FMT_BF16 || fmt == BBBP_FMT_F16, "bg_layer: fmt %d (-1 = no rounding)", fmt);
  BBBP_CHECK_ARG(!neg16 || fmt >= 0, "bg_layer: neg16 needs a 16-bit format");
  if (N == 0) return BBBP_OK;
  // images per block: every warp of a channel group takes two passes of 4 images
  const int imgs = 2 * conv::BGL_IMGS * (8 / (Cout / 32));
  const int smem = (Cin * (Cout + 1) + imgs * Cin) * (int)sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    cudaFuncSetAttribute(conv::bg_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  BBBP_CHECK_ARG(smem <= 200 * 1024, "bg_layer: %d -> %d channels need %d bytes of shared memory", Cin, Cout, smem);
  conv::bg_layer_kernel<<<ceil_div(N, imgs), 256, smem, as_stream(stream)>>>(wsum, bias, in, ld_in, Cin, Cout, out0, ld0, out1, ld1,
                                                                           static_cast<uint16_t*>(neg16), fmt, N, imgs);
  return launch_status("bg_layer");
}

extern "C" int bbbp_fc_weight_channel_sums(const float* w, float* out, int rows, int C, int HW, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(w && out && rows > 0 && C > 0 && HW > 0, "fc_weight_channel_sums: bad argument");
  conv::fc_weight_channel_sums_kernel<<<ceil_div(rows * C * 32, 256), 256, 0, as_stream(stream)>>>(w, out, rows * C, HW);
  return launch_status("fc_weight_channel_sums");
}

extern "C" int bbbp_conv1_from_image_bg16(int fmt, int split, const void* img_chw, int img_is_u8, const float* stats,
                                          const void* wprep, const float* bg_in, const float* tab, void* y_nhwc, int N, int H,
                                          int W, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(img_chw && wprep && bg_in && tab && y_nhwc, "conv1_from_image_bg: null operand");
  BBBP_CHECK_ARG(fmt == BBBP_FMT_F16 && (split == 1 || split == 2), "conv1_from_image_bg: built for fp16, split 1 or 2");
  BBBP_CHECK_ARG(!img_is_u8 || stats, "conv1_from_image_bg: uint8 input needs the per-image (mean, 1/std) table");
  BBBP_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 32 == 0 && W % 16 == 0,
                 "conv1_from_image_bg: H=%d must be a multiple of 32 and W=%d of 16", H, W);
  BBBP_CHECK_ARG(((uintptr_t)tab % 16) == 0 && ((uintptr_t)y_nhwc % 16) == 0, "conv1_from_image_bg: table and output must be 16-byte aligned");
  if (N == 0) return BBBP_OK;
  cudaStream_t s = as_stream(stream);
  const conv::BgArgs bg{bg_in, tab};
  if (img_is_u8) {
    if (split == 2)
      return conv::launch<1, 32, conv::SRC_CHW_U8, BBBP_FMT_F16, 2, 1>(img_chw, nullptr, stats, wprep, nullptr, y_nhwc, nullptr, N, H, W, s, bg);
    // one pass on exact integers (BG = 2): u - background byte is exact in fp16, so raw depictions need no (hi, lo) pair
    return conv::launch<1, 32, conv::SRC_CHW_U8, BBBP_FMT_F16, 1, 2>(img_chw, nullptr, stats, wprep, nullptr, y_nhwc, nullptr, N, H, W, s, bg);
  }
  if (split == 2)
    return conv::launch<1, 32, conv::SRC_CHW_F32, BBBP_FMT_F16, 2, 1>(img_chw, nullptr, nullptr, wprep, nullptr, y_nhwc, nullptr, N, H, W, s, bg);
  return conv::launch<1, 32, conv::SRC_CHW_F32, BBBP_FMT_F16, 1, 1>(img_chw, nullptr, nullptr, wprep, nullptr, y_nhwc, nullptr, N, H, W, s, bg);
}

extern "C" int bbbp_conv3x3_relu_pool_bg16(int fmt, const void* x_nhwc, const void* wprep, const void* neg_bg_in, const float* tab,
                                           void* y_nhwc, int N, int Cin_pad, int Cout, int H, int W, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(x_nhwc && wprep && neg_bg_in && tab && y_nhwc, "conv3x3_relu_pool_bg: null operand");
  BBBP_CHECK_ARG(fmt == BBBP_FMT_F16 && Cin_pad == 32 && Cout == 64, "conv3x3_relu_pool_bg: built for fp16, 32 -> 64 channels");
  BBBP_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H % 32 == 0 && W % 16 == 0,
                 "conv3x3_relu_pool_bg: H=%d must be a multiple of 32 and W=%d of 16", H, W);
  BBBP_CHECK_ARG(((uintptr_t)x_nhwc % 16) == 0 && ((uintptr_t)y_nhwc % 16) == 0 && ((uintptr_t)wprep % 16) == 0 &&
                     ((uintptr_t)tab % 16) == 0 && ((uintptr_t)neg_bg_in % 16) == 0,
                 "conv3x3_relu_pool_bg: operands must be 16-byte aligned");
  if (N == 0) return BBBP_OK;
  return conv::launch<4, 64, conv::SRC_NHWC_BF16, BBBP_FMT_F16, 1, 1>(x_nhwc, nullptr, nullptr, wprep, nullptr, y_nhwc, nullptr, N, H, W,
                                                                      as_stream(stream), conv::BgArgs{neg_bg_in, tab});
}

// debug: probe = device array of 16 uint64 counters (zero it first), or NULL to switch the probe off
extern "C" int bbbp_debug_conv_probe(void* probe) {
  unsigned long long* p = static_cast<unsigned long long*>(probe);
  cudaError_t e = cudaMemcpyToSymbol(conv::g_conv_probe, &p, sizeof(p));
  if (e != cudaSuccess) {
    set_error("debug_conv_probe: %s", cudaGetErrorString(e));
    return BBBP_ECUDA;
  }
  return BBBP_OK;
}
