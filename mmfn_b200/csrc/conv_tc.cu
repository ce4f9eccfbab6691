// TF32 tensor-core implicit-GEMM convolutions for the ResNet trunks (torchvision BasicBlock convs;
// call sites model_rad.py:523-525, :542-544, :560-562, :577-579), NHWC activations, KRSC filters.
//
// No im2col buffer exists anywhere: the activation operand is fetched by 4-D TMA boxes
// (channels x W x H x image) whose start coordinate carries the filter-tap offset; out-of-bounds
// rows/columns (the zero padding) are filled by the TMA unit, strided convolutions use the tensor
// map's traversal stride.
//   forward / data-gradient : M = 128 output pixels (image x rows x cols patch), N = Co tile,
//                             K loops over (tap, 32-channel chunk); both operands K-major.
//   weight-gradient         : M = Co tile, N = Ci tile, K loops over 32-pixel patches; both operands
//                             MN-major (pixels are the slow dimension of dY and x); one CTA per
//                             (tap, split) accumulates atomically into dW.
#include <type_traits>
#include "tc_kernel.cuh"

// norm.cu: stand-alone batch-statistics pass (used when a split-K launch cannot accumulate them in its epilogue)
int mmfn_bn_stats_launch(const float* x, int64_t M, int C, float* mean, float* rstd, float* running_mean, float* running_var,
                         float momentum, float eps, double* ws, cudaStream_t stream);

namespace {

struct ConvGeomTc {
  int N, H, W, C;        // input
  int Co, R, S, stride, pad;
  int Ho, Wo;
  int BW, BH, BI;        // pixel patch of one tile / k-block
  int tiles_w, tiles_h, tiles_n;
};

// DGRAD = true: the stride-1 DATA GRADIENT run as a forward convolution of dy -- here g.C = channels of dy (the conv's
// Co), g.Co = channels of dx (the conv's C) -- reading the UNTRANSFORMED KRSC filters w[k = co][tap][n = ci] as an MN-major
// B operand ([32 co][32 ci] boxes, SWIZZLE_128B_ATOM_32B) with mirrored taps: no CRSK filter copy is ever made.
template <int TBN, bool DGRAD = false, int EB_ = 32>
struct ConvFwdOp {
  static constexpr bool A_MN = false, B_MN = DGRAD;
  static constexpr int EB = EB_;                    // channels per 128-byte k-block: 32 (tf32) / 64 (bf16)
  using ET = ElemTraits<EB_>;
  ConvGeomTc g;
  int kb_per_split;          // split-K over (tap, channel-chunk) blocks for layers with few pixel tiles
  int w0, h0, i0, co0;
  __device__ void setup() {
    int t = blockIdx.y;
    int tw = t % g.tiles_w; t /= g.tiles_w;
    int th = t % g.tiles_h;
    int tn = t / g.tiles_h;
    w0 = tw * g.BW; h0 = th * g.BH; i0 = tn * g.BI;
    co0 = blockIdx.x * TBN;
  }
  __device__ int kb_begin() const { return blockIdx.z * kb_per_split; }
  __device__ int kb_end() const { return min(g.R * g.S * (g.C / EB), (int)(blockIdx.z + 1) * kb_per_split); }
  __device__ void load(int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar, const CUtensorMap* ta, const CUtensorMap* tb) const {
    int cch = g.C / EB;
    int tap = kb / cch, cc = kb - tap * cch;
    int r = tap / g.S, s = tap - r * g.S;
    tc::tma_load_4d(sa, ta, bar, cc * EB, w0 * g.stride - g.pad + s, h0 * g.stride - g.pad + r, i0);
    if constexpr (!DGRAD) tc::tma_load_2d(sb, tb, bar, tap * g.C + cc * EB, co0);
    else {
      const int ftap = g.R * g.S - 1 - tap;                              // mirrored tap of the original filter
#pragma unroll
      for (int j = 0; j < TBN / EB; ++j) tc::tma_load_2d(sb + j * ET::BOX_BYTES, tb, bar, ftap * g.Co + co0 + EB * j, cc * EB);
    }
  }
  __device__ bool out_row(int r, int64_t& off) const {
    int per = g.BH * g.BW;
    int img = r / per, rem = r - img * per;
    int hh = rem / g.BW, ww = rem - hh * g.BW;
    int n = i0 + img, h = h0 + hh, w = w0 + ww;
    off = (((int64_t)n * g.Ho + h) * g.Wo + w) * g.Co;
    return n < g.N && h < g.Ho && w < g.Wo;
  }
  __device__ int n_cols() const { return g.Co; }
  __device__ int col0() const { return co0; }
  __device__ bool first_split() const { return blockIdx.z == 0; }
};

// Data gradient of a STRIDE-2 convolution without zero insertion: the input pixels split into four parity classes
// (h % 2, w % 2); a class only ever meets the filter taps r = (h + pad) % 2 (mod 2) (and likewise s), so each class is a
// small stride-1 implicit GEMM over dy with 1 / 2 / 2 / 4 of the 9 taps of a 3x3 filter (1 / 0 / 0 / 0 of a 1x1):
//   dx[n, 2y+ph, 2x+pw, :] = sum_{(r,s) in class, co} dy[n, y + (ph+pad-r)/2, x + (pw+pad-s)/2, co] * w[co, r, s, :]
// blockIdx.z = parity class; M = 128 pixels of the (H/2, W/2) sub-grid; B = the KRSC filters, MN-major.
// A class without taps (1x1 filters, odd pixels) still runs its epilogue and writes zeros (+ residual).
template <int TBN, int EB_ = 32>
struct ConvDgradS2Op {
  static constexpr bool A_MN = false, B_MN = true;        // B = the KRSC filters themselves, [EB co][EB ci] boxes
  static constexpr int EB = EB_;
  using ET = ElemTraits<EB_>;
  ConvGeomTc g;              // g.H, g.W: dx (= conv input) size; g.Ho, g.Wo: dy size; g.C: dx channels; g.Co: dy channels
  int unused_;
  int w0, h0, i0, ci0, ph, pw, r0, s0, nr, ns;
  __device__ void setup() {
    int t = blockIdx.y;
    int tw = t % g.tiles_w; t /= g.tiles_w;
    int th = t % g.tiles_h;
    int tn = t / g.tiles_h;
    w0 = tw * g.BW; h0 = th * g.BH; i0 = tn * g.BI;
    ci0 = blockIdx.x * TBN;
    ph = blockIdx.z >> 1; pw = blockIdx.z & 1;
    r0 = (ph + g.pad) & 1; s0 = (pw + g.pad) & 1;          // first tap of the class; taps step by 2
    nr = r0 < g.R ? (g.R - r0 + 1) / 2 : 0;
    ns = s0 < g.S ? (g.S - s0 + 1) / 2 : 0;
  }
  __device__ int kb_begin() const { return 0; }
  __device__ int kb_end() const { return nr * ns * (g.Co / EB); }
  __device__ void load(int kb, uint8_t* sa, uint8_t* sb, uint64_t* bar, const CUtensorMap* ta, const CUtensorMap* tb) const {
    int cch = g.Co / EB;
    int tap = kb / cch, cc = kb - tap * cch;
    int r = r0 + 2 * (tap / ns), s = s0 + 2 * (tap % ns);
    int oy = (ph + g.pad - r) / 2, ox = (pw + g.pad - s) / 2;   // exact: numerator is even (C++ division of negatives: -2/2)
    tc::tma_load_4d(sa, ta, bar, cc * EB, w0 + ox, h0 + oy, i0);
#pragma unroll
    for (int j = 0; j < TBN / EB; ++j) tc::tma_load_2d(sb + j * ET::BOX_BYTES, tb, bar, (r * g.S + s) * g.C + ci0 + EB * j, cc * EB);
  }
  __device__ bool out_row(int r, int64_t& off) const {
    int per = g.BH * g.BW;
    int img = r / per, rem = r - img * per;
    int hh = rem / g.BW, ww = rem - hh * g.BW;
    int n = i0 + img, h = 2 * (h0 + hh) + ph, w = 2 * (w0 + ww) + pw;
    off = (((int64_t)n * g.H + h) * g.W + w) * g.C;
    return n < g.N && h < g.H && w < g.W;
  }
  __device__ int n_cols() const { return g.C; }
  __device__ int col0() const { return ci0; }
  __device__ bool first_split() const { return true; }
};

template <int TBN, int EB_ = 32>
struct ConvWgradOp {
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int EB = EB_;                    // pixels per k-block (= channels per MN box)
  using ET = ElemTraits<EB_>;
  ConvGeomTc g;
  int splitk, pb_per_split;
  int tap, r_tap, s_tap, co0, ci0, pb0, pb1;
  __device__ void setup() {
    tap = blockIdx.z / splitk;
    int split = blockIdx.z - tap * splitk;
    r_tap = tap / g.S; s_tap = tap - r_tap * g.S;
    co0 = blockIdx.y * tc::TBM;
    ci0 = blockIdx.x * TBN;
    int npb = g.tiles_w * g.tiles_h * g.tiles_n;
    pb0 = split * pb_per_split;
    pb1 = min(npb, pb0 + pb_per_split);
  }
  __device__ int kb_begin() const { return pb0; }
  __device__ int kb_end() const { return pb1; }
  __device__ void load(int pb, uint8_t* sa, uint8_t* sb, uint64_t* bar, const CUtensorMap* ta, const CUtensorMap* tb) const {
    int t = pb;
    int tw = t % g.tiles_w; t /= g.tiles_w;
    int th = t % g.tiles_h;
    int tn = t / g.tiles_h;
    int w0 = tw * g.BW, h0 = th * g.BH, i0 = tn * g.BI;
    for (int i = 0; i < tc::TBM / EB; ++i) tc::tma_load_4d(sa + i * ET::BOX_BYTES, ta, bar, co0 + EB * i, w0, h0, i0);
    for (int j = 0; j < TBN / EB; ++j)
      tc::tma_load_4d(sb + j * ET::BOX_BYTES, tb, bar, ci0 + EB * j, w0 * g.stride - g.pad + s_tap, h0 * g.stride - g.pad + r_tap, i0);
  }
  __device__ bool out_row(int r, int64_t& off) const {
    int co = co0 + r;
    off = ((int64_t)co * g.R * g.S + tap) * g.C;
    return co < g.Co;
  }
  __device__ int n_cols() const { return g.C; }
  __device__ int col0() const { return ci0; }
  __device__ bool first_split() const { return true; }
};

// Train-mode BatchNorm statistics of the convolution output, accumulated by the epilogue (tc_kernel.cuh) when the
// launch is not split over K; otherwise by a separate reduction pass (norm.cu).
struct BnStatsArgs { double* ws; float* mean; float* rstd; float* rmean; float* rvar; float momentum, eps; };
static inline void set_bn(tc::Epilogue& e, const BnStatsArgs* bn, long long rows) {
  if (!bn) return;
  e.bn_ws = bn->ws; e.bn_mean = bn->mean; e.bn_rstd = bn->rstd; e.bn_rmean = bn->rmean; e.bn_rvar = bn->rvar;
  e.bn_eps = bn->eps; e.bn_momentum = bn->momentum; e.bn_rows = rows;
}

int check_tc_geom(const ConvGeomTc& g, const char* what, int EB = 32) {
  MMFN_CHECK_ARG(g.N > 0 && g.H > 0 && g.W > 0 && g.C > 0 && g.Co > 0 && g.R > 0 && g.S > 0 && g.stride > 0 && g.pad >= 0,
                 "%s: bad sizes", what);
  MMFN_CHECK_ARG(g.Ho == (g.H + 2 * g.pad - g.R) / g.stride + 1 && g.Wo == (g.W + 2 * g.pad - g.S) / g.stride + 1,
                 "%s: inconsistent output size", what);
  MMFN_CHECK_ARG(g.C % EB == 0, "%s: input channels must be a multiple of %d", what, EB);
  MMFN_CHECK_ARG(g.Co % 4 == 0, "%s: output channels must be a multiple of 4", what);
  return 0;
}

// activation tensor (N,H,W,C) as a 4-D map (C, W, H, N) with a (32, bw*stride, bh*stride, bi) box
int make_act_tmap(CUtensorMap* m, const void* x, int N, int H, int W, int C, int bw, int bh, int bi, int stride, bool swz32, int EB = 32) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
  uint32_t box[4] = {(uint32_t)EB, (uint32_t)(bw * stride), (uint32_t)(bh * stride), (uint32_t)bi};
  uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  if (EB == 64) return mmfn_make_tmap_bf16(m, x, 4, dims, strides, box, es);
  return mmfn_make_tmap_f32(m, static_cast<const float*>(x), 4, dims, strides, box, es, swz32);
}

// KRSC filters w(Co,R,S,C) as the MN-major B operand of a data-gradient GEMM: a 2-D map with inner = (tap, ci)
// (contiguous, R*S*C long) and outer = co; boxes of [32 co][32 ci] in the SWIZZLE_128B_ATOM_32B pattern.
int make_krsc_b_tmap(CUtensorMap* m, const void* w, int Co, int R, int S, int C, int EB = 32) {
  uint64_t dims[2] = {(uint64_t)R * S * C, (uint64_t)Co}, strides[2] = {1, (uint64_t)R * S * C};
  uint32_t box[2] = {(uint32_t)EB, (uint32_t)EB};
  if (EB == 64) return mmfn_make_tmap_bf16(m, w, 2, dims, strides, box, nullptr);
  return mmfn_make_tmap_f32(m, static_cast<const float*>(w), 2, dims, strides, box, nullptr, true);
}
// KRSC filters as the K-major B operand of a forward convolution: [tbn co][EB (tap, ci)] boxes
int make_krsc_fwd_tmap(CUtensorMap* m, const void* w, int Co, int R, int S, int C, int tbn, int EB) {
  uint64_t dims[2] = {(uint64_t)R * S * C, (uint64_t)Co}, strides[2] = {1, (uint64_t)R * S * C};
  uint32_t box[2] = {(uint32_t)EB, (uint32_t)tbn};
  if (EB == 64) return mmfn_make_tmap_bf16(m, w, 2, dims, strides, box, nullptr);
  return mmfn_make_tmap_f32(m, static_cast<const float*>(w), 2, dims, strides, box, nullptr, false);
}

// ---------------------------------------------------------------- 3x3 stride-1 convolutions with input-halo reuse
// The implicit GEMM above re-fetches every input pixel once per filter tap (9 x for 3x3), and with fp32 operands these
// kernels run at the L2 -> SM bandwidth cap.  Here ONE (16+2) x (8+2) pixel patch of a 32-channel slice is loaded per
// slice (one 4-D TMA box, zero padding = OOB fill) and all nine taps issue their MMAs from it: the A descriptor of tap
// (r, s) simply starts (r * 10 + s) rows into the patch with a group stride SBO = 10 rows.  (A K-major SWIZZLE_128B
// operand may start at any 128-byte row of a TMA-written tile and use any 128-byte-multiple group stride: the swizzle
// is a function of the absolute shared-memory address -- tools/exp/desc_shift_probe.cu.)  Output tile = 16 rows x
// 8 cols, so a UMMA row group (8 consecutive M rows) is 8 consecutive pixels of one image row.
//   A ring: 2 x 23 KB patches (warp 2 lane 0 produces, then joins the epilogue); B ring: 64 KB of [TBN co x 32 ch] filter
//   tiles, one per (slice, tap) (warp 0).  A traffic drops 9 x 16 KB -> 23 KB per slice.
constexpr int PT_BH = 16, PT_BW = 8, PT_PH = PT_BH + 2, PT_PW = PT_BW + 2;
constexpr int PT_A_BYTES = PT_PH * PT_PW * 128;            // 23 040
constexpr int PT_A_STAGE = 23 * 1024;

// DEEP = false: 3 patches + 32 KB of filter tiles = 101 KB, two CTAs per SM (grids of more than one CTA per SM).
// DEEP = true : 5 patches + 64 KB of filter tiles = 179 KB, one CTA per SM: a patch is consumed in ~0.6 us (9 taps) but
//               takes > 1 us to arrive (180 scattered 128-byte rows), so single-wave grids need more patches in flight.
template <int TBN, bool DEEP>
struct PatchSmem {
  static constexpr int B_BYTES = TBN * 128;
  static constexpr int A_STAGES = DEEP ? 5 : 3;
  static constexpr int B_STAGES = (DEEP ? 65536 : 32768) / B_BYTES;
  static constexpr int B_OFF = A_STAGES * PT_A_STAGE;
  static constexpr int BAR_OFF = B_OFF + B_STAGES * B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(PT_A_STAGE >= PT_A_BYTES && 8 * 32 * 36 * 4 + 128 * 8 <= BAR_OFF, "patch conv smem layout");
};

template <int TBN, bool DGRAD, int EB_ = 32>
struct ConvPatchOp {
  static constexpr bool A_MN = false, B_MN = DGRAD;
  static constexpr int EB = EB_;
  ConvGeomTc g;              // forward roles (for DGRAD: in = dy, out = dx, g.C = dy channels, g.Co = dx channels)
  int w0, h0, img, co0;
  __device__ void setup() {
    int t = blockIdx.y;
    int tw = t % g.tiles_w; t /= g.tiles_w;
    int th = t % g.tiles_h;
    img = t / g.tiles_h;
    w0 = tw * PT_BW; h0 = th * PT_BH;
    co0 = blockIdx.x * TBN;
  }
  __device__ bool out_row(int r, int64_t& off) const {
    int h = h0 + (r >> 3), w = w0 + (r & 7);
    off = (((int64_t)img * g.Ho + h) * g.Wo + w) * g.Co;
    return h < g.Ho && w < g.Wo;
  }
  __device__ int n_cols() const { return g.Co; }
  __device__ int col0() const { return co0; }
  __device__ bool first_split() const { return true; }
};

template <int TBN, bool DGRAD, bool DEEP, int EB = 32>
__global__ void __launch_bounds__(tc::TC_THREADS, DEEP ? 1 : 2)
conv3x3_patch_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     ConvPatchOp<TBN, DGRAD, EB> op, tc::Epilogue e) {
  using L = PatchSmem<TBN, DEEP>;
  using ET = ElemTraits<EB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* a_empty = a_full + L::A_STAGES;
  uint64_t* b_full = a_empty + L::A_STAGES;
  uint64_t* b_empty = b_full + L::B_STAGES;
  uint64_t* tmem_full = b_empty + L::B_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  op.setup();
  const int nsl = op.g.C / EB;                             // EB-channel (128-byte) slices of the reduction

  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < L::A_STAGES; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < L::B_STAGES; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, TBN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {                                 // ===== filter-tile producer =====
      int st = 0; uint32_t ph = 0;
      for (int sl = 0; sl < nsl; ++sl)
        for (int tap = 0; tap < 9; ++tap) {
          tc::mbar_wait(&b_empty[st], ph ^ 1);
          uint8_t* sb = smem + L::B_OFF + st * L::B_BYTES;
          tc::mbar_expect_tx(&b_full[st], L::B_BYTES);
          if constexpr (!DGRAD) tc::tma_load_2d(sb, &tmB, &b_full[st], tap * op.g.C + sl * EB, op.co0);
          else {
            const int ftap = 8 - tap;                      // mirrored tap of the original filter
#pragma unroll
            for (int j = 0; j < TBN / EB; ++j)
              tc::tma_load_2d(sb + j * ET::BOX_BYTES, &tmB, &b_full[st], ftap * op.g.Co + op.co0 + EB * j, sl * EB);
          }
          if (++st == L::B_STAGES) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {                                 // ===== MMA issuer =====
      const uint32_t idesc = ET::idesc(tc::TBM, TBN, false, DGRAD);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      for (int sl = 0; sl < nsl; ++sl) {
        tc::mbar_wait(&a_full[sa], pa);
        tc::tc_fence_after();
        const uint32_t patch = tc::smem_u32(smem + sa * PT_A_STAGE);
        for (int tap = 0; tap < 9; ++tap) {
          tc::mbar_wait(&b_full[sb], pb);
          tc::tc_fence_after();
          const int r = tap / 3, s = tap - 3 * r;
          const uint32_t a0 = patch + (uint32_t)(r * PT_PW + s) * 128u;
          const uint32_t b0 = tc::smem_u32(smem + L::B_OFF + sb * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < tc::TBK / tc::UMMA_K; ++k) {
            const uint64_t ad = tc::smem_desc(a0 + k * 32, 16, PT_PW * 128, 2);
            const uint64_t bd = DGRAD ? ET::mn_desc(b0 + k * ET::MN_STEP) : tc::smem_desc_kmajor(b0 + k * 32);
            ET::mma(tmem_base, ad, bd, idesc, (sl | tap | k) ? 1u : 0u);
          }
          tc::mma_commit(&b_empty[sb]);
          if (++sb == L::B_STAGES) { sb = 0; pb ^= 1; }
        }
        tc::mma_commit(&a_empty[sa]);
        if (++sa == L::A_STAGES) { sa = 0; pa ^= 1; }
      }
      tc::mma_commit(tmem_full);
    }
  } else {
    if (warp == 2) {
      if (tc::elect_one()) {                               // ===== patch producer (then an epilogue warp like the rest) =====
        int st = 0; uint32_t ph = 0;
        for (int sl = 0; sl < nsl; ++sl) {
          tc::mbar_wait(&a_empty[st], ph ^ 1);
          tc::mbar_expect_tx(&a_full[st], PT_A_BYTES);
          tc::tma_load_4d(smem + st * PT_A_STAGE, &tmA, &a_full[st], sl * EB, op.w0 - 1, op.h0 - 1, op.img);
          if (++st == L::A_STAGES) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
    }
    tc::tc_epilogue<ConvPatchOp<TBN, DGRAD, EB>, TBN, false>(op, e, smem, tmem_full, tmem_base, 0, nsl);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TBN);
  __shared__ bool bn_last;
  tc::tc_bn_finalize(e, op.g.Co, &bn_last);
}

// x: conv input (fwd) or dy (dgrad), NHWC with Cin channels; out NHWC with Cout channels, same H x W (3x3, stride 1, pad 1)
template <bool DGRAD, int EB = 32>
int launch_conv3x3_patch(const void* x, const void* w, float* out, const float* res, int N, int H, int W, int Cin, int Cout,
                         cudaStream_t stream, const char* what, const BnStatsArgs* bn = nullptr) {
  ConvGeomTc g{N, H, W, Cin, Cout, 3, 3, 1, 1, H, W};
  g.BW = PT_BW; g.BH = PT_BH; g.BI = 1;
  g.tiles_w = (W + PT_BW - 1) / PT_BW; g.tiles_h = (H + PT_BH - 1) / PT_BH; g.tiles_n = N;
  const int ptiles = g.tiles_w * g.tiles_h * g.tiles_n;
  MMFN_CHECK_ARG(ptiles <= 65535, "%s: too many pixel tiles", what);
  CUtensorMap ta, tb;
  if (int rc = make_act_tmap(&ta, x, N, H, W, Cin, PT_PW, PT_PH, 1, 1, false, EB)) return rc;
  // 128-wide filter tiles as soon as they still give ~100 CTAs: a TF32 MMA re-reads its 4 KB A operand from shared
  // memory for every 128 x N x 8 step, so N = 64 is bound by operand bandwidth (6 KB per 33 clk of math), N = 128 much less
  const int tbn = (Cout % 128 == 0 && ptiles * (Cout / 128) >= 100) ? 128 : (Cout <= 64 || ptiles * ((Cout + 127) / 128) < 148) ? 64 : 128;
  if (DGRAD) {
    if (int rc = make_krsc_b_tmap(&tb, w, Cin, 3, 3, Cout, EB)) return rc;      // w is (Co_conv = Cin here, 3, 3, C_conv = Cout)
  } else {
    if (int rc = make_krsc_fwd_tmap(&tb, w, Cout, 3, 3, Cin, tbn, EB)) return rc;
  }
  tc::Epilogue e{out, nullptr, res, nullptr, 1.f, 0, 0, 0.f, 0, nullptr};
  set_bn(e, bn, (long long)N * H * W);
  auto go = [&](auto tbn_tag, auto deep_tag) -> int {
    constexpr int TBN = decltype(tbn_tag)::value;
    constexpr bool DEEP = decltype(deep_tag)::value;
    using L = PatchSmem<TBN, DEEP>;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t ce = cudaFuncSetAttribute(conv3x3_patch_kernel<TBN, DGRAD, DEEP, EB>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
      if (ce != cudaSuccess) { mmfn_set_error("%s: smem attribute: %s", what, cudaGetErrorString(ce)); return (int)ce; }
      attr_set = true;
    }
    ConvPatchOp<TBN, DGRAD, EB> op{g};
    conv3x3_patch_kernel<TBN, DGRAD, DEEP, EB><<<dim3((Cout + TBN - 1) / TBN, ptiles, 1), tc::TC_THREADS, L::TOTAL, stream>>>(ta, tb, op, e);
    return mmfn_launch_status(what);
  };
  const bool deep = ptiles * ((Cout + tbn - 1) / tbn) <= 148 && Cin / EB > 3;  // single wave and more than 3 slices
  if (tbn == 64) return deep ? go(std::integral_constant<int, 64>{}, std::true_type{}) : go(std::integral_constant<int, 64>{}, std::false_type{});
  return deep ? go(std::integral_constant<int, 128>{}, std::true_type{}) : go(std::integral_constant<int, 128>{}, std::false_type{});
}

static inline bool patch_conv_ok(int R, int S, int stride, int pad, int H, int W) {
  return R == 3 && S == 3 && stride == 1 && pad == 1 && H >= 16 && W >= 8;
}

__global__ void krsc_to_crsk_kernel(const float* __restrict__ w, float* __restrict__ wt, int Co, int R, int S, int C, int flip) {
  int64_t n = (int64_t)Co * R * S * C;
  int RS = R * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int co = (int)(i % Co);
    int64_t t = i / Co;
    int rs = (int)(t % RS);
    int c = (int)(t / RS);
    int src_rs = flip ? (RS - 1 - rs) : rs;
    wt[i] = w[((int64_t)co * RS + src_rs) * C + c];
  }
}

}  // namespace

// wt[c][r][s][co] = w[co][r][s][c]; with flip != 0 the taps are mirrored (r,s -> R-1-r, S-1-s), which
// turns a stride-1 data-gradient into a plain forward convolution of dy with wt.
MMFN_API int mmfn_filter_krsc_to_crsk(const float* w, float* wt, int Co, int R, int S, int C, int flip,
                                      cudaStream_t stream) {
  MMFN_CHECK_ARG(w && wt && Co > 0 && R > 0 && S > 0 && C > 0, "filter permute: bad args");
  int64_t n = (int64_t)Co * R * S * C;
  krsc_to_crsk_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(w, wt, Co, R, S, C, flip);
  return mmfn_launch_status("krsc_to_crsk");
}

template <int EB>
static int conv_fwd_impl(const void* x, const void* w, float* y, const float* res,
                         int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                         int Ho, int Wo, cudaStream_t stream, const BnStatsArgs* bn = nullptr) {
  const char* what = EB == 64 ? "conv_fwd_bf16" : "conv_fwd_tf32";
  MMFN_CHECK_ARG(!bn || (bn->ws && bn->mean && bn->rstd && !res && Co % 64 == 0), "%s: batch statistics need ws / mean / rstd, no residual, Co %% 64 == 0", what);
  MMFN_CHECK_ARG(x && w && y, "%s: null pointer", what);
  ConvGeomTc g{N, H, W, C, Co, R, S, stride, pad, Ho, Wo};
  if (int rc = check_tc_geom(g, what, EB)) return rc;
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0, "%s: operands must be 16-byte aligned", what);
  MMFN_CHECK_ARG(Wo >= 8 && Ho >= 8, "%s: output must be at least 8x8", what);
  if (patch_conv_ok(R, S, stride, pad, H, W)) return launch_conv3x3_patch<false, EB>(x, w, y, res, N, H, W, C, Co, stream, what, bn);
  // 128-pixel tile: 8 rows x 16 cols of one image, or two whole 8x8 maps
  g.BW = Wo >= 16 ? 16 : 8;
  g.BH = 8;
  g.BI = tc::TBM / (g.BW * g.BH);
  MMFN_CHECK_ARG(g.BW * g.BH * g.BI == tc::TBM && g.BW * stride <= 256 && g.BH * stride <= 256, "%s: unsupported tile", what);
  g.tiles_w = (Wo + g.BW - 1) / g.BW; g.tiles_h = (Ho + g.BH - 1) / g.BH; g.tiles_n = (N + g.BI - 1) / g.BI;
  CUtensorMap ta, tb;
  if (int rc = make_act_tmap(&ta, x, N, H, W, C, g.BW, g.BH, g.BI, stride, false, EB)) return rc;
  // 128-wide filter tiles unless that leaves SMs idle (deep layers at small batch): then twice as many 64-wide
  // tiles, which halves both the per-CTA epilogue and the split-K factor needed to fill the chip
  const int tbn = (Co <= 64 || g.tiles_w * g.tiles_h * g.tiles_n * ((Co + 127) / 128) < 148) ? 64 : 128;
  if (int rc = make_krsc_fwd_tmap(&tb, w, Co, R, S, C, tbn, EB)) return rc;
  int ptiles = g.tiles_w * g.tiles_h * g.tiles_n;
  MMFN_CHECK_ARG(ptiles <= 65535, "%s: too many pixel tiles", what);
  // deep layers at small batch have few output tiles (B=16: 64 / 32 CTAs at 16x16 / 8x8): split the
  // (tap, channel) reduction across CTAs and accumulate atomically into a zeroed output
  const int nkb = R * S * (C / EB);
  const int ctas = ptiles * ((Co + tbn - 1) / tbn);
  int splitk = 1;
  if (ctas < 148) splitk = max(1, min(nkb / 8, (2 * 148) / ctas));   // whole grid co-resident: 2 CTAs per SM
  int kb_per = (nkb + splitk - 1) / splitk;
  splitk = (nkb + kb_per - 1) / kb_per;
  if (splitk > 1) {
    cudaError_t ce = cudaMemsetAsync(y, 0, sizeof(float) * (size_t)N * Ho * Wo * Co, stream);
    if (ce != cudaSuccess) { mmfn_set_error("%s: memset: %s", what, cudaGetErrorString(ce)); return (int)ce; }
  }
  tc::Epilogue e{y, nullptr, res, nullptr, 1.f, 0, splitk > 1 ? 2 : 0, 0.f, 0, mmfn_tc_trace_ptr()};
  if (splitk == 1) set_bn(e, bn, (long long)N * Ho * Wo);       // partial sums of a K split are not the output: separate pass below
  int rc;
  if (tbn == 64) {
    ConvFwdOp<64, false, EB> op{g, kb_per};
    rc = tc::launch<ConvFwdOp<64, false, EB>, 64, 4>(ta, tb, op, e, dim3((Co + 63) / 64, ptiles, splitk), stream, what);
  } else {
    ConvFwdOp<128, false, EB> op{g, kb_per};
    rc = tc::launch<ConvFwdOp<128, false, EB>, 128, 3>(ta, tb, op, e, dim3((Co + 127) / 128, ptiles, splitk), stream, what);
  }
  if (rc == 0 && bn && splitk > 1)
    rc = mmfn_bn_stats_launch(y, (int64_t)N * Ho * Wo, Co, bn->mean, bn->rstd, bn->rmean, bn->rvar, bn->momentum, bn->eps, bn->ws, stream);
  return rc;
}

// Data gradient on the tensor cores straight from the KRSC filters w(Co,R,S,C) (no transposed / mirrored filter copy):
// dx(N,H,W,C) = dgrad of y = conv(x, w, stride, pad) [+ res], from dy(N,Ho,Wo,Co).  stride 1 (any R, S, pad < R) or
// stride 2 (H, W even).  Co % EB == 0, C % 4 == 0 (bf16: C % 8).
template <int EB>
static int conv2d_dgrad_s2(const void* dy, const void* wt, float* dx, const float* res,
                           int N, int H, int W, int C, int Co, int R, int S, int pad,
                           int Ho, int Wo, cudaStream_t stream, const char* what) {
  MMFN_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && Co > 0 && R > 0 && S > 0 && pad >= 0, "%s: bad sizes", what);
  MMFN_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && Ho == (H + 2 * pad - R) / 2 + 1 && Wo == (W + 2 * pad - S) / 2 + 1,
                 "%s: H, W must be even and consistent with Ho, Wo", what);
  const int Hs = H / 2, Ws = W / 2;                    // one parity class of dx
  MMFN_CHECK_ARG(Hs >= 8 && Ws >= 8, "%s: dx must be at least 16x16", what);
  ConvGeomTc g{N, H, W, C, Co, R, S, 2, pad, Ho, Wo};
  g.BW = Ws >= 16 ? 16 : 8;
  g.BH = 8;
  g.BI = tc::TBM / (g.BW * g.BH);
  g.tiles_w = (Ws + g.BW - 1) / g.BW; g.tiles_h = (Hs + g.BH - 1) / g.BH; g.tiles_n = (N + g.BI - 1) / g.BI;
  CUtensorMap ta, tb;
  if (int rc = make_act_tmap(&ta, dy, N, Ho, Wo, Co, g.BW, g.BH, g.BI, 1, false, EB)) return rc;
  const int ptiles = g.tiles_w * g.tiles_h * g.tiles_n;
  MMFN_CHECK_ARG(ptiles <= 65535, "%s: too many pixel tiles", what);
  const int tbn = (C <= 64 || ptiles * 4 * ((C + 127) / 128) < 148) ? 64 : 128;
  if (int rc = make_krsc_b_tmap(&tb, wt, Co, R, S, C, EB)) return rc;
  tc::Epilogue e{dx, nullptr, res, nullptr, 1.f, 0, 0, 0.f, 0, mmfn_tc_trace_ptr()};
  if (tbn == 64) {
    ConvDgradS2Op<64, EB> op{g, 0};
    return tc::launch<ConvDgradS2Op<64, EB>, 64, 4>(ta, tb, op, e, dim3((C + 63) / 64, ptiles, 4), stream, what);
  }
  ConvDgradS2Op<128, EB> op{g, 0};
  return tc::launch<ConvDgradS2Op<128, EB>, 128, 3>(ta, tb, op, e, dim3((C + 127) / 128, ptiles, 4), stream, what);
}

template <int EB>
static int conv_dgrad_impl(const void* dy, const void* w, float* dx, const float* res,
                           int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                           int Ho, int Wo, cudaStream_t stream) {
  const char* what = EB == 64 ? "conv_dgrad_bf16" : "conv_dgrad_tf32";
  MMFN_CHECK_ARG(dy && w && dx, "%s: null pointer", what);
  MMFN_CHECK_ARG(stride == 1 || stride == 2, "%s: stride must be 1 or 2", what);
  MMFN_CHECK_ARG(Co % EB == 0 && C % (EB == 64 ? 8 : 4) == 0, "%s: Co %% %d == 0 and 16-byte channel rows required", what, EB);
  MMFN_CHECK_ARG((((uintptr_t)dy | (uintptr_t)w) & 15) == 0, "%s: operands must be 16-byte aligned", what);
  if (stride == 2) return conv2d_dgrad_s2<EB>(dy, w, dx, res, N, H, W, C, Co, R, S, pad, Ho, Wo, stream, what);
  // stride 1: a forward convolution of dy (N,Ho,Wo,Co) -> dx (N,H,W,C) with mirrored taps and padding R-1-pad
  MMFN_CHECK_ARG(pad < R && pad < S && Ho == H + 2 * pad - R + 1 && Wo == W + 2 * pad - S + 1, "%s: inconsistent sizes", what);
  MMFN_CHECK_ARG(W >= 8 && H >= 8, "%s: dx must be at least 8x8", what);
  if (patch_conv_ok(R, S, 1, pad, H, W)) return launch_conv3x3_patch<true, EB>(dy, w, dx, res, N, H, W, Co, C, stream, what);
  ConvGeomTc g{N, Ho, Wo, Co, C, R, S, 1, R - 1 - pad, H, W};       // roles as a forward conv: in = dy, out = dx
  g.BW = W >= 16 ? 16 : 8;
  g.BH = 8;
  g.BI = tc::TBM / (g.BW * g.BH);
  g.tiles_w = (W + g.BW - 1) / g.BW; g.tiles_h = (H + g.BH - 1) / g.BH; g.tiles_n = (N + g.BI - 1) / g.BI;
  CUtensorMap ta, tb;
  if (int rc = make_act_tmap(&ta, dy, N, Ho, Wo, Co, g.BW, g.BH, g.BI, 1, false, EB)) return rc;
  const int tbn = (C <= 64 || g.tiles_w * g.tiles_h * g.tiles_n * ((C + 127) / 128) < 148) ? 64 : 128;
  if (int rc = make_krsc_b_tmap(&tb, w, Co, R, S, C, EB)) return rc;
  int ptiles = g.tiles_w * g.tiles_h * g.tiles_n;
  MMFN_CHECK_ARG(ptiles <= 65535, "%s: too many pixel tiles", what);
  const int nkb = R * S * (Co / EB);
  const int ctas = ptiles * ((C + tbn - 1) / tbn);
  int splitk = 1;
  if (ctas < 148) splitk = max(1, min(nkb / 8, (2 * 148) / ctas));
  int kb_per = (nkb + splitk - 1) / splitk;
  splitk = (nkb + kb_per - 1) / kb_per;
  if (splitk > 1) {
    cudaError_t ce = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * C, stream);
    if (ce != cudaSuccess) { mmfn_set_error("%s: memset: %s", what, cudaGetErrorString(ce)); return (int)ce; }
  }
  tc::Epilogue e{dx, nullptr, res, nullptr, 1.f, 0, splitk > 1 ? 2 : 0, 0.f, 0, mmfn_tc_trace_ptr()};
  if (tbn == 64) {
    ConvFwdOp<64, true, EB> op{g, kb_per};
    return tc::launch<ConvFwdOp<64, true, EB>, 64, 4>(ta, tb, op, e, dim3((C + 63) / 64, ptiles, splitk), stream, what);
  }
  ConvFwdOp<128, true, EB> op{g, kb_per};
  return tc::launch<ConvFwdOp<128, true, EB>, 128, 3>(ta, tb, op, e, dim3((C + 127) / 128, ptiles, splitk), stream, what);
}

template <int EB>
static int conv_wgrad_impl(const void* dy, const void* x, float* dw,
                           int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                           int Ho, int Wo, int splitk, cudaStream_t stream) {
  const char* what = EB == 64 ? "conv_wgrad_bf16" : "conv_wgrad_tf32";
  MMFN_CHECK_ARG(dy && x && dw, "%s: null pointer", what);
  ConvGeomTc g{N, H, W, C, Co, R, S, stride, pad, Ho, Wo};
  if (int rc = check_tc_geom(g, what, EB)) return rc;
  MMFN_CHECK_ARG(Co % (EB == 64 ? 8 : 4) == 0, "%s: dy channel rows must be multiples of 16 bytes", what);
  MMFN_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, "%s: operands must be 16-byte aligned", what);
  MMFN_CHECK_ARG(Wo >= 8 && Ho >= 4, "%s: output must be at least 4x8", what);
  g.BW = Wo >= 16 ? 16 : 8;                        // k-block = EB output pixels: a BW x BH patch of one image
  g.BH = EB / g.BW;
  g.BI = 1;
  g.tiles_w = (Wo + g.BW - 1) / g.BW; g.tiles_h = (Ho + g.BH - 1) / g.BH; g.tiles_n = N;
  CUtensorMap ta, tb;
  if (int rc = make_act_tmap(&ta, dy, N, Ho, Wo, Co, g.BW, g.BH, g.BI, 1, true, EB)) return rc;
  if (int rc = make_act_tmap(&tb, x, N, H, W, C, g.BW, g.BH, g.BI, stride, true, EB)) return rc;
  const int tbn = C <= 64 ? 64 : 128;
  int npb = g.tiles_w * g.tiles_h * g.tiles_n;
  int co_tiles = (Co + tc::TBM - 1) / tc::TBM, ci_tiles = (C + tbn - 1) / tbn;
  if (splitk <= 0) {
    int ctas = co_tiles * ci_tiles * R * S;
    splitk = max(1, min(npb / 8, (2 * 148) / ctas));
  }
  int pb_per = (npb + splitk - 1) / splitk;
  splitk = (npb + pb_per - 1) / pb_per;
  MMFN_CHECK_ARG(R * S * splitk <= 65535, "%s: too many splits", what);
  tc::Epilogue e{dw, nullptr, nullptr, nullptr, 1.f, 0, 2, 0.f, 0, mmfn_tc_trace_ptr()};
  dim3 grid(ci_tiles, co_tiles, R * S * splitk);
  if (tbn == 64) {
    ConvWgradOp<64, EB> op{g, splitk, pb_per};
    return tc::launch<ConvWgradOp<64, EB>, 64, 4>(ta, tb, op, e, grid, stream, what);
  }
  ConvWgradOp<128, EB> op{g, splitk, pb_per};
  return tc::launch<ConvWgradOp<128, EB>, 128, 3>(ta, tb, op, e, grid, stream, what);
}

// y(N,Ho,Wo,Co) = conv(x(N,H,W,C), w(Co,R,S,C)) [+ res]; TF32 multiply, FP32 accumulate.  C % 32 == 0.
MMFN_API int mmfn_conv2d_fwd_tf32(const float* x, const float* w, float* y, const float* res,
                                  int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                  int Ho, int Wo, cudaStream_t stream) {
  return conv_fwd_impl<32>(x, w, y, res, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream);
}

// mmfn_conv2d_fwd_tf32 + the train-mode BatchNorm statistics of its output (conv -> BatchNorm2d of a torchvision
// BasicBlock): per-channel mean / rstd (biased variance, eps) and the momentum update of running_mean / running_var
// (unbiased variance; nullable) are produced by the convolution's own epilogue + last-CTA finalize -- no separate pass
// over y.  ws: the BatchNorm scratch of mmfn_bn_train_fwd (zero on entry, left zero).  Follow with mmfn_bn_apply.
MMFN_API int mmfn_conv2d_fwd_bn_tf32(const float* x, const float* w, float* y,
                                     int N, int H, int W, int C, int Co, int R, int S, int stride, int pad, int Ho, int Wo,
                                     double* ws, float* mean, float* rstd, float* running_mean, float* running_var,
                                     float momentum, float eps, cudaStream_t stream) {
  BnStatsArgs bn{ws, mean, rstd, running_mean, running_var, momentum, eps};
  return conv_fwd_impl<32>(x, w, y, nullptr, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream, &bn);
}

// bf16-operand variant of mmfn_conv2d_fwd_bn_tf32 (x and w BF16, y fp32).
MMFN_API int mmfn_conv2d_fwd_bn_bf16(const void* x, const void* w, float* y,
                                     int N, int H, int W, int C, int Co, int R, int S, int stride, int pad, int Ho, int Wo,
                                     double* ws, float* mean, float* rstd, float* running_mean, float* running_var,
                                     float momentum, float eps, cudaStream_t stream) {
  BnStatsArgs bn{ws, mean, rstd, running_mean, running_var, momentum, eps};
  return conv_fwd_impl<64>(x, w, y, nullptr, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream, &bn);
}

// Data gradient of mmfn_conv2d_fwd_tf32 straight from the KRSC filters (see conv_dgrad_impl): stride 1 or 2.
MMFN_API int mmfn_conv2d_dgrad_tf32(const float* dy, const float* w, float* dx, const float* res,
                                    int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                    int Ho, int Wo, cudaStream_t stream) {
  return conv_dgrad_impl<32>(dy, w, dx, res, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream);
}

// dw(Co,R,S,C) += dy^T * im2col(x), atomically; TF32 multiply, FP32 accumulate.  C % 32 == 0, Co % 4 == 0.
MMFN_API int mmfn_conv2d_wgrad_tf32(const float* dy, const float* x, float* dw,
                                    int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                    int Ho, int Wo, int splitk, cudaStream_t stream) {
  return conv_wgrad_impl<32>(dy, x, dw, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, splitk, stream);
}

// BASELINE configs[2] (torch.autocast(bfloat16) over the torchvision BasicBlock convolutions, model_rad.py:523-525,
// :542-544, :560-562, :577-579): x (N,H,W,C) and the KRSC filter shadow w are BF16, multiplied with tcgen05 kind::f16,
// accumulated in fp32; y (pre-BatchNorm, feeds the batch statistics) and res are fp32.  C % 64 == 0.
MMFN_API int mmfn_conv2d_fwd_bf16(const void* x, const void* w, float* y, const float* res,
                                  int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                  int Ho, int Wo, cudaStream_t stream) {
  return conv_fwd_impl<64>(x, w, y, res, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream);
}

// dx (fp32, + res) from dy (BF16, written by the BatchNorm backward) and the BF16 filter shadow.  Co % 64 == 0.
MMFN_API int mmfn_conv2d_dgrad_bf16(const void* dy, const void* w, float* dx, const float* res,
                                    int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                    int Ho, int Wo, cudaStream_t stream) {
  return conv_dgrad_impl<64>(dy, w, dx, res, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, stream);
}

// dw (fp32 master gradient, accumulated atomically) from dy and x in BF16; k-blocks of 64 output pixels.  C % 64 == 0.
MMFN_API int mmfn_conv2d_wgrad_bf16(const void* dy, const void* x, float* dw,
                                    int N, int H, int W, int C, int Co, int R, int S, int stride, int pad,
                                    int Ho, int Wo, int splitk, cudaStream_t stream) {
  return conv_wgrad_impl<64>(dy, x, dw, N, H, W, C, Co, R, S, stride, pad, Ho, Wo, splitk, stream);
}

MMFN_DEFINE_RNG_BINDER(conv_tc)
