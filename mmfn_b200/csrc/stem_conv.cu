// Direct 7x7 / stride 2 / pad 3 stem convolutions (3 RGB or 2 BEV channels -> 64; torchvision ResNet conv1 as called at
// model_rad.py:512 image, :518 LiDAR) and their weight gradient, WITHOUT a column matrix in HBM.
//
// The im2col + GEMM formulation moved 4x the algorithmic bytes (a (B*128*128) x 192 column matrix written, read by the
// GEMM and read again by the weight gradient).  Here a CTA stages the input patch of a 2 x 64 output tile in shared
// memory (9 x 133 x C floats) and gathers the implicit-GEMM operand from it: with NHWC input the (s, c) taps of one
// filter row are CONTIGUOUS in a patch row, so the K index k = r * 7C + (s C + c) maps to the patch offset
// r * row_stride + (k mod 7C) on top of the pixel's base 2 oh * row_stride + 2 ow * C.  K = 49 C is padded to a
// multiple of 8 with zero filter taps.  Contractions: warp-level mma.sync m16n8k8 TF32 (operands rounded to nearest
// TF32 when staged), fp32 accumulate -- K = 147 / 98 and 64 output channels are too small for 128-wide tcgen05 tiles
// to pay, and the forward is bound by writing its 64-channel output.  HBM traffic = input + output (+ filters).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int CO = 64;                          // output channels of both stems
constexpr int TH = 2, TW = 64, TP = TH * TW;    // output tile: 2 rows x 64 columns = 128 pixels
constexpr int NTHREADS = 256;

template <int C>
struct Geom {
  static constexpr int RC = 7 * C;              // one filter row
  static constexpr int K = 49 * C;
  static constexpr int KS = (K + 7) / 8;        // k-steps of 8
  static constexpr int KP = KS * 8;
  static constexpr int PH = 2 * TH + 5;         // 9 patch rows
  static constexpr int PS = (2 * TW + 5) * C;   // floats per patch row (133 C) = row stride
  static constexpr int WS = KP + 4;             // filter row stride in shared memory: 156 / 108 = 28 / 12 (mod 32) -> the
                                                // B fragments (8 filters x 4 taps per load) hit 32 distinct banks
  static constexpr int DS = CO + 8;             // gradient tile row stride (weight gradient): 72 = 8 (mod 32)
  static constexpr int PWORDS = (PH * PS + 3) / 4 * 4;   // patch words, rounded up: what follows is accessed as 16-byte vectors
  __device__ static __forceinline__ int koff(int k) { return k < K ? (k / RC) * PS + (k % RC) : 0; }
};

// The zero-padded input patch of a tile travels global -> registers -> shared memory in two halves, so that the loads
// of the NEXT tile are in flight while the current tile's MMAs run (all loads of a thread are independent).
template <int C>
struct Patch {
  using G = Geom<C>;
  static constexpr int NV = (G::PH * G::PS + NTHREADS - 1) / NTHREADS;
  float v[NV];
  __device__ __forceinline__ void fetch(const float* __restrict__ x, int tile, int tiles_w, int tiles_h, int H, int W) {
    const int tw = tile % tiles_w, t2 = tile / tiles_w;
    const int th = t2 % tiles_h, n = t2 / tiles_h;
    const int ih0 = 2 * th * TH - 3, iwc0 = (2 * tw * TW - 3) * C, WC = W * C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * NTHREADS;
      const int pr = i / G::PS, pc = i - pr * G::PS;
      const int ih = ih0 + pr, iwc = iwc0 + pc;
      v[j] = (i < G::PH * G::PS && ih >= 0 && ih < H && iwc >= 0 && iwc < WC) ? __ldg(x + ((int64_t)n * H + ih) * WC + iwc) : 0.f;
    }
  }
  __device__ __forceinline__ void stage(uint32_t* __restrict__ Psm) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * NTHREADS;
      if (i < G::PH * G::PS) Psm[i] = to_tf32(v[j]);
    }
  }
};

// z (N, Ho, Wo, 64) = conv(x (N, H, W, C), w (64, 7, 7, C)).  Persistent CTAs (filters staged once), 8 warps: warp =
// 32 pixels x 32 channels (2 x 4 MMA tiles); per k-step 8 gathered A words + 8 filter words feed 8 MMAs.
template <int C>
__global__ void __launch_bounds__(NTHREADS, 2)
stem_conv7_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ z,
                      int H, int W, int Ho, int Wo, int tiles_w, int tiles_h, int ntiles) {
  using G = Geom<C>;
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* Wsm = sm;                            // [64][WS]
  uint32_t* Psm = sm + CO * G::WS;               // [PH][PS]
  for (int i = threadIdx.x; i < CO * G::KP; i += NTHREADS) {
    const int n = i / G::KP, k = i - n * G::KP;
    Wsm[n * G::WS + k] = k < G::K ? to_tf32(__ldg(w + n * G::K + k)) : 0u;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  // patch offsets of a thread's two k columns (t, t + 4) per k-step, 16 bits each: one shared-memory word per k-step
  __shared__ uint32_t Ksm[G::KS * 4];
  if (threadIdx.x < G::KS * 4) {
    const int ks = threadIdx.x >> 2, tt = threadIdx.x & 3;
    Ksm[threadIdx.x] = (uint32_t)G::koff(ks * 8 + tt) | ((uint32_t)G::koff(ks * 8 + tt + 4) << 16);
  }
  int base[2][2];                                // patch offset of pixel rows g / g + 8 of the warp's two m-tiles
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = wm * 32 + mt * 16 + g + h * 8;
      base[mt][h] = 2 * (p / TW) * G::PS + 2 * (p % TW) * C;
    }
  }
  const uint32_t* wrow = Wsm + (wn * 32 + g) * G::WS + t;
  Patch<C> patch;
  if ((int)blockIdx.x < ntiles) patch.fetch(x, blockIdx.x, tiles_w, tiles_h, H, W);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, t2 = tile / tiles_w;
    const int th = t2 % tiles_h, n = t2 / tiles_h;
    const int oh0 = th * TH, ow0 = tw * TW;
    __syncthreads();                              // the previous tile's fragments are read (and the filters are staged)
    patch.stage(Psm);
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) patch.fetch(x, tile + gridDim.x, tiles_w, tiles_h, H, W);
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll 4
    for (int ks = 0; ks < G::KS; ++ks) {
      uint32_t a[2][4], b[4][2];
      const uint32_t kk = Ksm[ks * 4 + t];
      const int k0 = (int)(kk & 0xffffu), k1 = (int)(kk >> 16);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        a[mt][0] = Psm[base[mt][0] + k0];
        a[mt][1] = Psm[base[mt][1] + k0];
        a[mt][2] = Psm[base[mt][0] + k1];
        a[mt][3] = Psm[base[mt][1] + k1];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = wrow[nt * 8 * G::WS + ks * 8];
        b[nt][1] = wrow[nt * 8 * G::WS + ks * 8 + 4];
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt][0], b[nt][1]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = wm * 32 + mt * 16 + g + h * 8;
        const int oh = oh0 + p / TW, ow = ow0 + p % TW;
        if (oh >= Ho || ow >= Wo) continue;
        float* zp = z + (((int64_t)n * Ho + oh) * Wo + ow) * CO + wn * 32 + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<float2*>(zp + nt * 8) = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
      }
    }
  }
}

__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) { return mmfn_unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(p))); }

// dw (64, 7, 7, C) += dz^T (64 x pixels) * patches (pixels x K).  Each persistent CTA keeps the whole 64 x KP result in
// registers (warp = 32 filters x a quarter of the k-tiles) while it walks its tiles -- the reduction dimension is the
// pixels, 8 per MMA -- and adds it to dw with one atomic per element at the end.
template <int C, typename TDZ>
__global__ void __launch_bounds__(NTHREADS, 2)
stem_conv7_wgrad_kernel(const float* __restrict__ x, const TDZ* __restrict__ dz, float* __restrict__ dw,
                        int H, int W, int Ho, int Wo, int tiles_w, int tiles_h, int ntiles) {
  using G = Geom<C>;
  constexpr int NTG = (G::KS + 3) / 4;           // k-tiles per warp group
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* Psm = sm;                            // [PH][PS]
  uint32_t* Dsm = sm + G::PWORDS;                // [128 pixels][DS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mp = warp & 1, ng = warp >> 1;
  int kb[NTG];
#pragma unroll
  for (int j = 0; j < NTG; ++j) kb[j] = G::koff(min((ng * NTG + j) * 8 + g, G::KP - 1));
  float acc[2][NTG][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < NTG; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
  Patch<C> patch;
  if ((int)blockIdx.x < ntiles) patch.fetch(x, blockIdx.x, tiles_w, tiles_h, H, W);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, t2 = tile / tiles_w;
    const int th = t2 % tiles_h, n = t2 / tiles_h;
    const int oh0 = th * TH, ow0 = tw * TW;
    __syncthreads();
    patch.stage(Psm);
    if (tile + (int)gridDim.x < ntiles) patch.fetch(x, tile + gridDim.x, tiles_w, tiles_h, H, W);
#pragma unroll 4
    for (int i = threadIdx.x; i < TP * (CO / 4); i += NTHREADS) {
      const int p = i / (CO / 4), c4 = i % (CO / 4);
      const int oh = oh0 + p / TW, ow = ow0 + p % TW;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oh < Ho && ow < Wo) v = load4(dz + (((int64_t)n * Ho + oh) * Wo + ow) * CO + c4 * 4);
      *reinterpret_cast<uint4*>(Dsm + p * G::DS + c4 * 4) = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
    __syncthreads();
#pragma unroll 4
    for (int ps = 0; ps < TP / 8; ++ps) {
      const int p0 = ps * 8 + t, p1 = p0 + 4;
      const int po0 = 2 * (p0 / TW) * G::PS + 2 * (p0 % TW) * C, po1 = 2 * (p1 / TW) * G::PS + 2 * (p1 % TW) * C;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int ch = (mp * 2 + mt) * 16 + g;
        a[mt][0] = Dsm[p0 * G::DS + ch];
        a[mt][1] = Dsm[p0 * G::DS + ch + 8];
        a[mt][2] = Dsm[p1 * G::DS + ch];
        a[mt][3] = Dsm[p1 * G::DS + ch + 8];
      }
#pragma unroll
      for (int j = 0; j < NTG; ++j) {
        if ((ng * NTG + j) < G::KS) {              // warp-uniform
          const uint32_t b0 = Psm[po0 + kb[j]], b1 = Psm[po1 + kb[j]];
          mma_tf32(acc[0][j], a[0], b0, b1);
          mma_tf32(acc[1][j], a[1], b0, b1);
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int j = 0; j < NTG; ++j) {
      const int nt = ng * NTG + j;
      if (nt >= G::KS) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = (mp * 2 + mt) * 16 + g + (e >> 1) * 8, k = nt * 8 + 2 * t + (e & 1);
        if (k < G::K) atomicAdd(dw + ch * G::K + k, acc[mt][j][e]);
      }
    }
  }
}

template <int C> constexpr int fwd_smem() { return (CO * Geom<C>::WS + Geom<C>::PH * Geom<C>::PS) * 4; }
template <int C> constexpr int wgrad_smem() { return (Geom<C>::PWORDS + TP * Geom<C>::DS) * 4; }

template <typename Kern>
int opt_in_smem(Kern kern, int bytes) {
  return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

// z (N, Ho, Wo, 64) = conv7x7 / stride 2 / pad 3 of x (N, H, W, C), C = 3 (image stem, model_rad.py:512) or 2 (LiDAR BEV
// stem, :518), filters w_krsc (64, 7, 7, C); Ho = (H - 1) / 2 + 1.  TF32 multiply (operands rounded to nearest), fp32
// accumulate; no column matrix: HBM traffic is x + z.
MMFN_API int mmfn_conv2d_stem7_fwd(const float* x, const float* w_krsc, float* z, int N, int H, int W, int C,
                                   cudaStream_t stream) {
  MMFN_CHECK_ARG(x && w_krsc && z, "conv2d_stem7_fwd: null pointer");
  MMFN_CHECK_ARG(N > 0 && H > 0 && W > 0 && (C == 2 || C == 3), "conv2d_stem7_fwd: C must be 2 or 3");
  MMFN_CHECK_ARG(((uintptr_t)z & 7) == 0, "conv2d_stem7_fwd: z must be 8-byte aligned");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
  const int64_t nt = (int64_t)N * tiles_w * tiles_h;
  MMFN_CHECK_ARG(nt < (int64_t)1 << 31, "conv2d_stem7_fwd: too many tiles");
  const int grid = (int)(nt < 296 ? nt : 296);
  int rc;
  if (C == 3) {
    if ((rc = opt_in_smem(stem_conv7_fwd_kernel<3>, fwd_smem<3>()))) return rc;
    stem_conv7_fwd_kernel<3><<<grid, NTHREADS, fwd_smem<3>(), stream>>>(x, w_krsc, z, H, W, Ho, Wo, tiles_w, tiles_h, (int)nt);
  } else {
    if ((rc = opt_in_smem(stem_conv7_fwd_kernel<2>, fwd_smem<2>()))) return rc;
    stem_conv7_fwd_kernel<2><<<grid, NTHREADS, fwd_smem<2>(), stream>>>(x, w_krsc, z, H, W, Ho, Wo, tiles_w, tiles_h, (int)nt);
  }
  return mmfn_launch_status("conv2d_stem7_fwd");
}

// dw_krsc (64, 7, 7, C) += weight gradient of the stem convolution from dz (N, Ho, Wo, 64) (fp32, or bf16 when dz_bf16)
// and the network input x (N, H, W, C).  Reads dz and x once; atomic accumulation into dw (fp32).
MMFN_API int mmfn_conv2d_stem7_wgrad(const void* dz, int dz_bf16, const float* x, float* dw_krsc, int N, int H, int W, int C,
                                     cudaStream_t stream) {
  MMFN_CHECK_ARG(dz && x && dw_krsc, "conv2d_stem7_wgrad: null pointer");
  MMFN_CHECK_ARG(N > 0 && H > 0 && W > 0 && (C == 2 || C == 3), "conv2d_stem7_wgrad: C must be 2 or 3");
  MMFN_CHECK_ARG(((uintptr_t)dz & (dz_bf16 ? 7 : 15)) == 0, "conv2d_stem7_wgrad: dz alignment");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
  const int64_t nt = (int64_t)N * tiles_w * tiles_h;
  MMFN_CHECK_ARG(nt < (int64_t)1 << 31, "conv2d_stem7_wgrad: too many tiles");
  const int grid = (int)(nt < 296 ? nt : 296);
  int rc;
#define MMFN_STEM_WGRAD(CC, T)                                                                                          \
  do {                                                                                                                  \
    if ((rc = opt_in_smem(stem_conv7_wgrad_kernel<CC, T>, wgrad_smem<CC>()))) return rc;                                \
    stem_conv7_wgrad_kernel<CC, T><<<grid, NTHREADS, wgrad_smem<CC>(), stream>>>(x, (const T*)dz, dw_krsc, H, W, Ho, Wo, \
                                                                                 tiles_w, tiles_h, (int)nt);            \
  } while (0)
  if (C == 3 && dz_bf16) MMFN_STEM_WGRAD(3, __nv_bfloat16);
  else if (C == 3) MMFN_STEM_WGRAD(3, float);
  else if (dz_bf16) MMFN_STEM_WGRAD(2, __nv_bfloat16);
  else MMFN_STEM_WGRAD(2, float);
#undef MMFN_STEM_WGRAD
  return mmfn_launch_status("conv2d_stem7_wgrad");
}
