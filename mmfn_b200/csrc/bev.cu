// LiDAR point cloud -> 2-channel BEV pillar histogram (reference:
// team_code/mmfn_utils/datasets/dataloader.py:271-293 lidar_to_histogram_features).
//
// One CTA owns one (frame, x-strip) slab of the 256x256 grid -- BOTH height channels -- and keeps its
// counters in shared memory as packed u16 fields; every CTA streams the frame's points ONCE (L2-resident
// after the first CTA touches them; one pass serves both channels), bins them with integer arithmetic only,
// and finally writes its two channel slabs once, coalesced, already clamped and scaled.  DRAM traffic is
// therefore the algorithmic N*stride*4 + 2*256*256*4 bytes per frame; there are no global atomics and no
// separate memset/finalize.  L2 read amplification = number of strips (chosen so that ~256 CTAs exist).
#include "common.cuh"

namespace {

constexpr int GRID = 256;

// clamp(count, 5) / 5 as the reference computes it: float32(k / 5.0) == float32(k) * float32(0.2) for k = 0..5
__device__ __forceinline__ float fifth(uint32_t count) { return __fmul_rn((float)min(count, 5u), 0.2f); }

template <int STRIPS>
__global__ void __launch_bounds__(1024)
bev_scatter_kernel(const float* __restrict__ pts, int n_pts, int pt_stride,
                   float* __restrict__ out) {
  constexpr int ROWS = GRID / STRIPS;              // x-bins owned by this CTA
  constexpr int WORDS = ROWS * GRID / 2;           // words per channel: two u16 counters per word
  extern __shared__ uint32_t cnt[];                // [2 channels][WORDS]
  const int strip = blockIdx.x % STRIPS;
  const int frame = blockIdx.x / STRIPS;
  for (int i = threadIdx.x; i < 2 * WORDS; i += blockDim.x) cnt[i] = 0u;
  __syncthreads();

  const float* p = pts + (int64_t)frame * n_pts * pt_stride;
  const int x_lo = strip * ROWS;
  // Every CTA of a frame scans all of its points, so the common case -- a point outside this strip -- must cost two
  // compares: ix = floor(x*8) + 128 lies in [x_lo, x_lo + ROWS) iff x lies in [(x_lo-128)/8, (x_lo+ROWS-128)/8) (x*8
  // and the bounds are exact in fp32); the inclusive right edge x == 16 belongs to the last strip.  NaN fails both.
  const float xs_lo = (float)(x_lo - 128) * 0.125f, xs_hi = (float)(x_lo + ROWS - 128) * 0.125f;
  const bool last_strip = strip == STRIPS - 1;
  // eight points per thread are requested before the first is binned: the loop is latency-bound otherwise
  // (one dependent L2 round trip per point and thread)
  constexpr int U = 8;
  for (int i0 = threadIdx.x; i0 < n_pts; i0 += U * blockDim.x) {
    float px[U], py[U], pz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * blockDim.x;
      px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);          // NaN: rejected below
      if (i < n_pts) {
        if (pt_stride == 4) {
          float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
          px[u] = v.x; py[u] = v.y; pz[u] = v.z;
        } else {
          const float* q = p + (int64_t)i * pt_stride;
          px[u] = __ldg(q); py[u] = __ldg(q + 1); pz[u] = __ldg(q + 2);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float x = px[u], y = py[u], z = pz[u];
      if (!(x >= xs_lo && (x < xs_hi || (last_strip && x <= 16.0f)))) continue;
      // closed range test also rejects NaN; x*8 / y*8 are exact in fp32.
      if (!(y >= -24.0f && y <= 8.0f)) continue;
      // channel 0: z <= -2, channel 1: z > -2; NaN z matches neither.
      const bool lo = z <= -2.0f, hi = z > -2.0f;
      if (!(lo || hi)) continue;
      int ix = (int)floorf(x * 8.0f) + 128;
      int iy = (int)floorf(y * 8.0f) + 192;
      ix = min(ix, GRID - 1);                        // right-most edge is inclusive
      iy = min(iy, GRID - 1);
      ix -= x_lo;
      if (ix < 0 || ix >= ROWS) continue;
      const int bin = ix * GRID + iy;
      uint32_t* word = cnt + (hi ? WORDS : 0) + (bin >> 1);
      const uint32_t shift = (bin & 1) * 16;
      // counts are clamped at 5 downstream: stop incrementing once a field reached 5 so
      // a u16 field can never carry into its neighbour (<= 4 + blockDim.x increments).
      if (((*(volatile uint32_t*)word >> shift) & 0xffffu) >= 5u) continue;
      atomicAdd(word, 1u << shift);
    }
  }
  __syncthreads();

  // float32(k / 5.0) for k = 0..5 == float32(k) * 0.2f bit for bit (checked for all six values); a dynamically indexed
  // local array would live in local memory: 256 dependent LDLs per thread made the conversion the slowest part.
#pragma unroll
  for (int chan = 0; chan < 2; ++chan) {
    float* o = out + (((int64_t)frame * 2 + chan) * GRID + x_lo) * GRID;
    const uint32_t* c = cnt + chan * WORDS;
    for (int i = threadIdx.x; i < WORDS / 2; i += blockDim.x) {       // 4 bins = one 16-byte store per thread
      const uint2 w = reinterpret_cast<const uint2*>(c)[i];
      reinterpret_cast<float4*>(o)[i] = make_float4(fifth(w.x & 0xffffu), fifth(w.x >> 16),
                                                    fifth(w.y & 0xffffu), fifth(w.y >> 16));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One-visit variant (two launches).  The strip kernel above re-reads every point once per strip: with ~256 CTAs that is
// 8-16 passes over the sweep through L2 (134 MB at 16 or 64 frames) and the kernel runs at the L2 read rate, ~25 us
// whatever the frame count (profiles/r02_bev_bench_strip_v2b.json).  Here every point is read ONCE:
//   1. bev_count_kernel: a CTA owns a chunk of one frame's points and counts them with packed-u16 `red.global` atomics
//      into a per-frame counter grid that lives in L2 (2 x 256 x 256 u16 = 256 KB per frame);
//   2. bev_convert_kernel: one thread per 8 bins turns the counters into the clamped / scaled fp32 grid and re-zeroes
//      them, so the scratch is zero again for the next call (no memset).
// (A single launch whose last CTA per frame converted the frame measured 38 us at 16 frames, 31 us of it in that
// one-SM-per-frame tail; counting alone takes 6.7 us -- profiles/r02_bev_variants.txt.)
// A u16 field cannot overflow: at most 5 are added per warp instruction with the warp-level dedupe, 1 per point without.
// ws layout: [frames][65536] u32 counters (two bins per word).
constexpr int CHUNK_THREADS = 512, CHUNK_PTS = 4;       // 2048 points per CTA, all loads in flight

template <bool DEDUPE>
__global__ void __launch_bounds__(CHUNK_THREADS)
bev_count_kernel(const float* __restrict__ pts, int n_pts, int pt_stride, int chunks, uint32_t* __restrict__ ws) {
  const int frame = blockIdx.x / chunks, chunk = blockIdx.x - frame * chunks;
  uint32_t* cnt = ws + (int64_t)frame * (GRID * GRID);                 // [2 channels][65536 bins / 2]
  const float* p = pts + (int64_t)frame * n_pts * pt_stride;
  const int base = chunk * (CHUNK_THREADS * CHUNK_PTS);
  float px[CHUNK_PTS], py[CHUNK_PTS], pz[CHUNK_PTS];
#pragma unroll
  for (int u = 0; u < CHUNK_PTS; ++u) {
    const int i = base + u * CHUNK_THREADS + threadIdx.x;
    px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);               // NaN: rejected below
    if (i < n_pts) {
      if (pt_stride == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        px[u] = v.x; py[u] = v.y; pz[u] = v.z;
      } else {
        const float* q = p + (int64_t)i * pt_stride;
        px[u] = __ldg(q); py[u] = __ldg(q + 1); pz[u] = __ldg(q + 2);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CHUNK_PTS; ++u) {
    const float x = px[u], y = py[u], z = pz[u];
    // closed range test also rejects NaN; x*8 / y*8 are exact in fp32.  channel 0: z <= -2, channel 1: z > -2 (NaN: neither)
    const bool lo = z <= -2.0f, hi = z > -2.0f;
    int bin = -1;
    if (x >= -16.0f && x <= 16.0f && y >= -24.0f && y <= 8.0f && (lo || hi)) {
      int ix = (int)floorf(x * 8.0f) + 128;
      int iy = (int)floorf(y * 8.0f) + 192;
      ix = min(ix, GRID - 1);                                          // right-most edge is inclusive
      iy = min(iy, GRID - 1);
      bin = (hi ? GRID * GRID : 0) + ix * GRID + iy;
    }
    if constexpr (DEDUPE) {
      // dense pillars: the lanes of a warp that hit the same bin elect ONE leader, which adds min(count, 5)
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1))
        atomicAdd(cnt + (bin >> 1), (uint32_t)min(__popc(peers), 5) << ((bin & 1) * 16));
    } else {
      if (bin >= 0) atomicAdd(cnt + (bin >> 1), 1u << ((bin & 1) * 16));   // result unused: red.global.add
    }
  }
}

// counters -> clamp(count, 5) / 5 as fp32, counters re-zeroed: one thread per 4 words (8 bins, two 16-byte stores)
__global__ void __launch_bounds__(256)
bev_convert_kernel(uint4* __restrict__ cnt, float4* __restrict__ out, int64_t n4) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uint4 w = __ldcg(cnt + i);
  cnt[i] = make_uint4(0u, 0u, 0u, 0u);
  out[2 * i] = make_float4(fifth(w.x & 0xffffu), fifth(w.x >> 16), fifth(w.y & 0xffffu), fifth(w.y >> 16));
  out[2 * i + 1] = make_float4(fifth(w.z & 0xffffu), fifth(w.z >> 16), fifth(w.w & 0xffffu), fifth(w.w >> 16));
}

}  // namespace

// One-visit BEV scatter (see bev_count_kernel / bev_convert_kernel).  ws: frames * 65536 u32 of scratch that is ZERO on
// entry (zero it once after allocation; every call leaves it zero again) -- mmfn_workspace_bytes(MMFN_WS_BEV).
MMFN_API int mmfn_bev_scatter_ws(const float* pts, int frames, int n_pts, int pt_stride,
                                 float* out, void* ws, cudaStream_t stream) {
  MMFN_CHECK_ARG(out && ws && (pts || n_pts == 0 || frames == 0), "bev_scatter_ws: null pointer");
  MMFN_CHECK_ARG(frames >= 0 && n_pts >= 0 && n_pts < 65536, "bev_scatter_ws: 0 <= n_pts < 65536 (u16 pillar counters)");
  MMFN_CHECK_ARG(pt_stride >= 3, "bev_scatter_ws: pt_stride must be >= 3 (x,y,z,...)");
  MMFN_CHECK_ARG(pt_stride != 4 || ((uintptr_t)pts & 15) == 0, "bev_scatter_ws: xyzi rows must be 16B aligned");
  MMFN_CHECK_ARG((((uintptr_t)out | (uintptr_t)ws) & 15) == 0, "bev_scatter_ws: out / ws must be 16-byte aligned");
  if (frames == 0) return 0;
  const int per = CHUNK_THREADS * CHUNK_PTS;
  const int chunks = n_pts > 0 ? (n_pts + per - 1) / per : 0;
  MMFN_CHECK_ARG((int64_t)frames * chunks <= 0x7fffffff, "bev_scatter_ws: too many CTAs");
  // (the warp-level dedupe variant <true> measured 12.4 vs 11.9 us at 16 frames: __match_any_sync costs what it saves)
  if (chunks > 0)
    bev_count_kernel<false><<<frames * chunks, CHUNK_THREADS, 0, stream>>>(pts, n_pts, pt_stride, chunks, static_cast<uint32_t*>(ws));
  const int64_t n4 = (int64_t)frames * (GRID * GRID / 4);
  bev_convert_kernel<<<(unsigned)ceil_div64(n4, 256), 256, 0, stream>>>(static_cast<uint4*>(ws), reinterpret_cast<float4*>(out), n4);
  return mmfn_launch_status("bev_scatter_ws");
}

// pts: (frames, n_pts, pt_stride) f32 device; out: (frames, 2, 256, 256) f32 device.
MMFN_API int mmfn_bev_scatter(const float* pts, int frames, int n_pts, int pt_stride,
                              float* out, int strips, cudaStream_t stream) {
  MMFN_CHECK_ARG(out && (pts || n_pts == 0 || frames == 0), "bev_scatter: null pointer");
  MMFN_CHECK_ARG(frames >= 0 && n_pts >= 0, "bev_scatter: negative size");
  MMFN_CHECK_ARG(pt_stride >= 3, "bev_scatter: pt_stride must be >= 3 (x,y,z,...)");
  MMFN_CHECK_ARG(pt_stride != 4 || ((uintptr_t)pts & 15) == 0, "bev_scatter: xyzi rows must be 16B aligned");
  if (frames == 0) return 0;
  if (strips <= 0) {                                // ~256 CTAs (two 1024-thread CTAs per SM): 148 SMs covered, one wave
    strips = frames >= 64 ? 4 : frames >= 24 ? 8 : 16;
  }
  MMFN_CHECK_ARG(((uintptr_t)out & 15) == 0, "bev_scatter: out must be 16-byte aligned");
  dim3 grid(frames * strips);
  size_t smem = (size_t)2 * (GRID / strips) * GRID / 2 * sizeof(uint32_t);      // both channels of the strip
#define MMFN_BEV_CASE(S)                                                                         \
  case S: {                                                                                      \
    cudaError_t ce = cudaFuncSetAttribute(bev_scatter_kernel<S>,                                 \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (ce != cudaSuccess) { mmfn_set_error("bev_scatter: smem attr: %s", cudaGetErrorString(ce)); return (int)ce; } \
    bev_scatter_kernel<S><<<grid, 1024, smem, stream>>>(pts, n_pts, pt_stride, out);             \
  } break;
  switch (strips) {
    MMFN_BEV_CASE(2) MMFN_BEV_CASE(4) MMFN_BEV_CASE(8) MMFN_BEV_CASE(16)
    case 1: mmfn_set_error("bev_scatter: strips = 1 needs 256 KB of shared memory (both channels of a frame); use 2, 4, 8 or 16"); return MMFN_BAD_ARG;
    default: mmfn_set_error("bev_scatter: strips must be 2, 4, 8 or 16"); return MMFN_BAD_ARG;
  }
#undef MMFN_BEV_CASE
  return mmfn_launch_status("bev_scatter");
}
