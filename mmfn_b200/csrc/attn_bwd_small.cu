// Attention backward of fusion transformers 1-3 (heads of 16 / 32 / 64 dims, T = 128 / 192 tokens) in ONE launch,
// bf16 configuration (reference: autograd through SelfAttention.forward, model_rad.py:96-105).
//
// The per-op chain is five launches per block -- dPd = dY V^T, softmax backward, dQ = dS K on the critical path, dV =
// Pd^T dY and dK = dS^T Q on fork streams and a join -- ~33 us of latency per block for a few MFLOP.  Here one CTA owns
// a (sample, head): q, k, v, dy (T x hs) and the saved P, Pd (T x T, bf16) are copied into shared memory with coalesced
// cp.async, then
//   phase 1, warp = 16 queries:  dPd = dy v^T (mma.sync, scores in registers), dS = scale (dPd o Pd - P rowsum(dPd o Pd))
//            -- the dropout mask never has to be regenerated: dP o P = dPd o Pd -- written over P in shared memory;
//            dq = dS k from the same registers;
//   phase 2, warp = 16 keys:     dk = dS^T q and dv = Pd^T dy, the transposed operands fetched with ldmatrix.trans.
// No atomics, no cross-CTA traffic, deterministic.  TF32 (fp32 P: 2 x 147 KB) does not fit and keeps the per-op chain.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void ldm_x4(uint32_t saddr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldm_x4_t(uint32_t saddr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// T = 16 NQ tokens, NQ warps.  Shared memory (bf16): q, k, v, dy [T][HS + 8]; SB (P, then dS) and PB (Pd) [T][T + 8]
// (row strides are odd multiples of 16 bytes: ldmatrix and the fragment accesses are conflict-free).
template <int HS, int NQ>
__global__ void __launch_bounds__(NQ * 32, 1)
attn_bwd_small_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ P,
                      const __nv_bfloat16* __restrict__ Pd, __nv_bfloat16* __restrict__ dqkv, int C, int nh, float scale) {
  constexpr int T = NQ * 16, LH = HS + 8, LT = T + 8, NTH = NQ * 32;
  // 64-dim heads: q, k, v do not fit next to the two T x T tiles -- they take turns in ONE buffer (v for dPd, then k for
  // dQ, then q for dK), two extra L2 round trips per CTA
  constexpr bool STAGED = HS > 32;
  extern __shared__ __align__(16) __nv_bfloat16 sm[];
  __nv_bfloat16* ds = sm;
  __nv_bfloat16* vs = ds + T * LH;
  __nv_bfloat16* ks = STAGED ? vs : vs + T * LH;
  __nv_bfloat16* qs = STAGED ? vs : ks + T * LH;
  __nv_bfloat16* SB = (STAGED ? vs : qs) + T * LH;
  __nv_bfloat16* PB = SB + T * LT;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)b * T;
  // ---- loads (16-byte chunks)
  {
    constexpr int CH = HS / 8;                            // chunks per head row
    for (int i = threadIdx.x; i < T * CH; i += NTH) {
      const int r = i / CH, c = i - r * CH;
      const __nv_bfloat16* src = qkv + (row0 + r) * 3 * C + h * HS + c * 8;
      if constexpr (!STAGED) {
        cp_async16(ks + r * LH + c * 8, src);
        cp_async16(qs + r * LH + c * 8, src + C);
      }
      cp_async16(vs + r * LH + c * 8, src + 2 * C);
      cp_async16(ds + r * LH + c * 8, dy + (row0 + r) * C + h * HS + c * 8);
    }
    constexpr int CT = T / 8;
    const long long pbase = ((long long)b * nh + h) * T * T;
    for (int i = threadIdx.x; i < T * CT; i += NTH) {
      const int r = i / CT, c = i - r * CT;
      cp_async16(SB + r * LT + c * 8, P + pbase + (long long)r * T + c * 8);
      cp_async16(PB + r * LT + c * 8, Pd + pbase + (long long)r * T + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const uint32_t qs_s = (uint32_t)__cvta_generic_to_shared(qs), ks_s = (uint32_t)__cvta_generic_to_shared(ks);
  const uint32_t vs_s = (uint32_t)__cvta_generic_to_shared(vs), ds_s = (uint32_t)__cvta_generic_to_shared(ds);
  const uint32_t SB_s = (uint32_t)__cvta_generic_to_shared(SB), PB_s = (uint32_t)__cvta_generic_to_shared(PB);
  const int l7 = lane & 7, l3 = (lane >> 3) & 1, l4 = lane >> 4;

  // ================= phase 1: this warp's 16 queries =================
  {
    const int q0 = warp * 16;
    uint32_t ady[HS / 16][4];                             // A fragments of dy: rows q0 .. q0 + 15, 16 dims per step
#pragma unroll
    for (int s = 0; s < HS / 16; ++s)
      ldm_x4(ds_s + (uint32_t)((q0 + l7 + l3 * 8) * LH + s * 16 + l4 * 8) * 2u, ady[s]);
    float dp[T / 8][4];
#pragma unroll
    for (int j = 0; j < T / 8; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < T / 8; j += 2) {                  // dPd[q, key] = sum_dim dy[q, dim] v[key, dim]: B rows = keys
#pragma unroll
      for (int s = 0; s < HS / 16; ++s) {
        uint32_t bb[4];
        ldm_x4(vs_s + (uint32_t)((j * 8 + l4 * 8 + l7) * LH + s * 16 + l3 * 8) * 2u, bb);
        mma_bf16(dp[j], ady[s], bb[0], bb[1]);
        mma_bf16(dp[j + 1], ady[s], bb[2], bb[3]);
      }
    }
    // r = rowsum(dPd o Pd); dS = scale (dPd o Pd - P r), written over P
    const uint32_t* SBw = reinterpret_cast<const uint32_t*>(SB);
    const uint32_t* PBw = reinterpret_cast<const uint32_t*>(PB);
    const int w0 = (q0 + g) * (LT / 2) + t, w1 = w0 + 8 * (LT / 2);
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int j = 0; j < T / 8; ++j) {
      const float2 pd0 = unpack_bf16(PBw[w0 + j * 4]), pd1 = unpack_bf16(PBw[w1 + j * 4]);
      dp[j][0] *= pd0.x; dp[j][1] *= pd0.y; dp[j][2] *= pd1.x; dp[j][3] *= pd1.y;
      r0 += dp[j][0] + dp[j][1]; r1 += dp[j][2] + dp[j][3];
    }
    r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    uint32_t* SBo = reinterpret_cast<uint32_t*>(SB);
    uint32_t dsp[T / 8][2];                               // packed dS: [tile][row half]
#pragma unroll
    for (int j = 0; j < T / 8; ++j) {
      const float2 p0 = unpack_bf16(SBw[w0 + j * 4]), p1 = unpack_bf16(SBw[w1 + j * 4]);
      dsp[j][0] = pack_bf16(scale * (dp[j][0] - p0.x * r0), scale * (dp[j][1] - p0.y * r0));
      dsp[j][1] = pack_bf16(scale * (dp[j][2] - p1.x * r1), scale * (dp[j][3] - p1.y * r1));
      SBo[w0 + j * 4] = dsp[j][0];
      SBo[w1 + j * 4] = dsp[j][1];
    }
    if constexpr (STAGED) {                               // v is done: k takes its place
      __syncthreads();
      for (int i = threadIdx.x; i < T * (HS / 8); i += NTH) {
        const int r = i / (HS / 8), c = i - r * (HS / 8);
        cp_async16(ks + r * LH + c * 8, qkv + (row0 + r) * 3 * C + h * HS + c * 8);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    // dq[q, dim] = sum_key dS[q, key] k[key, dim]: A from the registers, B = k read transposed
    float dq[HS / 8][4];
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < T / 16; ++kk) {
      const uint32_t a[4] = {dsp[2 * kk][0], dsp[2 * kk][1], dsp[2 * kk + 1][0], dsp[2 * kk + 1][1]};
#pragma unroll
      for (int jp = 0; jp < HS / 16; ++jp) {
        uint32_t bb[4];
        ldm_x4_t(ks_s + (uint32_t)((kk * 16 + l3 * 8 + l7) * LH + (jp * 2 + l4) * 8) * 2u, bb);
        mma_bf16(dq[jp * 2], a, bb[0], bb[1]);
        mma_bf16(dq[jp * 2 + 1], a, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) {
      __nv_bfloat16* o = dqkv + (row0 + q0 + g) * 3 * C + C + h * HS + j * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(o) = pack_bf16(dq[j][0], dq[j][1]);
      *reinterpret_cast<uint32_t*>(o + 8LL * 3 * C) = pack_bf16(dq[j][2], dq[j][3]);
    }
  }
  __syncthreads();
  if constexpr (STAGED) {                                 // k is done: q takes its place
    for (int i = threadIdx.x; i < T * (HS / 8); i += NTH) {
      const int r = i / (HS / 8), c = i - r * (HS / 8);
      cp_async16(qs + r * LH + c * 8, qkv + (row0 + r) * 3 * C + C + h * HS + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  // ================= phase 2: this warp's 16 keys =================
  {
    const int k0 = warp * 16;
    float dk[HS / 8][4], dv[HS / 8][4];
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) { dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; }
#pragma unroll 2
    for (int qb = 0; qb < T / 16; ++qb) {
      // A[m = key][k = query] = X[query][key] (X = dS, Pd), read transposed: matrices (queries 0-7 | 8-15) x (keys 0-7 | 8-15)
      uint32_t as[4], ap[4];
      const uint32_t aoff = (uint32_t)((qb * 16 + l4 * 8 + l7) * LT + k0 + l3 * 8) * 2u;
      ldm_x4_t(SB_s + aoff, as);
      ldm_x4_t(PB_s + aoff, ap);
#pragma unroll
      for (int jp = 0; jp < HS / 16; ++jp) {
        // B[k = query][n = dim] = q[query][dim] / dy[query][dim], read transposed
        uint32_t bq[4], bd[4];
        const uint32_t boff = (uint32_t)((qb * 16 + l3 * 8 + l7) * LH + (jp * 2 + l4) * 8) * 2u;
        ldm_x4_t(qs_s + boff, bq);
        ldm_x4_t(ds_s + boff, bd);
        mma_bf16(dk[jp * 2], as, bq[0], bq[1]);
        mma_bf16(dk[jp * 2 + 1], as, bq[2], bq[3]);
        mma_bf16(dv[jp * 2], ap, bd[0], bd[1]);
        mma_bf16(dv[jp * 2 + 1], ap, bd[2], bd[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) {
      __nv_bfloat16* o = dqkv + (row0 + k0 + g) * 3 * C + h * HS + j * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(o) = pack_bf16(dk[j][0], dk[j][1]);
      *reinterpret_cast<uint32_t*>(o + 8LL * 3 * C) = pack_bf16(dk[j][2], dk[j][3]);
      *reinterpret_cast<uint32_t*>(o + 2 * C) = pack_bf16(dv[j][0], dv[j][1]);
      *reinterpret_cast<uint32_t*>(o + 2 * C + 8LL * 3 * C) = pack_bf16(dv[j][2], dv[j][3]);
    }
  }
}

template <int HS, int NQ>
int launch_attn_bwd_small(const void* qkv, const void* dy, const void* P, const void* Pd, void* dqkv, int B, int C, int nh, cudaStream_t stream) {
  constexpr int T = NQ * 16;
  const int smem = ((HS > 32 ? 2 : 4) * T * (HS + 8) + 2 * T * (T + 8)) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_bwd_small_kernel<HS, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { mmfn_set_error("attention_bwd_small: shared memory attribute (%d B): %s", smem, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  attn_bwd_small_kernel<HS, NQ><<<dim3(nh, B), NQ * 32, smem, stream>>>(
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)P, (const __nv_bfloat16*)Pd, (__nv_bfloat16*)dqkv, C, nh,
      rsqrtf((float)HS));
  return mmfn_launch_status("attention_bwd_small");
}


// ------------------------------------------------------------------------------------------------------------------
// TF32 configuration (fp32 tensors, tf32 m16n8k8 MMAs): two fp32 T x T tiles do not fit in one CTA, so a CLUSTER OF TWO
// CTAs owns a (sample, head): CTA r holds P / Pd of queries [r T/2, (r+1) T/2), computes their dS and dQ, then the
// partial dK / dV of ALL keys over its queries; the halves are exchanged through distributed shared memory (each CTA
// finishes the keys of its own half).  q, k, v take turns in one buffer.  Operand fragments that need a transposed view
// are fetched with 32-bit shared loads (ldmatrix.trans works on 16-bit elements only); strides keep them conflict-free.
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NW warps per CTA, TQ = 16 NW queries per CTA, T = 2 TQ tokens.  Shared memory (fp32): dy [T][HS + 4], staged q / k / v
// [T][HS + 4], SB (P -> dS) and PB (Pd) [TQ][T + 8].
template <int HS, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
attn_bwd_small_tf32_kernel(const float* __restrict__ qkv, const float* __restrict__ dy, const float* __restrict__ P,
                           const float* __restrict__ Pd, float* __restrict__ dqkv, int C, int nh, float scale) {
  constexpr int TQ = NW * 16, T = 2 * TQ, LH = HS + 4, LT = T + 8, NTH = NW * 32;
  extern __shared__ __align__(16) float smf[];
  float* ds = smf;
  float* xb = ds + T * LH;
  float* SB = xb + T * LH;
  float* PB = SB + TQ * LT;
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  const int r = (int)cluster.block_rank();                // which half of the queries
  const int h = blockIdx.x >> 1, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int l7 = lane & 7, l3 = (lane >> 3) & 1, l4 = lane >> 4;
  const long long row0 = (long long)b * T;
  auto load_head = [&](float* dst, const float* src, long long ld) {     // [T][HS] slice of a (rows, ld) matrix
    for (int i = threadIdx.x; i < T * (HS / 4); i += NTH) {
      const int rr = i / (HS / 4), c = i - rr * (HS / 4);
      cp_async16(dst + rr * LH + c * 4, src + (row0 + rr) * ld + c * 4);
    }
  };
  load_head(ds, dy + h * HS, C);
  load_head(xb, qkv + 2 * C + h * HS, 3LL * C);           // v
  {
    const long long pbase = (((long long)b * nh + h) * T + r * TQ) * T;
    for (int i = threadIdx.x; i < TQ * (T / 4); i += NTH) {
      const int rr = i / (T / 4), c = i - rr * (T / 4);
      cp_async16(SB + rr * LT + c * 4, P + pbase + (long long)rr * T + c * 4);
      cp_async16(PB + rr * LT + c * 4, Pd + pbase + (long long)rr * T + c * 4);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t ds_s = (uint32_t)__cvta_generic_to_shared(ds), xb_s = (uint32_t)__cvta_generic_to_shared(xb);
  const uint32_t* xbw = reinterpret_cast<const uint32_t*>(xb);
  const uint32_t* dsw = reinterpret_cast<const uint32_t*>(ds);

  // ================= phase 1: this warp's 16 queries =================
  const int lq = warp * 16, gq = r * TQ + lq;             // local / in-sample query row
  float dsv[T / 8][4];                                    // dPd, then dS (C-fragment layout)
  {
    uint32_t ady[HS / 8][4];
#pragma unroll
    for (int s = 0; s < HS / 8; ++s)
      ldm_x4(ds_s + (uint32_t)((gq + l7 + l3 * 8) * LH + s * 8 + l4 * 4) * 4u, ady[s]);
#pragma unroll
    for (int j = 0; j < T / 8; ++j) dsv[j][0] = dsv[j][1] = dsv[j][2] = dsv[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < T / 8; j += 2) {
#pragma unroll
      for (int s = 0; s < HS / 8; ++s) {
        uint32_t bb[4];
        ldm_x4(xb_s + (uint32_t)((j * 8 + l4 * 8 + l7) * LH + s * 8 + l3 * 4) * 4u, bb);
        mma_tf32(dsv[j], ady[s], bb[0], bb[1]);
        mma_tf32(dsv[j + 1], ady[s], bb[2], bb[3]);
      }
    }
    float r0 = 0.f, r1 = 0.f;
    float* sb0 = SB + (lq + g) * LT + 2 * t;
    const float* pb0 = PB + (lq + g) * LT + 2 * t;
#pragma unroll
    for (int j = 0; j < T / 8; ++j) {
      const float2 pd0 = *reinterpret_cast<const float2*>(pb0 + j * 8), pd1 = *reinterpret_cast<const float2*>(pb0 + 8 * LT + j * 8);
      dsv[j][0] *= pd0.x; dsv[j][1] *= pd0.y; dsv[j][2] *= pd1.x; dsv[j][3] *= pd1.y;
      r0 += dsv[j][0] + dsv[j][1]; r1 += dsv[j][2] + dsv[j][3];
    }
    r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
#pragma unroll
    for (int j = 0; j < T / 8; ++j) {
      const float2 p0 = *reinterpret_cast<const float2*>(sb0 + j * 8), p1 = *reinterpret_cast<const float2*>(sb0 + 8 * LT + j * 8);
      dsv[j][0] = scale * (dsv[j][0] - p0.x * r0); dsv[j][1] = scale * (dsv[j][1] - p0.y * r0);
      dsv[j][2] = scale * (dsv[j][2] - p1.x * r1); dsv[j][3] = scale * (dsv[j][3] - p1.y * r1);
      *reinterpret_cast<float2*>(sb0 + j * 8) = make_float2(dsv[j][0], dsv[j][1]);
      *reinterpret_cast<float2*>(sb0 + 8 * LT + j * 8) = make_float2(dsv[j][2], dsv[j][3]);
    }
  }
  __syncthreads();                                        // v is done: k takes its place
  load_head(xb, qkv + h * HS, 3LL * C);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  {
    // dq = dS k: A = the score fragments (MMA k index t <-> key 8j + 2t, t + 4 <-> key 8j + 2t + 1), B = k rows 8j + 2t (+1)
    float dq[HS / 8][4];
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < T / 8; ++j) {
      const uint32_t a[4] = {__float_as_uint(dsv[j][0]), __float_as_uint(dsv[j][2]), __float_as_uint(dsv[j][1]), __float_as_uint(dsv[j][3])};
      const uint32_t* kr = xbw + (j * 8 + 2 * t) * LH + g;
#pragma unroll
      for (int jn = 0; jn < HS / 8; ++jn) mma_tf32(dq[jn], a, kr[jn * 8], kr[LH + jn * 8]);
    }
#pragma unroll
    for (int j = 0; j < HS / 8; ++j) {
      float* o = dqkv + (row0 + gq + g) * 3 * C + C + h * HS + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(o) = make_float2(dq[j][0], dq[j][1]);
      *reinterpret_cast<float2*>(o + 8LL * 3 * C) = make_float2(dq[j][2], dq[j][3]);
    }
  }
  __syncthreads();                                        // k is done: q takes its place; every warp's dS is in SB
  load_head(xb, qkv + C + h * HS, 3LL * C);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ================= phase 2: partial dK / dV of two 16-key blocks over this CTA's queries =================
  float acc[2][2][HS / 8][4];                             // [mine / theirs][dk / dv]
#pragma unroll
  for (int w2 = 0; w2 < 2; ++w2) {
    const int k0 = ((w2 == 0 ? r : 1 - r) * NW + warp) * 16;
#pragma unroll
    for (int jn = 0; jn < HS / 8; ++jn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[w2][0][jn][i] = acc[w2][1][jn][i] = 0.f;
    }
    const uint32_t* sbw = reinterpret_cast<const uint32_t*>(SB) + k0 + g;
    const uint32_t* pbw = reinterpret_cast<const uint32_t*>(PB) + k0 + g;
#pragma unroll 2
    for (int qs = 0; qs < TQ / 8; ++qs) {
      // A[m = key][k = query] = X[query][key]: a0 (key g, query t), a1 (key g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
      const int o0 = (qs * 8 + t) * LT, o1 = o0 + 4 * LT;
      const uint32_t as[4] = {sbw[o0], sbw[o0 + 8], sbw[o1], sbw[o1 + 8]};
      const uint32_t ap[4] = {pbw[o0], pbw[o0 + 8], pbw[o1], pbw[o1 + 8]};
      const int qrow = (r * TQ + qs * 8 + t) * LH + g;    // B[k = query][n = dim]: rows t, t + 4 of the step
#pragma unroll
      for (int jn = 0; jn < HS / 8; ++jn) {
        mma_tf32(acc[w2][0][jn], as, xbw[qrow + jn * 8], xbw[qrow + 4 * LH + jn * 8]);
        mma_tf32(acc[w2][1][jn], ap, dsw[qrow + jn * 8], dsw[qrow + 4 * LH + jn * 8]);
      }
    }
  }
  // ---- exchange: the other CTA finishes "theirs"; its partial for "mine" comes back the same way
  __syncthreads();                                        // SB is free: it carries the exported fragments [warp][dk/dv][tile][lane][4]
  float4* xp = reinterpret_cast<float4*>(SB);
#pragma unroll
  for (int w = 0; w < 2; ++w) {
#pragma unroll
    for (int jn = 0; jn < HS / 8; ++jn)
      xp[((warp * 2 + w) * (HS / 8) + jn) * 32 + lane] = make_float4(acc[1][w][jn][0], acc[1][w][jn][1], acc[1][w][jn][2], acc[1][w][jn][3]);
  }
  cluster.sync();
  const float4* xpeer = cluster.map_shared_rank(xp, 1 - r);
  const int k0 = (r * NW + warp) * 16;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
#pragma unroll
    for (int jn = 0; jn < HS / 8; ++jn) {
      const float4 o = xpeer[((warp * 2 + w) * (HS / 8) + jn) * 32 + lane];
      float* dst = dqkv + (row0 + k0 + g) * 3 * C + (w == 0 ? 0 : 2 * C) + h * HS + jn * 8 + 2 * t;
      *reinterpret_cast<float2*>(dst) = make_float2(acc[0][w][jn][0] + o.x, acc[0][w][jn][1] + o.y);
      *reinterpret_cast<float2*>(dst + 8LL * 3 * C) = make_float2(acc[0][w][jn][2] + o.z, acc[0][w][jn][3] + o.w);
    }
  }
  cluster.sync();                                         // the peer may still be reading this CTA's shared memory
}

template <int HS, int NW>
int launch_attn_bwd_small_tf32(const void* qkv, const void* dy, const void* P, const void* Pd, void* dqkv, int B, int C, int nh, cudaStream_t stream) {
  constexpr int TQ = NW * 16, T = 2 * TQ;
  const int smem = (2 * T * (HS + 4) + 2 * TQ * (T + 8)) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_bwd_small_tf32_kernel<HS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce != cudaSuccess) { mmfn_set_error("attention_bwd_small_tf32: shared memory attribute (%d B): %s", smem, cudaGetErrorString(ce)); return (int)ce; }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nh, B, 1);
  cfg.blockDim = dim3(NW * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t ce = cudaLaunchKernelEx(&cfg, attn_bwd_small_tf32_kernel<HS, NW>, (const float*)qkv, (const float*)dy, (const float*)P,
                                      (const float*)Pd, (float*)dqkv, C, nh, rsqrtf((float)HS));
  if (ce != cudaSuccess) { mmfn_set_error("attention_bwd_small_tf32: launch: %s", cudaGetErrorString(ce)); return (int)ce; }
  return mmfn_launch_status("attention_bwd_small_tf32");
}

}  // namespace

// Whole attention backward of one transformer block for small heads, bf16: qkv (B*T, 3C) [key | query | value], dy (B*T, C)
// gradient of the attention output, P / Pd (B, nh, T, T) saved probabilities before / after dropout (the same tensor when
// there was no dropout) -> dqkv (B*T, 3C), every element written.  Head size C / nh in {16, 32, 64}, T in {128, 192}.
MMFN_API int mmfn_attention_bwd_small_bf16(const void* qkv, const void* dy, const void* P, const void* Pd, void* dqkv,
                                           int B, int T, int C, int nh, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && dy && P && Pd && dqkv, "attention_bwd_small: null pointer");
  MMFN_CHECK_ARG(B >= 1 && B <= 65535 && nh >= 1 && C % nh == 0 && (C / nh == 16 || C / nh == 32 || C / nh == 64) && (T == 128 || T == 192),
                 "attention_bwd_small: needs head size 16, 32 or 64 and T in {128, 192}");
  MMFN_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)dy | (uintptr_t)P | (uintptr_t)Pd | (uintptr_t)dqkv) & 15) == 0, "attention_bwd_small: 16-byte alignment");
  const int hs = C / nh;
  if (hs == 16) return T == 192 ? launch_attn_bwd_small<16, 12>(qkv, dy, P, Pd, dqkv, B, C, nh, stream)
                                : launch_attn_bwd_small<16, 8>(qkv, dy, P, Pd, dqkv, B, C, nh, stream);
  if (hs == 32) return T == 192 ? launch_attn_bwd_small<32, 12>(qkv, dy, P, Pd, dqkv, B, C, nh, stream)
                                : launch_attn_bwd_small<32, 8>(qkv, dy, P, Pd, dqkv, B, C, nh, stream);
  return T == 192 ? launch_attn_bwd_small<64, 12>(qkv, dy, P, Pd, dqkv, B, C, nh, stream)
                  : launch_attn_bwd_small<64, 8>(qkv, dy, P, Pd, dqkv, B, C, nh, stream);
}

// The same for the TF32 configuration: fp32 tensors, tf32 tensor-core products; head size 16 or 32, T in {128, 192}.
// Launches 2 * nh * B CTAs in clusters of two (see the kernel comment).
MMFN_API int mmfn_attention_bwd_small_tf32(const float* qkv, const float* dy, const float* P, const float* Pd, float* dqkv,
                                           int B, int T, int C, int nh, cudaStream_t stream) {
  MMFN_CHECK_ARG(qkv && dy && P && Pd && dqkv, "attention_bwd_small_tf32: null pointer");
  MMFN_CHECK_ARG(B >= 1 && B <= 65535 && nh >= 1 && C % nh == 0 && (C / nh == 16 || C / nh == 32) && (T == 128 || T == 192),
                 "attention_bwd_small_tf32: needs head size 16 or 32 and T in {128, 192}");
  MMFN_CHECK_ARG((((uintptr_t)qkv | (uintptr_t)dy | (uintptr_t)P | (uintptr_t)Pd | (uintptr_t)dqkv) & 15) == 0, "attention_bwd_small_tf32: 16-byte alignment");
  if (C / nh == 16) return T == 192 ? launch_attn_bwd_small_tf32<16, 6>(qkv, dy, P, Pd, dqkv, B, C, nh, stream)
                                    : launch_attn_bwd_small_tf32<16, 4>(qkv, dy, P, Pd, dqkv, B, C, nh, stream);
  return T == 192 ? launch_attn_bwd_small_tf32<32, 6>(qkv, dy, P, Pd, dqkv, B, C, nh, stream)
                  : launch_attn_bwd_small_tf32<32, 4>(qkv, dy, P, Pd, dqkv, B, C, nh, stream);
}
