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
