"""Thin torch-tensor wrappers over the C ABI (include/mmfn_b200.h).

torch supplies device memory and the current stream only; every arithmetic op below is a
kernel of libmmfn_b200.so.  Tensors are fp32 CUDA tensors; activations are NHWC.
"""
import os

import torch

from ._lib import lib, MmfnError  # noqa: F401


def _p(t):
    return 0 if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name="tensor", bf16_ok=False):
    if t is not None and not (t.is_cuda and (t.dtype == torch.float32 or (bf16_ok and t.dtype == torch.bfloat16))):
        raise MmfnError(f"{name}: expected a CUDA float32{'/bfloat16' if bf16_ok else ''} tensor, got {t.device}/{t.dtype}")


BF = torch.bfloat16


def twin(t):
    """The bf16 twin a producer kernel wrote next to an fp32 activation (attribute `.h`), or None.  In the bf16
    configuration trunk activations exist twice: fp32 for the element-wise consumers (residual adds, BatchNorm backward,
    pooling), bf16 for the convolution that reads them through TMA."""
    return getattr(t, "h", None)


def _with_twin(t, h):
    if h is not None:
        t.h = h
    return t


def to_bf16(x):
    """fp32 -> bf16 copy (n % 4 == 0)."""
    assert x.is_contiguous() and x.dtype == torch.float32 and x.numel() % 4 == 0
    y = torch.empty(x.shape, device=x.device, dtype=BF)
    lib().f32_to_bf16(_p(x), _p(y), x.numel(), _st())
    return y


TF32 = True   # large aligned GEMMs/convs run on the tcgen05 TF32 path; False = exact fp32 SIMT everywhere
# Attention backward dP -> dS -> dQ in one tcgen05 kernel (attention_bwd_dq) instead of GEMM + softmax_bwd + GEMM.
# Off by default: (1) measured on B200 at B=16 it shortens the step by < 0.1 ms, because the dK GEMM still has to wait
# for its dS (9.9 us + 9 us vs 9.1 + 4.5 + 9.0 us with dQ || dK); (2) it takes the softmax row term from the
# FlashAttention identity delta_i = sum_d dY_id Y_id, which under TF32 rounding of Y differs from sum_j dP_ij P_ij by
# ~1e-3 relative -- enough to dominate the (second-order small) query/key gradients of a freshly initialised model and
# to fail the per-tensor gradient-norm check of tests/test_gpu_parity.py.  The kernel itself is covered op-level.
FUSED_ATTN_BWD = False
# bf16 tensor-core operands (BASELINE configs[2]): activations that feed MMAs and a bf16 shadow of the weights are
# stored in bf16 (tcgen05 kind::f16, fp32 accumulate); residual stream, statistics, gradients of parameters, master
# weights and optimizer stay fp32.  Implies TF32 for everything the bf16 kernels do not cover (stems, VectorNet, GAT, head).
BF16 = False
BF16_ATTN = True      # bf16 configuration: attention core on the bf16 kernel (False: TF32 kernel on an fp32 qkv buffer)
import os as _os
if _os.environ.get("MMFN_BF16_ATTN") is not None:          # A/B switches for measurement runs
    BF16_ATTN = _os.environ["MMFN_BF16_ATTN"] != "0"


def set_precision(mode):
    """"fp32": exact-fp32 SIMT kernels everywhere (tests);  "tf32": production configs[1];  "bf16": configs[2]."""
    global TF32, BF16
    if mode not in ("fp32", "tf32", "bf16"):
        raise MmfnError(f"unknown precision {mode!r}")
    TF32 = mode != "fp32"
    BF16 = mode == "bf16"


def _major(t):
    """-> (is_mn_major, pitch) when `t` (.., rows, red) is TMA-addressable as a 2-D matrix (pitches are multiples of
    16 bytes: 4 fp32 / 8 bf16 elements), else None."""
    rows, red = t.shape[-2], t.shape[-1]
    sr, sc = t.stride(-2), t.stride(-1)
    al = 8 if t.dtype == BF else 4
    if t.data_ptr() % 16:
        return None
    if sc == 1 and sr % al == 0 and sr >= red:
        return (0, sr)
    if sr == 1 and sc % al == 0 and sc >= rows:
        return (1, sc)
    return None


def gemm(A, B, C, *, bias=None, res=None, mask=None, alpha=1.0, act=0, accum=0,
         drop_p=0.0, seed=0, splitk=1):
    """C[..., M, N] (+)= alpha * A[..., M, K] @ B[..., N, K]^T  with up to two leading batch dims.

    A and B may be arbitrary strided views; C needs unit column stride.  res/mask index like C.
    """
    for t in (A, B, C, mask):
        _chk(t, bf16_ok=True)
    for t in (bias, res):
        _chk(t)
    bf_in = A.dtype == BF
    if bf_in != (B.dtype == BF):
        raise MmfnError("gemm: A and B must have the same element type")
    if mask is not None and (mask.dtype == BF) != bf_in:
        raise MmfnError("gemm: the mask tensor must have the operands' element type")
    nbd = C.dim() - 2
    assert A.dim() == C.dim() and B.dim() == C.dim() and 0 <= nbd <= 2
    M, K = A.shape[-2], A.shape[-1]
    N = B.shape[-2]
    assert B.shape[-1] == K and C.shape[-2] == M and C.shape[-1] == N, (A.shape, B.shape, C.shape)
    assert C.stride(-1) == 1 or N == 1
    bs = list(C.shape[:nbd])
    nb = [1] * (2 - nbd) + bs

    def bstr(t):
        s = [t.stride(i) if t.shape[i] > 1 else 0 for i in range(nbd)]
        for i in range(nbd):
            assert t.shape[i] == bs[i] or t.shape[i] == 1
        return [0] * (2 - nbd) + s

    a_b, b_b, c_b = bstr(A), bstr(B), bstr(C)
    for t in (res, mask):
        if t is not None:
            assert t.shape == C.shape and t.stride() == C.stride()
    nbt = nb[0] * nb[1]
    es = 2.0 if bf_in else 4.0
    lib().next_work = (2.0 * M * N * K * nbt, nbt * (es * (M * K + N * K) + C.element_size() * M * N), M, N, K, nbt)
    if bf_in:
        # bf16 operands: tcgen05 kind::f16 only -- there is no other path for 2-byte tensors
        am, bm = _major(A), _major(B)
        ok = (am is not None and bm is not None and C.stride(-1) == 1 and all(x % 8 == 0 for x in a_b + b_b)
              and all(x > 0 for i, x in enumerate(a_b + b_b) if ([nb[0], nb[1]] * 2)[i] > 1))
        if not ok:
            raise MmfnError(f"gemm(bf16): operands are not TMA-addressable (shapes {tuple(A.shape)} x {tuple(B.shape)}, "
                            f"strides {A.stride()} / {B.stride()})")
        if accum == 1:
            accum = 2
        linear = act == 0 and mask is None and drop_p == 0
        if accum == 0 and linear and splitk == 1 and K >= 4096 and C.dtype != BF and ((M + 127) // 128) * ((N + 63) // 64) * nbt < 64:
            C.zero_()
            accum = 2
        lib().gemm_bf16(_p(A), am[1], am[0], a_b[0], a_b[1], _p(B), bm[1], bm[0], b_b[0], b_b[1],
                        _p(C), int(C.dtype == BF), C.stride(-2), c_b[0], c_b[1], M, N, K, nb[0], nb[1],
                        _p(bias), _p(res), _p(mask), float(alpha), int(act), int(accum), float(drop_p),
                        int(seed), int(splitk) if splitk > 1 else 0, _st())
        return C
    if C.dtype == BF:
        # fp32 operands (TF32 multiply), bf16 result: attention-gradient products feeding a bf16 GEMM
        am, bm = _major(A), _major(B)
        if not (TF32 and am is not None and bm is not None and act == 0 and mask is None and drop_p == 0 and accum == 0
                and bias is None and res is None and all(x % 4 == 0 for x in a_b + b_b)):
            raise MmfnError("gemm: a bf16 result from fp32 operands needs the plain TF32 tensor-core path")
        lib().gemm_tf32_out(_p(A), am[1], am[0], a_b[0], a_b[1], _p(B), bm[1], bm[0], b_b[0], b_b[1],
                            _p(C), 1, C.stride(-2), c_b[0], c_b[1], M, N, K, nb[0], nb[1], float(alpha), _st())
        return C
    # (M >= 16: the per-sample matrices of the waypoint head / VectorNet tail (M = batch) also take the tensor-core kernel --
    #  TMA zero-fills the rest of the 128-row tile; the SIMT kernel needs ~18 us for these latency-bound shapes)
    if TF32 and M >= 16 and K >= 16 and M * N * K * nbt >= (1 << 18):
        am, bm = _major(A), _major(B)
        bs_ok = all(x % 4 == 0 for x in a_b + b_b)
        if nbt > 1:      # TMA cannot express broadcast (stride-0) batches
            bs_ok = bs_ok and all(x > 0 for i, x in enumerate(a_b + b_b) if ([nb[0], nb[1]] * 2)[i] > 1)
        if am is not None and bm is not None and C.stride(-1) == 1 and bs_ok and nbt * max(1, splitk) <= 4096:
            if accum == 1:          # plain += is a single-writer RMW; tensor-core path accumulates atomically
                accum = 2
            linear = act == 0 and mask is None and drop_p == 0
            if accum == 0 and linear and splitk == 1 and K >= 4096 and ((M + 127) // 128) * ((N + 63) // 64) * nbt < 64:
                # few output tiles, long reduction (VectorNet generator dgrad at B >= 64: 64x64 <- K = 262144):
                # zero C and let the library split K over the SMs with atomic accumulation
                C.zero_()
                accum = 2
            lib().gemm_tf32(_p(A), am[1], am[0], a_b[0], a_b[1], _p(B), bm[1], bm[0], b_b[0], b_b[1],
                            _p(C), C.stride(-2), c_b[0], c_b[1], M, N, K, nb[0], nb[1],
                            _p(bias), _p(res), _p(mask), float(alpha), int(act), int(accum), float(drop_p),
                            int(seed), int(splitk) if splitk > 1 else 0, _st())
            return C
    if splitk == 1 and act == 0 and mask is None and drop_p == 0 and K >= 512:
        # skinny outputs with a long reduction (generator dgrad: 16x64 <- K=262144; polyline wgrad: 64x7 <-
        # K=18432; radar GAT weight gradients: 5x162 <- K=1296): split K over CTAs so the reduction is not one
        # CTA's serial loop
        tiles = ((M + 63) // 64) * ((N + 63) // 64) * nbt
        if tiles < 64:
            splitk = max(1, min(K // 128, 296 // tiles))
            if splitk > 1:
                if accum == 0:
                    C.zero_()
                accum = 2
    lib().gemm_f32(_p(A), A.stride(-2), A.stride(-1), a_b[0], a_b[1],
                   _p(B), B.stride(-2), B.stride(-1), b_b[0], b_b[1],
                   _p(C), C.stride(-2), c_b[0], c_b[1],
                   M, N, K, nb[0], nb[1], _p(bias), _p(res), _p(mask),
                   float(alpha), int(act), int(accum), float(drop_p), int(seed), int(splitk), _st())
    return C


def colsum_(x2d, out):
    """out[n] += sum_m x2d[m, n]  (x2d fp32 or bf16)"""
    assert x2d.stride(1) == 1
    fn = lib().colsum_bf16 if x2d.dtype == BF else lib().colsum_f32
    fn(_p(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], _p(out), _st())


# ------------------------------------------------------------------ convolution (NHWC / KRSC)
def conv_out_hw(H, W, R, S, stride, pad):
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


def _conv_work(N, H, W, C, Co, R, S, Ho, Wo):
    """algorithmic (flops, bytes) of one conv pass: 2*MACs; input + filter + output read/written once"""
    return (2.0 * N * Ho * Wo * Co * R * S * C, 4.0 * (N * H * W * C + Co * R * S * C + N * Ho * Wo * Co),
            N, H, C, Co, R, Ho)


def _tc_conv_ok(C, Co, Ho, Wo):
    return TF32 and C % 32 == 0 and Co % 32 == 0 and Ho >= 8 and Wo >= 8


def bf16_conv_ok(C, Co, Ho, Wo):
    """geometries the bf16 implicit-GEMM kernels take (every BasicBlock convolution of the three trunks)"""
    return BF16 and C % 64 == 0 and Co % 64 == 0 and Ho >= 8 and Wo >= 8


def conv2d_fwd(x, w_krsc, stride, pad, res=None):
    N, H, W, C = x.shape
    Co, R, S, C2 = w_krsc.shape
    assert C2 == C and x.is_contiguous() and w_krsc.is_contiguous()
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
    lib().next_work = _conv_work(N, H, W, C, Co, R, S, Ho, Wo)
    if x.dtype == BF:
        assert w_krsc.dtype == BF and C % 64 == 0 and Ho >= 8 and Wo >= 8, "bf16 convolution: unsupported geometry"
        lib().conv2d_fwd_bf16(_p(x), _p(w_krsc), _p(y), _p(res), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
        return y
    if _tc_conv_ok(C, Co, Ho, Wo):
        lib().conv2d_fwd_tf32(_p(x), _p(w_krsc), _p(y), _p(res), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
    else:
        assert res is None
        lib().conv2d_fwd_f32(_p(x), _p(w_krsc), _p(y), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
    return y


# BatchNorm backward: ReLU mask recomputed from z (conv -> BN -> ReLU without a residual) / read from the bf16 twin of y
BN_MASK_FROM_Z = _os.environ.get("MMFN_BN_MASK_Z", "1") != "0"
# bf16 configuration: the activation between the two convolutions of a BasicBlock exists only as bf16 (no fp32 copy)
BF16_ONLY_INNER = _os.environ.get("MMFN_BF16_ONLY_INNER", "1") != "0"
FUSE_BN_STATS = _os.environ.get("MMFN_FUSE_BN", "1") != "0"   # train-mode BatchNorm statistics from the convolution epilogue
BN_FUSE_MAX_CTAS = int(_os.environ.get("MMFN_FUSE_BN_MAX_CTAS", "512"))


def conv_bn_fusable(x, w_krsc, stride, pad):
    """conv -> train-mode BatchNorm pairs whose batch statistics the tensor-core convolution accumulates itself: feature
    maps above the single-launch BatchNorm threshold on the TF32 / bf16 implicit-GEMM path."""
    N, H, W, C = x.shape
    Co, R, S, _ = w_krsc.shape
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    tc_ok = (x.dtype == BF and C % 64 == 0) or (x.dtype == torch.float32 and _tc_conv_ok(C, Co, Ho, Wo))
    # every epilogue warp adds its column sums with fp64 atomics: measured +5.7 us on a 128-CTA layer-3 convolution (the
    # separate reduction pass costs ~13 us) but +16.8 us on a 1024-CTA layer-1 convolution (profiles/r02_ncu_full_kernels.json)
    ctas = ((N * Ho * Wo + 127) // 128) * ((Co + 127) // 128)
    return (FUSE_BN_STATS and tc_ok and Co % 64 == 0 and Co <= BN_WS_MAX_C and Ho >= 8 and Wo >= 8
            and N * Ho * Wo > BN_SMALL_ROWS and ctas <= BN_FUSE_MAX_CTAS)


def conv2d_fwd_bn(x, w_krsc, stride, pad, running_mean, running_var, momentum=0.1, eps=1e-5):
    """z = conv(x, w) plus the batch statistics of z (mean, rstd; running statistics updated) from the same launch.
    -> (z, mean, rstd); follow with bn_apply."""
    N, H, W, C = x.shape
    Co, R, S, C2 = w_krsc.shape
    assert C2 == C and x.is_contiguous() and w_krsc.is_contiguous() and x.dtype == w_krsc.dtype
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    z = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
    mean = torch.empty(Co, device=x.device, dtype=torch.float32)
    rstd = torch.empty(Co, device=x.device, dtype=torch.float32)
    lib().next_work = _conv_work(N, H, W, C, Co, R, S, Ho, Wo)
    fn = lib().conv2d_fwd_bn_bf16 if x.dtype == BF else lib().conv2d_fwd_bn_tf32
    fn(_p(x), _p(w_krsc), _p(z), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _p(_bn_ws(x.device)), _p(mean), _p(rstd),
       _p(running_mean), _p(running_var), momentum, eps, _st())
    return z, mean, rstd


def bn_apply(x, gamma, beta, mean, rstd, res=None, relu=False, want16=False, only16=False):
    """only16: write just the bf16 result (returned as a bf16 tensor that is its own twin): for activations whose only
    readers are bf16 convolutions."""
    C = x.shape[-1]
    M = x.numel() // C
    y = None if only16 else torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=BF) if (want16 or only16) else None
    lib().bn_apply(_p(x), _p(y), M, C, _p(gamma), _p(beta), _p(mean), _p(rstd), _p(res), int(relu), _p(y16), _st())
    if only16:
        y16.h = y16
        return y16
    return _with_twin(y, y16)


def stem_uses_im2col(x, w_krsc):
    """The two stems (7x7/2 on 3 / 2 channels) cannot use the implicit-GEMM tensor-core path (channels % 32); with
    TF32 enabled they run as im2col + ONE dense tensor-core GEMM instead of the SIMT gather-GEMM."""
    return TF32 and x.shape[-1] < 32 and w_krsc.shape[0] % 4 == 0


def conv2d_fwd_im2col(x, w_krsc, stride, pad, w_pad=None, bf16=False):
    """-> (y, col, w_pad): col (N*Ho*Wo, Kp) is kept for the weight gradient; w_pad (Co, Kp) is the zero-padded filter
    matrix (allocate once by passing None, then pass it back in).  bf16 (bf16 configuration, 7x7 stems only): the column
    matrix and the padded filters are bf16 (Kp = a multiple of the 64-element k-block), the GEMM runs on kind::f16."""
    N, H, W, C = x.shape
    Co, R, S, _ = w_krsc.shape
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    K = R * S * C
    blk = 64 if bf16 else 32
    Kp = (K + blk - 1) // blk * blk
    if w_pad is None or w_pad.shape[1] != Kp:
        w_pad = torch.zeros((Co, Kp), device=x.device, dtype=torch.float32)
    lib().copy2d_f32(_p(w_krsc), K, _p(w_pad), Kp, Co, K, 0, _st())
    y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
    if bf16:
        assert R == 7 and S == 7 and C in (2, 3)
        col = torch.empty((N * Ho * Wo, Kp), device=x.device, dtype=BF)
        lib().im2col_stem_bf16(_p(x), _p(col), N, H, W, C, stride, pad, Ho, Wo, Kp, _st())
        gemm(col, to_bf16(w_pad), y.view(N * Ho * Wo, Co))
    else:
        col = torch.empty((N * Ho * Wo, Kp), device=x.device, dtype=torch.float32)
        lib().im2col_nhwc(_p(x), _p(col), N, H, W, C, R, S, stride, pad, Ho, Wo, Kp, _st())
        gemm(col, w_pad, y.view(N * Ho * Wo, Co))
    lib().next_work = None
    return y, col, w_pad


# Stem convolutions (7x7 / 2, 3 or 2 input channels) as a direct tensor-core kernel that gathers its operand from a
# shared-memory input patch: no column matrix in HBM.  False = im2col + dense GEMM (conv2d_fwd_im2col).
DIRECT_STEM_CONV = os.environ.get("MMFN_DIRECT_STEM", "1") != "0"


def stem_conv_direct_ok(x, w_krsc, stride, pad):
    return (TF32 and DIRECT_STEM_CONV and x.dtype == torch.float32 and x.shape[-1] in (2, 3) and stride == 2 and pad == 3
            and tuple(w_krsc.shape[:3]) == (64, 7, 7))


def conv2d_stem7_fwd(x, w_krsc):
    N, H, W, C = x.shape
    assert x.is_contiguous() and w_krsc.is_contiguous() and tuple(w_krsc.shape) == (64, 7, 7, C)
    Ho, Wo = conv_out_hw(H, W, 7, 7, 2, 3)
    z = torch.empty((N, Ho, Wo, 64), device=x.device, dtype=torch.float32)
    lib().next_work = _conv_work(N, H, W, C, 64, 7, 7, Ho, Wo)
    lib().conv2d_stem7_fwd(_p(x), _p(w_krsc), _p(z), N, H, W, C, _st())
    return z


def conv2d_stem7_wgrad_(dz, x, dw_krsc):
    """dw_krsc (64, 7, 7, C) += weight gradient; dz (N, Ho, Wo, 64) fp32 or bf16."""
    N, H, W, C = x.shape
    Ho, Wo = conv_out_hw(H, W, 7, 7, 2, 3)
    assert dz.is_contiguous() and x.is_contiguous() and dw_krsc.is_contiguous() and tuple(dz.shape) == (N, Ho, Wo, 64)
    lib().next_work = _conv_work(N, H, W, C, 64, 7, 7, Ho, Wo)
    lib().conv2d_stem7_wgrad(_p(dz), int(dz.dtype == BF), _p(x), _p(dw_krsc), N, H, W, C, _st())


def conv2d_wgrad_im2col_(dy, col, dw_krsc):
    """dw_krsc (Co, R, S, C) += dy^T col  -- one split-K tensor-core GEMM over all output pixels."""
    Co = dw_krsc.shape[0]
    K = dw_krsc.numel() // Co
    M = col.shape[0]
    gemm(dy.view(M, Co).t(), col[:, :K].t(), dw_krsc.view(Co, K), accum=2)


def filter_crsk(w_krsc, flip=False):
    Co, R, S, C = w_krsc.shape
    wt = torch.empty((C, R, S, Co), device=w_krsc.device, dtype=torch.float32)
    lib().filter_krsc_to_crsk(_p(w_krsc), _p(wt), Co, R, S, C, int(flip), _st())
    return wt


def _tc_dgrad_ok(C, Co, H, W, stride, Ho, Wo):
    if not (TF32 and C % 32 == 0 and Co % 32 == 0):
        return False
    if stride == 1:
        return H >= 8 and W >= 8
    return stride == 2 and H == 2 * Ho and W == 2 * Wo and H >= 16 and W >= 16


def conv2d_dgrad(dy, w_krsc, x_shape, stride, pad, res=None):
    """Data gradient of conv2d_fwd.  Tensor-core path: stride 1 = forward convolution of dy with mirrored taps;
    stride 2 = four parity classes of dx, each a small stride-1 implicit GEMM over dy with only the taps that can
    reach it (no zero insertion, exact MAC count).  Both read the KRSC filters as they are."""
    N, H, W, C = x_shape
    Co, R, S, _ = w_krsc.shape
    _, Ho, Wo, _ = dy.shape
    dx = torch.empty(x_shape, device=dy.device, dtype=torch.float32)
    lib().next_work = _conv_work(N, H, W, C, Co, R, S, Ho, Wo)
    if dy.dtype == BF:
        assert w_krsc.dtype == BF and Co % 64 == 0, "bf16 data gradient: unsupported geometry"
        lib().conv2d_dgrad_bf16(_p(dy), _p(w_krsc), _p(dx), _p(res), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
        return dx
    if _tc_dgrad_ok(C, Co, H, W, stride, Ho, Wo):
        lib().conv2d_dgrad_tf32(_p(dy), _p(w_krsc), _p(dx), _p(res), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
        return dx
    wt = filter_crsk(w_krsc)
    lib().conv2d_dgrad_f32(_p(dy), _p(wt), _p(dx), _p(res), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, _st())
    return dx


def conv2d_wgrad_(dy, x, dw_krsc, stride, pad):
    """dw_krsc += wgrad"""
    N, H, W, C = x.shape
    Co, R, S, _ = dw_krsc.shape
    _, Ho, Wo, _ = dy.shape
    assert dw_krsc.is_contiguous() and dy.is_contiguous() and x.is_contiguous()
    lib().next_work = _conv_work(N, H, W, C, Co, R, S, Ho, Wo)
    if dy.dtype == BF:
        assert x.dtype == BF and dw_krsc.dtype == torch.float32
        lib().conv2d_wgrad_bf16(_p(dy), _p(x), _p(dw_krsc), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, 0, _st())
        return
    if _tc_conv_ok(C, Co, Ho, Wo):
        lib().conv2d_wgrad_tf32(_p(dy), _p(x), _p(dw_krsc), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, 0, _st())
    else:
        lib().conv2d_wgrad_f32(_p(dy), _p(x), _p(dw_krsc), N, H, W, C, Co, R, S, stride, pad, Ho, Wo, 0, _st())


# ------------------------------------------------------------------ normalisation
_ws = {}
BN_WS_MAX_C = 2048
BN_SMALL_ROWS = 1024     # norm.cu: feature maps with at most this many rows take the single-launch kernel


def set_bn_small_rows(rows):
    """Move the single-launch BatchNorm threshold (library + this module's mirror of it)."""
    global BN_SMALL_ROWS
    lib().set_bn_small_rows(int(rows), 0)
    BN_SMALL_ROWS = int(rows)


if _os.environ.get("MMFN_BN_SMALL_ROWS"):
    set_bn_small_rows(int(_os.environ["MMFN_BN_SMALL_ROWS"]))


F32, TF32_T, BF16_T = 0, 1, 2                     # MMFN_F32 / MMFN_TF32 / MMFN_BF16 of the header
WS_BN, WS_STEM_IM2COL, WS_STEM_FILTER, WS_ATTN_PROB, WS_GRU_SAVED, WS_BEV = range(6)


def workspace_bytes(op, dtype=F32, a=0, b=0, c=0, d=0):
    """mmfn_workspace_bytes: the library states the scratch size, the caller (torch) allocates it."""
    out = torch.zeros(1, dtype=torch.int64)
    lib().workspace_bytes(op, dtype, a, b, c, d, out.data_ptr())
    lib().launches -= 1                           # host-only query, not a kernel
    return int(out.item())


def _bn_ws(dev):
    """fp64 scratch for the BatchNorm partial sums: one per stream, since trunks run concurrently.  Zeroed ONCE here;
    the reduction kernel's last CTA leaves it zero again (size from mmfn_workspace_bytes, C <= BN_WS_MAX_C)."""
    key = (dev, torch.cuda.current_stream().cuda_stream)
    if key not in _ws:
        _ws[key] = torch.zeros(workspace_bytes(WS_BN, F32, BN_WS_MAX_C) // 8, device=dev, dtype=torch.float64)
    return _ws[key]


def bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5, res=None, relu=False, want16=False):
    """want16: also write the bf16 twin of y (y.h) in the same pass."""
    C = x.shape[-1]
    assert C <= BN_WS_MAX_C
    M = x.numel() // C
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=BF) if want16 else None
    mean = torch.empty(C, device=x.device, dtype=torch.float32)
    rstd = torch.empty(C, device=x.device, dtype=torch.float32)
    lib().bn_train_fwd(_p(x), _p(y), M, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                       momentum, eps, _p(mean), _p(rstd), _p(res), int(relu), _p(_bn_ws(x.device)), _p(y16), _st())
    if M <= BN_SMALL_ROWS:
        lib().launches -= 1
    return _with_twin(y, y16), mean, rstd


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, eps=1e-5, res=None, relu=False, want16=False):
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=BF) if want16 else None
    mean = torch.empty(C, device=x.device, dtype=torch.float32)
    rstd = torch.empty(C, device=x.device, dtype=torch.float32)
    lib().bn_eval_fwd(_p(x), _p(y), M, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var), eps,
                      _p(mean), _p(rstd), _p(res), int(relu), _p(y16), _st())
    return _with_twin(y, y16), mean, rstd


def bn_train_bwd(dy, x, yout, mean, rstd, gamma, dgamma, dbeta, want_dres=False, out_bf16=False, relu_beta=None):
    """out_bf16: dx (the gradient of the convolution output) is written as bf16 -- it only feeds wgrad / dgrad MMAs.
    ReLU mask: from yout (the forward output, fp32 or its bf16 twin), or -- relu_beta given, yout None: the ReLU followed
    the BatchNorm directly -- recomputed from x with the BatchNorm bias (no third tensor is read)."""
    C = x.shape[-1]
    assert C <= BN_WS_MAX_C
    assert yout is None or relu_beta is None
    M = x.numel() // C
    dx = torch.empty(x.shape, device=x.device, dtype=BF if out_bf16 else torch.float32)
    dres = torch.empty_like(x) if want_dres else None
    lib().bn_train_bwd(_p(dy), _p(x), _p(yout), int(yout is not None and yout.dtype == BF), int(relu_beta is not None),
                       _p(mean), _p(rstd), _p(gamma), _p(relu_beta), M, C, _p(dx), int(out_bf16), _p(dres),
                       _p(dgamma), _p(dbeta), _p(_bn_ws(x.device)), _st())
    if M <= BN_SMALL_ROWS:
        lib().launches -= 1
    return dx, dres


def layernorm_fwd(x2d, gamma, beta, act=0, eps=1e-5, out=None, out_bf16=False):
    M, C = x2d.shape
    assert x2d.is_contiguous()
    if out is None:
        out = torch.empty(x2d.shape, device=x2d.device, dtype=BF if out_bf16 else torch.float32)
    y = out
    mean = torch.empty(M, device=x2d.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x2d.device, dtype=torch.float32)
    lib().layernorm_fwd(_p(x2d), _p(gamma), _p(beta), _p(y), int(y.dtype == BF), _p(mean), _p(rstd), M, C, eps, act, _st())
    return y, mean, rstd


def layernorm_bwd(dy, x2d, gamma, beta, mean, rstd, dgamma, dbeta, act=0, dres=None, parts=3, drop=None, drop_bf16=False):
    """parts bit 0: dx (returned), bit 1: dgamma/dbeta accumulation.  drop=(p, seed): also return
    dx * dropout_mask(p, seed) -- the gradient entering the dropout of the next residual branch (drop_bf16: as a bf16
    tensor, written even for p == 0 because it is the operand of that branch's bf16 GEMMs)."""
    M, C = x2d.shape
    assert dy.is_contiguous() and x2d.is_contiguous()
    dx = torch.empty_like(x2d) if parts & 1 else None
    dxd, p, seed = None, 0.0, 0
    if drop is not None and parts & 1 and (drop[0] > 0 or drop_bf16):
        dxd = torch.empty(x2d.shape, device=x2d.device, dtype=BF if drop_bf16 else torch.float32)
        p, seed = drop[0], drop[1]
    lib().layernorm_bwd(_p(dy), _p(x2d), _p(gamma), _p(beta), _p(mean), _p(rstd), _p(dres), _p(dx),
                        _p(dgamma), _p(dbeta), M, C, act, parts, _p(dxd), int(drop_bf16 and dxd is not None), float(p), int(seed), _st())
    if drop is not None:
        return dx, (dxd if dxd is not None else dx)
    return dx


# ------------------------------------------------------------------ layout / pooling
def nchw_to_nhwc(x, mean=None, std=None):
    B, C, H, W = x.shape
    assert x.is_contiguous()
    y = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
    if x.dtype == torch.uint8:
        lib().nchw_u8_to_nhwc_f32(x.data_ptr(), _p(y), B, C, H, W, _p(mean), _p(std), _st())
    else:
        _chk(x, "nchw input")
        lib().nchw_to_nhwc_f32(_p(x), _p(y), B, C, H, W, _p(mean), _p(std), _st())
    return y


def transpose(x3d):
    """(nb, R, C) -> (nb, C, R)"""
    nb, R, Cc = x3d.shape
    assert x3d.is_contiguous()
    y = torch.empty((nb, Cc, R), device=x3d.device, dtype=torch.float32)
    lib().transpose_f32(_p(x3d), _p(y), nb, R, Cc, _st())
    return y


def maxpool_fwd(x, want16=False):
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
    y16 = torch.empty((B, Ho, Wo, C), device=x.device, dtype=BF) if want16 else None
    idx = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.uint8)
    lib().maxpool3x3s2_fwd(_p(x), _p(y), idx.data_ptr(), B, H, W, C, _p(y16), _st())
    return _with_twin(y, y16), idx


def maxpool_bwd(dy, idx, x_shape):
    B, H, W, C = x_shape
    dx = torch.empty(x_shape, device=dy.device, dtype=torch.float32)
    lib().maxpool3x3s2_bwd(_p(dy), idx.data_ptr(), _p(dx), B, H, W, C, _st())
    return dx


# Stem tail (bn1 -> relu -> maxpool) as fused kernels: the pre-pool activation y is never written, the backward reduces
# over the pooled pixels and makes one pass over z.  False = bn_train_fwd / maxpool_fwd / maxpool_bwd / bn_train_bwd.
FUSE_STEM_TAIL = os.environ.get("MMFN_FUSE_STEM", "1") != "0"


def stem_bn_relu_maxpool_fwd(z, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5, want16=False):
    """-> (out (+ .h bf16 twin), (idx, zmax), mean, rstd); idx carries the ReLU mask in bit 7, zmax is z at the arg-max
    (see the C header): the pair is what stem_bn_relu_maxpool_bwd wants back."""
    B, H, W, C = z.shape
    assert C <= BN_WS_MAX_C and z.is_contiguous()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((B, Ho, Wo, C), device=z.device, dtype=torch.float32)
    out16 = torch.empty((B, Ho, Wo, C), device=z.device, dtype=BF) if want16 else None
    idx = torch.empty((B, Ho, Wo, C), device=z.device, dtype=torch.uint8)
    zmax = torch.empty((B, Ho, Wo, C), device=z.device, dtype=torch.float32)
    mean = torch.empty(C, device=z.device, dtype=torch.float32)
    rstd = torch.empty(C, device=z.device, dtype=torch.float32)
    lib().stem_bn_relu_maxpool_fwd(_p(z), B, H, W, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var), momentum, eps,
                                   _p(mean), _p(rstd), _p(out), _p(out16), idx.data_ptr(), _p(zmax), _p(_bn_ws(z.device)), _st())
    return _with_twin(out, out16), (idx, zmax), mean, rstd


def stem_bn_relu_maxpool_bwd(dout, idx, z, mean, rstd, gamma, dgamma, dbeta, out_bf16=False):
    """-> dz (B, H, W, C), fp32 or bf16; dgamma / dbeta accumulated."""
    B, H, W, C = z.shape
    assert dout.is_contiguous() and dout.dtype == torch.float32
    dz = torch.empty(z.shape, device=z.device, dtype=BF if out_bf16 else torch.float32)
    idx, zmax = idx
    lib().stem_bn_relu_maxpool_bwd(_p(dout), idx.data_ptr(), _p(z), _p(zmax), _p(mean), _p(rstd), _p(gamma), B, H, W, C, _p(dz),
                                   int(out_bf16), _p(dgamma), _p(dbeta), _p(_bn_ws(z.device)), _st())
    return dz


def _four(ts):
    ts = list(ts) + [None] * (4 - len(ts))
    return [_p(t) for t in ts]


def tokens_fwd(feats, pos_emb, vel_w, vel_b, velocity, drop_p=0.0, seed=0):
    B, H, W, C = feats[0].shape
    nmod = len(feats)
    tok = torch.empty((B, nmod * 64, C), device=feats[0].device, dtype=torch.float32)
    f = _four(feats)
    lib().tokens_fwd(f[0], f[1], f[2], f[3], nmod, B, H, W, C, _p(pos_emb), _p(vel_w), _p(vel_b),
                     _p(velocity), _p(tok), drop_p, seed, _st())
    return tok


def tokens_bwd_(dtok, dfeats, shape, velocity, dpos, dvel_w, dvel_b, drop_p=0.0, seed=0):
    """dfeats[m] += pool_bwd(dtok) (entries may be None); parameter grads accumulated."""
    B, H, W, C = shape
    f = _four(dfeats)
    lib().tokens_bwd(_p(dtok), f[0], f[1], f[2], f[3], len(dfeats), B, H, W, C, _p(velocity),
                     _p(dpos), _p(dvel_w), _p(dvel_b), drop_p, seed, _st())


def upsample_add_fwd(feat, tok, m, align_corners=True, want16=None):
    """feat + bilinear upsample of tokens [m*64, (m+1)*64) viewed as an 8x8 map (F.interpolate semantics).
    want16 (default: the bf16 configuration): also write the bf16 twin the next layer's first convolution reads."""
    B, H, W, C = feat.shape
    out = torch.empty_like(feat)
    out16 = torch.empty(feat.shape, device=feat.device, dtype=BF) if (BF16 if want16 is None else want16) else None
    lib().upsample_add_fwd(_p(feat), _p(tok), _p(out), m, tok.shape[1], B, H, W, C, int(align_corners), _p(out16), _st())
    return _with_twin(out, out16)


def upsample_add_bwd_(dA, dtok, m, align_corners=True):
    B, H, W, C = dA.shape
    lib().upsample_add_bwd(_p(dA), _p(dtok), m, dtok.shape[1], B, H, W, C, int(align_corners), _st())


def pool_sum_fwd(feats, tok):
    B, _, _, C = feats[0].shape
    fused = torch.empty((B, C), device=tok.device, dtype=torch.float32)
    f = _four(feats)
    lib().pool_sum_fwd(f[0], f[1], f[2], f[3], len(feats), _p(tok), B, C, _p(fused), _st())
    return fused


def pool_sum_bwd(dfused, nmod):
    B, C = dfused.shape
    dfe = [torch.empty((B, 8, 8, C), device=dfused.device, dtype=torch.float32) for _ in range(nmod)]
    dtok = torch.empty((B, nmod * 64, C), device=dfused.device, dtype=torch.float32)
    f = _four(dfe)
    lib().pool_sum_bwd(_p(dfused), f[0], f[1], f[2], f[3], nmod, _p(dtok), B, C, _st())
    return dfe, dtok


# ------------------------------------------------------------------ softmax family
def attention_fwd(qkv, B, T, C, nh, drop_p=0.0, seed=0, y_bf16=False):
    """Fused tcgen05 attention forward on the (B*T, 3C) [key|query|value] buffer.
    -> y (B*T, C) (bf16 when y_bf16), P (B,nh,T,T) softmax probabilities, Pd (= P after dropout; P itself when drop_p == 0)."""
    assert qkv.is_contiguous() and qkv.shape == (B * T, 3 * C)
    y = torch.empty((B * T, C), device=qkv.device, dtype=BF if y_bf16 else torch.float32)
    P = torch.empty((B, nh, T, T), device=qkv.device, dtype=torch.float32)
    Pd = torch.empty_like(P) if drop_p > 0 else None
    hs = C // nh
    lib().next_work = (4.0 * B * nh * T * T * hs, 4.0 * (B * T * 4 * C + B * nh * T * T), B, T, C, nh)
    lib().attention_fwd_tf32(_p(qkv), _p(y), int(y_bf16), _p(P), _p(Pd), B, T, C, nh, float(drop_p), int(seed), _st())
    return y, P, (Pd if Pd is not None else P)


def attention_fwd_bf16(qkv, B, T, C, nh, drop_p=0.0, seed=0, save_probs=True):
    """Fused bf16 attention forward (mmfn_attention_fwd_bf16) on the bf16 (B*T, 3C) [key|query|value] buffer.
    -> y (B*T, C) bf16, P / Pd (B,nh,T,T) bf16 (None when save_probs is False), stats (B,nh,T,2) fp32 row max / sum."""
    assert qkv.is_contiguous() and qkv.shape == (B * T, 3 * C) and qkv.dtype == BF
    y = torch.empty((B * T, C), device=qkv.device, dtype=BF)
    P = torch.empty((B, nh, T, T), device=qkv.device, dtype=BF) if save_probs else None
    Pd = torch.empty_like(P) if (save_probs and drop_p > 0) else None
    stats = torch.empty((B, nh, T, 2), device=qkv.device, dtype=torch.float32)
    hs = C // nh
    lib().next_work = (4.0 * B * nh * T * T * hs, 2.0 * (B * T * 4 * C) + (2.0 * B * nh * T * T * (2 if Pd is not None else 1) if save_probs else 0.0),
                       B, T, C, nh)
    lib().attention_fwd_bf16(_p(qkv), _p(y), _p(P), _p(Pd), _p(stats), B, T, C, nh, float(drop_p), int(seed), _st())
    return y, P, (Pd if Pd is not None else P), stats


def attention_bf16_ok(T, C, nh):
    return BF16 and C % nh == 0 and (C // nh) in (16, 32, 64, 128) and T in (128, 192, 256)


def attention_bwd_dq(qkv, dy, y, P, dqkv, B, T, C, nh, drop_p=0.0, seed=0):
    """Fused critical half of the attention backward: dPd = dY V^T (TMEM only), dS = softmax'(P, dPd o mask),
    dQ = dS K written into the query slice of dqkv.  Returns dS (B,nh,T,T) for the dK GEMM."""
    assert qkv.is_contiguous() and dy.is_contiguous() and y.is_contiguous() and P.is_contiguous() and dqkv.is_contiguous()
    assert dy.shape == (B * T, C) and y.shape == (B * T, C) and dqkv.shape == qkv.shape
    dS = torch.empty_like(P)
    hs = C // nh
    lib().next_work = (4.0 * B * nh * T * T * hs, 4.0 * (B * T * 5 * C + 2 * B * nh * T * T), B, T, C, nh)
    lib().attention_bwd_dq_tf32(_p(qkv), _p(dy), _p(y), _p(P), _p(dS), _p(dqkv), B, T, C, nh, float(drop_p), int(seed), _st())
    return dS


def attention_fwd_ok(T, C, nh):
    return TF32 and C % nh == 0 and (C // nh) in (16, 32, 64, 128) and T % 32 == 0 and 32 <= T <= 256


def softmax_fwd(s, scale, drop_p=0.0, seed=0):
    cols = s.shape[-1]
    rows = s.numel() // cols
    p = torch.empty_like(s)
    pd = torch.empty_like(s) if drop_p > 0 else None
    lib().softmax_fwd(_p(s), _p(p), _p(pd), rows, cols, scale, drop_p, seed, _st())
    return p, (pd if pd is not None else p)


def softmax_bwd(p, dpd, scale, drop_p=0.0, seed=0):
    """p bf16 (saved by attention_fwd_bf16): dS is written as bf16 too (operand of the dQ / dK GEMMs)"""
    cols = p.shape[-1]
    rows = p.numel() // cols
    ds = torch.empty_like(p)
    if p.dtype == BF:
        lib().softmax_bwd_bf16(_p(p), _p(dpd), _p(ds), rows, cols, scale, drop_p, seed, _st())
    else:
        lib().softmax_bwd(_p(p), _p(dpd), _p(ds), rows, cols, scale, drop_p, seed, _st())
    return ds


def gat_softmax_fwd(z, adj, alpha, drop_p=0.0, seed=0):
    cols = z.shape[-1]
    rows = z.numel() // cols
    att = torch.empty_like(z)
    attd = torch.empty_like(z)
    lib().gat_softmax_fwd(_p(z), _p(adj), _p(att), _p(attd), rows, cols, alpha, drop_p, seed, _st())
    return att, attd


def gat_softmax_bwd(z, adj, att, dattd, alpha, drop_p=0.0, seed=0):
    cols = z.shape[-1]
    rows = z.numel() // cols
    dz = torch.empty_like(z)
    lib().gat_softmax_bwd(_p(z), _p(adj), _p(att), _p(dattd), _p(dz), rows, cols, alpha, drop_p, seed, _st())
    return dz


def l2l_row0_fwd(qkv, lane_num_i32, heads, out):
    """qkv (B,L,3*heads*64); writes out (B, heads*64) (may be a column slice); returns prob."""
    B, L, _ = qkv.shape
    prob = torch.empty((B, heads, L), device=qkv.device, dtype=torch.float32)
    tmp = torch.empty((B, heads * 64), device=qkv.device, dtype=torch.float32)
    lib().l2l_row0_fwd(_p(qkv), lane_num_i32.data_ptr(), B, L, heads, 64, _p(prob), _p(tmp), _st())
    return prob, tmp


def l2l_row0_bwd(qkv, lane_num_i32, prob, dout, heads):
    B, L, _ = qkv.shape
    assert dout.is_contiguous()
    dqkv = torch.empty_like(qkv)
    lib().l2l_row0_bwd(_p(qkv), lane_num_i32.data_ptr(), _p(prob), _p(dout), B, L, heads, 64, _p(dqkv), _st())
    return dqkv


# ------------------------------------------------------------------ misc
def elu_fwd(x):
    y = torch.empty_like(x)
    lib().elu_fwd(_p(x), _p(y), x.numel(), _st())
    return y


def elu_bwd(dy, y):
    dx = torch.empty_like(y)
    lib().elu_bwd(_p(dy), _p(y), _p(dx), y.numel(), _st())
    return dx


def relu_bwd(dy, y):
    assert dy.is_contiguous() and y.is_contiguous()
    dx = torch.empty_like(y)
    lib().relu_bwd(_p(dy), _p(y), _p(dx), y.numel(), _st())
    return dx


def dropout(x, p, seed):
    if p <= 0:
        return x
    assert x.is_contiguous()
    y = torch.empty_like(x)
    lib().dropout_f32(_p(x), _p(y), x.numel(), p, seed, _st())
    return y


def axpy_(x, y, a=1.0):
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    lib().axpy_f32(_p(x), _p(y), a, x.numel(), _st())


def lane_to_vector(lane):
    """(B, L, P, 5) -> (B*L*(P-1), 7)"""
    B, L, P, F = lane.shape
    assert F == 5 and lane.is_contiguous()
    vec = torch.empty((B * L * (P - 1), 7), device=lane.device, dtype=torch.float32)
    lib().lane_to_vector(_p(lane), _p(vec), B * L, P, _st())
    return vec


def subgraph_pool_fwd(x, G, V):
    C = x.shape[-1]
    y = torch.empty((G * V, 2 * C), device=x.device, dtype=torch.float32)
    arg = torch.empty((G, C), device=x.device, dtype=torch.int32)
    lib().subgraph_pool_fwd(_p(x), G, V, C, _p(y), arg.data_ptr(), _st())
    return y, arg


def subgraph_pool_bwd(dy, arg, G, V):
    C = arg.shape[-1]
    dx = torch.empty((G * V, C), device=dy.device, dtype=torch.float32)
    lib().subgraph_pool_bwd(_p(dy), arg.data_ptr(), G, V, C, _p(dx), _st())
    return dx


def segmax_fwd(x, G, V):
    C = x.shape[-1]
    out = torch.empty((G, C), device=x.device, dtype=torch.float32)
    arg = torch.empty((G, C), device=x.device, dtype=torch.int32)
    lib().segmax_fwd(_p(x), G, V, C, _p(out), arg.data_ptr(), _st())
    return out, arg


def segmax_bwd(dout, arg, G, V):
    C = arg.shape[-1]
    dx = torch.empty((G * V, C), device=dout.device, dtype=torch.float32)
    lib().segmax_bwd(_p(dout), arg.data_ptr(), G, V, C, _p(dx), _st())
    return dx


def wgrad_n64_k7_(dy, x, dw):
    """dw (64, 7) += dy^T x (polyline input layer): one streaming pass instead of a split-K SIMT GEMM."""
    M = dy.shape[0]
    assert dy.is_contiguous() and x.is_contiguous() and dw.is_contiguous() and tuple(dw.shape) == (64, 7)
    assert tuple(dy.shape) == (M, 64) and tuple(x.shape) == (M, 7) and dy.dtype == x.dtype == dw.dtype == torch.float32
    lib().wgrad_n64_k7(_p(dy), _p(x), _p(dw), M, _st())


# whole-GPT kernels (csrc/gpt_small.cu): "1" = where they beat the per-op chain on B200 (n_embd 64: 1225 -> 908 us fwd+bwd
# at B=32 bf16, 1037 -> 965 us at B=16 TF32), "2" = also n_embd 128 in bf16 (a wash: 1379 -> 1357 us at B=32, slower at
# B=16; profiles/r02_gpt_bench_final.json), "0" = off
FUSE_GPT = int(_os.environ.get("MMFN_FUSE_GPT", "1"))


def gpt_small_ok(C, T, nh, n_layer, B=0):
    """whole-GPT forward kernel + row-local backward kernel: the narrow fusion transformers, tensor-core precisions only.
    n_embd 128 (bf16 only) is a wash inside the step at any batch (2 052 vs 2 059-2 067 samples/s at B=32) and stays on
    the per-op chain unless MMFN_FUSE_GPT=2."""
    if not FUSE_GPT or not (BF16 or TF32) or nh != 4 or T not in (128, 192) or not 1 <= n_layer <= 12:
        return False
    return C == 64 or (C == 128 and BF16 and int(FUSE_GPT) >= 2)


def gpt_small_fwd(x0, B, T, C, nh, layers, attn_p, resid_p, seed, eps=1e-5):
    """All transformer blocks of one fusion GPT in one launch.  x0 (B*T, C) fp32; layers: per block a 12-tuple of tensors
    (Wqkv, Wproj, Wfc1, Wfc2 in the operand type, their fp32 biases, ln1 gamma / beta, ln2 gamma / beta).
    -> dict of tensors stacked over the blocks: xout, x1 (fp32); h1, qkv, y, h2, a, P, Pd (operand type); mean1, rstd1,
    mean2, rstd2."""
    import ctypes
    Lr, M = len(layers), B * T
    bf = layers[0][0].dtype == BF
    dt = BF if bf else torch.float32
    dev = x0.device
    assert x0.is_contiguous() and x0.shape == (M, C) and x0.dtype == torch.float32
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    e = lambda *shape: torch.empty(shape, device=dev, dtype=dt)
    o = dict(xout=f(Lr, M, C), x1=f(Lr, M, C), h1=e(Lr, M, C), qkv=e(Lr, M, 3 * C), y=e(Lr, M, C), h2=e(Lr, M, C), a=e(Lr, M, 4 * C),
             P=e(Lr, B, nh, T, T), mean1=f(Lr, M), rstd1=f(Lr, M), mean2=f(Lr, M), rstd2=f(Lr, M))
    o["Pd"] = e(Lr, B, nh, T, T) if attn_p > 0 else None
    tab = (ctypes.c_void_p * (12 * Lr))()
    for l, tensors in enumerate(layers):
        assert len(tensors) == 12 and all(t.is_contiguous() for t in tensors) and all(t.dtype == dt for t in tensors[:4])
        for i, t in enumerate(tensors):
            tab[l * 12 + i] = t.data_ptr()
    hs = C // nh
    lib().next_work = (Lr * (2.0 * M * 12 * C * C + 4.0 * B * nh * T * T * hs), 0.0, B, T, C, nh)
    lib().gpt_small_fwd(_p(x0), B, T, C, nh, Lr, 2 if bf else 1, ctypes.addressof(tab), _p(o["xout"]), _p(o["x1"]), _p(o["h1"]),
                        _p(o["qkv"]), _p(o["y"]), _p(o["h2"]), _p(o["a"]), _p(o["P"]), _p(o["Pd"]), _p(o["mean1"]), _p(o["rstd1"]),
                        _p(o["mean2"]), _p(o["rstd2"]), float(attn_p), float(resid_p), int(seed), eps, _st())
    if o["Pd"] is None:
        o["Pd"] = o["P"]
    return o


FUSE_GPT_BWD = _os.environ.get("MMFN_FUSE_GPT_BWD", "1") != "0"


def gpt_small_transpose(tab_dev, n_layer, C, bf, out):
    """transposed operand-typed weight copies for gpt_small_bwd_rows; tab_dev: int64 device tensor (n_layer, 12) of pointers"""
    lib().gpt_small_transpose(tab_dev.data_ptr(), n_layer, C, 2 if bf else 1, _p(out), _st())


def gpt_small_bwd_rows(M, C, bf, a=None, b=None):
    """Row-local backward between two attention backwards (csrc/gpt_small.cu).  a = dict(dqkv, dx1, x, mean, rstd, gamma,
    wT) finishes a block; b = dict(dx2 (None when a is given), a, x1, mean, rstd, gamma, wT, p, seed_mlp, seed_proj)
    starts the block below.  -> dict of new tensors: dh1, dx (only without b) | dz, da, dh2, dx1, dzp, dy."""
    dev = (a or b)["wT"].device
    dt = BF if bf else torch.float32
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    e = lambda *shape: torch.empty(shape, device=dev, dtype=dt)
    o = {}
    if a is not None:
        o["dh1"] = f(M, C)
        if b is None:
            o["dx"] = f(M, C)
    if b is not None:
        o.update(dz=e(M, C), da=e(M, 4 * C), dh2=f(M, C), dx1=f(M, C), dzp=e(M, C), dy=e(M, C))
    A = a or {}
    Bk = b or {}
    g = lambda d, k: _p(d.get(k))
    lib().next_work = (2.0 * M * C * C * (3 * (a is not None) + 9 * (b is not None)), 0.0, M, C)
    lib().gpt_small_bwd_rows(M, C, 2 if bf else 1, int(a is not None), int(b is not None),
                             g(A, "dqkv"), g(A, "dx1"), g(A, "x"), g(A, "mean"), g(A, "rstd"), g(A, "gamma"), g(A, "wT"),
                             _p(o.get("dh1")), _p(o.get("dx")),
                             g(Bk, "dx2"), g(Bk, "a"), g(Bk, "x1"), g(Bk, "mean"), g(Bk, "rstd"), g(Bk, "gamma"), g(Bk, "wT"),
                             _p(o.get("dz")), _p(o.get("da")), _p(o.get("dh2")), _p(o.get("dx1")), _p(o.get("dzp")), _p(o.get("dy")),
                             float(Bk.get("p", 0.0)), int(Bk.get("seed_mlp", 0)), int(Bk.get("seed_proj", 0)), _st())
    return o


FUSE_ATTN_BWD_SMALL = _os.environ.get("MMFN_FUSE_ATTN_BWD_SMALL", "1") != "0"


def attention_bwd_small_ok(T, C, nh, p):
    """one-launch attention backward (csrc/attn_bwd_small.cu): bf16 heads of 16 / 32 / 64 dims, TF32 heads of 16 / 32, T in {128, 192}"""
    if not FUSE_ATTN_BWD_SMALL or C % nh or T not in (128, 192):
        return False
    if p.dtype == BF:
        return (C // nh) in (16, 32, 64)
    # TF32: a cluster of two CTAs per (sample, head); 64-dim heads do not fit
    return p.dtype == torch.float32 and TF32 and (C // nh) in (16, 32)


def attention_bwd_small(qkv, dy, P, Pd, B, T, C, nh):
    """qkv (B*T, 3C), dy (B*T, C), P / Pd (B, nh, T, T), all bf16 or all fp32 (TF32 products) -> dqkv (B*T, 3C)"""
    assert qkv.is_contiguous() and dy.is_contiguous() and P.is_contiguous() and Pd.is_contiguous()
    assert dy.dtype == qkv.dtype and P.dtype == qkv.dtype and Pd.dtype == qkv.dtype and P.shape == (B, nh, T, T)
    dqkv = torch.empty_like(qkv)
    es = qkv.element_size()
    lib().next_work = (8.0 * B * nh * T * T * (C // nh), es * (B * T * 7 * C + 2 * B * nh * T * T), B, T, C, nh)
    if qkv.dtype == BF:
        lib().attention_bwd_small_bf16(_p(qkv), _p(dy), _p(P), _p(Pd), _p(dqkv), B, T, C, nh, _st())
    else:
        assert qkv.dtype == torch.float32
        lib().attention_bwd_small_tf32(_p(qkv), _p(dy), _p(P), _p(Pd), _p(dqkv), B, T, C, nh, _st())
    return dqkv


FUSE_SUBGRAPH = _os.environ.get("MMFN_FUSE_SUBGRAPH", "1") != "0"
SUBGRAPH_FUSED_V = (9, 19)       # vectors per polyline the one-launch Subgraph forward is instantiated for


SUBGRAPH_MMA = os.environ.get("MMFN_SUBGRAPH_MMA", "1") != "0"   # TF32 configuration: tensor-core sub-graph kernel


def subgraph_fused_fwd(lane, layers, eps=1e-5):
    """Polyline Subgraph forward in ONE launch (csrc/vectornet.cu; reference model_rad.py:260-283, :369-382).
    lane (B, L, P, 5); layers = 3 x (W, b, gamma, beta).  -> dict with everything the unfused backward consumes:
    vec, y[3] (pre-LayerNorm), mean[3], rstd[3], x1, x2 ([h | max] inputs of layers 1 / 2), arg[3], tok (G,128), argf."""
    B, L, P, F = lane.shape
    G, V = B * L, P - 1
    assert F == 5 and lane.is_contiguous() and V in SUBGRAPH_FUSED_V and len(layers) == 3
    dev = lane.device
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    i = lambda *shape: torch.empty(shape, device=dev, dtype=torch.int32)
    o = dict(vec=f(G * V, 7), y=[f(G * V, 64) for _ in range(3)], mean=[f(G * V) for _ in range(3)], rstd=[f(G * V) for _ in range(3)],
             x1=f(G * V, 128), x2=f(G * V, 128), arg=[i(G, 64) for _ in range(3)], tok=f(G, 128), argf=i(G, 128))
    flat = [_p(t) for layer in layers for t in layer]
    lib().subgraph_fused_fwd(_p(lane), G, V, int(TF32 and SUBGRAPH_MMA), *flat, _p(o["vec"]), *[_p(t) for t in o["y"]],
                             _p(o["mean"][0]), _p(o["rstd"][0]), _p(o["mean"][1]), _p(o["rstd"][1]), _p(o["mean"][2]), _p(o["rstd"][2]),
                             _p(o["x1"]), _p(o["x2"]), *[t.data_ptr() for t in o["arg"]], _p(o["tok"]), o["argf"].data_ptr(),
                             eps, _st())
    return o


def radar_logsoftmax_fwd(v, B, C):
    y = torch.empty((B, 8, 8, C), device=v.device, dtype=torch.float32)
    lib().radar_logsoftmax_fwd(_p(v), _p(y), B, C, _st())
    return y


def radar_logsoftmax_bwd(dy, y):
    B, _, _, C = y.shape
    dv = torch.empty((B, 64, C), device=y.device, dtype=torch.float32)
    lib().radar_logsoftmax_bwd(_p(dy), _p(y), _p(dv), B, C, _st())
    return dv


_bev_ws = {}


def bev_scatter(points, strips=0):
    """points (frames, n, 3|4) f32 -> (frames, 2, 256, 256) f32, reference layout [c, xbin, ybin].
    strips == 0 (default): up to 48 frames the one-visit kernels (count: every point read once, packed-u16 `red.global`
    counters in an L2-resident scratch; convert: counters -> fp32 grid, scratch left zero; one scratch per (stream, frame
    count)) -- 11.9 us vs 24-28 us at 16 frames; above that the shared-memory strip kernel, which wins from 64 frames
    (33-36 vs 41 us).  strips in (2, 4, 8, 16): the strip kernel with that many strips per frame; strips == -1: one-visit."""
    _chk(points, "points")
    assert points.dim() == 3 and points.is_contiguous()
    F, n, s = points.shape
    out = torch.empty((F, 2, 256, 256), device=points.device, dtype=torch.float32)
    if ((strips == 0 and F <= 48) or strips == -1) and n < 65536 and F > 0:
        key = (points.device, torch.cuda.current_stream().cuda_stream, F)
        ws = _bev_ws.get(key)
        if ws is None:
            ws = _bev_ws[key] = torch.zeros(workspace_bytes(WS_BEV, F32, F) // 4, device=points.device, dtype=torch.int32)
        lib().bev_scatter_ws(_p(points), F, n, s, _p(out), _p(ws), _st())
        return out
    lib().bev_scatter(_p(points), F, n, s, _p(out), max(strips, 0), _st())
    return out


def gru_head_fwd(z0, target, w_ih, w_hh, b_ih, b_hh, w_out, b_out, steps):
    B = z0.shape[0]
    dev = z0.device
    pred = torch.empty((B, steps, 2), device=dev, dtype=torch.float32)
    saved = torch.empty((B, steps, 5, 64), device=dev, dtype=torch.float32)
    xin = torch.empty((B, steps, 2), device=dev, dtype=torch.float32)
    hlast = torch.empty((B, 64), device=dev, dtype=torch.float32)
    lib().gru_head_fwd(_p(z0), _p(target), _p(w_ih), _p(w_hh), _p(b_ih), _p(b_hh), _p(w_out), _p(b_out),
                       B, steps, _p(pred), _p(saved), _p(xin), _p(hlast), _st())
    return pred, (saved, xin, hlast)


def gru_head_bwd(dpred, ctx, w_ih, w_hh, w_out, dw_ih, dw_hh, db_ih, db_hh, dw_out, db_out):
    saved, xin, hlast = ctx
    B, steps = saved.shape[0], saved.shape[1]
    dz0 = torch.empty((B, 64), device=dpred.device, dtype=torch.float32)
    lib().gru_head_bwd(_p(dpred), _p(saved), _p(xin), _p(hlast), _p(w_ih), _p(w_hh), _p(w_out), B, steps,
                       _p(dz0), _p(dw_ih), _p(dw_hh), _p(db_ih), _p(db_hh), _p(dw_out), _p(db_out), _st())
    return dz0


def l1_loss(pred, gt, gscale=1.0, want_grad=True):
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred) if want_grad else None
    assert pred.is_contiguous() and gt.is_contiguous()
    lib().l1_loss(_p(pred), _p(gt), pred.numel(), _p(loss), _p(dpred), gscale, _st())
    return loss, dpred


def adamw_step_(p, g, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, grad_scale=1.0, p16=None):
    """p16: bf16 shadow of p, refreshed in the same pass (bf16 configuration)."""
    lib().adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay,
                     _p(state), grad_scale, _p(p16), _st())


def adamw_advance_(state, beta1=0.9, beta2=0.999):
    lib().adamw_advance(_p(state), beta1, beta2, _st())


def adamw_apply_(p, g, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, grad_scale=1.0, p16=None):
    """AdamW on one parameter range; the step count in `state` must already be advanced (adamw_advance_)."""
    lib().adamw_apply(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay,
                      _p(state), grad_scale, _p(p16), _st())


# ------------------------------------------------------------------ loader-side kernels (csrc/loader.cu)
def lidar_ego_transform(points, pose):
    """points (F, N, >= 3) fp32 raw sweeps, pose (F, 6) float64 [cos r1, sin r1, cos r2, sin r2, t1x - t2x, t1y - t2y]
    -> (F, N, 3) fp32: y flip + frame change in float64 (dataloader.py:229-239), the input of bev_scatter."""
    F_, N, S = points.shape
    assert points.is_contiguous() and points.dtype == torch.float32 and pose.dtype == torch.float64 and tuple(pose.shape) == (F_, 6)
    out = torch.empty((F_, N, 3), device=points.device, dtype=torch.float32)
    lib().lidar_ego_transform_f64(_p(points), S, pose.contiguous().data_ptr(), _p(out), F_, N, _st())
    return out


def bev_pack_u8(hist):
    """float32 histogram (values k / 5) -> uint8 counts k, same shape"""
    assert hist.is_contiguous() and hist.dtype == torch.float32
    out = torch.empty(hist.shape, device=hist.device, dtype=torch.uint8)
    lib().bev_pack_u8(_p(hist), out.data_ptr(), hist.numel(), _st())
    return out


def bev_unpack_u8(counts, out=None):
    """uint8 counts -> the float32 histogram the model consumes (bit-identical to bev_scatter's output)"""
    assert counts.is_contiguous() and counts.dtype == torch.uint8
    if out is None:
        out = torch.empty(counts.shape, device=counts.device, dtype=torch.float32)
    lib().bev_unpack_u8(counts.data_ptr(), _p(out), counts.numel(), _st())
    return out


def radar_adjacency(az64):
    """az64 (B, R) float64 azimuths -> (B, R, R) fp32 with adj[b, i, j] = az[b, j] - az[b, i] (dataloader.py:379-384)"""
    B, R = az64.shape
    assert az64.is_contiguous() and az64.dtype == torch.float64
    adj = torch.empty((B, R, R), device=az64.device, dtype=torch.float32)
    lib().radar_adjacency_f64(az64.data_ptr(), _p(adj), B, R, _st())
    return adj
