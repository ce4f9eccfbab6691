"""GPU op-level parity: every C-ABI kernel family against a plain fp32 torch reference evaluated on
the CPU, on the layer shapes the MMFN step uses.  Tolerances: exact-fp32 kernels 1e-4 relative to the
tensor's max; TF32 tensor-core kernels 3e-3 (10-bit mantissa inputs, fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def close(got, ref, tol):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1.0)
    assert err <= tol * scale, (err, scale)


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    from mmfn_b200 import ops
    yield
    ops.TF32 = True


# fwd / dgrad / wgrad of every distinct conv geometry of the three trunks (B=4 to keep the CPU reference fast)
CONVS = [(4, 64, 64, 64, 3, 1, 1), (4, 64, 64, 128, 3, 2, 1), (4, 64, 64, 128, 1, 2, 0), (4, 32, 128, 128, 3, 1, 1),
         (4, 32, 128, 256, 3, 2, 1), (4, 32, 128, 256, 1, 2, 0), (4, 16, 256, 256, 3, 1, 1), (4, 16, 256, 512, 3, 2, 1),
         (4, 16, 256, 512, 1, 2, 0), (4, 8, 512, 512, 3, 1, 1), (2, 256, 3, 64, 7, 2, 3), (2, 256, 2, 64, 7, 2, 3)]


@pytest.mark.parametrize("tf32", [False, True], ids=["fp32", "tf32"])
@pytest.mark.parametrize("geom", CONVS, ids=lambda g: f"N{g[0]}H{g[1]}C{g[2]}-{g[3]}R{g[4]}s{g[5]}")
def test_conv_fwd_dgrad_wgrad(geom, tf32):
    from mmfn_b200 import ops
    N, H, C, Co, R, stride, pad = geom
    if not tf32 and H * C * Co > 64 * 64 * 128:
        pytest.skip("exact SIMT path is covered on the smaller geometries")
    ops.TF32 = tf32
    tol = 3e-3 if tf32 else 1e-4
    x = torch.randn(N, C, H, H)
    w = torch.randn(Co, C, R, R) * (2.0 / (C * R * R)) ** 0.5
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=pad)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wk = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    close(ops.conv2d_fwd(xn, wk, stride, pad).permute(0, 3, 1, 2), yr, tol)
    if C >= 32:                                     # the stems never need a data gradient
        res = torch.randn_like(xn)
        dx = ops.conv2d_dgrad(dyn, wk, xn.shape, stride, pad, res=res)
        close(dx.permute(0, 3, 1, 2), xr.grad + res.cpu().permute(0, 3, 1, 2), tol)
    dw = torch.zeros_like(wk)
    ops.conv2d_wgrad_(dyn, xn, dw, stride, pad)
    close(dw.permute(0, 3, 1, 2), wr.grad, tol)


@pytest.mark.parametrize("C", [3, 2])
def test_stem_conv_as_im2col_tensor_core_gemm(C):
    """The 7x7/2 stems on the production path: im2col + dense TF32 GEMM (forward), split-K GEMM over the kept
    column matrix (weight gradient, accumulated into a packed (Co, R*S*C) filter gradient through ragged tiles)."""
    from mmfn_b200 import ops
    N, H, Co, R = 2, 256, 64, 7
    x = torch.randn(N, C, H, H)
    w = torch.randn(Co, C, R, R) * (2.0 / (C * R * R)) ** 0.5
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=3)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wk = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    assert ops.stem_uses_im2col(xn, wk)
    y, col, w_pad = ops.conv2d_fwd_im2col(xn, wk, 2, 3)
    close(y.permute(0, 3, 1, 2), yr, 3e-3)
    assert col.shape == (N * 128 * 128, (R * R * C + 31) // 32 * 32)
    assert torch.count_nonzero(col[:, R * R * C:]).item() == 0
    y2, _, w_pad2 = ops.conv2d_fwd_im2col(xn, wk, 2, 3, w_pad)          # re-used padded filter buffer
    assert w_pad2 is w_pad and torch.equal(y2, y)
    dw = torch.ones_like(wk)                                             # accumulates on top of existing gradient
    ops.conv2d_wgrad_im2col_(dy.permute(0, 2, 3, 1).contiguous().to(DEV), col, dw)
    close(dw.permute(0, 3, 1, 2) - 1.0, wr.grad, 3e-3)
    # generic (run-time filter geometry) column builder: 3x3 stride-1 on 4 channels
    x4 = torch.randn(2, 4, 20, 20)
    w4 = torch.randn(8, 4, 3, 3) * 0.2
    y4, _, _ = ops.conv2d_fwd_im2col(x4.permute(0, 2, 3, 1).contiguous().to(DEV), w4.permute(0, 2, 3, 1).contiguous().to(DEV), 1, 1)
    close(y4.permute(0, 3, 1, 2), F.conv2d(x4, w4, stride=1, padding=1), 3e-3)


@pytest.mark.parametrize("geom", [(2, 256, 256), (1, 37, 150), (3, 9, 13), (1, 130, 258)], ids=lambda g: "N%dH%dW%d" % g)
@pytest.mark.parametrize("C", [3, 2])
def test_stem_conv_direct_no_column_matrix(C, geom):
    """ops.conv2d_stem7_fwd / conv2d_stem7_wgrad_ (csrc/stem_conv.cu: 7x7 / 2 / pad 3, C -> 64, operand gathered from a
    shared-memory input patch, TF32 mma.sync) against torch fp32 conv2d: ragged tiles (output sizes that are not multiples
    of the 2 x 64 tile), images smaller than one tile, fp32 and bf16 gradients, accumulation on top of an existing dw."""
    from mmfn_b200 import ops
    N, H, W = geom
    x = torch.randn(N, C, H, W)
    w = torch.randn(64, C, 7, 7) * (2.0 / (C * 49)) ** 0.5
    xr, wr = x.clone(), w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=3)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wk = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    assert ops.stem_conv_direct_ok(xn, wk, 2, 3)
    z = ops.conv2d_stem7_fwd(xn, wk)
    assert z.shape == yr.permute(0, 2, 3, 1).shape
    close(z.permute(0, 3, 1, 2), yr, 3e-3)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    dw = torch.ones_like(wk)
    ops.conv2d_stem7_wgrad_(dyn, xn, dw)
    close(dw.permute(0, 3, 1, 2) - 1.0, wr.grad, 3e-3)
    dw16 = torch.zeros_like(wk)
    ops.conv2d_stem7_wgrad_(dyn.to(torch.bfloat16), xn, dw16)
    close(dw16.permute(0, 3, 1, 2), wr.grad, 1e-2)



@pytest.mark.parametrize("C,T", [(64, 192), (128, 192), (256, 192), (512, 256)])
def test_linear_gemms_and_epilogues_tf32(C, T):
    """qkv / proj / mlp GEMM shapes of one fusion-transformer block, forward and both backward products."""
    from mmfn_b200 import ops
    B = 4
    M = B * T
    x = torch.randn(M, C)
    for (N, K) in [(3 * C, C), (C, C), (4 * C, C), (C, 4 * C)]:
        a = torch.randn(M, K)
        w = torch.randn(N, K) * K ** -0.5
        bias, res, dy = torch.randn(N), torch.randn(M, N), torch.randn(M, N)
        A, W, Bi, Rs, dY = (t.to(DEV) for t in (a, w, bias, res, dy))
        out = torch.empty(M, N, device=DEV)
        ops.gemm(A, W, out, bias=Bi, res=Rs, act=1)
        close(out, torch.relu(a @ w.t() + bias) + res, 3e-3)
        dX = torch.empty(M, K, device=DEV)
        ops.gemm(dY, W.t(), dX, mask=A)                     # dgrad with the ReLU mask of the producer
        close(dX, (dy @ w) * (a > 0), 3e-3)
        dW = torch.ones(N, K, device=DEV)
        ops.gemm(dY.t(), A.t(), dW, accum=1)                # wgrad accumulates into the flat grad buffer
        close(dW, 1 + dy.t() @ a, 3e-3)
    assert x is not None


def test_skinny_gemm_with_long_reduction_splits_k():
    """64 x 64 outputs over K = 32768 (the VectorNet generator data gradient at B >= 64): the tensor-core path must
    split K over CTAs (zeroed output + atomic accumulation) and still add the bias exactly once."""
    from mmfn_b200 import ops
    M, N, K = 64, 64, 32768
    a, w, bias = torch.randn(M, K) * 0.1, torch.randn(N, K) * 0.1, torch.randn(N)
    out = torch.full((M, N), 7.0, device=DEV)                  # stale contents must be overwritten, not accumulated
    ops.gemm(a.to(DEV), w.to(DEV), out, bias=bias.to(DEV))
    close(out, a @ w.t() + bias, 3e-3)


@pytest.mark.parametrize("C,T", [(64, 192), (128, 192), (256, 192), (512, 256)])
def test_attention_products_on_strided_heads(C, T):
    from mmfn_b200 import ops
    B, nh = 3, 4
    hs = C // nh
    qkv = torch.randn(B * T, 3 * C)
    dy = torch.randn(B * T, C)
    dS = torch.randn(B, nh, T, T)
    heads = lambda t2d, i: t2d[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
    k, q, v = (heads(qkv, i) for i in range(3))
    P = torch.softmax(q @ k.transpose(-1, -2) / hs ** 0.5, -1)
    dyh = dy.view(B, T, nh, hs).permute(0, 2, 1, 3)
    g_qkv, g_dy, g_dS, g_P = qkv.to(DEV), dy.to(DEV), dS.to(DEV), P.to(DEV)
    gk, gq, gv = (heads(g_qkv, i) for i in range(3))
    S = torch.empty(B, nh, T, T, device=DEV)
    ops.gemm(gq, gk, S)
    close(S, q @ k.transpose(-1, -2), 3e-3)
    y = torch.empty(B * T, C, device=DEV)
    yh = y.view(B, T, nh, hs).permute(0, 2, 1, 3)
    ops.gemm(g_P, gv.transpose(-1, -2), yh)
    close(yh, P @ v, 3e-3)
    g_dyh = g_dy.view(B, T, nh, hs).permute(0, 2, 1, 3)
    dqkv = torch.zeros_like(g_qkv)
    dk, dq, dv = (heads(dqkv, i) for i in range(3))
    ops.gemm(g_P.transpose(-1, -2), g_dyh.transpose(-1, -2), dv)
    ops.gemm(g_dS, gk.transpose(-1, -2), dq)
    ops.gemm(g_dS.transpose(-1, -2), gq.transpose(-1, -2), dk)
    close(dv, P.transpose(-1, -2) @ dyh, 3e-3)
    close(dq, dS @ k, 3e-3)
    close(dk, dS.transpose(-1, -2) @ q, 3e-3)
    # softmax forward / backward (scale, no dropout)
    p, pd = ops.softmax_fwd(S, hs ** -0.5)
    close(p, torch.softmax(S.cpu() * hs ** -0.5, -1), 1e-5)
    Sr = S.cpu().clone().requires_grad_(True)
    torch.softmax(Sr * hs ** -0.5, -1).backward(dS)
    close(ops.softmax_bwd(p, g_dS, hs ** -0.5), Sr.grad, 1e-4)


@pytest.mark.parametrize("C,T,B", [(64, 192, 3), (128, 192, 3), (256, 192, 2), (512, 256, 2), (512, 256, 16)])
def test_attention_fwd_fused(C, T, B):
    """mmfn_attention_fwd_tf32 (S = QK^T in TMEM -> softmax -> dropout -> PV, one tcgen05 kernel; SelfAttention.forward,
    model_rad.py:96-105) against plain fp32 torch for the four transformer geometries (+ the benchmarked B=16 grid of
    transformer4), without and with dropout: P == softmax, Pd == P * mask / (1 - p) with the library's own mask,
    y == Pd V."""
    from mmfn_b200 import ops
    nh = 4
    hs = C // nh
    qkv = torch.randn(B * T, 3 * C)
    heads = lambda t2d, i: t2d[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
    k, q, v = (heads(qkv, i) for i in range(3))
    Pr = torch.softmax(q @ k.transpose(-1, -2) / hs ** 0.5, -1)
    yr = (Pr @ v).permute(0, 2, 1, 3).reshape(B * T, C)
    g_qkv = qkv.to(DEV)
    assert ops.attention_fwd_ok(T, C, nh)
    y, P, Pd = ops.attention_fwd(g_qkv, B, T, C, nh)
    assert Pd is P                                                  # no dropout: one probability tensor
    close(P, Pr, 2e-3)
    close(y, yr, 3e-3)
    close(P.sum(-1), torch.ones(B, nh, T), 1e-5)                    # rows are normalised in fp32 whatever the TF32 scores
    p_drop, seed = 0.1, 1234
    y2, P2, Pd2 = ops.attention_fwd(g_qkv, B, T, C, nh, p_drop, seed)
    mask = ops.dropout(torch.ones_like(P2), p_drop, seed)           # same (seed, index) hash as the kernel
    close(P2, Pr, 2e-3)
    assert torch.equal(Pd2, P2 * mask)
    assert abs((mask == 0).float().mean().item() - p_drop) < 0.01
    close(y2, ((Pr * mask.cpu()) @ v).permute(0, 2, 1, 3).reshape(B * T, C), 3e-3)
    # the saved tensors are exactly what the (unfused) backward consumes: dV = Pd^T dY through the batched GEMM
    dy = torch.randn(B * T, C, device=DEV)
    dv = torch.empty(B, nh, T, hs, device=DEV)
    ops.gemm(Pd2.transpose(-1, -2), dy.view(B, T, nh, hs).permute(0, 2, 1, 3).transpose(-1, -2), dv)
    close(dv, (Pr * mask.cpu()).transpose(-1, -2) @ dy.cpu().view(B, T, nh, hs).permute(0, 2, 1, 3), 3e-3)


@pytest.mark.parametrize("C,T", [(64, 192), (128, 192), (256, 192), (512, 256)])
def test_fused_attention_backward_dq_ds(C, T):
    """attention_bwd_dq (dP in TMEM -> dS -> dQ in one kernel) against torch autograd, without dropout, and against the
    unfused kernel sequence (batched GEMM + softmax_bwd + batched GEMM) with the SAME dropout masks (p = 0.1)."""
    import math
    from mmfn_b200 import ops
    B, nh = 3, 4
    hs = C // nh
    qkv = (torch.randn(B * T, 3 * C) * 0.5)
    dy = torch.randn(B * T, C)
    heads = lambda t2d, i: t2d[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
    qkv_r = qkv.clone().requires_grad_(True)
    k, q, v = (heads(qkv_r, i) for i in range(3))
    P = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hs), -1)
    y = (P @ v).permute(0, 2, 1, 3).reshape(B * T, C)
    (dS_ref,) = torch.autograd.grad(y, P, dy, retain_graph=True)          # = dP; dS below
    y.backward(dy)
    g_qkv, g_dy, g_P, g_y = qkv.to(DEV), dy.to(DEV), P.detach().contiguous().to(DEV), y.detach().to(DEV)
    dqkv = torch.zeros_like(g_qkv)
    dS = ops.attention_bwd_dq(g_qkv, g_dy, g_y, g_P, dqkv, B, T, C, nh)
    dP = dS_ref
    ds_ref = P.detach() * (dP - (dP * P.detach()).sum(-1, keepdim=True)) / math.sqrt(hs)
    close(dS, ds_ref, 3e-3)
    close(heads(dqkv, 1), heads(qkv_r.grad, 1), 3e-3)                      # query gradient
    assert torch.count_nonzero(heads(dqkv, 0)).item() == 0               # key / value slices untouched
    # with dropout: same hash as the stand-alone kernels
    p_drop, seed = 0.1, 11
    gk, gq, gv = (heads(g_qkv, i) for i in range(3))
    g_dyh = g_dy.view(B, T, nh, hs).permute(0, 2, 1, 3)
    Pd = ops.dropout(g_P, p_drop, seed)
    yd = torch.empty(B * T, C, device=DEV)
    ops.gemm(Pd, gv.transpose(-1, -2), yd.view(B, T, nh, hs).permute(0, 2, 1, 3))
    dPd = torch.empty(B, nh, T, T, device=DEV)
    ops.gemm(g_dyh, gv, dPd)
    dS_u = ops.softmax_bwd(g_P, dPd, 1.0 / math.sqrt(hs), p_drop, seed)
    dq_u = torch.zeros_like(g_qkv)
    ops.gemm(dS_u, gk.transpose(-1, -2), heads(dq_u, 1))
    dq_f = torch.zeros_like(g_qkv)
    dS_f = ops.attention_bwd_dq(g_qkv, g_dy, yd, g_P, dq_f, B, T, C, nh, p_drop, seed)
    close(dS_f, dS_u, 3e-3)
    close(dq_f, dq_u, 5e-3)


def test_batchnorm_train_forward_backward():
    from mmfn_b200 import ops
    # M = N*H*H <= 4096 rows: single-launch kernel; larger: two launches whose scratch must come back zeroed (the
    # big shapes run back to back on one stream); C = 96 exercises the non-power-of-two channel-quad indexing
    for (N, H, C) in [(4, 32, 128), (16, 8, 512), (2, 64, 64), (3, 72, 96), (2, 64, 64)]:
        x = torch.randn(N, C, H, H) * 2 + 0.5
        res = torch.randn(N, C, H, H)
        bn = torch.nn.BatchNorm2d(C).train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
        rm, rv = bn.running_mean.clone().to(DEV), bn.running_var.clone().to(DEV)
        xr, rr = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        yr = torch.relu(bn(xr) + rr)
        dy = torch.randn_like(yr)
        yr.backward(dy)
        xn, rn = (t.permute(0, 2, 3, 1).contiguous().to(DEV) for t in (x, res))
        g, b = bn.weight.data.to(DEV), bn.bias.data.to(DEV)
        y, mean, rstd = ops.bn_train_fwd(xn, g, b, rm, rv, res=rn, relu=True)
        close(y.permute(0, 3, 1, 2), yr, 1e-4)
        close(rm, bn.running_mean, 1e-5); close(rv, bn.running_var, 1e-5)
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dx, dres = ops.bn_train_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), xn, y, mean, rstd, g, dg, db, want_dres=True)
        close(dx.permute(0, 3, 1, 2), xr.grad, 1e-4)
        close(dres.permute(0, 3, 1, 2), rr.grad, 1e-5)
        close(dg, bn.weight.grad, 1e-4); close(db, bn.bias.grad, 1e-4)
        # the ReLU mask read from the bf16 twin of y instead of y: same signs, same result
        dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
        dg2, db2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dx2, dres2 = ops.bn_train_bwd(dyn, xn, y.to(torch.bfloat16), mean, rstd, g, dg2, db2, want_dres=True)
        assert torch.equal(dres2, dres)
        close(dx2, dx, 1e-6); close(dg2, dg, 1e-6); close(db2, db, 1e-6)
        # conv -> BN -> ReLU without a residual: the mask recomputed from x (relu_beta) equals the mask read from y
        y0, mean0, rstd0 = ops.bn_train_fwd(xn, g, b, rm.clone(), rv.clone(), relu=True)
        dga, dba, dgb, dbb = (torch.zeros(C, device=DEV) for _ in range(4))
        dxa, _ = ops.bn_train_bwd(dyn, xn, y0, mean0, rstd0, g, dga, dba)
        dxb, _ = ops.bn_train_bwd(dyn, xn, None, mean0, rstd0, g, dgb, dbb, relu_beta=b)
        close(dxa, dxb, 1e-6); close(dga, dgb, 1e-6); close(dba, dbb, 1e-6)


@pytest.mark.parametrize("geom", [(2, 64, 64, 64), (3, 31, 29, 8), (1, 8, 10, 128), (2, 33, 64, 96)], ids=lambda g: "N%dH%dW%dC%d" % g)
def test_stem_tail_bn_relu_maxpool_fused(geom):
    """ops.stem_bn_relu_maxpool_fwd / _bwd (BatchNorm -> ReLU -> MaxPool(3, 2, 1) of the ResNet stems, model_rad.py:513-521)
    against torch autograd, and against the unfused kernels of the library (same values, same arg-max taps).  A negative
    BatchNorm bias makes most windows tie at zero (first-tap rule + ReLU mask in bit 7 of the arg-max byte)."""
    from mmfn_b200 import ops
    N, H, W, C = geom
    for bias_shift in (0.0, -1.5):
        x = torch.randn(N, C, H, W) * 2 + 0.5
        bn = torch.nn.BatchNorm2d(C).train()
        with torch.no_grad():
            bn.weight.uniform_(-1.5, 1.5); bn.bias.normal_().add_(bias_shift)        # negative gammas too
        rm, rv = bn.running_mean.clone().to(DEV), bn.running_var.clone().to(DEV)
        xr = x.clone().requires_grad_(True)
        yr = F.max_pool2d(torch.relu(bn(xr)), 3, 2, 1)
        dy = torch.randn_like(yr)
        yr.backward(dy)
        z = x.permute(0, 2, 3, 1).contiguous().to(DEV)
        g, b = bn.weight.data.to(DEV), bn.bias.data.to(DEV)
        out, saved, mean, rstd = ops.stem_bn_relu_maxpool_fwd(z, g, b, rm, rv, want16=True)
        idx, zmax = saved
        close(out.permute(0, 3, 1, 2), yr, 1e-5)
        close(rm, bn.running_mean, 1e-5); close(rv, bn.running_var, 1e-5)
        assert torch.equal(out.h, out.to(torch.bfloat16))
        # unfused library path: identical pooled values; identical taps; bit 7 <=> pooled value is not positive
        y = ops.bn_apply(z, g, b, mean, rstd, relu=True)
        out2, idx2 = ops.maxpool_fwd(y)
        assert torch.equal(out, out2)
        assert torch.equal(idx & 0x7F, idx2)
        assert torch.equal((idx & 0x80) != 0, ~(out > 0))
        assert torch.equal(ops.bn_apply(zmax, g, b, mean, rstd, relu=True), out)       # zmax = z at the arg-max
        dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
        for bf in (False, True):
            dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            dz = ops.stem_bn_relu_maxpool_bwd(dyn, saved, z, mean, rstd, g, dg, db, out_bf16=bf)
            close(dz.permute(0, 3, 1, 2), xr.grad, 1e-2 if bf else 1e-4)
            close(dg, bn.weight.grad, 1e-4); close(db, bn.bias.grad, 1e-4)
            # the unfused chain of the library computes the same gradient
            dg2, db2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
            dz2, _ = ops.bn_train_bwd(ops.maxpool_bwd(dyn, idx2, y.shape), z, y, mean, rstd, g, dg2, db2, out_bf16=bf)
            close(dz, dz2, 1e-2 if bf else 1e-5); close(dg, dg2, 1e-5); close(db, db2, 1e-5)



def test_column_sums_scalar_and_16_byte_load_kernels():
    """ops.colsum_ (bias gradients): fp32 / bf16, the 16-byte-load kernel (rows >= 512, widths and pitches that are
    multiples of 4 / 8 elements) and the scalar kernel (everything else, e.g. an odd width or a column slice that starts at
    an unaligned element), accumulation on top of `out`."""
    from mmfn_b200 import ops
    for M, N, dt in [(4100, 512, torch.float32), (8192, 2048, torch.bfloat16), (3072, 192, torch.bfloat16), (600, 68, torch.float32),
                     (300, 512, torch.float32), (2000, 50, torch.bfloat16), (1024, 1536, torch.bfloat16)]:
        x = torch.randn(M, N, device=DEV).to(dt)
        out = torch.ones(N, device=DEV)
        ops.colsum_(x, out)
        close(out - 1.0, x.double().sum(0), 2e-5)
    # strided views: a column slice of a wider matrix (aligned start -> vector kernel, unaligned start -> scalar kernel)
    wide = torch.randn(2048, 3 * 256, device=DEV)
    for c0 in (256, 258):
        v = wide[:, c0: c0 + 256]
        out = torch.zeros(256, device=DEV)
        ops.colsum_(v, out)
        close(out, v.double().sum(0), 2e-5)


def test_segment_max_forward_backward_scalar_and_vector_kernels():
    """ops.segmax_fwd / segmax_bwd (polyline max-pool of VectorNet, model_rad.py:281-283) against torch: C % 4 == 0 takes
    the 16-byte-store backward, an odd width the scalar one."""
    from mmfn_b200 import ops
    for G, V, C in [(37, 9, 128), (6, 19, 128), (5, 19, 6), (300, 9, 64), (3, 7, 5)]:
        x = torch.randn(G, V, C)
        xr = x.clone().requires_grad_(True)
        out_r = xr.max(dim=1).values
        dout = torch.randn_like(out_r)
        out_r.backward(dout)
        out, arg = ops.segmax_fwd(x.view(G * V, C).to(DEV), G, V)
        close(out, out_r, 1e-6)
        dx = ops.segmax_bwd(dout.to(DEV), arg, G, V)
        assert torch.equal(dx.view(G, V, C).cpu(), xr.grad)


def test_polyline_pool_backward_and_input_layer_weight_gradient():
    """ops.subgraph_pool_fwd / _bwd ([h | max_v h] of model_rad.py:270-283; 16-byte and scalar backward kernels) against torch
    autograd, and ops.wgrad_n64_k7_ (weight gradient of the 7 -> 64 polyline input layer) against dy^T x in float64."""
    from mmfn_b200 import ops
    for G, V, C in [(41, 9, 64), (7, 19, 64), (5, 9, 6)]:
        x = torch.randn(G, V, C)
        xr = x.clone().requires_grad_(True)
        yr = torch.cat([xr, xr.max(dim=1, keepdim=True).values.expand(-1, V, -1)], dim=-1)
        dy = torch.randn_like(yr)
        yr.backward(dy)
        y, arg = ops.subgraph_pool_fwd(x.view(G * V, C).to(DEV), G, V)
        close(y.view(G, V, 2 * C), yr, 1e-6)
        dx = ops.subgraph_pool_bwd(dy.view(G * V, 2 * C).to(DEV), arg, G, V)
        close(dx.view(G, V, C), xr.grad, 1e-6)
    for M in (4096, 36864, 5001):
        dy, xin = torch.randn(M, 64, device=DEV), torch.randn(M, 7, device=DEV)
        dw = torch.ones(64, 7, device=DEV)
        ops.wgrad_n64_k7_(dy, xin, dw)
        close(dw - 1.0, dy.double().t() @ xin.double(), 1e-5)

def test_layernorm_variants():
    from mmfn_b200 import ops
    # rows >= 1024 with C % 128 == 0 take the 16-byte-load parameter-gradient kernel
    for C, act, M in [(64, 0, 300), (64, 1, 300), (128, 2, 300), (512, 0, 300), (256, 1, 2050), (512, 0, 4100), (128, 2, 1030),
                      (64, 1, 5003), (64, 0, 4099), (64, 2, 8192)]:       # C = 64, rows >= 4096: the half-warp-per-row kernels
        x = torch.randn(M, C) * 1.5 + 0.3
        ln = torch.nn.LayerNorm(C)
        with torch.no_grad():
            ln.weight.uniform_(0.5, 1.5); ln.bias.normal_()
        xr = x.clone().requires_grad_(True)
        yr = ln(xr)
        yr = torch.relu(yr) if act == 1 else (F.gelu(yr) if act == 2 else yr)
        dy, dres = torch.randn_like(yr), torch.randn_like(yr)
        yr.backward(dy)
        g, b = ln.weight.data.to(DEV), ln.bias.data.to(DEV)
        y, mean, rstd = ops.layernorm_fwd(x.to(DEV), g, b, act=act)
        close(y, yr, 1e-5)
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dx = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g, b, mean, rstd, dg, db, act=act, dres=dres.to(DEV))
        close(dx, xr.grad + dres, 1e-4)
        close(dg, ln.weight.grad, 1e-4); close(db, ln.bias.grad, 1e-4)
        # split form: data gradient (+ the dropout-masked copy for the next residual branch) and parameter
        # gradients as two independent launches
        dg2, db2 = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        assert ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g, b, mean, rstd, dg2, db2, act=act, parts=2) is None
        close(dg2, dg, 1e-5); close(db2, db, 1e-5)
        dx1, dxd = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g, b, mean, rstd, None, None, act=act, dres=dres.to(DEV),
                                     parts=1, drop=(0.25, 77))
        assert torch.equal(dx1, dx)
        assert torch.equal(dxd, ops.dropout(dx, 0.25, 77))
        dx1, dxd = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g, b, mean, rstd, None, None, act=act, parts=1, drop=(0.0, 0))
        assert dxd is dx1


def test_pool_token_upsample_kernels():
    from mmfn_b200 import ops
    B, C = 3, 32
    for H in (64, 32, 16, 8):
        feats = [torch.randn(B, C, H, H, requires_grad=True) for _ in range(3)]
        pos = torch.randn(1, 192, C, requires_grad=True)
        vw, vb = torch.randn(C, 1, requires_grad=True), torch.randn(C, requires_grad=True)
        vel = torch.rand(B) * 10
        pooled = [F.adaptive_avg_pool2d(f, (8, 8)) for f in feats]
        tokr = torch.cat([p.view(B, 1, C, 8, 8) for p in pooled], 1).permute(0, 1, 3, 4, 2).reshape(B, -1, C)
        tokr = pos + tokr + F.linear(vel.unsqueeze(1), vw, vb).unsqueeze(1)
        dtok = torch.randn_like(tokr)
        tokr.backward(dtok)
        fn = [f.detach().permute(0, 2, 3, 1).contiguous().to(DEV) for f in feats]
        tok = ops.tokens_fwd(fn, pos.detach()[0].to(DEV), vw.detach()[:, 0].contiguous().to(DEV), vb.detach().to(DEV), vel.to(DEV))
        close(tok, tokr, 1e-5)
        df = [torch.zeros_like(f) for f in fn]
        dpos, dvw, dvb = torch.zeros(192, C, device=DEV), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        ops.tokens_bwd_(dtok.to(DEV), df, fn[0].shape, vel.to(DEV), dpos, dvw, dvb)
        for m in range(3):
            close(df[m].permute(0, 3, 1, 2), feats[m].grad, 1e-5)
        close(dpos, pos.grad[0], 1e-4); close(dvw, vw.grad[:, 0], 1e-4); close(dvb, vb.grad, 1e-4)
        if H > 8:
            feat = torch.randn(B, C, H, H)
            tk = torch.randn(B, 192, C, requires_grad=True)
            g = tk[:, 64:128].view(B, 8, 8, C).permute(0, 3, 1, 2)
            outr = feat + F.interpolate(g, scale_factor=H // 8, mode="bilinear", align_corners=True)
            dA = torch.randn_like(outr)
            outr.backward(dA)
            out = ops.upsample_add_fwd(feat.permute(0, 2, 3, 1).contiguous().to(DEV), tk.detach().to(DEV), 1)
            close(out.permute(0, 3, 1, 2), outr, 1e-5)
            dtk = torch.zeros(B, 192, C, device=DEV)
            ops.upsample_add_bwd_(dA.permute(0, 2, 3, 1).contiguous().to(DEV), dtk, 1)
            close(dtk[:, 64:128], tk.grad[:, 64:128], 1e-4)
    x = torch.relu(torch.randn(2, 64, 128, 128))
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    y, idx = ops.maxpool_fwd(xn)
    close(y.permute(0, 3, 1, 2), yr, 0)
    close(ops.maxpool_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), idx, xn.shape).permute(0, 3, 1, 2), xr.grad, 1e-6)


def test_gat_vectornet_and_head_kernels():
    from mmfn_b200 import ops
    # radar GAT masked softmax incl. fully-masked rows (uniform 1/81) and the log-softmax relayout
    z, adj = torch.randn(3, 81, 81), torch.randn(3, 81, 81)
    adj[:, 5] = -1
    zr = z.clone().requires_grad_(True)
    e = F.leaky_relu(zr, 0.2)
    attr = torch.softmax(torch.where(adj > 0, e, torch.full_like(e, -9e15)), -1)
    d = torch.randn_like(attr)
    attr.backward(d)
    att, _ = ops.gat_softmax_fwd(z.to(DEV), adj.to(DEV), 0.2)
    close(att, attr, 1e-6)
    assert abs(att[0, 5].sum().item() - 1.0) < 1e-5 and abs(att[0, 5, 0].item() - 1 / 81) < 1e-6
    close(ops.gat_softmax_bwd(z.to(DEV), adj.to(DEV), att, d.to(DEV), 0.2), zr.grad, 1e-5)
    v = torch.randn(3, 256, 128)
    vr = v.clone().requires_grad_(True)
    yr = F.log_softmax(vr.view(3, 8, 8, 512).transpose(1, 3), dim=1)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    yy = ops.radar_logsoftmax_fwd(v.to(DEV), 3, 512)
    close(yy.permute(0, 3, 1, 2), yr, 1e-5)
    close(ops.radar_logsoftmax_bwd(dy.permute(0, 2, 3, 1).contiguous().to(DEV), yy).view(3, 256, 128), vr.grad, 1e-5)
    # VectorNet: lane-to-lane attention restricted to query row 0, with ragged lane counts (incl. 1 lane)
    B, L = 3, 40
    qkv = torch.randn(B, L, 384)
    ln = torch.tensor([40, 17, 1], dtype=torch.int32)
    qr = qkv.clone().requires_grad_(True)
    q, k, vv = [t.view(B, L, 2, 64).transpose(1, 2) for t in qr.chunk(3, -1)]
    dots = (q @ k.transpose(-1, -2) * 0.125).masked_fill((torch.arange(L)[None, :] >= ln[:, None]).view(B, 1, 1, L), -1e9)
    o = (torch.softmax(dots, -1) @ vv).transpose(1, 2).reshape(B, L, 128)[:, 0]
    do = torch.randn_like(o)
    o.backward(do)
    prob, out = ops.l2l_row0_fwd(qkv.to(DEV), ln.to(DEV), 2, None)
    close(out, o, 1e-5)
    close(ops.l2l_row0_bwd(qkv.to(DEV), ln.to(DEV), prob, do.contiguous().to(DEV), 2), qr.grad, 1e-5)
    # Subgraph segment max-pool + concat
    G, V, C = 12, 9, 64
    h = torch.randn(G * V, C)
    hr = h.clone().view(G, V, C).requires_grad_(True)
    yr2 = torch.cat([hr, hr.max(1)[0].unsqueeze(1).expand(G, V, C)], -1)
    dy2 = torch.randn_like(yr2)
    yr2.backward(dy2)
    yy2, arg = ops.subgraph_pool_fwd(h.to(DEV), G, V)
    close(yy2.view(G, V, 2 * C), yr2, 0)
    close(ops.subgraph_pool_bwd(dy2.reshape(G * V, 2 * C).to(DEV), arg, G, V).view(G, V, C), hr.grad, 1e-6)
    # GRU waypoint head + L1 loss + AdamW
    Bh = 5
    gru, lin = torch.nn.GRUCell(2, 64), torch.nn.Linear(64, 2)
    z0 = torch.randn(Bh, 64, requires_grad=True)
    tp, gt = torch.randn(Bh, 2) * 5, torch.randn(Bh, 4, 2)
    zc, xc, wps = z0, torch.zeros(Bh, 2), []
    for _ in range(4):
        zc = gru(xc + tp, zc); xc = lin(zc) + xc; wps.append(xc)
    predr = torch.stack(wps, 1)
    lossr = F.l1_loss(predr, gt, reduction="none").mean()
    lossr.backward()
    P = [p.detach().to(DEV) for p in (gru.weight_ih, gru.weight_hh, gru.bias_ih, gru.bias_hh, lin.weight, lin.bias)]
    pred, ctx = ops.gru_head_fwd(z0.detach().to(DEV), tp.to(DEV), *P, 4)
    close(pred, predr, 1e-5)
    loss, dpred = ops.l1_loss(pred, gt.to(DEV))
    close(loss, lossr, 1e-6)
    Gd = [torch.zeros_like(p) for p in P]
    close(ops.gru_head_bwd(dpred, ctx, P[0], P[1], P[4], *Gd), z0.grad, 1e-5)
    for g, p in zip(Gd, (gru.weight_ih, gru.weight_hh, gru.bias_ih, gru.bias_hh, lin.weight, lin.bias)):
        close(g, p.grad, 1e-5)
    n = 4096
    p0, g0 = torch.randn(n), torch.randn(n)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-4)
    p, g = p0.to(DEV), g0.to(DEV)
    m, v2, st = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(3, device=DEV)
    for _ in range(3):
        pr.grad = g0.clone(); opt.step()
        ops.adamw_step_(p, g, m, v2, st, 1e-4)
    close(p, pr, 1e-6)


def test_dropout_masks_are_consistent_between_kernels_and_passes():
    """The same (seed, index) hash drives the GEMM epilogue (SIMT and tcgen05) and the stand-alone kernel used
    in backward, so forward and backward see identical masks; keep-rate ~ 1-p."""
    from mmfn_b200 import ops
    M, N, K = 1024, 256, 128
    A, W = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV)
    c1, c2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.TF32 = True
    ops.gemm(A, W, c1, drop_p=0.1, seed=5)
    ops.TF32 = False
    ops.gemm(A, W, c2, drop_p=0.1, seed=5)
    ones = torch.ones(M, N, device=DEV)
    mask = ops.dropout(ones, 0.1, 5)
    assert torch.equal(c1 == 0, mask == 0) and torch.equal(c2 == 0, mask == 0)
    assert abs((mask == 0).float().mean().item() - 0.1) < 0.01
    assert torch.allclose(mask[mask != 0], torch.tensor(1 / 0.9, device=DEV))


def test_dropout_masks_follow_the_bound_rng_offset_in_every_translation_unit():
    """Round-1 advisor finding: norm.cu never bound the device-resident rng offset, so with an engine alive the
    LayerNorm-backward residual-dropout mask differed from the forward GEMM-epilogue mask.  Bump the process-wide
    offset to a non-zero value and check every dropout site against the stand-alone kernel (misc.cu)."""
    from mmfn_b200 import ops
    from mmfn_b200._lib import lib
    rng = lib().rng_tensor(torch.device(DEV))
    old = rng.clone()
    try:
        rng.add_(1000003 * 17)
        p, seed = 0.1, 77
        M, C = 384, 256
        ones = torch.ones(M, C, device=DEV)
        mask = ops.dropout(ones, p, seed)                             # misc.cu
        assert abs((mask == 0).float().mean().item() - p) < 0.02
        rng.add_(-1000003 * 17)
        mask0 = ops.dropout(ones, p, seed)
        rng.add_(1000003 * 17)
        assert not torch.equal(mask0, mask)                           # the offset really changes the masks
        # GEMM epilogues (tcgen05 and SIMT)
        A, W = torch.randn(M, 128, device=DEV), torch.randn(C, 128, device=DEV)
        for tf32 in (True, False):
            ops.TF32 = tf32
            c = torch.empty(M, C, device=DEV)
            ops.gemm(A, W, c, drop_p=p, seed=seed)
            assert torch.equal(c == 0, mask == 0), tf32
        ops.TF32 = True
        # LayerNorm backward's fused dropout copy (norm.cu)
        x, dy = torch.randn(M, C, device=DEV), torch.randn(M, C, device=DEV)
        g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
        _, mean, rstd = ops.layernorm_fwd(x, g, b)
        dx, dxd = ops.layernorm_bwd(dy, x, g, b, mean, rstd, None, None, parts=1, drop=(p, seed))
        assert torch.equal(dxd, dx * mask)
        # fused attention probabilities (attn_tc.cu) and the softmax kernels (attn.cu)
        B, T, Cc, nh = 2, 192, 128, 4
        qkv = torch.randn(B * T, 3 * Cc, device=DEV)
        _, P, Pd = ops.attention_fwd(qkv, B, T, Cc, nh, p, seed)
        pm = ops.dropout(torch.ones_like(P), p, seed)
        assert torch.equal(Pd, P * pm)
        S = torch.randn(B, nh, T, T, device=DEV)
        p2, pd2 = ops.softmax_fwd(S, 1.0, p, seed)
        assert torch.equal(pd2, p2 * pm)
    finally:
        rng.copy_(old)


def test_block_gradients_with_residual_dropout_and_bound_rng():
    """One transformer Block (model_rad.py:112-133) forward + backward with resid_pdrop = attn_pdrop = 0.1 and a
    non-zero rng offset bound, exact-fp32 kernels, against torch autograd on the SAME block with the library's own
    masks (ops.dropout of ones): the hand-written backward must regenerate exactly the forward's masks in every
    translation unit (GEMM epilogues, softmax, LayerNorm backward)."""
    import math
    from mmfn_b200 import ops
    from mmfn_b200._lib import lib
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.model_rad import Block, _Aux
    from mmfn_b200.params import ParamStore
    ops.TF32 = False
    rng = lib().rng_tensor(torch.device(DEV))
    old = rng.clone()
    try:
        rng.add_(1000003 * 5)
        cfg = GlobalConfig()
        st = ParamStore(cfg, DEV)
        torch.manual_seed(1)
        st.flat.copy_(torch.randn_like(st.flat) * 0.05)
        pre = "encoder.transformer2.blocks.0"
        C, B, T, nh, p = 128, 2, 192, 4, 0.1
        hs = C // nh
        blk = Block(st, pre, C, nh, p, p)
        blk.ln1.g.add_(1.0); blk.ln2.g.add_(1.0)
        x = torch.randn(B * T, C, device=DEV)
        w = torch.randn(B * T, C, device=DEV)                     # loss = <w, block(x)>
        seed = 4242
        out = blk.fwd(x, B, T, seed, True)
        st.flat_grad.zero_()
        dz = ops.dropout(w, p, seed + 2)                          # gradient entering the fc2 residual-branch dropout
        dx, _ = blk.bwd(w, dz, (0.0, 0))
        _Aux.join_all()
        torch.cuda.synchronize()
        # ---- torch autograd replica with the library's masks
        m_att = ops.dropout(torch.ones(B, nh, T, T, device=DEV), p, seed)
        m_proj = ops.dropout(torch.ones(B * T, C, device=DEV), p, seed + 1)
        m_fc2 = ops.dropout(torch.ones(B * T, C, device=DEV), p, seed + 2)
        leaves = {n: t.detach().clone().requires_grad_(True) for n, t in
                  dict(x=x, wqkv=blk.qkv.w, bqkv=blk.qkv.b, wp=blk.proj.w, bp=blk.proj.b, w1=blk.fc1.w, b1=blk.fc1.b,
                       w2=blk.fc2.w, b2=blk.fc2.b, g1=blk.ln1.g, c1=blk.ln1.b, g2=blk.ln2.g, c2=blk.ln2.b).items()}
        L = leaves
        h1 = F.layer_norm(L["x"], (C,), L["g1"], L["c1"])
        qkv = h1 @ L["wqkv"].t() + L["bqkv"]
        heads = lambda i: qkv[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
        k, q, v = heads(0), heads(1), heads(2)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hs), -1) * m_att
        y = (att @ v).permute(0, 2, 1, 3).reshape(B * T, C)
        x1 = L["x"] + (y @ L["wp"].t() + L["bp"]) * m_proj
        h2 = F.layer_norm(x1, (C,), L["g2"], L["c2"])
        x2 = x1 + (torch.relu(h2 @ L["w1"].t() + L["b1"]) @ L["w2"].t() + L["b2"]) * m_fc2
        close(out, x2, 1e-4)
        (x2 * w).sum().backward()
        close(dx, L["x"].grad, 2e-4)
        for name, got in dict(wqkv=blk.qkv.dw, bqkv=blk.qkv.db, wp=blk.proj.dw, bp=blk.proj.db, w1=blk.fc1.dw, b1=blk.fc1.db,
                              w2=blk.fc2.dw, b2=blk.fc2.db, g1=blk.ln1.dg, c1=blk.ln1.db, g2=blk.ln2.dg, c2=blk.ln2.db).items():
            ref = L[name].grad
            err = (got - ref).norm().item() / max(ref.norm().item(), 1e-9)
            assert err < 1e-3, (name, err)
    finally:
        rng.copy_(old)


# ------------------------------------------------------------------------------------------------ bf16 (BASELINE configs[2])
# Tolerances: operands are rounded to bf16 (8-bit mantissa, relative step 2^-8 = 3.9e-3), products accumulate in fp32.
# Against an fp32 torch reference evaluated ON THE SAME bf16-ROUNDED OPERANDS the only differences are accumulation order
# and the rounding of a bf16 result: 2e-3 (fp32 result) / 8e-3 (bf16 result) relative to the tensor's max.
def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(384, 64, 64), (4096, 2048, 512), (768, 192, 64), (8192, 512, 2048), (200, 72, 40)])
def test_gemm_bf16_all_majors_and_epilogues(M, N, K):
    """mmfn_gemm_bf16: K-major / MN-major operands (forward, data-gradient and weight-gradient forms of nn.Linear),
    fp32 and bf16 results, bias / ReLU / bf16 mask / dropout / residual epilogues, split-K accumulation."""
    from mmfn_b200 import ops
    A, Bm = torch.randn(M, K), torch.randn(N, K) * 0.1
    A16, B16 = _bf(A).to(DEV), _bf(Bm).to(DEV)
    Ar, Br = A16.float().cpu(), B16.float().cpu()
    ref = Ar @ Br.t()
    c = torch.empty(M, N, device=DEV)
    ops.gemm(A16, B16, c)
    close(c, ref, 2e-3)
    c16 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A16, B16, c16)
    close(c16, ref, 8e-3)
    # MN-major operands: A stored (K, M), B stored (K, N)
    At, Bt = A16.t().contiguous(), B16.t().contiguous()
    for a_op, b_op in ((At.t(), B16), (A16, Bt.t()), (At.t(), Bt.t())):
        c.zero_()
        ops.gemm(a_op, b_op, c)
        close(c, ref, 2e-3)
    if N % 8 == 0:
        bias, res = torch.randn(N), torch.randn(M, N)
        mask = _bf(torch.randn(M, N)).to(DEV)
        ops.gemm(A16, B16, c, bias=bias.to(DEV), res=res.to(DEV), act=1)
        close(c, torch.relu(ref + bias) + res, 2e-3)
        ops.gemm(A16, B16, c16, mask=mask)
        close(c16, ref * (mask.float().cpu() > 0), 8e-3)
        ops.gemm(A16, B16, c, bias=bias.to(DEV), drop_p=0.1, seed=9, res=res.to(DEV))
        dm = ops.dropout(torch.ones(M, N, device=DEV), 0.1, 9).cpu()
        close(c, (ref + bias) * dm + res, 2e-3)
    # weight-gradient form: dW (N, K) += dY^T X with the reduction over M rows, both operands MN-major, atomics
    dw = torch.zeros(N, K, device=DEV)
    dy16 = _bf(torch.randn(M, N)).to(DEV)
    ops.gemm(dy16.t(), A16.t(), dw, accum=1)
    close(dw, dy16.float().cpu().t() @ Ar, 2e-3)


@pytest.mark.parametrize("C,T", [(64, 192), (128, 192), (256, 192), (512, 256)])
def test_bf16_result_from_tf32_attention_products(C, T):
    """dV / dK / dQ written as bf16 head slices of dqkv by the TF32 batched GEMMs (mmfn_gemm_tf32_out)."""
    from mmfn_b200 import ops
    B, nh = 2, 4
    hs = C // nh
    qkv, dy, dS = torch.randn(B * T, 3 * C), torch.randn(B * T, C), torch.randn(B, nh, T, T)
    heads = lambda t2d, i: t2d[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
    k, q, v = (heads(qkv, i) for i in range(3))
    P = torch.softmax(q @ k.transpose(-1, -2) / hs ** 0.5, -1)
    dyh = dy.view(B, T, nh, hs).permute(0, 2, 1, 3)
    g_qkv, g_dy, g_dS, g_P = qkv.to(DEV), dy.to(DEV), dS.to(DEV), P.to(DEV)
    gk, gq, gv = (heads(g_qkv, i) for i in range(3))
    dqkv = torch.zeros(B * T, 3 * C, device=DEV, dtype=torch.bfloat16)
    dk, dq, dv = (heads(dqkv, i) for i in range(3))
    ops.gemm(g_P.transpose(-1, -2), g_dy.view(B, T, nh, hs).permute(0, 2, 1, 3).transpose(-1, -2), dv)
    ops.gemm(g_dS, gk.transpose(-1, -2), dq)
    ops.gemm(g_dS.transpose(-1, -2), gq.transpose(-1, -2), dk)
    close(dv, P.transpose(-1, -2) @ dyh, 8e-3)
    close(dq, dS @ k, 8e-3)
    close(dk, dS.transpose(-1, -2) @ q, 8e-3)


BF_CONVS = [g for g in CONVS if g[2] % 64 == 0] + [(16, 16, 256, 256, 3, 1, 1), (3, 64, 64, 64, 3, 1, 1), (5, 8, 512, 512, 3, 1, 1)]


@pytest.mark.parametrize("geom", BF_CONVS, ids=lambda g: f"N{g[0]}H{g[1]}C{g[2]}-{g[3]}R{g[4]}s{g[5]}")
def test_conv_bf16_fwd_dgrad_wgrad(geom):
    """mmfn_conv2d_{fwd,dgrad,wgrad}_bf16 on every BasicBlock geometry of the trunks against torch fp32 on the same
    bf16-rounded operands."""
    from mmfn_b200 import ops
    N, H, C, Co, R, stride, pad = geom
    x = _bf(torch.randn(N, C, H, H))
    w = _bf(torch.randn(Co, C, R, R) * (2.0 / (C * R * R)) ** 0.5)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=pad)
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wk = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    ops.BF16 = True
    try:
        y = ops.conv2d_fwd(xn, wk, stride, pad)
        close(y.permute(0, 3, 1, 2), yr, 2e-3)
        dy = _bf(torch.randn_like(yr))
        yr.backward(dy.float())
        dyn = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
        res = torch.randn(N, H, H, C, device=DEV)
        dx = ops.conv2d_dgrad(dyn, wk, xn.shape, stride, pad, res=res)
        close(dx.permute(0, 3, 1, 2), xr.grad + res.cpu().permute(0, 3, 1, 2), 2e-3)
        dw = torch.zeros(Co, R, R, C, device=DEV)
        ops.conv2d_wgrad_(dyn, xn, dw, stride, pad)
        close(dw.permute(0, 3, 1, 2), wr.grad, 2e-3)
    finally:
        ops.BF16 = False


def test_bf16_twins_and_typed_elementwise_outputs():
    """Producers of MMA operands write bf16 next to / instead of fp32 in the same pass: BatchNorm apply (twin),
    BatchNorm backward (bf16 dz), LayerNorm forward (bf16 out) and backward (bf16 dropout copy), max-pool and
    upsample-add twins, bf16 column sums, the AdamW weight shadow and the fused attention's bf16 output."""
    from mmfn_b200 import ops
    rnd = lambda t: t.to(torch.bfloat16).float()
    M, C = 4096, 128
    x = torch.randn(M, C, device=DEV)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    for rows in (M, 512):                                        # two-kernel and single-launch BatchNorm variants
        xx = x[:rows].contiguous().view(rows // 64, 8, 8, C)
        res = torch.randn_like(xx)
        y, mean, rstd = ops.bn_train_fwd(xx, g, b, rm.clone(), rv.clone(), res=res, relu=True, want16=True)
        y0, _, _ = ops.bn_train_fwd(xx, g, b, rm.clone(), rv.clone(), res=res, relu=True)
        assert torch.equal(y, y0) and torch.equal(ops.twin(y).float(), rnd(y)) and ops.twin(y0) is None
        dy = torch.randn_like(xx)
        dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        dz, dres = ops.bn_train_bwd(dy, xx, y, mean, rstd, g, dg, db, want_dres=True)
        dz16, dres16 = ops.bn_train_bwd(dy, xx, y, mean, rstd, g, dg.clone(), db.clone(), want_dres=True, out_bf16=True)
        assert dz16.dtype == torch.bfloat16 and torch.equal(dz16.float(), rnd(dz)) and torch.equal(dres, dres16)
    h, mean, rstd = ops.layernorm_fwd(x, g, b)
    h16, _, _ = ops.layernorm_fwd(x, g, b, out_bf16=True)
    assert torch.equal(h16.float(), rnd(h))
    dy = torch.randn(M, C, device=DEV)
    for p in (0.0, 0.1):
        dx, dxd = ops.layernorm_bwd(dy, x, g, b, mean, rstd, None, None, parts=1, drop=(p, 3))
        dx2, dxd16 = ops.layernorm_bwd(dy, x, g, b, mean, rstd, None, None, parts=1, drop=(p, 3), drop_bf16=True)
        assert torch.equal(dx, dx2) and dxd16.dtype == torch.bfloat16 and torch.equal(dxd16.float(), rnd(dxd))
    out = torch.zeros(C, device=DEV)
    ops.colsum_(h16, out)
    close(out, h16.float().sum(0), 1e-4)
    img = torch.randn(2, 32, 32, 64, device=DEV)
    mp, idx = ops.maxpool_fwd(img, want16=True)
    assert torch.equal(ops.twin(mp).float(), rnd(mp))
    tok = torch.randn(2, 192, 64, device=DEV)
    up = ops.upsample_add_fwd(img, tok, 1, want16=True)
    up0 = ops.upsample_add_fwd(img, tok, 1, want16=False)
    assert torch.equal(up, up0) and torch.equal(ops.twin(up).float(), rnd(up))
    n = 4096
    p0, gr = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    pa, pb = p0.clone(), p0.clone()
    sh = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        ma, va, sa = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(3, device=DEV)
        ops.adamw_step_(pa, gr, ma, va, sa, 1e-2)
        mb, vb, sb = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(3, device=DEV)
        ops.adamw_step_(pb, gr, mb, vb, sb, 1e-2, p16=sh)
    assert torch.equal(pa, pb) and torch.equal(sh.float(), rnd(pb))
    assert torch.equal(ops.to_bf16(p0).float(), rnd(p0))
    B, T, Cc, nh = 2, 192, 128, 4
    qkv = torch.randn(B * T, 3 * Cc, device=DEV)
    y, P, _ = ops.attention_fwd(qkv, B, T, Cc, nh)
    y16, P2, _ = ops.attention_fwd(qkv, B, T, Cc, nh, y_bf16=True)
    assert torch.equal(P, P2) and torch.equal(y16.float(), rnd(y))


@pytest.mark.parametrize("C,T,B", [(64, 192, 3), (128, 192, 3), (256, 192, 2), (512, 256, 2), (512, 256, 32), (256, 128, 2)])
def test_attention_fwd_bf16(C, T, B):
    """mmfn_attention_fwd_bf16 (K / V resident, two softmax threads per row, one exp2 per score, packed-bf16 registers)
    against fp32 torch on the same bf16-rounded q, k, v: probabilities, dropout consistency with the library's mask,
    y == Pd V on the STORED bf16 probabilities, saved row statistics, and the stats-only mode (no P store)."""
    from mmfn_b200 import ops
    nh = 4
    hs = C // nh
    qkv16 = _bf(torch.randn(B * T, 3 * C)).to(DEV)
    qkv = qkv16.float().cpu()
    heads = lambda t2d, i: t2d[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3)
    k, q, v = (heads(qkv, i) for i in range(3))
    S = q @ k.transpose(-1, -2) / hs ** 0.5
    Pr = torch.softmax(S, -1)
    y, P, Pd, stats = ops.attention_fwd_bf16(qkv16, B, T, C, nh)
    assert Pd is P and P.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
    close(P, Pr, 5e-3)
    close(P.float().sum(-1), torch.ones(B, nh, T), 5e-3)
    close(y, (Pr @ v).permute(0, 2, 1, 3).reshape(B * T, C), 1e-2)
    close(y, (P.float().cpu() @ v).permute(0, 2, 1, 3).reshape(B * T, C), 5e-3)      # exactly what the MMA consumed
    l2 = S * 1.4426950408889634
    close(stats[..., 0], l2.max(-1)[0], 1e-4)
    close(stats[..., 1], torch.exp2(l2 - l2.max(-1, keepdim=True)[0]).sum(-1), 2e-3)
    p_drop, seed = 0.1, 4321
    y2, P2, Pd2, _ = ops.attention_fwd_bf16(qkv16, B, T, C, nh, p_drop, seed)
    mask = ops.dropout(torch.ones(B, nh, T, T, device=DEV), p_drop, seed)
    assert torch.equal(P2, P)
    assert torch.count_nonzero(Pd2[mask == 0]).item() == 0             # exactly the library's mask ...
    close(Pd2, P2.float() * mask, 4e-3)                                # ... on the fp32 probabilities (one rounding each)
    close(y2, (Pd2.float().cpu() @ v).permute(0, 2, 1, 3).reshape(B * T, C), 5e-3)
    y3, P3, _, stats3 = ops.attention_fwd_bf16(qkv16, B, T, C, nh, p_drop, seed, save_probs=False)
    assert P3 is None and torch.equal(y3, y2) and torch.equal(stats3, stats)
    # backward pieces on the saved bf16 tensors: softmax backward -> bf16 dS
    dP = torch.randn(B, nh, T, T, device=DEV)
    dS = ops.softmax_bwd(P2, dP, hs ** -0.5, p_drop, seed)
    g = dP.cpu() * mask.cpu()
    Pf = P2.float().cpu()
    close(dS, hs ** -0.5 * Pf * (g - (g * Pf).sum(-1, keepdim=True)), 8e-3)


@pytest.mark.parametrize("bf16", [False, True], ids=["tf32", "bf16"])
@pytest.mark.parametrize("geom", [(4, 64, 64, 64, 3, 1, 1), (4, 64, 64, 128, 3, 2, 1), (4, 64, 64, 128, 1, 2, 0), (8, 32, 128, 128, 3, 1, 1),
                                  (16, 32, 128, 256, 3, 2, 1), (16, 16, 256, 256, 3, 1, 1), (40, 16, 256, 512, 3, 2, 1)],
                         ids=lambda g: f"N{g[0]}H{g[1]}C{g[2]}-{g[3]}R{g[4]}s{g[5]}")
def test_conv_epilogue_batchnorm_statistics(geom, bf16):
    """mmfn_conv2d_fwd_bn_*: per-channel mean / rstd / running statistics accumulated by the convolution epilogue
    (halo-reuse 3x3 kernel, generic implicit GEMM, and the split-K fallback to a separate pass) equal those of the
    stand-alone reduction over the same convolution output; bn_apply == the apply half of bn_train_fwd."""
    from mmfn_b200 import ops
    N, H, C, Co, R, stride, pad = geom
    x = torch.randn(N, H, H, C, device=DEV) * 2 + 0.5
    w = torch.randn(Co, R, R, C, device=DEV) * (2.0 / (C * R * R)) ** 0.5
    if bf16:
        x, w = x.to(torch.bfloat16), w.to(torch.bfloat16)
    g, b = torch.rand(Co, device=DEV) + 0.5, torch.randn(Co, device=DEV)
    rm0, rv0 = torch.randn(Co, device=DEV), torch.rand(Co, device=DEV) + 0.5
    ops.BF16 = bf16
    old_cap, ops.BN_FUSE_MAX_CTAS = ops.BN_FUSE_MAX_CTAS, 1 << 20          # exercise the epilogue path on every geometry
    try:
        assert ops.conv_bn_fusable(x, w, stride, pad)
        z0 = ops.conv2d_fwd(x, w, stride, pad)
        rm_a, rv_a = rm0.clone(), rv0.clone()
        res = torch.randn_like(z0)
        y0, mean0, rstd0 = ops.bn_train_fwd(z0, g, b, rm_a, rv_a, res=res, relu=True, want16=bf16)
        rm_b, rv_b = rm0.clone(), rv0.clone()
        z1, mean1, rstd1 = ops.conv2d_fwd_bn(x, w, stride, pad, rm_b, rv_b)
        y1 = ops.bn_apply(z1, g, b, mean1, rstd1, res=res, relu=True, want16=bf16)
        z2, mean2, rstd2 = ops.conv2d_fwd_bn(x, w, stride, pad, rm_b.clone(), rv_b.clone())       # scratch left zero: repeatable
        for got, ref in ((mean1, mean0), (rstd1, rstd0), (rm_b, rm_a), (rv_b, rv_a), (mean2, mean0), (rstd2, rstd0)):
            assert torch.allclose(got, ref, rtol=2e-5, atol=2e-6), (got - ref).abs().max().item()
        close(z1, z0, 1e-6 if z1.shape[1] * z1.shape[2] * N > 0 else 0)     # same kernel, atomics order only for split-K
        close(y1, y0, 2e-5)
        if bf16:
            assert ops.twin(y1).dtype == torch.bfloat16 and torch.equal(ops.twin(y1).float(), y1.to(torch.bfloat16).float())
    finally:
        ops.BF16 = False
        ops.BN_FUSE_MAX_CTAS = old_cap



@pytest.mark.parametrize("C", [3, 2])
def test_stem_conv_bf16_column_matrix(C):
    """bf16 configuration of the 7x7/2 stems: bf16 im2col matrix + kind::f16 GEMMs (forward, weight gradient) against
    torch fp32 on bf16-rounded operands."""
    from mmfn_b200 import ops
    N, H, Co = 2, 256, 64
    x = _bf(torch.randn(N, C, H, H))
    w = _bf(torch.randn(Co, C, 7, 7) * 0.1)
    xr, wr = x.float(), w.float().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=3)
    xn = x.float().permute(0, 2, 3, 1).contiguous().to(DEV)
    wk = w.float().permute(0, 2, 3, 1).contiguous().to(DEV)
    y, col, w_pad = ops.conv2d_fwd_im2col(xn, wk, 2, 3, None, bf16=True)
    assert col.dtype == torch.bfloat16 and col.shape[1] % 64 == 0
    close(y.permute(0, 3, 1, 2), yr, 3e-3)
    dy = _bf(torch.randn_like(yr))
    yr.backward(dy.float())
    dw = torch.zeros(Co, 7, 7, C, device=DEV)
    ops.conv2d_wgrad_im2col_(dy.permute(0, 2, 3, 1).contiguous().to(DEV), col, dw)
    close(dw.permute(0, 3, 1, 2), wr.grad, 3e-3)


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
@pytest.mark.parametrize("T,C", [(192, 64), (192, 128), (128, 64), (128, 128), (192, 256), (128, 256)])
@pytest.mark.parametrize("drop", [0.0, 0.1])
def test_attention_bwd_small_matches_the_formulas(T, C, drop, prec):
    """csrc/attn_bwd_small.cu (dPd, softmax backward, dQ, dK, dV of one block in one launch; bf16: one CTA per head, TF32:
    a cluster of two) against the same formulas in fp32 torch on the same operands: dV = Pd^T dY,
    dS = scale (dPd o Pd - P rowsum(dPd o Pd)), dQ = dS K, dK = dS^T Q.  Tolerance: bf16 rounding of dS and of the outputs /
    TF32 operand rounding."""
    from mmfn_b200 import ops
    B, nh = 3, 4
    hs = C // nh
    if prec == "tf32" and hs == 64:
        pytest.skip("64-dim heads: bf16 only (two fp32 half-tiles + operands exceed shared memory)")
    gen = torch.Generator(device=DEV).manual_seed(3)
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    qkv = (torch.randn(B * T, 3 * C, device=DEV, generator=gen) * 0.7).to(dt)
    dy = torch.randn(B * T, C, device=DEV, generator=gen).to(dt)
    k, q, v = (qkv[:, i * C:(i + 1) * C].float().view(B, T, nh, hs).permute(0, 2, 1, 3) for i in range(3))
    scale = hs ** -0.5
    P = torch.softmax(q @ k.transpose(-1, -2) * scale, -1).to(dt)
    if drop > 0:
        mask = (torch.rand(B, nh, T, T, device=DEV, generator=gen) >= drop).float() / (1 - drop)
        Pd = (P.float() * mask).to(dt)
    else:
        Pd = P
    got = ops.attention_bwd_small(qkv, dy, P, Pd, B, T, C, nh).float()
    dyh = dy.float().view(B, T, nh, hs).permute(0, 2, 1, 3)
    dPd = dyh @ v.transpose(-1, -2)
    w = dPd * Pd.float()
    dS = scale * (w - P.float() * w.sum(-1, keepdim=True))
    if prec == "bf16":
        dS = dS.to(dt).float()
    dq, dk, dv = dS @ k, dS.transpose(-1, -2) @ q, Pd.float().transpose(-1, -2) @ dyh
    ref = torch.cat([t.permute(0, 2, 1, 3).reshape(B * T, C) for t in (dk, dq, dv)], dim=1)
    for i, name in enumerate(("dk", "dq", "dv")):
        a, b = got[:, i * C:(i + 1) * C], ref[:, i * C:(i + 1) * C]
        err = (a - b).abs().max().item() / b.abs().max().item()
        assert err < (1.5e-2 if prec == "bf16" else 3e-3), (name, err)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("variant", ["plain", "bias_relu", "res_drop", "dgrad_mask", "out16"])
@pytest.mark.parametrize("M,N,K", [(8192, 2048, 512), (8192, 512, 2048), (4200, 1152, 320), (6144, 1024, 256)])
def test_persistent_gemm_is_bit_identical_to_the_one_tile_kernel(M, N, K, variant, prec):
    """Large plain GEMMs (>= 2 output tiles per SM) run on the persistent kernel (one CTA per SM, TMA ring across tiles,
    two TMEM accumulators, two epilogue groups).  Same MMA order per tile and the same epilogue code, so the result must
    equal the one-tile-per-CTA kernel's BIT FOR BIT -- bias / ReLU / residual / dropout / mask / bf16 output, K-major and
    MN-major B operand, ragged last row tile -- and both must match torch within the operand precision."""
    from mmfn_b200 import ops
    from mmfn_b200._lib import lib
    if prec == "tf32" and variant == "out16":
        pytest.skip("bf16 results of TF32 products are covered by test_gemm_tf32_out")
    gen = torch.Generator(device=DEV).manual_seed(1)
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    ops.set_precision(prec)
    try:
        A = torch.randn(M, K, device=DEV, generator=gen).to(dt)
        W = (torch.randn(N, K, device=DEV, generator=gen) * 0.05).to(dt)
        bias = torch.randn(N, device=DEV, generator=gen)
        res = torch.randn(M, N, device=DEV, generator=gen)
        kw, Bop = {}, W
        out_dt = torch.float32
        if variant == "bias_relu":
            kw = dict(bias=bias, act=1)
        elif variant == "res_drop":
            kw = dict(bias=bias, res=res, drop_p=0.1, seed=77)
        elif variant == "dgrad_mask":                     # dx = dy W with the ReLU mask of the producer: B is a transposed view
            Wt = (torch.randn(K, N, device=DEV, generator=gen) * 0.05).to(dt)
            Bop = Wt.t()
            kw = dict(mask=(torch.randn(M, N, device=DEV, generator=gen)).to(dt))
        elif variant == "out16":
            kw = dict(bias=bias, act=1)
            out_dt = torch.bfloat16
        outs = []
        for persist in (0, 1):
            lib().set_gemm_persist(persist)
            C = torch.full((M, N), float("nan"), device=DEV, dtype=out_dt)
            ops.gemm(A, Bop, C, **kw)
            torch.cuda.synchronize()
            outs.append(C)
        assert torch.equal(outs[0], outs[1])
        if variant in ("plain", "bias_relu", "out16"):
            ref = A.float() @ (Bop.float().t() if variant != "dgrad_mask" else Bop.float().t())
            if "bias" in kw:
                ref = torch.relu(ref + bias)
            close(outs[1], ref, 2e-2 if out_dt == torch.bfloat16 else (1e-2 if prec == "bf16" else 3e-3))
    finally:
        lib().set_gemm_persist(1)
        ops.set_precision("tf32")
