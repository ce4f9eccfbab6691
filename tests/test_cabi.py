"""CPU: the C-ABI library loads and exports exactly what include/mmfn_b200.h declares; the header
is in sync with the sources; the product package never imports the oracle."""
import ctypes
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_in_sync_with_sources():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_header.py"), "--check"]) == 0


def test_library_exports_every_declared_symbol():
    from mmfn_b200._lib import LIB_PATH, parse_header
    protos = parse_header()
    assert len(protos) >= 45
    dll = ctypes.CDLL(LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), name
    dll.mmfn_version.restype = ctypes.c_int
    assert dll.mmfn_version() >= 100
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mmfn_\w+)", out))
    assert exported == set(protos), exported ^ set(protos)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any launch (so this runs on a CPU-only box)."""
    from mmfn_b200._lib import lib, MmfnError
    import pytest
    with pytest.raises(MmfnError, match="pt_stride"):
        lib().bev_scatter(16, 1, 10, 2, 16, 0, None)
    with pytest.raises(MmfnError, match="null"):
        lib().gemm_f32(0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 4, 4, 4, 1, 1, 0, 0, 0, 1.0, 0, 0, 0.0, 0, 1, None)
    # entry points added for the tensor-core data gradients, the stems, the fused attention backward and the bucketed
    # optimizer: shape / alignment violations are reported before any launch (fake 16-byte aligned pointers)
    P = 4096
    with pytest.raises(MmfnError, match="stride must be 1 or 2"):
        lib().conv2d_dgrad_tf32(P, P, P, 0, 2, 32, 32, 64, 64, 3, 3, 3, 1, 11, 11, None)
    with pytest.raises(MmfnError, match="H, W must be even"):
        lib().conv2d_dgrad_tf32(P, P, P, 0, 2, 31, 31, 64, 64, 3, 3, 2, 1, 16, 16, None)
    with pytest.raises(MmfnError, match="Kp must be a multiple of 4"):
        lib().im2col_nhwc(P, P, 1, 8, 8, 3, 7, 7, 2, 3, 4, 4, 150, None)
    with pytest.raises(MmfnError, match="multiple of 32, <= 256"):
        lib().attention_bwd_dq_tf32(P, P, P, P, P, P, 1, 100, 64, 4, 0.0, 0, None)
    with pytest.raises(MmfnError, match="head size"):
        lib().attention_bwd_dq_tf32(P, P, P, P, P, P, 1, 192, 96, 4, 0.0, 0, None)
    with pytest.raises(MmfnError, match="multiple of 4"):
        lib().adamw_apply(P, P, P, P, 10, 1e-4, 0.9, 0.999, 1e-8, 0.01, P, 1.0, 0, None)
    # bf16 tensor-core entry points (BASELINE configs[2]) and the size query
    with pytest.raises(MmfnError, match="multiples of 16 bytes"):
        lib().gemm_bf16(P, 12, 0, 0, 0, P, 64, 0, 0, 0, P, 0, 64, 0, 0, 128, 64, 64, 1, 1, 0, 0, 0, 1.0, 0, 0, 0.0, 0, 1, None)
    with pytest.raises(MmfnError, match="cannot be accumulated"):
        lib().gemm_bf16(P, 64, 0, 0, 0, P, 64, 0, 0, 0, P, 1, 64, 0, 0, 128, 64, 64, 1, 1, 0, 0, 0, 1.0, 0, 2, 0.0, 0, 1, None)
    with pytest.raises(MmfnError, match="multiple of 64"):
        lib().conv2d_fwd_bf16(P, P, P, 0, 2, 32, 32, 32, 64, 3, 3, 1, 1, 32, 32, None)
    with pytest.raises(MmfnError, match="Co % 64"):
        lib().conv2d_dgrad_bf16(P, P, P, 0, 2, 32, 32, 64, 96, 3, 3, 1, 1, 32, 32, None)
    with pytest.raises(MmfnError, match="unknown op"):
        lib().workspace_bytes(99, 0, 1, 1, 1, 1, P, )
    with pytest.raises(MmfnError, match="multiple of 4"):
        lib().f32_to_bf16(P, P, 6, None)
    with pytest.raises(MmfnError, match="bad args"):
        lib().copy2d_f32(P, 4, P, 8, 2, 6, 0, None)
    # one-launch polyline sub-graph: instantiated for 9 / 19 vectors per polyline only
    with pytest.raises(MmfnError, match="V must be 9 or 19"):
        lib().subgraph_fused_fwd(P, 4, 12, 0, *([P] * 12), *([P] * 17), 1e-5, None)
    with pytest.raises(MmfnError, match="null output"):
        lib().subgraph_fused_fwd(P, 4, 9, 0, *([P] * 12), 0, *([P] * 16), 1e-5, None)
    # whole-GPT kernels and the one-launch attention backward: supported geometries only
    with pytest.raises(MmfnError, match="needs C in"):
        lib().gpt_small_fwd(P, 2, 192, 256, 4, 8, 2, P, *([P] * 13), 0.0, 0.0, 0, 1e-5, None)
    with pytest.raises(MmfnError, match="multiple of 64"):
        lib().gpt_small_bwd_rows(100, 64, 2, 1, 1, *([P] * 9), *([P] * 13), 0.0, 0, 0, None)
    with pytest.raises(MmfnError, match="head size 16, 32 or 64"):
        lib().attention_bwd_small_bf16(P, P, P, P, P, 2, 192, 512, 4, None)
    with pytest.raises(MmfnError, match="head size 16 or 32"):
        lib().attention_bwd_small_tf32(P, P, P, P, P, 2, 192, 256, 4, None)
    # stem kernels, BatchNorm mask modes, loader kernels, skinny weight gradient: argument checks precede every launch
    with pytest.raises(MmfnError, match="C must be 2 or 3"):
        lib().conv2d_stem7_fwd(P, P, P, 2, 256, 256, 4, None)
    with pytest.raises(MmfnError, match="dz alignment"):
        lib().conv2d_stem7_wgrad(P + 8, 0, P, P, 2, 256, 256, 3, None)
    with pytest.raises(MmfnError, match="null pointer"):
        lib().stem_bn_relu_maxpool_fwd(P, 2, 128, 128, 64, P, P, 0, 0, 0.1, 1e-5, P, P, P, 0, P, 0, P, None)      # zmax missing
    with pytest.raises(MmfnError, match="C % 4 == 0"):
        lib().stem_bn_relu_maxpool_bwd(P, P, P, P, P, P, P, 2, 128, 128, 62, P, 0, P, P, P, None)
    with pytest.raises(MmfnError, match="relu_from_z needs beta and no yout"):
        lib().bn_train_bwd(P, P, P, 0, 1, P, P, P, P, 4096, 64, P, 0, 0, P, P, P, None)
    with pytest.raises(MmfnError, match="null pointer"):
        lib().bn_apply(P, 0, 4096, 64, P, P, P, P, 0, 1, 0, None)                      # neither y nor its bf16 twin
    with pytest.raises(MmfnError, match="bad sizes"):
        lib().lidar_ego_transform_f64(P, 2, P, P, 1, 1024, None)
    with pytest.raises(MmfnError, match="n % 4 == 0"):
        lib().bev_unpack_u8(P, P, 10, None)
    with pytest.raises(MmfnError, match="bad args"):
        lib().radar_adjacency_f64(P, P, 0, 81, None)
    with pytest.raises(MmfnError, match="16-byte aligned"):
        lib().wgrad_n64_k7(P + 4, P, P, 4096, None)
    # host-side setting: returns the previous threshold
    import ctypes
    old = ctypes.c_int(-1)
    lib().set_bn_small_rows(2048, ctypes.addressof(old))
    assert old.value == 1024
    lib().set_bn_small_rows(old.value, 0)
    with pytest.raises(MmfnError, match="rows must be >= 0"):
        lib().set_bn_small_rows(-1, 0)


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "mmfn_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), path


def test_sass_is_sm100a():
    from mmfn_b200._lib import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:200]


def test_param_spec_matches_reference_state_dict():
    import json
    import torch
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import ParamStore, param_spec, is_unused
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    spec = param_spec(GlobalConfig())
    assert [k for k, _, _ in spec] == list(gold.keys())
    for k, shape, kind in spec:
        assert list(shape) == gold[k][0], k
        assert gold[k][1] == ("torch.int64" if kind == "nbt" else "torch.float32")
    st = ParamStore(GlobalConfig(), "cpu")
    root = torch.nn.Module()
    st.register(root)
    sd = root.state_dict()
    assert list(sd.keys()) == list(gold.keys())
    assert sum(p.numel() for p in root.parameters()) == 104939738
    assert sum(1 for k, _ in root.named_parameters() if is_unused(k)) == 21
    # conv filters are stored KRSC but presented with the reference (K,C,R,S) shape
    w = dict(root.named_parameters())["encoder.image_encoder.features.layer1.0.conv1.weight"]
    assert tuple(w.shape) == (64, 64, 3, 3) and w.stride() == (576, 1, 192, 64)
    # load_state_dict round trip keeps the aliasing with the flat buffer
    from mmfn_b200 import synthetic
    root.load_state_dict(synthetic.fill_golden_weights(sd, 1))
    k = "encoder.transformer4.blocks.7.mlp.2.bias"
    assert torch.equal(st.p(k), synthetic.fill_golden_weights({k: sd[k]}, 1)[k])


def test_flat_layout_buckets_are_contiguous():
    """[late | early | never-used]: the early gradient bucket (head, transformer4, radar GAT, layer4 of every trunk)
    is ONE contiguous 16-byte-aligned range, fused q/k/v stay adjacent, never-used parameters sit behind n_active."""
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import ParamStore, is_early_bucket, is_mid_bucket, is_unused
    for variant in ("rad", "vec", "img", "transfuser"):
        st = ParamStore(GlobalConfig(), "cpu", variant)
        assert 0 < st.n_late < st.n_mid < st.n_active <= st.n_total
        assert st.n_late % 4 == 0 and st.n_mid % 4 == 0 and st.n_active % 4 == 0
        n_early = 0
        for k, off in st.offsets.items():
            n = 1
            for d in st.shapes[k]:
                n *= d
            assert off % 4 == 0, k
            if is_unused(k, variant):
                assert off >= st.n_active, k
            elif is_early_bucket(k):
                assert st.n_mid <= off and off + n <= st.n_active, k
                n_early += n
            elif is_mid_bucket(k):
                assert st.n_late <= off and off + n <= st.n_mid, k
            else:
                assert off + n <= st.n_late, k
        assert n_early > 0.4 * st.n_active            # the first overlap-able bucket carries > 40 % of the gradient bytes
        assert st.n_late < 0.35 * st.n_active         # ... and less than 35 % is left for the exposed exchange after backward
        qkv = st.fused([f"encoder.transformer1.blocks.0.attn.{n}.weight" for n in ("key", "query", "value")])
        assert tuple(qkv.shape) == (192, 64)
