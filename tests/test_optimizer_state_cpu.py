"""CPU: the flat AdamW state of TrainEngine converts to / from the state_dict of a real torch.optim.AdamW over the
reference-ordered parameters (what phase2_train_net.py:213,:219 saves as best_optim.pth and :297-302 loads), whatever
the internal flat layout (late / early / never-used groups) is."""
import types

import torch

from mmfn_b200.config import GlobalConfig
from mmfn_b200.params import ParamStore, is_unused
from mmfn_b200.trainer import load_optimizer_state_dict, optimizer_state_dict


def _stub_engine(step):
    st = ParamStore(GlobalConfig(), "cpu")
    root = torch.nn.Module()
    st.register(root)
    g = torch.Generator().manual_seed(1)
    model = types.SimpleNamespace(_param_items=list(root.named_parameters()), VARIANT="rad")
    b1, b2 = 0.9, 0.999
    eng = types.SimpleNamespace(st=st, model=model, m=torch.randn(st.n_active, generator=g),
                                v=torch.rand(st.n_active, generator=g), lr=1e-4, betas=(b1, b2), eps=1e-8, wd=0.01,
                                state=torch.tensor([float(step), 1 - b1 ** step, 1 - b2 ** step]))
    return eng, root


def test_state_dict_loads_into_torch_adamw_and_round_trips():
    eng, root = _stub_engine(step=3)
    sd = optimizer_state_dict(eng)
    params = list(root.parameters())
    names = [k for k, _ in root.named_parameters()]
    assert sd["param_groups"][0]["params"] == list(range(len(params)))
    opt = torch.optim.AdamW(params, lr=123.0)
    opt.load_state_dict(sd)                                   # torch validates group sizes / indices here
    assert opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[0]["weight_decay"] == 0.01
    n_unused = 0
    for k, p in zip(names, params):
        if is_unused(k):
            assert p not in opt.state                          # never received a gradient: torch keeps no state
            n_unused += 1
            continue
        s = opt.state[p]
        assert float(s["step"]) == 3.0 and s["exp_avg"].shape == p.shape and s["exp_avg_sq"].shape == p.shape
    assert n_unused == 21
    # a conv filter's moments come back in the reference (K, C, R, S) shape and map onto the KRSC flat storage
    k = "encoder.image_encoder.features.layer2.0.conv1.weight"
    p = dict(root.named_parameters())[k]
    off = eng.st.offsets[k]
    want = eng.m[off: off + p.numel()].view(128, 3, 3, 64).permute(0, 3, 1, 2)
    assert torch.equal(opt.state[p]["exp_avg"], want)
    # and back: a fresh engine restored from torch's own state_dict holds the same flat moments and step state
    eng2, _ = _stub_engine(step=0)
    eng2.m.zero_(); eng2.v.zero_()
    load_optimizer_state_dict(eng2, opt.state_dict())
    live = torch.zeros(eng.st.n_active, dtype=torch.bool)
    for k2 in names:
        if not is_unused(k2):
            o = eng.st.offsets[k2]
            n = 1
            for d in eng.st.shapes[k2]:
                n *= d
            live[o: o + n] = True
    assert torch.equal(eng2.m[live], eng.m[live]) and torch.equal(eng2.v[live], eng.v[live])
    assert float(eng2.state[0]) == 3.0
    assert torch.allclose(eng2.state[1:], eng.state[1:], rtol=1e-4)      # loader widens fp32 betas like the device kernel


def test_empty_state_before_the_first_step():
    eng, root = _stub_engine(step=0)
    sd = optimizer_state_dict(eng)
    assert sd["state"] == {}
    torch.optim.AdamW(list(root.parameters()), lr=1e-4).load_state_dict(sd)
