"""Split residual gradients (models/resnet.py: SPLIT_RESGRAD, ops/bn_act.py: split / bn_act_backward2) and the PyTorch
emulation of the BN kernels' ABI (ops/bn_act.py: _Emu), which drives the real autograd op on CPU."""
import copy
import itertools

import pytest
import torch
import torch.nn as nn

import pytorch_distributed_b200.models.resnet as R
from pytorch_distributed_b200.models import create_model
from pytorch_distributed_b200.ops.bn_act import bn_act, bn_act_reference


@pytest.mark.parametrize("relu,res,split,dt", list(itertools.product([True, False], [True, False], [0, 1, 2],
                                                                      [torch.float32, torch.bfloat16])))
def test_emulated_op_matches_composition(relu, res, split, dt):
    """split: 0 = one output, 1 = both aliases consumed (two gradients -> bn_act_backward2), 2 = only the second alias."""
    torch.manual_seed(0)
    N, C, H, W = 4, 16, 5, 6
    x0 = torch.randn(N, C, H, W).to(dt).contiguous(memory_format=torch.channels_last)
    r0 = torch.randn(N, C, H, W).to(dt).contiguous(memory_format=torch.channels_last)
    w0, b0 = torch.rand(C) + 0.5, torch.randn(C)
    outs = []
    for mode in ("emulate", "ref"):
        x = x0.clone().requires_grad_()
        r = r0.clone().requires_grad_() if res else None
        w, b = w0.clone().requires_grad_(), b0.clone().requires_grad_()
        rm, rv, nbt = torch.zeros(C), torch.ones(C), torch.zeros((), dtype=torch.long)
        if mode == "emulate":
            y = bn_act(x, w, b, rm, rv, r, relu, True, 0.1, 1e-5, fused="emulate", num_batches_tracked=nbt, split=split > 0)
            if split:
                assert y[0].data_ptr() == y[1].data_ptr() and y[0].grad_fn is y[1].grad_fn
        else:
            y = bn_act_reference(x.float(), w, b, rm, rv, None if r is None else r.float(), relu, True, 0.1, 1e-5).to(dt)
            nbt += 1
            if split:
                y = (y, y)
        if split == 0:
            ya = y
            loss = (ya.float() * 1.5).sin().sum()
        elif split == 1:
            ya, yb = y
            loss = (ya.float() * 1.5).sin().sum() + (yb.float() * 0.7 + 0.3).cos().sum()
        else:
            ya, yb = y
            loss = (yb.float() * 0.7 + 0.3).cos().sum()
        loss.backward()
        outs.append((ya.detach().float(), x.grad.float(), None if r is None else r.grad.float(), w.grad, b.grad, rm, rv, nbt))
    tol = 1e-4 if dt == torch.float32 else 6e-2
    for i, (a, e) in enumerate(zip(*outs)):
        if a is None:
            continue
        a, e = a.float(), e.float()
        assert (a - e).abs().max().item() / (e.abs().max().item() + 1e-6) < tol, i


def _stack(fused):
    torch.manual_seed(1)
    blocks = [R.Bottleneck(32, 16, downsample=R._Downsample(32, 64, 1, fused), fused=fused), R.Bottleneck(64, 16, fused=fused),
              R.Bottleneck(64, 16, stride=2, downsample=R._Downsample(64, 64, 2, fused), fused=fused), R.Bottleneck(64, 16, fused=fused)]
    return nn.Sequential(*blocks).to(memory_format=torch.channels_last)


def _run(net, x, split):
    R.SPLIT_RESGRAD = split
    try:
        net.train()
        xin = x.clone().requires_grad_()
        out = R._pair(net(xin))[0]
        (out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum().backward()
        return out.detach(), xin.grad, {n: p.grad.clone() for n, p in net.named_parameters()}
    finally:
        R.SPLIT_RESGRAD = False


def test_split_gradients_through_bottleneck_stack():
    """Four bottlenecks (two with projection shortcuts): reference composition vs the emulated fused op, with and without the
    split; every block boundary except the last then reaches bn_act_backward2 with two gradients."""
    x = torch.randn(6, 32, 12, 12).contiguous(memory_format=torch.channels_last)
    ref = _run(_stack(False), x, False)
    calls = {"two": 0}
    from pytorch_distributed_b200.ops import bn_act as B
    orig = B._Emu.bn_act_backward2

    def counted(*a, **k):
        calls["two"] += 1
        return orig(*a, **k)

    B._Emu.bn_act_backward2 = staticmethod(counted)
    try:
        emu = _run(_stack("emulate"), x, False)
        assert calls["two"] == 0
        spl = _run(_stack("emulate"), x, True)
        assert calls["two"] == 3                    # blocks 0, 1, 2 feed two consumers; block 3's second alias is unused
    finally:
        B._Emu.bn_act_backward2 = staticmethod(orig)
    for got in (emu, spl):
        assert torch.allclose(got[0], ref[0], rtol=1e-4, atol=1e-5)
        assert torch.allclose(got[1], ref[1], rtol=1e-3, atol=1e-5)
        for n in ref[2]:
            scale = ref[2][n].abs().max().item() + 1e-8
            assert (got[2][n] - ref[2][n]).abs().max().item() / scale < 2e-3, n


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_full_model_forward_is_unchanged_by_split(arch):
    """Whole network with the emulated kernels: identical forward, gradients equal up to the fp32 conditioning of a
    randomly initialised ResNet (both agree with an fp64 oracle only to a few percent)."""
    torch.manual_seed(0)
    base = create_model(arch, num_classes=10, fused_bn=False).float().to(memory_format=torch.channels_last)
    for mod in base.modules():
        if isinstance(mod, R.BNAct):
            mod.fused = "emulate"
    x = torch.randn(8, 3, 64, 64).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,))
    res = []
    for split in (False, True):
        m = copy.deepcopy(base).train()
        R.SPLIT_RESGRAD = split
        try:
            out = m(x)
            torch.nn.functional.cross_entropy(out, y).backward()
        finally:
            R.SPLIT_RESGRAD = False
        res.append((out.detach(), {n: p.grad.clone() for n, p in m.named_parameters()}, dict(m.named_buffers())))
    assert torch.equal(res[0][0], res[1][0])
    for n, b in res[0][2].items():
        assert torch.equal(b, res[1][2][n]), n
    for n, g in res[0][1].items():
        assert (g - res[1][1][n]).abs().max().item() / (g.abs().max().item() + 1e-8) < 0.1, n
