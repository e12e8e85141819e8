"""fp32 oracle for whole-model numerics checks of the fused bf16 paths.

A randomly initialised 50-layer network amplifies rounding differences between two equally valid bf16 evaluation orders by
10 % and more (tests/ROUND notes, docs/ROUND_NOTES.md "Pitfalls"), so "variant vs default" comparisons at full depth are
meaningless.  These helpers compare every path against a plain PyTorch fp32 evaluation (no TF32, no fused kernels) of the same
weights on a SHALLOW bottleneck ResNet (one block per stage: stem, projection shortcuts, strided 3x3, every fused op) and
accept a variant when its error against the oracle is of the same order as the default path's.
"""
import contextlib
import copy

import torch


def small_resnet(num_classes=64, layers=(1, 1, 1, 1), seed=0, fused_bn=None):
    from pytorch_distributed_b200.models.resnet import Bottleneck, ResNet
    torch.manual_seed(seed)
    return ResNet(Bottleneck, list(layers), num_classes=num_classes, fused_bn=fused_bn)


@contextlib.contextmanager
def model_flags(**flags):
    import pytorch_distributed_b200.models.resnet as R
    old = {k: getattr(R, k) for k in flags}
    for k, v in flags.items():
        setattr(R, k, v)
    try:
        yield
    finally:
        for k, v in old.items():
            setattr(R, k, v)


def step(model, x, y):
    out = model(x)
    torch.nn.functional.cross_entropy(out.float(), y).backward()
    return out.float().detach(), {n: p.grad.float().detach().clone() for n, p in model.named_parameters()}


def fp32_oracle(base, x, y):
    """Plain PyTorch fp32 (BatchNorm via F.batch_norm, cuDNN without TF32) over the bf16-rounded weights of ``base``."""
    m = copy.deepcopy(base).float()
    for mod in m.modules():
        if hasattr(mod, "fused"):
            mod.fused = False
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with model_flags(FUSED_CONV1X1=False, SPLIT_RESGRAD=False, STEM_GEMM=False):
            return step(m.train(), x.float(), y)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rel_err(a, ref):
    return ((a - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def compare(variant, default, oracle, factor=2.5, floor=1e-2):
    """variant / default / oracle = (out, grads).  Returns a list of violations (empty = pass)."""
    bad = []
    ev, ed = rel_err(variant[0], oracle[0]), rel_err(default[0], oracle[0])
    if ev > max(factor * ed, floor):
        bad.append(("output", ev, ed))
    for n in oracle[1]:
        ev, ed = rel_err(variant[1][n], oracle[1][n]), rel_err(default[1][n], oracle[1][n])
        if ev > max(factor * ed, floor):
            bad.append((n, ev, ed))
    return bad
