"""Module surgery: give *any* model of the zoo the fused NHWC BatchNorm(+ReLU) kernels.

The reference accepts every torchvision architecture through ``--arch`` (/root/reference/distributed.py:21-23,136-139).
Only the ResNet family is re-implemented natively (``models/resnet.py``); for the rest, ``fuse_bn_relu`` rewrites the
``BatchNorm2d -> ReLU`` pairs that sit next to each other inside ``nn.Sequential`` containers (VGG-BN ``features``,
torchvision's ``Conv2dNormActivation`` blocks used by RegNet / ShuffleNet / GoogLeNet-style stems, ...) into one
:class:`~pytorch_distributed_b200.models.resnet.BNAct` (which runs ``csrc/bn_act.cu`` on CUDA + channels_last and the plain
PyTorch composition everywhere else) followed by ``nn.Identity``.  Parameter / buffer names - and therefore checkpoints -
are unchanged.  Only ``nn.Sequential`` parents are touched: there the "BN output feeds the ReLU and nothing else"
property is structural, not an assumption about somebody's ``forward``.
"""
from __future__ import annotations

import torch.nn as nn

from .resnet import BNAct


def _to_bnact(bn: nn.BatchNorm2d, relu: bool) -> BNAct:
    new = BNAct(bn.num_features, relu=relu, eps=bn.eps, momentum=bn.momentum, affine=bn.affine,
                track_running_stats=bn.track_running_stats)
    new.training = bn.training
    if bn.affine:                      # share the very same Parameter objects (optimizers / DDP hooks stay valid)
        new.weight, new.bias = bn.weight, bn.bias
    if bn.track_running_stats:
        new._buffers["running_mean"] = bn.running_mean
        new._buffers["running_var"] = bn.running_var
        new._buffers["num_batches_tracked"] = bn.num_batches_tracked
    return new


def fuse_bn_relu(module: nn.Module) -> int:
    """In-place rewrite; returns the number of fused BatchNorm2d -> ReLU pairs."""
    fused = 0
    for child in module.children():
        fused += fuse_bn_relu(child)
    if isinstance(module, nn.Sequential):
        names = list(module._modules.keys())
        for a, b in zip(names, names[1:]):
            bn, act = module._modules[a], module._modules[b]
            if type(bn) is nn.BatchNorm2d and type(act) is nn.ReLU and bn.affine:
                module._modules[a] = _to_bnact(bn, relu=True)
                module._modules[b] = nn.Identity()
                fused += 1
    return fused
