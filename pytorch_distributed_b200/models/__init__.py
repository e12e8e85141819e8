"""Model zoo front-end.

The reference exposes every lowercase callable of ``torchvision.models`` as an
``--arch`` choice (/root/reference/distributed.py:21-23,134-139).  We keep that
surface, but the ResNet family (the benchmarked models) is implemented natively
in :mod:`.resnet` with NHWC fused BN(+add)+ReLU blocks; every other name is
served by torchvision.
"""
from __future__ import annotations

from . import resnet as _resnet

_NATIVE = dict(_resnet.FACTORIES)


def _torchvision_names():
    try:
        import torchvision.models as tvm
    except Exception:  # pragma: no cover - torchvision is in the image
        return []
    return sorted(n for n, v in tvm.__dict__.items()
                  if n.islower() and not n.startswith("__") and callable(v))


def available_models():
    return sorted(set(_torchvision_names()) | set(_NATIVE))


def create_model(arch: str, pretrained: bool = False, num_classes: int = 1000, fused_bn: bool | None = None,
                 native: bool = True):
    """Build ``arch``.  ``fused_bn=None`` => fused kernels whenever they can run (CUDA, NHWC)."""
    if native and arch in _NATIVE:
        if pretrained:
            raise RuntimeError("--pretrained needs network access to download weights; load a local "
                               "checkpoint with --resume instead")
        print("=> creating model '{}'".format(arch))
        return _NATIVE[arch](num_classes=num_classes, fused_bn=fused_bn)
    import torchvision.models as tvm
    if pretrained:
        print("=> using pre-trained model '{}'".format(arch))
        return tvm.__dict__[arch](pretrained=True)
    print("=> creating model '{}'".format(arch))
    kwargs = {} if num_classes == 1000 else {"num_classes": num_classes}
    model = tvm.__dict__[arch](**kwargs)
    if fused_bn:                       # explicit --fused-bn: rewrite Sequential BN -> ReLU pairs of zoo models (models/surgery.py)
        from .surgery import fuse_bn_relu
        n = fuse_bn_relu(model)
        print("=> fused %d BatchNorm+ReLU pairs of '%s'" % (n, arch))
    return model
