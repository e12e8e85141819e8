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


def find_pretrained_file(arch: str):
    """Local torchvision-format weights for ``arch``: $PTD_PRETRAINED (a file), $PTD_PRETRAINED_DIR/<arch>*.pth, or torch
    hub's cache ($TORCH_HOME/hub/checkpoints/<arch>-<hash>.pth - where ``models.resnet50(pretrained=True)`` would have put it)."""
    import glob
    import os
    f = os.environ.get("PTD_PRETRAINED", "")
    if f and os.path.isfile(f):
        return f
    dirs = [os.environ.get("PTD_PRETRAINED_DIR", "")]
    try:
        import torch.hub
        dirs.append(os.path.join(torch.hub.get_dir(), "checkpoints"))
    except Exception:  # noqa: BLE001
        pass
    for d in dirs:
        if d and os.path.isdir(d):
            hits = sorted(glob.glob(os.path.join(d, arch + "-*.pth")) + glob.glob(os.path.join(d, arch + ".pth")) +
                          glob.glob(os.path.join(d, arch + "*.pth.tar")))
            if hits:
                return hits[0]
    return None


def load_pretrained(model, arch: str):
    """``--pretrained`` (/root/reference/distributed.py:134-136) for the native ResNets: the parameter / buffer names equal
    torchvision's, so its published state dicts load directly.  Looks for a local file first, then lets torchvision
    download (which needs network access)."""
    import torch
    path = find_pretrained_file(arch)
    if path is not None:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        if isinstance(sd, dict) and "state_dict" in sd:       # a checkpoint.pth.tar of this framework / the reference
            sd = sd["state_dict"]
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
    else:
        try:
            import torchvision.models as tvm
            sd = tvm.get_model_weights(arch).DEFAULT.get_state_dict(progress=False)
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("--pretrained: no local weights for '%s' (set PTD_PRETRAINED=<file> or PTD_PRETRAINED_DIR=<dir with "
                               "%s-*.pth>, or pre-populate torch hub's cache) and the download failed: %s" % (arch, arch, e)) from e
    own = model.state_dict()
    if "fc.weight" in sd and "fc.weight" in own and sd["fc.weight"].shape != own["fc.weight"].shape:
        sd = {k: v for k, v in sd.items() if not k.startswith("fc.")}      # different class count: keep the trunk only
        print("=> pretrained classifier dropped (num_classes differs)")
    missing, unexpected = model.load_state_dict(sd, strict=False)
    bad = [k for k in missing if not k.startswith("fc.")]
    if bad or unexpected:
        raise RuntimeError("--pretrained: state dict does not match '%s' (missing %s, unexpected %s)" % (arch, bad[:5], list(unexpected)[:5]))
    return model


def create_model(arch: str, pretrained: bool = False, num_classes: int = 1000, fused_bn: bool | None = None,
                 native: bool = True):
    """Build ``arch``.  ``fused_bn=None`` => fused kernels whenever they can run (CUDA, NHWC)."""
    if native and arch in _NATIVE:
        model = _NATIVE[arch](num_classes=num_classes, fused_bn=fused_bn)
        if pretrained:
            print("=> using pre-trained model '{}'".format(arch))
            return load_pretrained(model, arch)
        print("=> creating model '{}'".format(arch))
        return model
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
