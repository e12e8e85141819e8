"""``bn_act``: BatchNorm2d (+ residual add) (+ ReLU) as one autograd op.

CUDA + channels_last + C % 8 == 0  -> hand-written NHWC kernels (``csrc/bn_act.cu``): statistics pass + one fused
apply pass forward, reduce pass + one fused apply pass backward.
Anything else (CPU, NCHW, odd channel counts) -> the plain PyTorch composition below, which is also the numerical
oracle in ``tests/test_bn_act.py``.

The per-call fp32 accumulators ([2C] sums for forward, [2C] for backward) come from a per-device workspace that is
zeroed ONCE per training step (one memset for all ~100 slices of a ResNet-50) instead of one ``zeros()`` per layer.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def bn_act_reference(x, weight, bias, running_mean, running_var, residual=None, relu=True, training=True, momentum=0.1,
                     eps=1e-5):
    """Plain PyTorch semantics: relu(batch_norm(x) + residual)."""
    w = weight if weight is None or weight.dtype == x.dtype or x.dtype == torch.float32 else weight
    y = F.batch_norm(x, running_mean, running_var, w, bias, training, momentum, eps)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


class _Workspace:
    """Bump allocator over one zeroed fp32 buffer per device; ``reset`` = a single memset per step."""

    def __init__(self, device, capacity: int = 1 << 20):
        self.buf = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.used = 0
        self.generation = 0

    def reset(self):
        if self.used:
            self.buf[: self.used].zero_()
        self.used = 0
        self.generation += 1

    def take(self, n: int):
        n = (n + 31) // 32 * 32
        if self.used + n > self.buf.numel():
            return torch.zeros(n, dtype=torch.float32, device=self.buf.device), -1
        s = self.buf[self.used: self.used + n]
        self.used += n
        return s, self.generation


_workspaces = {}


def workspace(device) -> _Workspace:
    ws = _workspaces.get(device)
    if ws is None:
        ws = _Workspace(device)
        _workspaces[device] = ws
    return ws


def begin_step(device) -> None:
    """Call once before a training forward pass: recycles (zeroes) the accumulator slices of the previous step."""
    if device.type == "cuda":
        workspace(device).reset()


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, nbt, training, momentum, eps, relu, need_grad, pre=None):
        from .. import _ext
        C = _ext.lib()
        nc = x.size(1)
        ws = workspace(x.device)
        stats_ready = pre is not None          # (work[4C], generation): sums already reduced by the producing GEMM
        if stats_ready:
            work, gen = pre
        elif training:
            work, gen = ws.take(4 * nc)
        else:
            work, gen = torch.empty(0, dtype=torch.float32, device=x.device), -1
        _ext.note_launch(1 if (stats_ready or not training) else 2)
        y, saved, mask = C.bn_act_forward(x, residual, weight, bias, running_mean, running_var, nbt, training, momentum, eps, relu,
                                          need_grad, work[: 2 * nc] if training else work, stats_ready)
        ctx.relu = relu
        ctx.has_res = residual is not None
        ctx.work = work[2 * nc:] if training else None
        ctx.gen = gen
        ctx.ws = ws
        if need_grad:
            if not training:
                raise RuntimeError("fused bn_act: backward through eval-mode batch norm is not supported")
            ctx.save_for_backward(x, mask if relu else None, weight, saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import _ext
        C = _ext.lib()
        x, mask, weight, saved = ctx.saved_tensors
        work = ctx.work
        if work is None or (ctx.gen != -1 and ctx.gen != ctx.ws.generation):
            work = torch.zeros(2 * x.size(1), dtype=torch.float32, device=x.device)   # slice was recycled: use a fresh one
        _ext.note_launch(2)
        dx, dres, dw, db = C.bn_act_backward(dy, x, mask, weight, saved, ctx.relu, ctx.has_res, work)
        return dx, (dres if ctx.has_res else None), dw, db, None, None, None, None, None, None, None, None, None


def _can_fuse(x, weight, residual, running_mean=True) -> bool:
    return (x.is_cuda and x.dim() == 4 and x.size(1) % 8 == 0 and x.size(1) <= 8192 and weight is not None and running_mean is not None
            and x.dtype in (torch.float32, torch.bfloat16, torch.float16)
            and x.is_contiguous(memory_format=torch.channels_last)
            and (residual is None or (residual.is_contiguous(memory_format=torch.channels_last) and residual.dtype == x.dtype
                                      and residual.shape == x.shape)))


def bn_act(x, weight, bias, running_mean, running_var, residual: Optional[torch.Tensor] = None, relu: bool = True,
           training: bool = True, momentum: float = 0.1, eps: float = 1e-5, fused: Optional[bool] = None,
           num_batches_tracked: Optional[torch.Tensor] = None):
    """relu(batch_norm(x) + residual).  ``fused=None`` picks the CUDA kernels whenever the layout allows it."""
    ok = _can_fuse(x, weight, residual, running_mean)
    use = ok if fused is None else (fused and ok)
    if not use:
        if weight is not None and x.is_cuda and weight.dtype != torch.float32 and x.dtype != weight.dtype:
            weight, bias = weight.to(x.dtype), bias.to(x.dtype)
        if training and num_batches_tracked is not None:
            num_batches_tracked.add_(1)
        if not x.is_cuda and x.dtype != torch.float32:
            # CPU batch_norm wants one dtype for activations and statistics: do the math in fp32 (test / debug path)
            y = bn_act_reference(x.float(), None if weight is None else weight.float(), None if bias is None else bias.float(),
                                 running_mean, running_var, None if residual is None else residual.float(), relu, training,
                                 momentum, eps)
            return y.to(x.dtype)
        return bn_act_reference(x, weight, bias, running_mean, running_var, residual, relu, training, momentum, eps)
    need_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or bias.requires_grad
                                             or (residual is not None and residual.requires_grad))
    if need_grad and not training:      # backward through frozen (eval-mode) statistics: rare, use the composition
        return bn_act_reference(x, weight.to(x.dtype) if weight.dtype != torch.float32 else weight,
                                bias.to(x.dtype) if bias.dtype != torch.float32 else bias, running_mean, running_var, residual, relu,
                                training, momentum, eps)
    return _BnActFn.apply(x, residual, weight, bias, running_mean, running_var, num_batches_tracked, training, float(momentum),
                          float(eps), relu, need_grad)
