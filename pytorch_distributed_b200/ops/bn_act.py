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
    if device.type == "cuda" or device in _workspaces:
        workspace(device).reset()


class _Emu:
    """Pure PyTorch (fp32 math) stand-in for the ``csrc/bn_act.cu`` entry points, same signatures and tensor contracts
    (channels_last activations seen as a row-major [M, C] matrix, 1 mask byte per 8 channels, ``saved`` = mean | invstd).
    It documents what the kernels compute and lets the CPU tests drive the autograd plumbing of :class:`_BnActFn`
    (``bn_act(..., fused="emulate")``)."""

    @staticmethod
    def _rows(t):
        return t.permute(0, 2, 3, 1).reshape(-1, t.size(1))

    @staticmethod
    def _like(rows, ref):
        n, c, h, w = ref.shape
        return rows.reshape(n, h, w, c).permute(0, 3, 1, 2).to(ref.dtype).contiguous(memory_format=torch.channels_last)

    @staticmethod
    def _unpack(mask, m, c):
        bits = (mask.view(m, c // 8, 1).to(torch.int32) >> torch.arange(8, device=mask.device, dtype=torch.int32)) & 1
        return bits.reshape(m, c).bool()

    @staticmethod
    def bn_act_forward(x, residual, weight, bias, rm, rv, nbt, training, momentum, eps, relu, need_mask, work, stats_ready):
        xr = _Emu._rows(x).float()
        m, c = xr.shape
        saved = mask = None
        if training:
            if stats_ready:
                mean = work[:c] / m
                var = (work[c:2 * c] / m - mean * mean).clamp_min(0)
            else:
                mean, var = xr.mean(0), xr.var(0, unbiased=False)
            invstd = torch.rsqrt(var + eps)
            saved = torch.cat([mean, invstd])
            if rm is not None:
                rm.mul_(1 - momentum).add_(momentum * mean)
                rv.mul_(1 - momentum).add_(momentum * var * (m / (m - 1) if m > 1 else 1.0))
            if nbt is not None:
                nbt.add_(1)
        else:
            mean, invstd = rm, torch.rsqrt(rv + eps)
        v = (xr - mean) * (invstd * weight.float()) + bias.float()
        if residual is not None:
            v = v + _Emu._rows(residual).float()
        if relu:
            if need_mask:
                pos = (v > 0).view(m, c // 8, 8).to(torch.int32)
                mask = (pos << torch.arange(8, device=x.device, dtype=torch.int32)).sum(-1).to(torch.uint8).reshape(-1)
            v = v.clamp_min(0)
        return _Emu._like(v, x), saved, mask

    @staticmethod
    def _bwd_from_g(g, x, weight, saved):
        xr = _Emu._rows(x).float()
        m, c = xr.shape
        mean, invstd = saved[:c], saved[c:]
        xhat = (xr - mean) * invstd
        sdz, sdzx = g.sum(0), (g * xhat).sum(0)
        dx = (weight.float() * invstd) * (g - sdz / m - xhat * sdzx / m)
        return _Emu._like(dx, x), sdzx.to(weight.dtype), sdz.to(weight.dtype)

    @staticmethod
    def bn_act_backward(dy, x, mask, weight, saved, relu, has_res, work):
        g = _Emu._rows(dy).float()
        if relu:
            g = g * _Emu._unpack(mask, *g.shape)
        dx, dw, db = _Emu._bwd_from_g(g, x, weight, saved)
        dres = None
        if has_res:
            dres = dy if not relu else _Emu._like(g, x)
        return dx, dres, dw, db

    @staticmethod
    def bn_act_backward2(dy_a, dy_b, x, mask, weight, saved, relu, work):
        g = (_Emu._rows(dy_a).float() + _Emu._rows(dy_b).float()).to(x.dtype).float()      # rounded like an eager add
        if relu:
            g = g * _Emu._unpack(mask, *g.shape)
        dx, dw, db = _Emu._bwd_from_g(g, x, weight, saved)
        return dx, _Emu._like(g, x), dw, db


def _kernels(x):
    if x.is_cuda:
        from .. import _ext
        return _ext.lib()
    return _Emu


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, nbt, training, momentum, eps, relu, need_grad, pre=None,
                split=False):
        from .. import _ext
        C = _kernels(x)
        ctx.set_materialize_grads(False)
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
        if split:       # two aliases of one buffer: each consumer's gradient arrives separately in backward (no autograd add)
            return y, y.view_as(y)
        return y

    @staticmethod
    def backward(ctx, dy, dy2=None):
        from .. import _ext
        none = (None,) * 10
        if dy is None:
            dy, dy2 = dy2, None
        if dy is None:                          # neither alias was used
            return (None, None, None, None) + none
        x, mask, weight, saved = ctx.saved_tensors
        C = _kernels(x)
        work = ctx.work
        if work is None or (ctx.gen != -1 and ctx.gen != ctx.ws.generation):
            work = torch.zeros(2 * x.size(1), dtype=torch.float32, device=x.device)   # slice was recycled: use a fresh one
        _ext.note_launch(2)
        if dy2 is not None:                     # add + mask + reductions in one pass; g doubles as the residual gradient
            dx, dres, dw, db = C.bn_act_backward2(dy, dy2, x, mask, weight, saved, ctx.relu, work)
        else:
            dx, dres, dw, db = C.bn_act_backward(dy, x, mask, weight, saved, ctx.relu, ctx.has_res, work)
        return (dx, (dres if ctx.has_res else None), dw, db) + none


def _can_fuse(x, weight, residual, running_mean=True, emulate=False) -> bool:
    return ((x.is_cuda or emulate) and x.dim() == 4 and x.size(1) % 8 == 0 and x.size(1) <= 8192 and weight is not None and running_mean is not None
            and x.dtype in (torch.float32, torch.bfloat16, torch.float16)
            and x.is_contiguous(memory_format=torch.channels_last)
            and (residual is None or (residual.is_contiguous(memory_format=torch.channels_last) and residual.dtype == x.dtype
                                      and residual.shape == x.shape)))


def bn_act(x, weight, bias, running_mean, running_var, residual: Optional[torch.Tensor] = None, relu: bool = True,
           training: bool = True, momentum: float = 0.1, eps: float = 1e-5, fused=None,
           num_batches_tracked: Optional[torch.Tensor] = None, split: bool = False):
    """relu(batch_norm(x) + residual).  ``fused=None`` picks the CUDA kernels whenever the layout allows it;
    ``fused="emulate"`` runs the same autograd op over the PyTorch emulation of the kernels (CPU tests).
    ``split=True`` returns the result twice - two aliases of one buffer for the two consumers of a residual block's
    output - so that backward receives their gradients separately and fuses the add (``bn_act_backward2``)."""
    y = _bn_act(x, weight, bias, running_mean, running_var, residual, relu, training, momentum, eps, fused, num_batches_tracked, split)
    if split and not isinstance(y, tuple):
        return y, y
    return y


def _bn_act(x, weight, bias, running_mean, running_var, residual, relu, training, momentum, eps, fused, num_batches_tracked, split):
    ok = _can_fuse(x, weight, residual, running_mean, emulate=(fused == "emulate"))
    use = ok if fused is None else (bool(fused) and ok)
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
                          float(eps), relu, need_grad, None, bool(split and need_grad))
