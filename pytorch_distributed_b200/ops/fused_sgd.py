"""FusedSGD: SGD with momentum / weight decay as hand-written sm_100a kernels (``csrc/optim.cu``).

Drop-in for ``torch.optim.SGD`` as used at /root/reference/distributed.py:153-156 and, together with
:mod:`..parallel.amp`, for apex's patched optimizer + ``amp_C`` kernels (/root/reference/apex_distributed.py:211-216,330).

Three execution modes, picked automatically:
  * **flat / arena** - the parameters belong to one of our data-parallel engines: the reduced gradients are read
    directly from the symmetric wire arena (no write-back into ``p.grad``), master weights / momentum / model copy are
    flat buffers with the arena's layout, and the whole step is ONE streaming kernel.
  * **multi-tensor** - CUDA parameters without an engine: chunked multi-tensor-apply kernel over ``p.grad``.
  * **reference** - CPU tensors: plain PyTorch math (also the numerical oracle for the tests).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.optim import Optimizer


def sgd_reference_step(p, g, buf, lr, momentum, weight_decay, dampening, nesterov, first):
    """torch.optim.SGD semantics on plain tensors (fp32 math)."""
    g = g.float()
    if weight_decay != 0:
        g = g.add(p.float(), alpha=weight_decay)
    if momentum != 0:
        if first:
            buf.copy_(g)
        else:
            buf.mul_(momentum).add_(g, alpha=1 - dampening)
        g = g.add(buf, alpha=momentum) if nesterov else buf
    p.add_(g.to(p.dtype), alpha=-lr)


class FusedSGD(Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, flat: Optional[bool] = None,
                 overlap_backward: bool = False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self._steps = 0
        self._flat = None           # _FlatState when bound to an engine
        self._hyper = {}            # group index -> (device tensor, cached python tuple)
        self._amp = None            # LossScaler (set by amp.initialize)
        self._want_flat = flat
        # overlap_backward: in flat mode the update of each gradient bucket is enqueued on the communication stream right
        # behind that bucket's all-reduce (inside loss.backward()), so after the last gradient only the small tail bucket
        # remains; step() then only joins.  Contract: exactly one backward per step() and no gradient surgery between them
        # (the reference loop, /root/reference/distributed.py:267-269).  Inactive under a loss scaler (an overflow found in
        # a later bucket must be able to cancel the whole step).
        self._overlap = bool(overlap_backward)
        self._ov_active = False
        self._ov_applied = 0
        self._ov_first = False
        self._bind_refused = False  # an engine was found but declined (mixed dtypes, other parameter list): do not retry
        self._try_bind()

    # ------------------------------------------------------------------ engine binding (flat mode)
    def _try_bind(self):
        if self._want_flat is False or len(self.param_groups) != 1:
            return
        params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not params or not all(p.is_cuda for p in params):
            return
        engines = {getattr(p, "_ptd_engine", None) for p in params}
        if len(engines) != 1:
            return
        ref = engines.pop()
        eng = ref() if ref is not None else None
        if eng is None or not getattr(eng, "supports_flat_optimizer", False):
            return
        self._flat = eng.bind_flat_optimizer(self, params)
        if self._flat is None:
            self._bind_refused = True
        elif self._overlap and getattr(eng, "supports_overlap_optimizer", False):
            eng.set_overlap_optimizer(self)

    # ------------------------------------------------------------------ overlap mode (called by the gradient engine)
    def _prepare_overlap(self) -> None:
        """First bucket of a backward pass, on the compute stream: decide whether this step is applied bucket by bucket
        and push the hyper-parameters to the device before the side stream forks off."""
        if self._ov_active and self._ov_applied:
            raise RuntimeError("FusedSGD(overlap_backward=True): a second backward pass started before step() - the update of the "
                               "previous pass has already been applied bucket by bucket; use overlap_backward=False "
                               "(--no-overlap-optimizer) for gradient accumulation")
        self._ov_applied = 0
        self._ov_active = self._flat is not None and self._amp is None
        if self._ov_active:
            self._ov_first = self._steps == 0
            self._hyper_tensor(0, self.param_groups[0], self._flat.master.device)

    @torch.no_grad()
    def _apply_slice(self, off: int, n: int) -> None:
        """SGD update of flat elements [off, off + n) - the bucket whose all-reduce was just enqueued on this stream."""
        if not self._ov_active:
            return
        fs = self._flat
        from .. import _ext
        _ext.note_launch()
        copy = fs.model_copy[off:off + n] if fs.model_copy is not None else None
        _ext.lib().fused_sgd_flat(fs.engine.grad_arena()[off:off + n], fs.master[off:off + n], fs.momentum[off:off + n], copy,
                                  self._hyper[0][0], None, bool(self.param_groups[0]["nesterov"]), self._ov_first)
        self._ov_applied += n

    @property
    def is_flat(self) -> bool:
        return self._flat is not None

    # ------------------------------------------------------------------ hyper-parameters on the device
    def _hyper_tensor(self, gi: int, group, device):
        gmul = 1.0
        vals = (float(group["lr"]), float(group["momentum"]), float(group["weight_decay"]), float(group["dampening"]))
        ent = self._hyper.get(gi)
        if ent is None:
            t = torch.tensor(list(vals) + [gmul, 0, 0, 0], dtype=torch.float32, device=device)
            self._hyper[gi] = [t, vals]
            return t
        if ent[1] != vals:
            # only lr..dampening are host-owned; slot 4 (gradient multiplier) belongs to the loss scaler kernel
            ent[0][:4].copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=True)
            ent[1] = vals
        return ent[0]

    def refresh_hyper(self) -> None:
        """Push changed lr/momentum/weight-decay to the device copies (needed when ``step`` is replayed from a CUDA graph)."""
        for gi, ent in self._hyper.items():
            self._hyper_tensor(gi, self.param_groups[gi], ent[0].device)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        first = self._steps == 0
        amp = self._amp
        if self._flat is None and not self._bind_refused:
            # the engine may have been created after this optimizer (apex order: amp -> DDP), and a resumed optimizer
            # has _steps > 0 before its first step: bind whenever still unbound (existing momentum is carried over)
            self._try_bind()
        if self._flat is not None:
            fs = self._flat
            fs.engine.wait_for_gradients()
            group = self.param_groups[0]
            if self._ov_active and self._ov_applied == fs.master.numel():
                self._ov_applied = 0       # every bucket was updated behind its all-reduce during backward: nothing left to do
                self._steps += 1
                return loss
            if self._ov_applied:
                raise RuntimeError("FusedSGD(overlap_backward=True): only %d of %d elements were updated during backward" %
                                   (self._ov_applied, fs.master.numel()))
            hyper = self._hyper_tensor(0, group, fs.master.device)
            if amp is not None:
                amp.attach_hyper(hyper)
            from .. import _ext
            _ext.note_launch()
            _ext.lib().fused_sgd_flat(fs.engine.grad_arena(), fs.master, fs.momentum, fs.model_copy, hyper,
                                      amp.found_inf if amp is not None else None, bool(group["nesterov"]), first)
            if amp is not None:
                amp.update()
        else:
            for gi, group in enumerate(self.param_groups):
                self._step_group(gi, group, first, amp)
            if amp is not None:
                amp.update()
        self._steps += 1
        return loss

    def _step_group(self, gi, group, first, amp):
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return
        bufs = []
        for p in params:
            st = self.state[p]
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
                st["_fresh"] = True
            bufs.append(st["momentum_buffer"])
        if params[0].is_cuda and all(p.dtype in (torch.float32, torch.bfloat16, torch.float16) for p in params):
            from .. import _ext
            C = _ext.lib()
            hyper = self._hyper_tensor(gi, group, params[0].device)
            if amp is not None:
                amp.attach_hyper(hyper)
            fresh = [p for p in params if self.state[p].pop("_fresh", False)]

            def master_of(p):
                """fp32 tensor the kernel updates for a low-precision parameter (multi-tensor mode of a bf16 / fp16 model without
                a flat engine, e.g. under hvd.DistributedOptimizer); the parameter itself is refreshed from it as the model copy."""
                st = self.state[p]
                if "master" not in st:
                    init = getattr(p, "_ptd_master_init", None)         # fp32 values stashed by amp.cast_model
                    st["master"] = (init if init is not None else p.detach().float()).clone(memory_format=torch.preserve_format)
                    if init is not None:
                        del p._ptd_master_init
                return st["master"]

            fs = set(fresh) if (fresh and len(fresh) != len(params)) else None      # rare: first gradient later than the others
            for first_flag, sub in ((True, fresh), (False, [p for p in params if p not in fs])) if fs is not None else ((bool(fresh), params),):
                full = [p for p in sub if p.dtype == torch.float32]
                low = [p for p in sub if p.dtype != torch.float32]
                if full:
                    C.fused_sgd_multi([p.grad for p in full], full, [self.state[p]["momentum_buffer"] for p in full], [], hyper,
                                      amp.found_inf if amp is not None else None, bool(group["nesterov"]), first_flag)
                if low:
                    C.fused_sgd_multi([p.grad for p in low], [master_of(p) for p in low], [self.state[p]["momentum_buffer"] for p in low],
                                      [p.data for p in low], hyper, amp.found_inf if amp is not None else None, bool(group["nesterov"]),
                                      first_flag)
        else:
            if amp is not None and amp.host_found_inf():
                return
            gmul = amp.host_inv_scale() if amp is not None else 1.0
            for p, buf in zip(params, bufs):
                fresh = self.state[p].pop("_fresh", False)
                g = p.grad if gmul == 1.0 else p.grad.float() * gmul
                if p.dtype == torch.float32:
                    sgd_reference_step(p, g, buf, group["lr"], group["momentum"], group["weight_decay"], group["dampening"],
                                       group["nesterov"], fresh)
                else:
                    st = self.state[p]
                    if "master" not in st:
                        st["master"] = p.detach().float().clone()
                    sgd_reference_step(st["master"], g, buf, group["lr"], group["momentum"], group["weight_decay"],
                                       group["dampening"], group["nesterov"], fresh)
                    p.copy_(st["master"])

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)

    def load_state_dict(self, state_dict):
        """Standard ``Optimizer.load_state_dict``; in flat mode the loaded momentum is copied INTO the flat buffer
        (the kernel reads that buffer, not the per-parameter tensors) and the state entries are re-pointed at its views."""
        super().load_state_dict(state_dict)
        loaded = False
        if self._flat is not None:
            eng = self._flat.engine
            with torch.no_grad():
                for i, p in enumerate(eng.params):
                    buf = self.state.get(p, {}).get("momentum_buffer")
                    off, cnt = eng.param_elem_off[i], p.numel()
                    view = self._flat.momentum[off:off + cnt].as_strided(p.size(), p.stride())
                    if buf is not None and buf.data_ptr() != view.data_ptr():
                        view.copy_(buf.to(view.dtype))
                        loaded = True
                    self.state[p]["momentum_buffer"] = view
        else:
            loaded = any("momentum_buffer" in st for st in self.state.values())
        if loaded:
            self._steps = max(self._steps, 1)      # do not re-initialise the momentum on the next step
