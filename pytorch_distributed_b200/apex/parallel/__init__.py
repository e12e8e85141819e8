"""``apex.parallel.DistributedDataParallel(model)`` equivalent (/root/reference/apex_distributed.py:217).

apex's wrapper takes no ``device_ids``, learns its buckets (``message_size`` elements, default 1e7) from gradient
arrival and all-reduces flat buffers on a side stream.  Here it is a thin front over the same
:class:`~pytorch_distributed_b200.parallel.ddp.GradientEngine` as our DDP: fixed reverse-order buckets of
``message_size`` elements, the fused peer-memory kernel, fp16/bf16 wire chosen from the amp state, and - when amp's
dynamic loss scaling is on - the non-finite test folded into the all-reduce so all ranks skip the same steps.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...parallel import amp as _amp
from ...parallel.comm import make_communicator
from ...parallel.ddp import GradientEngine, _float_buffers, sync_module_states

_WIRE = {torch.float16: "fp16", torch.bfloat16: "bf16", torch.float32: "fp32"}


class DistributedDataParallel(nn.Module):
    """apex.parallel.DistributedDataParallel arguments and what they do here:

    ``message_size``              bucket size in ELEMENTS (apex default 1e7), fixed reverse-order buckets
    ``delay_allreduce``           True: nothing is launched from the hooks, every bucket goes out at the end of backward
    ``gradient_average``          divide by the world size (default) or leave the sum
    ``gradient_predivide_factor`` apex divides by f before and multiplies by f/world after the all-reduce to keep an fp16 sum in
                                  range; the fused kernel pre-scales in the pack pass and accumulates in fp32 (in the switch),
                                  so only the net factor matters: 1/world with averaging, 1/f without
    ``allreduce_always_fp32``     fp32 wire format
    ``retain_allreduce_buffers``  ``self.allreduce_buffers`` = the per-bucket flat slices of the symmetric gradient arena
    ``num_allreduce_streams`` > 1, ``allreduce_communicators``, ``allreduce_trigger_params``, ``shared_param`` are rejected:
    there is one communication stream and one (symmetric-memory) communicator by design.
    """

    def __init__(self, module: nn.Module, message_size: int = 10000000, delay_allreduce: bool = False, shared_param=None,
                 allreduce_trigger_params=None, retain_allreduce_buffers: bool = False, allreduce_always_fp32: bool = False,
                 num_allreduce_streams: int = 1, allreduce_communicators=None, gradient_average: bool = True,
                 gradient_predivide_factor: float = 1.0, comm="auto", wire_dtype=None, process_group=None):
        super().__init__()
        if shared_param is not None:
            raise ValueError("shared_param is no longer supported as an option (apex removed it as well); use delay_allreduce=True")
        if allreduce_trigger_params is not None:
            raise NotImplementedError("allreduce_trigger_params: buckets are fixed by message_size here, custom triggers are not supported")
        if num_allreduce_streams != 1:
            raise NotImplementedError("num_allreduce_streams=%r: the fused data plane uses exactly one communication stream" %
                                      (num_allreduce_streams,))
        if allreduce_communicators is not None:
            raise NotImplementedError("allreduce_communicators: the symmetric-memory communicator is created internally")
        self.module = module
        params = [p for p in module.parameters() if p.requires_grad]
        device = params[0].device
        if isinstance(comm, str):
            comm = make_communicator(comm, group=process_group, device=device)
        self.comm = comm
        scaler = _amp.current_scaler()
        if wire_dtype is None:
            if allreduce_always_fp32 or device.type != "cuda":
                wire_dtype = "fp32"
            elif scaler is not None:
                wire_dtype = _WIRE[_amp._amp_state.half_dtype]
            else:
                wire_dtype = "bf16"
        check_inf = scaler is not None and scaler.dynamic and getattr(comm, "backend", "") == "fused"
        if check_inf:
            scaler.rebind_found_inf(comm.found_inf)
        sync_module_states(module, comm, root=0)
        esz = 2 if wire_dtype != "fp32" else 4
        world = comm.world
        scale = (1.0 / world) if gradient_average else (1.0 / float(gradient_predivide_factor))
        self.engine = GradientEngine(params, comm, wire_dtype=wire_dtype, bucket_cap_mb=message_size * esz / float(1 << 20),
                                     first_bucket_mb=message_size * esz / float(1 << 20), check_inf=check_inf,
                                     average=gradient_average, delay_allreduce=delay_allreduce, scale=scale, tail_bucket_mb=None)
        self.delay_allreduce = delay_allreduce
        self.gradient_average = gradient_average
        self.gradient_predivide_factor = gradient_predivide_factor
        self.retain_allreduce_buffers = retain_allreduce_buffers
        self._buffers_f = _float_buffers(module)

    @property
    def allreduce_buffers(self):
        """Flat per-bucket views of the reduced gradients (apex ``retain_allreduce_buffers``); fused data plane only."""
        eng = self.engine
        if not self.retain_allreduce_buffers or not eng.fused:
            return []
        flat = eng.grad_arena()
        return [flat[b.elem_off:b.elem_off + b.region_elems] for b in eng.buckets]

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


class Reducer:
    """apex.parallel.Reducer: manual ``reduce()`` of a module's gradients (average across ranks)."""

    def __init__(self, module_or_grads_list, comm="auto"):
        if isinstance(module_or_grads_list, nn.Module):
            self.module = module_or_grads_list
            params = list(self.module.parameters())
        else:
            self.module = None
            params = list(module_or_grads_list)
        self.params = params
        self.comm = make_communicator(comm, device=params[0].device) if isinstance(comm, str) else comm
        if self.module is not None:
            sync_module_states(self.module, self.comm, root=0)

    def reduce(self):
        grads = [p.grad for p in self.params if p.grad is not None]
        self.comm.all_reduce_(grads, average=True)


def flatten(tensors):
    """apex_C.flatten: one contiguous 1-D tensor holding the given dense tensors back to back."""
    return torch.cat([t.contiguous().view(-1) for t in tensors]) if len(tensors) else torch.empty(0)


def unflatten(flat, tensors):
    """apex_C.unflatten: views of ``flat`` shaped like ``tensors``."""
    out, off = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat.narrow(0, off, n).view_as(t))
        off += n
    return out
