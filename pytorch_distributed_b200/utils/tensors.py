"""Small tensor helpers."""
import torch


def is_dense(t: torch.Tensor) -> bool:
    """True when ``t`` covers its storage range exactly once in some permutation (flat elementwise kernels may treat
    it as ``numel`` consecutive elements).  Fast paths first; the general check is torch's own predicate."""
    if t.is_contiguous():
        return True
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return True
    if t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d):
        return True
    from torch._prims_common import is_non_overlapping_and_dense
    return bool(is_non_overlapping_and_dense(t))
