"""``apex.optimizers`` namespace: apex's FusedSGD is this framework's :class:`~pytorch_distributed_b200.ops.fused_sgd.FusedSGD`
(flat arena mode under a data-parallel engine, multi-tensor kernel otherwise; unscale / overflow-skip fused when amp is active)."""
from ..ops.fused_sgd import FusedSGD  # noqa: F401
