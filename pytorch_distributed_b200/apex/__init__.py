"""apex-compatible namespace: ``from pytorch_distributed_b200.apex import amp`` and
``from pytorch_distributed_b200.apex.parallel import DistributedDataParallel`` stand in for the two apex imports of
/root/reference/apex_distributed.py:21-22 (apex itself is not installable offline; this is a from-scratch equivalent)."""
from ..parallel import amp  # noqa: F401
from . import optimizers, parallel  # noqa: F401
