"""Training meters, accuracy and LR schedule.

Output contract mirrors the reference scripts (identical in all six):
  AverageMeter        /root/reference/distributed.py:333-354
  ProgressMeter       /root/reference/distributed.py:357-371
  adjust_learning_rate /root/reference/distributed.py:374-378
  accuracy            /root/reference/distributed.py:381-395

Deviation: ``accuracy`` uses a rank-counting formulation instead of
``topk`` + ``view(-1)`` (the reference form raises on torch >= 1.7 for k=5).
On CUDA the driver uses the fused kernel in ``ops.metrics`` instead; this file
is the plain-PyTorch oracle and the CPU path.
"""
from __future__ import annotations

import torch


class AverageMeter:
    """Tracks the latest value and the running (count-weighted) average."""

    def __init__(self, name: str, fmt: str = ":f"):
        self.name = name
        self.fmt = fmt
        self.reset()

    def reset(self) -> None:
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n: int = 1) -> None:
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self) -> str:
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(
            name=self.name, val=self.val, avg=self.avg)


class ProgressMeter:
    """Prints ``prefix[ cur/total]\\tmeter\\tmeter...`` lines."""

    def __init__(self, num_batches: int, meters, prefix: str = ""):
        width = len(str(num_batches // 1))
        self._fmt = "[{:" + str(width) + "d}/" + ("{:" + str(width) + "d}").format(num_batches) + "]"
        self.meters = list(meters)
        self.prefix = prefix

    def line(self, batch: int) -> str:
        return "\t".join([self.prefix + self._fmt.format(batch)] + [str(m) for m in self.meters])

    def display(self, batch: int) -> None:
        print(self.line(batch), flush=True)


def step_lr(base_lr: float, epoch: int, step: int = 30, gamma: float = 0.1) -> float:
    return base_lr * (gamma ** (epoch // step))


def adjust_learning_rate(optimizer, epoch: int, args) -> float:
    """LR = args.lr decayed 10x every 30 epochs, written into every param group."""
    lr = step_lr(args.lr, epoch)
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr


@torch.no_grad()
def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """Top-k accuracy in percent; returns a list of 1-element tensors.

    A sample is top-k correct iff fewer than k logits are strictly greater than
    the target-class logit (ties resolved in favour of the target, which only
    differs from ``topk`` on exact float ties).
    """
    batch = target.size(0)
    out = output.float()
    tgt_logit = out.gather(1, target.view(-1, 1))
    rank = (out > tgt_logit).sum(dim=1)
    res = []
    for k in topk:
        correct_k = (rank < k).float().sum(0, keepdim=True)
        res.append(correct_k.mul_(100.0 / batch))
    return res
