"""Shared command line for every entrypoint.

Flag surface = /root/reference/distributed.py:25-102 (the same block is copied
into all six reference scripts; ``--local_rank`` exists only in distributed.py:73
and apex_distributed.py:76, ``--dist-file`` only in distributed_slurm_main.py:102).
Reference-compatible defaults are kept; everything else is additive.

Deviations: both ``--local_rank`` and ``--local-rank`` are accepted
and fall back to $LOCAL_RANK; ``-j/--workers`` is honoured.
"""
from __future__ import annotations

import argparse
import os


def model_names():
    from .models import available_models
    return available_models()


def build_parser(entry: str = "distributed") -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="B200-native ImageNet training (%s)" % entry)
    p.add_argument("--data", metavar="DIR", default=os.environ.get("IMAGENET_DIR", ""),
                   help="path to dataset (DIR/train, DIR/val ImageFolder trees); "
                        "empty or --synthetic => synthetic ImageNet-shaped data")
    p.add_argument("-a", "--arch", metavar="ARCH", default="resnet18", choices=model_names(),
                   help="model architecture (default: resnet18)")
    p.add_argument("-j", "--workers", default=4, type=int, metavar="N",
                   help="number of data loading workers (default: 4)")
    p.add_argument("--epochs", default=90, type=int, metavar="N", help="number of total epochs to run")
    p.add_argument("--start-epoch", default=0, type=int, metavar="N",
                   help="manual epoch number (useful on restarts)")
    p.add_argument("-b", "--batch-size", default=3200, type=int, metavar="N",
                   help="mini-batch size (default: 3200): total batch of all GPUs on the node")
    p.add_argument("--lr", "--learning-rate", default=0.1, type=float, metavar="LR",
                   help="initial learning rate", dest="lr")
    p.add_argument("--momentum", default=0.9, type=float, metavar="M", help="momentum")
    p.add_argument("--wd", "--weight-decay", default=1e-4, type=float, metavar="W",
                   help="weight decay (default: 1e-4)", dest="weight_decay")
    p.add_argument("-p", "--print-freq", default=10, type=int, metavar="N",
                   help="print frequency (default: 10)")
    p.add_argument("-e", "--evaluate", dest="evaluate", action="store_true",
                   help="evaluate model on validation set")
    p.add_argument("--pretrained", dest="pretrained", action="store_true", help="use pre-trained model")
    p.add_argument("--seed", default=None, type=int, help="seed for initializing training.")

    if entry in ("distributed", "apex_distributed"):
        p.add_argument("--local_rank", "--local-rank", default=-1, type=int, dest="local_rank",
                       help="local rank injected by the launcher (falls back to $LOCAL_RANK)")
    if entry == "distributed_slurm_main":
        p.add_argument("--dist-file", default=None, type=str, help="shared file for file:// rendezvous")
    if entry == "dataparallel":
        p.add_argument("--gpus", default=None, type=str,
                       help="comma separated device ids (default: all visible; the reference hard-codes 0,1,2,3)")
    if entry == "apex_distributed":
        p.add_argument("--opt-level", default="O1", choices=["O0", "O1", "O2", "O3"],
                       help="amp optimisation level (reference passes none => O1)")
        p.add_argument("--loss-scale", default="dynamic", help="'dynamic' or a float")
    if entry == "horovod_distributed":
        p.add_argument("--compression", default="fp16", choices=["none", "fp16", "bf16"],
                       help="wire compression for the DistributedOptimizer (reference: fp16)")

    # ---- additive, B200-native knobs (all have reference-compatible defaults) ----
    x = p.add_argument_group("b200")
    x.add_argument("--synthetic", action="store_true", help="force synthetic ImageNet-shaped data")
    x.add_argument("--synthetic-size", default=None, type=int, metavar="N",
                   help="images per synthetic epoch (default: 1,281,167 train / 50,000 val; "
                        "--steps-per-epoch overrides)")
    x.add_argument("--steps-per-epoch", default=None, type=int, help="cap iterations per epoch (tests/bench)")
    x.add_argument("--val-steps", default=None, type=int, help="cap validation iterations")
    x.add_argument("--image-size", default=224, type=int)
    x.add_argument("--num-classes", default=1000, type=int)
    x.add_argument("--comm", default="auto", choices=["auto", "fused", "nccl", "gloo"],
                   help="gradient data plane: fused = sm_100a peer-memory kernels, nccl/gloo = library all-reduce")
    x.add_argument("--wire-dtype", default="bf16", choices=["bf16", "fp16", "fp32"],
                   help="gradient wire format of the fused all-reduce")
    x.add_argument("--bucket-cap-mb", default=8.0, type=float,
                   help="gradient bucket size cap in MiB of wire data (torch's default is 25; NVSwitch has no per-link cost, "
                        "so smaller buckets only buy earlier overlap; first bucket 1 MiB, tail bucket 1 MiB)")
    x.add_argument("--precision", default=None, choices=["fp32", "bf16", "fp16"],
                   help="compute precision (default: bf16 on CUDA, fp32 on CPU)")
    x.add_argument("--channels-last", dest="channels_last", action="store_true", default=None)
    x.add_argument("--no-channels-last", dest="channels_last", action="store_false")
    x.add_argument("--fused-bn", dest="fused_bn", action="store_true", default=None,
                   help="use the hand-written NHWC BN(+add)+ReLU kernels in the ResNet family")
    x.add_argument("--no-fused-bn", dest="fused_bn", action="store_false")
    x.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                   help="fused = hand-written multi-tensor SGD kernel; torch = torch.optim.SGD")
    x.add_argument("--cuda-graph", action="store_true", help="capture the train step in a CUDA graph")
    x.add_argument("--overlap-optimizer", dest="overlap_optimizer", action="store_true", default=True,
                   help="fused optimizer: update each gradient bucket right behind its all-reduce, inside backward (default)")
    x.add_argument("--no-overlap-optimizer", dest="overlap_optimizer", action="store_false")
    x.add_argument("--bucket-view", action="store_true",
                   help="DDP gradient_as_bucket_view: p.grad are views of the symmetric arena (no write-back / pack pass)")
    x.add_argument("--device", default=None, help="cuda|cpu (default: cuda if available)")
    x.add_argument("--dist-backend", default=None, help="control-plane backend (default nccl on CUDA, gloo on CPU)")
    x.add_argument("--dist-url", default=None, help="override rendezvous URL")
    x.add_argument("--world-size", default=None, type=int, help="processes to spawn (mp/hvd self-launch)")
    x.add_argument("--resume", default="", type=str, help="checkpoint to resume from (extension; SURVEY Q10)")
    x.add_argument("--checkpoint-dir", default=".", type=str)
    x.add_argument("--log-jsonl", default="", type=str, help="append machine-readable step records here")
    x.add_argument("--quiet", action="store_true")
    return p


def resolve_local_rank(args) -> int:
    lr = getattr(args, "local_rank", -1)
    if lr is None or lr < 0:
        lr = int(os.environ.get("LOCAL_RANK", "0"))
    args.local_rank = lr
    return lr


def parse_args(entry: str, argv=None):
    args = build_parser(entry).parse_args(argv)
    args.entry = entry
    return args
