"""Small launch targets for `ncu` captures of the collective / optimizer kernels (K1..K6).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> -o gpurun_out/<name> python tools/ncu_targets.py <mode>

modes
  single   one GPU, world 1: K1 two-shot / K1b one-shot (pack + local reduce phases, 8 MB and 32 MB of bf16), K4 metrics,
           K6 flat SGD over 25.6 M parameters, the normalise/cast/NHWC kernel.  No peers => no NVLink traffic; DRAM / issue behaviour.
  local    ONE process over all visible GPUs (the DataParallel engine): K2' push (pack + multimem.st), K5 pack + reduce-to-caller
           (multimem.ld_reduce pulled by the root).  These kernels carry no flags (events order the devices), so ncu's kernel replay
           is safe and the NVLink / L2 counters of the in-switch paths can be read.
  rank     one RANK of a multi-process job (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT / LOCAL_RANK set by hand): K1 two-shot and
           K2 broadcast with their in-kernel flag barriers.  Only rank 0 runs under ncu; the other ranks run the same script plainly.
           Replay is idempotent because the flags are monotonic sequence numbers (a replayed pass sees `flag >= seq` already true)
           and every launch is fenced by a host barrier, so the peers hold still while rank 0 replays.
Run ncu with `--profile-from-start off`: every launch of interest is preceded by 3 warm-up launches outside the profiler range
(cudaProfilerStart/Stop bracket exactly one launch of each kernel).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def profiled(fn, warm=3, before=None, devices=None):
    """`warm` launches outside the profiler range, then exactly one inside it."""
    def sync():
        for d in (devices or [torch.cuda.current_device()]):
            torch.cuda.synchronize(d)
    for _ in range(warm):
        if before:
            before()
        fn()
    sync()
    if before:
        before()
    torch.cuda.profiler.start()
    fn()
    sync()
    torch.cuda.profiler.stop()


def single():
    from pytorch_distributed_b200 import _ext
    from pytorch_distributed_b200.parallel.comm import KIND_ONE_SHOT, KIND_TWO_SHOT, FusedCommunicator
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    comm = FusedCommunicator(device=dev, arena_bytes=1 << 30)
    C = _ext.lib()
    for n in (1 << 22, 1 << 24):                      # 8 MB / 32 MB of bf16 wire data
        src = torch.randn(n, device=dev).bfloat16()
        plan = comm.make_plan([n], "bf16")
        profiled(lambda: comm.run(plan, [src], KIND_TWO_SHOT, comm.misc_channel, scale=1.0, writeback=False))
        view = plan.arena_tensor()[:n]
        profiled(lambda: comm.run(plan, [view], KIND_TWO_SHOT, comm.misc_channel, scale=0.5, writeback=False, prepacked=True))
    src = torch.randn(1 << 17, device=dev).bfloat16()
    p1 = comm.make_plan([src.numel()], "bf16", double_buffer=True)
    profiled(lambda: comm.run(p1, [src], KIND_ONE_SHOT, comm.misc_channel, scale=1.0, writeback=False))
    logits = torch.randn(256, 1000, device=dev).bfloat16()
    target = torch.randint(0, 1000, (256,), device=dev)
    out = torch.zeros(4, device=dev)
    one = torch.tensor(1.0, device=dev)
    profiled(lambda: comm.metrics(logits, target, one, out))
    n = 25_600_000
    grad = torch.randn(n, device=dev).bfloat16()
    master, mom, copy = torch.randn(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev, dtype=torch.bfloat16)
    hyper = torch.tensor([0.1, 0.9, 1e-4, 0.0, 1.0, 0, 0, 0], device=dev)
    profiled(lambda: C.fused_sgd_flat(grad, master, mom, copy, hyper, None, False, False))
    raw = torch.randint(0, 255, (256, 3, 224, 224), device=dev, dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev) * 255
    std = torch.tensor([0.229, 0.224, 0.225], device=dev) * 255
    profiled(lambda: C.normalize_nhwc(raw, mean, std, 1, True))
    torch.cuda.synchronize()
    print("single done")


def local():
    from pytorch_distributed_b200.parallel.comm import KIND_PACK, KIND_PUSH, KIND_REDUCE
    from pytorch_distributed_b200.parallel.dp import LocalCommunicator, _TensorSet
    devices = list(range(torch.cuda.device_count()))
    comm = LocalCommunicator(devices, 1 << 30)
    print("local engine over %d devices, nvls=%s" % (len(devices), comm.nvls), flush=True)
    n = 12_800_000                                     # 25.6 MB of bf16: one DDP-sized bucket
    per_dev = []
    for d in devices:
        with torch.cuda.device(d):
            per_dev.append([torch.randn(n, device="cuda:%d" % d).bfloat16()])
    ts = _TensorSet(comm, per_dev, "bf16")
    def pack_all():
        for r in range(len(devices)):
            with torch.cuda.device(devices[r]):
                ts.launch(KIND_PACK, r)
        for d in devices:
            torch.cuda.synchronize(d)

    def reduce_root():
        with torch.cuda.device(0):
            ts.launch(KIND_REDUCE, 0, writeback=False)

    def push_root():
        with torch.cuda.device(0):
            ts.launch(KIND_PUSH, 0)

    torch.cuda.set_device(0)
    pack_all()
    profiled(lambda: (pack_all(), reduce_root()), devices=devices)      # K5: one pack per device + the in-switch pull by the root
    profiled(push_root, devices=devices)                                # K2'

    print("local done")


def rank():
    import torch.distributed as dist
    r, local_rank, world = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"])), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from pytorch_distributed_b200.parallel.comm import KIND_TWO_SHOT, FusedCommunicator
    comm = FusedCommunicator(device=dev, arena_bytes=1 << 30, timeout_ms=60000)

    def fence():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    for n in (1 << 22, 1 << 24):
        src = torch.randn(n, device=dev).bfloat16()
        plan = comm.make_plan([n], "bf16")
        view = plan.arena_tensor()[:n]
        profiled(lambda: comm.run(plan, [src], KIND_TWO_SHOT, comm.misc_channel, scale=1.0 / world, writeback=False), before=fence)
        profiled(lambda: comm.run(plan, [view], KIND_TWO_SHOT, comm.misc_channel, scale=1.0 / world, writeback=False, prepacked=True),
                 before=fence)
    big = [torch.randn(1 << 23, device=dev)]
    profiled(lambda: comm.broadcast_(big, root=0), before=fence)
    fence()
    comm.check()
    print("rank %d done (nvls=%s)" % (r, comm.nvls), flush=True)
    dist.destroy_process_group()


def stem():
    """The fused stem at the benchmark shape: im2col + tcgen05 GEMM (+BN statistics) + BN/ReLU/MaxPool forward, and the
    two-pass quad backward (ncu -k regex:stem_)."""
    import torch.nn as nn
    from pytorch_distributed_b200.models.resnet import BNAct
    from pytorch_distributed_b200.ops.bn_act import begin_step
    from pytorch_distributed_b200.ops.stem_conv import stem_conv_bn_relu_maxpool
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    conv = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False).to(dev).bfloat16().to(memory_format=torch.channels_last)
    bn = BNAct(64).to(dev).train()
    x = torch.randn(256, 3, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)

    def step():
        begin_step(dev)
        y = stem_conv_bn_relu_maxpool(x, conv, bn)
        y.backward(torch.ones_like(y))
        conv.weight.grad = None
        bn.weight.grad = None
        bn.bias.grad = None

    profiled(step)
    print("stem done")


if __name__ == "__main__":
    {"single": single, "local": local, "rank": rank, "stem": stem}[sys.argv[1]]()
