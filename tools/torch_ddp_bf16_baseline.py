"""Context row for the ablation table (NOT the reference arm): stock PyTorch at the same precision / layout as this framework.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/torch_ddp_bf16_baseline.py [--steps 20 --warmup 5]

torchvision ResNet-50, channels_last, bf16 autocast, torch.nn.parallel.DistributedDataParallel over NCCL, torch.optim.SGD -
i.e. what a user gets from the reference's distributed.py by adding the two standard lines for bf16 + NHWC, none of this
repo's code.  It separates "bf16 + NHWC" (a PyTorch switch) from what the hand-written kernels, the fused data plane, the
flat optimizer and the CUDA-graph step earn on top.  Same timing protocol as bench.py (device-resident inputs, CUDA events,
barrier + synchronize on both sides, max over ranks); prints one JSON line.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--batch-per-gpu", type=int, default=256)
    p.add_argument("--arch", default="resnet50")
    a = p.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if "RANK" not in os.environ:
        os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29544")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torchvision
    torch.backends.cudnn.benchmark = True
    model = torchvision.models.__dict__[a.arch]().to(dev).to(memory_format=torch.channels_last)
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss().to(dev)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pool = [(torch.randn(a.batch_per_gpu, 3, 224, 224, generator=g).to(dev).contiguous(memory_format=torch.channels_last),
             torch.randint(0, 1000, (a.batch_per_gpu,), generator=g).to(dev)) for _ in range(4)]

    def step(i):
        x, y = pool[i % 4]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
            loss = crit(out, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({"impl": "torch DDP + bf16 autocast + channels_last (context row, stock PyTorch)", "value": a.batch_per_gpu * world * a.steps / (ms / 1e3),
                          "unit": "images/s", "n_gpus": world, "ms_per_step": ms / a.steps, "steps": a.steps, "warmup": a.warmup,
                          "final_loss": float(loss.item())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
