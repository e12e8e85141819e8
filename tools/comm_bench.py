"""Micro-benchmark of the fused collectives against NCCL (run under torchrun on N GPUs).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/comm_bench.py > profiles/comm_bench_8gpu.md

Per size: device time (CUDA events, max over ranks, median of `reps`) of
  K1 two-shot  : pack(fp32->wire, x1/world) + reduce-scatter + all-gather in ONE kernel (NVLS and P2P variants)
  NCCL         : what torch DDP does: div_ + all_reduce on a flat fp32 (or bf16) bucket
and the bus bandwidth  2*(W-1)/W * wire_bytes / t  with the roofline fraction against the measured 8-rank NCCL bus
bandwidth reference of the profiling guide (725 GB/s) and the per-direction peer-copy number (770 GB/s).
Latency section: K4 metrics kernel vs accuracy()+barrier+3 x all_reduce (the reference's per-iteration sync).
"""
import os
import statistics
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps, device, sync):
    times = []
    for _ in range(reps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(device)
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    return statistics.median(times)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pytorch_distributed_b200.parallel.comm import KIND_TWO_SHOT, FusedCommunicator
    from pytorch_distributed_b200.utils.meters import accuracy
    comm = FusedCommunicator(device=dev, arena_bytes=1 << 30)

    def sync():
        torch.cuda.synchronize(dev)
        dist.barrier()

    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    say("# Fused collectives vs NCCL, %d x B200 (symm=%s, nvls=%s)\n" % (world, comm.symm_backend, comm.nvls))
    say("Device-timed (CUDA events), max over ranks, median of 7 after 3 warm-ups; one bucket per launch, 32 CTAs x 512 threads.\n")
    say("| elements | wire | fused NVLS us | fused P2P us | NCCL (div+all_reduce) us | fused busbw GB/s | frac of 725 GB/s | NCCL busbw GB/s |")
    say("|---:|---|---:|---:|---:|---:|---:|---:|")
    for n in (1 << 14, 1 << 18, 1 << 20, 1 << 22, 6_000_000, 1 << 24, 25_600_000, 1 << 26):
        for wire, wbytes in (("bf16", 2), ("fp32", 4)):
            src = torch.randn(n, device=dev)
            plan = comm.make_plan([n], wire)
            flat = src.clone() if wire == "fp32" else src.bfloat16()

            def fused(nvls):
                comm.run(plan, [src], KIND_TWO_SHOT, comm.misc_channel, scale=1.0 / world, writeback=False, nvls=nvls)

            def nccl():
                flat.div_(world)
                dist.all_reduce(flat)

            for _ in range(3):
                fused(True); fused(False); nccl()
            t_nvls = timed(lambda: fused(True), 7, dev, sync) if comm.nvls else float("nan")
            t_p2p = timed(lambda: fused(False), 7, dev, sync)
            t_nccl = timed(nccl, 7, dev, sync)
            best = min(t for t in (t_nvls, t_p2p) if t == t)
            bus = 2 * (world - 1) / world * n * wbytes / (best * 1e-3) / 1e9
            busn = 2 * (world - 1) / world * n * wbytes / (t_nccl * 1e-3) / 1e9
            say("| %d | %s | %.1f | %.1f | %.1f | %.0f | %.2f | %.0f |" % (n, wire, t_nvls * 1e3, t_p2p * 1e3, t_nccl * 1e3, bus, bus / 725.0, busn))
    # ---- K1 with a wider grid (64 CTAs) for the big messages, where a 32-CTA pack phase cannot stream HBM fast enough
    say("\n## K1 two-shot NVLS: 32 vs 64 CTAs\n")
    say("| elements | wire | 32 CTAs us | 64 CTAs us |\n|---:|---|---:|---:|")
    for n in (1 << 24, 25_600_000, 1 << 26):
        wide = FusedCommunicator(device=dev, arena_bytes=1 << 30)      # plans are never freed: a fresh 1 GiB arena per size
        for wire in ("bf16", "fp32"):
            src = torch.randn(n, device=dev)
            plans = {c: wide.make_plan([n], wire, max_ctas=c) for c in (32, 64)}
            ts = {}
            for c, pl in plans.items():
                fn = lambda pl=pl: wide.run(pl, [src], KIND_TWO_SHOT, wide.misc_channel, scale=1.0 / world, writeback=False, nvls=wide.nvls)
                for _ in range(3):
                    fn()
                ts[c] = timed(fn, 7, dev, sync)
            say("| %d | %s | %.1f | %.1f |" % (n, wire, ts[32] * 1e3, ts[64] * 1e3))
    # ---- K2: broadcast of a ResNet-50-shaped tensor list from rank 0 (DDP constructor / per-forward buffer sync)
    from pytorch_distributed_b200.models import create_model
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model("resnet50").to(dev)
    bcomm = FusedCommunicator(device=dev, arena_bytes=1 << 30)
    params = [p.data for p in model.parameters()]
    bufs = [b for b in model.buffers() if b.is_floating_point()]
    say("\n## K2 broadcast from rank 0 (ResNet-50 tensor lists)\n")
    say("| tensors | elements | fused us | NCCL (flatten + broadcast + unflatten) us |\n|---:|---:|---:|---:|")
    for name, ts_ in (("parameters", params), ("BN buffers", bufs)):
        def fused_b(ts_=ts_):
            bcomm.broadcast_(ts_, root=0)

        def nccl_b(ts_=ts_):
            flat = torch._utils._flatten_dense_tensors(ts_)
            dist.broadcast(flat, src=0)
            for t, f in zip(ts_, torch._utils._unflatten_dense_tensors(flat, ts_)):
                t.copy_(f)

        for _ in range(3):
            fused_b(); nccl_b()
        tf, tn = timed(fused_b, 7, dev, sync), timed(nccl_b, 7, dev, sync)
        say("| %d (%s) | %d | %.1f | %.1f |" % (len(ts_), name, sum(t.numel() for t in ts_), tf * 1e3, tn * 1e3))
    # ---- latency: per-iteration metric synchronisation
    logits = torch.randn(256, 1000, device=dev).bfloat16()
    target = torch.randint(0, 1000, (256,), device=dev)
    loss = torch.tensor(1.0, device=dev)
    out = torch.zeros(4, device=dev)

    def k4():
        comm.metrics(logits, target, loss, out)

    def ref():
        a1, a5 = accuracy(logits, target, (1, 5))
        dist.barrier()
        for t in (loss.clone(), a1, a5):
            dist.all_reduce(t)
            t /= world

    for _ in range(5):
        k4(); ref()
    t4 = timed(k4, 15, dev, sync)
    tr = timed(ref, 15, dev, sync)
    say("\n## Metric synchronisation per iteration (logits 256x1000 bf16)\n")
    say("| path | device us |\n|---|---:|")
    say("| K4 `metrics_kernel` (top-k counting + LL all-reduce, 1 launch) | %.1f |" % (t4 * 1e3))
    say("| reference sequence: accuracy() + barrier + 3 x (clone, all_reduce, div) via NCCL | %.1f |" % (tr * 1e3))
    comm.check()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
