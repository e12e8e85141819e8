"""CPU tier for the wrapper APIs: hvd shim over gloo, apex namespace, prefetcher, metric pipeline, train step, reduce_mean."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HVD = r'''
import os, sys, torch
sys.path.insert(0, %r)
import pytorch_distributed_b200.parallel.hvd as hvd
hvd.init(device="cpu")
r, n = hvd.rank(), hvd.size()
assert n == 2 and hvd.local_rank() == r and hvd.is_initialized()
t = torch.full((5,), float(r + 1))
out = hvd.allreduce(t, name="barrier")                 # out-of-place, averaged (SURVEY Q4: really returned)
assert torch.allclose(out, torch.full((5,), 1.5)) and torch.allclose(t, torch.full((5,), float(r + 1)))
hvd.allreduce_(t, average=False)
assert torch.allclose(t, torch.full((5,), 3.0))
assert torch.allclose(hvd.allreduce(torch.full((2,), float(r + 1)), op=hvd.Sum), torch.full((2,), 3.0))
assert torch.allclose(hvd.allreduce(torch.full((2,), float(r + 1)), op=hvd.Average), torch.full((2,), 1.5))
g = hvd.allgather(torch.full((r + 1, 2), float(r)))                      # ragged first dimension: 1 row from rank 0, 2 from rank 1
assert g.shape == (3, 2) and g[:, 0].tolist() == [0.0, 1.0, 1.0]
assert hvd.broadcast_object({"lr": 0.1 * (r + 1)}, root_rank=1) == {"lr": 0.2}
hvd.barrier()
h = hvd.allreduce_async_(torch.full((3,), float(r)), average=True)
assert hvd.poll(h) in (True, False)
assert torch.allclose(hvd.synchronize(h), torch.full((3,), 0.5))
m = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.BatchNorm1d(8), torch.nn.Linear(8, 2))
torch.manual_seed(10 + r)
for p in m.parameters():
    torch.nn.init.normal_(p)
hvd.broadcast_parameters(m.state_dict(), root_rank=0)
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
ref = flat.clone(); torch.distributed.broadcast(ref, 0)
assert torch.equal(flat, ref)
opt = torch.optim.SGD(m.parameters(), lr=0.1 * (r + 1), momentum=0.9)
hvd.broadcast_optimizer_state(opt, root_rank=0)
assert opt.param_groups[0]["lr"] == 0.1
opt = hvd.DistributedOptimizer(opt, named_parameters=m.named_parameters(), compression=hvd.Compression.fp16)
assert isinstance(opt, torch.optim.SGD)
for it in range(int(os.environ.get("PTD_TEST_ITERS", "3"))):
    torch.manual_seed(100 + r + it)
    x, y = torch.randn(6, 4), torch.randint(0, 2, (6,))
    opt.zero_grad()
    torch.nn.functional.cross_entropy(m(x), y).backward()
    opt.step()
eng = opt._ptd_engine_obj
if os.environ.get("HOROVOD_AUTOTUNE") == "1":
    assert eng._tuner is not None and eng._tuner.done and eng.cycle_bytes in eng._tuner.cands
    eng.write_timeline()
if os.environ.get("PTD_HVD_STATIC") == "1":           # frozen after the first complete step; the hooks launched steps 2 and 3 themselves
    assert eng._schedule is not None and sorted(i for g in eng._schedule for i in g) == list(range(len(eng.params)))
else:
    assert eng._schedule is None
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
print("HVD-SUM {} {:.10f}".format(r, flat.double().sum().item()))
lo, hi = flat.clone(), flat.clone()
torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN); torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
assert torch.allclose(lo, hi, atol=1e-6), (lo - hi).abs().max()      # averaged gradients => identical weights on all ranks
try:
    hvd.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=[("a", p) for p in m.parameters()])
    raise SystemExit("duplicate names must be rejected")
except ValueError:
    pass
from pytorch_distributed_b200.utils.dist_ops import reduce_mean
assert abs(reduce_mean(torch.tensor(float(r)), 2).item() - 0.5) < 1e-6
print("HVD-OK", r)
torch.distributed.destroy_process_group()
'''


def test_hvd_api_over_gloo(tmp_path):
    """Dynamic (fusion queue + dispatcher thread) and static (frozen schedule, hooks launch the groups) modes give the same weights."""
    script = tmp_path / "hvd_check.py"
    script.write_text(HVD % ROOT)
    sums = {}
    for static, port in (("0", "29731"), ("1", "29733")):
        env = dict(os.environ, OMP_NUM_THREADS="1", PTD_HVD_STATIC=static)
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", port, str(script)], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
        assert p.stdout.count("HVD-OK") == 2
        sums[static] = sorted(re.findall(r"HVD-SUM \d (-?\d+\.\d+)", p.stdout))      # ranks may interleave their lines
    assert sums["0"] == sums["1"] and len(sums["0"]) == 2


def test_hvd_autotune_and_timeline_over_gloo(tmp_path):
    """HOROVOD_AUTOTUNE=1 sweeps the cycle budget (same winner on every rank), then the static schedule freezes;
    HOROVOD_TIMELINE writes the fusion-queue trace."""
    import json
    script = tmp_path / "hvd_check.py"
    script.write_text(HVD % ROOT)
    tl = tmp_path / "timeline.json"
    env = dict(os.environ, OMP_NUM_THREADS="1", PTD_HVD_STATIC="1", HOROVOD_AUTOTUNE="1", HOROVOD_AUTOTUNE_STEPS_PER_SAMPLE="2",
               HOROVOD_TIMELINE=str(tl), PTD_TEST_ITERS="24")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29735", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert p.stdout.count("HVD-OK") == 2 and "[hvd autotune]" in p.stdout
    tr = json.load(open(tmp_path / "timeline.rank0.json"))
    assert tr["traceEvents"] and {"name", "ts", "dur", "args"} <= set(tr["traceEvents"][0]) and tr["stats"]["groups"] >= 1


def test_fusion_queue_cycle_budget_and_wait_idle():
    """C++ FusionQueue: groups close at the cycle budget (deterministically), wait_idle blocks without polling, done entries are reaped."""
    import threading
    import time
    from pytorch_distributed_b200 import _ext
    C = _ext.lib()
    q = C.FusionQueue(64 << 20, 5.0, 1000)
    q.enable_timeline(True)
    hs = [q.enqueue("t%d" % i, 400, i) for i in range(7)]      # 400+400+400 >= 1000 closes a group every 3 tensors
    g1, g2 = q.next_group(10.0), q.next_group(10.0)
    assert g1 == hs[0:3] and g2 == hs[3:6] and q.next_group(1.0) == []
    q.flush()
    g3 = q.next_group(10.0)
    assert g3 == hs[6:7] and q.pending() == 7
    t0 = time.time()
    assert q.wait_idle(50.0) is False and time.time() - t0 >= 0.04        # still outstanding: times out

    def finish():
        time.sleep(0.05)
        q.mark_done(g1 + g2 + g3)
    th = threading.Thread(target=finish)
    th.start()
    assert q.wait_idle(5000.0) is True and q.pending() == 0
    th.join()
    tl = q.timeline()
    assert len(tl) == 7 and tl[0][0] == "t0" and tl[0][2] == 1 and tl[3][2] == 2 and tl[6][2] == 3
    q.set_cycle_bytes(0)
    for i in range(5):
        q.enqueue("u%d" % i, 400, i)
    assert q.next_group(1.0) == []                              # budget off: only the 64 MiB threshold or a flush closes a group
    q.shutdown()


def test_apex_namespace_surface():
    from pytorch_distributed_b200 import apex
    from pytorch_distributed_b200.apex import amp
    from pytorch_distributed_b200.apex.parallel import DistributedDataParallel, Reducer
    assert callable(amp.initialize) and callable(amp.scale_loss) and callable(amp.master_params)
    assert callable(amp.state_dict) and callable(amp.load_state_dict)
    assert DistributedDataParallel.__name__ == "DistributedDataParallel" and Reducer is not None and apex.amp is amp
    from pytorch_distributed_b200.apex.optimizers import FusedSGD
    from pytorch_distributed_b200.ops.fused_sgd import FusedSGD as Own
    assert FusedSGD is Own
    import pytest as _pt
    lin = torch.nn.Linear(2, 2)
    for bad in (dict(num_allreduce_streams=2), dict(allreduce_trigger_params=[lin.weight]), dict(shared_param=True)):
        with _pt.raises((NotImplementedError, ValueError)):
            DistributedDataParallel(lin, comm="gloo", **bad)


def test_amp_o2_casts_in_place_and_keeps_fp32_masters_cpu():
    from pytorch_distributed_b200.models import create_model
    from pytorch_distributed_b200.ops.fused_sgd import FusedSGD
    from pytorch_distributed_b200.parallel import amp
    m = create_model("resnet18", num_classes=4)
    opt = FusedSGD(m.parameters(), lr=0.1, momentum=0.9)
    ids = [id(p) for p in m.parameters()]
    m, opt = amp.initialize(m, opt, opt_level="O2", half_dtype=torch.bfloat16, verbosity=0)
    assert [id(p) for p in m.parameters()] == ids                      # same Parameter objects: the optimizer stays valid
    assert all(p.dtype == torch.bfloat16 for p in m.parameters())
    assert m.bn1.running_mean.dtype == torch.float32
    x = torch.randn(2, 3, 32, 32)
    with amp.scale_loss(m(x).float().sum(), opt) as sl:
        sl.backward()
    opt.step()
    masters = list(amp.master_params(opt))
    assert len(masters) == len(ids) and all(t.dtype == torch.float32 for t in masters)
    amp._amp_state.enabled = False
    amp._amp_state.scaler = None


def test_prefetcher_cpu_iter_and_next_and_limit():
    from pytorch_distributed_b200.utils.data import DataPrefetcher, SyntheticLoader, data_prefetcher
    assert data_prefetcher is DataPrefetcher
    loader = SyntheticLoader(4, 5, image_size=8, num_classes=3, pin=False)
    pf = DataPrefetcher(loader, "cpu", dtype=torch.float32, limit=3)
    got = list(pf)
    assert len(pf) == 3 and len(got) == 3 and got[0][0].shape == (4, 3, 8, 8) and got[0][1].dtype == torch.int64
    assert pf.h2d_bytes == 3 * loader.bytes_per_step
    pf2 = DataPrefetcher(SyntheticLoader(2, 2, image_size=8, pin=False, raw_uint8=True), "cpu", normalize="imagenet255")
    a, b = pf2.next()
    assert a.dtype == torch.float32 and abs(float(a.mean())) < 3.0
    pf2.next()
    assert pf2.next() == (None, None)


def test_metric_pipeline_and_train_step_cpu():
    from pytorch_distributed_b200 import cli, driver
    from pytorch_distributed_b200.utils.meters import AverageMeter
    torch.manual_seed(0)
    args = cli.parse_args("distributed", ["--device", "cpu", "--synthetic", "-b", "4"])
    st = driver.Strategy()
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 5))
    st.autocast, st.input_dtype = None, torch.float32
    meters = (AverageMeter("Loss"), AverageMeter("Acc@1"), AverageMeter("Acc@5"))
    mp = driver.MetricPipeline(None, torch.device("cpu"), meters, reduce=False)
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    step = driver.TrainStep(st, model, torch.nn.CrossEntropyLoss(), opt, mp, use_graph=True)   # no CUDA => stays eager
    x, y = torch.randn(4, 3, 2, 2), torch.tensor([0, 1, 2, 3])
    losses = []
    for _ in range(20):
        step(x, y)
        losses.append(meters[0].val)
    mp.drain()
    assert step.graph is None and losses[-1] < losses[0] * 0.5 and meters[2].val == 100.0 and meters[0].count == 80


def test_pretrained_loads_local_torchvision_state_dict(tmp_path, monkeypatch):
    """--pretrained (reference distributed.py:134-136): a torchvision-format state dict found locally loads into the native ResNet."""
    import torch
    import torchvision
    from pytorch_distributed_b200.models import create_model
    tv = torchvision.models.resnet18()
    f = tmp_path / "resnet18-deadbeef.pth"
    torch.save(tv.state_dict(), f)
    monkeypatch.setenv("PTD_PRETRAINED_DIR", str(tmp_path))
    m = create_model("resnet18", pretrained=True)
    for (n1, a), (n2, b) in zip(tv.state_dict().items(), m.state_dict().items()):
        assert n1 == n2 and torch.equal(a, b), n1
    m10 = create_model("resnet18", pretrained=True, num_classes=10)      # other class count: trunk only
    assert torch.equal(m10.conv1.weight, tv.conv1.weight) and m10.fc.weight.shape[0] == 10
    monkeypatch.delenv("PTD_PRETRAINED_DIR")
    monkeypatch.setenv("TORCH_HOME", str(tmp_path / "empty"))
    import pytest
    with pytest.raises(RuntimeError, match="no local weights"):
        create_model("resnet34", pretrained=True)
