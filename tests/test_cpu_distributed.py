"""Distributed-plumbing tier on CPU (gloo, world_size 2): every entrypoint runs a few synthetic iterations
(BASELINE.json config 1), plus gradient parity of the bucket engine against torch DDP."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["-a", "resnet18", "-b", "8", "--synthetic", "--steps-per-epoch", "2", "--epochs", "1", "--image-size", "32",
          "--num-classes", "10", "-p", "1", "--device", "cpu"]


def _env(extra=None):
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": "1", "PYTHONPATH": ROOT})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    env.update(extra or {})
    return env


def _run(cmd, extra_env=None, timeout=600, cwd=ROOT):
    p = subprocess.run(cmd, env=_env(extra_env), cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return p.stdout


def _torchrun(script, n, args, port):
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
            "--master-port", str(port), os.path.join(ROOT, script)] + args


def _check_output(out, world, tmp_path, expect_ckpt=True):
    assert len(re.findall(r"Epoch: \[0\]\[0/2\]\tTime", out)) == world      # every rank prints, like the reference
    assert len(re.findall(r"Test: \[0/2\]\tTime", out)) == world
    assert len(re.findall(r" \* Acc@1 \d+\.\d{3} Acc@5 \d+\.\d{3}", out)) == world
    # reduced metrics are identical on all ranks
    lines = sorted(set(l.split("\tLoss ")[1] for l in out.splitlines() if l.startswith("Epoch: [0][1/2]")))
    assert len(lines) == 1, lines
    if expect_ckpt:
        ck = torch.load(os.path.join(str(tmp_path), "checkpoint.pth.tar"), weights_only=False)
        assert ck["epoch"] == 1 and ck["arch"] == "resnet18" and "best_acc1" in ck
        assert not any(k.startswith("module.") for k in ck["state_dict"])


def test_distributed_py_gloo_world2(tmp_path):
    out = _run(_torchrun("distributed.py", 2, COMMON + ["--checkpoint-dir", str(tmp_path)], 29711))
    _check_output(out, 2, tmp_path)


def test_multiprocessing_distributed_spawn_world2(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "multiprocessing_distributed.py")] + COMMON +
               ["--world-size", "2", "--checkpoint-dir", str(tmp_path)])
    _check_output(out, 2, tmp_path)


def test_apex_distributed_gloo_world2(tmp_path):
    out = _run(_torchrun("apex_distributed.py", 2, COMMON + ["--checkpoint-dir", str(tmp_path), "--opt-level", "O1"], 29713))
    _check_output(out, 2, tmp_path)


def test_horovod_distributed_selfspawn_world2(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "horovod_distributed.py")] + COMMON +
               ["--world-size", "2", "--checkpoint-dir", str(tmp_path)])
    _check_output(out, 2, tmp_path)


def test_slurm_entrypoint_fake_env(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "distributed_slurm_main.py")] + COMMON +
               ["--world-size", "2", "--dist-file", str(tmp_path / "rdzv"), "--checkpoint-dir", str(tmp_path)],
               extra_env={"SLURM_PROCID": "0", "SLURM_NPROCS": "1", "SLURM_JOBID": "77"})
    _check_output(out, 2, tmp_path)
    assert os.path.exists(tmp_path / "distributed.csv")


def test_slurm_two_nodes_global_rank(tmp_path):
    """Two fake Slurm tasks (one 'node' each, one worker per node): the process group must be built from the GLOBAL rank
    (node_rank * ngpus + gpu), not the local one - with local ranks both tasks would register as rank 0 and hang."""
    procs = []
    for node in (0, 1):
        env = _env({"SLURM_PROCID": str(node), "SLURM_NPROCS": "2", "SLURM_JOBID": "78"})
        d = tmp_path / ("node%d" % node)
        d.mkdir()
        cmd = [sys.executable, os.path.join(ROOT, "distributed_slurm_main.py")] + COMMON + \
              ["--world-size", "1", "--dist-file", str(tmp_path / "rdzv2"), "--checkpoint-dir", str(d)]
        procs.append(subprocess.Popen(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, o[-2000:] + "\n" + e[-3000:]
        outs.append(o)
    _check_output("\n".join(outs), 2, tmp_path / "node0")
    assert not os.path.exists(tmp_path / "node1" / "checkpoint.pth.tar")      # only global rank 0 saves (Q11)


def test_dataparallel_cpu_passthrough(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "dataparallel.py")] + COMMON + ["--checkpoint-dir", str(tmp_path)])
    assert " * Acc@1" in out and os.path.exists(tmp_path / "dataparallel.csv")
    assert os.path.exists(tmp_path / "checkpoint.pth.tar")


def test_evaluate_flag_skips_training(tmp_path):
    out = _run(_torchrun("distributed.py", 1, COMMON + ["-e", "--checkpoint-dir", str(tmp_path)], 29717))
    assert "Epoch:" not in out and " * Acc@1" in out and not os.path.exists(tmp_path / "checkpoint.pth.tar")


def test_resume_roundtrip(tmp_path):
    _run(_torchrun("distributed.py", 1, COMMON + ["--checkpoint-dir", str(tmp_path)], 29718))
    out = _run(_torchrun("distributed.py", 1, COMMON + ["--checkpoint-dir", str(tmp_path), "--epochs", "2",
                                                        "--resume", str(tmp_path / "checkpoint.pth.tar")], 29719))
    assert "=> loaded checkpoint" in out and "Epoch: [1]" in out and "Epoch: [0]" not in out


PARITY = r'''
import copy, os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
from pytorch_distributed_b200.models import create_model
from pytorch_distributed_b200.parallel.ddp import DistributedDataParallel
from pytorch_distributed_b200.ops.fused_sgd import FusedSGD
torch.manual_seed(3 + rank)          # different init per rank: the ctor must broadcast rank 0's weights
base = create_model("resnet18", num_classes=10)
own = DistributedDataParallel(copy.deepcopy(base), comm="gloo", bucket_cap_mb=2.0)
ref = torch.nn.parallel.DistributedDataParallel(copy.deepcopy(own.module))
assert len(own.engine.buckets) > 3
crit = torch.nn.CrossEntropyLoss()
oo = FusedSGD(own.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
orf = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
for it in range(2):
    x = torch.randn(4, 3, 32, 32); y = torch.randint(0, 10, (4,))
    for m, o in ((own, oo), (ref, orf)):
        o.zero_grad(); crit(m(x), y).backward()
    for (n, a), b in zip(own.module.named_parameters(), ref.module.parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-6), (it, n, (a.grad - b.grad).abs().max())
    oo.step(); orf.step()
    with torch.no_grad():
        for a, b in zip(ref.module.parameters(), own.module.parameters()): a.copy_(b)
        for a, b in zip(ref.module.buffers(), own.module.buffers()): a.copy_(b)
flat = torch.cat([p.detach().reshape(-1) for p in own.parameters()])
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi)
with own.no_sync():
    crit(own(torch.randn(2, 3, 32, 32)), torch.randint(0, 10, (2,))).backward()
print("PARITY-OK", rank)
dist.destroy_process_group()
'''


def test_bucket_engine_gradients_match_torch_ddp_gloo(tmp_path):
    script = tmp_path / "parity.py"
    script.write_text(PARITY % ROOT)
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29721", str(script)])
    assert out.count("PARITY-OK") == 2


def test_reference_arm_reports_unavailable_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = _run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert '"impl": "reference"' in out and '"unavailable"' in out
