"""Single-GPU runs of the entrypoints (what a 1-GPU box can exercise): eager and CUDA-graph training steps,
checkpoint contract, evaluation-only mode."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
COMMON = ["-a", "resnet18", "-b", "32", "--synthetic", "--steps-per-epoch", "8", "--val-steps", "2", "--epochs", "1", "--image-size", "64",
          "-p", "1", "--lr", "0.01"]


def _run(cmd, timeout=900):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = "0"
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-5000:]
    return p.stdout


def _losses(out):
    return [float(x) for x in re.findall(r"Epoch: \[0\]\[\d+/8\].*?Loss (\d\.\d+e[+-]\d+)", out)]


@pytest.mark.parametrize("graph", [False, True])
def test_distributed_py_single_gpu(tmp_path, graph):
    args = COMMON + ["--checkpoint-dir", str(tmp_path), "--seed", "1"] + (["--cuda-graph"] if graph else [])
    out = _run([sys.executable, os.path.join(ROOT, "distributed.py")] + args)
    ls = _losses(out)
    assert len(ls) == 8 and all(v == v and v < 100 for v in ls), ls
    assert " * Acc@1" in out
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    assert ck["epoch"] == 1 and all(v.dtype == torch.float32 for v in ck["state_dict"].values() if v.is_floating_point())


def test_cuda_graph_matches_eager_losses(tmp_path):
    """Same seed, same synthetic data: the captured step must reproduce the eager trajectory (bf16 tolerance)."""
    outs = []
    for graph in (False, True):
        args = COMMON + ["--checkpoint-dir", str(tmp_path), "--seed", "3", "--no-fused-bn"] + (["--cuda-graph"] if graph else [])
        outs.append(_losses(_run([sys.executable, os.path.join(ROOT, "distributed.py")] + args)))
    assert len(outs[0]) == len(outs[1]) == 8
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert abs(a - b) <= 0.05 * max(1.0, abs(a)), (outs[0], outs[1])


def test_apex_o2_single_gpu(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "apex_distributed.py")] + COMMON + ["--opt-level", "O2", "--cuda-graph",
                                                                                   "--checkpoint-dir", str(tmp_path)])
    ls = _losses(out)
    assert len(ls) == 8 and all(v == v and v < 100 for v in ls), ls


def test_dataparallel_single_gpu(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "dataparallel.py")] + COMMON + ["--gpus", "0", "--checkpoint-dir", str(tmp_path)])
    assert " * Acc@1" in out and os.path.exists(tmp_path / "dataparallel.csv")


def test_dataparallel_single_gpu_applies_gradients():
    """DataParallel over ONE device with the flat FusedSGD: the optimizer reads the gradient arena, which only the
    end-of-backward pack fills - weights must move and the loss on a fixed batch must go down."""
    from pytorch_distributed_b200.models import create_model
    from pytorch_distributed_b200.ops.fused_sgd import FusedSGD
    from pytorch_distributed_b200.parallel.dp import DataParallel
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = create_model("resnet18", num_classes=10).to(dev).to(memory_format=torch.channels_last)
    dp = DataParallel(model, device_ids=[0])
    opt = FusedSGD(dp.parameters(), lr=0.05, momentum=0.9)
    assert opt.is_flat
    x = torch.randn(16, 3, 64, 64, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (16,), device=dev)
    w0 = model.fc.weight.detach().clone()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(dp(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert not torch.equal(w0, model.fc.weight.detach()), "weights did not move"
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize("flags", [["--bucket-view"], ["--no-overlap-optimizer"], ["--bucket-view", "--cuda-graph"]])
def test_distributed_py_engine_modes_match_default(tmp_path, flags):
    """gradient_as_bucket_view (in-place accumulation into the arena, K1 without a pack pass) and the non-overlapped optimizer
    must reproduce the default trajectory (same seed, same data)."""
    outs = []
    for extra in ([], flags):
        args = COMMON + ["--checkpoint-dir", str(tmp_path), "--seed", "5", "--no-fused-bn"] + extra
        outs.append(_losses(_run([sys.executable, os.path.join(ROOT, "distributed.py")] + args)))
    assert len(outs[0]) == len(outs[1]) == 8
    for a, b in zip(outs[0][:5], outs[1][:5]):
        assert abs(a - b) <= 0.05 * max(1.0, abs(a)), (outs[0], outs[1])


def test_evaluate_only(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "distributed.py")] + COMMON + ["-e", "--checkpoint-dir", str(tmp_path)])
    assert "Epoch:" not in out and " * Acc@1" in out
