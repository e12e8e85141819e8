"""Cross-GPU kernel tests (K1-K4, DDP parity): spawn torchrun over the visible GPUs."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _run(script, nproc, extra_env=None, args=(), timeout=900):
    env = dict(os.environ)
    env.update(extra_env or {})
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)] + list(args)
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-4000:] + "\n" + p.stderr[-4000:]
    return p.stdout


@pytest.mark.parametrize("nvls", ["1", "0"])
def test_collectives_and_ddp_parity(nvls):
    n = min(torch.cuda.device_count(), 8)
    out = _run("tests/mp_gpu_checks.py", n, {"PTD_NVLS": nvls})
    assert out.count("PASS rank") == n, out[-3000:]


def test_distributed_entrypoint_two_gpus(tmp_path):
    out = _run("distributed.py", 2, args=["-a", "resnet18", "-b", "32", "--synthetic", "--steps-per-epoch", "4", "--epochs", "1",
                                          "--image-size", "64", "-p", "1", "--checkpoint-dir", str(tmp_path)])
    assert " * Acc@1" in out
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    assert ck["epoch"] == 1 and ck["arch"] == "resnet18"
    assert all(v.dtype == torch.float32 for v in ck["state_dict"].values() if v.is_floating_point())
