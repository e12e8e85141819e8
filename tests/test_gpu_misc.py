"""Remaining native surfaces: multi_tensor_axpby, LL scalar all-reduce at world 1, p2p copy on one device, NVTX/poison switches."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def C():
    from pytorch_distributed_b200 import _ext
    return _ext.lib()


def test_multi_tensor_axpby():
    x = [torch.randn(1000, device="cuda"), torch.randn(7, 9, device="cuda").half()]
    y = [torch.randn(1000, device="cuda").bfloat16(), torch.randn(7, 9, device="cuda")]
    out = [torch.empty(1000, device="cuda"), torch.empty(7, 9, device="cuda")]
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    C().multi_tensor_axpby(x, y, out, 0.5, -2.0, flag)
    for a, b, o in zip(x, y, out):
        torch.testing.assert_close(o, 0.5 * a.float() - 2.0 * b.float(), rtol=1e-6, atol=1e-6)
    assert flag.item() == 0
    y[0][3] = float("nan")
    C().multi_tensor_axpby(x, y, out, 1.0, 1.0, flag)
    assert flag.item() == 1


def test_p2p_copy_multi_same_device():
    src = [torch.randn(1000, device="cuda"), torch.randn(33, 7, device="cuda").bfloat16(), torch.arange(5, device="cuda", dtype=torch.float32)]
    dst = [torch.empty_like(s) for s in src]
    C().p2p_copy_multi(src, dst, 0)
    for s, d in zip(src, dst):
        assert torch.equal(s, d)


def test_nvtx_and_poison_switches_run(tmp_path):
    env = dict(os.environ, PTD_NVTX="1", PTD_DEBUG_POISON="1", CUDA_VISIBLE_DEVICES="0")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "distributed.py"), "-a", "resnet18", "-b", "16", "--synthetic", "--steps-per-epoch", "4",
                        "--val-steps", "1", "--epochs", "1", "--image-size", "64", "-p", "1", "--lr", "0.01", "--checkpoint-dir", str(tmp_path)],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    import re
    vals = [float(x) for x in re.findall(r"Loss (\d\.\d+e[+-]\d+)", p.stdout)]
    assert vals and all(v == v for v in vals)      # poisoning the consumed arena must not leak NaNs into training
