"""The real-data path (ImageFolder + DistributedSampler + DataLoader workers) on a tiny generated dataset."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_tree(root, classes=2, per_class=6, size=40):
    from PIL import Image
    import numpy as np
    rng = np.random.default_rng(0)
    for split in ("train", "val"):
        for c in range(classes):
            d = os.path.join(root, split, "class%d" % c)
            os.makedirs(d, exist_ok=True)
            for i in range(per_class):
                arr = (rng.random((size, size, 3)) * 255).astype("uint8")
                arr[..., c % 3] = 255 if split == "train" else arr[..., c % 3]
                Image.fromarray(arr).save(os.path.join(d, "img%d.png" % i))


@pytest.mark.parametrize("entry,extra", [("distributed.py", []), ("apex_distributed.py", ["--opt-level", "O1"])])
def test_imagefolder_pipeline_gloo_world2(tmp_path, entry, extra):
    pytest.importorskip("PIL")
    data = tmp_path / "data"
    _make_tree(str(data))
    env = dict(os.environ, OMP_NUM_THREADS="1", PYTHONPATH=ROOT)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", os.path.join(ROOT, entry), "--data", str(data), "-a", "resnet18", "-b", "4", "--epochs", "1",
           "--image-size", "32", "--num-classes", "2", "-j", "1", "-p", "1", "--device", "cpu", "--checkpoint-dir", str(tmp_path)] + extra
    p = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    # 12 train images / (2 ranks * batch 2) = 3 iterations per rank
    assert p.stdout.count("Epoch: [0][0/3]") == 2 and p.stdout.count(" * Acc@1") == 2
    assert os.path.exists(tmp_path / "checkpoint.pth.tar")
