"""Failure handling: a rank that dies mid-epoch must take the whole job down quickly, with a diagnostic, instead of
leaving the survivors blocked in a collective (reference launcher: /root/reference/multiprocessing_distributed.py:110-135,
``mp.spawn(..., join=True)``)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(extra):
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": "1", "PYTHONPATH": ROOT})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    env.update(extra)
    return env


def test_killed_rank_aborts_the_job_quickly(tmp_path):
    cmd = [sys.executable, os.path.join(ROOT, "multiprocessing_distributed.py"), "-a", "resnet18", "-b", "8", "--synthetic",
           "--steps-per-epoch", "200", "--epochs", "1", "--image-size", "32", "--num-classes", "10", "-p", "1", "--device", "cpu",
           "--world-size", "2", "--checkpoint-dir", str(tmp_path)]
    t0 = time.time()
    p = subprocess.run(cmd, env=_env({"PTD_TEST_KILL_RANK": "1", "PTD_TEST_KILL_STEP": "3"}), cwd=ROOT, capture_output=True, text=True,
                       timeout=300)
    dt = time.time() - t0
    assert p.returncode != 0, "the job must fail when a rank dies"
    assert dt < 120, "tear-down took %.0f s" % dt
    err = p.stderr + p.stdout
    assert "SIGKILL" in err or "signal 9" in err.lower() or "terminated" in err, err[-2000:]
    assert not os.path.exists(tmp_path / "checkpoint.pth.tar")
