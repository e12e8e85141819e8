"""Native shard loader (csrc/host/loader.cpp + utils/shards.py): format, resampling numerics against a PyTorch fp32
reference, sampler semantics, crop statistics, ring-buffer recycling, and the entrypoint path on shards."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pytorch_distributed_b200 import _hostext
from pytorch_distributed_b200.utils import shards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(tmp_path, n=37, split="train", per_shard=16, sizes=((40, 56), (64, 48), (33, 33)), seed=0):
    """Random-noise records of three shapes in shards of ``per_shard``; label = index % 5."""
    rng = np.random.default_rng(seed)
    paths, w, images = [], None, []
    for i in range(n):
        if w is None or w.full:
            if w is not None:
                w.close()
            paths.append(str(tmp_path / ("%s-%05d.ptds" % (split, len(paths)))))
            w = shards.ShardWriter(paths[-1], min(per_shard, n - i))
        h, wd = sizes[i % len(sizes)]
        img = rng.integers(0, 256, (h, wd, 3), dtype=np.uint8)
        images.append(img)
        w.add(img, i % 5)
    w.close()
    return paths, images


def _ref_resample(img_hwc, box, out, flip=False):
    """fp32 reference: crop (integer box) then antialiased bilinear resize, as torchvision's crop -> resize."""
    x0, y0, bw, bh = box
    t = torch.from_numpy(img_hwc[y0:y0 + bh, x0:x0 + bw]).permute(2, 0, 1)[None].float()
    r = F.interpolate(t, size=(out, out), mode="bilinear", antialias=True, align_corners=False)[0]
    if flip:
        r = r.flip(-1)
    return r


def test_shard_format_roundtrip(tmp_path):
    paths, images = _write(tmp_path)
    assert len(paths) == 3
    idx = [e for p in paths for e in shards.read_index(p)]
    assert [(e[1], e[2]) for e in idx] == [im.shape[:2] for im in images]
    assert [e[3] for e in idx] == [i % 5 for i in range(len(images))]
    with open(paths[0], "rb") as f:
        raw = f.read()
    off, h, w, _ = idx[0]
    assert np.array_equal(np.frombuffer(raw[off:off + h * w * 3], np.uint8).reshape(h, w, 3), images[0])
    with pytest.raises(RuntimeError):
        _hostext.lib().ShardLoader([paths[0] + ".missing"], 4, 8, 8, True, 0, 0, 1, 1, 3, False, True)


@pytest.mark.parametrize("box,out", [((0, 0, 56, 40), 24), ((5, 3, 30, 30), 30), ((7, 2, 20, 28), 64), ((0, 0, 56, 40), 7)])
def test_resample_matches_fp32_reference(box, out):
    L = _hostext.lib()
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    src = torch.from_numpy(img).contiguous()
    for flip in (False, True):
        dst = torch.zeros(3, out, out, dtype=torch.uint8)
        L.resample(src.data_ptr(), 40, 56, float(box[0]), float(box[1]), float(box[2]), float(box[3]), True, flip, dst.data_ptr(), out, out)
        ref = _ref_resample(img, box, out, flip)
        err = (dst.float() - ref).abs().max().item()
        assert err <= 0.51, (box, out, flip, err)       # rounding to uint8 only


def test_identity_box_is_exact():
    L = _hostext.lib()
    img = np.random.default_rng(2).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    src = torch.from_numpy(img)
    dst = torch.zeros(3, 32, 32, dtype=torch.uint8)
    L.resample(src.data_ptr(), 32, 32, 0.0, 0.0, 32.0, 32.0, True, False, dst.data_ptr(), 32, 32)
    assert torch.equal(dst, src.permute(2, 0, 1))


def test_center_crop_matches_resize_then_crop(tmp_path):
    """val transform = Resize(out * 256 / 224) + CenterCrop(out) (/root/reference/distributed.py:183-188)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    # smooth the noise: the two orders of (resize, crop) agree up to where the filter footprint is sampled
    t = torch.from_numpy(img).permute(2, 0, 1)[None].float()
    t = F.avg_pool2d(t, 5, 1, 2).round().clamp(0, 255)
    img = t[0].permute(1, 2, 0).to(torch.uint8).numpy()
    with shards.ShardWriter(str(tmp_path / "val-00000.ptds"), 1) as w:
        w.add(img, 3)
    out = 56
    ld = shards.ShardLoader([str(tmp_path / "val-00000.ptds")], 1, out, train=False, workers=1, pin=False)
    (x, y), = list(ld)
    assert y.tolist() == [3]
    short = 64                                            # 56 * 256 / 224
    r = F.interpolate(t, size=(short, int(128 * short / 96)), mode="bilinear", antialias=True, align_corners=False)[0]
    top, left = round((r.shape[1] - out) / 2.0), round((r.shape[2] - out) / 2.0)      # torchvision CenterCrop geometry
    ref = r[:, top:top + out, left:left + out]
    assert (x[0].float() - ref).abs().max().item() <= 0.51


def test_sampler_covers_dataset_once_per_epoch(tmp_path):
    paths, images = _write(tmp_path, n=37)
    world, seen = 2, []
    orders = []
    for rank in range(world):
        ld = shards.ShardLoader(paths, 8, 16, train=True, seed=7, rank=rank, world=world, workers=3, pin=False, with_ids=True)
        assert ld.num_records == 37 and len(ld) == 3       # ceil(37 / 2) = 19 samples -> 3 batches of <= 8
        ids = []
        for x, y in ld:
            assert x.shape[1:] == (3, 16, 16) and x.dtype == torch.uint8 and y.dtype == torch.int64
            assert torch.equal(y, ld.last_ids % 5)
            ids += ld.last_ids.tolist()
        assert len(ids) == 19
        assert ids == ld._L.epoch_order(0)
        orders.append(ids)
        seen += ids
    assert sorted(set(seen)) == list(range(37)) and len(seen) == 38          # padded by one wrapped sample
    ld.sampler.set_epoch(1)
    ids1 = [i for _ in ld for i in ld.last_ids.tolist()]
    assert ids1 != orders[1] and len(ids1) == 19
    ld.sampler.set_epoch(0)                                                 # deterministic in (seed, epoch)
    assert [i for _ in ld for i in ld.last_ids.tolist()] == orders[1]


def test_drop_last_gives_full_batches_only(tmp_path):
    paths, _ = _write(tmp_path, n=37)
    ld = shards.ShardLoader(paths, 8, 16, train=True, seed=7, rank=1, world=2, workers=2, pin=False, drop_last=True)
    sizes = [x.shape[0] for x, _ in ld]
    assert sizes == [8, 8] and len(ld) == 2               # 37 // 2 = 18 samples -> two full batches


def test_batches_do_not_depend_on_thread_count(tmp_path):
    paths, _ = _write(tmp_path, n=23)
    outs = []
    for workers in (1, 5):
        ld = shards.ShardLoader(paths, 4, 20, train=True, seed=11, workers=workers, pin=False)
        outs.append(torch.cat([x.clone() for x, _ in ld]))
    assert torch.equal(outs[0], outs[1])


def test_random_resized_crop_statistics(tmp_path):
    paths, _ = _write(tmp_path, n=4)
    ld = shards.ShardLoader(paths, 2, 16, train=True, seed=5, workers=1, pin=False)
    W, H = 500, 375
    p = np.array([ld._L.crop_params(0, i, W, H) for i in range(4000)])
    x0, y0, w, h, flip = p.T
    assert (x0 >= 0).all() and (y0 >= 0).all() and (x0 + w <= W).all() and (y0 + h <= H).all()
    frac, ratio = w * h / (W * H), w / h
    from torchvision.transforms import RandomResizedCrop                              # same distribution as torchvision's
    torch.manual_seed(0)
    tv = [RandomResizedCrop.get_params(torch.zeros(3, H, W), (0.08, 1.0), (3 / 4, 4 / 3)) for _ in range(4000)]
    tv_frac = np.array([th * tw / (W * H) for _, _, th, tw in tv])
    tv_lr = np.array([np.log(tw / th) for _, _, th, tw in tv])
    assert frac.min() >= 0.079 and frac.max() <= 1.0 and abs(frac.mean() - tv_frac.mean()) < 0.02
    assert abs(np.percentile(frac, 25) - np.percentile(tv_frac, 25)) < 0.03
    assert abs(np.log(ratio).mean() - tv_lr.mean()) < 0.02 and abs(np.log(ratio).std() - tv_lr.std()) < 0.02
    assert ratio.min() >= 0.74 and ratio.max() <= 1.34
    assert 0.45 < flip.mean() < 0.55
    assert len({tuple(r) for r in p[:, :4]}) > 3900


def test_ring_slots_are_recycled_after_two_batches(tmp_path):
    paths, _ = _write(tmp_path, n=40)
    ld = shards.ShardLoader(paths, 4, 12, train=False, workers=2, depth=3, pin=False, with_ids=True)
    held = []
    for x, y in ld:
        held.append((x, x.clone()))
        if len(held) >= 2:                       # the previous batch must still be intact while the next one is drawn
            assert torch.equal(held[-2][0], held[-2][1])
    assert len(held) == 10


def test_prefetcher_normalises_shard_batches(tmp_path):
    from pytorch_distributed_b200.utils.data import IMAGENET_MEAN, IMAGENET_STD, DataPrefetcher
    paths, _ = _write(tmp_path, n=9)
    ld = shards.ShardLoader(paths, 4, 16, train=False, workers=2, pin=False)
    raw = torch.cat([x.clone() for x, _ in ld]).float()
    pf = DataPrefetcher(ld, "cpu", dtype=torch.float32, normalize="imagenet255")
    got = torch.cat([x for x, _ in pf])
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    assert torch.allclose(got, (raw / 255.0 - mean) / std, atol=1e-5)


def test_writer_skips_unreadable_and_converts_modes(tmp_path):
    from PIL import Image
    d = tmp_path / "raw" / "train" / "a"
    d.mkdir(parents=True)
    Image.fromarray(np.full((20, 30), 77, dtype=np.uint8), mode="L").save(str(d / "grey.png"))
    Image.fromarray(np.zeros((12, 12, 3), dtype=np.uint8)).save(str(d / "rgb.png"))
    (d / "broken.png").write_bytes(b"not an image")
    msgs = []
    paths = shards.write_shards(str(tmp_path / "raw" / "train"), str(tmp_path / "out"), "train", max_side=16, log=msgs.append)
    idx = shards.read_index(paths[0])
    assert len(idx) == 2 and any("broken.png" in m for m in msgs)
    assert sorted((e[1], e[2]) for e in idx) == [(12, 12), (16, 24)]          # grey 20x30 -> RGB, short side capped at 16


def test_entrypoint_trains_from_shards(tmp_path):
    """ImageFolder -> tools/make_shards.py -> distributed.py --data <shards> under torchrun (gloo, world 2)."""
    from PIL import Image
    rng = np.random.default_rng(0)
    for split, n in (("train", 6), ("val", 4)):
        for c in ("cat", "dog"):
            d = tmp_path / "raw" / split / c
            d.mkdir(parents=True)
            for i in range(n):
                Image.fromarray(rng.integers(0, 256, (48 + 8 * i, 64, 3), dtype=np.uint8)).save(str(d / ("%d.png" % i)))
    out = tmp_path / "shards"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_shards.py"), str(tmp_path / "raw"), str(out),
                        "--max-side", "40", "--per-shard", "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert len(shards.find_shards(str(out), "train")) == 3 and len(shards.find_shards(str(out), "val")) == 2
    assert min(e[1:3] for e in shards.read_index(shards.find_shards(str(out), "train")[0]))[0] <= 48
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "distributed.py"), "--data", str(out), "-a", "resnet18", "-b", "4",
           "--epochs", "1", "--image-size", "32", "--num-classes", "2", "-j", "2", "--device", "cpu", "-p", "1",
           "--checkpoint-dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("Epoch: [0][0/3]") == 2          # 12 train images / 2 ranks / batch 2
    assert r.stdout.count(" * Acc@1") == 2
    assert (tmp_path / "checkpoint.pth.tar").exists()


def test_abandoned_epoch_and_restart(tmp_path):
    paths, _ = _write(tmp_path, n=64)
    ld = shards.ShardLoader(paths, 4, 12, train=True, seed=3, workers=3, depth=3, pin=False, with_ids=True)
    first = []
    for i, (x, y) in enumerate(ld):
        first.append(ld.last_ids.tolist())
        if i == 2:
            break                                   # consumer walks away mid-epoch (e.g. --steps-per-epoch)
    again = [ld.last_ids.tolist() for _ in ld]      # same epoch number -> same order, from the start
    assert again[:3] == first and len(again) == 16
    ld.close()


def test_corrupt_shard_is_rejected(tmp_path):
    paths, _ = _write(tmp_path, n=5)
    bad = tmp_path / "train-99999.ptds"
    data = bytearray(open(paths[0], "rb").read())
    data[:8] = b"NOTASHRD"
    bad.write_bytes(bytes(data))
    with pytest.raises(RuntimeError, match="not a PTDSHRD1 shard"):
        shards.ShardLoader([str(bad)], 2, 8, pin=False)
    trunc = tmp_path / "train-99998.ptds"
    trunc.write_bytes(open(paths[0], "rb").read()[:200])
    with pytest.raises(RuntimeError, match="corrupt shard record|truncated shard index"):
        shards.ShardLoader([str(trunc)], 2, 8, pin=False)
