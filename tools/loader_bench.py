#!/usr/bin/env python
"""Host-side throughput of the native shard loader vs the PIL/torchvision transform chain on the same pixels.

    python tools/loader_bench.py --records 2048 --threads 1,2,4,8

Synthetic records (short side 256, 4:3) so that only the transform cost is measured; JPEG decoding - which the
ImageFolder path pays on every sample and the shard path paid once offline - is reported separately.
"""
import argparse
import io
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pytorch_distributed_b200.utils import shards  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--threads", default="1,2,4,8")
    ap.add_argument("--epochs", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    tmp = tempfile.mkdtemp(prefix="ptds_bench_")
    base = rng.integers(0, 256, (64, 86, 3), dtype=np.uint8)
    from PIL import Image
    proto = np.asarray(Image.fromarray(base).resize((341, 256), Image.BICUBIC))      # smooth, photo-like spectrum
    path = os.path.join(tmp, "train-00000.ptds")
    with shards.ShardWriter(path, a.records) as w:
        for i in range(a.records):
            w.add(np.roll(proto, i, axis=1), i % 1000)
    print("| pipeline | threads / procs | images/s | per core |")
    print("|---|---:|---:|---:|")
    for t in [int(x) for x in a.threads.split(",")]:
        for train in (True, False):
            ld = shards.ShardLoader([path], a.batch, a.size, train=train, workers=t, depth=4, pin=False)
            n = 0
            for _ in ld:                      # warm-up epoch (page cache, thread start)
                pass
            t0 = time.perf_counter()
            for e in range(a.epochs):
                ld.sampler.set_epoch(e + 1)
                for x, y in ld:
                    n += x.shape[0]
            dt = time.perf_counter() - t0
            print("| native shards, %s | %d | %.0f | %.0f |" % ("train (RRC + flip)" if train else "val (resize + centre crop)", t, n / dt,
                                                               n / dt / t))
    # the reference's per-sample work on the same pixels, one process: PIL transforms (+ JPEG decode, quality 90)
    import torchvision.transforms as T
    tf = T.Compose([T.RandomResizedCrop(a.size), T.RandomHorizontalFlip(), T.PILToTensor()])
    img = Image.fromarray(proto)
    buf = io.BytesIO()
    img.save(buf, format="JPEG", quality=90)
    jpeg = buf.getvalue()
    torch.set_num_threads(1)
    for name, fn in (("PIL transforms only", lambda: tf(img)),
                     ("PIL JPEG decode + transforms (ImageFolder path)", lambda: tf(Image.open(io.BytesIO(jpeg)).convert("RGB")))):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        k = 400
        for _ in range(k):
            fn()
        dt = time.perf_counter() - t0
        print("| %s | 1 | %.0f | %.0f |" % (name, k / dt, k / dt))


if __name__ == "__main__":
    main()
