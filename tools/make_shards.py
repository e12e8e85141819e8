#!/usr/bin/env python
"""Decode an ImageFolder dataset (DIR/train, DIR/val) once into pre-decoded ``*.ptds`` shards for the native loader.

    python tools/make_shards.py /data/imagenet /data/imagenet-ptds --max-side 256 --workers 32
    torchrun --nproc-per-node 8 distributed.py --data /data/imagenet-ptds -a resnet50 -b 2048 -j 12

Images keep their aspect ratio (shorter side reduced to ``--max-side``), so RandomResizedCrop still samples from the
whole picture.  ImageNet train at max-side 256 is ~330 GB of raw pixels: it is read through mmap / the page cache.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pytorch_distributed_b200.utils import shards  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("src", help="ImageFolder root with train/ and val/")
    ap.add_argument("dst", help="output directory for the shards")
    ap.add_argument("--splits", default="train,val")
    ap.add_argument("--max-side", type=int, default=256, help="cap of the shorter image side (0 = keep the original size)")
    ap.add_argument("--per-shard", type=int, default=4096, help="records per shard file")
    ap.add_argument("--workers", type=int, default=0, help="decoder processes (0 = in-process)")
    a = ap.parse_args()
    for split in a.splits.split(","):
        d = os.path.join(a.src, split)
        if not os.path.isdir(d):
            print("skip %s: %s does not exist" % (split, d))
            continue
        t0 = time.time()
        paths = shards.write_shards(d, a.dst, split, a.max_side, a.per_shard, a.workers, log=print)
        n = sum(len(shards.read_index(p)) for p in paths)
        size = sum(os.path.getsize(p) for p in paths)
        print("%s: %d images -> %d shards, %.1f MB, %.1f s" % (split, n, len(paths), size / 1e6, time.time() - t0))


if __name__ == "__main__":
    main()
