"""Horovod-style entrypoint (reference: /root/reference/horovod_distributed.py, start.sh:4 `horovodrun -np 4 ...`).

There is no MPI / horovodrun on the target image, so the script accepts either launcher:
    python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 horovod_distributed.py -a resnet50 --synthetic
    python horovod_distributed.py --world-size 8 -a resnet50 --synthetic          (self-spawn)
The ``hvd`` API (init / broadcast_parameters / broadcast_optimizer_state / DistributedOptimizer with fp16
compression) is pytorch_distributed_b200.parallel.hvd: a C++ fusion queue + the fused NVLink all-reduce.
"""
import os

from pytorch_distributed_b200 import cli, driver, launch


def worker(local_rank, nprocs, args):
    driver.seed_everything(args)
    driver.main_worker(local_rank, nprocs, args)


def main():
    args = cli.parse_args("horovod_distributed")
    env = launch.torchrun_env()
    if env is not None:
        args.nprocs = env[2]
        worker(env[1], env[2], args)
        return
    args.nprocs = launch.default_nprocs(args)
    port = launch.pick_port(launch.DEFAULT_PORT + 1)
    envs = {"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "WORLD_SIZE": str(args.nprocs)}
    launch.spawn(_HvdSpawn(), args.nprocs, args, extra_env=envs)


class _HvdSpawn:
    """Picklable spawn target: exports RANK before hvd.init() reads the environment."""

    def __call__(self, i, n, a):
        os.environ["RANK"] = str(i)
        worker(i, n, a)


if __name__ == "__main__":
    main()
