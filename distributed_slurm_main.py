"""Multi-node DDP under Slurm (reference: /root/reference/distributed_slurm_main.py, start.sh:5).

    srun -N2 --gres gpu:8 python distributed_slurm_main.py --dist-file dist_file -a resnet50 --synthetic

One task per node; each task ``mp.spawn``s one worker per local GPU; global rank = SLURM_PROCID * ngpus + gpu;
``file://<dist_file>.<SLURM_JOBID>`` rendezvous.  Within a node the fused NVLink data plane is used; across nodes the
gradient engine falls back to the library collectives (``--comm nccl``) because peer mappings do not span nodes.
"""
import os

import torch

from pytorch_distributed_b200 import cli, driver, launch


class _Worker:
    def __init__(self, node_rank, ngpus, world, url):
        self.node_rank, self.ngpus, self.world, self.url = node_rank, ngpus, world, url

    def __call__(self, gpu, nprocs, args):
        rank = self.node_rank * self.ngpus + gpu
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(self.world), LOCAL_RANK=str(gpu))
        args.dist_url = self.url
        args.global_rank = rank
        if self.world > self.ngpus and args.comm == "auto":
            args.comm = "nccl" if torch.cuda.is_available() else "gloo"
        driver.seed_everything(args)
        driver.main_worker(gpu, self.world, args, driver.SlurmStrategy())


def main():
    args = cli.parse_args("distributed_slurm_main")
    ngpus = launch.default_nprocs(args)
    node_rank, n_nodes, world, url = launch.slurm_topology(args, ngpus)
    args.nprocs = world
    launch.spawn(_Worker(node_rank, ngpus, world, url), ngpus, args)


if __name__ == "__main__":
    main()
