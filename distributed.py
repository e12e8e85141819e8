"""DistributedDataParallel entrypoint (reference: /root/reference/distributed.py, launched by start.sh:2).

    python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 distributed.py -a resnet50 -b 2048 --synthetic

One process per GPU; gradients are all-reduced during backward by the fused NVLink kernel; this script is also the
"distributed evaluation" demo (sharded validation + one-kernel metric all-reduce).
"""
from pytorch_distributed_b200 import cli, driver, launch


def main():
    args = cli.parse_args("distributed")
    env = launch.torchrun_env()
    args.nprocs = env[2] if env else 1            # world size from the launcher, not device_count()
    local_rank = cli.resolve_local_rank(args)
    driver.seed_everything(args)
    if env is None and not args.dist_url:
        args.dist_url = launch.tcp_url()           # plain `python distributed.py` => single-process group
    driver.main_worker(local_rank, args.nprocs, args)


if __name__ == "__main__":
    main()
