"""Mixed-precision DDP entrypoint (reference: /root/reference/apex_distributed.py, start.sh:3).

    python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 apex_distributed.py -a resnet50 -b 2048 \
        --opt-level O2 --synthetic

``amp.initialize`` + apex-style ``DistributedDataParallel`` + ``amp.scale_loss`` + the CUDA-stream prefetcher, all
re-implemented natively (pytorch_distributed_b200.apex): dynamic loss scaling, overflow skip and the SGD update are
one fused kernel; the non-finite test is folded into the gradient all-reduce.
"""
from pytorch_distributed_b200 import cli, driver, launch


def main():
    args = cli.parse_args("apex_distributed")
    env = launch.torchrun_env()
    args.nprocs = env[2] if env else 1
    local_rank = cli.resolve_local_rank(args)
    driver.seed_everything(args)
    if env is None and not args.dist_url:
        args.dist_url = launch.tcp_url()
    driver.main_worker(local_rank, args.nprocs, args)


if __name__ == "__main__":
    main()
