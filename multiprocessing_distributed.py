"""Self-launching DDP entrypoint (reference: /root/reference/multiprocessing_distributed.py, start.sh:1).

    python multiprocessing_distributed.py -a resnet50 -b 2048 --synthetic

``mp.spawn`` one worker per GPU, TCP rendezvous on 127.0.0.1:23456 (a free port is chosen if it is taken).
"""
from pytorch_distributed_b200 import cli, driver, launch


def worker(local_rank, nprocs, args):
    driver.seed_everything(args)                   # the reference seeds inside the worker for this script (:120-128)
    driver.main_worker(local_rank, nprocs, args)


def main():
    args = cli.parse_args("multiprocessing_distributed")
    args.nprocs = launch.default_nprocs(args)
    if not args.dist_url:
        args.dist_url = launch.tcp_url(launch.pick_port(launch.DEFAULT_PORT))
    launch.spawn(worker, args.nprocs, args)


if __name__ == "__main__":
    main()
