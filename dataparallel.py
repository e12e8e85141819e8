"""Single-process multi-GPU entrypoint (reference: /root/reference/dataparallel.py, README:86 `python dataparallel.py`).

    python dataparallel.py -a resnet50 -b 2048 --synthetic [--gpus 0,1,2,3]

One process drives every GPU through pytorch_distributed_b200.parallel.dp.DataParallel: persistent replicas, a
multicast parameter broadcast before forward and an in-switch gradient reduce onto GPU0 after backward.  Writes the
reference's per-epoch ``dataparallel.csv``.
"""
from pytorch_distributed_b200 import cli, driver


def main():
    args = cli.parse_args("dataparallel")
    driver.seed_everything(args)
    args.nprocs = 1
    driver.main_worker(0, 1, args)


if __name__ == "__main__":
    main()
