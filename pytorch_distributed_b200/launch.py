"""Process topology helpers (the L6 layer of SURVEY section 1).

  * ``spawn``            - ``mp.spawn`` self-launch with TCP rendezvous (/root/reference/multiprocessing_distributed.py:110-135;
                           port 23456 is kept as the default but a free port is picked when it is taken).
  * ``slurm_topology``   - node rank / world from ``SLURM_*`` + ``file://`` rendezvous (/root/reference/distributed_slurm_main.py:124-140).
  * ``torchrun_env``     - RANK / LOCAL_RANK / WORLD_SIZE as set by ``torch.distributed.run`` (start.sh:2-3 used the older
                           ``torch.distributed.launch``; both spellings of ``--local_rank`` are accepted by the CLI).
"""
from __future__ import annotations

import os
import socket
from typing import Callable, Optional

import torch
import torch.multiprocessing as mp

DEFAULT_PORT = 23456


def port_is_free(port: int, host: str = "127.0.0.1") -> bool:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        try:
            s.bind((host, port))
            return True
        except OSError:
            return False


def pick_port(preferred: int = DEFAULT_PORT) -> int:
    if port_is_free(preferred):
        return preferred
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun_env():
    """(rank, local_rank, world_size) from the launcher's environment, or None when not launched by it."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        return int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ["WORLD_SIZE"])
    return None


def default_nprocs(args) -> int:
    if getattr(args, "world_size", None):
        return int(args.world_size)
    if (args.device or "cuda").startswith("cuda") and torch.cuda.is_available():
        return torch.cuda.device_count()
    return 1


def _spawn_entry(i: int, fn: Callable, nprocs: int, args, env: dict):
    os.environ.update(env)
    os.environ["LOCAL_RANK"] = str(i)
    fn(i, nprocs, args)


def spawn(fn: Callable, nprocs: int, args, extra_env: Optional[dict] = None) -> None:
    """``mp.spawn(fn, nprocs=nprocs, args=(nprocs, args))`` with error propagation (join=True)."""
    env = dict(extra_env or {})
    if nprocs == 1:
        _spawn_entry(0, fn, 1, args, env)
        return
    mp.spawn(_spawn_entry, nprocs=nprocs, args=(fn, nprocs, args, env), join=True)


def tcp_url(port: Optional[int] = None) -> str:
    return "tcp://127.0.0.1:%d" % (port or pick_port())


def slurm_topology(args, ngpus_per_node: int):
    """Returns (node_rank, n_nodes, world_size, dist_url) from the Slurm environment."""
    node_rank = int(os.environ.get("SLURM_PROCID", "0"))
    n_nodes = int(os.environ.get("SLURM_NPROCS", "1"))
    job = os.environ.get("SLURM_JOBID", "0")
    if args.dist_file is not None:
        url = "file://{}.{}".format(os.path.realpath(args.dist_file), job)
    else:
        url = args.dist_url or tcp_url()
    return node_rank, n_nodes, n_nodes * ngpus_per_node, url
