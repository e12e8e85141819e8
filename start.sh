# Canonical launch lines (reference: /root/reference/start.sh); add --synthetic when no ImageNet tree is available.
python multiprocessing_distributed.py -a resnet50 -b 2048
python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 distributed.py -a resnet50 -b 2048
python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 apex_distributed.py -a resnet50 -b 2048 --opt-level O2
python -m torch.distributed.run --nproc_per_node=8 --master-addr 127.0.0.1 horovod_distributed.py -a resnet50 -b 2048
python dataparallel.py -a resnet50 -b 2048
srun -N2 --gres gpu:8 python distributed_slurm_main.py --dist-file dist_file -a resnet50 -b 2048
