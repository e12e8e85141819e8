#!/bin/bash
# ncu captures of the collective / optimizer kernels (K1..K6).  2-GPU box:  gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_session_ncu.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" | head -80 > $O/ncu_nvl_metric_names.txt
# (1) one GPU, world 1: K1 / K1b / K4 / K6 / normalise - DRAM, issue and occupancy behaviour
CUDA_VISIBLE_DEVICES=0 timeout 600 $NCU -o $O/ncu_single python tools/ncu_targets.py single > $O/ncu_single.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 600 $NCU -k regex:"stem_" -o $O/ncu_stem python tools/ncu_targets.py stem > $O/ncu_stem.log 2>&1
# (2) one process, two GPUs: K2' push and K5 reduce-to-caller (no in-kernel flags => replay-safe), NVLink / multicast traffic
timeout 600 $NCU -k regex:"pack_only|reduce_to_caller|push_kernel" -o $O/ncu_local python tools/ncu_targets.py local > $O/ncu_local.log 2>&1
# (3) K1 / K2 with their flag barriers: rank 1 runs plainly, ONLY rank 0 is under ncu (replay is idempotent, see tools/ncu_targets.py)
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29911 WORLD_SIZE=2 PTD_COMM_TIMEOUT_MS=60000
RANK=1 LOCAL_RANK=1 timeout 600 python tools/ncu_targets.py rank > $O/ncu_rank1.log 2>&1 &
R1=$!
RANK=0 LOCAL_RANK=0 timeout 600 $NCU -k regex:"fused_allreduce|fused_broadcast" -o $O/ncu_rank0 python tools/ncu_targets.py rank > $O/ncu_rank0.log 2>&1
wait $R1
ls -la $O/*.ncu-rep
tail -n 3 $O/ncu_single.log $O/ncu_stem.log $O/ncu_local.log $O/ncu_rank0.log $O/ncu_rank1.log
# re-measure the latency-bound collectives after the unroll / granularity changes
P=29950; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/comm_bench.py k1small k2 > $O/ncu_comm_bench2.md 2> $O/ncu_comm_bench2.err
timeout 200 python tools/comm_bench.py local > $O/ncu_comm_local2.md 2> $O/ncu_comm_local2.err
cat $O/ncu_comm_bench2.md $O/ncu_comm_local2.md; tail -2 $O/ncu_comm_bench2.err $O/ncu_comm_local2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
