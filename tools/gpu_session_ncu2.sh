#!/bin/bash
# ncu counters for the CROSS-GPU kernels (K1, K2, K2', K5).  ncu's multi-pass kernel replay has to save / restore device memory and fails
# ("UnknownError") on peer-mapped / multicast allocations, so every run here collects a metric list that fits in ONE pass.
#   gpurun --gpus 2 --timeout 1200 -- bash tools/gpu_session_ncu2.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
NCU="ncu --clock-control none --profile-from-start off --csv"
i=0
for M in "nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum" "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum" "lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,gpu__time_duration.sum" "sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size"; do
  i=$((i+1))
  timeout 300 $NCU --metrics $M -k regex:"pack_only|reduce_to_caller|push_kernel" --log-file $O/ncu2_local_$i.csv python tools/ncu_targets.py local > $O/ncu2_local_$i.log 2>&1
  export MASTER_ADDR=127.0.0.1 MASTER_PORT=$((29920+i)) WORLD_SIZE=2 PTD_COMM_TIMEOUT_MS=60000
  RANK=1 LOCAL_RANK=1 timeout 300 python tools/ncu_targets.py rank > $O/ncu2_rank1_$i.log 2>&1 &
  R1=$!
  RANK=0 LOCAL_RANK=0 timeout 300 $NCU --metrics $M -k regex:"fused_allreduce|fused_broadcast" --log-file $O/ncu2_rank0_$i.csv python tools/ncu_targets.py rank > $O/ncu2_rank0_$i.log 2>&1
  wait $R1
  echo "set $i: local $(grep -c -i error $O/ncu2_local_$i.csv) errors, rank $(grep -c -i error $O/ncu2_rank0_$i.csv) errors; $(tail -n 1 $O/ncu2_rank0_$i.log | cut -c1-120)"
done
head -c 3000 $O/ncu2_local_1.csv; echo; head -c 3000 $O/ncu2_rank0_1.csv
# hvd with bf16 parameters + masters in the multi-tensor optimizer (changed after session B2)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29960 bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --entry horovod_distributed > $O/ncu2_bench_hvd.json 2> $O/ncu2_bench_hvd.err
echo "hvd: $(grep -o '"ms_per_step": [0-9.]*' $O/ncu2_bench_hvd.json) $(tail -n 1 $O/ncu2_bench_hvd.err | cut -c1-200)"
timeout 600 python -m pytest tests/test_gpu_entrypoints.py tests/test_gpu_fused_paths.py -q -k "horovod" 2>&1 | tail -4
