#!/bin/bash
# Final multi-GPU sanity of the frozen code (2 GPUs): cross-GPU checks with NVLS + one bench run of the headline entry.
mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tests/mp_gpu_checks.py > gpurun_out/f2_mpchecks.log 2>&1; grep -o "PASS rank [0-9]" gpurun_out/f2_mpchecks.log | sort | uniq | wc -l; grep -i "assert\|Error" gpurun_out/f2_mpchecks.log | head -3 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/f2_bench2.json 2> gpurun_out/f2_bench2.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/f2_bench2.json | head -4; tail -1 gpurun_out/f2_bench2.err | cut -c1-200
