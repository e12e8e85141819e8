#!/bin/bash
# 8-GPU call (charged 8x): BASELINE configs 3-5 at full width, DataParallel parity at 8 devices, the collective micro-benchmarks.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1200 -- bash tools/gpu_session_8gpu.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 --entry apex_distributed --opt-level O2 > gpurun_out/bench8_apex.json 2> gpurun_out/bench8_apex.err
timeout 300 $TR --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 --entry horovod_distributed > gpurun_out/bench8_hvd.json 2> gpurun_out/bench8_hvd.err
PTD_HVD_STATIC=1 timeout 300 $TR --master-port 29604 bench.py --gpus 8 --steps 20 --warmup 5 --entry horovod_distributed > gpurun_out/bench8_hvdstatic.json 2> gpurun_out/bench8_hvdstatic.err
timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --entry dataparallel > gpurun_out/bench8_dp.json 2> gpurun_out/bench8_dp.err
PTD_TEST_DP_GPUS=8 timeout 300 python -m pytest tests/test_gpu_entrypoints.py -q -k "dataparallel_matches" 2>&1 | tail -5 > gpurun_out/dp8_parity.log
timeout 400 $TR --master-port 29603 tools/comm_bench.py > gpurun_out/comm_bench_8gpu_v3.md 2> gpurun_out/comm_bench_v3.err
for f in apex hvd hvdstatic dp; do echo "$f: $(grep -o '"value": [0-9.]*' gpurun_out/bench8_$f.json | head -1) $(tail -n 2 gpurun_out/bench8_$f.err | tr '\n' ' ' | cut -c1-300)"; done
cat gpurun_out/dp8_parity.log
tail -n 25 gpurun_out/comm_bench_8gpu_v3.md
