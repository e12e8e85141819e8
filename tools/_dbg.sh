mkdir -p gpurun_out; export PYTHONPATH=$PWD
for nv in 1 0; do PTD_NVLS=$nv timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$nv tests/mp_gpu_checks.py > gpurun_out/dbg_mpchecks_nvls$nv.log 2>&1; grep -v "^W0\|Warn" gpurun_out/dbg_mpchecks_nvls$nv.log | grep -i "info\|PASS\|Error\|assert" | head -8; done
timeout 900 python -m pytest tests/test_gpu_entrypoints.py -q -k "dataparallel" 2>&1 | tail -5
timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --entry dataparallel > gpurun_out/dbg_bench_dp.json 2> gpurun_out/dbg_bench_dp.err; grep -o '"ms_per_step": [0-9.]*\|"host_enqueue_ms_per_step": [0-9.]*' gpurun_out/dbg_bench_dp.json; tail -1 gpurun_out/dbg_bench_dp.err | cut -c1-200
