#!/bin/bash
# Last sanity after the final (host-side) edits: smoke(), the single-GPU entrypoint tests, one short bench run.
mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_single_entry.py -q 2>&1 | tail -2
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --skip-e2e > gpurun_out/h_bench1.json 2> gpurun_out/h_bench1.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/h_bench1.json; tail -1 gpurun_out/h_bench1.err | cut -c1-200
