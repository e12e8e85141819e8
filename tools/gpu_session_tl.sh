#!/bin/bash
mkdir -p gpurun_out; export PYTHONPATH=$PWD
PTD_TIMELINE=gpurun_out/g_tl timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/g_bench1.json; ls -la gpurun_out/g_tl*
