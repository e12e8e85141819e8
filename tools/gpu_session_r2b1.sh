#!/bin/bash
# Round-2 call B1 (1 GPU): the whole single-GPU test tier (incl. the opt-in kernels), then 1-GPU A/B of the new engine modes.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > $O/b1_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e > $O/b1_bench_$tag.json 2> $O/b1_bench_$tag.err; }
run base PTD_NOOP=1
run both PTD_SPLIT_RESGRAD=1 PTD_STEM_GEMM=1
run both_noov PTD_BENCH_ARGS=--no-overlap-optimizer
run both_bv PTD_BENCH_ARGS=--bucket-view
run both_cap25 "PTD_BENCH_ARGS=--bucket-cap-mb 25"
PTD_TIMELINE=$O/b1_tl timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e > $O/b1_bench_tl.json 2> $O/b1_bench_tl.err
echo "== tests"; tail -n 25 $O/b1_tests.log
for t in base both both_noov both_bv both_cap25 tl; do echo "$t: $(grep -o '"ms_per_step": [0-9.]*' $O/b1_bench_$t.json | head -1) $(tail -n 2 $O/b1_bench_$t.err | cut -c1-300)"; done
