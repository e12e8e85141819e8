#!/bin/bash
# Round-2 call A (1 GPU): reference arm + own arm at N=1, validation of the opt-in kernels written without hardware, A/B of the flags,
# a kernel timeline of the default step.   gpurun --timeout 1500 -- bash tools/gpu_session_r2a.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/a_ref1.json 2> $O/a_ref1.err
PTD_TIMELINE=$O/a_tl timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/a_own1.json 2> $O/a_own1.err
timeout 600 python -m pytest tests/test_gpu_fused_paths.py -q -k "not horovod" 2>&1 | tail -40 > $O/a_exp_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e > $O/a_bench_$tag.json 2> $O/a_bench_$tag.err; }
run nosplit PTD_SPLIT_RESGRAD=0
run nostem PTD_STEM_GEMM=0
run neither PTD_SPLIT_RESGRAD=0 PTD_STEM_GEMM=0
timeout 200 python tools/stem_gemm_probe.py 256 > $O/a_stem_gemm_probe.md 2>&1
echo "== ref"; cat $O/a_ref1.json; tail -n 3 $O/a_ref1.err
echo "== own"; cat $O/a_own1.json; tail -n 3 $O/a_own1.err
echo "== exp tests"; tail -n 15 $O/a_exp_tests.log
for t in split stem both; do echo "$t: $(grep -o '"ms_per_step": [0-9.]*' $O/a_bench_$t.json | head -1) $(tail -n 2 $O/a_bench_$t.err | cut -c1-300)"; done
tail -n 20 $O/a_stem_gemm_probe.md
