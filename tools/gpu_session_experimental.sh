#!/bin/bash
# First GPU call of the next session: validate the opt-in paths written without hardware access and measure what they buy.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- bash tools/gpu_session_experimental.sh
# Everything lands in gpurun_out/ (scratch); copy the summaries worth keeping into profiles/.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
PTD_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -q 2>&1 | tail -25 > gpurun_out/exp_tests.log
timeout 300 python tools/stem_gemm_probe.py 256 > gpurun_out/stem_gemm_probe.md 2>&1
timeout 300 python tools/kernel_bench.py > gpurun_out/kernel_bench_v2.md 2>&1
run() {   # tag, env assignments...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e > gpurun_out/bench_exp_$tag.json 2> gpurun_out/bench_exp_$tag.err
}
run base PTD_NOOP=1
run split PTD_SPLIT_RESGRAD=1
run stem PTD_STEM_GEMM=1
run both PTD_SPLIT_RESGRAD=1 PTD_STEM_GEMM=1
# horovod entry, eager vs static schedule + CUDA graph (single GPU exercises hooks, queue and capture; 2+ GPUs: tools/gpu_session_8gpu.sh)
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e --entry horovod_distributed > gpurun_out/bench_exp_hvd_eager.json 2> gpurun_out/bench_exp_hvd_eager.err
PTD_HVD_STATIC=1 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-e2e --entry horovod_distributed > gpurun_out/bench_exp_hvd_static.json 2> gpurun_out/bench_exp_hvd_static.err
tail -n 3 gpurun_out/exp_tests.log
cat gpurun_out/stem_gemm_probe.md
grep -E "split gradients|ATen add" gpurun_out/kernel_bench_v2.md
for t in base split stem both hvd_eager hvd_static; do echo "$t: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_exp_$t.json)"; done
