#!/bin/sh
# compute-sanitizer over the kernels rewritten in round 2 (shared-memory staged im2col, stem backward at 64 registers).
#   gpurun --timeout 150 -- 'sh tools/sanitize_r2.sh > gpurun_out/sanitize_r2.log 2>&1'
export PYTHONPATH=$PWD
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused_paths.py -q -x -k "im2col" -p no:cacheprovider 2>&1 | tail -6
echo "memcheck(im2col) exit $?"
timeout 60 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fused_paths.py tests/test_gpu_stem.py -q -x -k "(im2col and shape0) or (stem_forward_backward and dtype0 and shape0)" -p no:cacheprovider 2>&1 | tail -6
echo "racecheck(im2col, stem) exit $?"
