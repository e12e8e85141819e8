#!/bin/sh
# Race / memory checking of the single-GPU kernels (SURVEY section 5 "race detection"): run on a GPU box, e.g.
#   gpurun --timeout 900 -- 'sh tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (BN combine rows, stem, GEMM staging).
set -x
K='bn_act_forward_backward and 256 or fused_sgd_flat_matches or normalize_kernel or metrics_kernel or multi_tensor_scale'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "$K" -p no:cacheprovider
echo "memcheck exit $?"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stem.py -q -x -k "stem_forward_backward and dtype0 and shape0" -p no:cacheprovider
echo "racecheck(stem) exit $?"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "bn_act_forward_backward and 64 and dtype0" -p no:cacheprovider
echo "racecheck(bn) exit $?"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tcgen05.py -q -x -k "shape0 or shape1" -p no:cacheprovider
echo "memcheck(tcgen05) exit $?"
