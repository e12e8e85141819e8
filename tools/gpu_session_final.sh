#!/bin/bash
# Final verification on a fresh 1-GPU box: what the driver runs at round end (GPU test tier, smoke(), both bench arms).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/f_ref1.json 2> $O/f_ref1.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/f_own1.json 2> $O/f_own1.err
tail -n 8 $O/f_tests.log; tail -n 2 $O/f_smoke.log
grep -o '"value": [0-9.]*' $O/f_ref1.json | head -1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/f_own1.json | head -4
