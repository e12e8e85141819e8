#!/bin/bash
# Round-2 call C (8 GPUs, charged 8x): BASELINE configs 2-5 at full width, the cross-GPU tests at world 8, the ablation table, the
# collective micro-benchmarks.  Ordered by priority; items are skipped once BUDGET_S seconds have elapsed.
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_session_r2c.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
T0=$(date +%s)
BUDGET_S=${BUDGET_S:-560}
left() { now=$(date +%s); [ $((now - T0)) -lt $BUDGET_S ]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
P=29800
bench() { tag=$1; shift; left || { echo "skip $tag (budget)"; return; }; P=$((P+1)); timeout 240 $TR --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 "$@" > $O/c_$tag.json 2> $O/c_$tag.err; echo "$tag: $(grep -o '"value": [0-9.]*' $O/c_$tag.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' $O/c_$tag.json | head -1) $(grep -o '"final_loss": [0-9.]*' $O/c_$tag.json) $(tail -n 1 $O/c_$tag.err | cut -c1-200)"; }
PTD_TIMELINE=$O/c_tl bench own
bench ref --impl reference
bench apex --entry apex_distributed --skip-e2e
bench hvd --entry horovod_distributed --skip-e2e
left && { PTD_PYPROFILE=$O/c_dp_pyprofile.txt timeout 240 python bench.py --gpus 8 --steps 20 --warmup 5 --skip-e2e --entry dataparallel > $O/c_dp.json 2> $O/c_dp.err; echo "dp: $(grep -o '"value": [0-9.]*' $O/c_dp.json | head -1) $(grep -o '"host_enqueue_ms_per_step": [0-9.]*' $O/c_dp.json) $(tail -n 1 $O/c_dp.err | cut -c1-200)"; }
left && { P=$((P+1)); timeout 400 $TR --master-port $P tests/mp_gpu_checks.py > $O/c_mpchecks.log 2>&1; echo "mp_gpu_checks world 8: $(grep -c 'PASS rank' $O/c_mpchecks.log) PASS; $(grep -m1 '\[info\]' $O/c_mpchecks.log)"; grep -i "assert\|error" $O/c_mpchecks.log | head -3 | cut -c1-300; }
left && { P=$((P+1)); timeout 300 $TR --master-port $P tools/comm_bench.py k1 k1small k2 k4 > $O/c_comm_bench.md 2> $O/c_comm_bench.err; tail -n 2 $O/c_comm_bench.err | cut -c1-200; }
bench nccl --comm nccl --skip-e2e
left && { PTD_TEST_DP_GPUS=8 timeout 300 python -m pytest tests/test_gpu_entrypoints.py -q -k "dataparallel_matches" 2>&1 | tail -4 > $O/c_dp8_parity.log; cat $O/c_dp8_parity.log; }
left && { timeout 200 python tools/comm_bench.py local > $O/c_comm_local.md 2> $O/c_comm_local.err; tail -n 2 $O/c_comm_local.err | cut -c1-200; }
left && { P=$((P+1)); timeout 240 $TR --master-port $P tools/torch_ddp_bf16_baseline.py > $O/c_torch_bf16.json 2> $O/c_torch_bf16.err; cat $O/c_torch_bf16.json; }
bench nofusedbn --no-fused-bn --skip-e2e
bench torchopt --optimizer torch --skip-e2e
bench nograph --no-cuda-graph --skip-e2e
PTD_DEFERRED_BCAST=0 PTD_METRICS_SIDE=0 PTD_BENCH_ARGS="--no-overlap-optimizer --bucket-cap-mb 25" bench r1engine --skip-e2e
left && { P=$((P+1)); timeout 200 $TR --master-port $P tools/comm_bench.py ctas > $O/c_comm_ctas.md 2> $O/c_comm_ctas.err; }
echo "elapsed $(( $(date +%s) - T0 )) s"
cat $O/c_comm_bench.md $O/c_comm_local.md $O/c_comm_ctas.md 2>/dev/null | head -150
