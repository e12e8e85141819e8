#!/bin/bash
# Round-2 call B2 (2 GPUs): the cross-GPU test tier, then 2-GPU benches of every entry + ablations of the side-stream changes,
# a timeline of one 2-GPU step and the small comm_bench sections.   gpurun --gpus 2 --timeout 2400 -- bash tools/gpu_session_r2b2.sh
mkdir -p gpurun_out
export PYTHONPATH=$PWD
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_entrypoints.py tests/test_gpu_fused_paths.py tests/test_gpu_single_entry.py -m gpu -q -k "not bn_backward2 and not im2col" 2>&1 | tail -60 > $O/b2_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29700
run() { tag=$1; shift; P=$((P+1)); env "$@" timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e $EXTRA > $O/b2_bench_$tag.json 2> $O/b2_bench_$tag.err; }
EXTRA=""
run base PTD_NOOP=1
run nodefer PTD_DEFERRED_BCAST=0
run nomside PTD_METRICS_SIDE=0
run old PTD_DEFERRED_BCAST=0 PTD_METRICS_SIDE=0 "PTD_BENCH_ARGS=--no-overlap-optimizer --bucket-cap-mb 25"
run noov PTD_BENCH_ARGS=--no-overlap-optimizer
run cap25 "PTD_BENCH_ARGS=--bucket-cap-mb 25"
run cap4 "PTD_BENCH_ARGS=--bucket-cap-mb 4"
run bv PTD_BENCH_ARGS=--bucket-view
run ctas16 PTD_MAX_CTAS=16
EXTRA="--comm nccl"; run nccl PTD_NOOP=1; EXTRA=""
EXTRA="--entry apex_distributed"; run apex PTD_NOOP=1
EXTRA="--entry horovod_distributed"; run hvd PTD_NOOP=1
EXTRA="--entry horovod_distributed"; run hvd_dyn PTD_HVD_STATIC=0
EXTRA=""
timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --entry dataparallel > $O/b2_bench_dp.json 2> $O/b2_bench_dp.err
PTD_DP_GRAPH=0 timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e --entry dataparallel > $O/b2_bench_dp_eager.json 2> $O/b2_bench_dp_eager.err
P=$((P+1)); PTD_TIMELINE=$O/b2_tl timeout 300 $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --skip-e2e > $O/b2_bench_tl.json 2> $O/b2_bench_tl.err
P=$((P+1)); timeout 300 $TR --master-port $P tools/comm_bench.py k1small ctas k2 k4 > $O/b2_comm_bench.md 2> $O/b2_comm_bench.err
timeout 200 python tools/comm_bench.py local > $O/b2_comm_local.md 2> $O/b2_comm_local.err
echo "== tests"; tail -n 30 $O/b2_tests.log
for t in base nodefer nomside old noov cap25 cap4 bv ctas16 apex hvd hvd_dyn dp dp_eager tl; do echo "$t: $(grep -o '"ms_per_step": [0-9.]*' $O/b2_bench_$t.json | head -1) $(tail -n 2 $O/b2_bench_$t.err | cut -c1-300)"; done
cat $O/b2_comm_bench.md; tail -3 $O/b2_comm_bench.err; cat $O/b2_comm_local.md; tail -3 $O/b2_comm_local.err
