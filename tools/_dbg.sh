mkdir -p gpurun_out; export PYTHONPATH=$PWD
PTD_PYPROFILE=gpurun_out/c2_dp_pyprofile.txt timeout 240 python bench.py --gpus 8 --steps 20 --warmup 5 --entry dataparallel > gpurun_out/c2_dp.json 2> gpurun_out/c2_dp.err; cat gpurun_out/c2_dp.json; tail -2 gpurun_out/c2_dp.err | cut -c1-300; head -12 gpurun_out/c2_dp_pyprofile.txt
PTD_TEST_DP_GPUS=8 timeout 200 python -m pytest tests/test_gpu_entrypoints.py -q -k "dataparallel_matches" 2>&1 | tail -3
