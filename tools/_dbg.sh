mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_fused_paths.py tests/test_gpu_stem.py -q -k "not horovod" 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:stem_ --csv --log-file gpurun_out/d_stem_times.csv python tools/ncu_targets.py stem > gpurun_out/d_stem.log 2>&1; grep -o '"[^"]*stem[^"(]*(\|"gpu__time_duration.sum","ns","[0-9,.]*"' gpurun_out/d_stem_times.csv | paste - - | cut -c1-160
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/d_bench1.json 2> gpurun_out/d_bench1.err; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' gpurun_out/d_bench1.json | head -4; tail -1 gpurun_out/d_bench1.err | cut -c1-200
