#!/bin/bash
# Round 2, seventh GPU call: deterministic matcher, block-sparse dense term (N = 72 vs oracle and reference CUDA), shim, loop bench with stage profile on a tracked
# stream segment, launch list of the timed pass, full ncu capture of the batch stencil inside the loop.
O=gpurun_out/r2g; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests/test_sift_gpu.py tests/test_reference_classes_shim.py tests/test_solver_gpu.py tests/test_solver_vs_reference_gpu.py tests/test_frame_loop_gpu.py tests/test_tsdf_fast_gpu.py -q -m gpu > $O/pytest.log 2>&1
tail -8 $O/pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --trace $O/trace_loop.txt > $O/bench_loop.json 2> $O/bench_loop.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_loop_20.json 2> $O/bench_loop_20.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_loop.csv python bench.py --steps 20 --warmup 5 --cuda-profiler --no-cpu-baseline > $O/launches_loop.log 2>&1
python scripts/ncu_summary.py $O/launches_loop.csv > $O/launches_loop.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stencil_multi_kernel -c 2 -o $O/ncu_stencil_multi_loop -f python bench.py --steps 20 --warmup 5 --cuda-profiler --no-cpu-baseline > $O/ncu_stencil_multi_loop.log 2>&1
ncu -i $O/ncu_stencil_multi_loop.ncu-rep --page raw --csv > $O/ncu_stencil_multi_loop_raw.csv 2>/dev/null
ls -la $O
