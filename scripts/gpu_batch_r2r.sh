#!/bin/bash
# Round 2: workspaces reserved at loop creation (no cudaFree / cudaMalloc in the steady state), describe kernel with a warp per cell and a grid over the
# features that exist, tiled ingest / cache kernels: GPU tests of the touched rows, loop bench at both settings, per-step wall times, launch list.
O=gpurun_out/r2r; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_zz_sift_detect_gpu.py tests/test_ingest_gpu.py tests/test_cache_gpu.py tests/test_frame_loop_gpu.py tests/test_sift_gpu.py tests/test_solver_gpu.py tests/test_reference_classes_shim.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
BF_LOOP_STEPTIMES=$O/steptimes.txt timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_100.json 2> $O/bench_100.err; tail -c 300 $O/bench_100.err; head -c 300 $O/bench_100.json; echo
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_20.json 2> $O/bench_20.err; head -c 300 $O/bench_20.json; echo
sort -k2 -n -r $O/steptimes.txt | head -6
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_loop.csv python bench.py --steps 20 --warmup 5 --cuda-profiler --no-cpu-baseline > $O/launches_loop.log 2>&1
python scripts/ncu_summary.py $O/launches_loop.csv > $O/launches_loop.txt 2>&1; head -24 $O/launches_loop.txt
