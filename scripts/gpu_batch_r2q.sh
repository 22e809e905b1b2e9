#!/bin/bash
# Round 2: where does the look-ahead loop lose time at 100 steps?  Per-step host wall times; tiled ingest / cache kernels on hardware.
O=gpurun_out/r2q; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_cache_gpu.py tests/test_frame_loop_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
BF_LOOP_STEPTIMES=$O/steptimes_ahead.txt timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ahead.json 2> $O/bench_ahead.err; head -c 300 $O/bench_ahead.json; echo
BF_LOOP_OVERLAP=0 BF_LOOP_STEPTIMES=$O/steptimes_ahead_nooverlap.txt timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ahead_nooverlap.json 2> $O/bench_ahead_nooverlap.err; head -c 300 $O/bench_ahead_nooverlap.json; echo
sort -k2 -n -r $O/steptimes_ahead.txt | head -12
sort -k2 -n -r $O/steptimes_ahead_nooverlap.txt | head -12
