#!/bin/bash
# Round 2, look-ahead step + warp-cooperative Kabsch filter: their GPU tests, the loop bench with look-ahead on / off, launch list of the timed pass.
O=gpurun_out/r2p; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_frame_loop_gpu.py tests/test_filter_gpu.py tests/test_verify_filters_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ahead.json 2> $O/bench_ahead.err; tail -c 300 $O/bench_ahead.err; head -c 400 $O/bench_ahead.json; echo
BF_LOOP_AHEAD=0 timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_noahead.json 2> $O/bench_noahead.err; head -c 400 $O/bench_noahead.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_ahead_20.json 2> $O/bench_ahead_20.err; head -c 400 $O/bench_ahead_20.json; echo
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_loop.csv python bench.py --steps 20 --warmup 5 --cuda-profiler --no-cpu-baseline > $O/launches_loop.log 2>&1
python scripts/ncu_summary.py $O/launches_loop.csv > $O/launches_loop.txt 2>&1; head -60 $O/launches_loop.txt
