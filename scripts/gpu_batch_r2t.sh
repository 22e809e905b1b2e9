#!/bin/bash
# Round 2: SIFT level kernel with batched staging / interleaved tap chains, fuse kernel reading shared memory only in its serial part: tests, detection timing, loop bench, launch list.
O=gpurun_out/r2t; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_zz_sift_detect_gpu.py tests/test_fuse_gpu.py tests/test_frame_loop_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 600 python scripts/sift_detect_timing.py > $O/sift_detect_timing.jsonl 2> $O/sift_detect_timing.err; cat $O/sift_detect_timing.jsonl | cut -c1-300
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_100.json 2> $O/bench_100.err; tail -c 300 $O/bench_100.err; head -c 300 $O/bench_100.json; echo
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_loop.csv python bench.py --steps 20 --warmup 5 --cuda-profiler --no-cpu-baseline > $O/launches_loop.log 2>&1
python scripts/ncu_summary.py $O/launches_loop.csv > $O/launches_loop.txt 2>&1; head -12 $O/launches_loop.txt; grep "sift_fuse\|gn_iter\|dense_build" $O/launches_loop.txt
