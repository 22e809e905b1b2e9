#!/bin/bash
# Round 2: packed 32-bit top-2 in the tcgen05 read-back; Kabsch filter with insertion sort + incremental points; loop bench.
O=gpurun_out/r2l; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_sift_gpu.py -x -q -m gpu > $O/pytest_sift_tma.log 2>&1; echo "sift tma rc=$?"; tail -3 $O/pytest_sift_tma.log
BF_SIFT_MATCH=tc timeout 300 python -m pytest tests/test_sift_gpu.py -x -q -m gpu > $O/pytest_sift_tc.log 2>&1; echo "sift tc rc=$?"; tail -2 $O/pytest_sift_tc.log
for v in tma tc; do BF_SIFT_MATCH=$v timeout 300 python scripts/sift_match_timing.py > $O/sift_match_timing_$v.jsonl 2> $O/sift_match_timing_$v.err; echo "== $v"; cat $O/sift_match_timing_$v.jsonl; tail -c 300 $O/sift_match_timing_$v.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sift_best_tma_kernel -s 240 -c 2 -o $O/ncu_sift_tma -f python scripts/sift_match_timing.py > $O/ncu_sift_tma.log 2>&1
ncu -i $O/ncu_sift_tma.ncu-rep --page raw --csv > $O/ncu_sift_tma_raw.csv 2>/dev/null
timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_frame_loop_gpu.py tests/test_reference_classes_shim.py tests/test_solver_gpu.py -q -m gpu > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_loop.json 2> $O/bench_loop.err; tail -c 300 $O/bench_loop.err
ls $O | head -3
