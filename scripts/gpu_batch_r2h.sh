#!/bin/bash
# Round 2, eighth GPU call: tcgen05 matcher (tests under a timeout, then A/B timing against the mma.sync sweep), N = 72 dense tests at 2 x 15 PCG,
# two-stream frame loop (equivalence test + bench A/B), configs[3] sweep at N = 1.
O=gpurun_out/r2h; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_sift_gpu.py -x -q -m gpu > $O/pytest_sift_tc.log 2>&1; echo "sift tc rc=$?"; tail -3 $O/pytest_sift_tc.log
BF_SIFT_MATCH=mma timeout 300 python -m pytest tests/test_sift_gpu.py -x -q -m gpu > $O/pytest_sift_mma.log 2>&1; echo "sift mma rc=$?"; tail -2 $O/pytest_sift_mma.log
timeout 300 python scripts/sift_match_timing.py > $O/sift_match_timing_tc.jsonl 2> $O/sift_match_timing_tc.err; cat $O/sift_match_timing_tc.jsonl
BF_SIFT_MATCH=mma timeout 300 python scripts/sift_match_timing.py > $O/sift_match_timing_mma.jsonl 2>&1; cat $O/sift_match_timing_mma.jsonl
timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_solver_vs_reference_gpu.py tests/test_frame_loop_gpu.py tests/test_reference_classes_shim.py tests/test_filter_gpu.py -q -m gpu > $O/pytest.log 2>&1
tail -6 $O/pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_loop.json 2> $O/bench_loop.err
BF_LOOP_OVERLAP=0 timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_loop_1stream.json 2> $O/bench_loop_1stream.err
timeout 900 python bench.py --workload sweep --steps 48 --warmup 8 > $O/sweep_n1.json 2> $O/sweep_n1.err
tail -c 600 $O/sweep_n1.err
ls -la $O
