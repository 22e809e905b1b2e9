#!/bin/bash
# Round 2: ray cast on hardware (parity vs oracle, timing), fuse kernel with parallel root selection, solver barrier variants at small N, loop bench.
O=gpurun_out/r2s; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_raycast_gpu.py tests/test_fuse_gpu.py tests/test_frame_loop_gpu.py -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
timeout 600 python scripts/raycast_timing.py > $O/raycast_timing.jsonl 2> $O/raycast_timing.err; cat $O/raycast_timing.jsonl; tail -3 $O/raycast_timing.err
timeout 600 python scripts/solver_timing.py > $O/solver_timing_coop.jsonl 2>&1; cat $O/solver_timing_coop.jsonl
BF_SOLVER_BARRIER=cluster timeout 600 python scripts/solver_timing.py > $O/solver_timing_cluster.jsonl 2>&1; cat $O/solver_timing_cluster.jsonl
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_100.json 2> $O/bench_100.err; tail -c 300 $O/bench_100.err; head -c 300 $O/bench_100.json; echo
BF_SOLVER_BARRIER=cluster timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_100_cluster.json 2> $O/bench_100_cluster.err; head -c 300 $O/bench_100_cluster.json; echo
