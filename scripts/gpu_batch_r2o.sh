#!/bin/bash
# Round 2, driver-like check on one GPU: the whole -m gpu suite with -x, smoke(), the default bench (with its reference_cuda and cpu_baseline legs) at the driver's
# short setting and at the default, the reference arm.
O=gpurun_out/r2o; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/bench_20.json 2> $O/bench_20.err; tail -c 400 $O/bench_20.err; head -c 300 $O/bench_20.json; echo
timeout 1500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 400 $O/bench_default.err; head -c 300 $O/bench_default.json; echo
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; head -c 300 $O/bench_reference.json; echo
