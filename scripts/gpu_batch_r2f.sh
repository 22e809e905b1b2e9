#!/bin/bash
# Round 2, sixth GPU call: stencil trims + unified pair update + batch cull (tests, A/B of the cull), frame-loop trace (tracking diagnosis), shim run test.
O=gpurun_out/r2f; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_tsdf_fast_gpu.py tests/test_tsdf_gpu.py tests/test_reference_classes_shim.py tests/test_frame_loop_gpu.py -x -q -m gpu > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --trace $O/trace_loop.txt > $O/bench_loop.json 2> $O/bench_loop.err
BF_TSDF_BATCH_CULL=0 timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_loop_nocull.json 2> $O/bench_loop_nocull.err
timeout 300 python bench.py --workload ops --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ops.json 2> $O/bench_ops.err
BF_TSDF_BATCH_CULL=0 timeout 300 python bench.py --workload ops --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ops_nocull.json 2> $O/bench_ops_nocull.err
ls -la $O
