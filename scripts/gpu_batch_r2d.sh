#!/bin/bash
# Round 2, fourth GPU call: first run of the frame loop; VerifyTrajectory / fuseToGlobal tests; batch stencil with the heavy-first deal.
O=gpurun_out/r2d; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_frame_loop_gpu.py -m gpu -q -s -x -p no:cacheprovider > $O/pytest_loop.log 2>&1; echo "pytest rc=$?" >> $O/pytest_loop.log
timeout 600 python -m pytest tests/test_verify_filters_gpu.py tests/test_fuse_gpu.py tests/test_tsdf_fast_gpu.py -m gpu -q -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_batch.json 2> $O/bench_batch.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:"stencil_multi_kernel" -s 12 -c 2 -o $O/ncu_stencil_multi -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_stencil_multi.log 2>&1
ncu -i $O/ncu_stencil_multi.ncu-rep --page raw --csv > $O/ncu_stencil_multi_raw.csv 2>/dev/null
ls -la $O
