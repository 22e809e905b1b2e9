#!/bin/bash
# Round 2, third GPU call: batch re-integration (multi-op stencil) -- tests, A/B bench, ncu; first hardware run of VerifyTrajectory and fuseToGlobal.
O=gpurun_out/r2c; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_tsdf_fast_gpu.py tests/test_tsdf_gpu.py tests/test_verify_filters_gpu.py tests/test_fuse_gpu.py -m gpu -q -s -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_batch.json 2> $O/bench_batch.err
BF_TSDF_BATCH=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_nobatch.json 2> $O/bench_nobatch.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-ba > $O/bench_batch_noba.json 2> $O/bench_batch_noba.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:"stencil_multi_kernel" -s 12 -c 2 -o $O/ncu_stencil_multi -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_stencil_multi.log 2>&1
timeout 400 $NCU -k regex:"alloc_kernel" -s 150 -c 2 -o $O/ncu_alloc -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_alloc.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/launches_bench.log 2>&1
for f in stencil_multi alloc; do ncu -i $O/ncu_$f.ncu-rep --page raw --csv > $O/ncu_${f}_raw.csv 2>/dev/null; done
ls -la $O
