#!/bin/bash
# Round 2, second GPU call: re-run of the SIFT detection tests (level sigma now host-evaluated), the fast-stencil tests after its
# restructuring (branch-free batched probes, dynamic block deal), A/B benches (exact / fast / fast + block cull), ncu of the stencil.
O=gpurun_out/r2b; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_zz_sift_detect_gpu.py tests/test_tsdf_fast_gpu.py tests/test_tsdf_gpu.py tests/test_tsdf_vs_reference_gpu.py -m gpu -q -s -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_fast.json 2> $O/bench_fast.err
BF_TSDF_CULL=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_fast_cull.json 2> $O/bench_fast_cull.err
BF_TSDF_ARITH=exact timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_exact.json 2> $O/bench_exact.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:"stencil_fast_kernel" -s 150 -c 3 -o $O/ncu_stencil_fast -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_stencil_fast.log 2>&1
BF_TSDF_CULL=1 timeout 400 $NCU -k regex:"stencil_fast_kernel" -s 150 -c 3 -o $O/ncu_stencil_fast_cull -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_stencil_fast_cull.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/launches_bench.log 2>&1
for f in stencil_fast stencil_fast_cull; do ncu -i $O/ncu_$f.ncu-rep --page raw --csv > $O/ncu_${f}_raw.csv 2>/dev/null; done
ls -la $O
