#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (not -x), the dense-only diagnosis, first hardware runs / timings of SIFT detection and the
# reference-kernel comparisons, a bench line, and ncu captures of the kernels VERDICT asks for.
O=gpurun_out/r2a; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi > $O/nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=20 --ignore=tests/test_tsdf_fast_gpu.py > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
timeout 600 python -m pytest tests/test_tsdf_fast_gpu.py -m gpu -q -s -p no:cacheprovider > $O/pytest_fast.log 2>&1; echo "pytest rc=$?" >> $O/pytest_fast.log
timeout 300 python scripts/dense_only_diag.py > $O/dense_diag.jsonl 2> $O/dense_diag.err
timeout 240 python scripts/sift_detect_timing.py > $O/sift_detect_timing.jsonl 2> $O/sift_detect_timing.err
timeout 240 python scripts/ref_siftmgr_compare.py > $O/ref_siftmgr_compare.log 2>&1
timeout 240 python scripts/ref_imageutil_compare.py > $O/ref_imageutil_compare.log 2>&1
timeout 240 python scripts/solver_timing.py > $O/solver_timing.jsonl 2> $O/solver_timing.err
timeout 240 python scripts/sift_match_timing.py > $O/sift_match_timing.jsonl 2> $O/sift_match_timing.err
timeout 400 python bench.py --steps 100 --warmup 10 > $O/bench_fast.json 2> $O/bench_fast.err
BF_TSDF_ARITH=exact timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_exact.json 2> $O/bench_exact.err
# ncu: --set full on one launch of each kernel named by the verdict
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gn_iteration_kernel -s 30 -c 1 -o $O/ncu_gn500 -f python scripts/solver_timing.py > $O/ncu_gn500.log 2>&1
timeout 300 $NCU -k regex:gn_iteration_kernel -s 68 -c 1 -o $O/ncu_gn -f python scripts/solver_timing.py > $O/ncu_gn.log 2>&1
timeout 300 $NCU -k regex:sift_best_kernel -s 244 -c 2 -o $O/ncu_match -f python scripts/sift_match_timing.py > $O/ncu_match.log 2>&1
timeout 300 $NCU -k regex:"sift_level_kernel|sift_describe_kernel|sift_orient_kernel" -s 30 -c 6 -o $O/ncu_detect -f python scripts/sift_detect_timing.py > $O/ncu_detect.log 2>&1
timeout 400 $NCU -k regex:"stencil_fast_kernel" -s 150 -c 3 -o $O/ncu_stencil_fast -f python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_stencil_fast.log 2>&1
for f in gn500 gn match detect stencil_fast; do ncu -i $O/ncu_$f.ncu-rep --page raw --csv > $O/ncu_${f}_raw.csv 2>/dev/null; done
ls -la $O
