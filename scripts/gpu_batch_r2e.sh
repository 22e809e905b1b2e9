#!/bin/bash
# Round 2, fifth GPU call: the bench on the frame loop (headline workload) + legacy ops bench with the in-kernel timer; launch list of the loop.
O=gpurun_out/r2e; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_loop.json 2> $O/bench_loop.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_loop_20.json 2> $O/bench_loop_20.err
timeout 300 python bench.py --workload ops --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ops.json 2> $O/bench_ops.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 1500 --csv --log-file $O/launches_loop.csv python bench.py --steps 20 --warmup 5 --preroll 300 --no-cpu-baseline > $O/launches_loop.log 2>&1
python scripts/ncu_summary.py $O/launches_loop.csv > $O/launches_loop.txt 2>&1
ls -la $O
