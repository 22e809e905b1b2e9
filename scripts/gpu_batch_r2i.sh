#!/bin/bash
# Round 2, multi-GPU call (4 GPUs): configs[3] sweep at N = 1, 2, 4; the frame-loop bench at N = 1 (with the reference_cuda leg), 2, 4.
O=gpurun_out/r2i; mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --workload sweep --steps 48 --warmup 8 > $O/sweep_n1.json 2> $O/sweep_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --workload sweep --gpus 2 --steps 48 --warmup 8 > $O/sweep_n2.json 2> $O/sweep_n2.err
timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --workload sweep --gpus 4 --steps 48 --warmup 8 > $O/sweep_n4.json 2> $O/sweep_n4.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 900 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err
for f in sweep_n1 sweep_n2 sweep_n4 bench_n1 bench_n2 bench_n4; do echo "== $f"; tail -c 300 $O/$f.err; head -c 200 $O/$f.json; echo; done
