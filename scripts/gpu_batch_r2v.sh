#!/bin/bash
# Round 2, eight GPUs: the configs[3] sweep (hash sharded by block owner) and the loop bench as the driver launches them.
O=gpurun_out/r2v; mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 48 --warmup 8 --workload sweep > $O/sweep_n8.json 2> $O/sweep_n8.err; tail -c 300 $O/sweep_n8.err; head -c 300 $O/sweep_n8.json; echo
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_loop_n8.json 2> $O/bench_loop_n8.err; tail -c 300 $O/bench_loop_n8.err; head -c 300 $O/bench_loop_n8.json; echo
