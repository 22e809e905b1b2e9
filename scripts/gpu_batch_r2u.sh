#!/bin/bash
# Round 2, two GPUs: the loop bench with the look-ahead step under torchrun (as the driver launches it), the configs[3] sweep, the sharded-PCG test.
O=gpurun_out/r2u; mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_loop_n2.json 2> $O/bench_loop_n2.err; tail -c 400 $O/bench_loop_n2.err; head -c 400 $O/bench_loop_n2.json; echo
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 48 --warmup 8 --workload sweep > $O/sweep_n2.json 2> $O/sweep_n2.err; tail -c 300 $O/sweep_n2.err; head -c 300 $O/sweep_n2.json; echo
timeout 600 python -m pytest tests/test_solver_peers_gpu.py -q -m gpu > $O/pytest_peers.log 2>&1; tail -3 $O/pytest_peers.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err; head -c 300 $O/bench_loop_n1.json; echo
