#!/bin/bash
# Round 2, second multi-GPU call (4 GPUs): configs[3] sweep with the pipelined exchange at N = 1, 2, 4; the sharded-PCG pytest.
O=gpurun_out/r2n; mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --workload sweep --steps 48 --warmup 8 > $O/sweep_n1.json 2> $O/sweep_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --workload sweep --gpus 2 --steps 48 --warmup 8 > $O/sweep_n2.json 2> $O/sweep_n2.err
timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --workload sweep --gpus 4 --steps 48 --warmup 8 > $O/sweep_n4.json 2> $O/sweep_n4.err
timeout 600 python -m pytest tests/test_solver_peers_gpu.py -q -m gpu > $O/pytest_peers.log 2>&1; tail -2 $O/pytest_peers.log
for f in sweep_n1 sweep_n2 sweep_n4; do echo "== $f"; tail -c 300 $O/$f.err; grep -c '^{' $O/$f.json; done
