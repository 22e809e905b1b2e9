#!/bin/bash
# Round 2, multi-GPU call (4 GPUs): matcher with 8 read-back warps (tests + timing); sharded PCG check at 2 and 4 ranks; configs[3] sweep at N = 1, 2, 4;
# frame-loop bench at N = 2, 4.
O=gpurun_out/r2m; mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_sift_gpu.py -x -q -m gpu > $O/pytest_sift_tma.log 2>&1; echo "sift tma rc=$?"; tail -2 $O/pytest_sift_tma.log
timeout 300 python scripts/sift_match_timing.py > $O/sift_match_timing_tma.jsonl 2> $O/sift_match_timing_tma.err; cat $O/sift_match_timing_tma.jsonl
timeout 300 $TR --nproc-per-node 2 --master-port 29521 scripts/solver_peers_check.py > $O/peers_n2.jsonl 2> $O/peers_n2.err; echo "peers2 rc=$?"; cat $O/peers_n2.jsonl; tail -c 400 $O/peers_n2.err
timeout 300 $TR --nproc-per-node 4 --master-port 29522 scripts/solver_peers_check.py > $O/peers_n4.jsonl 2> $O/peers_n4.err; echo "peers4 rc=$?"; cat $O/peers_n4.jsonl; tail -c 400 $O/peers_n4.err
timeout 600 python bench.py --workload sweep --steps 48 --warmup 8 > $O/sweep_n1.json 2> $O/sweep_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29511 bench.py --workload sweep --gpus 2 --steps 48 --warmup 8 > $O/sweep_n2.json 2> $O/sweep_n2.err
timeout 600 $TR --nproc-per-node 4 --master-port 29512 bench.py --workload sweep --gpus 4 --steps 48 --warmup 8 > $O/sweep_n4.json 2> $O/sweep_n4.err
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 900 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err
for f in sweep_n1 sweep_n2 sweep_n4 bench_n2 bench_n4; do echo "== $f"; tail -c 300 $O/$f.err; head -c 160 $O/$f.json; echo; done
