#!/bin/bash
# Multi-GPU pass: P2P halo test + bench at N ranks (default line + extras).  gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh <tag> <N>'
TAG=${1:-r2n}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/test_p2p_halo.py > $OUT/p2p_halo_test.log 2>&1; echo "rc=$?" >> $OUT/p2p_halo_test.log
grep -E "P2P_HALO|rc=" $OUT/p2p_halo_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_FLAGS} > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"
tail -c 4000 $OUT/bench_n$N.json; grep -v "^W\|^\[W\|warn" $OUT/bench_n$N.err | tail -15
