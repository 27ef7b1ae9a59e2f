#!/bin/bash
# One GPU pass: parity tests, bench lines, ncu launch list and a full capture of the dominant kernels.
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_pass.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --extra --no-cpu > $OUT/bench_extra.json 2> $OUT/bench_extra.err; tail -c 4000 $OUT/bench_extra.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    -k regex:'local_|bilateral_kernel|harris_|pyr_|reduce_|point_kernel' python bench.py --steps 2 --warmup 3 --extra --no-cpu > $OUT/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'local_tma_f32_kernel|local_tiled_kernel' -s 9 -c 3 \
    -o $OUT/prof_local python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
ls -la $OUT
