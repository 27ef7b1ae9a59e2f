#!/bin/bash
# Short GPU pass: probes, parity tests, bench lines (+ optional ncu launch list).  gpurun --timeout 1200 -- 'bash tools/gpu_pass2.sh <tag> [ncu]'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for p in tools/probe/fp32x2_probe; do [ -x $p ] && timeout 60 $p > $OUT/$(basename $p).txt 2>&1; done
cat $OUT/fp32x2_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 2500 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --extra --no-cpu > $OUT/bench_extra.json 2> $OUT/bench_extra.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_extra.json"))
    for k, v in d.get("operators", {}).items():
        print(f"{k:40s} {v['Gpx_s']:9.1f} Gpx/s {v['ms']:8.3f} ms  hbm_frac {v['hbm_frac']:.3f}")
except Exception as e:
    print("bench_extra parse failed", e)
PY
tail -3 $OUT/bench_extra.err
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches.csv \
    -k regex:'local_|bilateral_kernel|harris_|pyr_|reduce_|point_|hist' python bench.py --steps 2 --warmup 3 --extra --no-cpu --no-e2e > $OUT/ncu_launches.log 2>&1
fi
ls -la $OUT
