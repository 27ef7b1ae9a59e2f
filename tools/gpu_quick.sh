#!/bin/bash
# quick GPU pass: parity tests + the extra-operator bench.  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [pytest -k expr]'
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --extra --no-cpu --no-e2e > $OUT/bench_extra.json 2> $OUT/bench_extra.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_extra.json"))
    print(d["metric"], d["value"], d["roofline"]["frac"])
    for k, v in d.get("operators", {}).items():
        print(f"{k:40s} {v['Gpx_s']:9.1f} Gpx/s {v['ms']:8.3f} ms  hbm_frac {v.get('hbm_frac', 0):.3f}")
except Exception as e:
    print("bench_extra parse failed", e)
PY
tail -3 $OUT/bench_extra.err
