#!/bin/bash
# Full ncu capture WITH the source page exported (per-line instruction / stall counters) for the issue-bound kernels.
# gpurun --timeout 1500 -- 'bash tools/ncu_src.sh <tag> "<kernel regex>|<kernel regex>..."'
TAG=${1:-r2}
RE=${2:-local_tiled_kernel|harris_fused3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
IFS='|' read -ra KS <<< "$RE"
for k in "${KS[@]}"; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$k" -s ${NCU_SKIP:-3} -c 1 -f -o $OUT/full_$k \
      python bench.py --steps 1 --warmup 3 --extra --no-cpu --no-e2e --no-graph > $OUT/ncu_$k.log 2>&1
  ncu -i $OUT/full_$k.ncu-rep --page raw --csv > $OUT/full_$k.raw.csv 2>/dev/null
  ncu -i $OUT/full_$k.ncu-rep --page source --csv > $OUT/full_$k.source.csv 2>/dev/null
  python tools/ncu_keys.py $OUT/full_$k.raw.csv > $OUT/full_$k.txt 2>&1
  echo "== $k"; head -40 $OUT/full_$k.txt
done
ls -la $OUT
