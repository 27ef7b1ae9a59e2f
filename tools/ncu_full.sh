#!/bin/bash
# Full ncu captures (one launch each) of the kernels behind bench.py --extra; raw pages exported on the box.
# gpurun --timeout 1500 -- 'bash tools/ncu_full.sh <tag> "<kernel regex>"'
TAG=${1:-r1}
RE=${2:-pyr_down_fused|pyr_up_half|harris_fused|bilateral_kernel|local_tiled_kernel|binning_kernel|point_stream|reduce_mms}
OUT=gpurun_out/$TAG
mkdir -p $OUT
IFS='|' read -ra KS <<< "$RE"
for k in "${KS[@]}"; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s ${NCU_SKIP:-3} -c 1 -f -o $OUT/full_$k \
      python bench.py --steps 1 --warmup 3 --extra --no-cpu --no-e2e --no-graph > $OUT/ncu_$k.log 2>&1
  ncu -i $OUT/full_$k.ncu-rep --page raw --csv > $OUT/full_$k.raw.csv 2>/dev/null
  python tools/ncu_keys.py $OUT/full_$k.raw.csv > $OUT/full_$k.txt 2>&1
  echo "== $k"; head -40 $OUT/full_$k.txt
done
ls -la $OUT
