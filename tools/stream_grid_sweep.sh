#!/bin/bash
# sweep HB_STREAM_CTAS_PER_SM (0 = one CTA per chunk) for the streaming kernels (point copy, reduction, histogram)
for c in 8 16 32 64 0; do
  echo -n "ctas_per_sm=$c: "
  HB_STREAM_CTAS_PER_SM=$c timeout 200 python bench.py --steps 3 --warmup 3 --extra --no-cpu --no-e2e | python -c "
import sys,json; d=json.loads(sys.stdin.read())['operators']
print(' '.join(f\"{k.split('_')[0]+'_'+k.split('_')[1]}={d[k]['Gpx_s']:.0f}\" for k in ('point_copy_f32_8192','C3_reduce_minmaxsum_f32_8192','hist256_f32_8192','C5_pyramid8_f32_16384')))"
done
