for c in 10 12 14 16 18 20 24; do
  echo -n "ctas_per_sm=$c: "
  HB_STREAM_CTAS_PER_SM=$c timeout 100 python bench.py --steps 3 --warmup 3 --extra --no-cpu --no-e2e | python -c "
import sys,json; d=json.loads(sys.stdin.read())['operators']
print(' '.join(f\"{k}={d[k]['Gpx_s']:.0f}\" for k in ('C3_reduce_minmaxsum_f32_8192','hist256_f32_8192')))"
done
