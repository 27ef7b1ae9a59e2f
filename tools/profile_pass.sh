#!/bin/bash
# One profiling pass for profiles/: the ncu launch list of the bench command plus one `ncu --set full` capture per kernel
# family (raw page + the summary tools/ncu_keys.py prints; source page for the two issue-bound kernels).
#   gpurun --timeout 1500 -- 'bash tools/profile_pass.sh <tag>'
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --extra --no-cpu --no-e2e --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches.csv \
    -k regex:'local_|bilateral_kernel|harris_|pyr_|reduce_|point_|binning_|halo_' $BENCH > $OUT/ncu_launches.log 2>&1
capture() {  # kernel regex, launches to skip, tag, source page?
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c 1 -f -o $OUT/full_$3 $BENCH > $OUT/ncu_$3.log 2>&1
  ncu -i $OUT/full_$3.ncu-rep --page raw --csv > $OUT/full_$3.raw.csv 2>/dev/null
  [ "$4" = "src" ] && ncu -i $OUT/full_$3.ncu-rep --page source --csv > $OUT/full_$3.source.csv 2>/dev/null
  python tools/ncu_keys.py $OUT/full_$3.raw.csv > $OUT/full_$3.txt 2>&1
  [ "$4" = "src" ] && python tools/ncu_opmix.py $OUT/full_$3.source.csv > $OUT/full_$3.opmix.txt 2>&1
  rm -f $OUT/full_$3.ncu-rep   # gpurun copies at most 64 MiB back; the raw / source CSV pages carry what profiles/ keeps
  echo "== $3"; grep -E "Kernel Name|gpu__time_duration|dram__bytes|issue_active|registers_per_thread" $OUT/full_$3.txt
}
capture local_tma_f32_kernel 3 local_tma
capture local_pair_kernel 3 local_pair src
capture harris_fused3 1 harris_fused3 src
capture pyr_down_fused 0 pyr_down_fused_L0
capture pyr_up_half 6 pyr_up_half_L0
capture reduce_mms 2 reduce_mms
capture reduce_int 1 reduce_int
capture point_stream 2 point_stream
capture binning_kernel 2 binning
capture bilateral_kernel 0 bilateral
ls -la $OUT
