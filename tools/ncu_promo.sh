#!/bin/bash
export PYTHONPATH=.
mkdir -p gpurun_out/promo
for promo in 0 2 3; do
  HB_TMA_L2PROMO=$promo ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'local_tma' -s 9 -c 3 --csv --log-file gpurun_out/promo/p$promo.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
  echo promo=$promo; grep -E "local_tma" gpurun_out/promo/p$promo.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste - - - - - | head -3
done
ncu --set full --clock-control none --import-source on -k regex:'local_tma' -s 9 -c 3 -o gpurun_out/promo/prof_v3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/promo
