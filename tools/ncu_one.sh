#!/bin/bash
# One `ncu --set full` capture (raw + source pages, summary, opcode mix) of one kernel while a python command runs.
#   gpurun --timeout 600 -- 'bash tools/ncu_one.sh <tag> <kernel regex> <launches to skip> <name> python tools/ab_harris.py x'
TAG=$1; RE=$2; SKIP=$3; NAME=$4; shift 4
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -f -o $OUT/full_$NAME "$@" > $OUT/ncu_$NAME.log 2>&1
ncu -i $OUT/full_$NAME.ncu-rep --page raw --csv > $OUT/full_$NAME.raw.csv 2>/dev/null
ncu -i $OUT/full_$NAME.ncu-rep --page source --csv > $OUT/full_$NAME.source.csv 2>/dev/null
python tools/ncu_keys.py $OUT/full_$NAME.raw.csv > $OUT/full_$NAME.txt 2>&1
python tools/ncu_opmix.py $OUT/full_$NAME.source.csv > $OUT/full_$NAME.opmix.txt 2>&1
rm -f $OUT/full_$NAME.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes|issue_active|registers_per_thread|warps_active" $OUT/full_$NAME.txt
head -30 $OUT/full_$NAME.opmix.txt
