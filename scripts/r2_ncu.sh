#!/bin/bash
# ncu captures of one bench step, exported as CSV on the GPU box (the .ncu-rep files are too large to bring back).
# usage: scripts/r2_ncu.sh <tag> [source-kernel-skip-count ...]
TAG=$1; shift
mkdir -p gpurun_out
# every tcgen05 conv launch of the 2nd step: --set full, raw page
ncu --set full --clock-control none -k regex:conv_tc -s 14 -c 14 --csv --page raw python scripts/profile_step.py 2 > gpurun_out/${TAG}_raw.csv 2> gpurun_out/${TAG}_ncu.log
for SKIP in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:halo2 -s $SKIP -c 1 -o /tmp/${TAG}_src_$SKIP python scripts/profile_step.py 2 >> gpurun_out/${TAG}_ncu.log 2>&1
  ncu -i /tmp/${TAG}_src_$SKIP.ncu-rep --page source --csv > gpurun_out/${TAG}_src_$SKIP.csv 2>> gpurun_out/${TAG}_ncu.log
  rm -f /tmp/${TAG}_src_$SKIP.ncu-rep
done
ls -la gpurun_out/${TAG}_*
