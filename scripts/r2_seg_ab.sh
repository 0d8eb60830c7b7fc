#!/bin/bash
# per-layer promotion-period variants (DCSCN_SEG) on the bench workload: step time, per-launch times, error vs fp64
mkdir -p gpurun_out
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 && echo SMOKE_OK) || { echo SMOKE_FAIL; tail -5 gpurun_out/smoke.log; exit 1; }
echo "== two-pass A1+B1"; DCSCN_PAIR_STREAM=0 ORACLE_TILES=3 timeout 200 python scripts/r2_halo_ab.py halo=3 2>&1 | tail -2
for S in "" "A1+B1=2" "A1+B1=3" "A1+B1=6" "A1+B1=8" "A1+B1=11" "A1+B1=21" ""; do
  echo "== SEG=$S"
  DCSCN_SEG="$S" ORACLE_TILES=3 timeout 200 python scripts/r2_halo_ab.py halo=3 2>&1 | tail -2
done | tee gpurun_out/seg16.log
