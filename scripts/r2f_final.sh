#!/bin/bash
# Final evidence run of round 2 (GPU box): ncu --set full of one step's tcgen05 launches, ncu launch list of the bench
# command, the GPU test suite, smoke(), the bench line and the isolated-UMMA table.  Outputs under gpurun_out/r2f_*.
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:conv_tc -s 14 -c 14 --csv --page raw python scripts/profile_step.py 2 > gpurun_out/r2f_raw.csv 2> gpurun_out/r2f_ncu.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 2 > gpurun_out/r2f_bench_under_ncu.json 2>> gpurun_out/r2f_ncu.log
(timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12) > gpurun_out/r2f_pytest.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 100 python scripts/umma_probe.py > gpurun_out/r2f_umma_probe.txt 2>&1
tail -3 gpurun_out/r2f_pytest.log; tail -2 gpurun_out/r2f_smoke.log; ls -la gpurun_out/r2f_*; tail -c 300 gpurun_out/r2f_bench.json
