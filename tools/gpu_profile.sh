#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run, then --set full on the fused SA kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --batch 1024"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --quick --batch 512"
timeout 300 $CMD2 > gpurun_out/plain_full.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sa_tc -s 9 -c 3 -f -o gpurun_out/prof_tc $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
