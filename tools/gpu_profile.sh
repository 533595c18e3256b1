#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run, then --set full on one forward's seven product kernels
# (FPS, ball query, fused SA x3 ...) at the bench batch.  Each ncu pass only after the same command has exited 0 alone.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --batch 1024"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --quick --batch ${1:-2048}"
timeout 300 $CMD2 > gpurun_out/plain_full.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'sa_tc|fps_kernel|ball_query_kernel' -s 21 -c 7 -f -o gpurun_out/prof_all $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null; echo "raw rc=$?"
