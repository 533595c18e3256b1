#!/bin/bash
# launch list of the bench on the final build + ncu --set full of the training BatchNorm kernels (one layer, tools/bn_probe.py)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick --batch 1024"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/bn_probe.py 5 > gpurun_out/bn_probe.txt 2>&1 && cat gpurun_out/bn_probe.txt &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_fwd|bn_bwd' -s 12 -c 12 -f -o gpurun_out/prof_bn python tools/bn_probe.py 1 > gpurun_out/ncu_bn.log 2>&1
echo "bn full rc=$?"; tail -n 2 gpurun_out/ncu_bn.log
ncu -i gpurun_out/prof_bn.ncu-rep --page raw --csv > gpurun_out/prof_bn_raw.csv 2>/dev/null; echo "raw rc=$?"
