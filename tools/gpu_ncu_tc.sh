#!/bin/bash
# ncu --set full capture of the fused SA tensor-core kernel (3 launches: sa1, sa2, sa3 of one forward)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --quick --batch ${1:-512}"
timeout 300 $CMD > gpurun_out/plain_full.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sa_tc -s 9 -c 3 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_full.log
