#!/bin/bash
# bf16 streaming kernel (sa3): TMA multicast of the weight chunks over clusters of 2 / 4 / 8 CTAs vs none
mkdir -p gpurun_out
for cs in 1 2 4 8; do
  PN2_TC_CLUSTER=$cs timeout 90 python bench.py --config 0 --quick --steps 10 --warmup 3 > gpurun_out/r02zd_cs$cs.json 2> gpurun_out/r02zd_cs$cs.err; echo "cs=$cs rc=$?"
done
