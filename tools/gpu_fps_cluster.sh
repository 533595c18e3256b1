#!/bin/bash
# cluster FPS variants (PN2_FPS_CLUSTER, see csrc/pn2_fps.cu): parity subset under each, then the configs[2] microbench
mkdir -p gpurun_out
for v in 4 default; do
  if [ $v == default ]; then unset PN2_FPS_CLUSTER; else export PN2_FPS_CLUSTER=$v; fi
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fps" > gpurun_out/t_fps_v$v.log 2>&1; echo "variant $v fps tests rc=$?"; tail -n 3 gpurun_out/t_fps_v$v.log
done
for v in 4 2 default; do
  if [ $v == default ]; then unset PN2_FPS_CLUSTER; else export PN2_FPS_CLUSTER=$v; fi
  timeout 120 python tools/fps_micro.py
done 2>&1 | tee gpurun_out/fps_micro.txt
