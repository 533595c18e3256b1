#!/bin/bash
# round-2 A/B of the streamed-weight variants (weight replicas, gather depth, split chunk width); writes gpurun_out/r02u_*
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 90 python bench.py --config 0 --quick --steps 10 --warmup 3 $EXTRA > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err; echo "$name rc=$?"; }
EXTRA="--precision bf16"
run bf16_base PN2_TC_GATHER4=0
run bf16_r8 PN2_W_REPLICAS=8 PN2_TC_GATHER4=0
run bf16_r8_g4 PN2_W_REPLICAS=8
run bf16_g4 PN2_W_REPLICAS=1
EXTRA="--precision fp32"
run fp32_r8 PN2_W_REPLICAS=8
run fp32_r8_k16 PN2_W_REPLICAS=8 PN2_F32_KBLK=16
run fp32_k16 PN2_F32_KBLK=16
(PN2_W_REPLICAS=8 PN2_F32_KBLK=16 timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "test_sa_fused_vs_oracle or test_sa_fused_bf16" > gpurun_out/r02u_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02u_parity.log)
tail -2 gpurun_out/r02u_parity.log
