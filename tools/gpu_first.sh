#!/bin/bash
# first GPU pass: each risky group in its own process under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not bf16 and not umma" > gpurun_out/t_kernels_f32.log 2>&1; echo "kernels_f32 rc=$?"
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "umma" > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "bf16" > gpurun_out/t_kernels_bf16.log 2>&1; echo "kernels_bf16 rc=$?"
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "not bf16" > gpurun_out/t_models_f32.log 2>&1; echo "models_f32 rc=$?"
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -k "bf16" > gpurun_out/t_models_bf16.log 2>&1; echo "models_bf16 rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -n 5 gpurun_out/t_*.log gpurun_out/smoke.log
