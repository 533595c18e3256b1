#!/bin/bash
# round-2 GPU pass: multi-radius ball query parity, configs[1] bench, ncu launch list + full capture of one bf16 forward
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bench_shapes.py tests/test_gpu_models.py -m gpu -x -q -k "multi_radius or batched_launch_shapes or per_trip or msg or few_points" > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02y_tests.log)
tail -3 gpurun_out/r02y_tests.log
timeout 120 python bench.py --config 1 > gpurun_out/r02y_bench_cfg1.json 2> gpurun_out/r02y_bench_cfg1.err; echo "cfg1 rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches_bench_b1024.csv python bench.py --batch 1024 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r02y_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|ball_query|sa_tc|fc_head" --launch-skip 8 --launch-count 8 -f -o gpurun_out/r02y_full python tools/prof_case.py bf16 2048 pointnet2_cls_ssg 2 > gpurun_out/r02y_ncu_full.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r02y_full.ncu-rep --page raw --csv > gpurun_out/r02y_full_raw.csv 2>/dev/null; ls -la gpurun_out/r02y_full*
