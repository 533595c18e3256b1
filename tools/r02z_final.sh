#!/bin/bash
# round-2 final pass on one GPU: the whole -m gpu suite, smoke, every single-GPU bench config, the ncu launch list
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02z_tests.log)
tail -3 gpurun_out/r02z_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 150 python bench.py > gpurun_out/r02z_bench_cfg0.json 2> gpurun_out/r02z_bench_cfg0.err; echo "cfg0 rc=$?"
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_ref_cfg0.json 2> gpurun_out/r02z_ref_cfg0.err; echo "ref rc=$?"
timeout 100 python bench.py --config 2 > gpurun_out/r02z_bench_cfg2.json 2> gpurun_out/r02z_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 150 python bench.py --config 3 > gpurun_out/r02z_bench_cfg3.json 2> gpurun_out/r02z_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 150 python bench.py --config 4 > gpurun_out/r02z_bench_cfg4.json 2> gpurun_out/r02z_bench_cfg4.err; echo "cfg4 rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_bench_b1024.csv python bench.py --batch 1024 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/r02z_ncu_bench.log 2>&1; echo "launch list rc=$?"
