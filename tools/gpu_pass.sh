#!/bin/bash
# One GPU pass: parity tests, smoke, bench, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout 300 python bench.py --steps 2 --warmup 3 --quick --batch 1024 > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --quick --batch 1024 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
fi
