#!/bin/bash
# quick iteration on the tensor-core SA kernels: bf16 parity subset (own process, own timeout), then a short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "bf16" > gpurun_out/t_bf16.log 2>&1; echo "bf16 kernels rc=$?"; tail -n 6 gpurun_out/t_bf16.log
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "bf16" > gpurun_out/t_bf16m.log 2>&1; echo "bf16 models rc=$?"; tail -n 4 gpurun_out/t_bf16m.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick ${BENCH_ARGS} > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
    for k,v in d["kernels"].items(): print(f"  {k:50s} {v['ms_per_launch']:8.3f} ms  {100*v['share']:5.1f}%")
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
