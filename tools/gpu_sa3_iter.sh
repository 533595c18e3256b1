#!/bin/bash
# wide-stack (ring) kernel iteration: parity of the group_all cases under each cluster size, then timing
for cs in 1 4; do
  echo "== PN2_TC_CLUSTER=$cs"
  PN2_TC_CLUSTER=$cs timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "bf16 and (sa3 or trans_sa2 or msg_all or odd or bigk or msg_sa2_k128)" 2>&1 | tail -2
  PN2_TC_CLUSTER=$cs timeout 120 python bench.py --steps 5 --warmup 3 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), [round(v['ms_per_launch'],3) for k,v in d['kernels'].items() if k.startswith('sa_fused')])"
done
