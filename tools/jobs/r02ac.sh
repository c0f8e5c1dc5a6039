#!/bin/bash
# Round-2 GPU job AC: PAIR flavour of the fused kernel (8-lane groups own two adjacent pixels, the shared tap column stays
# in registers): parity under JAF_WF_PAIR=1, then same-box A/B against the shipped wide kernel.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
JAF_WF_PAIR=1 timeout 900 python -m pytest tests -m gpu -x -q -k "wide_lane or random_configurations or hot_kernel or k1_is_exactly or host_pipeline" > $O/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ac_pytest.log
tail -6 $O/r02ac_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for v in 0 1 3 0 1; do
  for w in "--flow dense" "--flow hard"; do
    echo "== JAF_WF_PAIR=$v $w"
    JAF_WF_PAIR=$v $B $w 2>> $O/r02ac_err.log | tee -a $O/r02ac_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"
  done
done
tail -5 $O/r02ac_err.log
