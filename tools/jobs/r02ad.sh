#!/bin/bash
# Round-2 GPU job AD: no-shuffle flavour of the wide kernel (every lane prepares all K references): parity, then same-box A/B.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
JAF_WF_WIDE_NOSHFL=1 timeout 900 python -m pytest tests -m gpu -x -q -k "wide_lane or random_configurations or hot_kernel or k1_is_exactly or host_pipeline" > $O/r02ad_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02ad_pytest.log
tail -6 $O/r02ad_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for v in 0 1 3 0 1; do
  for w in "--flow dense" "--flow hard"; do
    echo "== JAF_WF_WIDE_NOSHFL=$v $w"
    JAF_WF_WIDE_NOSHFL=$v $B $w 2>> $O/r02ad_err.log | tee -a $O/r02ad_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"
  done
done
tail -5 $O/r02ad_err.log
