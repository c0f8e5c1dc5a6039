#!/bin/bash
# Round-2 GPU job U: K = 5..8 split-K wide kernel (8-lane groups, halves split the references): parity + 512^2 K=8 A/B.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "warp_fuse or full_size" > $O/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02u_pytest.log
tail -6 $O/r02u_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload scaled_512_k8_c64"
run() { local name=$1; shift; echo "== $name"; env "$@" $B $FL 2>> $O/r02u_err.log | tee -a $O/r02u_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"; }
for FL in "--flow dense" "--flow hard"; do
echo "#### $FL"
run splitk_rows16 X=1
run wide2 JAF_WF_WIDE8_SPLITK=0
run splitk_rows8 JAF_WF_SPLITK_ROWS=8
run splitk_rows32 JAF_WF_SPLITK_ROWS=32
run splitk_rows16_again X=1
run wide2_again JAF_WF_WIDE8_SPLITK=0
done
tail -5 $O/r02u_err.log
