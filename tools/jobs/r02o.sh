#!/bin/bash
# Round-2 GPU job O: RGB taps shared between x-adjacent lanes by shuffle (A/B against the previous build on the same box).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02o_pytest.log
tail -8 $O/r02o_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for v in new base new base; do
  L=$PWD/jafpro_b200/libjafpro_b200.so; [ $v = base ] && L=$PWD/jafpro_b200/libjafpro_b200_base.so
  for w in "--flow dense" "--flow hard" "--workload rgb_only_256_k4" "--workload rgb_only_256_k4 --flow hard" "--workload scaled_512_k8_c64"; do
    echo "== $v $w"
    JAFPRO_B200_LIB=$L $B $w 2>> $O/r02o_err.log | tee -a $O/r02o_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"
  done
done
tail -5 $O/r02o_err.log
