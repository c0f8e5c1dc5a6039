#!/bin/bash
# Round-2 GPU job Z: what-if probes: se gather dropped but its unpack kept (taps4) / se gather kept but its unpack shared (taps5).
# Upper bound of what ANY tap-sharing scheme (x-neighbour shuffles, pixel pairs, row reuse) could buy.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for v in base taps4 taps5 base taps4 taps5; do
  L=$PWD/jafpro_b200/libjafpro_b200.so; [ $v != base ] && L=$PWD/jafpro_b200/libjafpro_b200_$v.so
  for w in "--flow dense" "--flow hard"; do
    echo "== $v $w"
    JAFPRO_B200_LIB=$L $B $w 2>> $O/r02z_err.log | tee -a $O/r02z_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"
  done
done
tail -5 $O/r02z_err.log
