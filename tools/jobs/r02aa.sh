#!/bin/bash
# Round-2 GPU job AA: feature taps shared between x-adjacent pixel groups by shuffle (per-group coherence test, predicated
# fallback gathers) in the wide kernel: parity + A/B (xtap = with the reference loop not unrolled, unroll1 = that alone,
# xtapfull = unrolled, spills).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
JAFPRO_B200_LIB=$PWD/jafpro_b200/libjafpro_b200_xtap.so timeout 900 python -m pytest tests -m gpu -x -q -k "warp_fuse or full_size" > $O/r02aa_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02aa_pytest.log
tail -4 $O/r02aa_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for v in base xtap unroll1 xtapfull base xtap; do
  L=$PWD/jafpro_b200/libjafpro_b200.so; [ $v != base ] && L=$PWD/jafpro_b200/libjafpro_b200_$v.so
  for w in "--flow dense" "--flow hard"; do
    echo "== $v $w"
    JAFPRO_B200_LIB=$L $B $w 2>> $O/r02aa_err.log | tee -a $O/r02aa_ab.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline',{})
print(d.get('value'), d.get('ms_per_step'), 'frac', r.get('frac'), r.get('kernel'), (d.get('clocks') or {}).get('sm_mhz'))"
  done
done
tail -5 $O/r02aa_err.log
