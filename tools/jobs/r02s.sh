#!/bin/bash
# Round-2 GPU job S: ncu --set full of the persistent grouped ConvLSTM kernel (12 ch @ 200^2 and 24 ch @ 100^2 launches).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_convlstm_grouped_p -s 1 -c 3 -o $O/r02s_cgp_full -f \
  python tools/probes/convlstm_grouped_once.py > $O/r02s_ncu.log 2>&1
tail -3 $O/r02s_ncu.log
for kb in 28 48 64 96; do
echo "== STAGE_KB=$kb"; JAF_CG_STAGE_KB=$kb timeout 300 python tools/bench_convlstm_small.py 2>> $O/r02s_err.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:48], d.get('grouped_tc_ms'), d.get('grouped_ms'))" | tail -3
done
