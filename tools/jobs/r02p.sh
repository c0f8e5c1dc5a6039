#!/bin/bash
# Round-2 GPU job P: per-role timers of the un-swapped grouped ConvLSTM kernel on the two small levels (48@25^2, 96@13^2)
# + full test run and bench of the batch-1 sequence with parallel graph lanes.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
JAFPRO_B200_LIB=$PWD/jafpro_b200/libjafpro_b200_cgprof.so timeout 120 python - > $O/r02p_cgprof.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0,'.')
from jafpro_b200 import ops
for Ch, S in [(48, 25), (96, 13)]:
    x, h, c = (torch.randn(24, 1, Ch, S, S, device="cuda") for _ in range(3))
    w = torch.randn(24, 4 * Ch, 2 * Ch, 3, 3, device="cuda") * 0.05
    wp = ops.convlstm_gpack_weight(w, Ch, Ch)
    for _ in range(2):
        ops.convlstm_step_grouped(x, h, c, wp, None, Ch, Ch)
        torch.cuda.synchronize()
    print("----", Ch, S, flush=True)
PY
grep -v "^$" $O/r02p_cgprof.log | tail -30
timeout 900 python -m pytest tests -m gpu -x -q -k "frame_graph or rgb or warp_fuse" > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p_pytest.log
tail -5 $O/r02p_pytest.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency > $O/r02p_bench_c1_latency.json 2>> $O/r02p_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench_c1_latency.json').read().strip().splitlines()[-1])
print(json.dumps(d["latency_us"]))
PY
tail -5 $O/r02p_err.log
