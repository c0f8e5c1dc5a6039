#!/bin/bash
# Round-2 GPU job K: grouped ConvLSTM launch shapes (quarter-SM CTAs), batch-1 sequence with small pose tiles.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "convlstm or from_poses or frame_graph" > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest.log
tail -4 $O/r02k_pytest.log
for mode in 0 1 2 3; do
  JAF_CG_MODE=$mode timeout 600 python tools/bench_convlstm_small.py > $O/r02k_convlstm_mode$mode.jsonl 2>> $O/r02k_err.log
  echo "== mode $mode"; python - $O/r02k_convlstm_mode$mode.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:48], d.get("grouped_tc_ms"), d.get("grouped_ms"), d.get("max_abs_err_vs_fp64"))
PY
done
JAF_CG_AUTO_QUARTER=1 timeout 600 python tools/bench_convlstm_small.py > $O/r02k_convlstm_autoq.jsonl 2>> $O/r02k_err.log
echo "== auto quarter"; tail -1 $O/r02k_convlstm_autoq.jsonl
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency --no-cpu > $O/r02k_bench_c1_latency.json 2>> $O/r02k_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_c1_latency.json').read().strip().splitlines()[-1]); print(json.dumps(d["latency_us"]))
PY
tail -5 $O/r02k_err.log
