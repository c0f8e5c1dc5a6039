#!/bin/bash
# Round-2 GPU job R: un-swapped grouped ConvLSTM kernel with a dedicated weight-stream warp + deep ring (A/B by environment).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "convlstm" > $O/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02r_pytest.log
tail -4 $O/r02r_pytest.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python tools/bench_convlstm_small.py > $O/r02r_convlstm_$name.jsonl 2>> $O/r02r_err.log
  echo "== $name ($*)"
  python - $O/r02r_convlstm_$name.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:48], d.get("grouped_tc_ms"), d.get("grouped_ms"), d.get("max_abs_err_vs_fp64"))
PY
}
run default X=1
run stage8 JAF_CG_STAGE_KB=8
tail -5 $O/r02r_err.log
run stage12 JAF_CG_STAGE_KB=12
run ring3 JAF_CG_RING=3
