#!/bin/bash
# Round-2 GPU job N: persistent swapped grouped ConvLSTM (flattened K units, dense weight rows): parity + pyramid timing.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "convlstm" > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02n_pytest.log
tail -15 $O/r02n_pytest.log
timeout 600 python tools/bench_convlstm_small.py > $O/r02n_convlstm.jsonl 2>> $O/r02n_err.log
python - $O/r02n_convlstm.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:48], d.get("grouped_tc_ms"), d.get("grouped_ms"), d.get("max_abs_err_vs_fp64"), d.get("TFLOPs_useful"))
PY
tail -5 $O/r02n_err.log
