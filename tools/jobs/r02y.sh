#!/bin/bash
# Round-2 GPU job Y: un-swapped grouped ConvLSTM: per-pair staging barriers, channel-major K order (staging overlaps the MMAs).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "convlstm" > $O/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y_pytest.log
tail -4 $O/r02y_pytest.log
for i in 1 2; do
timeout 600 python tools/bench_convlstm_small.py 2>> $O/r02y_err.log | tee $O/r02y_convlstm.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:48], d.get('grouped_tc_ms'), d.get('grouped_ms'), d.get('max_abs_err_vs_fp64'))"
done
tail -5 $O/r02y_err.log
