#!/bin/bash
# Round-2 GPU job L: operand-swapped grouped ConvLSTM kernel (parity vs fp64 + pyramid timing, A/B against the old kernel).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "convlstm" > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -15 $O/r02l_pytest.log
timeout 600 python tools/bench_convlstm_small.py > $O/r02l_convlstm_swap.jsonl 2>> $O/r02l_err.log
JAF_CG_SWAP=0 timeout 600 python tools/bench_convlstm_small.py > $O/r02l_convlstm_noswap.jsonl 2>> $O/r02l_err.log
for f in swap noswap; do echo "== $f"; python - $O/r02l_convlstm_$f.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["config"][:48], d.get("grouped_tc_ms"), d.get("grouped_ms"), d.get("max_abs_err_vs_fp64"), d.get("TFLOPs_useful"))
PY
done
tail -5 $O/r02l_err.log
