#!/bin/bash
# Round-2 GPU job V: full GPU test suite + smoke + the driver's bench command (e2e step = the whole 240-frame step).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02v_pytest.log
tail -6 $O/r02v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/r02v_bench_default.json 2>> $O/r02v_err.log
timeout 600 python bench.py --no-cpu --flow hard --steps 20 > $O/r02v_bench_hard.json 2>> $O/r02v_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c4_convlstm > $O/r02v_bench_c4_convlstm.json 2>> $O/r02v_err.log
for f in $O/r02v_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{}); e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), e.get("frames_per_step"), "per_frame_refs", (e.get("per_frame_refs") or {}).get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -5 $O/r02v_err.log
