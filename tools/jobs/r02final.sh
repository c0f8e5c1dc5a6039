#!/bin/bash
# Round-2 final GPU job (1 GPU): full GPU test suite, smoke, every bench workload with the final build, the CPU arm,
# the launch list of a bench run and of the pose-driven step.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02final_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02final_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02final_pytest.log
tail -5 $O/r02final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="timeout 900 python bench.py"
$B > $O/r02final_bench_default.json 2>> $O/r02final_err.log
$B --impl reference --steps 3 --warmup 1 > $O/r02final_bench_reference.json 2>> $O/r02final_err.log
$B --no-cpu --steps 20 --flow hard > $O/r02final_bench_hard.json 2>> $O/r02final_err.log
$B --no-cpu --steps 20 --flow smpl > $O/r02final_bench_smpl.json 2>> $O/r02final_err.log
$B --no-cpu --steps 20 --flow perm > $O/r02final_bench_perm.json 2>> $O/r02final_err.log
$B --steps 10 --workload scaled_512_k8_c64 > $O/r02final_bench_512k8.json 2>> $O/r02final_err.log
$B --steps 20 --workload rgb_only_256_k4 > $O/r02final_bench_rgbonly.json 2>> $O/r02final_err.log
$B --steps 10 --warmup 5 --workload c1_latency > $O/r02final_bench_c1_latency.json 2>> $O/r02final_err.log
$B --steps 10 --warmup 5 --workload c3_flow > $O/r02final_bench_c3_flow.json 2>> $O/r02final_err.log
$B --steps 10 --warmup 5 --workload c4_convlstm > $O/r02final_bench_c4_convlstm.json 2>> $O/r02final_err.log
timeout 600 python tools/bench_convlstm_small.py > $O/r02final_convlstm_grouped.jsonl 2>> $O/r02final_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_raster|k_warp_fuse|k_convlstm|k_flow|k_mask" -c 400 --csv --log-file $O/r02final_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02final_bench_under_ncu.log 2>&1
for f in $O/r02final_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(d.get("impl"), d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
    if "latency_us" in d: print(json.dumps(d["latency_us"]))
except Exception as ex:
    print("unparsed", ex)
PY
done
python - <<'PY'
import json
for l in open('gpurun_out/r02final_convlstm_grouped.jsonl'):
    d=json.loads(l); print(d["config"][:48], d.get("grouped_tc_ms"), d.get("grouped_ms"), d.get("torch_ref_ops_tf32_ms"), d.get("torch_tf32_ms"))
PY
tail -5 $O/r02final_err.log
