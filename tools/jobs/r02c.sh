#!/bin/bash
# Round-2 GPU job C: tile-rows A/B on dense flows, from_poses breakdown under ncu, latency sequence.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
tail -5 $O/r02c_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
for r in 8 12; do JAF_WF_WIDE_ROWS_PER_CTA=$r $B > $O/r02c_bench_dense_rows$r.json 2>> $O/r02c_err.log; done
JAF_WF_WIDE_ROWS_PER_CTA=4 $B --flow hard > $O/r02c_bench_hard_rows4.json 2>> $O/r02c_err.log
JAF_WF_WIDE_ROWS_PER_CTA=8 $B --workload scaled_512_k8_c64 > $O/r02c_bench_512k8_rows8.json 2>> $O/r02c_err.log
JAF_WF_WIDE_ROWS_PER_CTA=8 $B --workload scaled_512_k8_c64 --flow hard > $O/r02c_bench_512k8_hard_rows8.json 2>> $O/r02c_err.log
$B --flow smpl > $O/r02c_bench_smpl.json 2>> $O/r02c_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency --no-cpu > $O/r02c_bench_c1_latency.json 2>> $O/r02c_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_raster|k_warp_fuse" --csv --log-file $O/r02c_launches_from_poses.csv \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02c_prof_from_poses.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_raster_scatter|k_warp_fuse_nhwc" -s 4 -c 2 -o $O/r02c_from_poses_full -f \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02c_ncu_from_poses.log 2>&1
for f in $O/r02c_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
    if "latency_us" in d: print(json.dumps(d["latency_us"]))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -5 $O/r02c_err.log
