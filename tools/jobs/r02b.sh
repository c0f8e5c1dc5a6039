#!/bin/bash
# Round-2 GPU job B: parity suite with the pose-driven kernel + host pipelines, from_poses A/B, latency workload, ncu.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
tail -15 $O/r02b_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5"
$B > $O/r02b_bench_default.json 2> $O/r02b_bench_default.err
JAF_WF_MINB_POSES=4 $B --no-cpu > $O/r02b_bench_poses_minb4.json 2>> $O/r02b_err.log
JAF_WF_ROWS_PER_CTA=8 $B --no-cpu > $O/r02b_bench_poses_rows8.json 2>> $O/r02b_err.log
JAF_WF_ROWS_PER_CTA=32 $B --no-cpu > $O/r02b_bench_poses_rows32.json 2>> $O/r02b_err.log
$B --no-cpu --flow smpl > $O/r02b_bench_smpl.json 2>> $O/r02b_err.log
JAF_WF_WIDE_ROWS_PER_CTA=8 $B --no-cpu --flow hard > $O/r02b_bench_hard_rows8.json 2>> $O/r02b_err.log
JAF_WF_WIDE_ROWS_PER_CTA=32 $B --no-cpu --flow hard > $O/r02b_bench_hard_rows32.json 2>> $O/r02b_err.log
$B --no-cpu --flow hard > $O/r02b_bench_hard.json 2>> $O/r02b_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency > $O/r02b_bench_c1_latency.json 2>> $O/r02b_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file $O/r02b_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02b_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_warp_fuse_nhwc" -s 40 -c 1 -o $O/r02b_poses_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02b_ncu_poses.log 2>&1
for f in $O/r02b_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), (e.get("application") or {}).get("matches_device_path"), "from_poses", (d.get("from_poses") or {}).get("value"), (d.get("from_poses") or {}).get("kernel"), d.get("clocks"))
    if "latency_us" in d: print(json.dumps(d["latency_us"]))
    if e.get("pinned_memcpy_probe"): print(e["pinned_memcpy_probe"])
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -5 $O/r02b_err.log
