#!/bin/bash
# Round-2 GPU job D: flattened raster scatter + compacted pose phase 0 (parity + A/B), get_texture parity, tile rows.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
tail -8 $O/r02d_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
$B > $O/r02d_bench_default.json 2>> $O/r02d_err.log
JAF_RASTER_FLAT=0 $B > $O/r02d_bench_flat0.json 2>> $O/r02d_err.log
JAF_WF_WIDE_ROWS_PER_CTA=4 $B > $O/r02d_bench_dense_rows4.json 2>> $O/r02d_err.log
JAF_WF_WIDE_ROWS_PER_CTA=8 $B --flow perm > $O/r02d_bench_perm_rows8.json 2>> $O/r02d_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c3_flow --no-cpu > $O/r02d_bench_c3_flow.json 2>> $O/r02d_err.log
JAF_RASTER_FLAT=0 timeout 600 python bench.py --steps 10 --warmup 5 --workload c3_flow --no-cpu > $O/r02d_bench_c3_flow_flat0.json 2>> $O/r02d_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency --no-cpu > $O/r02d_bench_c1_latency.json 2>> $O/r02d_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_raster|k_warp_fuse" --csv --log-file $O/r02d_launches_from_poses.csv \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02d_prof_from_poses.log 2>&1
for f in $O/r02d_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
    if "latency_us" in d: print(json.dumps(d["latency_us"]))
    if "reference_cuda" in d: print(d["reference_cuda"])
except Exception as ex:
    print("unparsed", ex)
PY
done
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02d_launches_from_poses.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>iv:
        try: d[r[ik][:80]].append(float(r[iv].replace(',','')))
        except: pass
for k,v in d.items(): print(f"{len(v):4d} x avg {sum(v)/len(v)/1e3:9.1f} us  {k}")
PY
tail -5 $O/r02d_err.log
