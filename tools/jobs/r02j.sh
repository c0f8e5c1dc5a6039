#!/bin/bash
# Round-2 GPU job J: new defaults (4 CTAs/SM skip flavours, 32x32 pose tiles, merged clear), RGB-only pose calls, sequence.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02j_pytest.log
tail -8 $O/r02j_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5"
$B > $O/r02j_bench_default.json 2>> $O/r02j_err.log
$B --no-cpu --flow smpl > $O/r02j_bench_smpl.json 2>> $O/r02j_err.log
$B --no-cpu --flow hard > $O/r02j_bench_hard.json 2>> $O/r02j_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c1_latency > $O/r02j_bench_c1_latency.json 2>> $O/r02j_err.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload c3_flow > $O/r02j_bench_c3_flow.json 2>> $O/r02j_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_raster|k_warp_fuse" --csv --log-file $O/r02j_launches_from_poses.csv \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02j_prof_from_poses.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_raster_scatter_flat|k_warp_fuse_nhwc" -s 4 -c 2 -o $O/r02j_from_poses_full -f \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02j_ncu_from_poses.log 2>&1
for f in $O/r02j_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
    if "latency_us" in d: print(json.dumps(d["latency_us"]))
except Exception as ex:
    print("unparsed", ex)
PY
done
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02j_launches_from_poses.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>iv:
        try: d[r[ik][:80]].append(float(r[iv].replace(',','')))
        except: pass
for k,v in d.items(): print(f"{len(v):4d} x avg {sum(v)/len(v)/1e3:9.1f} us  {k}")
PY
tail -5 $O/r02j_err.log
