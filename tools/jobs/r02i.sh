#!/bin/bash
# Round-2 GPU job I: list-driven phases for pixel-level visibility (parity + from_poses / smpl A/B).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest.log
tail -8 $O/r02i_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
$B > $O/r02i_bench_default.json 2>> $O/r02i_err.log
JAF_WF_MINB_POSES=4 $B > $O/r02i_bench_poses_minb4.json 2>> $O/r02i_err.log
$B --flow smpl > $O/r02i_bench_smpl.json 2>> $O/r02i_err.log
JAF_WF_MINB_SKIP=4 $B --flow smpl > $O/r02i_bench_smpl_minb4.json 2>> $O/r02i_err.log
JAF_WF_MINB_POSES=4 JAF_WF_ROWS_PER_CTA=8 $B > $O/r02i_bench_poses_minb4_rows8.json 2>> $O/r02i_err.log
JAF_WF_MINB_POSES=4 JAF_WF_ROWS_PER_CTA=32 $B > $O/r02i_bench_poses_minb4_rows32.json 2>> $O/r02i_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_raster|k_warp_fuse" --csv --log-file $O/r02i_launches_from_poses.csv \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02i_prof_from_poses.log 2>&1
for f in $O/r02i_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), (d.get("from_poses") or {}).get("kernel"), d.get("clocks"))
except Exception as ex:
    print("unparsed", ex)
PY
done
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02i_launches_from_poses.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>iv:
        try: d[r[ik][:80]].append(float(r[iv].replace(',','')))
        except: pass
for k,v in d.items(): print(f"{len(v):4d} x avg {sum(v)/len(v)/1e3:9.1f} us  {k}")
PY
tail -5 $O/r02i_err.log
