#!/bin/bash
# Round-2 GPU job E: full parity suite, final defaults, profiles (launch lists + ncu --set full), ConvLSTM pyramid.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02e_pytest.log
tail -8 $O/r02e_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5"
$B > $O/r02e_bench_default.json 2>> $O/r02e_err.log
JAF_WF_MINB_POSES=4 $B --no-cpu > $O/r02e_bench_poses_minb4.json 2>> $O/r02e_err.log
JAF_WF_MINB_POSES=4 JAF_WF_ROWS_PER_CTA=8 $B --no-cpu > $O/r02e_bench_poses_minb4_rows8.json 2>> $O/r02e_err.log
$B --no-cpu --flow hard > $O/r02e_bench_hard.json 2>> $O/r02e_err.log
$B --no-cpu --flow smpl > $O/r02e_bench_smpl.json 2>> $O/r02e_err.log
$B --no-cpu --flow perm > $O/r02e_bench_perm.json 2>> $O/r02e_err.log
$B --no-cpu --workload scaled_512_k8_c64 > $O/r02e_bench_512k8.json 2>> $O/r02e_err.log
$B --no-cpu --workload rgb_only_256_k4 > $O/r02e_bench_rgbonly.json 2>> $O/r02e_err.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02e_bench_reference.json 2>> $O/r02e_err.log
timeout 600 python tools/bench_convlstm_small.py > $O/r02e_convlstm_small.jsonl 2>> $O/r02e_err.log
# profiles
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_warp_fuse_nhwc_wide" -c 30 --csv --log-file $O/r02e_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02e_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_warp_fuse_nhwc_wide" -s 4 -c 1 -o $O/r02e_wf_full -f \
  python tools/prof_step.py --what warp_fuse --reps 6 > $O/r02e_ncu_wf.log 2>&1
python tools/ncu_traffic.py $O/r02e_wf_full.ncu-rep $O/r02e_bench_default.json > $O/r02e_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_warp_fuse_nhwc_wide" -s 4 -c 1 -o $O/r02e_wf_hard_full -f \
  python tools/prof_step.py --what warp_fuse --flow hard --reps 6 > $O/r02e_ncu_wf_hard.log 2>&1
python tools/ncu_traffic.py $O/r02e_wf_hard_full.ncu-rep $O/r02e_bench_hard.json >> $O/r02e_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_raster|k_warp_fuse_nhwc<" -s 6 -c 3 -o $O/r02e_from_poses_full -f \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02e_ncu_from_poses.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_raster" -s 6 -c 3 -o $O/r02e_cal_flow_full -f \
  python tools/prof_step.py --what cal_flow --reps 3 > $O/r02e_ncu_cal_flow.log 2>&1
cp profiles/r02_traffic.json $O/r02e_traffic.json 2>/dev/null
cat $O/r02e_traffic.log
for f in $O/r02e_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print("unparsed", ex)
PY
done
cat $O/r02e_convlstm_small.jsonl | tail -3
tail -5 $O/r02e_err.log
