#!/bin/bash
# Round-2 GPU job G: parity suite (rounds-of-four K=5..8 kernel, shard loader), 512^2 K=8 A/B, pose-driven ncu capture.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log
tail -8 $O/r02g_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu"
$B --workload scaled_512_k8_c64 > $O/r02g_bench_512k8.json 2>> $O/r02g_err.log
JAF_WF_WIDE8_ROUNDS=0 $B --workload scaled_512_k8_c64 > $O/r02g_bench_512k8_rounds0.json 2>> $O/r02g_err.log
$B --workload scaled_512_k8_c64 --flow hard > $O/r02g_bench_512k8_hard.json 2>> $O/r02g_err.log
JAF_WF_WIDE_ROWS_PER_CTA=16 $B --workload scaled_512_k8_c64 > $O/r02g_bench_512k8_rows16.json 2>> $O/r02g_err.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_raster_scatter_flat|k_warp_fuse_nhwc" -s 4 -c 2 -o $O/r02g_from_poses_full -f \
  python tools/prof_step.py --what from_poses --reps 3 > $O/r02g_ncu_from_poses.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_warp_fuse_nhwc_wide2r" -s 2 -c 1 -o $O/r02g_wf512_full -f \
  python tools/prof_step.py --what warp_fuse --workload scaled_512_k8_c64 --videos-per-gpu 4 --reps 4 > $O/r02g_ncu_wf512.log 2>&1
for f in $O/r02g_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -5 $O/r02g_err.log
