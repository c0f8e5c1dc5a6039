#!/bin/bash
# Round-2 GPU job A: parity suite, hot-kernel A/B (merged RGB, occupancy), hard flows, secondary workloads, ncu.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
B="timeout 600 python bench.py --steps 20 --warmup 5"
$B > $O/r02a_bench_default.json 2> $O/r02a_bench_default.err
JAF_WF_RGB_MERGE=0 $B --no-cpu > $O/r02a_bench_merge0.json 2>> $O/r02a_err.log
JAF_WF_RGB_MERGE=1 JAF_WF_WIDE_MINB=3 $B --no-cpu > $O/r02a_bench_merge1_minb3.json 2>> $O/r02a_err.log
JAF_WF_RGB_MERGE=0 JAF_WF_WIDE_MINB=3 $B --no-cpu > $O/r02a_bench_merge0_minb3.json 2>> $O/r02a_err.log
$B --no-cpu --flow hard > $O/r02a_bench_hard.json 2>> $O/r02a_err.log
JAF_WF_RGB_MERGE=0 $B --no-cpu --flow hard > $O/r02a_bench_hard_merge0.json 2>> $O/r02a_err.log
$B --no-cpu --flow perm > $O/r02a_bench_perm.json 2>> $O/r02a_err.log
$B --no-cpu --flow smpl > $O/r02a_bench_smpl.json 2>> $O/r02a_err.log
$B --no-cpu --workload scaled_512_k8_c64 > $O/r02a_bench_512k8.json 2>> $O/r02a_err.log
$B --no-cpu --workload scaled_512_k8_c64 --flow hard > $O/r02a_bench_512k8_hard.json 2>> $O/r02a_err.log
$B --no-cpu --workload rgb_only_256_k4 > $O/r02a_bench_rgbonly.json 2>> $O/r02a_err.log
for w in c1_latency c3_flow c4_convlstm; do
  timeout 600 python bench.py --steps 10 --warmup 5 --workload $w > $O/r02a_bench_$w.json 2>> $O/r02a_err.log
done
# ncu: launch list of the default bench + full captures of the hot kernel and of the raster kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02a_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_warp_fuse_nhwc -s 4 -c 1 -o $O/r02a_wf_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu --e2e-frames 30 > $O/r02a_ncu_wf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 9 -c 3 -o $O/r02a_raster_full -f \
  python bench.py --steps 5 --warmup 3 --no-cpu --workload c3_flow > $O/r02a_ncu_raster.log 2>&1
ls -la $O | tail -30
tail -3 $O/r02a_pytest.log
for f in $O/r02a_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print(d.get("value"), d.get("unit"), "frac", r.get("frac"), r.get("kernel"), "e2e", (d.get("e2e") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print("unparsed", e)
PY
done
