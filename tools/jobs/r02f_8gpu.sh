#!/bin/bash
# Round-2 GPU job F (8 GPUs of one box): two-device tests, the driver's torchrun command at N=8 for the headline and for
# BASELINE config 5 (512^2, K=8, 64 videos x 30 frames sharded by video over 8 GPUs).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02f_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "second_device or every_device or host_pipeline" > $O/r02f_pytest_2gpu.log 2>&1
tail -4 $O/r02f_pytest_2gpu.log
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
$T bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02f_bench_n8.json 2> $O/r02f_bench_n8.err
$T bench.py --gpus 8 --steps 20 --warmup 5 --workload scaled_512_k8_c64 > $O/r02f_bench_512k8_n8.json 2> $O/r02f_bench_512k8_n8.err
$T bench.py --gpus 8 --steps 20 --warmup 5 --flow hard > $O/r02f_bench_hard_n8.json 2> $O/r02f_bench_hard_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02f_bench_n2.json 2> $O/r02f_bench_n2.err
for f in $O/r02f_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    e=d.get("e2e") or {}
    print(d.get("n_gpus"), d.get("value"), d.get("unit"), "frac(rank0)", r.get("frac"), "e2e", e.get("value"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
    print(e.get("pinned_memcpy_probe"))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -3 $O/r02f_bench_n8.err
