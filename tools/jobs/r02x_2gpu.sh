#!/bin/bash
# Round-2 GPU job X (2 GPUs of one box): two-device tests + the driver's torchrun command at N=2 with the final build
# (e2e step = the whole per-GPU step).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "second_device or every_device or host_pipeline or two_device or devices" > $O/r02x_pytest_2gpu.log 2>&1
tail -4 $O/r02x_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02x_bench_n2.json 2> $O/r02x_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r02x_bench_reference_n2.json 2> $O/r02x_bench_reference_n2.err
for f in $O/r02x_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print(d.get("impl"), d.get("n_gpus"), d.get("value"), d.get("unit"), "frac(rank0)", r.get("frac"), "e2e", e.get("value"), e.get("frames_per_step"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -3 $O/r02x_bench_n2.err
