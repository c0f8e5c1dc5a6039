#!/bin/bash
# Round-2 final 8-GPU job: the driver's torchrun command at N = 8 / 4 / 2 and the plain command at N = 1 on ONE box, final
# build (e2e step = the whole per-GPU step), + BASELINE config 5 (512^2, K=8, 64 videos x 30 frames over 8 GPUs).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02g8_gpus.txt
T="timeout 900 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02g8_bench_n1.json 2> $O/r02g8_bench_n1.err
$T --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02g8_bench_n2.json 2> $O/r02g8_bench_n2.err
$T --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r02g8_bench_n4.json 2> $O/r02g8_bench_n4.err
$T --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02g8_bench_n8.json 2> $O/r02g8_bench_n8.err
$T --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 5 --workload scaled_512_k8_c64 > $O/r02g8_bench_512k8_n8.json 2> $O/r02g8_bench_512k8_n8.err
$T --nproc-per-node 8 --master-port 29545 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/r02g8_bench_reference_n8.json 2> $O/r02g8_bench_reference_n8.err
for f in $O/r02g8_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(d.get("impl"), d.get("n_gpus"), d.get("value"), d.get("unit"), "frac(rank0)", r.get("frac"), "e2e", e.get("value"), e.get("frames_per_step"), "app", (e.get("application") or {}).get("value"), "from_poses", (d.get("from_poses") or {}).get("value"), d.get("clocks"))
    print("   ", e.get("pinned_memcpy_probe"))
except Exception as ex:
    print("unparsed", ex)
PY
done
tail -3 $O/r02g8_bench_n8.err
