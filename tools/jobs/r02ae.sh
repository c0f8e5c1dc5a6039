#!/bin/bash
# Round-2 GPU job AE: ncu --set full of the stand-alone RGB kernel (taps shared between x-adjacent lanes) and, on the same
# box, of the shipped fused kernel (final build), dense flows.  Only the text summaries travel back.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --clock-control none -k regex:k_warp_fuse_rgb -s 2 -c 1 -o /tmp/ncu/rgb -f \
  python tools/prof_step.py --what warp_fuse --workload rgb_only_256_k4 --reps 4 > $O/r02ae_ncu_rgb.log 2>&1
python tools/ncu_summary.py /tmp/ncu/rgb.ncu-rep > $O/r02ae_rgb_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_warp_fuse_nhwc_wide -s 2 -c 1 -o /tmp/ncu/wf -f \
  python tools/prof_step.py --what warp_fuse --reps 4 > $O/r02ae_ncu_wf.log 2>&1
python tools/ncu_summary.py /tmp/ncu/wf.ncu-rep > $O/r02ae_wf_summary.txt 2>&1
head -40 $O/r02ae_rgb_summary.txt
