#!/bin/bash
# Round-2 GPU job AB: ncu --set full of every level of the grouped ConvLSTM pyramid with the final kernels
# (k_convlstm_grouped_p: 12@200^2, 24@100^2, 24@50^2; k_convlstm_grouped: 48@25^2, 96@13^2).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_convlstm_grouped -c 10 -o $O/r02ab_cg_full -f \
  python tools/probes/convlstm_grouped_once.py > $O/r02ab_ncu.log 2>&1
tail -3 $O/r02ab_ncu.log
