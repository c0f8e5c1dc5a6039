#!/usr/bin/env python
"""Launch k_convlstm_grouped twice per pyramid level of Accumulate_LSTM_no_loss (24 parts, B=1) — the workload for
`ncu --set full -k regex:k_convlstm_grouped` (see profiles/r01_convlstm_grouped_ncu.txt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from jafpro_b200 import ops  # noqa: E402

for Ch, S in [(12, 200), (24, 100), (24, 50), (48, 25), (96, 13)]:
    x, h, c = (torch.randn(24, 1, Ch, S, S, device="cuda") for _ in range(3))
    w = torch.randn(24, 4 * Ch, 2 * Ch, 3, 3, device="cuda") * 0.05
    wp = ops.convlstm_gpack_weight(w, Ch, Ch)
    for _ in range(2):
        ops.convlstm_step_grouped(x, h, c, wp, None, Ch, Ch)
    torch.cuda.synchronize()
