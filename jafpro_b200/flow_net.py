"""Drop-in for the arithmetic of ``Propagation3DFlowNet.forward`` (src/flow_net.py:87-99):
visibility mask (:91) and confidence blend (:98).  The confidence U-Net itself
(``CompositeWeightUnet``, :6-58) is a generic cuDNN conv stack outside this path and is injected."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class Propagation3DFlowNet(nn.Module):
    def __init__(self, composite_unet: nn.Module, use_tgt_dp=False):
        super().__init__()
        self.composite_unet = composite_unet
        self.use_tgt_dp = use_tgt_dp

    def forward(self, x):
        fake_tgt, tsf_image, tgt_IUV, use_IUV = x['fake_tgt'], x['tsf_image'], x['tgt_IUV'], x['use_IUV']
        use_mask, tgt_smpl_mask = x['use_mask'], x['tgt_smpl_mask']
        if use_mask:
            tsf_image, _ = ops.mask_blend(tsf_image.contiguous(), tgt_smpl_mask.contiguous())  # :91
        if not use_IUV:
            cated_input = torch.cat([tsf_image, fake_tgt], dim=1)
        else:
            cated_input = torch.cat([tsf_image, fake_tgt, tgt_IUV], dim=1)
        weight = self.composite_unet(cated_input)
        _, pred = ops.mask_blend(tsf_image.contiguous(), None, fake_tgt.contiguous(), weight.contiguous())  # :98
        return {'pred_target': pred, 'weight': weight}
