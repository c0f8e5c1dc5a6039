"""Drop-in for the warping part of ``SpatioTempoCRN.forward`` (src/crn_model.py:457-566): per pyramid level the
reference nearest-downsamples the flow, adds / subtracts it to the level's base grid and calls ``F.grid_sample`` twice
(``padding_mode='border'``).  ``warp_level`` does one level in one launch, ``warp_pyramid`` all of them."""
from __future__ import annotations

from . import ops


def warp_level(prev_pool, pool, grid, flow, align_corners: bool = False):
    """-> (warped_prev_pool, warped_pool) of one level (e.g. src/crn_model.py:457-466 for level 6):
    warped_prev_pool = grid_sample(prev_pool, grid + flow_s), warped_pool = grid_sample(pool, grid - flow_s)."""
    return ops.flow_warp_pair(None if prev_pool is None else prev_pool.contiguous(),
                              None if pool is None else pool.contiguous(), grid.contiguous(), flow.contiguous(),
                              align_corners)


def warp_pyramid(prev_pools, pools, grid_list, flow, align_corners: bool = False):
    """All levels: prev_pools[i], pools[i] [B,C_i,h_i,w_i], grid_list[i] [B,2,h_i,w_i], flow [B,2,H,W]
    -> [(warped_prev_pool_i, warped_pool_i)]."""
    return [warp_level(pp, p, g, flow, align_corners) for pp, p, g in zip(prev_pools, pools, grid_list)]
