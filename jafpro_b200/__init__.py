"""jafpro_b200 — B200-native (sm_100a) implementation of JAFPro's appearance warp-and-fuse
hot path, kept as a drop-in for the reference's Python call sites:

    reference                                          here
    -------------------------------------------------  ------------------------------------------
    neural_renderer.rasterize_face_index_map_and_...   jafpro_b200.neural_renderer (same names)
    src.nmr.SMPLRenderer.render_fim_wim / cal_bc_...   jafpro_b200.nmr.SMPLRenderer
    src.cal_flow.float_estimate                        jafpro_b200.cal_flow.float_estimate
    src.flow_net.Propagation3DFlowNet (:87-99)         jafpro_b200.flow_net.Propagation3DFlowNet
    src.networks.Downsampler_mask K-reduction          jafpro_b200.fusion.softmax_fuse / warp_fuse
    src.convLSTM.ConvLSTMCell / ConvLSTM               jafpro_b200.convLSTM

All compute goes through the C ABI of include/jafpro_b200.h (libjafpro_b200.so, hand-written
CUDA).  There is no CPU fallback: importing the ops without the built library raises.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
__all__ = ["ops", "neural_renderer", "nmr", "cal_flow", "flow_net", "fusion", "convLSTM", "synth", "dist"]
