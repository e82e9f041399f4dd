"""flowmse_b200: B200-native reverse-ODE sampling hot path of FlowSE (reference: seongq/flowmse).

Host side mirrors the reference's plugin surface (registries, get_white_box_solver, NCSNpp, VFModel); all arithmetic
runs in libflowse.so (hand-written sm_100a CUDA behind the C ABI of include/flowse.h).  No CPU fallback.
"""
__all__ = ["ncsnpp_spec", "checkpoint"]
