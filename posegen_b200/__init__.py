"""posegen_b200 — B200-native A-NeRF volumetric renderer behind PoseGen's RayCaster API.

Host side (Python, mirrors the reference's interface for the render path):
  posegen_b200.raycaster   RayCaster / create_raycaster drop-in (core/raycasters.py)
  posegen_b200.render      render / batchify_rays / render_path callers (core/trainer.py, run_nerf.py)
  posegen_b200.engine      handle on the C-ABI context (include/posegen_b200.h)
  posegen_b200.synthetic   synthetic poses / cameras / weights for tests and bench
  posegen_b200.dist        pose/image sharding across ranks
Device side: posegen_b200/csrc (hand-written sm_100a CUDA) -> lib/libposegen_b200.so.
"""
__version__ = "0.1.0"
