"""pointdreamer_b200 — B200-native (sm_100a) project -> DDNM-inpaint -> unproject path.

Host side mirrors the reference's operator surface (YuQiao0303/PointDreamer:
pointdreamer/ours_utils.py, pointdreamer/unproject.py, models/DDNM/ddnm_inpainting.py,
demo.py:colorize_one_mesh); all arithmetic runs in hand-written CUDA kernels behind the
C ABI in include/pdr.h.
"""
__version__ = "0.1.0"
