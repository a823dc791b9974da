"""simseg_b200 — B200-native implementation of SimSeg's data-parallel hot path.

Python mirror of the reference's model interface (``simseg/models/pipelines/clip.py``) on top of
hand-written sm_100a CUDA kernels exported through the C ABI in ``include/simseg_b200.h``.
"""
__version__ = "0.1.0"
