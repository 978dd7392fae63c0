"""u2seg_b200 — B200-native (sm_100a) implementation of U2Seg's two data-parallel hot paths:
the Panoptic-FPN detector step and the k-means pseudo-label clustering. Host code mirrors the
reference's operator interface; all device work goes through libu2b200.so (include/u2b200.h)."""
__version__ = "0.1.0"
