"""squad-mortar-helper_b200 — B200 (sm_100a) backend for squad-mortar-helper's per-frame
map-analysis vision path (the `vision-common::Vision` trait).  See DESIGN.md / INTEGRATION.md.

    csrc/      CUDA kernels + the extern "C" ABI of include/smh_b200.h
    capi.py    ctypes declarations of that ABI
    vision.py  host-side mirror of `trait Vision` (B200Vision)
    synth.py   deterministic synthetic captures
    shard.py   frame-wise sharding across GPUs (no collectives on the data path)
    build.py   in-tree nvcc build
"""
from . import build, capi, shard, synth, vision  # noqa: F401
from .vision import B200Vision, DebugView, VisionError  # noqa: F401
