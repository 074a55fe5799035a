#!/usr/bin/env python3
"""Time the batch with and without the marker icon template match (bench workload).  usage: python tools/icons_time.py [frames]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

W, H = 2560, 1440
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
v = B200Vision.init(0)
v.load_map_markers(np.load(ROOT / "tests" / "golden" / "marker_templates_22.npy"))
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(v, synth.SynthParams(W, H), 0, n, frames.data_ptr())
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream


def timed(steps=5, **kw):
    for _ in range(2):
        v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, stream=st, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, stream=st, **kw)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


base = timed()
print(f"whole path: {base:.3f} ms / {n} frames")
nol = timed(skip_lsd=True)
print(f"pixels only: {nol:.3f} ms")
ic = timed(filter_icons=True, skip_lsd=True)
print(f"pixels + pack + icons: {ic:.3f} ms (icons stage ~{ic - nol:.3f} ms incl. pack)")
full = timed(filter_icons=True)
print(f"whole path with icons: {full:.3f} ms")
v.close()
