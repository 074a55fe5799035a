#!/usr/bin/env python3
"""Timeline of smh_b200_process_host_batch chunks (run with SMH_B200_PIPE_TRACE=1).  usage: e2e_trace.py [frames] [chunk]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import smh_b200  # noqa
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision, roi as roi_of

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
W, H = 2560, 1440
v = B200Vision.init(0)
dev = torch.empty((ne, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(v, synth.SynthParams(W, H), 0, ne, dev.data_ptr())
host = torch.empty((ne, H, W, 4), dtype=torch.uint8).pin_memory()
host.copy_(dev)
roi = roi_of(W, H)
_, _, mw, mh = roi["map"]; _, _, qw, qh = roi["brq"]
outs = {"ui_map": torch.empty((ne, mh, mw, 4), dtype=torch.uint8).pin_memory(),
        "ocr_out": torch.empty((ne, qh, qw), dtype=torch.uint8).pin_memory(),
        "scales": torch.empty((ne, qh, qw), dtype=torch.uint8).pin_memory(),
        "lines": torch.empty((ne, 32, 4), dtype=torch.float32).pin_memory(),
        "n_lines": torch.empty((ne,), dtype=torch.int32).pin_memory(),
        "red_pixels": torch.empty((ne,), dtype=torch.int32).pin_memory(),
        "map_open": torch.empty((ne,), dtype=torch.uint8).pin_memory()}
ptrs = {k: t.data_ptr() for k, t in outs.items()}
for it in range(3):
    t0 = time.perf_counter()
    v.process_host_batch(None, frames_ptr=host.data_ptr(), shape=(ne, H, W), max_gap=15, grayscale=True, chunk=chunk,
                         want=("ui_map", "ocr_out", "scales"), out=dict(ptrs))
    dt = time.perf_counter() - t0
    print(f"iteration {it}: {ne / dt:.0f} frames/s ({dt * 1e3:.1f} ms)", file=sys.stderr)
