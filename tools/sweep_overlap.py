#!/usr/bin/env python3
"""Sweep the chunk size / stream count of the chunked smh_b200_process_batch (bench workload, one process).
usage: python tools/sweep_overlap.py [frames] > gpurun_out/sweep_overlap.txt"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

W, H = 2560, 1440
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
v = B200Vision.init(0)
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(v, synth.SynthParams(W, H), 0, n, frames.data_ptr())
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream


def timed(steps=6, **kw):
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


print(f"serial: {timed(serial=True):.3f} ms / {n} frames", flush=True)
print(f"pixels only: {timed(skip_lsd=True):.3f} ms", flush=True)
for lanes in (2, 3, 4, 6):
    for chunk in (32, 64, 96, 148, 222, 296, 444):
        os.environ["SMH_B200_BATCH_CHUNK"] = str(chunk)
        os.environ["SMH_B200_BATCH_LANES"] = str(lanes)
        ms = timed()
        print(f"lanes {lanes} chunk {chunk:4d}: {ms:.3f} ms  {n / ms:.1f} kframes/s", flush=True)
v.close()
