#!/usr/bin/env python3
"""Experiment: consecutive steps on two handles / two streams (step k+1's pixel pass under step k's line-search tail)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

W, H = 2560, 1440
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
vs = [B200Vision.init(0) for _ in range(3)]
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(vs[0], synth.SynthParams(W, H), 0, n, frames.data_ptr())
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(3)]


def run(k_handles, steps=12):
    for i in range(k_handles * 2):
        vs[i % k_handles].process_batch(frames.data_ptr(), n, W, H, max_gap=15, stream=streams[i % k_handles].cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in streams[:k_handles]:
        s.wait_event(a)
    for i in range(steps):
        vs[i % k_handles].process_batch(frames.data_ptr(), n, W, H, max_gap=15, stream=streams[i % k_handles].cuda_stream)
    for s in streams[:k_handles]:
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for k in (1, 2, 3):
    ms = run(k)
    print(f"{k} handle(s)/stream(s): {ms:.3f} ms per step of {n} frames -> {n / ms:.1f} kframes/s", flush=True)
