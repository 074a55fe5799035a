#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 on one GPU (the bench line is config 3): 4K frames, and the line-search stress scenes
of SURVEY.md 8(d) — frames/s with one batch in flight plus the line search's work counters.
usage: python tools/config_bench.py > gpurun_out/configs.jsonl"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

v = B200Vision.init(0)
st = torch.cuda.current_stream().cuda_stream


def arr(ptr, nbytes, dt):
    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda").cpu().numpy().view(dt)


def run(name, W, H, n, kw, steps=5):
    p = synth.SynthParams(W, H, **kw)
    frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(v, p, 0, n, frames.data_ptr())
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
        v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, stream=st)
    torch.cuda.synchronize()
    px = lsd = 0.0
    for _ in range(steps):
        ev[0].record()
        v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, skip_lsd=True, stream=st)
        ev[1].record()
        v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, skip_pixels=True, stream=st)
        ev[2].record()
        torch.cuda.synchronize()
        px += ev[0].elapsed_time(ev[1]) / steps
        lsd += ev[1].elapsed_time(ev[2]) / steps
    b = v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, lsd_stats=True, skip_pixels=True, stream=st)
    torch.cuda.synchronize()
    stats = arr(b.lsd_stats, n * 24, np.uint64).reshape(n, 3).astype(np.float64)
    nl = arr(b.n_lines, n * 4, np.uint32)
    print(json.dumps({"config": name, "resolution": [W, H], "frames": n, "synth": p.as_dict(), "frames_per_s": n / ((px + lsd) * 1e-3),
                      "k_pixels_ms": px, "k_lsd_ms": lsd, "lines_per_frame": float(nl.mean()), "seeds_per_frame": float(stats[:, 0].mean()),
                      "seeds_pruned_per_frame": float(stats[:, 2].mean()), "ray_steps_per_frame": float(stats[:, 1].mean()),
                      "ray_steps_per_s": float(stats[:, 1].sum()) / (lsd * 1e-3)}), flush=True)
    del frames
    torch.cuda.empty_cache()


run("config 3: 1440p dense, one batch in flight", 2560, 1440, 1024, synth.DENSE)
run("config 4 frame size: 4K dense (per-GPU shard of the 4096-frame batch at 8 GPUs = 512 frames)", 3840, 2160, 512, synth.DENSE)
run("config 5: 1440p LSD stress, 2000 segments", 2560, 1440, 256, synth.STRESS_SEGMENTS)
run("config 5: 1440p LSD stress, noise blips ahead of the lines", 2560, 1440, 256, dict(n_segments=24, n_blips=400, n_labels=2, blip_band_pct=12))
v.close()
