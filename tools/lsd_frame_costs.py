#!/usr/bin/env python3
"""Per-frame line-search work (stats counters) against the white-pixel count of the mask (bench workload)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import smh_b200  # noqa: F401
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

W, H, n = 2560, 1440, 512
v = B200Vision.init(0)
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(v, synth.SynthParams(W, H), 0, n, frames.data_ptr())
st = torch.cuda.current_stream().cuda_stream
b = v.process_batch(frames.data_ptr(), n, W, H, max_gap=15, lsd_stats=True, stream=st)
torch.cuda.synchronize()


def arr(ptr, nbytes, dt):
    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda").cpu().numpy().view(dt)


stats = arr(b.lsd_stats, n * 24, np.uint64).reshape(n, 3).astype(np.float64)
mask = v.download_plane(b.lsd_image, 0, n)
white = (mask == 255).reshape(n, -1).sum(axis=1).astype(np.float64)
nl = arr(b.n_lines, n * 4, np.uint32).astype(np.float64)
for name, col in (("seeds", stats[:, 0]), ("ray steps", stats[:, 1]), ("pruned", stats[:, 2]), ("seeds - pruned", stats[:, 0] - stats[:, 2]), ("lines", nl)):
    print(f"{name:16s} mean {col.mean():10.1f} min {col.min():9.0f} max {col.max():9.0f} cv {col.std() / col.mean():.2f}  corr(white) {np.corrcoef(col, white)[0, 1]:+.2f}")
print("white pixels     mean %.0f cv %.2f" % (white.mean(), white.std() / white.mean()))
