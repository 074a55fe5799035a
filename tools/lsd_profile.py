#!/usr/bin/env python3
"""Per-phase cycle shares of k_lsd (library built with -DLSD_PROFILE; see tools/build_variants.sh).
usage: SMH_B200_LIB=build/variants/lib_prof.so python tools/lsd_profile.py [frames]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import smh_b200  # noqa: E402,F401
from squad_mortar_helper_b200 import capi, synth  # noqa: E402
from squad_mortar_helper_b200.vision import B200Vision  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
W, H = 2560, 1440
v = B200Vision.init(0)
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
synth.frames_device(v, synth.SynthParams(W, H), 0, n, frames.data_ptr())
st = torch.cuda.current_stream().cuda_stream
L = capi.lib()
out = (C.c_uint64 * 16)()
for it in range(3):
    v.process_batch(frames.data_ptr(), n, W, H, stream=st)
    torch.cuda.synchronize()
    L.smh_b200_debug_lsd_profile(v.handle, out, 1)
names = {0: "tile-row load+strike", 1: "seed pick", 2: "dismissed (dead cluster)", 3: "ring scan", 4: "window load",
         5: "ray marking", 6: "ray list", 7: "march", 8: "accept+restrike+row end", 9: "get_centre", 15: "barrier after ring scan"}
tot = sum(out[k] for k in names) or 1
for k, nm in names.items():
    print(f"{nm:28s} {out[k] / tot * 100:5.1f}%  {out[k] / n / 1.965e3:9.1f} us/frame")
print(f"total {tot / n / 1.965e3:.1f} us/frame (cycles of thread 0 at 1.965 GHz)")
print(f"per frame: {out[10] / n:.1f} seeds reach the window load, {out[11] / n:.0f} rays listed")
