#!/usr/bin/env python3
"""Time the reference's GPU backend (its kernels in its launch order, oracle/_ref/libsmh_refgpu.so) on the
bench workload: the "GPU kernel to beat" of SURVEY.md 8(d).  usage: tools/refgpu_time.py [n_frames]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import smh_b200  # noqa: F401,E402
from oracle import refgpu_py as R  # noqa: E402
from squad_mortar_helper_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = synth.frames_host(synth.SynthParams(2560, 1440), 0, n)
with R.RefGpu() as r:
    print(json.dumps({"workload": f"{n} x 2560x1440 dense synthetic captures, host frames in, host results out", **r.time_frames(frames, True, 15, 1)}))
