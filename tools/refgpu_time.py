#!/usr/bin/env python3
"""The reference's GPU backend (its kernels in its launch order, oracle/_ref/libsmh_refgpu.so) on the bench workload — the
"GPU kernel to beat" of SURVEY.md 8(d): end-to-end frames/s (warm-up, >= 256 frames, median of 5 passes) and the
device time of every reference kernel (CUDA events around each launch).  usage: tools/refgpu_time.py [n_frames]
Under ncu (tools/gpu_visit_e.sh) the same script with `--one` runs a single frame so that every reference launch is captured."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import oracle_py as O  # noqa: E402
from oracle import refgpu_py as R  # noqa: E402

one = "--one" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = 1 if one else (int(args[0]) if args else 256)
distinct = min(n, 64)
frames = O.synth_frames(2560, 1440, 0, distinct)
with R.RefGpu() as r:
    if one:
        r.process_frame(frames[0], True, 0, 15)
        print("one frame through the reference kernels")
    else:
        import numpy as np
        reps = np.concatenate([frames] * (n // distinct)) if n > distinct else frames
        t = r.time_frames(reps, True, 15, warm=8, passes=5)
        k = r.kernel_times(frames[:32], True, 15)
        print(json.dumps({"workload": f"{len(reps)} x 2560x1440 dense synthetic captures ({distinct} distinct), host frames in, host results out",
                          **t, "kernels_us_per_frame": k}))
