"""Run the reference's K8 (mask_marker_lines) over all 2^24 colours on the GPU box; store the bit-packed table."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from oracle import refgpu_py as R
idx = np.arange(1 << 24, dtype=np.uint32)
rgb = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
with R.RefGpu() as r:
    raw, _ = r.mask_and_dilate(rgb)
bits = np.packbits((raw.reshape(-1) == 255).astype(np.uint8))
np.save(Path(__file__).resolve().parents[1] / "gpurun_out" / "k8_marker_table_bits.npy", bits)
print("marker colours per K8:", int((raw == 255).sum()))
