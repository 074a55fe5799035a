#!/usr/bin/env python3
"""Generates tests/golden/marker_templates_22.npy from the reference's bundled TGA icons
(vision-common/src/markers/resources/) with oracle/marker_templates.py (a restatement of the commented-out
`markers::load_markers`; image-crate arithmetic unpinned).  Run in the build container only —
/root/reference does not exist on the GPU box, the .npy travels."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import marker_templates as M  # noqa: E402

RES = Path("/root/reference/vision-common/src/markers/resources")
if __name__ == "__main__":
    size = 22  # src/vision/mod.rs:48
    t = M.load_markers(RES, size)
    np.save(Path(__file__).with_name(f"marker_templates_{size}.npy"), t)
    for k, name in enumerate(M.MARKER_FILES):
        a = t[k, ..., 3]
        print(name, "alpha>0:", int((a > 0).sum()), "of", size * size)
        print("\n".join("".join("#" if v > 128 else "+" if v else "." for v in row) for row in a))
