import sys, json
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import smh_b200
from oracle import refgpu_py as R
from conftest import sample_path, load_png_bgra
from helpers import oracle_run
f = load_png_bgra(sample_path("point_intersect.png"))
want = oracle_run(f, True, 0, 15)
with R.RefGpu() as r:
    got = r.process_frame(f, True, 0, 15)
qx,qy,qw,qh = want["roi"]["brq"]
rgb = f[qy:qy+qh, qx:qx+qw, 2::-1]
d = got["ocr_out"] != want["ocr_out"]
ys,xs = np.nonzero(d)
out=[]
for y,x in zip(ys,xs):
    out.append([int(y),int(x),rgb[y,x].tolist(),int(got["ocr_out"][y,x]),int(want["ocr_out"][y,x])])
json.dump({"qw":qw,"qh":qh,"diffs":out}, open("gpurun_out/dbg_refocr.json","w"))
print(len(out)); print(out[:50])
