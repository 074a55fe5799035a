import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import smh_b200
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision
from helpers import oracle_run
from test_gpu_batch import _results
from oracle import oracle_py as O
O.set_threads(8)
v = B200Vision.init(0)
W,H,n = 1280,720,6
frames = np.concatenate([synth.frames_host(synth.SynthParams(W, H, map_open=(i % 2 == 0)), i, 1) for i in range(n)])
dev = torch.from_numpy(frames).cuda()
for exh in (False, True):
    b = v.process_batch(dev.data_ptr(), n, W, H, max_gap=15, grayscale=False, lsd_stats=True, lsd_exhaustive=exh, stream=torch.cuda.current_stream().cuda_stream)
    res = _results(v, b, n)
    for i in range(0,n,2):
        want = oracle_run(frames[i], False, 0, 15)
        nl = int(res["n_lines"][i]); gl = res["lines"][i,:nl]; wl = want["lines"]
        same = gl.shape==wl.shape and np.array_equal(gl.view(np.uint32), wl.view(np.uint32))
        print("exh",exh,"frame",i,"n",nl,len(wl),"same",same,"stats",res["stats"][i], want["lsd_stats"])
        if not same:
            m=min(nl,len(wl))
            for k in range(m):
                if not np.array_equal(gl[k].view(np.uint32), wl[k].view(np.uint32)): print("  first diff line",k,gl[k],wl[k]); break
            mask = want["lsd_image"]
            # check find_longest_line API at the differing line's p0
            k0=k
            from oracle import oracle_py as OO
            print("  oracle fll at gpu p0:", OO.find_longest_line_in(mask, float(gl[k0][0]), float(gl[k0][1]), 15.0))
            print("  oracle fll at ora p0:", OO.find_longest_line_in(mask, float(wl[k0][0]), float(wl[k0][1]), 15.0))
