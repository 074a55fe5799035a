import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import smh_b200
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision
v = B200Vision.init(0)
for (w, h) in ((1920, 1080), (2560, 1440), (3840, 2160)):
    for fr in (0, 1):
        frame = synth.frames_host(synth.SynthParams(w, h), fr, 1)[0]
        dev = torch.from_numpy(frame[None]).cuda()
        b = v.process_batch(dev.data_ptr(), 1, w, h, max_gap=15, lsd_stats=True, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        class _H: pass
        hd = _H(); hd.__cuda_array_interface__ = {"shape": (24,), "typestr": "|u1", "data": (int(b.lsd_stats), False), "version": 2}
        st = torch.as_tensor(hd, device="cuda").cpu().numpy().view(np.uint64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for it in range(6):
            e0.record()
            v.process_batch(dev.data_ptr(), 1, w, h, max_gap=15, skip_pixels=True, stream=torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(w, h, "frame", fr, "seeds", int(st[0]), "steps", int(st[1]), "pruned", int(st[2]), "lsd ms", round(min(ts), 3))
