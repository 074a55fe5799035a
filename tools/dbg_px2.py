import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import smh_b200
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision
from helpers import oracle_run
W, H = 1280, 720
v = B200Vision.init(0)
p = synth.SynthParams(W, H)
host = synth.frames_host(p, 100, 2)
dev = torch.from_numpy(host).cuda()
flags = dict(skip_lsd=True) if len(sys.argv) > 1 and sys.argv[1] == "px" else {}
b = v.process_batch(dev.data_ptr(), 2, W, H, max_gap=15, grayscale=True, stream=torch.cuda.current_stream().cuda_stream, **flags)
torch.cuda.synchronize()
print("ok")
