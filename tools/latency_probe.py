"""Single-frame latency through the trait calls (bench.py's latency leg alone), with the line search's CTA width."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import smh_b200
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision
import bench
v = B200Vision.init(0)
out = bench.run_latency(v, synth, torch)
print(os.environ.get("SMH_B200_LSD_LATENCY_WARPS", "8"), json.dumps({k: {"wall": round(x["wall_ms"], 3), "dev": round(x["device_ms"], 3), "lsd": x["stages_host_ms_device_ms"]["find_marker_lines"], "crop": x["stages_host_ms_device_ms"]["crop_to_map"]} for k, x in out.items()}))
