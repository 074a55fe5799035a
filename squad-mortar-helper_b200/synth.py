"""Deterministic synthetic captures — stand-in for src/capture.rs in tests and
benchmarks.  The CPU and CUDA rasterisers share csrc/smh_synth.h and are byte-identical."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi

SEED_BASE = 0x534D4800_00000000  # "SMH"

# named scenes of BASELINE.json's configs
# "dense SL markers + lines": 8 squad-leader markers with lines plus a dozen stray blips gives
# ~200-300 line-search seeds per frame — the worst bundled screenshot has 230
DENSE = dict(n_segments=8, n_blips=12, n_labels=2)
# LSD stress (config 5), two flavours: "thousands of segments" saturates the mask and exits after
# 32 accepted lines; "blips first" puts 5000 blips above a few lines so every blip pixel costs a
# full 3600-ray search before the 32-line exit can trigger
STRESS_SEGMENTS = dict(n_segments=2000, n_blips=0, n_labels=2)
STRESS_BLIPS = dict(n_segments=24, n_blips=5000, n_labels=2, blip_band_pct=12)
PLAIN = dict(n_segments=0, n_blips=0, n_labels=0)


@dataclass
class SynthParams:
    W: int
    H: int
    seed: int = SEED_BASE
    n_segments: int = 8
    n_blips: int = 12
    n_labels: int = 2
    map_open: bool = True
    blip_band_pct: int = 100

    def c(self) -> capi.SynthParams:
        return capi.SynthParams(self.seed, self.W, self.H, self.n_segments, self.n_blips, self.n_labels,
                                1 if self.map_open else 0, self.blip_band_pct, 0)

    def as_dict(self) -> dict:
        return {"W": self.W, "H": self.H, "seed": hex(self.seed), "n_segments": self.n_segments,
                "n_blips": self.n_blips, "n_labels": self.n_labels, "map_open": self.map_open,
                "blip_band_pct": self.blip_band_pct}


def frames_host(p: SynthParams, first: int, n: int) -> np.ndarray:
    """CPU rasteriser -> [n, H, W, 4] uint8 BGRA. Needs no GPU."""
    out = np.empty((n, p.H, p.W, 4), np.uint8)
    cp = p.c()
    rc = capi.lib().smh_b200_synth_frames_host(C.byref(cp), first, n, out.ctypes.data)
    if rc != 0:
        raise capi.VisionError(rc, "synth_frames_host: bad parameters")
    return out


def frames_device(vision, p: SynthParams, first: int, n: int, d_out: int, stream: int = 0) -> None:
    """CUDA rasteriser into device memory at d_out (n*H*W*4 bytes)."""
    cp = p.c()
    capi.check(vision.handle, capi.lib().smh_b200_synth_frames_device(vision.handle, C.byref(cp), first, n, d_out, stream))
