"""Marker icon templates — restatement of `markers::load_markers` (vision-common/src/markers/mod.rs:56-152,
commented out in the reference).  TEST INFRASTRUCTURE ONLY: in the reference this is host-side work done
once per marker size with the `image` crate; the CUDA library receives the finished templates as data
(smh_b200_load_map_markers), exactly what vision-gpu uploads (vision-gpu/src/gpumarkers.rs:24-52).

Third-party arithmetic restated from the published algorithm, sources NOT under /root/reference — the template
GENERATION (TGA decode + Gaussian resize of the `image` crate) is **parity unpinned**: the reference never runs it
(the stage is commented out) and the crate cannot be built here; the MATCH against finished templates is pinned
against the reference's kernels K11/K12 (tests/test_gpu_refcuda.py::test_template_match_kernels_k11_k12):
  image 0.23.14 (Cargo.lock:1487)  TGA decoder (type 10 RLE true-colour, bottom-left origin flipped),
                                   imageops::resize with FilterType::Gaussian: vertical pass into f32
                                   pixels, then horizontal pass, clamp, round-to-nearest;
                                   kernel gaussian(x, 0.5), support 3.0.
Pipeline per icon (markers/mod.rs:96-109): decode -> resize_exact(size, size, Gaussian) -> rgba8 ->
non-marker-colour pixels := (0,0,0,0) -> pixels listed by corners.tga := (0,0,0,0).
"""
from __future__ import annotations

import math
import struct
from pathlib import Path

import numpy as np

# order of the markers! macro (markers/mod.rs:68-74)
MARKER_FILES = ("map_commandmarker_squad_attack.TGA", "map_commandmarker_squad_build.TGA",
                "map_commandmarker_squad_defend.TGA", "map_commandmarker_squad_move.TGA",
                "map_commandmarker_squad_observe.TGA")
CORNERS_FILE = "corners.tga"
F = np.float32


def decode_tga(data: bytes) -> np.ndarray:
    """-> RGBA [H, W, 4] uint8, top row first."""
    (id_len, cmap_type, img_type, _cs, _cl, _cb, _xo, _yo, w, h, bpp, desc) = struct.unpack("<BBBHHBHHHHBB", data[:18])
    assert cmap_type == 0 and img_type in (2, 10) and bpp in (24, 32), (cmap_type, img_type, bpp)
    pos = 18 + id_len
    nb = bpp // 8
    px = np.zeros((w * h, nb), np.uint8)
    if img_type == 2:
        px[:] = np.frombuffer(data, np.uint8, w * h * nb, pos).reshape(-1, nb)
    else:
        i = 0
        while i < w * h:
            hdr = data[pos]; pos += 1
            cnt = (hdr & 0x7F) + 1
            if hdr & 0x80:
                px[i:i + cnt] = np.frombuffer(data, np.uint8, nb, pos)
                pos += nb
            else:
                px[i:i + cnt] = np.frombuffer(data, np.uint8, cnt * nb, pos).reshape(cnt, nb)
                pos += cnt * nb
            i += cnt
    img = px.reshape(h, w, nb)
    out = np.empty((h, w, 4), np.uint8)
    out[..., 0], out[..., 1], out[..., 2] = img[..., 2], img[..., 1], img[..., 0]  # BGR(A) -> RGB(A)
    out[..., 3] = img[..., 3] if nb == 4 else 255
    if not desc & 0x20:  # screen-origin bit clear: first row is the bottom one
        out = out[::-1]
    return np.ascontiguousarray(out)


def _gaussian(x: np.float32) -> np.float32:
    r = F(0.5)
    norm = F(1.0) / (F(math.sqrt(F(2.0) * F(math.pi))) * r)
    return F(norm * F(np.exp(F(-(x * x)) / (F(2.0) * r * r))))


def _weights(n_in: int, n_out: int):
    """Per output index: (left, normalised f32 weights) of image 0.23's sampler."""
    ratio = F(n_in) / F(n_out)
    sratio = F(1.0) if ratio < 1.0 else ratio
    support = F(3.0) * sratio
    out = []
    for o in range(n_out):
        c = F(F(o) + F(0.5)) * ratio
        left = int(min(max(math.floor(F(c - support)), 0), n_in - 1))
        right = int(min(max(math.ceil(F(c + support)), left + 1), n_in))
        c = F(c - F(0.5))
        ws = [_gaussian(F(F(F(i) - c) / sratio)) for i in range(left, right)]
        s = F(0.0)
        for w in ws:
            s = F(s + w)
        out.append((left, [F(w / s) for w in ws]))
    return out


def resize_gaussian(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """imageops::resize(img, nw, nh, Gaussian) for RGBA8: vertical pass (f32 result), horizontal pass, round."""
    h, w = img.shape[:2]
    src = img.astype(np.float32)
    tmp = np.zeros((nh, w, 4), np.float32)
    for oy, (top, ws) in enumerate(_weights(h, nh)):
        t = np.zeros((w, 4), np.float32)
        for i, wt in enumerate(ws):
            t = (t + src[top + i] * wt).astype(np.float32)
        tmp[oy] = t
    out = np.zeros((nh, nw, 4), np.uint8)
    for ox, (left, ws) in enumerate(_weights(w, nw)):
        t = np.zeros((nh, 4), np.float32)
        for i, wt in enumerate(ws):
            t = (t + tmp[:, left + i] * wt).astype(np.float32)
        t = np.clip(t, F(0.0), F(255.0))
        # FloatNearest: f32::round, half away from zero
        out[:, ox] = np.floor(t + F(0.5)).astype(np.uint8)
    return out


def _is_marker_img(rgb: np.ndarray) -> np.ndarray:
    from . import oracle_py as O
    L = O.lib()
    flat = rgb.reshape(-1, 3)
    return np.array([bool(L.smh_oracle_is_marker(int(r), int(g), int(b))) for r, g, b in flat], bool).reshape(rgb.shape[:2])


def process_marker(data: bytes, size: int, corners=None) -> np.ndarray:
    """markers/mod.rs:96-109 -> RGBA [size, size, 4]."""
    img = resize_gaussian(decode_tga(data), size, size)
    img[~_is_marker_img(img[..., :3])] = 0
    if corners is not None:
        img[corners] = 0
    return img


def load_markers(resources: Path, size: int) -> np.ndarray:
    """markers/mod.rs:111-119 -> [5, size, size, 4] uint8 (UnfilteredMarkers layout)."""
    resources = Path(resources)
    corners = process_marker((resources / CORNERS_FILE).read_bytes(), size)[..., 3] != 0  # FilteredMarkers: alpha != 0
    return np.stack([process_marker((resources / f).read_bytes(), size, corners) for f in MARKER_FILES])
