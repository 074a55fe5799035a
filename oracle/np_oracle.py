"""Independent numpy/float32 restatement of the vision-cpu semantics.  TEST INFRASTRUCTURE ONLY.

Second, differently-structured statement of SURVEY.md Appendix A used to cross-check
oracle/smh_oracle.c (whole-array float32 ops instead of per-pixel scalar code, shift-based
dilation instead of a distance transform, ray marching vectorised over the 3600 angles).
It shares no code with the C oracle.  PARITY STATUS: this restatement itself is pinned only by its agreement with
the C oracle (which is pinned against the reference's own compiled CUDA kernels, see smh_oracle.h).

Reference lines restated (relative to /root/reference):
  ROI          vision-common/src/screen.rs:55-66, consts/mod.rs:7-19, vision-cpu/src/lib.rs:137-145
  luma         image 0.23.14 rgb_to_luma (call sites vision-cpu/src/lib.rs:154,224,242)
  hsv/marker   util/src/image.rs:159-187, vision-common/src/markers/mod.rs:17-54
  ocr          vision-cpu/src/lib.rs:39-53,173-231
  scales       vision-cpu/src/lib.rs:233-251
  mask+dilate  vision-cpu/src/lib.rs:357-375 (imageproc dilate_mut L1 k=1 = 3x3 cross)
  rays         vision-cpu/src/lib.rs:387-449
  find_lines   vision-common/src/lsd.rs:5-105
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

BTN_COLOR = np.array([217, 67, 49], np.int16)
TEAMS = ((105, 100, 100), (285, 46, 85), (158, 60, 91))


def _round_half_away(v: float) -> int:
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def roi(W: int, H: int):
    rh = lambda f: _round_half_away(f * float(H))
    bw, bh = rh(0.236132813), rh(0.038205295)
    bx, by = W - rh(0.0078125) - bw, H - rh(0.0078125) - bh
    w0, h, x0 = rh(0.864930556), rh(0.761078559), rh(0.018522135)
    y = H - rh(0.07421875) - h
    w = W - w0
    x = W - x0 - w
    if min(bx, by, x, y) < 0 or w <= 0 or x + w >= W or y + h >= H:
        return None
    return {"button": (bx, by, bw, bh), "map": (x, y, w, h), "brq": (x + w // 2, y + h // 2, w // 2, h // 2)}


def luma8(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., i].astype(F) for i in range(3))
    l = (F(0.2126) * r + F(0.7152) * g) + F(0.0722) * b
    return np.trunc(l).astype(np.uint8)


def hsv(rgb: np.ndarray):
    with np.errstate(invalid="ignore", divide="ignore"):
        r, g, b = (rgb[..., i].astype(F) / F(255.0) for i in range(3))
        mx = np.maximum(r, np.maximum(g, b))
        mn = np.minimum(r, np.minimum(g, b))
        d = mx - mn
        h_r = F(60.0) * np.fmod((g - b) / d, F(6.0))
        h_g = F(60.0) * (((b - r) / d) + F(2.0))
        h_b = F(60.0) * (((r - g) / d) + F(4.0))
        h = np.where(mx == mn, F(0.0), np.where(mx == r, h_r, np.where(mx == g, h_g, h_b))).astype(F)
        h = np.fmod(h, F(360.0))
        h = np.where(h < 0, h + F(360.0), h).astype(F)
        s = (F(100.0) * d) / mx
        v = F(100.0) * mx
    s = np.where(np.isnan(s), F(0), s)
    return (np.trunc(h).astype(np.int32), np.clip(np.trunc(s), 0, 255).astype(np.int32),
            np.clip(np.trunc(v), 0, 255).astype(np.int32))


def is_marker(rgb: np.ndarray) -> np.ndarray:
    h, s, v = hsv(rgb)
    ok = np.zeros(h.shape, bool)
    for mh, ms, mv in TEAMS:
        sat = (np.abs(ms - s) <= 15) | (np.abs(s - (ms - 50)) <= 15)
        ok |= (np.abs(mh - h) <= 15) & sat & (np.abs(mv - v) <= 15)
    return ok & (s >= 35)


def button_red_pixels(frame_bgra: np.ndarray, r) -> int:
    bx, by, bw, bh = r["button"]
    rgb = frame_bgra[by:by + bh, bx:bx + bw, 2::-1].astype(np.int16)
    return int((np.abs(BTN_COLOR - rgb) <= 25).all(axis=-1).sum())


def crop_to_map(frame_bgra: np.ndarray, grayscale: bool):
    H, W = frame_bgra.shape[:2]
    r = roi(W, H)
    red = button_red_pixels(frame_bgra, r)
    bw, bh = r["button"][2:]
    if F(red) / F(bw * bh) < F(0.65):
        return None
    x, y, w, h = r["map"]
    qx, qy, qw, qh = r["brq"]
    cropped_map = frame_bgra[y:y + h, x:x + w, 2::-1].copy()
    cropped_brq = frame_bgra[qy:qy + qh, qx:qx + qw, 2::-1].copy()
    ui = np.empty((h, w, 4), np.uint8)
    if grayscale:
        ui[..., :3] = luma8(cropped_map)[..., None]
    else:
        ui[..., :3] = cropped_map
    ui[..., 3] = 255
    return {"ui_map": ui, "xywh": [x, y, w, h], "cropped_map": cropped_map, "cropped_brq": cropped_brq,
            "red_pixels": red}


def ocr_preprocess(brq: np.ndarray) -> np.ndarray:
    h, w = brq.shape[:2]
    c = brq.astype(np.int32)
    mono = np.zeros((h, w), np.int32)
    for a in range(3):
        for b in range(3):
            mono += np.abs(c[..., a] - c[..., b])
    bright = c.min(axis=-1)
    core = (mono <= 3) & (bright >= 200)
    edge = (mono <= 48) & (bright >= 130)
    # neighbour search window: xx in [x-3 (sat), min(x+3, w-3)], yy likewise
    near = np.zeros((h, w), bool)
    ys, xs = np.nonzero(core)
    for yy, xx in zip(ys, xs):
        if xx > w - 3 or yy > h - 3:
            continue  # never inside any window (upper clip)
        # pixel (x,y) sees (xx,yy) iff x-3<=xx<=x+3 (and xx<=w-3) => x in [xx-3, xx+3]
        near[max(yy - 3, 0):yy + 4, max(xx - 3, 0):xx + 4] = True
    keep = core | (edge & near)
    return np.where(keep, 255 - luma8(brq).astype(np.int32), 255).astype(np.uint8)


def find_scales_preprocess(brq: np.ndarray, start_y: int, prev: np.ndarray | None = None) -> np.ndarray:
    h, w = brq.shape[:2]
    out = np.zeros((h, w), np.uint8) if prev is None else prev.copy()
    if start_y < h:
        out[start_y:] = np.where(luma8(brq[start_y:]) != 0, 255, 0)
    return out


def mask_marker_lines(cropped_map: np.ndarray):
    m = is_marker(cropped_map)
    d = m.copy()
    d[1:, :] |= m[:-1, :]
    d[:-1, :] |= m[1:, :]
    d[:, 1:] |= m[:, :-1]
    d[:, :-1] |= m[:, 1:]
    return (m * 255).astype(np.uint8), (d * 255).astype(np.uint8)


_RAYS = None


def ray_table() -> np.ndarray:
    """cos/sin via the C library's cosf/sinf (what Rust's f32::cos/sin call on linux-gnu)."""
    global _RAYS
    if _RAYS is None:
        import ctypes
        libm = ctypes.CDLL("libm.so.6")
        libm.cosf.restype = libm.sinf.restype = ctypes.c_float
        libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
        rpd = F(np.float32(3.14159265358979323846) / F(180.0))
        t = np.empty((3600, 2), F)
        for k in range(3600):
            th = F(F(k) / F(10.0)) * rpd
            t[k, 0] = libm.cosf(float(th))
            t[k, 1] = libm.sinf(float(th))
        _RAYS = t
    return _RAYS


def _as_u32(a: np.ndarray) -> np.ndarray:
    a = np.where(np.isnan(a), 0, a)
    return np.clip(np.trunc(a), 0, 4294967295).astype(np.int64)


def find_longest_line(img: np.ndarray, px, py, max_gap):
    h, w = img.shape
    t = ray_table()
    dx, dy = t[:, 0], t[:, 1]
    px, py, max_gap = F(px), F(py), F(max_gap)
    n = 3600
    x = np.full(n, px, F); y = np.full(n, py, F)
    xo = np.zeros(n, F); yo = np.zeros(n, F)
    gn = np.zeros(n, F); gx = np.zeros(n, F); gy = np.zeros(n, F)
    alive = np.ones(n, bool)
    while True:
        inb = alive & (x >= 0) & (y >= 0) & (x < F(w)) & (y < F(h))
        alive = inb
        if not alive.any():
            break
        idx = np.nonzero(alive)[0]
        v = img[y[idx].astype(np.int64), x[idx].astype(np.int64)]
        white = v == 255
        brk = (~white) & (gn[idx] >= max_gap)
        first = (~white) & ~brk & (gn[idx] == 0)
        more = (~white) & ~brk & ~first
        iw = idx[white]; gn[iw] = 0; gx[iw] = 0; gy[iw] = 0
        ib = idx[brk]; x[ib] = gx[ib]; y[ib] = gy[ib]; alive[ib] = False
        i1 = idx[first]; gn[i1] = 1; gx[i1] = x[i1]; gy[i1] = y[i1]
        im = idx[more]; gn[im] += F(1)
        adv = idx[~brk]
        xo[adv] += dx[adv]; yo[adv] += dy[adv]
        x[adv] = xo[adv] + px; y[adv] = yo[adv] + py
    xi, yi = _as_u32(x), _as_u32(y)
    ok = (xi < w) & (yi < h)
    zero = np.zeros(n, bool)
    zero[ok] = img[yi[ok], xi[ok]] == 0
    ex = np.where(zero, x - dx, px).astype(F)
    ey = np.where(zero, y - dy, py).astype(F)
    ln = ((px - ex) * (px - ex) + (py - ey) * (py - ey)).astype(F)
    m = ln.max()
    k = int(np.nonzero(ln == m)[0][-1])  # right-biased reduce: last index among maxima
    return np.array([px, py, ex[k], ey[k]], F), F(m)


def get_centre(img: np.ndarray, ptx, pty):
    h, w = img.shape
    ptx, pty = F(ptx), F(pty)

    def at(fx, fy):
        xi = min(int(max(fx, 0)), w - 1); yi = min(int(max(fy, 0)), h - 1)
        return img[yi, xi]

    left = ptx
    while left > 0 and abs(F(left - ptx)) < 5 and at(left, pty) == 255:
        left = F(left - F(1))
    right = ptx
    while right < F(w - 1) and abs(F(right - ptx)) < 5 and at(right, pty) == 255:
        right = F(right + F(1))
    up = pty
    while up > 0 and abs(F(up - pty)) < 5 and at(ptx, up) == 255:
        up = F(up - F(1))
    down = pty
    while down < F(h - 1) and abs(F(down - pty)) < 5 and at(ptx, down) == 255:
        down = F(down + F(1))
    return F((left + right) / F(2)), F((up + down) / F(2))


def find_lines(img: np.ndarray, max_gap: int, stats: dict | None = None):
    h, w = img.shape
    lines = []
    seeds = 0
    ys, xs = np.nonzero(img == 255)  # raster order
    for yi, xi in zip(ys, xs):
        x, y = F(xi), F(yi)
        sup = False
        for l in lines:
            dx, dy = F(l[2] - l[0]), F(l[3] - l[1])
            if dx == 0 and dy == 0:
                nx, ny = l[0], l[1]
            else:
                u = F(F(F(F(x - l[0]) * dx) + F(F(y - l[1]) * dy)) / F(F(dx * dx) + F(dy * dy)))
                nx, ny = F(l[0] + F(u * dx)), F(l[1] + F(u * dy))
            ex, ey = F(x - nx), F(y - ny)
            if F(F(ex * ex) + F(ey * ey)) < F(50.0):
                sup = True
                break
        if sup:
            continue
        cx, cy = get_centre(img, x, y)
        seeds += 1
        line, ln = find_longest_line(img, cx, cy, max_gap)
        if ln > F(2500.0):
            c1x, c1y = get_centre(img, line[2], line[3])
            lines.append(np.array([line[0], line[1], c1x, c1y], F))
            if len(lines) == 32:
                break
    if stats is not None:
        stats["seeds"] = seeds
    return np.array(lines, F).reshape(-1, 4)


# ---- next rows: independent restatements, written from the reference source with a
#      different structure than oracle/smh_oracle.c (vectorised lookups instead of pixel walks) ----------

def calc_meters_to_px_ratio(img: np.ndarray, scales):
    """src/vision/mpx_ratio.rs:3-133.  img: L8 [h, w]; scales: [(meters, (x, y))].  -> None | float."""
    h, w = img.shape
    black = np.zeros((h + 4, w), bool)          # D16: rows past the image are non-black
    black[:h] = img == 0
    # vbar[y, x]: (x, y) .. (x, y + 3) all black (the chained `(y..y - 4).rev()` range is empty)
    vbar = black[:h] & black[1:h + 1] & black[2:h + 2] & black[3:h + 3]
    found = []
    for meters, (x, y0) in scales:
        if y0 < 4 or x >= w:
            continue
        max_off = _round_half_away((20.0 / 640.0) * w)
        for y in range(y0, min(h, y0 + max_off)):
            if not black[y, x]:
                continue
            rs = np.flatnonzero(vbar[y, x:])
            if rs.size == 0 or x + rs[0] == 0:
                continue
            right = x + int(rs[0]) - 1
            ls = np.flatnonzero(vbar[y, :x])
            if ls.size == 0 or ls[-1] == 0:
                continue
            left = int(ls[-1]) + 1
            width = (right - left) & 0xFFFFFFFF
            if width < 10:
                continue
            found.append(float(meters) / float(width))
            break
    if not found:
        return None
    if len(found) == 1:
        return found[0]
    s = found[0]
    for v in found[1:]:
        s = s + v
    return s / float(len(found))


def low_edginess_map(view_bgra: np.ndarray) -> np.ndarray:
    """bool [h, w]: get_edginess(x, y) <= 0.01 (src/vision/find_minimap.rs:8-46,66) for interior pixels,
    through the literal f32 expression `max as f32 / 765.0 <= 0.01`."""
    v = view_bgra[..., :3].astype(np.int32)
    h, w = v.shape[:2]
    mx = np.zeros((h, w), np.int32)
    c = v[1:-1, 1:-1]
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            q = v[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
            mx[1:-1, 1:-1] = np.maximum(mx[1:-1, 1:-1], np.abs(c - q).sum(axis=2))
    low = (mx.astype(np.float32) / np.float32(765.0)) <= np.float32(0.01)
    low[0, :] = low[-1, :] = False
    low[:, 0] = low[:, -1] = False
    return low


def find_minimap(view_bgra: np.ndarray):
    """src/vision/find_minimap.rs:51-146 on an already cropped view.  -> None | (left, right, top, bottom)."""
    h, w = view_bgra.shape[:2]
    if w < 3 or h < 3:
        return None
    low = low_edginess_map(view_bgra)
    sx, sy = w // 2, h // 2

    def edge(horizontal: bool, cod: int) -> int:
        c_max = (w if horizontal else h) - 3
        oc_full = h if horizontal else w
        s_c, s_oc = (sx, sy) if horizontal else (sy, sx)
        need = (oc_full - s_oc) // 2 - 1
        oc_max = oc_full - 3
        pos = s_c
        while True:
            pos += cod
            if pos > c_max or pos < 0:
                return c_max + 2
            if pos < 3:
                return 0
            if not (low[s_oc, pos] if horizontal else low[pos, s_oc]):
                continue
            # the run goes against the search direction along the other axis
            lo, hi = (s_oc + 1, s_oc + need) if cod < 0 else (s_oc - need, s_oc - 1)
            if need > 0:
                if lo < 3 or hi > oc_max:
                    continue
                run = low[lo:hi + 1, pos] if horizontal else low[pos, lo:hi + 1]
                if not run.all():
                    continue
            return pos - cod

    return edge(True, -1), edge(True, 1), edge(False, -1), edge(False, 1)


# ---- marker icon template match (A20 / N4; the reference carries this stage commented out) -------------

def filter_map_marker_icons(cropped_map: np.ndarray, templates: np.ndarray):
    """vision-cpu/src/lib.rs:270-277 (candidates) + 282-355 on the ISOLATED cropped_map [h,w,3]; templates
    [n,size,size,4] RGBA.  -> (None | (x, y, sad), cropped_map after erase + stamp).  Vectorised over the
    candidates: an independent restatement of oracle/smh_oracle.c's loops (ties: raster-first, D17)."""
    h, w = cropped_map.shape[:2]
    n, size = templates.shape[:2]
    out = cropped_map.copy()
    if w < size + 2 or h < size + 2:
        return None, out
    cand = is_marker(cropped_map)
    cand[:size] = False; cand[h - size - 1:] = False
    cand[:, :size] = False; cand[:, w - size - 1:] = False
    ys, xs = np.nonzero(cand)  # raster order
    if len(ys) == 0:
        return None, out
    m16 = cropped_map.astype(np.int32)
    sad = np.zeros((len(ys), n), np.uint32)
    for my in range(size):
        for mx in range(size):
            t = templates[:, my, mx].astype(np.int32)  # [n, 4]
            if not t[:, 3].any():
                continue
            q = m16[ys + my, xs + mx]  # [c, 3]
            ad = np.abs(q[:, None, :] - t[None, :, :3]).sum(axis=2).astype(np.float32)  # [c, n]
            a = (t[:, 3].astype(np.float32) / np.float32(255.0)).astype(np.float32)
            sad += (ad * a[None, :]).astype(np.float32).astype(np.uint32)  # alpha == 0 contributes 0
    best = sad.min(axis=1)
    k = int(np.argmin(best))  # first minimum in raster order
    x, y, s = int(xs[k]), int(ys[k]), int(best[k])
    out[y:min(y + size, h - 1), x:min(x + size, w - 1)] = 0
    px, py = x + size // 2, y + _round_half_away(float(np.float32(size) * np.float32(0.70)))
    out[max(py - 2, 0):py + 2, max(px - 2, 0):px + 2] = (0, 255, 0)
    return (x, y, s), out
