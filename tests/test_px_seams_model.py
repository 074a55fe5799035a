"""The seam protocol of the fused pixel pass, checked on the CPU as an algorithm (no CUDA code runs here).

k_pixels2 (csrc/kernels_px.cuh) works in strips of 128 padded columns and reads no pixel of another strip: per (strip,
side, row) it leaves one flag byte — bit 0 the outermost pixel is a marker pixel, bits 1-3 the OCR core flags of the
outermost three pixels, bits 4-6 the OCR edge pixels among them that found no core pixel inside the strip — and
k_px_seams finishes the pixels either side of a seam from those bytes alone.  This file restates both halves in numpy /
Python (strip-local pass with its flag bytes, then a line-by-line port of k_px_seams' decisions) and checks on random flag
images that the result equals the global definition: the 3x3-cross dilation of the marker mask and "an edge pixel keeps its
value iff a core pixel lies in its 7x7 window" (vision-cpu/src/lib.rs:196-216, 372) over the whole map."""
import numpy as np
import pytest

R = 3  # SMH_OCR_RADIUS


def _global(m, c, e):
    """dilated mask; set of edge pixels with a core pixel in their 7x7 window"""
    h, w = m.shape
    d = m.copy()
    d[:, 1:] |= m[:, :-1]; d[:, :-1] |= m[:, 1:]; d[1:, :] |= m[:-1, :]; d[:-1, :] |= m[1:, :]
    cp = np.zeros((h + 2 * R, w + 2 * R), bool)
    cp[R:R + h, R:R + w] = c
    near = np.zeros((h, w), bool)
    for dy in range(2 * R + 1):
        for dx in range(2 * R + 1):
            near |= cp[dy:dy + h, dx:dx + w]
    return d, e & near


def _strips_and_seams(m, c, e, pad, bqh):
    h, w = m.shape
    nstrips = (pad + w + 127) // 128
    d = np.zeros((h, w), bool)
    found = np.zeros((h, w), bool)
    flags = np.zeros((nstrips, 2, h), np.uint8)
    for s in range(nstrips):
        x0, x1 = max(128 * s - pad, 0), min(128 * (s + 1) - pad, w)   # map columns of the strip
        ms, cs, es = m[:, x0:x1], c[:, x0:x1], e[:, x0:x1]
        ds, fs = _global(ms, cs, es)                                      # the strip as if nothing lay beside it
        d[:, x0:x1] = ds
        found[:, x0:x1] = fs
        pend = es & ~fs
        full_left, full_right = 128 * s - pad >= 0, 128 * (s + 1) - pad <= w   # the strip's outermost columns exist
        for y in range(h):
            if full_left:
                flags[s, 0, y] = int(ms[y, 0]) | sum(int(cs[y, k]) << (1 + k) for k in range(min(3, x1 - x0))) \
                    | sum(int(pend[y, k]) << (4 + k) for k in range(min(3, x1 - x0)))
            if full_right:
                n = x1 - x0
                flags[s, 1, y] = int(ms[y, n - 1]) | sum(int(cs[y, n - 3 + k]) << (1 + k) for k in range(3)) \
                    | sum(int(pend[y, n - 3 + k]) << (4 + k) for k in range(3))
    # ---- k_px_seams, decision by decision --------------------------------------------------------------------------
    for b in range(1, nstrips):
        X0 = 128 * b
        L, Rr = flags[b - 1, 1], flags[b, 0]
        for j in range(h):
            lb, rb = int(L[j]), int(Rr[j])
            if not (lb | rb):
                continue
            for side, byte in ((0, lb), (1, rb)):
                if byte & 1:
                    i = (X0 - 1 if side else X0) - pad
                    if 0 <= i < w:
                        d[j, i] = True
            pl, pr = (lb >> 4) & 7, (rb >> 4) & 7
            if not (pl | pr):
                continue
            cl = cr = 0
            for jj in range(max(j - R, bqh), min(j + R, min(2 * bqh, h) - 1) + 1):
                cl |= (int(L[jj]) >> 1) & 7
                cr |= (int(Rr[jj]) >> 1) & 7
            fix = 0
            for k in range(3):
                if (pl >> k) & 1 and cr & ((2 << k) - 1):
                    fix |= 1 << k
                if (pr >> k) & 1 and cl & (7 & ~((1 << k) - 1)):
                    fix |= 8 << k
            for k in range(6):
                if (fix >> k) & 1:
                    found[j, X0 - 3 + k - pad] = True
    return d, found


@pytest.mark.parametrize("seed", range(12))
def test_strips_plus_seams_equal_the_global_definition(seed):
    rng = np.random.default_rng(seed)
    pad = int(rng.integers(0, 4))
    w = int(rng.choice([130, 255, 256, 257, 300, 384, 400, 515]))
    h = int(rng.integers(12, 40))
    bqw, bqh = w // 2, h // 2
    density = float(rng.choice([0.01, 0.05, 0.3]))
    m = rng.random((h, w)) < density
    inq = np.zeros((h, w), bool)
    inq[bqh:2 * bqh, bqw:2 * bqw] = True
    # OCR core pixels: inside the quadrant, below the clips of D10 (column x - bqw <= bqw - R, row y - bqh <= bqh - R)
    clip = np.zeros((h, w), bool)
    clip[bqh:bqh + max(bqh - R + 1, 0), bqw:bqw + max(bqw - R + 1, 0)] = True
    c = (rng.random((h, w)) < density) & inq & clip
    e = (rng.random((h, w)) < 2 * density) & inq & ~c
    want_d, want_f = _global(m, c, e)
    got_d, got_f = _strips_and_seams(m, c, e, pad, bqh)
    assert np.array_equal(got_d, want_d), np.argwhere(got_d != want_d)[:5]
    assert np.array_equal(got_f, want_f), np.argwhere(got_f != want_f)[:5]


def test_content_only_at_the_seams():
    """Marker / core / edge pixels only in the three columns either side of every seam: each needs its partner across it."""
    rng = np.random.default_rng(99)
    for pad in range(4):
        w, h = 400, 30
        bqw, bqh = w // 2, h // 2
        near = np.zeros(w, bool)
        for s in range(1, (pad + w + 127) // 128):
            cx = 128 * s - pad
            near[max(cx - 3, 0):min(cx + 3, w)] = True
        m = (rng.random((h, w)) < 0.4) & near[None, :]
        inq = np.zeros((h, w), bool)
        inq[bqh:2 * bqh, bqw:2 * bqw] = True
        clip = np.zeros((h, w), bool)
        clip[bqh:bqh + bqh - R + 1, bqw:bqw + bqw - R + 1] = True
        c = (rng.random((h, w)) < 0.15) & inq & clip & near[None, :]
        e = (rng.random((h, w)) < 0.5) & inq & ~c & near[None, :]
        want_d, want_f = _global(m, c, e)
        got_d, got_f = _strips_and_seams(m, c, e, pad, bqh)
        assert np.array_equal(got_d, want_d) and np.array_equal(got_f, want_f), pad
        assert want_f.any() and (want_d & ~m).any()
