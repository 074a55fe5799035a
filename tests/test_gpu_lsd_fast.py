"""k_lsd2 (one warp per frame, exact ray filter) against the oracle's find_lines: scenes from sparse to stress,
max_gap over the whole range the fast path takes (1..32), sizes from tiny to 4K, through the trait call and
the batch call.  Bit-exact f32 end points."""
import numpy as np
import pytest
import torch

import smh_b200  # noqa: F401
from helpers import oracle_run
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vision():
    with B200Vision.init(0) as v:
        yield v


def _lines_of(vision, frame, gap):
    vision.load_frame(frame)
    assert vision.crop_to_map(True) is not None
    vision.isolate_map_markers(); vision.mask_marker_lines()
    return vision.find_marker_lines(gap)


SCENES = [
    (1280, 720, dict(n_segments=5, n_blips=6, n_labels=1)),
    (1920, 1080, synth.DENSE),
    (2560, 1440, synth.DENSE),
    (1024, 768, dict(n_segments=12, n_blips=30, n_labels=2)),
    (1280, 720, dict(n_segments=60, n_blips=0, n_labels=0)),            # > 32 lines: early exit
    (1280, 720, dict(n_segments=3, n_blips=300, n_labels=0)),           # blips everywhere: dead clusters
    (1600, 1024, dict(n_segments=24, n_blips=200, n_labels=2, blip_band_pct=15)),
]


@pytest.mark.parametrize("W,H,kw", SCENES, ids=[f"{w}x{h}-{k['n_segments']}s{k['n_blips']}b" for w, h, k in SCENES])
@pytest.mark.parametrize("gap", [15, 22, 1, 2, 3, 7, 16, 31, 32])
def test_trait_path_matches_oracle(vision, W, H, kw, gap):
    if gap not in (15, 22) and (W, H) == (2560, 1440):
        pytest.skip("the large frame runs the production gaps only")
    O.set_threads(8)
    for seed in (0, 7):
        frame = synth.frames_host(synth.SynthParams(W, H, **kw), seed, 1)[0]
        got = _lines_of(vision, frame, gap)
        want = oracle_run(frame, True, 0, gap)["lines"]
        assert got.shape == want.shape, f"seed {seed}: {len(got)} lines, oracle {len(want)}\n{got}\n{want}"
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"seed {seed}\n{got}\n{want}"


def test_hand_made_masks_through_the_batch_call(vision):
    """Lines along and through the borders, discs, dashes at exactly max_gap: painted into the map ROI in a
    marker colour so that the dilated mask is known, compared with the oracle frame by frame."""
    W, H = 1280, 720
    roi = O.roi(W, H)
    mx, my, mw, mh = roi["map"]
    base = synth.frames_host(synth.SynthParams(W, H, **synth.PLAIN), 0, 1)[0]
    col = np.array([2, 253, 64, 255], np.uint8)  # BGRA of the alpha marker colour
    frames = []
    rng = np.random.default_rng(4)
    for k in range(6):
        f = base.copy()
        m = f[my:my + mh, mx:mx + mw]
        if k == 0:
            m[2:5, 0:300] = col; m[0:400, 1:4] = col; m[mh - 4:mh - 1, 50:mw] = col; m[100:mh, mw - 3:mw] = col
        elif k == 1:
            for t in range(0, min(mw, mh) - 3):
                m[t:t + 3, t:t + 3] = col
            for t in range(0, min(mw, mh) - 3):
                m[t:t + 3, mw - 4 - t:mw - 1 - t] = col
        elif k == 2:
            for x in range(20, mw - 20, 19):
                m[200:203, x:x + 4] = col           # dashes: 4 on, 15 off
            for y in range(20, mh - 20, 21):
                m[y:y + 4, 300:303] = col           # 4 on, 17 off: broken at gap 15, whole at 22
        elif k == 3:
            yy, xx = np.mgrid[0:mh, 0:mw]
            for _ in range(12):
                cx, cy, r = rng.integers(30, mw - 30), rng.integers(30, mh - 30), rng.integers(3, 26)
                m[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = col
        elif k == 4:
            for _ in range(40):
                x0, y0 = rng.integers(5, mw - 70), rng.integers(5, mh - 70)
                for t in range(rng.integers(20, 64)):
                    m[y0 + t // 2:y0 + t // 2 + 3, x0 + t:x0 + t + 3] = col
        else:
            sel = rng.random((mh, mw)) < 0.02
            m[sel] = col
        frames.append(f)
    frames = np.stack(frames)
    dev = torch.from_numpy(frames).cuda()
    for gap in (15, 22, 4):
        b = vision.process_batch(dev.data_ptr(), len(frames), W, H, max_gap=gap, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        n_lines = _dev(b.n_lines, len(frames) * 4).view(np.uint32)
        lines = _dev(b.lines, len(frames) * 512).view(np.float32).reshape(len(frames), 32, 4)
        for i in range(len(frames)):
            want = oracle_run(frames[i], True, 0, gap)["lines"]
            got = lines[i, :n_lines[i]]
            assert got.shape == want.shape, f"frame {i} gap {gap}: {len(got)} lines, oracle {len(want)}"
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"frame {i} gap {gap}\n{got}\n{want}"


def _dev(ptr, nbytes):
    class _H:
        pass
    h = _H()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda").cpu().numpy()


def test_stress_scenes_and_counters(vision):
    """BASELINE config 5 scenes through the fast path; its seed counter equals the exhaustive search's (= the oracle's)."""
    W, H = 2560, 1440
    for kw in (synth.STRESS_SEGMENTS, dict(n_segments=24, n_blips=400, n_labels=2, blip_band_pct=12)):
        p = synth.SynthParams(W, H, **kw)
        n = 2
        dev = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
        synth.frames_device(vision, p, 0, n, dev.data_ptr())
        b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, lsd_stats=True, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        n_lines = _dev(b.n_lines, n * 4).view(np.uint32)
        lines = _dev(b.lines, n * 512).view(np.float32).reshape(n, 32, 4)
        stats = _dev(b.lsd_stats, n * 24).view(np.uint64).reshape(n, 3)
        host = dev.cpu().numpy()
        for i in range(n):
            want = oracle_run(host[i], True, 0, 15)
            got = lines[i, :n_lines[i]]
            assert got.shape == want["lines"].shape
            assert np.array_equal(got.view(np.uint32), want["lines"].view(np.uint32))
            assert int(stats[i, 0]) == want["lsd_stats"][0], "seeds visited"
            assert int(stats[i, 1]) <= want["lsd_stats"][1], "the filter may only remove ray steps"
