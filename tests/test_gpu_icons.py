"""Marker icon template match on the device (SURVEY.md A20 / N4 — disabled in the reference) against the
oracle: per-frame trait order and the batch flag; bit-exact position, SAD, edited mask and lines."""
import numpy as np
import pytest
import torch

import smh_b200  # noqa: F401
from helpers import marker_templates, stamp_icon
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.capi import SMH_E_STATE, VisionError
from squad_mortar_helper_b200.vision import B200Vision, DebugView

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vision():
    O.set_threads(8)
    v = B200Vision.init(0)
    yield v
    v.close()


def _oracle(frame, templates, max_gap=15):
    o = O.Oracle()
    o.load_frame(frame)
    assert o.crop_to_map(True) is not None
    o.isolate_map_markers()
    o.load_map_markers(templates)
    hit = o.filter_map_marker_icons()
    view = o.get_debug_view(3)
    o.mask_marker_lines()
    return hit, o.lsd_image(), o.find_marker_lines(max_gap), view


def _frames(W, H, n, first, icons=True, **kw):
    t = marker_templates()
    fr = synth.frames_host(synth.SynthParams(W, H, **kw), first, n).copy()
    roi = O.roi(W, H)
    _, _, w, h = roi["map"]
    rng = np.random.default_rng(first)
    if icons:
        for i in range(n):
            for _ in range(int(rng.integers(1, 3))):
                stamp_icon(fr[i], roi["map"], t[int(rng.integers(0, 5))], int(rng.integers(30, w - 60)), int(rng.integers(30, h - 60)))
    return fr, t


@pytest.mark.parametrize("W,H", [(1280, 720), (1920, 1080), (2560, 1440)])
def test_trait_order_matches_oracle(vision, W, H):
    fr, t = _frames(W, H, 3, 40)
    vision.load_map_markers(t)
    for i, frame in enumerate(fr):
        want_hit, want_mask, want_lines, want_view = _oracle(frame, t)
        vision.load_frame(frame)
        assert vision.crop_to_map(True) is not None
        vision.isolate_map_markers()
        hit = vision.filter_map_marker_icons()
        assert hit == want_hit and hit is not None, (i, hit, want_hit)
        with pytest.raises(VisionError) as e:
            vision.filter_map_marker_icons()
        assert e.value.code == SMH_E_STATE
        vision.mask_marker_lines()
        assert np.array_equal(vision.lsd_image(), want_mask), f"frame {i}: edited mask"
        lines = vision.find_marker_lines(15)
        assert lines.shape == want_lines.shape and np.array_equal(lines.view(np.uint32), want_lines.view(np.uint32))
        assert np.array_equal(vision.get_debug_view(DebugView.LSDPreprocess), want_view)
        assert np.array_equal(vision.get_debug_view(DebugView.LSDInput)[..., 0], want_mask)


def test_no_candidates_and_unfiltered_frames_are_untouched(vision):
    W, H = 1280, 720
    fr, t = _frames(W, H, 1, 3, icons=False, n_segments=0, n_blips=0)
    vision.load_map_markers(t)
    vision.load_frame(fr[0])
    assert vision.crop_to_map(True) is not None
    before = vision.lsd_image().copy()
    assert vision.filter_map_marker_icons() is None
    assert np.array_equal(vision.lsd_image(), before)


def test_batch_flag_matches_oracle(vision):
    W, H, n = 1920, 1080, 12
    fr, t = _frames(W, H, n, 70)
    fr[5] = synth.frames_host(synth.SynthParams(W, H, map_open=False), 75, 1)[0]
    vision.load_map_markers(t)
    dev = torch.from_numpy(fr).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for env in (None, ("5", "2")):
        mp = pytest.MonkeyPatch()
        if env:
            mp.setenv("SMH_B200_BATCH_CHUNK", env[0]); mp.setenv("SMH_B200_BATCH_LANES", env[1])
        b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, filter_icons=True, stream=st)
        mp.undo()
        torch.cuda.synchronize()

        def arr(ptr, nbytes, dt):
            class _H:
                pass
            h = _H()
            h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
            return torch.as_tensor(h, device="cuda").cpu().numpy().view(dt)
        found, xy, sad = arr(b.icon_found, n, np.uint8), arr(b.icon_xy, n * 8, np.uint32).reshape(n, 2), arr(b.icon_sad, n * 4, np.uint32)
        lines, nl = arr(b.lines, n * 512, np.float32).reshape(n, 32, 4), arr(b.n_lines, n * 4, np.uint32)
        mask = vision.download_plane(b.lsd_image, 0, n)
        for i in range(n):
            if i == 5:
                assert not found[i] and nl[i] == 0
                continue
            want_hit, want_mask, want_lines, _ = _oracle(fr[i], t)
            assert found[i] and (int(xy[i, 0]), int(xy[i, 1]), int(sad[i])) == want_hit, (i, xy[i], sad[i], want_hit)
            assert np.array_equal(mask[i], want_mask), f"frame {i}: edited mask"
            assert nl[i] == len(want_lines) and np.array_equal(lines[i, :nl[i]].view(np.uint32), want_lines.view(np.uint32))
    # without the flag the same frames give the unedited result
    b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, stream=st)
    torch.cuda.synchronize()
    assert b.icon_found is None


def test_host_batch_with_icons(vision):
    W, H, n = 1280, 720, 9
    fr, t = _frames(W, H, n, 120)
    vision.load_map_markers(t)
    out = vision.process_host_batch(fr, max_gap=15, grayscale=True, chunk=4, want=("ui_map", "lsd_image", "icons"))
    for i in range(n):
        want_hit, want_mask, want_lines, _ = _oracle(fr[i], t)
        assert out["icon_found"][i] and (int(out["icon_xy"][i, 0]), int(out["icon_xy"][i, 1]), int(out["icon_sad"][i])) == want_hit
        assert np.array_equal(out["lsd_image"][i], want_mask)
        m = int(out["n_lines"][i])
        assert m == len(want_lines) and np.array_equal(out["lines"][i, :m].view(np.uint32), want_lines.view(np.uint32))


def test_icon_goldens(vision):
    """The committed goldens of the stage (tests/golden/golden_icons.json) through the trait-order calls."""
    import importlib.util
    import json
    from pathlib import Path
    from helpers import sha
    gdir = Path(__file__).with_name("golden")
    spec = importlib.util.spec_from_file_location("make_golden_icons", gdir / "make_golden_icons.py")
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    gold = json.loads((gdir / "golden_icons.json").read_text())
    vision.load_map_markers(marker_templates())
    for case in G.CASES:
        rec = gold[case[0]]
        vision.load_frame(G.frame_of(case))
        assert vision.crop_to_map(True) is not None
        vision.isolate_map_markers()
        hit = vision.filter_map_marker_icons()
        assert (list(hit) if hit else None) == rec["hit"], case[0]
        vision.mask_marker_lines()
        assert sha(vision.lsd_image()) == rec["lsd_image_sha"]
        lines = vision.find_marker_lines(15)
        assert [[f"{int(b):08x}" for b in row] for row in lines.view(np.uint32)] == rec["lines_bits"]
