"""The template-match stage the reference carries commented out: the C oracle against
the independent numpy restatement, the template fixture against a fresh run of the generator, and
hand-checkable cases."""
from pathlib import Path

import numpy as np
import pytest

import smh_b200  # noqa: F401
from helpers import marker_templates, stamp_icon
from oracle import np_oracle as N
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth

RES = Path("/root/reference/vision-common/src/markers/resources")


def _oracle_filter(frame, templates):
    o = O.Oracle()
    o.load_frame(frame)
    assert o.crop_to_map(True) is not None
    o.isolate_map_markers()
    before = o.cropped_map()
    o.load_map_markers(templates)
    hit = o.filter_map_marker_icons()
    after = o.cropped_map()
    o.mask_marker_lines()
    return o, before, hit, after


def test_templates_fixture_shape_and_content():
    t = marker_templates()
    assert t.shape == (5, 22, 22, 4) and t.dtype == np.uint8
    # every surviving pixel is marker-coloured, everything else is fully transparent black
    vis = t[..., 3] != 0
    assert N.is_marker(t[..., :3])[vis].all() and (t[~vis] == 0).all()
    assert (vis.reshape(5, -1).sum(axis=1) > 30).all()


@pytest.mark.skipif(not RES.is_dir(), reason="reference tree not mounted")
def test_templates_fixture_is_what_the_generator_makes():
    from oracle import marker_templates as M
    assert np.array_equal(M.load_markers(RES, 22), marker_templates())


@pytest.mark.parametrize("W,H,seed", [(1280, 720, 5), (1920, 1080, 9)])
def test_c_oracle_equals_numpy_restatement(W, H, seed):
    t = marker_templates()
    frame = synth.frames_host(synth.SynthParams(W, H), seed, 1)[0].copy()
    roi = O.roi(W, H)
    stamp_icon(frame, roi["map"], t[seed % 5], 200 + seed, 150)
    o, before, hit, after = _oracle_filter(frame, t)
    nhit, nafter = N.filter_map_marker_icons(before, t)
    assert hit == nhit and hit is not None
    assert np.array_equal(after, nafter)
    # the mask that follows sees the erased icon and the green square
    raw, dil = N.mask_marker_lines(nafter)
    assert np.array_equal(o.lsd_image(), dil)
    x, y, _ = hit
    assert (o.raw_mask()[y + 15 - 2:y + 15 + 2, x + 11 - 2:x + 11 + 2] == 255).all()


def test_exact_icon_is_found_near_its_position():
    """An icon drawn from a template matches with a small SAD at the marker-coloured pixel nearest its
    top-left corner (the candidate must itself be marker-coloured, vision-cpu/src/lib.rs:269-277)."""
    W, H = 1280, 720
    t = marker_templates()
    frame = synth.frames_host(synth.SynthParams(W, H, n_segments=0, n_blips=0), 0, 1)[0].copy()
    roi = O.roi(W, H)
    stamp_icon(frame, roi["map"], t[3], 300, 200)
    _, before, hit, after = _oracle_filter(frame, t)
    assert hit is not None
    x, y, sad = hit
    assert abs(x - 300) <= 22 and abs(y - 200) <= 22
    # erased square and 4x4 green stamp (lib.rs:335-351)
    assert (after[y:y + 22, x:x + 22].reshape(-1, 3).sum(axis=1) == 0).sum() >= 22 * 22 - 16
    assert (after[y + 15 - 2:y + 15 + 2, x + 11 - 2:x + 11 + 2] == (0, 255, 0)).all()


def test_no_candidates_no_match():
    W, H = 1280, 720
    frame = synth.frames_host(synth.SynthParams(W, H, n_segments=0, n_blips=0), 1, 1)[0]
    _, before, hit, after = _oracle_filter(frame, marker_templates())
    assert hit is None and np.array_equal(before, after)


def test_oracle_reproduces_the_icon_goldens():
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_golden_icons", Path(__file__).with_name("golden") / "make_golden_icons.py")
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    gold = json.loads((Path(__file__).with_name("golden") / "golden_icons.json").read_text())
    from helpers import sha
    for case in G.CASES:
        hit, mask, lines = G.run_oracle(G.frame_of(case))
        rec = gold[case[0]]
        assert (list(hit) if hit else None) == rec["hit"], case[0]
        assert sha(mask) == rec["lsd_image_sha"] and int((mask == 255).sum()) == rec["white_px"]
        assert [[f"{int(b):08x}" for b in row] for row in lines.view(np.uint32)] == rec["lines_bits"]
