"""The callers either side of the batch path that round 2 added: several GPUs behind one call
(smh_b200_multi_process_host_batch), streaming over a ring of host buffers (smh_b200_process_host_stream), the L8
ui_map hand-over (SMH_BATCH_UI_L8), the Timeshares ABI, and the batch semantics of the scales rows above start_y."""
import numpy as np
import pytest
import torch

import smh_b200  # noqa: F401
from helpers import oracle_run
from squad_mortar_helper_b200 import capi, synth
from squad_mortar_helper_b200.vision import B200Vision, B200VisionMulti

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vision():
    with B200Vision.init(0) as v:
        yield v


def _same(a: dict, b: dict, n):
    for k in ("ui_map", "ocr_out", "scales", "lsd_image", "n_lines", "red_pixels", "map_open"):
        if k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    for i in range(n):
        m = int(a["n_lines"][i])
        assert np.array_equal(a["lines"][i, :m].view(np.uint32), b["lines"][i, :m].view(np.uint32)), ("lines", i)


def test_multi_gpu_call_equals_one_gpu(vision):
    """Frames split over the GPUs of the box (contiguous ranges, a host thread + handle each) give, in the caller's
    frame order, exactly what one GPU gives — and what the oracle gives."""
    W, H, n = 1280, 720, 23
    frames = np.concatenate([synth.frames_host(synth.SynthParams(W, H, map_open=(i % 7 != 3)), 40 + i, 1) for i in range(n)])
    start_y = (np.arange(n, dtype=np.uint32) * 5) % 40
    want = ("ui_map", "ocr_out", "scales", "lsd_image")
    one = vision.process_host_batch(frames, start_y=start_y, max_gap=15, grayscale=True, chunk=4, want=want)
    n_dev = torch.cuda.device_count()
    for devs in ([0], list(range(min(n_dev, 2))), list(range(n_dev))):
        with B200VisionMulti(devs) as m:
            got = m.process_host_batch(frames, start_y=start_y, max_gap=15, grayscale=True, chunk=4, want=want)
            _same(one, got, n)
            assert len(m.last_times()) == len(devs) and all(t > 0 for t in m.last_times())
            cover = []
            for g in range(len(devs)):
                a, c = m.shard_range(n, g)
                cover += list(range(a, a + c))
            assert cover == list(range(n))
    for i in (0, 3, 11, 22):
        w_ = oracle_run(frames[i], True, int(start_y[i]), 15)
        assert bool(one["map_open"][i]) == w_["map_open"]
        if w_["map_open"]:
            assert np.array_equal(one["lsd_image"][i], w_["lsd_image"]) and np.array_equal(one["ui_map"][i], w_["ui_map"])
            assert np.array_equal(one["lines"][i, :one["n_lines"][i]].view(np.uint32), w_["lines"].view(np.uint32))
    if n_dev < 2:
        pytest.skip("one GPU visible: the 2-GPU equality runs on the multi-GPU box (gpurun --gpus 2)")


def test_stream_over_a_ring_of_host_buffers(vision):
    """smh_b200_process_host_stream: frame f is ring[f % R], its results land at index f % R."""
    W, H, R, n = 1280, 720, 16, 40
    ring = synth.frames_host(synth.SynthParams(W, H), 7, R)
    want = ("ui_map", "ocr_out", "scales", "lsd_image")
    ref = vision.process_host_batch(ring, max_gap=15, grayscale=True, chunk=4, want=want)
    got = vision.process_host_batch(ring, max_gap=15, grayscale=True, chunk=4, want=want, ring=R, n_total=n)
    _same(ref, got, R)   # every slot was written by a frame with the slot's content
    with pytest.raises(capi.VisionError):
        vision.process_host_batch(ring[:8], max_gap=15, chunk=4, ring=8, n_total=n)  # ring < 4 chunks


def test_ui_map_as_its_l_plane(vision):
    W, H, n = 1280, 720, 6
    frames = synth.frames_host(synth.SynthParams(W, H), 3, n)
    a = vision.process_host_batch(frames, max_gap=15, grayscale=True, chunk=2, want=("ui_map",))
    b = vision.process_host_batch(frames, max_gap=15, grayscale=True, chunk=2, want=("ui_map",), ui_l8=True)
    assert b["ui_map"].shape == a["ui_map"].shape[:3]
    assert np.array_equal(a["ui_map"][..., 0], b["ui_map"]) and np.array_equal(a["ui_map"][..., 2], b["ui_map"])
    assert (a["ui_map"][..., 3] == 255).all()
    with pytest.raises(capi.VisionError):
        vision.process_host_batch(frames, grayscale=False, chunk=2, want=("ui_map",), ui_l8=True)


def test_batch_scales_rows_above_start_y_are_zero_whatever_the_chunking(vision):
    """Every frame of a batch starts from a fresh scales buffer (D9): rows above start_y[f] are zero, whatever frame
    used the pipeline slot before and however the batch is chunked."""
    W, H, n = 1280, 720, 13
    frames = synth.frames_host(synth.SynthParams(W, H, n_labels=3), 0, n)
    start_y = (np.arange(n, dtype=np.uint32) * 17 + 5) % 120
    outs = [vision.process_host_batch(frames, start_y=start_y, max_gap=15, chunk=c, want=("scales",))["scales"].copy() for c in (2, 5, 13)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    for i in range(n):
        assert not outs[0][i, :start_y[i]].any()
        w_ = oracle_run(frames[i], True, int(start_y[i]), 15)   # a fresh oracle: zeros above start_y as well
        assert np.array_equal(outs[0][i], w_["scales"])
    dev = torch.from_numpy(frames).cuda()
    b = vision.process_batch(dev.data_ptr(), n, W, H, start_y=start_y, max_gap=15, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(vision.download_plane(b.scales, 0, n), outs[0])


def test_timeshares_of_a_frame_through_the_trait(vision):
    """The reference's Timeshares (vision-common/src/debug.rs:3-29): eleven stages, wall-clock + device time."""
    frame = synth.frames_host(synth.SynthParams(1920, 1080), 1, 1)[0]
    for _ in range(2):  # the second frame: no allocation, table build or first-use initialisation inside a stage
        res = vision.process(frame, grayscale=True, detect_markers=True, max_gap=15,
                             ocr=lambda img: ([(100, (img.shape[1] // 2, img.shape[0] // 2))], img.shape[0] // 3))
        assert res is not None
    ts = vision.timeshares()
    assert list(ts) == capi.STAGES and len(ts) == 11
    for name in ("load_frame", "crop_to_map", "find_minimap", "ocr_preprocess", "ocr", "find_scales_preprocess",
                 "calc_meters_to_px_ratio", "find_marker_lines"):
        assert ts[name][0] > 0.0, (name, ts)
    assert ts["isolate_map_markers"][0] >= 0.0 and ts["mask_marker_lines"][0] >= 0.0  # (no device work left in them: ~100 ns)
    for name in ("load_frame", "crop_to_map", "find_minimap", "find_marker_lines"):
        assert 0.0 < ts[name][1] < 500.0, (name, ts)
    assert ts["isolate_map_markers"][1] == 0.0 and ts["ocr"][1] == 0.0
