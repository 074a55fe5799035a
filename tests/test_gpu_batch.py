"""Batch extension (N frames per launch), device-resident and host-resident inputs, plus the
size-independent properties used at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

import smh_b200  # noqa: F401
from helpers import assert_same, gpu_run, oracle_run
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth
from squad_mortar_helper_b200.vision import B200Vision

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vision():
    O.set_threads(8)
    v = B200Vision.init(0)
    yield v
    v.close()


def _fetch(v, b, n):
    """Download every plane of a device batch into packed host arrays."""
    out = {"ui_map": v.download_plane(b.ui_map, 0, n), "lsd_image": v.download_plane(b.lsd_image, 0, n),
           "ocr_out": v.download_plane(b.ocr_out, 0, n), "scales": v.download_plane(b.scales, 0, n)}

    return out


def _device_array(ptr, nbytes):
    """Wrap a raw device pointer owned by the library as a torch uint8 tensor (plumbing only)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda")


def _results(v, b, n):
    torch.cuda.synchronize()
    out = _fetch(v, b, n)
    out["lines"] = _device_array(b.lines, n * 128 * 4).cpu().numpy().view(np.float32).reshape(n, 32, 4)
    out["n_lines"] = _device_array(b.n_lines, n * 4).cpu().numpy().view(np.uint32)
    out["red_pixels"] = _device_array(b.red_pixels, n * 4).cpu().numpy().view(np.uint32)
    out["map_open"] = _device_array(b.map_open, n).cpu().numpy()
    if b.lsd_stats:
        out["stats"] = _device_array(b.lsd_stats, n * 24).cpu().numpy().view(np.uint64).reshape(n, 3)
    return out


def _frame_result(res, i):
    n = int(res["n_lines"][i])
    return {"map_open": bool(res["map_open"][i]), "red_pixels": int(res["red_pixels"][i]),
            "ui_map": res["ui_map"][i], "lsd_image": res["lsd_image"][i], "ocr_out": res["ocr_out"][i],
            "scales": res["scales"][i], "lines": res["lines"][i, :n].copy()}


@pytest.mark.parametrize("W,H,n", [(1280, 720, 9), (1920, 1080, 5), (2560, 1440, 20)])
def test_device_batch_matches_oracle_per_frame(vision, W, H, n):
    p = synth.SynthParams(W, H)
    host = synth.frames_host(p, 100, n)
    dev = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(vision, p, 100, n, dev.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host), "CUDA rasteriser != CPU rasteriser"
    start_y = np.arange(n, dtype=np.uint32) * 7 % 50
    b = vision.process_batch(dev.data_ptr(), n, W, H, start_y=start_y, max_gap=15, grayscale=True, lsd_stats=True,
                             lsd_exhaustive=True, stream=torch.cuda.current_stream().cuda_stream)
    res = _results(vision, b, n)
    # default mode (isolated-blob pruning on) must give the same lines with fewer marched rays
    b2 = vision.process_batch(dev.data_ptr(), n, W, H, start_y=start_y, max_gap=15, grayscale=True, lsd_stats=True,
                              stream=torch.cuda.current_stream().cuda_stream)
    res2 = _results(vision, b2, n)
    assert np.array_equal(res["n_lines"], res2["n_lines"]) and np.array_equal(res["lines"].view(np.uint32), res2["lines"].view(np.uint32))
    assert np.array_equal(res["stats"][:, 0], res2["stats"][:, 0]) and (res2["stats"][:, 1] <= res["stats"][:, 1]).all()
    assert (res["stats"][:, 2] == 0).all()
    for i in range(n):
        want = oracle_run(host[i], True, int(start_y[i]), 15)
        got = _frame_result(res, i)
        got["xywh"] = want["xywh"]
        assert_same(got, want, f"batch frame {i}", scales_from=int(start_y[i]))
        assert tuple(int(x) for x in res["stats"][i][:2]) == tuple(want["lsd_stats"]), "seed / ray-step counters"


def test_batch_with_closed_maps_and_rgb(vision):
    W, H, n = 1280, 720, 6
    frames = np.concatenate([synth.frames_host(synth.SynthParams(W, H, map_open=(i % 2 == 0)), i, 1) for i in range(n)])
    dev = torch.from_numpy(frames).cuda()
    b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, grayscale=False,
                             stream=torch.cuda.current_stream().cuda_stream)
    res = _results(vision, b, n)
    for i in range(n):
        want = oracle_run(frames[i], False, 0, 15)
        assert bool(res["map_open"][i]) == want["map_open"] == (i % 2 == 0)
        assert int(res["red_pixels"][i]) == want["red_pixels"]
        if want["map_open"]:
            got = _frame_result(res, i); got["xywh"] = want["xywh"]
            assert_same(got, want, f"frame {i}")
        else:
            assert res["n_lines"][i] == 0


@pytest.mark.parametrize("gray,W,H,n,chunk", [(True, 1920, 1080, 11, 4), (False, 1920, 1080, 7, 3), (True, 1366, 768, 23, 5)])
def test_host_batch_end_to_end(vision, gray, W, H, n, chunk):
    """Host frames in, host results out.  With a grayscale ui_map only its L plane crosses the bus and the
    library's host threads widen it into the caller's RGBA buffer; the colour ui_map crosses as RGBA."""
    p = synth.SynthParams(W, H)
    frames = synth.frames_host(p, 7, n)
    start_y = (np.arange(n, dtype=np.uint32) * 13) % 90
    out = vision.process_host_batch(frames, start_y=start_y, max_gap=15, grayscale=gray, chunk=chunk,
                                    want=("ui_map", "ocr_out", "scales", "lsd_image"))
    for i in range(n):
        want = oracle_run(frames[i], gray, int(start_y[i]), 15)
        got = _frame_result(out, i); got["xywh"] = want["xywh"]
        # pipelined chunks reuse plane buffers: rows above start_y hold an earlier frame's data
        sy = int(start_y[i])
        got["scales"] = got["scales"].copy(); got["scales"][:sy] = 0
        assert_same(got, want, f"host batch frame {i}")


def test_full_size_properties_1440p(vision):
    """BASELINE config 3 shape (1440p, dense) at a batch big enough to fill the GPU several times:
    (1) batch == per-frame trait path, (2) idempotent across runs, (3) the line search alone
    (SKIP_PIXELS) reproduces the fused run, (4) structural invariants of every frame."""
    W, H, n = 2560, 1440, 320
    p = synth.SynthParams(W, H)
    dev = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(vision, p, 0, n, dev.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, stream=st)
    r1 = _results(vision, b, n)
    lines1, nl1, mask1 = r1["lines"].copy(), r1["n_lines"].copy(), r1["lsd_image"].copy()
    b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, stream=st)
    r2 = _results(vision, b, n)
    assert np.array_equal(nl1, r2["n_lines"]) and np.array_equal(mask1, r2["lsd_image"])
    for i in range(n):
        assert np.array_equal(lines1[i, :nl1[i]].view(np.uint32), r2["lines"][i, :nl1[i]].view(np.uint32))
    b = vision.process_batch(0, n, W, H, max_gap=15, skip_pixels=True, stream=st)
    r3 = _results(vision, b, n)
    assert np.array_equal(nl1, r3["n_lines"])
    # invariants: ui_map alpha 255 and grey; mask is 0/255; ocr in {255} U (255-luma); open flags
    assert (r1["ui_map"][..., 3] == 255).all() and (r1["ui_map"][..., 0] == r1["ui_map"][..., 2]).all()
    assert np.isin(mask1, (0, 255)).all() and np.isin(r1["scales"], (0, 255)).all()
    assert r1["map_open"].all() and (nl1 <= 32).all() and (nl1 > 0).all()
    # every accepted line is longer than 50 px before its end point was re-centred, and lies in the map
    w, h = b.lsd_image.width, b.lsd_image.height
    for i in range(0, n, 37):
        L = lines1[i, :nl1[i]]
        assert (L[:, [0, 2]] >= 0).all() and (L[:, [0, 2]] <= w).all() and (L[:, [1, 3]] >= 0).all() and (L[:, [1, 3]] <= h).all()
    # a sample of frames through the per-frame trait path and the oracle
    host = dev[:n:64].cpu().numpy()
    for k, i in enumerate(range(0, n, 64)):
        want = oracle_run(host[k], True, 0, 15)
        got = _frame_result(r1, i); got["xywh"] = want["xywh"]
        assert_same(got, want, f"frame {i}")
        assert_same(gpu_run(vision, host[k], True, 0, 15), want, f"trait path frame {i}")


@pytest.mark.parametrize("chunk,lanes", [(16, 3), (40, 2), (7, 8)])
def test_chunked_streams_equal_one_launch(vision, monkeypatch, chunk, lanes):
    """smh_b200_process_batch runs a large batch as chunks on internal streams (pixel pass of one chunk
    under the line search of another); every plane, line, flag and counter must equal the
    one-launch-per-kernel form (SMH_BATCH_SERIAL), with ragged last chunks and per-frame start_y."""
    W, H, n = 1280, 720, 101
    frames = np.concatenate([synth.frames_host(synth.SynthParams(W, H, map_open=(i % 5 != 3)), 300 + i, 1) for i in range(n)])
    dev = torch.from_numpy(frames).cuda()
    st = torch.cuda.current_stream().cuda_stream
    start_y = (np.arange(n, dtype=np.uint32) * 11) % 60
    a = vision.process_batch(dev.data_ptr(), n, W, H, start_y=start_y, max_gap=15, lsd_stats=True, minimap=True, serial=True, stream=st)
    ra = _results(vision, a, n)
    mm_a = _device_array(a.minimap, n * 16).cpu().numpy().copy(), _device_array(a.minimap_found, n).cpu().numpy().copy()
    ra = {k: v.copy() for k, v in ra.items()}
    monkeypatch.setenv("SMH_B200_BATCH_CHUNK", str(chunk))
    monkeypatch.setenv("SMH_B200_BATCH_LANES", str(lanes))
    l0 = vision.launch_count()
    b = vision.process_batch(dev.data_ptr(), n, W, H, start_y=start_y, max_gap=15, lsd_stats=True, minimap=True, stream=st)
    per_chunk = 4  # pixel pass (k_pixels2 + the seams between strips), minimap, line search
    assert vision.launch_count() - l0 == per_chunk * ((n + chunk - 1) // chunk), "expected one launch of each kernel per chunk"
    rb = _results(vision, b, n)
    mm_b = _device_array(b.minimap, n * 16).cpu().numpy(), _device_array(b.minimap_found, n).cpu().numpy()
    for k in ra:
        if k == "lines":
            for i in range(n):
                m = int(ra["n_lines"][i])
                assert np.array_equal(ra[k][i, :m].view(np.uint32), rb[k][i, :m].view(np.uint32)), (k, i)
        else:
            assert np.array_equal(ra[k], rb[k]), k
    assert np.array_equal(mm_a[1], mm_b[1]) and np.array_equal(mm_a[0], mm_b[0])
    assert (ra["map_open"] == (np.arange(n) % 5 != 3)).all()


def test_4k_frame_and_lsd_stress(vision):
    """BASELINE config 4 frame size and config 5 (LSD stress) scenes, checked against the oracle."""
    for W, H, kw, n in ((3840, 2160, synth.DENSE, 2), (2560, 1440, synth.STRESS_SEGMENTS, 1),
                        (2560, 1440, dict(n_segments=24, n_blips=400, n_labels=2, blip_band_pct=12), 1)):
        p = synth.SynthParams(W, H, **kw)
        dev = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
        synth.frames_device(vision, p, 0, n, dev.data_ptr())
        b = vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, lsd_stats=True, lsd_exhaustive=True,
                                 stream=torch.cuda.current_stream().cuda_stream)
        res = _results(vision, b, n)
        host = dev.cpu().numpy()
        for i in range(n):
            want = oracle_run(host[i], True, 0, 15)
            got = _frame_result(res, i); got["xywh"] = want["xywh"]
            assert_same(got, want, f"{W}x{H} {kw} frame {i}")
            assert tuple(int(x) for x in res["stats"][i][:2]) == tuple(want["lsd_stats"])


def test_two_handles_two_streams_in_flight(vision):
    """bench.py keeps two batches in flight: steps alternate between two handles on two CUDA streams.
    Interleaved calls must give each handle exactly the single-stream result."""
    W, H, n = 1280, 720, 64
    p = synth.SynthParams(W, H)
    dev = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(vision, p, 900, n, dev.data_ptr())
    torch.cuda.synchronize()
    ref = _results(vision, vision.process_batch(dev.data_ptr(), n, W, H, max_gap=15, stream=torch.cuda.current_stream().cuda_stream), n)
    ref = {k: v.copy() for k, v in ref.items()}
    other = B200Vision.init(0)
    try:
        hs, ss = [vision, other], [torch.cuda.Stream(), torch.cuda.Stream()]
        last = [None, None]
        for i in range(6):
            last[i % 2] = hs[i % 2].process_batch(dev.data_ptr(), n, W, H, max_gap=15, stream=ss[i % 2].cuda_stream)
        torch.cuda.synchronize()
        for k in range(2):
            got = _results(hs[k], last[k], n)
            for key in ("ui_map", "lsd_image", "ocr_out", "scales", "n_lines", "red_pixels", "map_open"):
                assert np.array_equal(got[key], ref[key]), (k, key)
            for i in range(n):
                m = int(ref["n_lines"][i])
                assert np.array_equal(got["lines"][i, :m].view(np.uint32), ref["lines"][i, :m].view(np.uint32))
    finally:
        other.close()
