"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, host-only entry points work, and compute entry points fail loudly without a
GPU (no fallback)."""
import ctypes
import re

import numpy as np
import pytest

import smh_b200  # noqa: F401
from conftest import ROOT, has_gpu
from oracle import oracle_py as O
from squad_mortar_helper_b200 import capi, synth
from squad_mortar_helper_b200.vision import B200Vision, VisionError, roi


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "smh_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(smh_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(capi.SYMBOLS), "capi.SYMBOLS out of sync with include/smh_b200.h"
    L = ctypes.CDLL(str(capi.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), f"libsmh_b200.so does not export {name}"
    capi.lib()


def test_rust_sketch_binds_every_declared_symbol():
    """rust/vision-b200/src/ffi.rs (never compiled here) names every entry point of the header, and nothing else."""
    header = (ROOT / "include" / "smh_b200.h").read_text()
    declared = set(re.findall(r"\b(smh_b200_[a-z0-9_]+)\s*\(", header))
    ffi = (ROOT / "rust" / "vision-b200" / "src" / "ffi.rs").read_text()
    bound = set(re.findall(r"pub fn (smh_b200_[a-z0-9_]+)", ffi))
    assert bound == declared, (sorted(declared - bound), sorted(bound - declared))


def test_library_carries_sm100a_code_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    deps = subprocess.run(["ldd", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    for f in (ROOT / "squad-mortar-helper_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            assert "oracle" not in f.read_text().replace("smh_oracle", "oracle").lower().replace(
                "oracle inputs", "").replace("cpu oracle (", "") or f.name in ("smh_synth.h",), f


def test_roi_matches_oracle_for_many_sizes():
    rng = np.random.default_rng(1)
    sizes = [(1024, 768), (1280, 720), (1280, 1024), (1366, 768), (1600, 900), (1600, 1024), (1920, 1080),
             (1920, 1200), (2560, 1080), (2560, 1440), (3440, 1440), (3840, 2160), (5120, 1440), (7680, 4320)]
    sizes += [(int(rng.integers(300, 6000)), int(rng.integers(200, 3000))) for _ in range(300)]
    for W, H in sizes:
        assert roi(W, H) == O.roi(W, H), (W, H)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(VisionError) as e:
        B200Vision.init(0)
    assert e.value.code == capi.SMH_E_NOGPU and "no CPU fallback" in str(e.value)


def test_synthetic_generator_is_deterministic():
    p = synth.SynthParams(1280, 720)
    a = synth.frames_host(p, 3, 2)
    b = synth.frames_host(p, 4, 1)
    assert np.array_equal(a[1], b[0]), "frame k must depend on (seed + k) only"
    assert not np.array_equal(a[0], a[1])
    assert (a[..., 3] == 255).all()
    # background never satisfies the marker predicate; every painted marker pixel does
    r = O.roi(1280, 720)
    x, y, w, h = r["map"]
    plain = synth.frames_host(synth.SynthParams(1280, 720, **synth.PLAIN), 0, 1)[0]
    tm = O.table_marker()
    rgb = plain[y:y + h, x:x + w].astype(np.uint32)
    assert not tm[(rgb[..., 2] << 16) | (rgb[..., 1] << 8) | rgb[..., 0]].any()
    rgb = a[0][y:y + h, x:x + w].astype(np.uint32)
    assert tm[(rgb[..., 2] << 16) | (rgb[..., 1] << 8) | rgb[..., 0]].sum() > 500


def test_host_widen_l8_to_rgba_matches_numpy():
    """The host half of the grayscale ui_map transfer (csrc/smh_host_widen.cpp): (L) -> (L, L, L, 255), any
    length and alignment (vector body + scalar head and tail)."""
    L = ctypes.CDLL(str(capi.LIB_PATH))
    fn = L.smh_host_widen_l8_to_rgba
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    fn.restype = None
    rng = np.random.default_rng(3)
    for n, off in [(0, 0), (1, 0), (7, 1), (8, 0), (63, 3), (4096, 0), (100003, 5), (1 << 20, 2)]:
        src = rng.integers(0, 256, n + 16, dtype=np.uint8)
        dst = np.zeros(n + 16, np.uint32)
        fn(src.ctypes.data + off, dst.ctypes.data + 4 * off, n)
        want = src[off:off + n].astype(np.uint32) * 0x010101 | 0xFF000000
        assert np.array_equal(dst[off:off + n], want)
        assert not dst[:off].any() and not dst[off + n:].any()
