"""The product's shared per-pixel arithmetic (csrc/smh_pixel.h), compiled for the host, against
the oracle over ALL 2^24 colours: exact predicate, integer fast-reject, luma, OCR classes."""
import ctypes

import numpy as np

import smh_b200  # noqa: F401
from oracle import oracle_py as O
from squad_mortar_helper_b200 import build


def _tab(which):
    L = ctypes.CDLL(str(build.build_hostcheck()))
    L.smh_hostcheck_table.argtypes = [ctypes.c_uint32, ctypes.c_void_p]
    out = np.empty(1 << 24, np.uint8)
    L.smh_hostcheck_table(which, out.ctypes.data)
    return out


def test_marker_predicate_and_prefilter_exhaustive():
    O.set_threads(8)
    want = O.table_marker()
    assert np.array_equal(_tab(4), want), "exact predicate differs from the oracle"
    pre = _tab(5)
    assert not (want & (1 - pre)).any(), "the integer fast-reject drops a marker colour"
    assert np.array_equal(_tab(0), want)


def test_luma_exhaustive():
    assert np.array_equal(_tab(1), O.table_luma())


def test_ocr_classes_and_button_exhaustive():
    c = np.arange(1 << 24, dtype=np.int64)
    r, g, b = (c >> 16) & 255, (c >> 8) & 255, c & 255
    mono = 2 * (np.abs(r - g) + np.abs(r - b) + np.abs(g - b))   # sum over the 9 ordered pairs
    mn = np.minimum(r, np.minimum(g, b))
    core = (mono <= 3) & (mn >= 200)
    edge = ~core & (mono <= 48) & (mn >= 130)
    assert np.array_equal(_tab(3), (core * 1 + edge * 2).astype(np.uint8))
    red = (np.abs(217 - r) <= 25) & (np.abs(67 - g) <= 25) & (np.abs(49 - b) <= 25)
    assert np.array_equal(_tab(2), red.astype(np.uint8))
