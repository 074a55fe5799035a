"""include/smh_consts.h must agree with the reference's single source of constants."""
import re
from pathlib import Path

import pytest

from conftest import REFERENCE, ROOT

HEADER = (ROOT / "include" / "smh_consts.h").read_text()


def macro(name):
    m = re.search(rf"#define\s+{name}\s+([^\s/]+)", HEADER)
    assert m, name
    return float(m.group(1).rstrip("fu"))


MAP = {
    "CLOSE_DEPLOYMENT_BUTTON_COLOR": ["SMH_BTN_COLOR_R", "SMH_BTN_COLOR_G", "SMH_BTN_COLOR_B"],
    "CLOSE_DEPLOYMENT_BUTTON_TOLERANCE": "SMH_BTN_TOLERANCE",
    "CLOSE_DEPLOYMENT_BUTTON_RED_PIXEL_THRESHOLD": "SMH_BTN_RED_RATIO_THRESHOLD",
    "OCR_PREPROCESS_BRIGHTNESS_THRESHOLD": "SMH_OCR_BRIGHT",
    "OCR_PREPROCESS_MONOCHROMATICY_THRESHOLD": "SMH_OCR_MONO",
    "OCR_PREPROCESS_BRIGHTNESS_EDGE_THRESHOLD": "SMH_OCR_BRIGHT_EDGE",
    "OCR_PREPROCESS_SIMILARITY_EDGE_THRESHOLD": "SMH_OCR_MONO_EDGE",
    "OCR_PREPROCESS_DILATE_RADIUS": "SMH_OCR_RADIUS",
    "ALPHA_MARKER_COLOR_HSV": ["SMH_ALPHA_H", "SMH_ALPHA_S", "SMH_ALPHA_V"],
    "BRAVO_MARKER_COLOR_HSV": ["SMH_BRAVO_H", "SMH_BRAVO_S", "SMH_BRAVO_V"],
    "CHARLIE_MARKER_COLOR_HSV": ["SMH_CHARLIE_H", "SMH_CHARLIE_S", "SMH_CHARLIE_V"],
    "FIND_MARKER_HSV_HUE_TOLERANCE": "SMH_HUE_TOL",
    "FIND_MARKER_HSV_SAT_TOLERANCE": "SMH_SAT_TOL",
    "FIND_MARKER_HSV_VIB_TOLERANCE": "SMH_VIB_TOL",
    "FIND_MARKER_HSV_MIN_SAT": "SMH_MIN_SAT",
    "FIND_MARKER_PLAYER_DIR_ARC_SAT": "SMH_ARC_SAT",
    "MAP_MARKER_POI_LOCATION": "SMH_MAP_MARKER_POI_LOCATION",
}


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not mounted")
def test_header_matches_reference_toml():
    import tomllib
    toml = tomllib.loads((REFERENCE / "vision-common/src/consts/consts.toml").read_text())
    assert set(toml) == set(MAP), "a constant was added to / removed from consts.toml"
    for key, names in MAP.items():
        val = toml[key]["value"]
        if isinstance(names, list):
            assert [macro(n) for n in names] == [float(v) for v in val], key
        else:
            assert macro(names) == pytest.approx(float(val), abs=0), key


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not mounted")
def test_roi_fractions_match_reference_source():
    src = (REFERENCE / "vision-common/src/consts/mod.rs").read_text()
    for name in ("SMH_MAP_X_FRAC", "SMH_MAP_Y_BOTTOM_FRAC", "SMH_MAP_W_FRAC", "SMH_MAP_H_FRAC",
                 "SMH_BTN_X_RIGHT_FRAC", "SMH_BTN_W_FRAC", "SMH_BTN_H_FRAC"):
        lit = re.search(rf"#define\s+{name}\s+([0-9.]+)", HEADER).group(1)
        assert f"ScreenH({lit})" in src, (name, lit)


def test_survey_roi_table():
    # SURVEY.md section 8 table, validated against the bundled screenshots at survey time
    from squad_mortar_helper_b200.vision import roi
    import smh_b200  # noqa: F401
    table = {
        (1024, 768): ((650, 126, 360, 585), (830, 418, 180, 292), (837, 733, 181, 29)),
        (1280, 1024): ((867, 169, 394, 779), (1064, 558, 197, 389), (1030, 977, 242, 39)),
        (1600, 1024): ((867, 169, 714, 779), (1224, 558, 357, 389), (1350, 977, 242, 39)),
        (1920, 1080): ((914, 178, 986, 822), (1407, 589, 493, 411), (1657, 1031, 255, 41)),
        (2560, 1440): ((1219, 237, 1314, 1096), (1876, 785, 657, 548), (2209, 1374, 340, 55)),
        (3840, 2160): ((1828, 356, 1972, 1644), (2814, 1178, 986, 822), (3313, 2060, 510, 83)),
    }
    for (W, H), (m, q, b) in table.items():
        r = roi(W, H)
        assert r["map"] == m and r["brq"] == q and r["button"] == b, (W, H, r)
