/* smh_consts.h — algorithm constants of the map-analysis vision path.
 *
 * Single source for the C oracle, the CUDA kernels and the host code.  The
 * values restate /root/reference/vision-common/src/consts/consts.toml:1-67
 * (thresholds) and vision-common/src/consts/mod.rs:7-19 (relative ROI bounds);
 * tests/test_consts.py re-parses that TOML when the reference tree is mounted
 * and fails on any drift.
 */
#ifndef SMH_CONSTS_H
#define SMH_CONSTS_H

/* "Close deployment" button test — consts.toml:1-11 */
#define SMH_BTN_COLOR_R 217
#define SMH_BTN_COLOR_G 67
#define SMH_BTN_COLOR_B 49
#define SMH_BTN_TOLERANCE 25
#define SMH_BTN_RED_RATIO_THRESHOLD 0.65f

/* OCR preprocess — consts.toml:13-31 */
#define SMH_OCR_BRIGHT 200          /* OCR_PREPROCESS_BRIGHTNESS_THRESHOLD */
#define SMH_OCR_MONO 3              /* OCR_PREPROCESS_MONOCHROMATICY_THRESHOLD */
#define SMH_OCR_BRIGHT_EDGE 130     /* OCR_PREPROCESS_BRIGHTNESS_EDGE_THRESHOLD */
#define SMH_OCR_MONO_EDGE 48        /* OCR_PREPROCESS_SIMILARITY_EDGE_THRESHOLD */
#define SMH_OCR_RADIUS 3            /* OCR_PREPROCESS_DILATE_RADIUS */

/* Squad-leader marker colours, HSV (h 0..360, s 0..100, v 0..100) — consts.toml:33-43 */
#define SMH_ALPHA_H 105
#define SMH_ALPHA_S 100
#define SMH_ALPHA_V 100
#define SMH_BRAVO_H 285
#define SMH_BRAVO_S 46
#define SMH_BRAVO_V 85
#define SMH_CHARLIE_H 158
#define SMH_CHARLIE_S 60
#define SMH_CHARLIE_V 91

/* Marker colour windows — consts.toml:45-63 */
#define SMH_HUE_TOL 15
#define SMH_SAT_TOL 15
#define SMH_VIB_TOL 15
#define SMH_MIN_SAT 35
#define SMH_ARC_SAT 50              /* FIND_MARKER_PLAYER_DIR_ARC_SAT */

#define SMH_MAP_MARKER_POI_LOCATION 0.70f /* consts.toml:65-67 (disabled template stage) */

/* Relative ROI bounds, all fractions of the screen HEIGHT — consts/mod.rs:7-19 */
#define SMH_MAP_X_FRAC 0.018522135
#define SMH_MAP_Y_BOTTOM_FRAC 0.07421875
#define SMH_MAP_W_FRAC 0.864930556
#define SMH_MAP_H_FRAC 0.761078559
#define SMH_BTN_X_RIGHT_FRAC 0.0078125
#define SMH_BTN_Y_BOTTOM_FRAC 0.0078125
#define SMH_BTN_W_FRAC 0.236132813
#define SMH_BTN_H_FRAC 0.038205295

/* Line search — vision-common/src/lsd.rs:9,87,95,99 and vision-cpu/src/lib.rs:434 */
#define SMH_LSD_RAYS 3600
#define SMH_LSD_MAX_LINES 32
#define SMH_LSD_CENTRE_MAX_DIST 5.0f
#define SMH_LSD_SUPPRESS_DIST_SQ 50.0f
#define SMH_LSD_MIN_LEN_SQ 2500.0f

/* Scale bars — src/vision/mpx_ratio.rs:5-6,12 */
#define SMH_SCALE_MIN_WIDTH 10u
#define SMH_SCALE_MIN_VBAR_HEIGHT 4u
/* Minimap edge search — src/vision/find_minimap.rs:66 (f32); `max as f32 / 765.0 <= 0.01` holds exactly
 * for max <= 7 (tests/test_next_rows_cpu.py checks all 766 values) */
#define SMH_MINIMAP_EDGINESS_THRESHOLD 0.01f
#define SMH_MINIMAP_EDGINESS_MAX_SUM 7u

/* image 0.23.14 rgb_to_luma weights (f32) */
#define SMH_LUMA_R 0.2126f
#define SMH_LUMA_G 0.7152f
#define SMH_LUMA_B 0.0722f

#endif /* SMH_CONSTS_H */
