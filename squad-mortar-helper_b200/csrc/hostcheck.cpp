// hostcheck.cpp — CPU-only build of the shared per-pixel arithmetic (smh_pixel.h) so that the
// exhaustive 2^24-colour checks (exact predicate, integer fast-reject, luma, OCR classes) can
// run without a GPU.  Test helper only; never loaded by the product path.
// Build: g++ -O2 -ffp-contract=off -fPIC -shared hostcheck.cpp -o ../libsmh_hostcheck.so
#include "smh_pixel.h"

extern "C" {
// which: 0 marker (prefilter && exact), 1 luma8, 2 button-red, 3 ocr class, 4 exact only, 5 prefilter only
void smh_hostcheck_table(uint32_t which, uint8_t* out) {
    for (uint32_t c = 0; c < (1u << 24); c++) {
        const uint32_t r = c >> 16, g = (c >> 8) & 0xFFu, b = c & 0xFFu;
        uint8_t v;
        switch (which) {
            case 0: v = smh_is_marker(r, g, b); break;
            case 1: v = (uint8_t)smh_luma8(r, g, b); break;
            case 2: v = smh_button_red(r, g, b); break;
            case 3: v = (uint8_t)((smh_ocr_core(r, g, b) ? 1 : 0) | ((!smh_ocr_core(r, g, b) && smh_ocr_edge(r, g, b)) ? 2 : 0)); break;
            case 4: v = smh_is_marker_exact(r, g, b); break;
            default: v = smh_marker_prefilter(r, g, b); break;
        }
        out[c] = v;
    }
}
}
