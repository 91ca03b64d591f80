// Host-only property check of the index helpers shared by the kernels (medaka_b200/csrc/common.cuh): compiled and run by
// tests/test_library.py::test_layout_helpers_host.  Exit code 0 = all properties hold.
#include <cstdio>
#include <vector>
#include "../../medaka_b200/csrc/common.cuh"

using namespace mdk;

int main() {
    // tiled_row: a bijection of (window, t) onto [0, tiled_rows) restricted to real windows; 16 windows of a tile adjacent per t
    const int64_t B = 37, T = 11;
    const int64_t rows = tiled_rows(B, T);
    if (rows != 48 * T) { printf("tiled_rows %lld\n", (long long)rows); return 1; }
    std::vector<int> seen(rows, 0);
    for (int64_t w = 0; w < B; ++w)
        for (int64_t t = 0; t < T; ++t) {
            const int64_t r = tiled_row(w, t, T);
            if (r < 0 || r >= rows || seen[r]++) { printf("tiled_row collision at w=%lld t=%lld\n", (long long)w, (long long)t); return 2; }
            if (r / WT != (w / WT) * T + t || r % WT != w % WT) return 3;
        }
    // gi_quad_index: a bijection of (row, col) onto [0, rows * 768); one tile-step = one contiguous block of GI_TS_FLOATS;
    // within it [blk 6][quad 4][j 128][w4 4], so a (tile-step, direction) is a contiguous half block
    std::vector<char> hit((size_t)rows * GI_COLS, 0);
    for (int64_t r = 0; r < rows; ++r)
        for (int c = 0; c < GI_COLS; ++c) {
            const int64_t i = gi_quad_index(r, c);
            if (i < 0 || i >= rows * GI_COLS || hit[i]++) { printf("gi_quad_index collision r=%lld c=%d\n", (long long)r, c); return 4; }
            const int64_t ts = r / WT;
            if (i / GI_TS_FLOATS != ts) return 5;
            const int64_t in = i % GI_TS_FLOATS;
            const int blk = c / H, j = c % H, w = (int)(r % WT);
            if (in != (((int64_t)blk * 4 + w / 4) * H + j) * 4 + w % 4) return 6;
            if ((in / (GI_TS_FLOATS / 2)) != (c / G3)) return 7;      // direction halves
        }
    if (gate_scale(0) != GATE_SCALE_RZ || gate_scale(1) != GATE_SCALE_RZ || gate_scale(2) != GATE_SCALE_N) return 8;
    if (PLOG_TS_FLOATS != 80 || GI_TS_FLOATS != 16 * GI_COLS) return 9;
    printf("layout helpers OK\n");
    return 0;
}
