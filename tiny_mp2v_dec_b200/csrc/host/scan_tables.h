// Scan-order tables of ISO/IEC 13818-2 7.3 (Figures 7-2, 7-3), generated once at start-up.
//
// Roles in the reference (src/core/scan_c.cpp:4-57): g_shuffle[alt][i] = raster index of scan
// position i, g_scan[alt] = its inverse, g_scan_trans[alt][i] = the TRANSPOSED raster index, which
// is where parse_block stores coefficient i (mb_decoder.cpp:141) because the SSE2 IDCT runs its
// first pass down the columns of that layout.  The device kernel uses scan_trans; the host uses
// scan0/shuffle to turn transmitted quantiser matrices into scan-position-indexed W (decoder.cpp:169-191).
#pragma once
#include <cstdint>

namespace mp2v {

struct scan_tables_t {
    uint8_t shuffle[2][64];     // scan position -> raster index v*8+u
    uint8_t scan[2][64];        // raster index -> scan position
    uint8_t scan_trans[2][64];  // scan position -> transposed raster index u*8+v

    scan_tables_t() {
        // Figure 7-2: zig-zag = walk the 15 anti-diagonals, alternating direction
        int i = 0;
        for (int d = 0; d < 15; d++) {
            const int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7;
            for (int k = 0; k <= hi - lo; k++) {
                const int v = (d & 1) ? lo + k : hi - k;
                shuffle[0][i++] = (uint8_t)(v * 8 + (d - v));
            }
        }
        // Figure 7-3: alternate scan.  Column pairs are traversed in vertical runs of 4 / 2 / 8;
        // held as the scan position of every raster cell [v][u].
        static const uint8_t alt[8][8] = {
            { 0,  4,  6, 20, 22, 36, 38, 52}, { 1,  5,  7, 21, 23, 37, 39, 53},
            { 2,  8, 19, 24, 34, 40, 50, 54}, { 3,  9, 18, 25, 35, 41, 51, 55},
            {10, 17, 26, 30, 42, 46, 56, 60}, {11, 16, 27, 31, 43, 47, 57, 61},
            {12, 15, 28, 32, 44, 48, 58, 62}, {13, 14, 29, 33, 45, 49, 59, 63}};
        for (int v = 0; v < 8; v++)
            for (int u = 0; u < 8; u++) shuffle[1][alt[v][u]] = (uint8_t)(v * 8 + u);
        for (int a = 0; a < 2; a++)
            for (int p = 0; p < 64; p++) {
                const int r = shuffle[a][p];
                scan[a][r] = (uint8_t)p;
                scan_trans[a][p] = (uint8_t)(((r & 7) << 3) | (r >> 3));
            }
    }
};

inline const scan_tables_t& scan_tables() {
    static const scan_tables_t t;
    return t;
}

// quantiser_matrices of mp2v_picture_c::init (decoder.cpp:169-191): tx = matrix as transmitted
// (zig-zag order); W[i] = tx[ scan[0][ shuffle[alt][i] ] ], indexed by scan position i.
inline void build_scan_indexed_matrix(const uint8_t tx[64], int alternate_scan, uint8_t W[64]) {
    const scan_tables_t& t = scan_tables();
    for (int i = 0; i < 64; i++) W[i] = tx[t.scan[0][t.shuffle[alternate_scan ? 1 : 0][i]]];
}

}  // namespace mp2v
