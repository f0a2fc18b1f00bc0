/*
 * oracle/fast_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the prevPts provider of the LK path:
 *   feature_detection::FastDetector::detect      /root/reference/src/features.cpp:43-98
 *   utils::shi_tomasi_score                      /root/reference/src/utils.cpp:282-321
 * plus the three routines it calls in the external, un-vendored `fast` library (uzh-rpg/fast, Rosten's FAST;
 * `find_package(fast REQUIRED)`, reference CMakeLists.txt:17 -- NOT in /root/reference and not installed here):
 *   fast_corner_detect_10[_sse2]   a pixel p is a corner when >= 10 contiguous pixels of the 16-pixel Bresenham circle
 *                                  of radius 3 are all > p + b or all < p - b; x in [3, w-3), y in [3, h-3), raster order
 *   fast_corner_score_10           bisection for the largest b (bmin = b, bmax = 255) that still makes p a corner
 *   fast_nonmax_3x3                raster-order list walk; a corner survives when no 8-neighbour corner has a
 *                                  score >= its own (Rosten's Compare(X, Y) = (X >= Y))
 *
 * PARITY UNPINNED for this row: neither the reference nor the `fast` library can be built here and the reference has
 * no test or golden vector for it.  What is pinned: the circle geometry / threshold semantics of the corner test are
 * checked against cv2.FastFeatureDetector (TYPE_9_16, no NMS) by running this same code with arc length 9
 * (tests/test_oracle_fast.py).  Everything is written list-based like the reference, on purpose different from the
 * map-based CUDA formulation, so the two check each other.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_E_ARG (-1)

typedef struct { short x, y; } fast_xy;

static const int CX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int CY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* corner test at threshold b with arc length `arc` */
static int is_corner(const uint8_t* p, long stride, int b, int arc)
{
    const int cb = *p + b, c_b = *p - b;
    int bright = 0, dark = 0;
    for (int i = 0; i < 16; i++) {
        const int v = p[CY[i] * stride + CX[i]];
        if (v > cb) bright |= 1 << i;
        if (v < c_b) dark |= 1 << i;
    }
    for (int s = 0; s < 16; s++) {
        int ab = 1, ad = 1;
        for (int k = 0; k < arc; k++) {
            const int bit = 1 << ((s + k) & 15);
            ab &= (bright & bit) != 0;
            ad &= (dark & bit) != 0;
        }
        if (ab || ad) return 1;
    }
    return 0;
}

/* fast_corner_detect_<arc>: raster-order corner list; returns the count (writes at most cap entries) */
int orc_fast_detect(const uint8_t* img, int w, int h, long stride, int b, int arc, short* out_xy, int cap)
{
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++)
            if (is_corner(img + y * stride + x, stride, b, arc)) {
                if (n < cap) { out_xy[2 * n] = (short)x; out_xy[2 * n + 1] = (short)y; }
                n++;
            }
    return n;
}

/* fast_corner_score_<arc>: the bisection of Rosten's generated code */
static int corner_score(const uint8_t* p, long stride, int bstart, int arc)
{
    int bmin = bstart, bmax = 255, b = (bmax + bmin) / 2;
    for (;;) {
        if (is_corner(p, stride, b, arc)) bmin = b; else bmax = b;
        if (bmin == bmax - 1 || bmin == bmax) return bmin;
        b = (bmin + bmax) / 2;
    }
}

void orc_fast_score(const uint8_t* img, long stride, const short* xy, int n, int b, int arc, int* scores)
{
    for (int i = 0; i < n; i++) scores[i] = corner_score(img + xy[2 * i + 1] * stride + xy[2 * i], stride, b, arc);
}

/* fast_nonmax_3x3: indices of the surviving corners, raster order; returns their number */
int orc_fast_nonmax(const short* xy, const int* scores, int n, int* keep)
{
    if (n < 1) return 0;
    const fast_xy* c = (const fast_xy*)xy;
    const int last_row = c[n - 1].y;
    int* row_start = (int*)malloc(sizeof(int) * (size_t)(last_row + 1));
    for (int i = 0; i <= last_row; i++) row_start[i] = -1;
    int prev_row = -1;
    for (int i = 0; i < n; i++)
        if (c[i].y != prev_row) { row_start[c[i].y] = i; prev_row = c[i].y; }
    int point_above = 0, point_below = 0, nk = 0;
    for (int i = 0; i < n; i++) {
        const int score = scores[i];
        const fast_xy pos = c[i];
        int suppressed = 0;
        if (i > 0 && c[i - 1].x == pos.x - 1 && c[i - 1].y == pos.y && scores[i - 1] >= score) continue;
        if (i < n - 1 && c[i + 1].x == pos.x + 1 && c[i + 1].y == pos.y && scores[i + 1] >= score) continue;
        if (pos.y != 0 && row_start[pos.y - 1] != -1) {
            if (c[point_above].y < pos.y - 1) point_above = row_start[pos.y - 1];
            for (; c[point_above].y < pos.y && c[point_above].x < pos.x - 1; point_above++) {}
            for (int j = point_above; c[j].y < pos.y && c[j].x <= pos.x + 1; j++) {
                const int x = c[j].x;
                if ((x == pos.x - 1 || x == pos.x || x == pos.x + 1) && scores[j] >= score) { suppressed = 1; break; }
            }
        }
        if (!suppressed && pos.y != last_row && row_start[pos.y + 1] != -1 && point_below < n) {
            if (c[point_below].y < pos.y + 1) point_below = row_start[pos.y + 1];
            for (; point_below < n && c[point_below].y == pos.y + 1 && c[point_below].x < pos.x - 1; point_below++) {}
            for (int j = point_below; j < n && c[j].y == pos.y + 1 && c[j].x <= pos.x + 1; j++) {
                const int x = c[j].x;
                if ((x == pos.x - 1 || x == pos.x || x == pos.x + 1) && scores[j] >= score) { suppressed = 1; break; }
            }
        }
        if (!suppressed) keep[nk++] = i;
    }
    free(row_start);
    return nk;
}

/* utils::shi_tomasi_score (src/utils.cpp:282-321), fp32 accumulators, 8x8 box, central differences */
float orc_shi_tomasi(const uint8_t* img, int cols, int rows, long stride, int u, int v)
{
    float dXX = 0.0f, dYY = 0.0f, dXY = 0.0f;
    const int halfbox_size = 4, box_size = 2 * halfbox_size, box_area = box_size * box_size;
    const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
    if (x_min < 1 || x_max >= cols - 1 || y_min < 1 || y_max >= rows - 1) return 0.0f;
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t* ptr_left = img + stride * y + x_min - 1;
        const uint8_t* ptr_right = img + stride * y + x_min + 1;
        const uint8_t* ptr_top = img + stride * (y - 1) + x_min;
        const uint8_t* ptr_bottom = img + stride * (y + 1) + x_min;
        for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
            const float dx = (float)(*ptr_right - *ptr_left);
            const float dy = (float)(*ptr_bottom - *ptr_top);
            dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
        }
    }
    dXX = (float)(dXX / (2.0 * box_area));
    dYY = (float)(dYY / (2.0 * box_area));
    dXY = (float)(dXY / (2.0 * box_area));
    /* the reference writes sqrt(<float expression>); with <cmath> this is the float overload */
    const float s = sqrtf((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY));
    return (float)(0.5 * (dXX + dYY - s));
}

/* FastDetector::detect (src/features.cpp:43-98).  levels[l]: continuous (w>>l) x (h>>l) box-pyramid images.
 * occupancy: grid_cols*grid_rows bytes or NULL (all free, as after reset_grid()).  Outputs in grid-cell order:
 * out_xy (level-0 coordinates), out_level, out_score; returns the number of features. */
int orc_fast_detector_arc(const uint8_t* const* levels, int w, int h, int n_levels, int cell_size, int fast_threshold,
                          double detection_threshold, const uint8_t* occupancy, int arc, int* out_xy, int* out_level, float* out_score)
{
    if (!levels || n_levels < 1 || cell_size < 1 || arc < 9 || arc > 12) return ORC_E_ARG;
    const int gc = (int)ceil((double)w / cell_size), gr = (int)ceil((double)h / cell_size);
    const int ncell = gc * gr;
    int* cx = (int*)calloc((size_t)ncell, sizeof(int));
    int* cy = (int*)calloc((size_t)ncell, sizeof(int));
    int* cl = (int*)calloc((size_t)ncell, sizeof(int));
    float* cs = (float*)malloc(sizeof(float) * (size_t)ncell);
    for (int k = 0; k < ncell; k++) cs[k] = (float)detection_threshold;
    for (int lvl = 0; lvl < n_levels; lvl++) {
        const int lw = w >> lvl, lh = h >> lvl, scale = 1 << lvl;
        const uint8_t* img = levels[lvl];
        const int cap = lw * lh;
        short* xy = (short*)malloc(sizeof(short) * 2 * (size_t)cap);
        const int n = orc_fast_detect(img, lw, lh, lw, fast_threshold, arc, xy, cap);
        int* scores = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
        int* keep = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
        orc_fast_score(img, lw, xy, n, fast_threshold, arc, scores);
        const int nk = orc_fast_nonmax(xy, scores, n, keep);
        for (int t = 0; t < nk; t++) {
            const int x = xy[2 * keep[t]], y = xy[2 * keep[t] + 1];
            const int k = ((y * scale) / cell_size) * gc + (x * scale) / cell_size;
            if (occupancy && occupancy[k]) continue;
            const float score = orc_shi_tomasi(img, lw, lh, lw, x, y);
            if (score > cs[k]) { cx[k] = x * scale; cy[k] = y * scale; cs[k] = score; cl[k] = lvl; }
        }
        free(xy); free(scores); free(keep);
    }
    int nf = 0;
    for (int k = 0; k < ncell; k++)
        if ((double)cs[k] > detection_threshold) {
            out_xy[2 * nf] = cx[k]; out_xy[2 * nf + 1] = cy[k]; out_level[nf] = cl[k]; out_score[nf] = cs[k];
            nf++;
        }
    free(cx); free(cy); free(cl); free(cs);
    return nf;
}

/* The reference's detector: arc length 10 (fast_corner_detect_10 / fast_corner_score_10, src/features.cpp:55-72).  The arc
 * parameter above exists so that the whole detector can be pinned against OpenCV's FAST-9 (tests/test_oracle_fast.py). */
int orc_fast_detector(const uint8_t* const* levels, int w, int h, int n_levels, int cell_size, int fast_threshold,
                      double detection_threshold, const uint8_t* occupancy, int* out_xy, int* out_level, float* out_score)
{
    return orc_fast_detector_arc(levels, w, h, n_levels, cell_size, fast_threshold, detection_threshold, occupancy, 10, out_xy,
                                 out_level, out_score);
}
