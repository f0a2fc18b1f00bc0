/*
 * oracle/lk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99) of the arithmetic on the 3DR pyramidal-LK hot path.  It exists only
 * so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path
 * against something that does not run on the GPU.  Nothing under 3dr_b200/ may link, import or call it.
 *
 * What it follows:
 *   - box pyramid:   /root/reference/src/utils.cpp:324-350 (halfSampleSSE2), 382-419 (reduce_to_half,
 *                    dispatch + scalar pointer walk), 421-430 (create_img_pyramid).
 *   - LK tracking:   the reference calls cv::calcOpticalFlowPyrLK at
 *                    /root/reference/src/initialization.cpp:608-613.  That function lives in OpenCV
 *                    (un-vendored, un-pinned `find_package(OpenCV REQUIRED)`, CMakeLists.txt:8,12;
 *                    3.x on the author's machine) and is NOT in /root/reference.  Its published
 *                    algorithm (modules/video/src/lkpyramid.cpp: calcOpticalFlowPyrLK,
 *                    buildOpticalFlowPyramid, calcScharrDeriv, LKTrackerInvoker; modules/imgproc/src/
 *                    pyramids.cpp: pyrDown) is restated here from SURVEY.md Appendix A.
 *
 * Pinning: the restatement is checked against opencv-python-headless 4.13.0 (`cv2`, the same C++
 * code path the reference calls) by tests/test_oracle.py when cv2 is importable, and against
 * the committed fixtures tests/golden/ (generated from cv2 by tests/golden/make_golden.py) otherwise.
 * The box pyramid is pinned to the reference itself: oracle/_ref (src/utils.cpp:282-430 compiled unmodified by
 * oracle/build_ref.sh) == this file bit for bit on every shape tests/test_oracle_vs_ref.py tries.
 * Pyramids and Scharr derivatives are bit-exact vs cv2; LK positions agree to a few 1e-3 px because
 * x86 OpenCV accumulates the 2x2 system in fp32 SIMD lanes while this file accumulates the same
 * integer products exactly (int64) and converts once (the GPU does exactly the same, so GPU == oracle
 * is expected to be bit-exact).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_E_ARG (-1)
#define ORC_E_UNSUPPORTED (-2)

#define ORC_BOX_AUTO_X86 0 /* reduce_to_half on x86: SSE2 rounding iff cols % 16 == 0 (utils.cpp:386-392) */
#define ORC_BOX_TRUNC 1    /* scalar / NEON arithmetic: (a+b+c+d)/4 (utils.cpp:411, 365-367)          */
#define ORC_BOX_SSE2 2     /* avg_epu8 then avg_epu16, both round up (utils.cpp:337-341)              */

/* ------------------------------------------------------------------------------------------------
 * Box pyramid (reference: src/utils.cpp)
 * ---------------------------------------------------------------------------------------------- */

/* Would the scalar pointer walk of reduce_to_half (utils.cpp:401-418) stay inside its buffers?
 * Returns the number of output rows it writes, or -1 when it reads past the input. */
static long box_walk_rows(int w, int h, long stride, int* reads_ok)
{
    long out_w = w / 2, top = 0, bottom = stride, end = stride * (long)h, rows = 0;
    *reads_ok = 1;
    while (bottom < end) {
        if (out_w > 0 && bottom + 2 * out_w - 1 >= end) *reads_ok = 0;
        top += 2 * out_w + stride;
        bottom += 2 * out_w + stride;
        rows++;
    }
    (void)top;
    return rows;
}

/* One level of utils::reduce_to_half.  `in` is h rows of `stride` bytes; `out` is (h/2) x (w/2)
 * continuous.  mode selects the rounding / code path as on the reference's build targets. */
int orc_box_half(const uint8_t* in, int w, int h, long stride, uint8_t* out, int mode)
{
    if (!in || !out || w < 2 || h < 2 || stride < w) return ORC_E_ARG;
    const int out_w = w / 2, out_h = h / 2;
    int sse2 = (mode == ORC_BOX_SSE2) || (mode == ORC_BOX_AUTO_X86 && (w % 16) == 0);
    if (sse2) {
        /* halfSampleSSE2(in, out, w, h): assumes a continuous buffer (advances by w, utils.cpp:346-347) */
        if ((w % 16) != 0) return ORC_E_ARG;
        if (stride != w) return ORC_E_UNSUPPORTED;
        for (int i = 0; i < out_h; i++) {
            const uint8_t* r0 = in + (long)(2 * i) * w;
            const uint8_t* r1 = r0 + w;
            for (int j = 0; j < out_w; j++) {
                unsigned v0 = (r0[2 * j] + r1[2 * j] + 1u) >> 1;         /* _mm_avg_epu8 (vertical)    */
                unsigned v1 = (r0[2 * j + 1] + r1[2 * j + 1] + 1u) >> 1;
                out[(long)i * out_w + j] = (uint8_t)((v0 + v1 + 1u) >> 1); /* _mm_avg_epu16 (horizontal) */
            }
        }
        return ORC_OK;
    }
    /* scalar walk, utils.cpp:401-418: `top` advances by 2 per output pixel and by `stride` per row, so
     * for odd w each output row starts one byte earlier than a plain 2x2 box filter would. */
    int reads_ok;
    long rows = box_walk_rows(w, h, stride, &reads_ok);
    if (rows > out_h || !reads_ok) return ORC_E_UNSUPPORTED; /* the reference overruns its buffers here */
    for (long i = 0; i < rows; i++) {
        long t = i * (2L * out_w + stride);
        for (int j = 0; j < out_w; j++, t += 2)
            out[i * out_w + j] =
                (uint8_t)(((unsigned)in[t] + in[t + 1] + in[t + stride] + in[t + stride + 1]) / 4);
    }
    return ORC_OK;
}

/* utils::create_img_pyramid (utils.cpp:421-430): level 0 is the caller's image (not copied here);
 * out_levels[l-1] receives level l, continuous (w>>l) x (h>>l), for l = 1..n_levels-1. */
int orc_box_pyramid(const uint8_t* img, int w, int h, long stride, int n_levels,
                    uint8_t* const* out_levels, int mode)
{
    const uint8_t* cur = img;
    long cur_stride = stride;
    for (int l = 1; l < n_levels; l++) {
        int rc = orc_box_half(cur, w, h, cur_stride, out_levels[l - 1], mode);
        if (rc != ORC_OK) return rc;
        w /= 2; h /= 2;
        cur = out_levels[l - 1];
        cur_stride = w;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * OpenCV-style pyramid pieces
 * ---------------------------------------------------------------------------------------------- */

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
static inline int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p; else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* cv::pyrDown for CV_8UC1: separable [1 4 6 4 1], un-normalised horizontal pass, vertical pass,
 * (sum + 128) >> 8, REFLECT_101 on source indices.  dst is ((w+1)/2) x ((h+1)/2) with dst_stride. */
int orc_pyrdown(const uint8_t* src, int w, int h, long src_stride, uint8_t* dst, long dst_stride)
{
    if (!src || !dst || w < 1 || h < 1) return ORC_E_ARG;
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* rows = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    if (!rows) return ORC_E_ARG;
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const uint8_t* s = src + (long)reflect101(2 * y - 2 + k, h) * src_stride;
            int* r = rows + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                r[x] = s[reflect101(2 * x - 2, w)] + 4 * s[reflect101(2 * x - 1, w)] + 6 * s[reflect101(2 * x, w)] +
                       4 * s[reflect101(2 * x + 1, w)] + s[reflect101(2 * x + 2, w)];
            }
        }
        for (int x = 0; x < dw; x++) {
            int v = rows[x] + 4 * rows[dw + x] + 6 * rows[2 * dw + x] + 4 * rows[3 * dw + x] + rows[4 * dw + x];
            dst[(long)y * dst_stride + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rows);
    return ORC_OK;
}

/* calcScharrDeriv for CV_8UC1 -> interleaved int16 (Ix, Iy), un-normalised, REFLECT_101 at the edge.
 * dst has dst_stride int16 elements per row (>= 2*w). */
int orc_scharr(const uint8_t* src, int w, int h, long src_stride, int16_t* dst, long dst_stride)
{
    if (!src || !dst || w < 1 || h < 1) return ORC_E_ARG;
    int* t0 = (int*)malloc(sizeof(int) * (size_t)(w + 2) * 2);
    if (!t0) return ORC_E_ARG;
    int* t1 = t0 + (w + 2);
    for (int y = 0; y < h; y++) {
        const uint8_t* s0 = src + (long)reflect101(y - 1, h) * src_stride;
        const uint8_t* s1 = src + (long)y * src_stride;
        const uint8_t* s2 = src + (long)reflect101(y + 1, h) * src_stride;
        for (int x = -1; x <= w; x++) {
            int xs = reflect101(x, w);
            t0[x + 1] = 3 * (s0[xs] + s2[xs]) + 10 * s1[xs];
            t1[x + 1] = s2[xs] - s0[xs];
        }
        int16_t* d = dst + (long)y * dst_stride;
        for (int x = 0; x < w; x++) {
            d[2 * x] = (int16_t)(t0[x + 2] - t0[x]);
            d[2 * x + 1] = (int16_t)(3 * (t1[x] + t1[x + 2]) + 10 * t1[x + 1]);
        }
    }
    free(t0);
    return ORC_OK;
}

/* Level sizes as buildOpticalFlowPyramid produces them, with its early stop: after level l is made,
 * if the next level would have w <= win_w or h <= win_h, maxLevel becomes l.  Returns the effective
 * maxLevel; ws/hs (max_level+1 entries) receive the level sizes. */
int orc_lk_level_sizes(int w, int h, int win_w, int win_h, int max_level, int* ws, int* hs)
{
    int level = 0;
    for (;; level++) {
        ws[level] = w; hs[level] = h;
        if (level == max_level) break;
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        w = nw; h = nh;
    }
    return level;
}

/* Total bytes needed for orc_build_lk_pyramid's img / deriv buffers. */
long orc_lk_pyramid_bytes(int w, int h, int win_w, int win_h, int max_level, long* deriv_elems)
{
    int ws[32], hs[32];
    if (max_level > 31) max_level = 31;
    int ml = orc_lk_level_sizes(w, h, win_w, win_h, max_level, ws, hs);
    long b = 0, d = 0;
    for (int l = 0; l <= ml; l++) { b += (long)ws[l] * hs[l]; d += (long)ws[l] * hs[l] * 2; }
    if (deriv_elems) *deriv_elems = d;
    return b;
}

/* Builds the Gaussian pyramid (levels packed back to back, each continuous) and, when deriv != NULL,
 * the Scharr derivatives of every level (packed likewise, 2 int16 per pixel).  Returns effective maxLevel. */
int orc_build_lk_pyramid(const uint8_t* img, int w, int h, long stride, int win_w, int win_h, int max_level,
                         uint8_t* levels, int16_t* deriv)
{
    int ws[32], hs[32];
    if (max_level > 31) max_level = 31;
    int ml = orc_lk_level_sizes(w, h, win_w, win_h, max_level, ws, hs);
    uint8_t* cur = levels;
    for (int y = 0; y < h; y++) memcpy(cur + (long)y * w, img + (long)y * stride, (size_t)w);
    for (int l = 1; l <= ml; l++) {
        uint8_t* nxt = cur + (long)ws[l - 1] * hs[l - 1];
        orc_pyrdown(cur, ws[l - 1], hs[l - 1], ws[l - 1], nxt, ws[l]);
        cur = nxt;
    }
    if (deriv) {
        const uint8_t* s = levels;
        int16_t* d = deriv;
        for (int l = 0; l <= ml; l++) {
            orc_scharr(s, ws[l], hs[l], ws[l], d, 2L * ws[l]);
            s += (long)ws[l] * hs[l];
            d += 2L * ws[l] * hs[l];
        }
    }
    return ml;
}

/* ------------------------------------------------------------------------------------------------
 * LKTrackerInvoker
 * ---------------------------------------------------------------------------------------------- */

#define W_BITS 14
#define ORC_TERM_COUNT 1
#define ORC_TERM_EPS 2
#define ORC_USE_INITIAL_FLOW 4
#define ORC_GET_MIN_EIGENVALS 8

typedef struct {
    const uint8_t* img; /* level image, continuous */
    const int16_t* der; /* level derivative (prev only) */
    int w, h;
} orc_level;

static inline int img_at(const orc_level* L, int x, int y) /* image padded with REFLECT_101 */
{
    return L->img[(long)reflect101(y, L->h) * L->w + reflect101(x, L->w)];
}
static inline int der_at(const orc_level* L, int x, int y, int c) /* derivative padded with zeros */
{
    if ((unsigned)x >= (unsigned)L->w || (unsigned)y >= (unsigned)L->h) return 0;
    return L->der[((long)y * L->w + x) * 2 + c];
}
static inline int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }
static inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_round(float v) { return (int)lrintf(v); } /* round-half-even, as _mm_cvtss_si32 */

static inline void bilinear_weights(float a, float b, int* iw00, int* iw01, int* iw10, int* iw11)
{
    *iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
    *iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
    *iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
    *iw11 = (1 << W_BITS) - *iw00 - *iw01 - *iw10;
}

/* Per-point trace, optional: iterations executed per level (for the algorithmic-bytes model) and a
 * per-level outcome code. trace_iters[k*32 + l], trace_code[k*32 + l]:
 * 0 = level not processed, 1 = prev window out of frame, 2 = rejected (minEig / D), 3 = iterated. */
typedef struct {
    int32_t* iters;
    uint8_t* code;
    float* pos; /* nextPts after the level, [k*32+l][2] */
} orc_trace;

static void track_point(const orc_level* P, const orc_level* N, int max_level, int k, const float* prev_pts,
                        float* next_pts, uint8_t* status, float* err, int win_w, int win_h, int max_count,
                        double eps2, int flags, double min_eig_thr, int16_t* Ibuf, int16_t* dbuf, orc_trace* tr)
{
    const float hwx = (win_w - 1) * 0.5f, hwy = (win_h - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    for (int level = max_level; level >= 0; level--) {
        const orc_level* I = &P[level];
        const orc_level* J = &N[level];
        const float sc = (float)(1. / (1 << level));
        float px = prev_pts[2 * k] * sc, py = prev_pts[2 * k + 1] * sc;
        float nx, ny;
        if (level == max_level) {
            if (flags & ORC_USE_INITIAL_FLOW) { nx = next_pts[2 * k] * sc; ny = next_pts[2 * k + 1] * sc; }
            else { nx = px; ny = py; }
        } else { nx = next_pts[2 * k] * 2.f; ny = next_pts[2 * k + 1] * 2.f; }
        next_pts[2 * k] = nx; next_pts[2 * k + 1] = ny;
        if (tr) { tr->iters[k * 32 + level] = 0; tr->code[k * 32 + level] = 1; }

        px -= hwx; py -= hwy;
        int ipx = cv_floor(px), ipy = cv_floor(py);
        if (ipx < -win_w || ipx >= I->w || ipy < -win_h || ipy >= I->h) {
            if (level == 0) { status[k] = 0; if (err) err[k] = 0; }
            if (tr) { tr->pos[(k * 32 + level) * 2] = nx; tr->pos[(k * 32 + level) * 2 + 1] = ny; }
            continue;
        }
        float a = px - ipx, b = py - ipy;
        int iw00, iw01, iw10, iw11;
        bilinear_weights(a, b, &iw00, &iw01, &iw10, &iw11);
        int64_t iA11 = 0, iA12 = 0, iA22 = 0;
        for (int y = 0; y < win_h; y++)
            for (int x = 0; x < win_w; x++) {
                int X = ipx + x, Y = ipy + y;
                int ival = descale(img_at(I, X, Y) * iw00 + img_at(I, X + 1, Y) * iw01 + img_at(I, X, Y + 1) * iw10 +
                                   img_at(I, X + 1, Y + 1) * iw11, W_BITS - 5);
                int ixval = descale(der_at(I, X, Y, 0) * iw00 + der_at(I, X + 1, Y, 0) * iw01 +
                                    der_at(I, X, Y + 1, 0) * iw10 + der_at(I, X + 1, Y + 1, 0) * iw11, W_BITS);
                int iyval = descale(der_at(I, X, Y, 1) * iw00 + der_at(I, X + 1, Y, 1) * iw01 +
                                    der_at(I, X, Y + 1, 1) * iw10 + der_at(I, X + 1, Y + 1, 1) * iw11, W_BITS);
                Ibuf[y * win_w + x] = (int16_t)ival;
                dbuf[(y * win_w + x) * 2] = (int16_t)ixval;
                dbuf[(y * win_w + x) * 2 + 1] = (int16_t)iyval;
                iA11 += ixval * ixval; iA12 += ixval * iyval; iA22 += iyval * iyval;
            }
        float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win_w * win_h);
        if (err && (flags & ORC_GET_MIN_EIGENVALS)) err[k] = minEig;
        /* LKTrackerInvoker keeps minEigThreshold as a float member: the test is float < float */
        if (minEig < (float)min_eig_thr || D < FLT_EPSILON) {
            if (level == 0) status[k] = 0;
            if (tr) { tr->code[k * 32 + level] = 2; tr->pos[(k * 32 + level) * 2] = nx; tr->pos[(k * 32 + level) * 2 + 1] = ny; }
            continue;
        }
        D = 1.f / D;
        nx -= hwx; ny -= hwy;
        float pdx = 0.f, pdy = 0.f;
        int j;
        for (j = 0; j < max_count; j++) {
            int inx = cv_floor(nx), iny = cv_floor(ny);
            if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) {
                if (level == 0) status[k] = 0;
                break;
            }
            a = nx - inx; b = ny - iny;
            bilinear_weights(a, b, &iw00, &iw01, &iw10, &iw11);
            int64_t ib1 = 0, ib2 = 0;
            for (int y = 0; y < win_h; y++)
                for (int x = 0; x < win_w; x++) {
                    int X = inx + x, Y = iny + y;
                    int diff = descale(img_at(J, X, Y) * iw00 + img_at(J, X + 1, Y) * iw01 + img_at(J, X, Y + 1) * iw10 +
                                       img_at(J, X + 1, Y + 1) * iw11, W_BITS - 5) - Ibuf[y * win_w + x];
                    ib1 += diff * dbuf[(y * win_w + x) * 2];
                    ib2 += diff * dbuf[(y * win_w + x) * 2 + 1];
                }
            float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
            float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
            nx += dx; ny += dy;
            next_pts[2 * k] = nx + hwx; next_pts[2 * k + 1] = ny + hwy;
            if (tr) tr->iters[k * 32 + level] = j + 1;
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) {
                next_pts[2 * k] -= dx * 0.5f; next_pts[2 * k + 1] -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (tr) {
            tr->code[k * 32 + level] = 3;
            tr->pos[(k * 32 + level) * 2] = next_pts[2 * k]; tr->pos[(k * 32 + level) * 2 + 1] = next_pts[2 * k + 1];
        }
        if (status[k] && err && level == 0 && !(flags & ORC_GET_MIN_EIGENVALS)) {
            float qx = next_pts[2 * k] - hwx, qy = next_pts[2 * k + 1] - hwy;
            int iqx = cv_floor(qx), iqy = cv_floor(qy);
            if (iqx < -win_w || iqx >= J->w || iqy < -win_h || iqy >= J->h) { status[k] = 0; continue; }
            float aa = qx - iqx, bb = qy - iqy;
            bilinear_weights(aa, bb, &iw00, &iw01, &iw10, &iw11);
            int64_t esum = 0; /* every |diff| is an integer, so an exact integer sum */
            for (int y = 0; y < win_h; y++)
                for (int x = 0; x < win_w; x++) {
                    int X = iqx + x, Y = iqy + y;
                    int diff = descale(img_at(J, X, Y) * iw00 + img_at(J, X + 1, Y) * iw01 + img_at(J, X, Y + 1) * iw10 +
                                       img_at(J, X + 1, Y + 1) * iw11, W_BITS - 5) - Ibuf[y * win_w + x];
                    esum += diff < 0 ? -diff : diff;
                }
            /* OpenCV: errval (fp32 running sum) * 1.f/(32*w*h), evaluated left to right. */
            err[k] = (float)esum * 1.f / (float)(32 * win_w * win_h);
        }
    }
}

/* cv::calcOpticalFlowPyrLK for two CV_8UC1 images.
 * next_pts must hold n points when flags has USE_INITIAL_FLOW; err may be NULL.
 * Returns the effective maxLevel (>= 0) or a negative ORC_E_* code.
 * trace_* may be NULL; when given they are n*32 entries (trace_pos n*32*2). */
int orc_calc_optical_flow_pyr_lk(const uint8_t* prev, long prev_stride, const uint8_t* next, long next_stride, int w, int h,
                                 const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n, int win_w,
                                 int win_h, int max_level, int crit_type, int crit_max_count, double crit_eps, int flags,
                                 double min_eig_thr, int nthreads, int32_t* trace_iters, uint8_t* trace_code, float* trace_pos)
{
    if (max_level < 0 || win_w <= 2 || win_h <= 2) return ORC_E_ARG; /* CV_Assert(maxLevel >= 0 && winSize > 2) */
    if (!prev || !next || w < 1 || h < 1 || n < 0) return ORC_E_ARG;
    if (n == 0) return 0;
    if (max_level > 31) max_level = 31;
    int max_count = (crit_type & ORC_TERM_COUNT) ? (crit_max_count < 0 ? 0 : crit_max_count > 100 ? 100 : crit_max_count) : 30;
    double eps = (crit_type & ORC_TERM_EPS) ? (crit_eps < 0 ? 0 : crit_eps > 10 ? 10 : crit_eps) : 0.01;
    eps *= eps;

    long dele = 0;
    long bytes = orc_lk_pyramid_bytes(w, h, win_w, win_h, max_level, &dele);
    uint8_t* pl = (uint8_t*)malloc((size_t)bytes);
    uint8_t* nl = (uint8_t*)malloc((size_t)bytes);
    int16_t* dl = (int16_t*)malloc(sizeof(int16_t) * (size_t)dele);
    if (!pl || !nl || !dl) { free(pl); free(nl); free(dl); return ORC_E_ARG; }
    int ml = orc_build_lk_pyramid(prev, w, h, prev_stride, win_w, win_h, max_level, pl, dl);
    orc_build_lk_pyramid(next, w, h, next_stride, win_w, win_h, max_level, nl, NULL);
    int ws[32], hs[32];
    orc_lk_level_sizes(w, h, win_w, win_h, max_level, ws, hs);
    orc_level P[32], N[32];
    long off = 0;
    for (int l = 0; l <= ml; l++) {
        P[l].img = pl + off; P[l].der = dl + 2 * off; P[l].w = ws[l]; P[l].h = hs[l];
        N[l].img = nl + off; N[l].der = NULL; N[l].w = ws[l]; N[l].h = hs[l];
        off += (long)ws[l] * hs[l];
    }
    for (int k = 0; k < n; k++) { status[k] = 1; if (err) err[k] = 0; }
    if (!(flags & ORC_USE_INITIAL_FLOW))
        for (int k = 0; k < 2 * n; k++) next_pts[k] = 0;
    orc_trace tr = {trace_iters, trace_code, trace_pos};
    orc_trace* trp = (trace_iters && trace_code && trace_pos) ? &tr : NULL;
    if (trp) { memset(trace_iters, 0, sizeof(int32_t) * 32 * (size_t)n); memset(trace_code, 0, 32 * (size_t)n); }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int16_t* Ibuf = (int16_t*)malloc(sizeof(int16_t) * (size_t)win_w * win_h * 3);
        int16_t* dbuf = Ibuf + (size_t)win_w * win_h;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 32)
#endif
        for (int k = 0; k < n; k++)
            track_point(P, N, ml, k, prev_pts, next_pts, status, err, win_w, win_h, max_count, eps, flags, min_eig_thr,
                        Ibuf, dbuf, trp);
        free(Ibuf);
    }
    (void)nthreads;
    free(pl); free(nl); free(dl);
    return ml;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
