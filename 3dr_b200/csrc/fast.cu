// the prevPts provider of the LK path on the device -- FAST-10 corners with their scores,
// 3x3 non-maximum suppression and the per-grid-cell Shi-Tomasi selection of the reference's
// feature_detection::FastDetector::detect (src/features.cpp:43-98; utils::shi_tomasi_score src/utils.cpp:282-321),
// consuming the Frame's box pyramid.  The FAST routines themselves live in the external `fast` library
// (uzh-rpg/fast, not vendored); they are restated here in a map-based, order-free form:
//   score map   : per pixel, 0 or the largest threshold b >= b0 for which >= 10 contiguous circle pixels are all
//                 brighter than p + b or all darker than p - b  (== fast_corner_detect_10 + fast_corner_score_10)
//   non-max     : a corner survives when all 8 neighbours have a strictly smaller score   (== fast_nonmax_3x3)
//   selection   : per grid cell the survivor with the largest Shi-Tomasi score, ties to the first in (level, raster)
//                 order, as one 64-bit atomicMax on (score bits, inverted order)
#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

// ARC = 10 is the reference's detector (fast_corner_detect_10); ARC = 9 exists so that the same kernels can be pinned against
// OpenCV's FAST-9 (tests/test_gpu_fast_cv2.py through dr3lk_debug_set_fast_arc) -- the `fast` library itself is absent here.
template <int ARC>
__device__ __forceinline__ unsigned run_arc(unsigned m)  // m: 16-bit circular mask; non-zero iff it has ARC contiguous set bits
{
    const unsigned mm = m | (m << 16);
    const unsigned t2 = mm & (mm >> 1), t4 = t2 & (t2 >> 2), t8 = t4 & (t4 >> 4);
    return (t8 & ((ARC == 10 ? t2 : mm) >> 8)) & 0xffffu;
}

// largest over the 16 arcs of ARC contiguous circle positions of the smallest d[] in the arc
template <int ARC>
__device__ __forceinline__ int best_arc_min(const int (&d)[16])
{
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m2[i] = min(d[i], d[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++) m8[i] = min(m4[i], m4[(i + 4) & 15]);
    int best = -1000;
#pragma unroll
    for (int i = 0; i < 16; i++) best = max(best, min(m8[i], ARC == 10 ? m2[(i + 8) & 15] : d[(i + 8) & 15]));
    return best;
}

template <int ARC>
__global__ void __launch_bounds__(256)
fast_score_kernel(const uint8_t* __restrict__ img, int w, int h, int b0, uint8_t* __restrict__ score)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    uint8_t out = 0;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const uint8_t* p = img + (long long)y * w + x;
        const int c = __ldg(p);
        // 16-pixel Bresenham circle of radius 3, clockwise from (0, +3)
        constexpr int CX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
        constexpr int CY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
        int d[16];
        unsigned bright = 0, dark = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            d[i] = (int)__ldg(p + CY[i] * w + CX[i]) - c;
            bright |= (d[i] > b0 ? 1u : 0u) << i;
            dark |= (d[i] < -b0 ? 1u : 0u) << i;
        }
        if (run_arc<ARC>(bright) | run_arc<ARC>(dark)) {
            // corner at threshold b  <=>  some arc has all d > b (or all -d > b)  <=>  b <= best - 1
            const int bb = best_arc_min<ARC>(d);
#pragma unroll
            for (int i = 0; i < 16; i++) d[i] = -d[i];
            const int bd = best_arc_min<ARC>(d);
            out = (uint8_t)min(max(bb, bd) - 1, 254);
        }
    }
    score[(long long)y * w + x] = out;
}

__global__ void __launch_bounds__(256)
fast_select_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ score, int w, int h, int level, int cell_size,
                   int grid_cols, float thr_f, double thr_d, const uint8_t* __restrict__ occupancy,
                   unsigned long long* __restrict__ cell_best)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return;
    const uint8_t* sp = score + (long long)y * w + x;
    const int s = sp[0];
    if (s == 0) return;
    // fast_nonmax_3x3: suppressed when a neighbouring corner has a score >= this one (non-corners hold 0)
    if (sp[-1] >= s || sp[1] >= s || sp[-w - 1] >= s || sp[-w] >= s || sp[-w + 1] >= s || sp[w - 1] >= s || sp[w] >= s || sp[w + 1] >= s)
        return;
    const int scale = 1 << level;
    const int k = ((y * scale) / cell_size) * grid_cols + (x * scale) / cell_size;
    if (occupancy && occupancy[k]) return;
    // utils::shi_tomasi_score: 8x8 box, central differences; the sums are exact integers below 2^24
    if (x - 4 < 1 || x + 4 >= w - 1 || y - 4 < 1 || y + 4 >= h - 1) return;  // score 0 never beats the threshold
    int sxx = 0, syy = 0, sxy = 0;
    for (int yy = y - 4; yy < y + 4; yy++) {
        const uint8_t* r = img + (long long)yy * w + (x - 4);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int dx = (int)__ldg(r + i + 1) - (int)__ldg(r + i - 1);
            const int dy = (int)__ldg(r + i + w) - (int)__ldg(r + i - w);
            sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
        }
    }
    const float dXX = __fmul_rn((float)sxx, 1.f / 128.f), dYY = __fmul_rn((float)syy, 1.f / 128.f), dXY = __fmul_rn((float)sxy, 1.f / 128.f);
    const float tr = __fadd_rn(dXX, dYY);
    const float det = __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY));
    const float rad = __fsub_rn(__fmul_rn(tr, tr), __fmul_rn(4.f, det));
    const float st = __fmul_rn(__fsub_rn(tr, __fsqrt_rn(rad)), 0.5f);
    if (!(st > thr_f) || !((double)st > thr_d)) return;
    const unsigned order = ((unsigned)level << 28) | (unsigned)(y * w + x);
    const unsigned long long key = ((unsigned long long)__float_as_uint(st) << 32) | (0xffffffffu - order);
    atomicMax(cell_best + k, key);
}

// one warp: cells in order -> compact list of features (level-0 coordinates, level, score)
__global__ void fast_gather_kernel(const unsigned long long* __restrict__ cell_best, int n_cells, const int* __restrict__ level_w,
                                   int* __restrict__ out_xy, int* __restrict__ out_level, float* __restrict__ out_score,
                                   int* __restrict__ n_out)
{
    const int lane = threadIdx.x;
    int base = 0;
    for (int k0 = 0; k0 < n_cells; k0 += 32) {
        const int k = k0 + lane;
        const unsigned long long key = k < n_cells ? cell_best[k] : 0ull;
        const unsigned m = __ballot_sync(0xffffffffu, key != 0ull);
        if (key != 0ull) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            const unsigned order = 0xffffffffu - (unsigned)(key & 0xffffffffu);
            const int lvl = (int)(order >> 28), idx = (int)(order & 0x0fffffffu), lw = level_w[lvl];
            out_xy[2 * pos] = (idx % lw) << lvl;
            out_xy[2 * pos + 1] = (idx / lw) << lvl;
            out_level[pos] = lvl;
            out_score[pos] = __uint_as_float((unsigned)(key >> 32));
        }
        base += __popc(m);
    }
    if (lane == 0) *n_out = base;
}

}  // namespace

void launch_fast_level(Launch& L, const uint8_t* img, uint8_t* score, int w, int h, int level, int fast_threshold, int cell_size,
                       int grid_cols, float thr_f, double thr_d, const uint8_t* occupancy, unsigned long long* cell_best, int arc)
{
    if (L.err != cudaSuccess) return;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    if (arc == 9) fast_score_kernel<9><<<grid, 256, 0, L.stream>>>(img, w, h, fast_threshold, score);
    else fast_score_kernel<10><<<grid, 256, 0, L.stream>>>(img, w, h, fast_threshold, score);
    fast_select_kernel<<<grid, 256, 0, L.stream>>>(img, score, w, h, level, cell_size, grid_cols, thr_f, thr_d, occupancy, cell_best);
    L.err = cudaGetLastError();
    L.launches += 2;
}

void launch_fast_gather(Launch& L, const unsigned long long* cell_best, int n_cells, const int* level_w, int* out_xy, int* out_level,
                        float* out_score, int* n_out)
{
    if (L.err != cudaSuccess) return;
    fast_gather_kernel<<<1, 32, 0, L.stream>>>(cell_best, n_cells, level_w, out_xy, out_level, out_score, n_out);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
