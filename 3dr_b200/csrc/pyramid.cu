// Pyramid kernels of the dr3lk path (sm_100a).
//
//  * pyr_level_kernel  -- one launch per pyramid level for a batch of images: stages a (2*TOX+4) x (2*TOY+4)
//    uint8 source tile (halo 2, REFLECT_101 by index reflection) in shared memory once and produces from it
//      - the next Gaussian level: OpenCV pyrDown, separable [1 4 6 4 1], (sum+128)>>8   (SURVEY.md A.2), and
//      - the Scharr derivative of the source level as packed int16x2 (Ix, Iy)            (SURVEY.md A.3),
//    i.e. the derivative is fused into the pyramid build: the level is read from HBM exactly once.
//  * box_half_kernel   -- utils::reduce_to_half of the reference (src/utils.cpp:382-419) including its
//    rounding modes and the odd-width pointer walk.
#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

constexpr int TOX = 64;            // output (down-sampled) tile
constexpr int TOY = 16;
constexpr int SW = 2 * TOX + 4;    // source tile incl. halo 2
constexpr int SH = 2 * TOY + 4;
constexpr int SPITCH = SW + 4;     // 136
constexpr int PYR_THREADS = 256;

template <bool DOWN, bool DERIV>
__global__ void __launch_bounds__(PYR_THREADS)
pyr_level_kernel(const uint8_t* __restrict__ src, int w, int h, int src_pitch, long long src_stride,
                 uint8_t* __restrict__ dst, int dst_pitch, long long dst_stride, int* __restrict__ deriv, int dpitch,
                 long long deriv_stride)
{
    __shared__ __align__(16) uint8_t tile[SH][SPITCH];
    __shared__ short hrow[DOWN ? SH : 1][DOWN ? TOX : 1];

    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * (2 * TOX) - 2;  // source coordinate of tile[0][0]
    const int y0 = blockIdx.y * (2 * TOY) - 2;
    const uint8_t* s = src + (long long)img * src_stride;

    const bool interior = x0 >= 0 && y0 >= 0 && x0 + SW <= w && y0 + SH <= h;
    if (interior) {
        for (int idx = tid; idx < SH * SW; idx += PYR_THREADS) {
            int ty = idx / SW, tx = idx - ty * SW;
            tile[ty][tx] = __ldg(s + (long long)(y0 + ty) * src_pitch + x0 + tx);
        }
    } else {
        for (int idx = tid; idx < SH * SW; idx += PYR_THREADS) {
            int ty = idx / SW, tx = idx - ty * SW;
            int sx = reflect101(x0 + tx, w), sy = reflect101(y0 + ty, h);
            tile[ty][tx] = __ldg(s + (long long)sy * src_pitch + sx);
        }
    }
    __syncthreads();

    if (DERIV) {
        // source pixel (sx, sy) sits at tile[sy - y0][sx - x0] = tile[ly + 2][lx + 2]
        const int lx = tid & (2 * TOX - 1);
        int* d = deriv + (long long)img * deriv_stride;
        for (int ly = tid / (2 * TOX); ly < 2 * TOY; ly += PYR_THREADS / (2 * TOX)) {
            const int sx = x0 + 2 + lx, sy = y0 + 2 + ly;
            if (sx < w && sy < h) {
                const uint8_t* r0 = &tile[ly + 1][lx + 1];
                const uint8_t* r1 = &tile[ly + 2][lx + 1];
                const uint8_t* r2 = &tile[ly + 3][lx + 1];
                int a0 = r0[0], a1 = r0[1], a2 = r0[2];
                int b0 = r1[0], b2 = r1[2];
                int c0 = r2[0], c1 = r2[1], c2 = r2[2];
                // t0 = 3*(up+down) + 10*mid (vertical smooth), t1 = down - up (vertical difference)
                int t0l = 3 * (a0 + c0) + 10 * b0, t0r = 3 * (a2 + c2) + 10 * b2;
                int t1l = c0 - a0, t1m = c1 - a1, t1r = c2 - a2;
                int ix = t0r - t0l;
                int iy = 3 * (t1l + t1r) + 10 * t1m;
                d[(long long)sy * dpitch + sx] = (ix & 0xffff) | (iy << 16);
            }
        }
    }

    if (DOWN) {
        for (int idx = tid; idx < SH * TOX; idx += PYR_THREADS) {
            int ty = idx / TOX, ox = idx - ty * TOX;
            const uint8_t* r = &tile[ty][2 * ox];
            hrow[ty][ox] = (short)(r[0] + 4 * r[1] + 6 * r[2] + 4 * r[3] + r[4]);
        }
        __syncthreads();
        const int dw = (w + 1) >> 1, dh = (h + 1) >> 1;
        uint8_t* o = dst + (long long)img * dst_stride;
        const int ox = tid & (TOX - 1);
        for (int oy = tid / TOX; oy < TOY; oy += PYR_THREADS / TOX) {
            const int gx = blockIdx.x * TOX + ox, gy = blockIdx.y * TOY + oy;
            if (gx < dw && gy < dh) {
                int v = hrow[2 * oy][ox] + 4 * hrow[2 * oy + 1][ox] + 6 * hrow[2 * oy + 2][ox] + 4 * hrow[2 * oy + 3][ox] +
                        hrow[2 * oy + 4][ox];
                o[(long long)gy * dst_pitch + gx] = (uint8_t)((v + 128) >> 8);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
box_half_kernel(const uint8_t* __restrict__ src, int out_w, int out_h, long long row_stride, long long src_img_stride,
                uint8_t* __restrict__ dst, long long dst_img_stride, int sse2_rounding)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= out_w || i >= out_h) return;
    const uint8_t* s = src + (long long)blockIdx.z * src_img_stride;
    // reduce_to_half's scalar walk: `top` advances 2 per pixel and `stride` per row (src/utils.cpp:410-417), so
    // output row i starts at flat offset i*(2*out_w + stride); for even widths this is the plain 2x2 box.
    const long long t = (long long)i * (2LL * out_w + row_stride) + 2 * j;
    const unsigned a = __ldg(s + t), b = __ldg(s + t + 1), c = __ldg(s + t + row_stride), d = __ldg(s + t + row_stride + 1);
    unsigned v;
    if (sse2_rounding) {
        const unsigned v0 = (a + c + 1u) >> 1, v1 = (b + d + 1u) >> 1;  // _mm_avg_epu8 then _mm_avg_epu16
        v = (v0 + v1 + 1u) >> 1;
    } else {
        v = (a + b + c + d) >> 2;
    }
    dst[(long long)blockIdx.z * dst_img_stride + (long long)i * out_w + j] = (uint8_t)v;
}

}  // namespace

void launch_pyr_level(Launch& L, const uint8_t* src, int w, int h, int src_pitch, long long src_stride, uint8_t* dst,
                      int dst_pitch, long long dst_stride, int* deriv, int dpitch, long long deriv_stride, int n_img)
{
    if (L.err != cudaSuccess || n_img <= 0) return;
    dim3 grid((w + 2 * TOX - 1) / (2 * TOX), (h + 2 * TOY - 1) / (2 * TOY), n_img);
    if (dst && deriv)
        pyr_level_kernel<true, true><<<grid, PYR_THREADS, 0, L.stream>>>(src, w, h, src_pitch, src_stride, dst, dst_pitch,
                                                                       dst_stride, deriv, dpitch, deriv_stride);
    else if (dst)
        pyr_level_kernel<true, false><<<grid, PYR_THREADS, 0, L.stream>>>(src, w, h, src_pitch, src_stride, dst, dst_pitch,
                                                                        dst_stride, nullptr, 0, 0);
    else if (deriv)
        pyr_level_kernel<false, true><<<grid, PYR_THREADS, 0, L.stream>>>(src, w, h, src_pitch, src_stride, nullptr, 0, 0,
                                                                        deriv, dpitch, deriv_stride);
    else
        return;
    L.err = cudaGetLastError();
    L.launches++;
}

void launch_box_half(Launch& L, const uint8_t* src, int w, int h, long long row_stride, long long src_img_stride,
                     uint8_t* dst, long long dst_img_stride, int n_img, int sse2_rounding)
{
    if (L.err != cudaSuccess || n_img <= 0) return;
    const int out_w = w / 2, out_h = h / 2;
    dim3 grid((out_w + 255) / 256, out_h, n_img);
    box_half_kernel<<<grid, 256, 0, L.stream>>>(src, out_w, out_h, row_stride, src_img_stride, dst, dst_img_stride,
                                                sse2_rounding);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
