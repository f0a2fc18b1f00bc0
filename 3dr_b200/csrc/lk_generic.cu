// Generic pyramidal-LK tracker kernel (any window size) -- warp per feature, all levels in one launch.
//
// Restates OpenCV's LKTrackerInvoker (the reference reaches it through
// cv::calcOpticalFlowPyrLK at src/initialization.cpp:608-613) with these choices:
//   * window values are the same exact integers as OpenCV's (14-bit weights, cvRound, CV_DESCALE);
//   * the 2x2 normal-equation sums are accumulated exactly in int64 and converted to fp32 once
//     (x86 OpenCV accumulates in fp32 SIMD lanes; see oracle/lk_oracle.c);
//   * every fp32/fp64 scalar step uses explicitly rounded intrinsics (no FMA contraction), so results are
//     bit-identical to the CPU oracle.
// This kernel is the fallback for window sizes without a specialised kernel in lk_fast.cu; it keeps the
// template window (I, Ix, Iy) in shared memory and gathers pixels straight from global memory.
#include <float.h>

#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

constexpr int GEN_WARPS = 4;

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Weights {
    int w00, w01, w10, w11;
};

// iw00 = cvRound((1-a)(1-b)*2^14) ... (Appendix A.4 step 3); cvRound == round-half-even == __float2int_rn
__device__ __forceinline__ Weights make_weights(float a, float b)
{
    Weights q;
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    q.w00 = __float2int_rn(__fmul_rn(__fmul_rn(oma, omb), 16384.f));
    q.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, omb), 16384.f));
    q.w10 = __float2int_rn(__fmul_rn(__fmul_rn(oma, b), 16384.f));
    q.w11 = (1 << W_BITS) - q.w00 - q.w01 - q.w10;
    return q;
}

__device__ __forceinline__ int img_px(const uint8_t* __restrict__ img, int pitch, int w, int h, int x, int y)
{
    return __ldg(img + (long long)reflect101(y, h) * pitch + reflect101(x, w));
}

// bilinear sample of the image window value with 5 fractional bits: (S + 2^8) >> 9
__device__ __forceinline__ int img_sample(const uint8_t* __restrict__ img, int pitch, int w, int h, int x, int y,
                                          const Weights& q, bool interior)
{
    int p00, p01, p10, p11;
    if (interior) {
        const uint8_t* r = img + (long long)y * pitch + x;
        p00 = __ldg(r); p01 = __ldg(r + 1); p10 = __ldg(r + pitch); p11 = __ldg(r + pitch + 1);
    } else {
        p00 = img_px(img, pitch, w, h, x, y); p01 = img_px(img, pitch, w, h, x + 1, y);
        p10 = img_px(img, pitch, w, h, x, y + 1); p11 = img_px(img, pitch, w, h, x + 1, y + 1);
    }
    return (p00 * q.w00 + p01 * q.w01 + p10 * q.w10 + p11 * q.w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
}

__device__ __forceinline__ int der_px(const int* __restrict__ der, int dpitch, int w, int h, int x, int y)
{
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0;  // derivatives are zero-padded
    return __ldg(der + (long long)y * dpitch + x);
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
lk_generic_kernel(const __grid_constant__ LKParams P)
{
    extern __shared__ int smem_dyn[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int f = blockIdx.x * GEN_WARPS + warp;
    if (f >= P.n_total) return;

    const int npix = P.win_w * P.win_h;
    // per-warp template: dxdy[npix] (packed s16x2) then I[npix] (s16)
    int* s_dxdy = smem_dyn + (size_t)warp * ((npix * 3 + 1) / 2 + 1);
    short* s_I = reinterpret_cast<short*>(s_dxdy + npix);

    // which frame pair does this feature belong to?
    const int pair = P.uniform_n > 0 ? f / P.uniform_n : __ldg(P.pair_idx + f);

    const float2 pp = sanitize_point(P.prev_pts[f]);
    float2 np = make_float2(0.f, 0.f);
    if (P.flags & DR3LK_USE_INITIAL_FLOW) np = sanitize_point(P.next_pts[f]);
    const float hwx = (P.win_w - 1) * 0.5f, hwy = (P.win_h - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const bool want_err = P.err != nullptr;
    const bool get_min_eig = (P.flags & DR3LK_GET_MIN_EIGENVALS) != 0;

    int status = 1;
    float err = 0.f;
    unsigned n_iters = 0, n_templates = 0, err_pass = 0;

    for (int level = P.max_level; level >= 0; --level) {
        const LevelDesc& L = P.lv[level];
        const uint8_t* imgI = L.prev + (unsigned long long)(unsigned)pair * L.prev_stride;
        const uint8_t* imgJ = L.next + (unsigned long long)(unsigned)pair * L.next_stride;
        const int* der = L.deriv + (unsigned long long)(unsigned)pair * L.deriv_stride;
        const int w = L.w, h = L.h;
        const float sc = __int_as_float((127 - level) << 23);  // (float)(1./(1 << level))

        float px = __fmul_rn(pp.x, sc), py = __fmul_rn(pp.y, sc);
        float nx, ny;
        if (level == P.max_level) {
            if (P.flags & DR3LK_USE_INITIAL_FLOW) { nx = __fmul_rn(np.x, sc); ny = __fmul_rn(np.y, sc); }
            else { nx = px; ny = py; }
        } else {
            nx = __fmul_rn(np.x, 2.f); ny = __fmul_rn(np.y, 2.f);
        }
        np.x = nx; np.y = ny;

        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -P.win_w || ipx >= w || ipy < -P.win_h || ipy >= h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        Weights q = make_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy));

        // ---- template window: I (5 fractional bits), Ix, Iy, and the Gram matrix ----
        n_templates++;
        const bool interiorI = ipx >= 0 && ipy >= 0 && ipx + P.win_w < w && ipy + P.win_h < h;
        long long a11 = 0, a12 = 0, a22 = 0;
        for (int i = lane; i < npix; i += 32) {
            const int y = i / P.win_w, x = i - y * P.win_w;
            const int X = ipx + x, Y = ipy + y;
            const int ival = img_sample(imgI, L.pitch_p, w, h, X, Y, q, interiorI);
            int d00, d01, d10, d11;
            if (interiorI) {
                const int* r = der + (long long)Y * L.dpitch + X;
                d00 = __ldg(r); d01 = __ldg(r + 1); d10 = __ldg(r + L.dpitch); d11 = __ldg(r + L.dpitch + 1);
            } else {
                d00 = der_px(der, L.dpitch, w, h, X, Y); d01 = der_px(der, L.dpitch, w, h, X + 1, Y);
                d10 = der_px(der, L.dpitch, w, h, X, Y + 1); d11 = der_px(der, L.dpitch, w, h, X + 1, Y + 1);
            }
            const int ix = ((short)d00 * q.w00 + (short)d01 * q.w01 + (short)d10 * q.w10 + (short)d11 * q.w11 +
                            (1 << (W_BITS - 1))) >> W_BITS;
            const int iy = ((d00 >> 16) * q.w00 + (d01 >> 16) * q.w01 + (d10 >> 16) * q.w10 + (d11 >> 16) * q.w11 +
                            (1 << (W_BITS - 1))) >> W_BITS;
            s_I[i] = (short)ival;
            s_dxdy[i] = (ix & 0xffff) | (iy << 16);
            a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
        }
        a11 = warp_sum_ll(a11); a12 = warp_sum_ll(a12); a22 = warp_sum_ll(a22);
        __syncwarp();

        const float A11 = __fmul_rn(__ll2float_rn(a11), FLT_SCALE);
        const float A12 = __fmul_rn(__ll2float_rn(a12), FLT_SCALE);
        const float A22 = __fmul_rn(__ll2float_rn(a22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * P.win_w * P.win_h));
        if (want_err && get_min_eig) err = minEig;
        if (minEig < P.min_eig_thr || D < FLT_EPSILON) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);

        nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_count; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -P.win_w || inx >= w || iny < -P.win_h || iny >= h) {
                if (level == 0) status = 0;
                break;
            }
            q = make_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny));
            const bool interiorJ = inx >= 0 && iny >= 0 && inx + P.win_w < w && iny + P.win_h < h;
            long long b1 = 0, b2 = 0;
            for (int i = lane; i < npix; i += 32) {
                const int y = i / P.win_w, x = i - y * P.win_w;
                const int diff = img_sample(imgJ, L.pitch_n, w, h, inx + x, iny + y, q, interiorJ) - s_I[i];
                const int dd = s_dxdy[i];
                b1 += diff * (int)(short)dd;
                b2 += diff * (dd >> 16);
            }
            b1 = warp_sum_ll(b1); b2 = warp_sum_ll(b2);
            n_iters++;
            const float fb1 = __fmul_rn(__ll2float_rn(b1), FLT_SCALE), fb2 = __fmul_rn(__ll2float_rn(b2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            np.x = __fadd_rn(nx, hwx); np.y = __fadd_rn(ny, hwy);
            const double dd2 = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy));
            if (dd2 <= P.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                np.x = __fsub_rn(np.x, __fmul_rn(dx, 0.5f));
                np.y = __fsub_rn(np.y, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }

        if (status && want_err && level == 0 && !get_min_eig) {
            const float qx = __fsub_rn(np.x, hwx), qy = __fsub_rn(np.y, hwy);
            const int iqx = __float2int_rd(qx), iqy = __float2int_rd(qy);
            if (iqx < -P.win_w || iqx >= w || iqy < -P.win_h || iqy >= h) {
                status = 0;
                continue;
            }
            q = make_weights(__fsub_rn(qx, (float)iqx), __fsub_rn(qy, (float)iqy));
            const bool interiorJ = iqx >= 0 && iqy >= 0 && iqx + P.win_w < w && iqy + P.win_h < h;
            long long es = 0;
            for (int i = lane; i < npix; i += 32) {
                const int y = i / P.win_w, x = i - y * P.win_w;
                const int diff = img_sample(imgJ, L.pitch_n, w, h, iqx + x, iqy + y, q, interiorJ) - s_I[i];
                es += diff < 0 ? -diff : diff;
            }
            es = warp_sum_ll(es);
            err_pass = 1;
            err = __fdiv_rn(__fmul_rn(__ll2float_rn(es), 1.f), (float)(32 * P.win_w * P.win_h));
        }
    }

    if (lane == 0) {
        P.next_out[f] = np;
        P.status[f] = (uint8_t)status;
        if (want_err) P.err[f] = err;
        if (P.stats) P.stats[f] = (n_iters & 0xffffu) | ((n_templates & 0xffu) << 16) | (err_pass << 24);
    }
}

}  // namespace

// pair index of every point: binary search in the offsets (only for ragged batches)
__global__ void pair_index_kernel(const int* __restrict__ offs, int batch, int n_total, int* __restrict__ pair_idx)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_total) return;
    int lo = 0, hi = batch;  // offs[lo] <= f < offs[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offs + mid) <= f) lo = mid; else hi = mid;
    }
    pair_idx[f] = lo;
}

void launch_pair_index(Launch& L, const int* pts_offset_dev, int batch, int n_total, int* pair_idx_dev)
{
    if (L.err != cudaSuccess || n_total <= 0) return;
    pair_index_kernel<<<(n_total + 255) / 256, 256, 0, L.stream>>>(pts_offset_dev, batch, n_total, pair_idx_dev);
    L.err = cudaGetLastError();
    L.launches++;
}

size_t lk_generic_smem_bytes(int win_w, int win_h)
{
    const int npix = win_w * win_h;
    return (size_t)GEN_WARPS * ((npix * 3 + 1) / 2 + 1) * sizeof(int);
}

void launch_lk_generic(Launch& L, const LKParams& p)
{
    if (L.err != cudaSuccess || p.n_total <= 0) return;
    const size_t smem = lk_generic_smem_bytes(p.win_w, p.win_h);
    if (smem > 48 * 1024) {
        L.err = cudaFuncSetAttribute(lk_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (L.err != cudaSuccess) return;
    }
    const int blocks = (p.n_total + GEN_WARPS - 1) / GEN_WARPS;
    lk_generic_kernel<<<blocks, GEN_WARPS * 32, smem, L.stream>>>(p);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
