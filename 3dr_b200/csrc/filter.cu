// the post-filter right behind the LK call (reference src/initialization.cpp:615-635): stable
// compaction of the tracked points, disparity norm and pinhole bearing vectors, one kernel.
#include <math.h>

#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

constexpr int FT = 1024;

// Pinhole::cam2world for a camera WITH distortion (reference src/camera.cpp:32-40): cv::undistortPoints on the float pixel
// with the float K / D the constructor builds (src/camera.cpp:19-20), no R / P, default criteria = 5 fixed-point iterations.
// OpenCV's cvUndistortPointsInternal works in double and rounds the result to float; the statement order below is its own
// (terms of the unused coefficients k5..k11 are kept as the +0 they evaluate to, so signed zeros come out the same).
struct Distortion {
    double k[5];  // (double)(float)d0 .. d4
    double fx, fy, cx, cy;  // (double)(float) of the intrinsics: _cvK is a float matrix
    int on;       // Pinhole::_distortion = fabs(d0) > 1e-7 (src/camera.cpp:17)
};

__device__ __forceinline__ void undistort_point(const Distortion& D, float uf, float vf, double& xo, double& yo)
{
    const double ifx = __ddiv_rn(1.0, D.fx), ify = __ddiv_rn(1.0, D.fy);
    const double u = (double)uf, v = (double)vf;
    double x = __dmul_rn(__dsub_rn(u, D.cx), ifx), y = __dmul_rn(__dsub_rn(v, D.cy), ify);
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double zero3 = __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(0.0, r2), 0.0), r2), 0.0), r2);  // ((k7 r2 + k6) r2 + k5) r2
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(D.k[4], r2), D.k[1]), r2), D.k[0]), r2));
        const double icdist = __ddiv_rn(__dadd_rn(1.0, zero3), den);
        if (icdist < 0) {
            x = __dmul_rn(__dsub_rn(u, D.cx), ifx);
            y = __dmul_rn(__dsub_rn(v, D.cy), ify);
            break;
        }
        const double z1 = __dmul_rn(0.0, r2), z2 = __dmul_rn(__dmul_rn(0.0, r2), r2);  // k8 r2, k9 r2 r2 (and k10, k11)
        const double xy2k2 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, D.k[2]), x), y), xy2k3 = __dmul_rn(__dmul_rn(__dmul_rn(2.0, D.k[3]), x), y);
        const double dX = __dadd_rn(__dadd_rn(__dadd_rn(xy2k2, __dmul_rn(D.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)))), z1), z2);
        const double dY = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(D.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), xy2k3), z1), z2);
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    xo = (double)(float)x;  // dst is CV_32FC2, cam2world promotes px.x / px.y back to double
    yo = (double)(float)y;
}

// Single CTA (N is a few thousand at most on this path): per-thread status counts over contiguous slices, block scan,
// then ordered scatter.  All double arithmetic uses explicitly rounded intrinsics so it matches a plain C evaluation.
__global__ void __launch_bounds__(FT)
filter_tracks_kernel(const float2* __restrict__ ref, const float2* __restrict__ cur, const uint8_t* __restrict__ status, int n,
                     double fx, double fy, double cx, double cy, const Distortion dist, float2* __restrict__ out_ref, float2* __restrict__ out_cur,
                     double* __restrict__ out_disp, double* __restrict__ out_bearing, int* __restrict__ n_kept)
{
    __shared__ int s_cnt[FT];
    const int t = threadIdx.x;
    const int per = (n + FT - 1) / FT;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    int c = 0;
    for (int i = lo; i < hi; i++) c += status[i] != 0;
    s_cnt[t] = c;
    __syncthreads();
    for (int o = 1; o < FT; o <<= 1) {  // Hillis-Steele inclusive scan
        const int v = t >= o ? s_cnt[t - o] : 0;
        __syncthreads();
        s_cnt[t] += v;
        __syncthreads();
    }
    int pos = s_cnt[t] - c;
    if (t == FT - 1) *n_kept = s_cnt[t];
    for (int i = lo; i < hi; i++) {
        if (!status[i]) continue;
        const float2 r = ref[i], q = cur[i];
        out_ref[pos] = r;
        out_cur[pos] = q;
        // Vector2d(ref.x - cur.x, ref.y - cur.y).norm(): float differences promoted to double
        const double dx = (double)__fsub_rn(r.x, q.x), dy = (double)__fsub_rn(r.y, q.y);
        out_disp[pos] = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        if (out_bearing) {
            double x, y;
            if (dist.on) undistort_point(dist, q.x, q.y, x, y);
            else { x = __ddiv_rn(__dsub_rn((double)q.x, cx), fx); y = __ddiv_rn(__dsub_rn((double)q.y, cy), fy); }
            const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), 1.0));
            out_bearing[3 * pos] = __ddiv_rn(x, nrm);
            out_bearing[3 * pos + 1] = __ddiv_rn(y, nrm);
            out_bearing[3 * pos + 2] = __ddiv_rn(1.0, nrm);
        }
        pos++;
    }
}

// InitHelper::CheckFundamental (src/initialization.cpp:171-249) for many hypotheses.  One CTA per
// hypothesis.  The two symmetric-transfer terms of every match (the expensive part: two fp32 divisions) are computed by
// all threads in parallel into shared memory, a chunk of matches at a time; one thread then adds the chunk's terms to
// the score strictly in match order, so the fp32 score is accumulated exactly like the reference's scalar loop.  A term
// the reference skips (chi > th) is stored as +0.0f: x + 0 == x bit for bit (terms that count are >= 2.15, never -0),
// and a NaN chi is not "> th", so it propagates exactly as in the reference.
constexpr int SF_THREADS = 128, SF_CHUNK = 1024;

__global__ void __launch_bounds__(SF_THREADS)
score_fundamental_kernel(const float* __restrict__ F, int n_hyp, const float2* __restrict__ p1, const float2* __restrict__ p2, int n,
                         float inv_sigma2, float* __restrict__ scores, uint8_t* __restrict__ inliers)
{
    __shared__ __align__(16) float2 terms[SF_CHUNK];
    const int hyp = blockIdx.x;
    const float* f = F + 9 * hyp;
    const float f11 = f[0], f12 = f[1], f13 = f[2], f21 = f[3], f22 = f[4], f23 = f[5], f31 = f[6], f32 = f[7], f33 = f[8];
    const float th = 3.841f, thScore = 5.991f;
    float score = 0.f;
    for (int i0 = 0; i0 < n; i0 += SF_CHUNK) {
        const int m = min(SF_CHUNK, n - i0);
        for (int i = threadIdx.x; i < m; i += SF_THREADS) {
            const float2 a = p1[i0 + i], b = p2[i0 + i];
            const float u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
            const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(f11, u1), __fmul_rn(f12, v1)), f13);
            const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(f21, u1), __fmul_rn(f22, v1)), f23);
            const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(f31, u1), __fmul_rn(f32, v1)), f33);
            const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
            const float sq1 = __fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
            const float chi1 = __fmul_rn(sq1, inv_sigma2);
            const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(f11, u2), __fmul_rn(f21, v2)), f31);
            const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(f12, u2), __fmul_rn(f22, v2)), f32);
            const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(f13, u2), __fmul_rn(f23, v2)), f33);
            const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
            const float sq2 = __fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
            const float chi2 = __fmul_rn(sq2, inv_sigma2);
            const bool out1 = chi1 > th, out2 = chi2 > th;
            terms[i] = make_float2(out1 ? 0.f : __fsub_rn(thScore, chi1), out2 ? 0.f : __fsub_rn(thScore, chi2));
            if (inliers) inliers[(size_t)hyp * n + i0 + i] = (out1 || out2) ? 0 : 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const float4* t4 = reinterpret_cast<const float4*>(terms);
            int i = 0;
            for (; i + 2 <= m; i += 2) {  // two matches (four terms) per 16-byte load, added in match order
                const float4 t = t4[i >> 1];
                score = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(score, t.x), t.y), t.z), t.w);
            }
            if (i < m) score = __fadd_rn(__fadd_rn(score, terms[i].x), terms[i].y);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) scores[hyp] = score;
}

}  // namespace

void launch_score_fundamental(Launch& L, const float* F, int n_hyp, const float* p1, const float* p2, int n, float inv_sigma2,
                              float* scores, uint8_t* inliers)
{
    if (L.err != cudaSuccess || n_hyp <= 0) return;
    score_fundamental_kernel<<<n_hyp, SF_THREADS, 0, L.stream>>>(F, n_hyp, (const float2*)p1, (const float2*)p2, n, inv_sigma2,
                                                                       scores, inliers);
    L.err = cudaGetLastError();
    L.launches++;
}

void launch_filter_tracks(Launch& L, const float* ref, const float* cur, const uint8_t* status, int n, double fx, double fy, double cx,
                          double cy, const double* dist5, float* out_ref, float* out_cur, double* out_disp, double* out_bearing, int* n_kept)
{
    if (L.err != cudaSuccess) return;
    Distortion D;
    D.on = dist5 != nullptr && fabs(dist5[0]) > 1e-7;
    for (int i = 0; i < 5; i++) D.k[i] = dist5 ? (double)(float)dist5[i] : 0.0;
    D.fx = (double)(float)fx; D.fy = (double)(float)fy; D.cx = (double)(float)cx; D.cy = (double)(float)cy;
    filter_tracks_kernel<<<1, FT, 0, L.stream>>>((const float2*)ref, (const float2*)cur, status, n, fx, fy, cx, cy, D, (float2*)out_ref,
                                               (float2*)out_cur, out_disp, out_bearing, n_kept);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
