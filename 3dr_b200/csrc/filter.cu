// the post-filter right behind the LK call (reference src/initialization.cpp:615-635): stable
// compaction of the tracked points, disparity norm and pinhole bearing vectors, one kernel.
#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

constexpr int FT = 1024;

// Single CTA (N is a few thousand at most on this path): per-thread status counts over contiguous slices, block scan,
// then ordered scatter.  All double arithmetic uses explicitly rounded intrinsics so it matches a plain C evaluation.
__global__ void __launch_bounds__(FT)
filter_tracks_kernel(const float2* __restrict__ ref, const float2* __restrict__ cur, const uint8_t* __restrict__ status, int n,
                     double fx, double fy, double cx, double cy, float2* __restrict__ out_ref, float2* __restrict__ out_cur,
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
            const double x = __ddiv_rn(__dsub_rn((double)q.x, cx), fx), y = __ddiv_rn(__dsub_rn((double)q.y, cy), fy);
            const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), 1.0));
            out_bearing[3 * pos] = __ddiv_rn(x, nrm);
            out_bearing[3 * pos + 1] = __ddiv_rn(y, nrm);
            out_bearing[3 * pos + 2] = __ddiv_rn(1.0, nrm);
        }
        pos++;
    }
}

// InitHelper::CheckFundamental (src/initialization.cpp:171-249) for many hypotheses.  One thread per
// hypothesis walks the matches in order, so the fp32 score is accumulated exactly like the reference's scalar loop; the
// matches are staged through shared memory in tiles so all hypotheses of a block read them once.
__global__ void __launch_bounds__(128)
score_fundamental_kernel(const float* __restrict__ F, int n_hyp, const float2* __restrict__ p1, const float2* __restrict__ p2, int n,
                         float inv_sigma2, float* __restrict__ scores, uint8_t* __restrict__ inliers)
{
    __shared__ float4 tile[256];
    const int hyp = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = hyp < n_hyp;
    float f11 = 0, f12 = 0, f13 = 0, f21 = 0, f22 = 0, f23 = 0, f31 = 0, f32 = 0, f33 = 0;
    if (active) {
        const float* f = F + 9 * hyp;
        f11 = f[0]; f12 = f[1]; f13 = f[2]; f21 = f[3]; f22 = f[4]; f23 = f[5]; f31 = f[6]; f32 = f[7]; f33 = f[8];
    }
    const float th = 3.841f, thScore = 5.991f;
    float score = 0.f;
    for (int i0 = 0; i0 < n; i0 += 256) {
        __syncthreads();
        for (int i = threadIdx.x; i < 256 && i0 + i < n; i += blockDim.x) {
            const float2 a = p1[i0 + i], b = p2[i0 + i];
            tile[i] = make_float4(a.x, a.y, b.x, b.y);
        }
        __syncthreads();
        if (!active) continue;
        const int m = min(256, n - i0);
        for (int i = 0; i < m; i++) {
            const float u1 = tile[i].x, v1 = tile[i].y, u2 = tile[i].z, v2 = tile[i].w;
            bool in = true;
            const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(f11, u1), __fmul_rn(f12, v1)), f13);
            const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(f21, u1), __fmul_rn(f22, v1)), f23);
            const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(f31, u1), __fmul_rn(f32, v1)), f33);
            const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
            const float sq1 = __fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
            const float chi1 = __fmul_rn(sq1, inv_sigma2);
            if (chi1 > th) in = false; else score = __fadd_rn(score, __fsub_rn(thScore, chi1));
            const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(f11, u2), __fmul_rn(f21, v2)), f31);
            const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(f12, u2), __fmul_rn(f22, v2)), f32);
            const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(f13, u2), __fmul_rn(f23, v2)), f33);
            const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
            const float sq2 = __fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
            const float chi2 = __fmul_rn(sq2, inv_sigma2);
            if (chi2 > th) in = false; else score = __fadd_rn(score, __fsub_rn(thScore, chi2));
            if (inliers) inliers[(size_t)hyp * n + i0 + i] = in ? 1 : 0;
        }
    }
    if (active) scores[hyp] = score;
}

}  // namespace

void launch_score_fundamental(Launch& L, const float* F, int n_hyp, const float* p1, const float* p2, int n, float inv_sigma2,
                              float* scores, uint8_t* inliers)
{
    if (L.err != cudaSuccess || n_hyp <= 0) return;
    score_fundamental_kernel<<<(n_hyp + 127) / 128, 128, 0, L.stream>>>(F, n_hyp, (const float2*)p1, (const float2*)p2, n, inv_sigma2,
                                                                       scores, inliers);
    L.err = cudaGetLastError();
    L.launches++;
}

void launch_filter_tracks(Launch& L, const float* ref, const float* cur, const uint8_t* status, int n, double fx, double fy, double cx,
                          double cy, float* out_ref, float* out_cur, double* out_disp, double* out_bearing, int* n_kept)
{
    if (L.err != cudaSuccess) return;
    filter_tracks_kernel<<<1, FT, 0, L.stream>>>((const float2*)ref, (const float2*)cur, status, n, fx, fy, cx, cy, (float2*)out_ref,
                                               (float2*)out_cur, out_disp, out_bearing, n_kept);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
