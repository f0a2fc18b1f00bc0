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

}  // namespace

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
