// Internal declarations shared by the dr3lk CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dr3lk.h"

namespace dr3lk {

constexpr int kMaxLevels = DR3LK_MAX_LEVELS;
constexpr int W_BITS = 14;  // OpenCV lkpyramid.cpp fixed-point weight bits

// One pyramid level of a batch of frame pairs, as the LK kernels see it.
struct LevelDesc {
    const uint8_t* prev;   // [batch][h][pitch_p]  level image of the previous frame
    const uint8_t* next;   // [batch][h][pitch_n]
    const int* deriv;      // [batch][h][dpitch]   Scharr (Ix, Iy) as packed int16x2, previous frame only
    unsigned prev_stride;  // bytes between consecutive pairs (< 4 GiB per image; pair * stride is a 32x32->64 multiply)
    unsigned next_stride;
    unsigned deriv_stride; // ints between consecutive pairs
    int pitch_p, pitch_n;    // bytes
    int dpitch;              // ints
    int w, h;
};

struct LKParams {
    LevelDesc lv[kMaxLevels];
    const float2* prev_pts;
    float2* next_pts;
    uint8_t* status;
    float* err;            // may be null
    uint32_t* stats;       // may be null
    const int* pair_idx;   // device, one frame-pair index per point (used when uniform_n == 0)
    int uniform_n;         // > 0: every pair has exactly this many points (pair = point / uniform_n)
    int fast_ok;           // all level images / derivatives are 16-B aligned (base, pitch, stride)
    int* work_counter;     // two device ints (zero-initialised): counter [work_epoch & 1] feeds the persistent warps of
    int work_epoch;        // this launch, which also re-zeroes the other one for the next launch (no memset per call)
    float eps2_lo, eps2_hi; // fp32 brackets of eps2: below lo / above hi the fp32 estimate of |delta|^2 decides
    double eps2;           // criteria.epsilon^2
    double min_eig_thr;
    int batch;
    int n_total;
    int max_level;         // effective
    int win_w, win_h;
    int max_count;
    int flags;
};

// ---- launchers (each returns the number of kernels it launched, or -1 after setting *err) ----
struct Launch {
    cudaStream_t stream;
    cudaError_t err;
    int launches;
};

// pyramid.cu
// One level step: reads the source level of n_prev previous-frame images and n_next next-frame images (same size and
// pitch), writes the 5-tap down-sampled next level of both (when `down`) and the Scharr derivative of the previous-frame
// source level (when deriv != nullptr) -- one launch for both pyramids.
struct PyrLevelArgs {
    const uint8_t* prev_src; const uint8_t* next_src;
    uint8_t* prev_dst; uint8_t* next_dst;
    unsigned prev_src_stride, next_src_stride, prev_dst_stride, next_dst_stride;  // bytes between images
    int* deriv; unsigned deriv_stride;  // ints between images
    int w, h, src_pitch, dst_pitch, dpitch;
    int n_prev, n_next;
    bool down;
};
void launch_pyr_level(Launch& L, const PyrLevelArgs& a);
void launch_box_half(Launch& L, const uint8_t* src, int w, int h, long long row_stride, long long src_img_stride,
                     uint8_t* dst, long long dst_img_stride, int n_img, int sse2_rounding);

// fast.cu
void launch_fast_level(Launch& L, const uint8_t* img, uint8_t* score, int w, int h, int level, int fast_threshold, int cell_size,
                       int grid_cols, float thr_f, double thr_d, const uint8_t* occupancy, unsigned long long* cell_best);
void launch_fast_gather(Launch& L, const unsigned long long* cell_best, int n_cells, const int* level_w, int* out_xy, int* out_level,
                        float* out_score, int* n_out);

// filter.cu
void launch_filter_tracks(Launch& L, const float* ref, const float* cur, const uint8_t* status, int n, double fx, double fy, double cx,
                          double cy, float* out_ref, float* out_cur, double* out_disp, double* out_bearing, int* n_kept);

void launch_score_fundamental(Launch& L, const float* F, int n_hyp, const float* p1, const float* p2, int n, float inv_sigma2,
                              float* scores, uint8_t* inliers);

// lk_generic.cu
void launch_lk_generic(Launch& L, const LKParams& p);
void launch_pair_index(Launch& L, const int* pts_offset_dev, int batch, int n_total, int* pair_idx_dev);
// lk_fast.cu -- returns false when the window size has no specialised kernel
bool launch_lk_fast(Launch& L, const LKParams& p);
bool lk_fast_supported(int win_w, int win_h);

// ---- device helpers ----
__device__ __forceinline__ int reflect101(int p, int len)
{
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

}  // namespace dr3lk
