// Internal declarations shared by the dr3lk CUDA translation units (sm_100a only).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dr3lk.h"

namespace dr3lk {

constexpr int kMaxLevels = DR3LK_MAX_LEVELS;
constexpr int W_BITS = 14;  // OpenCV lkpyramid.cpp fixed-point weight bits

// Apron ("padding") around every pyramid level the specialised LK kernels read.  Like OpenCV's
// buildOpticalFlowPyramid (level images padded by winSize with BORDER_REFLECT_101, derivatives with BORDER_CONSTANT 0,
// SURVEY.md Appendix A.2/A.3) every level image carries kApronX columns / apron_y(win_h) rows of reflected pixels on
// each side and every derivative level an apron of zeros, so the LK kernels stage windows and search regions with plain
// 16-byte row copies: no index reflection, no zero-fill predicates, no byte-wise border path.
constexpr int kApronX = 32;                                           // bytes; a multiple of 16 keeps x = 0 16-B aligned
__host__ __device__ constexpr int apron_y(int win_h) { return win_h + 4; }                    // rows: window + 1, + staging slack
__host__ __device__ constexpr int deriv_apron_x(int win_w) { return (win_w + 1 + 3) / 4 * 4; }  // ints (16-B multiple)

// One pyramid level of a batch of frame pairs, as the LK kernels see it.  prev / next / deriv point at pixel (0, 0) of
// image 0; with lk.fast_ok the aprons above exist around every image (negative coordinates are addressable).
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

// TMA descriptors of one level for the specialised LK kernels: boxes of the template window (previous image, u8), of its
// derivatives (int32 words) and of the search region (next image, u8) over the apron-carrying level allocations
// [batch][rows + 2 * apron_y][pitch]; coordinate (0, 0) is the first apron byte.  Only the first kTmaLevels levels can be
// described (kernel parameter space); deeper pyramids use the cp.async staging.
constexpr int kTmaLevels = 6;   // keeps LKParams below the classic 4 KB kernel-parameter limit
struct LevelTma {
    CUtensorMap prev, deriv, next;
};

struct LKParams {
    LevelDesc lv[kMaxLevels];
    LevelTma tma[kTmaLevels];
    int use_tma;           // the descriptors above are valid: stage with cp.async.bulk.tensor instead of cp.async
    const float2* prev_pts;
    float2* next_pts;        // initial estimates (DR3LK_USE_INITIAL_FLOW), device memory
    float2* next_out;        // results; == next_pts except on the latency path, where it is the mapped pinned mirror of the caller
    uint8_t* status;
    float* err;            // may be null
    uint32_t* stats;       // may be null
    const int* pair_idx;   // device, one frame-pair index per point (used when uniform_n == 0)
    int uniform_n;         // > 0: every pair has exactly this many points (pair = point / uniform_n)
    int fast_ok;           // every level is 16-B aligned (origin, pitch, stride) and carries the aprons (kApronX, apron_y, deriv_apron_x)
    int* work_counter;     // two device ints (zero-initialised): counter [work_epoch & 1] feeds the persistent warps of
    int work_epoch;        // this launch, which also re-zeroes the other one for the next launch (no memset per call)
    int fetch_n;           // features a warp reserves per atomic (set by the launcher)
    float eps2_lo, eps2_hi; // fp32 brackets of eps2: below lo / above hi the fp32 estimate of |delta|^2 decides
    double eps2;           // criteria.epsilon^2
    float min_eig_thr;     // OpenCV's LKTrackerInvoker keeps minEigThreshold as a float: float < float
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
    // Latency path: launch with programmatic stream serialization, so that this kernel's launch overlaps the tail of the
    // kernel before it in the stream.  Every kernel launched this way starts with grid_dependency_wait(), which returns
    // once the preceding grid has completed and its writes are visible; without the attribute the wait is a no-op.
    bool pdl = false;
};

__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// (An explicit early griddepcontrol.launch_dependents at the top of the pyramid kernels was measured too: 100.4 instead of
// 94.7 us per call -- the dependents then sit on the SMs spinning in their wait while the primary still needs them.  The
// implicit trigger at the primary's exit is what is used.)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(const Launch& L, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = L.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
    int dst_apron_x, dst_apron_y;  // > 0: also write the REFLECT_101 apron of the down-sampled level
    int src_apron_x, src_apron_y;  // apron the source level carries (0: none -- tiles at the image edge reflect indices)
};
void launch_pyr_level(Launch& L, const PyrLevelArgs& a);
// Level 0 into its apron-carrying scratch copy: dst(x, y) = src(reflect101(x), reflect101(y)) for -ax <= x < w + ax,
// -ay <= y < h + ay, for n_a images of set A followed by n_b images of set B (same geometry; one launch serves the previous
// and the next frames).  dst_* point at pixel (0, 0) of image 0 and are 16-B aligned (origin, pitch, stride); src is arbitrary.
void launch_pad_level0(Launch& L, const uint8_t* src_a, const uint8_t* src_b, size_t src_pitch, size_t src_stride, uint8_t* dst_a,
                       uint8_t* dst_b, int dst_pitch, size_t dst_stride, int w, int h, int ax, int ay, int n_a, int n_b);
void launch_box_half(Launch& L, const uint8_t* src, int w, int h, long long row_stride, long long src_img_stride,
                     uint8_t* dst, long long dst_img_stride, int n_img, int sse2_rounding);

// fast.cu
void launch_fast_level(Launch& L, const uint8_t* img, uint8_t* score, int w, int h, int level, int fast_threshold, int cell_size,
                       int grid_cols, float thr_f, double thr_d, const uint8_t* occupancy, unsigned long long* cell_best, int arc = 10);
void launch_fast_gather(Launch& L, const unsigned long long* cell_best, int n_cells, const int* level_w, int* out_xy, int* out_level,
                        float* out_score, int* n_out);

// filter.cu
void launch_filter_tracks(Launch& L, const float* ref, const float* cur, const uint8_t* status, int n, double fx, double fy, double cx,
                          double cy, const double* dist5 /* host, 5 doubles or null */, float* out_ref, float* out_cur, double* out_disp, double* out_bearing, int* n_kept);

void launch_score_fundamental(Launch& L, const float* F, int n_hyp, const float* p1, const float* p2, int n, float inv_sigma2,
                              float* scores, uint8_t* inliers);

// lk_generic.cu
void launch_lk_generic(Launch& L, const LKParams& p);
void launch_pair_index(Launch& L, const int* pts_offset_dev, int batch, int n_total, int* pair_idx_dev);
// lk_fast.cu -- returns false when the window size has no specialised kernel
bool launch_lk_fast(Launch& L, const LKParams& p);
bool lk_fast_supported(int win_w, int win_h);
// TMA box shapes of the specialised kernel for this window: {template bytes, derivative ints, search bytes} x rows
struct LkFastBoxes {
    int i_w, i_h, d_w, d_h, j_w, j_h;
};
bool lk_fast_boxes(int win_w, int win_h, LkFastBoxes* b);
bool lk_fast_check_read(unsigned long long out[4]);  // -DDR3LK_CHECKED builds only
bool pyramid_check_read(unsigned long long out[4]);

// ---- checked build (-DDR3LK_CHECKED): compute-sanitizer is not available on the GPU pool, so the kernels can be built with
// their own bounds checks -- LK: every staged rectangle must lie inside the apron-carrying level allocation it is copied
// from, every shared-memory load inside the region it reads (and inside the rows that were staged), every output index
// inside the batch; pyramid kernels: every store (level pixels, mirrored apron pixels, derivatives) inside the destination
// image's allocation.  Violations are counted in device memory, one counter set per translation unit, read and summed by
// dr3lk_debug_check_read; the default build has none of it.
#ifdef DR3LK_CHECKED
static __device__ unsigned long long g_check[4];  // [0] violations, [1] kind of the first, [2] its detail, [3] checks executed
static __device__ __forceinline__ void check_fail(int kind, long long info)
{
    if (atomicAdd(&g_check[0], 1ull) == 0) { g_check[1] = (unsigned long long)kind; g_check[2] = (unsigned long long)info; }
}
#define DR3LK_CHECK(cond, kind, info) do { if (!(cond)) check_fail(kind, (long long)(info)); } while (0)
#define DR3LK_CHECK_COUNT() do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_check[3], 1ull); } while (0)
// reads and resets this translation unit's counters
static inline bool check_read_tu(unsigned long long out[4])
{
    const unsigned long long zero[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out, g_check, sizeof(zero)) != cudaSuccess) return false;
    return cudaMemcpyToSymbol(g_check, zero, sizeof(zero)) == cudaSuccess;
}
#else
#define DR3LK_CHECK(cond, kind, info) do { } while (0)
#define DR3LK_CHECK_COUNT() do { } while (0)
static inline bool check_read_tu(unsigned long long*) { return false; }
#endif

// ---- device helpers ----
// A NaN coordinate: x86 OpenCV's cvFloor turns it into INT_MIN (cvttss2si's "integer indefinite"), which fails every bounds
// test, so the point is reported lost.  F2I on the GPU returns 0 for NaN, which would pass them -- map NaN to a huge
// negative coordinate once, when the point is loaded (infinities and huge values already saturate F2I and fail the tests).
__device__ __forceinline__ float2 sanitize_point(float2 p)
{
    if (!(p.x == p.x)) p.x = -1e30f;
    if (!(p.y == p.y)) p.y = -1e30f;
    return p;
}

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
