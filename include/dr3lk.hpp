// dr3lk.hpp -- C++14 host shim over the C ABI (dr3lk.h) that keeps the reference's call surface for this path.
//
//   dr3::calcOpticalFlowPyrLK(...)   same arguments, defaults, output sizing and error behaviour as
//                                    cv::calcOpticalFlowPyrLK, which the reference calls at src/initialization.cpp:608-613
//   dr3::create_img_pyramid(...)     utils::create_img_pyramid of the reference (include/utils.hpp:67, src/utils.cpp:421-430)
//
// No OpenCV headers are needed: images are passed as dr3::Image views (data / cols / rows / step -- exactly the
// cv::Mat fields, see dr3::view()), points as any 2-float struct (cv::Point2f works unchanged).
// Header-only; link with libdr3lk.so.  There is no CPU fallback: without a CUDA device the first call throws.
#ifndef DR3LK_HPP_
#define DR3LK_HPP_

#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "dr3lk.h"

namespace dr3 {

struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct Point2f {
    float x, y;
    Point2f(float x_ = 0.f, float y_ = 0.f) : x(x_), y(y_) {}
};
struct TermCriteria {
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria(int t = COUNT + EPS, int c = 30, double e = 0.01) : type(t), maxCount(c), epsilon(e) {}
};
enum { OPTFLOW_USE_INITIAL_FLOW = DR3LK_USE_INITIAL_FLOW, OPTFLOW_LK_GET_MIN_EIGENVALS = DR3LK_GET_MIN_EIGENVALS };

// Thrown where OpenCV would throw cv::Exception (CV_Assert failures) and on CUDA errors.
class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string& m) : std::runtime_error("dr3lk error " + std::to_string(c) + ": " + m), code(c) {}
};

// Non-owning view of a CV_8UC1 image.
struct Image {
    const uint8_t* data;
    int cols, rows;
    size_t step;
    Image() : data(nullptr), cols(0), rows(0), step(0) {}
    Image(const uint8_t* d, int c, int r, size_t s) : data(d), cols(c), rows(r), step(s) {}
    bool empty() const { return data == nullptr || cols <= 0 || rows <= 0; }
};
// view(cv::Mat) -- anything with data / cols / rows / step (size_t-convertible) members.
template <class Mat>
inline Image view(const Mat& m) { return Image(reinterpret_cast<const uint8_t*>(m.data), m.cols, m.rows, static_cast<size_t>(m.step)); }

// Owning continuous CV_8UC1 image (the levels create_img_pyramid allocates, like `cv::Mat(rows, cols, CV_8U)`).
struct OwnedImage {
    std::shared_ptr<std::vector<uint8_t>> buf;  // null for level 0: that level aliases the caller's buffer (shallow copy)
    const uint8_t* data;
    int cols, rows;
    size_t step;
    OwnedImage() : data(nullptr), cols(0), rows(0), step(0) {}
    operator Image() const { return Image(data, cols, rows, step); }
};
typedef std::vector<OwnedImage> ImgPyramid;  // reference: typedef std::vector<cv::Mat> ImgPyramid (include/global.hpp:33)

// One dr3lk_ctx per host thread (the reference calls the path from a single thread, src/handler.cpp:31-48).
class Context {
public:
    explicit Context(int device = 0) : ctx_(nullptr)
    {
        const int rc = dr3lk_create(&ctx_, device);
        if (rc != DR3LK_OK) throw Exception(rc, dr3lk_last_error(nullptr));
    }
    ~Context() { dr3lk_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    dr3lk_ctx* get() const { return ctx_; }
    void check(int rc) const
    {
        if (rc != DR3LK_OK) throw Exception(rc, dr3lk_last_error(ctx_));
    }
    // The context the free functions use when none is passed: one per calling thread, on the CUDA device named by the
    // environment variable DR3LK_DEVICE (default 0) -- the reference has no notion of a device, so the selection has to
    // live outside its call surface.
    static Context& thread_default()
    {
        static thread_local Context c(default_device());
        return c;
    }
    static int default_device()
    {
        const char* e = std::getenv("DR3LK_DEVICE");
        return e && *e ? std::atoi(e) : 0;
    }

private:
    dr3lk_ctx* ctx_;
};

// cv::calcOpticalFlowPyrLK.  P2f: any struct of two floats (x, y), e.g. cv::Point2f or dr3::Point2f.
template <class P2f>
inline void calcOpticalFlowPyrLK(const Image& prevImg, const Image& nextImg, const std::vector<P2f>& prevPts, std::vector<P2f>& nextPts,
                                 std::vector<unsigned char>& status, std::vector<float>& err, Size winSize = Size(21, 21), int maxLevel = 3,
                                 TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                 double minEigThreshold = 1e-4, Context* context = nullptr)
{
    static_assert(sizeof(P2f) == 2 * sizeof(float), "points must be two packed floats (x, y)");
    Context& c = context ? *context : Context::thread_default();
    if (maxLevel < 0 || winSize.width <= 2 || winSize.height <= 2)
        throw Exception(DR3LK_E_ARG, "(-215:Assertion failed) maxLevel >= 0 && winSize.width > 2 && winSize.height > 2");
    const size_t n = prevPts.size();
    if (n == 0) {  // OpenCV: nextPts / status / err are released
        nextPts.clear(); status.clear(); err.clear();
        return;
    }
    if (prevImg.empty() || nextImg.empty() || prevImg.cols != nextImg.cols || prevImg.rows != nextImg.rows)
        throw Exception(DR3LK_E_SIZE, "(-215:Assertion failed) prevPyr[level * lvlStep1].size() == nextPyr[level * lvlStep2].size()");
    if (flags & OPTFLOW_USE_INITIAL_FLOW) {
        if (nextPts.size() != n)
            throw Exception(DR3LK_E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints");
    } else {
        nextPts.resize(n);
    }
    status.resize(n);
    err.resize(n);
    c.check(dr3lk_calc_optical_flow_pyr_lk(c.get(), prevImg.data, prevImg.step, nextImg.data, nextImg.step, prevImg.cols, prevImg.rows,
                                           reinterpret_cast<const float*>(prevPts.data()), reinterpret_cast<float*>(nextPts.data()),
                                           status.data(), err.data(), static_cast<int>(n), winSize.width, winSize.height, maxLevel,
                                           criteria.type, criteria.maxCount, criteria.epsilon, flags, minEigThreshold));
}

// utils::create_img_pyramid(img_lvl_0, n_levels, pyr): pyr[0] aliases the input, pyr[i] is rows/2 x cols/2 of pyr[i-1].
// `mode` selects the rounding of utils::reduce_to_half; the default reproduces the reference on x86.
inline void create_img_pyramid(const Image& img_lvl_0, int n_levels, ImgPyramid& pyr, int mode = DR3LK_BOX_AUTO_X86, Context* context = nullptr)
{
    Context& c = context ? *context : Context::thread_default();
    pyr.assign(static_cast<size_t>(n_levels > 0 ? n_levels : 0), OwnedImage());
    if (n_levels <= 0) return;
    pyr[0].data = img_lvl_0.data; pyr[0].cols = img_lvl_0.cols; pyr[0].rows = img_lvl_0.rows; pyr[0].step = img_lvl_0.step;
    std::vector<uint8_t*> outs;
    int w = img_lvl_0.cols, h = img_lvl_0.rows;
    for (int l = 1; l < n_levels; l++) {
        w /= 2; h /= 2;
        pyr[l].buf = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(w > 0 ? w : 0) * static_cast<size_t>(h > 0 ? h : 0));
        pyr[l].data = pyr[l].buf->data(); pyr[l].cols = w; pyr[l].rows = h; pyr[l].step = static_cast<size_t>(w);
        outs.push_back(pyr[l].buf->data());
    }
    c.check(dr3lk_box_pyramid(c.get(), img_lvl_0.data, img_lvl_0.cols, img_lvl_0.rows, img_lvl_0.step, n_levels, outs.data(), mode));
}

// ---- a frame's LK pyramid built once on the device and reused across calls ----
class Pyramid {
public:
    Pyramid(const Image& img, Size winSize = Size(21, 21), int maxLevel = 3, Context* context = nullptr)
        : ctx_(context ? context : &Context::thread_default()), pyr_(nullptr), win_(winSize)
    {
        ctx_->check(dr3lk_pyramid_create(ctx_->get(), img.data, img.cols, img.rows, img.step, winSize.width, winSize.height, maxLevel, &pyr_));
    }
    // adopts a pyramid returned by the library (trackFrame)
    Pyramid(Context& context, dr3lk_pyramid* pyr, Size winSize) : ctx_(&context), pyr_(pyr), win_(winSize) {}
    ~Pyramid() { dr3lk_pyramid_destroy(pyr_); }
    Pyramid(const Pyramid&) = delete;
    Pyramid& operator=(const Pyramid&) = delete;
    const dr3lk_pyramid* get() const { return pyr_; }
    Context& context() const { return *ctx_; }
    Size winSize() const { return win_; }
    int levels() const { return dr3lk_pyramid_levels(pyr_); }

private:
    Context* ctx_;
    dr3lk_pyramid* pyr_;
    Size win_;
};

// calcOpticalFlowPyrLK on prebuilt pyramids (OpenCV's vector<Mat> pyramid input form); the window is the pyramids'.
template <class P2f>
inline void calcOpticalFlowPyrLK(const Pyramid& prevPyr, const Pyramid& nextPyr, const std::vector<P2f>& prevPts, std::vector<P2f>& nextPts,
                                 std::vector<unsigned char>& status, std::vector<float>& err, int maxLevel = 3,
                                 TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01), int flags = 0,
                                 double minEigThreshold = 1e-4)
{
    static_assert(sizeof(P2f) == 2 * sizeof(float), "points must be two packed floats (x, y)");
    Context& c = prevPyr.context();
    const size_t n = prevPts.size();
    if (n == 0) { nextPts.clear(); status.clear(); err.clear(); return; }
    if (flags & OPTFLOW_USE_INITIAL_FLOW) {
        if (nextPts.size() != n) throw Exception(DR3LK_E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints");
    } else {
        nextPts.resize(n);
    }
    status.resize(n);
    err.resize(n);
    const Size win = prevPyr.winSize();
    c.check(dr3lk_calc_optical_flow_pyr_lk_cached(c.get(), prevPyr.get(), nextPyr.get(), reinterpret_cast<const float*>(prevPts.data()),
                                                  reinterpret_cast<float*>(nextPts.data()), status.data(), err.data(), static_cast<int>(n),
                                                  win.width, win.height, maxLevel, criteria.type, criteria.maxCount, criteria.epsilon, flags,
                                                  minEigThreshold));
}

// Streaming form for the per-frame loop (src/handler.cpp:31-48): track from a frame that is already on the device into a
// NEW image in one call (one upload, the new frame's pyramid, LK, one download).  keepNext: 0 discard the new frame's
// pyramid, 1 keep its Gaussian levels, 2 keep it with derivatives (the previous frame of the next call).
template <class P2f>
inline std::unique_ptr<Pyramid> trackFrame(const Pyramid& prevPyr, const Image& nextImg, const std::vector<P2f>& prevPts,
                                           std::vector<P2f>& nextPts, std::vector<unsigned char>& status, std::vector<float>& err,
                                           int maxLevel = 3, TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01),
                                           int flags = 0, double minEigThreshold = 1e-4, int keepNext = 2)
{
    static_assert(sizeof(P2f) == 2 * sizeof(float), "points must be two packed floats (x, y)");
    Context& c = prevPyr.context();
    const size_t n = prevPts.size();
    if (flags & OPTFLOW_USE_INITIAL_FLOW) {
        if (nextPts.size() != n) throw Exception(DR3LK_E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints");
    } else {
        nextPts.resize(n);
    }
    status.resize(n);
    err.resize(n);
    const Size win = prevPyr.winSize();
    dr3lk_pyramid* out = nullptr;
    c.check(dr3lk_track_frame(c.get(), prevPyr.get(), nextImg.data, nextImg.step, reinterpret_cast<const float*>(prevPts.data()),
                              reinterpret_cast<float*>(nextPts.data()), status.data(), err.data(), static_cast<int>(n), win.width, win.height,
                              maxLevel, criteria.type, criteria.maxCount, criteria.epsilon, flags, minEigThreshold, keepNext,
                              keepNext ? &out : nullptr));
    return std::unique_ptr<Pyramid>(out ? new Pyramid(c, out, win) : nullptr);
}

// ---- the erase-by-status loop of src/initialization.cpp:615-635 in one call ----
// Removes the lost points from ref/cur (order kept), fills the disparities and the unit bearing vectors of the current
// points (3 doubles each; pass fx = 0 to skip them) -- Pinhole::cam2world, src/camera.cpp:25-41: `distortion` = the
// camera's d0..d4 (nullptr or |d0| <= 1e-7: the undistorted branch).
template <class P2f>
inline void filterTracks(std::vector<P2f>& kpsRef, std::vector<P2f>& kpsCur, const std::vector<unsigned char>& status,
                         std::vector<double>& disparities, std::vector<double>& bearings, double fx = 0, double fy = 0, double cx = 0,
                         double cy = 0, Context* context = nullptr, const double* distortion = nullptr)
{
    static_assert(sizeof(P2f) == 2 * sizeof(float), "points must be two packed floats (x, y)");
    Context& c = context ? *context : Context::thread_default();
    const int n = static_cast<int>(kpsRef.size());
    std::vector<P2f> r(kpsRef.size()), q(kpsRef.size());
    disparities.assign(kpsRef.size(), 0.0);
    bearings.assign(fx != 0 ? 3 * kpsRef.size() : 0, 0.0);
    int kept = 0;
    c.check(dr3lk_filter_tracks(c.get(), reinterpret_cast<const float*>(kpsRef.data()), reinterpret_cast<const float*>(kpsCur.data()),
                                status.data(), n, fx != 0 ? fx : 1.0, fy != 0 ? fy : 1.0, cx, cy, distortion, reinterpret_cast<float*>(r.data()),
                                reinterpret_cast<float*>(q.data()), disparities.data(), fx != 0 ? bearings.data() : nullptr, &kept));
    r.resize(kept); q.resize(kept); disparities.resize(kept);
    if (fx != 0) bearings.resize(3 * static_cast<size_t>(kept));
    kpsRef.swap(r); kpsCur.swap(q);
}

// ---- feature_detection::FastDetector::detect (src/features.cpp:43-98) ----
struct Corner {  // reference include/features.hpp:66-77
    int x, y, level;
    float score;
};
// FAST-10 corners of the frame's box pyramid, one per grid cell (best Shi-Tomasi score above detection_threshold), in
// cell order.  Defaults are the reference's Config values (src/config.cpp:9-12) and FAST threshold (features.cpp:57).
inline std::vector<Corner> fastDetect(const Image& img, int n_pyr_levels = 3, int cell_size = 30, double detection_threshold = 20.0,
                                      int fast_threshold = 20, const std::vector<unsigned char>* grid_occupancy = nullptr,
                                      int box_mode = DR3LK_BOX_AUTO_X86, Context* context = nullptr)
{
    Context& c = context ? *context : Context::thread_default();
    const size_t cells = static_cast<size_t>((img.cols + cell_size - 1) / cell_size) * static_cast<size_t>((img.rows + cell_size - 1) / cell_size);
    std::vector<int> xy(2 * cells), lv(cells);
    std::vector<float> sc(cells);
    int n = 0;
    c.check(dr3lk_fast_detect(c.get(), img.data, img.cols, img.rows, img.step, n_pyr_levels, cell_size, fast_threshold, detection_threshold,
                              box_mode, grid_occupancy ? grid_occupancy->data() : nullptr, xy.data(), lv.data(), sc.data(), &n));
    std::vector<Corner> out(static_cast<size_t>(n));
    for (int i = 0; i < n; i++) out[i] = Corner{xy[2 * i], xy[2 * i + 1], lv[i], sc[i]};
    return out;
}

}  // namespace dr3
#endif  // DR3LK_HPP_
