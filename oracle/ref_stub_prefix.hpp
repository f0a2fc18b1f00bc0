// oracle/_ref build, part 1 of 3 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference (kvmanohar22/3DR) cannot be built as a whole in this image: every TU includes opencv2/opencv.hpp and
// links OpenCV / Eigen / glog / Sophus / vikit / fast, none of which is installed.  But the functions of the hot path
// that live in the reference's OWN sources only touch a handful of members of cv::Mat / cv::KeyPoint / Eigen::Vector3d.
// oracle/build_ref.sh therefore compiles those functions UNMODIFIED, straight from /root/reference (the lines are piped
// from where they lie into g++'s stdin, nothing is copied into this repository), between this prefix -- minimal
// stand-ins for exactly the members they use -- and the C glue of ref_stub_suffix.hpp:
//
//   src/utils.cpp:282-430           shi_tomasi_score, halfSampleSSE2, reduce_to_half, create_img_pyramid
//   src/initialization.cpp:171-249  InitHelper::CheckFundamental
//   src/camera.cpp:25-41            Pinhole::cam2world(u, v)
//
// What is the reference's and what is ours: every arithmetic statement of those functions is the reference's.  Ours
// (and therefore NOT pinned by this build) are the stand-ins below: cv::Mat storage (16-byte aligned, like cv::Mat's
// allocator, so that reduce_to_half takes its SSE2 branch exactly when it would in the reference), Eigen's
// Vector3d::normalized() (restated from Eigen 3.3 Redux.h / Dot.h: squaredNorm of a 3-vector is (x*x + y*y) + z*z with
// SSE2 packets, then v / sqrt(z)), and cv::undistortPoints (restated from OpenCV's cvUndistortPointsInternal, default
// 5 fixed-point iterations; pinned separately against the cv2 wheel by tests/golden/make_golden.py).
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <utility>
#include <vector>

#if __SSE2__
#include <emmintrin.h>
#endif

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC2 13

namespace cv {

struct MatStep {
    size_t p[2];
};

// the members of cv::Mat the extracted functions touch: rows, cols, data, step.p[0], type(), at<float>(r, c), and the
// (rows, cols, type) / (rows, cols, type, void*) constructors; copies are shallow like cv::Mat's
struct Mat {
    int rows, cols;
    uint8_t* data;
    MatStep step;
    int type_;
    std::shared_ptr<void> owner;
    Mat() : rows(0), cols(0), data(nullptr), type_(CV_8U) { step.p[0] = step.p[1] = 0; }
    static size_t elem_size(int type) { return type == CV_8U ? 1 : (type == CV_32F ? 4 : 8); }
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type)
    {
        step.p[1] = elem_size(type);
        step.p[0] = (size_t)c * step.p[1];
        void* p = nullptr;
        if (posix_memalign(&p, 64, step.p[0] * (size_t)r + 64) != 0) abort();  // cv::Mat: CV_MALLOC_ALIGN = 64
        owner.reset(p, free);
        data = (uint8_t*)p;
    }
    Mat(int r, int c, int type, void* ext, size_t row_step = 0) : rows(r), cols(c), data((uint8_t*)ext), type_(type)
    {
        step.p[1] = elem_size(type);
        step.p[0] = row_step ? row_step : (size_t)c * step.p[1];
    }
    int type() const { return type_; }
    template <typename T>
    const T& at(int r, int c) const { return *(const T*)(data + step.p[0] * r + sizeof(T) * c); }
    template <typename T>
    T& at(int r, int c) { return *(T*)(data + step.p[0] * r + sizeof(T) * c); }
};

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
    Point2f pt;
    float size;
    KeyPoint() : size(0) {}
    KeyPoint(Point2f p, float s) : pt(p), size(s) {}
};

// OpenCV calib3d/imgproc undistortPoints for a 1x1 CV_32FC2 point, float 3x3 K, float 1x5 D, no R / P, default criteria
// (COUNT 5): a restatement (cvUndistortPointsInternal), see the header comment.
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D)
{
    const double fx = K.at<float>(0, 0), fy = K.at<float>(1, 1), cx = K.at<float>(0, 2), cy = K.at<float>(1, 2);
    const double ifx = 1. / fx, ify = 1. / fy;
    double k[5];
    for (int i = 0; i < 5; i++) k[i] = D.at<float>(0, i);
    const float* s = (const float*)src.data;
    float* d = (float*)dst.data;
    double x = s[0], y = s[1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {  // test: undistortPoints regression 14583
            x = (u - cx) * ifx;
            y = (v - cy) * ify;
            break;
        }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0 * r2 + 0 * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    d[0] = (float)x;
    d[1] = (float)y;
}

}  // namespace cv

// Eigen::Vector3d stand-in: operator[] and normalized() (see the header comment for the summation order)
struct Vector3d {
    double v[3];
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
    Vector3d normalized() const
    {
        const double z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
        Vector3d r = *this;
        if (z > 0.0) {
            const double n = sqrt(z);
            r.v[0] = v[0] / n;
            r.v[1] = v[1] / n;
            r.v[2] = v[2] / n;
        }
        return r;
    }
};

using std::vector;

namespace utils {
typedef std::vector<cv::Mat> ImgPyramid;  // include/global.hpp:33
}

namespace dr3 {

// include/camera.hpp: the members Pinhole::cam2world(u, v) reads
struct Pinhole {
    double _fx, _fy, _cx, _cy;
    bool _distortion;
    cv::Mat _cvK, _cvD;
    Vector3d cam2world(const double& u, const double& v) const;
};

namespace init {

// include/svo/initialization.hpp: the members InitHelper::CheckFundamental reads
struct InitHelper {
    typedef std::pair<int, int> Match;
    vector<cv::KeyPoint> mvKeys1, mvKeys2;
    vector<Match> mvMatches12;
    float CheckFundamental(const cv::Mat& F21, vector<bool>& vbMatchesInliers, float sigma);
};

}  // namespace init
}  // namespace dr3
