// tests/mock_opencv/opencv2/core.hpp -- TEST INFRASTRUCTURE.  A minimal stand-in for the handful of OpenCV core types that
// include/dr3lk_opencv.hpp and the reference's LK call site (src/initialization.cpp:593-635) touch, so that the OpenCV-typed
// shim can be compiled and exercised in an image without OpenCV headers.  Semantics follow OpenCV where the shim relies on
// them (shallow Mat copies, OutputArray::create resizing the wrapped std::vector, needed(), checkVector); everything else is
// absent.  This is NOT OpenCV and nothing in the product includes it.
#ifndef MOCK_OPENCV_CORE_HPP_
#define MOCK_OPENCV_CORE_HPP_
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

typedef unsigned char uchar;

namespace cv {

namespace Error { enum Code { StsError = -2, StsAssert = -215 }; }

class Exception : public std::runtime_error {
public:
    Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
    int code;
};
#define CV_Error(code, msg) throw cv::Exception(code, std::string(msg))
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(cv::Error::StsAssert, "(-215:Assertion failed) " #expr); } while (0)

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;
typedef Size Size2i;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };

struct MatStep {
    size_t p[2];
    operator size_t() const { return p[0]; }
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;
    Mat() : rows(0), cols(0), data(nullptr), type_(0) { step.p[0] = step.p[1] = 0; }
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type)
    {
        step.p[1] = elemSize(); step.p[0] = (size_t)c * step.p[1];
        void* p = nullptr;
        if (posix_memalign(&p, 64, step.p[0] * (size_t)(r > 0 ? r : 0) + 64) != 0) throw std::bad_alloc();
        owner_.reset(p, free);
        data = (uchar*)p;
    }
    Mat(int r, int c, int type, void* ext, size_t row_step = 0) : rows(r), cols(c), data((uchar*)ext), type_(type)
    {
        step.p[1] = elemSize(); step.p[0] = row_step ? row_step : (size_t)c * step.p[1];
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == CV_8U ? 1 : 4); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    bool isContinuous() const { return rows <= 1 || step.p[0] == (size_t)cols * elemSize(); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + step.p[0] * r); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + step.p[0] * r); }
    // number of elemChannels-channel elements of a vector-like matrix (N x 1 / 1 x N with that many channels, or N x elemChannels)
    int checkVector(int elemChannels, int depth_ = -1, bool requireContinuous = true) const
    {
        if (depth_ >= 0 && depth() != depth_) return -1;
        if (requireContinuous && !isContinuous()) return -1;
        if ((rows == 1 || cols == 1 || rows * cols == 0) && channels() == elemChannels) return rows * cols;
        if (channels() == 1 && cols == elemChannels) return rows;
        return -1;
    }

private:
    int type_;
    std::shared_ptr<void> owner_;
};

// the std::vector kinds the reference hands to calcOpticalFlowPyrLK, plus Mat
class _InputArray {
public:
    enum Kind { NONE, MAT, VEC_P2F, VEC_U8, VEC_F32 };
    _InputArray() : kind_(NONE), obj_(nullptr) {}
    _InputArray(const Mat& m) : kind_(MAT), obj_((void*)&m) {}
    _InputArray(const std::vector<Point2f>& v) : kind_(VEC_P2F), obj_((void*)&v) {}
    _InputArray(const std::vector<uchar>& v) : kind_(VEC_U8), obj_((void*)&v) {}
    _InputArray(const std::vector<float>& v) : kind_(VEC_F32), obj_((void*)&v) {}
    Mat getMat() const
    {
        switch (kind_) {
        case MAT: return *(const Mat*)obj_;
        // an empty std::vector<T> is an empty matrix that still has T's type (so checkVector() returns 0, not -1)
        case VEC_P2F: { auto& v = *(std::vector<Point2f>*)obj_; return Mat((int)v.size(), 1, CV_32FC2, v.data()); }
        case VEC_U8: { auto& v = *(std::vector<uchar>*)obj_; return Mat((int)v.size(), 1, CV_8UC1, v.data()); }
        case VEC_F32: { auto& v = *(std::vector<float>*)obj_; return Mat((int)v.size(), 1, CV_32FC1, v.data()); }
        default: return Mat();
        }
    }
    bool needed() const { return kind_ != NONE; }

protected:
    Kind kind_;
    void* obj_;
};

class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(std::vector<Point2f>& v) : _InputArray(v) {}
    _OutputArray(std::vector<uchar>& v) : _InputArray(v) {}
    _OutputArray(std::vector<float>& v) : _InputArray(v) {}
    void create(int rows, int cols, int type, int = -1, bool = false) const
    {
        const size_t n = (size_t)rows * cols;
        switch (kind_) {
        case MAT: { Mat& m = *(Mat*)obj_; if (m.rows != rows || m.cols != cols || m.type() != type) m = Mat(rows, cols, type); break; }
        case VEC_P2F: CV_Assert(type == CV_32FC2); ((std::vector<Point2f>*)obj_)->resize(n); break;
        case VEC_U8: CV_Assert(type == CV_8UC1); ((std::vector<uchar>*)obj_)->resize(n); break;
        case VEC_F32: CV_Assert(type == CV_32FC1); ((std::vector<float>*)obj_)->resize(n); break;
        default: break;
        }
    }
    void create(Size sz, int type, int i = -1, bool t = false) const { create(sz.height, sz.width, type, i, t); }
    void release() const
    {
        switch (kind_) {
        case MAT: *(Mat*)obj_ = Mat(); break;
        case VEC_P2F: ((std::vector<Point2f>*)obj_)->clear(); break;
        case VEC_U8: ((std::vector<uchar>*)obj_)->clear(); break;
        case VEC_F32: ((std::vector<float>*)obj_)->clear(); break;
        default: break;
        }
    }
};
typedef _OutputArray _InputOutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _InputOutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static const _OutputArray none; return none; }

}  // namespace cv
#endif
