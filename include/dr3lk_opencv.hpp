// dr3lk_opencv.hpp -- the reference's two call surfaces with the reference's OWN argument types (cv::Mat, cv::InputArray,
// std::vector<cv::Point2f>): with this header the swap at the call site is a namespace change,
//
//     cv::calcOpticalFlowPyrLK(...)      ->  dr3::calcOpticalFlowPyrLK(...)        (src/initialization.cpp:608-613)
//     utils::create_img_pyramid(...)     ->  dr3::utils::create_img_pyramid(...)   (src/frame.cpp:18, include/utils.hpp:67)
//
// and nothing else in init::Init::process_second_frame or the Frame constructor changes.  Signatures, default arguments,
// output sizing (status / err / nextPts are created like OpenCV creates them, `err` only when it is needed, outputs
// released for zero points) and the error type (cv::Exception through CV_Error / CV_Assert) are OpenCV's.
//
// Needs OpenCV's core headers, so it is only active when they are on the include path.  NOTE: this build image has no
// OpenCV C++ headers; the header is compiled and run in tests/test_host_shim.py against a minimal stand-in for the cv::
// types it touches (tests/mock_opencv/, test infrastructure) -- it has not been compiled against a real OpenCV tree.
#ifndef DR3LK_OPENCV_HPP_
#define DR3LK_OPENCV_HPP_

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#define DR3LK_HAVE_OPENCV 1
#endif
#endif

#ifdef DR3LK_HAVE_OPENCV
#include <opencv2/core.hpp>

#include <vector>

#include "dr3lk.hpp"

namespace dr3 {

// cv::calcOpticalFlowPyrLK (opencv2/video/tracking.hpp) for 8-bit single-channel images -- what the reference passes
inline void calcOpticalFlowPyrLK(cv::InputArray _prevImg, cv::InputArray _nextImg, cv::InputArray _prevPts, cv::InputOutputArray _nextPts,
                                 cv::OutputArray _status, cv::OutputArray _err, cv::Size winSize = cv::Size(21, 21), int maxLevel = 3,
                                 cv::TermCriteria criteria = cv::TermCriteria(cv::TermCriteria::COUNT + cv::TermCriteria::EPS, 30, 0.01),
                                 int flags = 0, double minEigThreshold = 1e-4)
{
    const cv::Mat prevImg = _prevImg.getMat(), nextImg = _nextImg.getMat(), prevPtsMat = _prevPts.getMat();
    CV_Assert(maxLevel >= 0 && winSize.width > 2 && winSize.height > 2);
    CV_Assert(prevImg.type() == CV_8UC1 && nextImg.type() == CV_8UC1);  // the LK path of the reference: 8-bit gray Frame images
    const int npoints = prevPtsMat.checkVector(2, CV_32F, true);
    CV_Assert(npoints >= 0);
    if (npoints == 0) {
        _nextPts.release();
        _status.release();
        _err.release();
        return;
    }
    CV_Assert(prevImg.size() == nextImg.size());
    if (!(flags & OPTFLOW_USE_INITIAL_FLOW)) _nextPts.create(prevPtsMat.size(), prevPtsMat.type(), -1, true);
    cv::Mat nextPtsMat = _nextPts.getMat();
    CV_Assert(nextPtsMat.checkVector(2, CV_32F, true) == npoints);
    _status.create(npoints, 1, CV_8U, -1, true);
    cv::Mat statusMat = _status.getMat();
    CV_Assert(statusMat.isContinuous());
    cv::Mat errMat;
    float* err = nullptr;
    if (_err.needed()) {
        _err.create(npoints, 1, CV_32F, -1, true);
        errMat = _err.getMat();
        CV_Assert(errMat.isContinuous());
        err = errMat.ptr<float>();
    }
    Context& c = Context::thread_default();
    const int rc = dr3lk_calc_optical_flow_pyr_lk(c.get(), prevImg.data, prevImg.step, nextImg.data, nextImg.step, prevImg.cols, prevImg.rows,
                                                  prevPtsMat.ptr<float>(), nextPtsMat.ptr<float>(), statusMat.data, err, npoints,
                                                  winSize.width, winSize.height, maxLevel, criteria.type, criteria.maxCount,
                                                  criteria.epsilon, flags, minEigThreshold);
    if (rc != DR3LK_OK) CV_Error(cv::Error::StsError, dr3lk_last_error(c.get()));
}

namespace utils {

typedef std::vector<cv::Mat> ImgPyramid;  // reference include/global.hpp:33

// utils::create_img_pyramid (src/utils.cpp:421-430): pyr[0] shares img_lvl_0's buffer, the other levels are fresh CV_8U
// matrices of rows/2 x cols/2, filled by the device version of reduce_to_half (same rounding path selection as on x86)
inline void create_img_pyramid(const cv::Mat& img_lvl_0, int n_levels, ImgPyramid& pyr)
{
    pyr.resize(n_levels);
    if (n_levels <= 0) return;
    pyr[0] = img_lvl_0;
    std::vector<uint8_t*> outs;
    for (int ii = 1; ii < n_levels; ++ii) {
        pyr[ii] = cv::Mat(pyr[ii - 1].rows / 2, pyr[ii - 1].cols / 2, CV_8U);
        outs.push_back(pyr[ii].data);
    }
    Context& c = Context::thread_default();
    const int rc = dr3lk_box_pyramid(c.get(), img_lvl_0.data, img_lvl_0.cols, img_lvl_0.rows, img_lvl_0.step, n_levels, outs.data(),
                                     DR3LK_BOX_AUTO_X86);
    if (rc != DR3LK_OK) CV_Error(cv::Error::StsError, dr3lk_last_error(c.get()));
}

}  // namespace utils
}  // namespace dr3

#endif  // DR3LK_HAVE_OPENCV
#endif  // DR3LK_OPENCV_HPP_
