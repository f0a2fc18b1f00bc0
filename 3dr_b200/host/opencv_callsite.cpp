// The reference's LK call site with the reference's own types, against include/dr3lk_opencv.hpp: the statements below are
// those of init::Init::process_second_frame (reference src/initialization.cpp:593-635) and of the Frame constructor
// (src/frame.cpp:13-20) with ONE change each -- the namespace of the callee:
//     cv::calcOpticalFlowPyrLK    -> dr3::calcOpticalFlowPyrLK
//     utils::create_img_pyramid   -> dr3::utils::create_img_pyramid
// Built by tests/test_host_shim.py with -I tests/mock_opencv (a stand-in for the few cv:: types involved: this image has
// no OpenCV headers) and compared with the output of init_frontend (the raw-pointer shim) on the same inputs.
//
// usage: opencv_callsite <ref.pgm> <cur.pgm> <points.txt> <out.txt>
#include <cstdio>
#include <cmath>
#include <fstream>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/dr3lk_opencv.hpp"

#ifndef DR3LK_HAVE_OPENCV
#error "opencv2/core.hpp (or the test stand-in) must be on the include path"
#endif

using std::vector;

static bool read_pgm(const std::string& path, cv::Mat& m)
{
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxv = 0;
    f >> magic >> w >> h >> maxv;
    if (!f || magic != "P5" || maxv != 255) return false;
    f.get();
    m = cv::Mat(h, w, CV_8UC1);
    f.read(reinterpret_cast<char*>(m.data), static_cast<std::streamsize>(w) * h);
    return static_cast<bool>(f);
}

int main(int argc, char** argv)
{
    if (argc != 5) return 2;
    cv::Mat img_ref, img_cur;
    if (!read_pgm(argv[1], img_ref) || !read_pgm(argv[2], img_cur)) return 2;
    vector<cv::Point2f> _kps_ref, _kps_cur;
    {
        std::ifstream f(argv[3]);
        float x, y;
        while (f >> x >> y) _kps_ref.emplace_back(x, y);
    }
    try {
        // Frame::Frame (src/frame.cpp:18): utils::create_img_pyramid(img, Config::n_pyr_levels(), _img_pyr)
        dr3::utils::ImgPyramid ref_img_pyr, cur_img_pyr;
        dr3::utils::create_img_pyramid(img_ref, 3, ref_img_pyr);
        dr3::utils::create_img_pyramid(img_cur, 3, cur_img_pyr);
        _kps_cur = _kps_ref;  // process_first_frame, src/initialization.cpp:578

        // ---- src/initialization.cpp:593-613, verbatim except for the callee's namespace ----
        const int klt_win_size = 30;
        const int klt_max_iter = 1000;
        const double klt_eps = 1e-3;
        vector<uchar> status;
        vector<float> error;
        cv::TermCriteria termcrit(cv::TermCriteria::COUNT+cv::TermCriteria::EPS,
                                  klt_max_iter, klt_eps);
        dr3::calcOpticalFlowPyrLK(ref_img_pyr[0],
                                  cur_img_pyr[0],
                                  _kps_ref, _kps_cur,
                                  status, error,
                                  cv::Size2i(klt_win_size, klt_win_size),
                                  4, termcrit, cv::OPTFLOW_USE_INITIAL_FLOW);

        // ---- src/initialization.cpp:615-635 (erase loop, disparities) ----
        auto kps_ref_itr = _kps_ref.begin();
        auto kps_cur_itr = _kps_cur.begin();
        vector<double> _disparities;
        size_t outlier_count = 0;
        for (size_t i = 0; kps_ref_itr != _kps_ref.end(); ++i) {
            if (!status[i]) {
                kps_ref_itr = _kps_ref.erase(kps_ref_itr);
                kps_cur_itr = _kps_cur.erase(kps_cur_itr);
                ++outlier_count;
                continue;
            }
            const double dx = kps_ref_itr->x - kps_cur_itr->x, dy = kps_ref_itr->y - kps_cur_itr->y;  // float differences
            _disparities.push_back(std::sqrt(dx * dx + dy * dy));
            ++kps_ref_itr;
            ++kps_cur_itr;
        }
        const double mean_disp = _disparities.empty() ? 0.0 : std::accumulate(_disparities.begin(), _disparities.end(), 0.0) / _disparities.size();
        std::ofstream out(argv[4]);
        out.precision(9);
        out << outlier_count << " " << _kps_ref.size() << " " << mean_disp << "\n";
        out << ref_img_pyr[1].cols << " " << ref_img_pyr[1].rows << " " << ref_img_pyr[2].cols << " " << ref_img_pyr[2].rows << "\n";
        unsigned long long s1 = 0, s2 = 0;
        for (int i = 0; i < ref_img_pyr[1].rows * ref_img_pyr[1].cols; i++) s1 += ref_img_pyr[1].data[i];
        for (int i = 0; i < ref_img_pyr[2].rows * ref_img_pyr[2].cols; i++) s2 += ref_img_pyr[2].data[i];
        out << s1 << " " << s2 << "\n";
        for (size_t i = 0; i < _kps_ref.size(); i++) out << _kps_ref[i].x << " " << _kps_ref[i].y << " " << _kps_cur[i].x << " " << _kps_cur[i].y << "\n";

        // the argument checks are OpenCV's: cv::Exception, same message
        bool threw = false;
        try {
            vector<cv::Point2f> a(3), b;
            dr3::calcOpticalFlowPyrLK(img_ref, img_cur, a, b, status, error, cv::Size(2, 2));
        } catch (const cv::Exception& e) {
            threw = std::string(e.what()).find("winSize.width > 2") != std::string::npos;
        }
        if (!threw) return 3;
        // zero points: outputs released
        vector<cv::Point2f> none, nout(5);
        dr3::calcOpticalFlowPyrLK(img_ref, img_cur, none, nout, status, error);
        if (!nout.empty() || !status.empty() || !error.empty()) return 4;
        // err not needed: cv::noArray()
        vector<cv::Point2f> one(1, cv::Point2f(100.f, 100.f)), onext;
        dr3::calcOpticalFlowPyrLK(img_ref, img_cur, one, onext, status, cv::noArray());
        if (onext.size() != 1 || status.size() != 1) return 5;
    } catch (const cv::Exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
