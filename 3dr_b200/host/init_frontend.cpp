// Host-side mirror of the reference's LK call site, written against the C++14 shim (include/dr3lk.hpp).
//
// It restates the front half of init::Init::process_second_frame (reference src/initialization.cpp:587-657): the
// KLT parameters (593-599), the calcOpticalFlowPyrLK call with OPTFLOW_USE_INITIAL_FLOW (608-613), the erase loop over
// !status (615-635, with the disparity norm), and the three DLOG lines (652-654) -- with dr3::calcOpticalFlowPyrLK in
// place of cv::calcOpticalFlowPyrLK, and dr3::create_img_pyramid in place of the Frame constructor's pyramid
// (src/frame.cpp:13-20).  The geometry that follows in the reference (RANSAC F, triangulation) is untouched host code and
// out of scope.  Used by tests/test_host_shim.py to show the drop-in works from compiled C++.
//
// usage: init_frontend <ref.pgm> <cur.pgm> <points.txt> <out.txt>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/dr3lk.hpp"

namespace {

struct Gray {
    std::vector<uint8_t> px;
    int cols = 0, rows = 0;
    dr3::Image view() const { return dr3::Image(px.data(), cols, rows, static_cast<size_t>(cols)); }
};

bool read_pgm(const std::string& path, Gray& g)
{
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv = 0;
    f >> magic >> g.cols >> g.rows >> maxv;
    if (!f || magic != "P5" || maxv != 255) return false;
    f.get();
    g.px.resize(static_cast<size_t>(g.cols) * g.rows);
    f.read(reinterpret_cast<char*>(g.px.data()), static_cast<std::streamsize>(g.px.size()));
    return static_cast<bool>(f);
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc != 5) {
        std::fprintf(stderr, "usage: %s ref.pgm cur.pgm points.txt out.txt\n", argv[0]);
        return 2;
    }
    Gray ref, cur;
    if (!read_pgm(argv[1], ref) || !read_pgm(argv[2], cur)) {
        std::fprintf(stderr, "cannot read the PGM inputs\n");
        return 2;
    }
    std::vector<dr3::Point2f> kps_ref, kps_cur;
    {
        std::ifstream f(argv[3]);
        float x, y;
        while (f >> x >> y) kps_ref.emplace_back(x, y);
    }
    try {
        // Frame constructor: 3-level box pyramid (Config::n_pyr_levels() == 3, src/config.cpp:11)
        dr3::ImgPyramid pyr_ref, pyr_cur;
        dr3::create_img_pyramid(ref.view(), 3, pyr_ref);
        dr3::create_img_pyramid(cur.view(), 3, pyr_cur);

        // process_first_frame: _kps_cur starts as a copy of _kps_ref (src/initialization.cpp:578)
        kps_cur = kps_ref;

        // process_second_frame (src/initialization.cpp:593-613)
        const int klt_win_size = 30;
        const int klt_max_iter = 1000;
        const double klt_eps = 1e-3;
        std::vector<unsigned char> status;
        std::vector<float> error;
        dr3::TermCriteria termcrit(dr3::TermCriteria::COUNT + dr3::TermCriteria::EPS, klt_max_iter, klt_eps);
        dr3::calcOpticalFlowPyrLK(dr3::Image(pyr_ref[0]), dr3::Image(pyr_cur[0]), kps_ref, kps_cur, status, error,
                                  dr3::Size(klt_win_size, klt_win_size), 4, termcrit, dr3::OPTFLOW_USE_INITIAL_FLOW);

        // erase !status, disparities (src/initialization.cpp:615-635)
        std::vector<double> disparities;
        size_t outlier_count = 0;
        auto ref_it = kps_ref.begin();
        auto cur_it = kps_cur.begin();
        for (size_t i = 0; ref_it != kps_ref.end(); ++i) {
            if (!status[i]) {
                ref_it = kps_ref.erase(ref_it);
                cur_it = kps_cur.erase(cur_it);
                ++outlier_count;
                continue;
            }
            // Vector2d(ref.x - cur.x, ref.y - cur.y).norm(): the differences are FLOAT subtractions (cv::Point2f members)
            // promoted to double, then sqrt(x*x + y*y) -- src/initialization.cpp:629-630; same as filter.cu / oracle/postfilter.py
            const float dxf = ref_it->x - cur_it->x, dyf = ref_it->y - cur_it->y;
            const double dx = dxf, dy = dyf;
            disparities.push_back(std::sqrt(dx * dx + dy * dy));
            ++ref_it;
            ++cur_it;
        }
        const double mean_disp = disparities.empty() ? 0.0 : std::accumulate(disparities.begin(), disparities.end(), 0.0) / disparities.size();
        std::printf("Outlier count from optical flow: %zu\n", outlier_count);
        std::printf("Average disparity: %.6fpx\n", mean_disp);
        std::printf("Total matches between ref and cur frames: %zu\n", kps_ref.size());

        std::ofstream out(argv[4]);
        out.precision(9);
        out << outlier_count << " " << kps_ref.size() << " " << mean_disp << "\n";
        out << pyr_ref[1].cols << " " << pyr_ref[1].rows << " " << pyr_ref[2].cols << " " << pyr_ref[2].rows << "\n";
        unsigned long long s1 = 0, s2 = 0;
        for (uint8_t v : *pyr_ref[1].buf) s1 += v;
        for (uint8_t v : *pyr_ref[2].buf) s2 += v;
        out << s1 << " " << s2 << "\n";
        for (size_t i = 0; i < kps_ref.size(); i++) out << kps_ref[i].x << " " << kps_ref[i].y << " " << kps_cur[i].x << " " << kps_cur[i].y << "\n";
    } catch (const dr3::Exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
