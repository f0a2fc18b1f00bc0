// oracle/_ref build, part 3 of 3 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C glue over the reference functions compiled between ref_stub_prefix.hpp and this file (see oracle/build_ref.sh).
// Nothing here computes: each entry point marshals plain buffers into the stand-in types and calls the reference.

extern "C" {

// utils::create_img_pyramid (src/utils.cpp:421-430) on a continuous w x h image.  The image is first copied into a
// fresh (64-byte aligned, continuous) cv::Mat, which is what cv::imread hands the Frame constructor (src/frame.cpp:13-20),
// so reduce_to_half takes its SSE2 branch exactly when it would in the reference (cols % 16 == 0).  out[l - 1] receives
// level l (rows/2 x cols/2 of the level above, continuous).  Returns 0, or -1 for the shapes on which the reference's
// scalar pointer walk leaves its buffers (e.g. odd cols together with odd rows) -- those are not run.
int ref_box_pyramid(const uint8_t* img, int w, int h, int n_levels, uint8_t* const* out)
{
    if (!img || w < 2 || h < 2 || n_levels < 1) return -2;
    for (int l = 0, cw = w, ch = h; l + 1 < n_levels; l++, cw /= 2, ch /= 2) {
        if (cw < 2 || ch < 2) return -2;
        if (cw % 16 == 0) continue;  // SSE2 branch: plain row pairs
        // the reference's scalar walk (src/utils.cpp:401-418) has no bounds check of its own; simulate its pointers and
        // refuse the shapes on which it would write more than rows/2 rows or read past the input
        long long top = 0, bottom = cw, rows = 0;
        const long long end = (long long)cw * ch;
        while (bottom < end) {
            if (bottom + 2 * (cw / 2) - 1 >= end) return -1;
            top += 2 * (cw / 2) + cw;
            bottom += 2 * (cw / 2) + cw;
            rows++;
        }
        if (rows > ch / 2) return -1;
    }
    cv::Mat lvl0(h, w, CV_8U);
    for (int y = 0; y < h; y++) memcpy(lvl0.data + (size_t)y * lvl0.step.p[0], img + (size_t)y * w, (size_t)w);
    utils::ImgPyramid pyr;
    utils::create_img_pyramid(lvl0, n_levels, pyr);
    for (int l = 1; l < n_levels; l++)
        for (int y = 0; y < pyr[l].rows; y++)
            memcpy(out[l - 1] + (size_t)y * pyr[l].cols, pyr[l].data + (size_t)y * pyr[l].step.p[0], (size_t)pyr[l].cols);
    return 0;
}

// utils::shi_tomasi_score (src/utils.cpp:282-321) at n integer positions of one image
void ref_shi_tomasi(const uint8_t* img, int w, int h, long step, const int* uv, int n, float* out)
{
    const cv::Mat m(h, w, CV_8UC1, (void*)img, (size_t)step);
    for (int i = 0; i < n; i++) out[i] = utils::shi_tomasi_score(m, uv[2 * i], uv[2 * i + 1]);
}

// InitHelper::CheckFundamental (src/initialization.cpp:171-249): matches are (i, i) over p1 / p2 as built at
// src/initialization.cpp:638-650; F is row-major 3x3 float.  inliers: n bytes.
float ref_check_fundamental(const float* F, const float* p1, const float* p2, int n, float sigma, uint8_t* inliers)
{
    dr3::init::InitHelper h;
    h.mvKeys1.reserve(n); h.mvKeys2.reserve(n); h.mvMatches12.reserve(n);
    for (int i = 0; i < n; i++) {
        h.mvKeys1.emplace_back(cv::KeyPoint(cv::Point2f(p1[2 * i], p1[2 * i + 1]), 2.0f));
        h.mvKeys2.emplace_back(cv::KeyPoint(cv::Point2f(p2[2 * i], p2[2 * i + 1]), 2.0f));
        h.mvMatches12.emplace_back(std::make_pair(i, i));
    }
    float Fm[9];
    memcpy(Fm, F, sizeof(Fm));
    const cv::Mat F21(3, 3, CV_32F, Fm);
    vector<bool> inl;
    const float score = h.CheckFundamental(F21, inl, sigma);
    for (int i = 0; i < n; i++) inliers[i] = inl[i] ? 1 : 0;
    return score;
}

// Pinhole::cam2world(u, v) (src/camera.cpp:25-41) for n pixels (uv as doubles, as Feature::px holds them); the members
// are set up as the constructor does (src/camera.cpp:8-21: _distortion = fabs(d0) > 1e-7, float _cvK / _cvD).
void ref_cam2world(double fx, double fy, double cx, double cy, const double* d, const double* uv, int n, double* out)
{
    dr3::Pinhole cam;
    cam._fx = fx; cam._fy = fy; cam._cx = cx; cam._cy = cy;
    cam._distortion = fabs(d[0]) > 1e-7;
    float K[9] = {(float)fx, 0.f, (float)cx, 0.f, (float)fy, (float)cy, 0.f, 0.f, 1.f};
    float D[5] = {(float)d[0], (float)d[1], (float)d[2], (float)d[3], (float)d[4]};
    cam._cvK = cv::Mat(3, 3, CV_32F, K);
    cam._cvD = cv::Mat(1, 5, CV_32F, D);
    for (int i = 0; i < n; i++) {
        const Vector3d b = cam.cam2world(uv[2 * i], uv[2 * i + 1]);
        out[3 * i] = b[0]; out[3 * i + 1] = b[1]; out[3 * i + 2] = b[2];
    }
}

}  // extern "C"
