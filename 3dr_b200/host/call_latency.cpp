// Per-call latency of the reference's call pattern as a compiled C++ caller sees it (no Python binding in the way):
// dr3::calcOpticalFlowPyrLK(prev, next, ...) on one frame pair, once with the frames in ordinary (pageable) memory and once
// in page-locked memory (dr3lk_host_alloc; rows at the aligned pitch and continuous), and the frame-to-frame
// form with the previous frame's pyramid kept on the device (dr3::Pyramid + dr3lk_track_frame).  bench.py --workload kitti
// runs it and puts the numbers into the `latency` block next to the ones measured through the Python binding.
//
// With more frames on the command line it also times the frame-to-frame chain over all of them (every
// frame uploaded once, its pyramid built once and used first as the next and then as the previous frame, surviving points
// carried forward), frames pageable and pinned.
//
// usage: call_latency <prev.pgm> <next.pgm> <points.txt> [calls [more frames.pgm ...]]      -> one JSON object on stdout
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/dr3lk.hpp"

namespace {

struct Gray {
    std::vector<uint8_t> px;
    int cols = 0, rows = 0;
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

template <class F>
double median_us(F&& call, int n)
{
    for (int i = 0; i < 20; i++) call();
    std::vector<double> t(static_cast<size_t>(n));
    for (int i = 0; i < n; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        call();
        t[static_cast<size_t>(i)] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s prev.pgm next.pgm points.txt [calls]\n", argv[0]);
        return 2;
    }
    Gray a, b;
    if (!read_pgm(argv[1], a) || !read_pgm(argv[2], b) || a.cols != b.cols || a.rows != b.rows) {
        std::fprintf(stderr, "cannot read the PGM inputs\n");
        return 2;
    }
    std::vector<dr3::Point2f> prev_pts;
    {
        std::ifstream f(argv[3]);
        float x, y;
        while (f >> x >> y) prev_pts.emplace_back(x, y);
    }
    const int calls = argc > 4 ? std::atoi(argv[4]) : 300;
    const int w = a.cols, h = a.rows;
    try {
        std::vector<dr3::Point2f> next_pts, ref_pts;
        std::vector<unsigned char> status, ref_status;
        std::vector<float> err;
        const dr3::Image pa(a.px.data(), w, h, static_cast<size_t>(w)), pb(b.px.data(), w, h, static_cast<size_t>(w));
        const double pageable = median_us([&] { dr3::calcOpticalFlowPyrLK(pa, pb, prev_pts, next_pts, status, err); }, calls);
        ref_pts = next_pts; ref_status = status;

        // the same two frames in page-locked memory, rows at the device pitch
        const size_t step = (static_cast<size_t>(w) + 15) / 16 * 16;
        uint8_t* ma = static_cast<uint8_t*>(dr3lk_host_alloc(step * h));
        uint8_t* mb = static_cast<uint8_t*>(dr3lk_host_alloc(step * h));
        if (!ma || !mb) throw dr3::Exception(DR3LK_E_CUDA, "dr3lk_host_alloc failed");
        for (int y = 0; y < h; y++) {
            std::memcpy(ma + y * step, a.px.data() + static_cast<size_t>(y) * w, static_cast<size_t>(w));
            std::memcpy(mb + y * step, b.px.data() + static_cast<size_t>(y) * w, static_cast<size_t>(w));
        }
        const dr3::Image qa(ma, w, h, step), qb(mb, w, h, step);
        const double pinned = median_us([&] { dr3::calcOpticalFlowPyrLK(qa, qb, prev_pts, next_pts, status, err); }, calls);
        bool same = next_pts.size() == ref_pts.size() && status == ref_status &&
                    std::memcmp(next_pts.data(), ref_pts.data(), ref_pts.size() * sizeof(dr3::Point2f)) == 0;

        // ... and as CONTINUOUS page-locked images (row step == width: what cv::Mat(rows, cols, CV_8U, pinned_ptr) is)
        uint8_t* ca = static_cast<uint8_t*>(dr3lk_host_alloc(a.px.size()));
        uint8_t* cb = static_cast<uint8_t*>(dr3lk_host_alloc(b.px.size()));
        if (!ca || !cb) throw dr3::Exception(DR3LK_E_CUDA, "dr3lk_host_alloc failed");
        std::memcpy(ca, a.px.data(), a.px.size());
        std::memcpy(cb, b.px.data(), b.px.size());
        const dr3::Image ra(ca, w, h, static_cast<size_t>(w)), rb(cb, w, h, static_cast<size_t>(w));
        const double pinned_cont = median_us([&] { dr3::calcOpticalFlowPyrLK(ra, rb, prev_pts, next_pts, status, err); }, calls);
        same = same && status == ref_status && std::memcmp(next_pts.data(), ref_pts.data(), ref_pts.size() * sizeof(dr3::Point2f)) == 0;
        dr3lk_host_free(ca);
        dr3lk_host_free(cb);

        // frame-to-frame form: the previous frame's pyramid is on the device, one upload per call
        dr3::Context& ctx = dr3::Context::thread_default();
        dr3::Pyramid prev(qa, dr3::Size(21, 21), 3, &ctx);
        const double streaming = median_us([&] {
            status.resize(prev_pts.size()); err.resize(prev_pts.size()); next_pts.resize(prev_pts.size());
            dr3lk_pyramid* keep = nullptr;
            const int rc = dr3lk_track_frame(ctx.get(), prev.get(), mb, step, reinterpret_cast<const float*>(prev_pts.data()),
                                             reinterpret_cast<float*>(next_pts.data()), status.data(), err.data(), static_cast<int>(prev_pts.size()),
                                             21, 21, 3, 3, 30, 0.01, 0, 1e-4, 0, &keep);
            if (rc != DR3LK_OK) throw dr3::Exception(rc, dr3lk_last_error(ctx.get()));
        }, calls);
        const bool same2 = status == ref_status && std::memcmp(next_pts.data(), ref_pts.data(), ref_pts.size() * sizeof(dr3::Point2f)) == 0;
        // the chain over all frames given: pyramid of frame 0, then one dr3lk_track_frame per new frame
        double chain_pageable_ms = 0., chain_pinned_ms = 0.;
        size_t survivors[2] = {0, 0};
        const int n_frames = argc > 5 ? 2 + (argc - 5) : 0;
        if (n_frames > 2) {
            std::vector<Gray> fr(static_cast<size_t>(n_frames));
            fr[0] = a; fr[1] = b;
            for (int i = 2; i < n_frames; i++)
                if (!read_pgm(argv[3 + i], fr[static_cast<size_t>(i)]) || fr[static_cast<size_t>(i)].cols != w || fr[static_cast<size_t>(i)].rows != h)
                    throw dr3::Exception(DR3LK_E_ARG, "cannot read a chain frame");
            std::vector<uint8_t*> pin(static_cast<size_t>(n_frames));
            for (int i = 0; i < n_frames; i++) {
                pin[static_cast<size_t>(i)] = static_cast<uint8_t*>(dr3lk_host_alloc(step * h));
                if (!pin[static_cast<size_t>(i)]) throw dr3::Exception(DR3LK_E_CUDA, "dr3lk_host_alloc failed");
                for (int y = 0; y < h; y++)
                    std::memcpy(pin[static_cast<size_t>(i)] + y * step, fr[static_cast<size_t>(i)].px.data() + static_cast<size_t>(y) * w, static_cast<size_t>(w));
            }
            for (int kind = 0; kind < 2; kind++) {
                auto chain = [&] {
                    const uint8_t* f0 = kind ? pin[0] : fr[0].px.data();
                    const size_t st = kind ? step : static_cast<size_t>(w);
                    dr3lk_pyramid* cur_pyr = nullptr;
                    ctx.check(dr3lk_pyramid_create(ctx.get(), f0, w, h, st, 21, 21, 3, &cur_pyr));
                    std::vector<dr3::Point2f> cur = prev_pts, nxt;
                    for (int i = 1; i < n_frames; i++) {
                        nxt.resize(cur.size()); status.resize(cur.size()); err.resize(cur.size());
                        dr3lk_pyramid* keep = nullptr;
                        const uint8_t* fi = kind ? pin[static_cast<size_t>(i)] : fr[static_cast<size_t>(i)].px.data();
                        const int rc = dr3lk_track_frame(ctx.get(), cur_pyr, fi, st, reinterpret_cast<const float*>(cur.data()),
                                                         reinterpret_cast<float*>(nxt.data()), status.data(), err.data(), static_cast<int>(cur.size()), 21, 21, 3,
                                                         3, 30, 0.01, 0, 1e-4, 2, &keep);
                        dr3lk_pyramid_destroy(cur_pyr);
                        cur_pyr = keep;
                        if (rc != DR3LK_OK) throw dr3::Exception(rc, dr3lk_last_error(ctx.get()));
                        size_t k = 0;
                        for (size_t j = 0; j < cur.size(); j++)
                            if (status[j]) nxt[k++] = nxt[j];
                        nxt.resize(k);
                        cur.swap(nxt);
                    }
                    dr3lk_pyramid_destroy(cur_pyr);
                    survivors[kind] = cur.size();
                };
                (kind ? chain_pinned_ms : chain_pageable_ms) = median_us(chain, std::max(10, calls / 10)) / 1e3;
            }
            for (uint8_t* m : pin) dr3lk_host_free(m);
        }
        dr3lk_host_free(ma);
        dr3lk_host_free(mb);
        const bool ok = same && same2 && survivors[0] == survivors[1];
        std::printf("{\"points\": %zu, \"calls\": %d, \"c_abi_call_us_pageable\": %.2f, \"c_abi_call_us_pinned\": %.2f, "
                    "\"c_abi_call_us_pinned_continuous\": %.2f, \"c_abi_track_frame_us_pinned\": %.2f, ", prev_pts.size(), calls, pageable, pinned,
                    pinned_cont, streaming);
        if (n_frames > 2)
            std::printf("\"chain_frames\": %d, \"chain_ms_pageable\": %.4f, \"chain_ms_pinned\": %.4f, \"chain_survivors\": %zu, ", n_frames,
                        chain_pageable_ms, chain_pinned_ms, survivors[1]);
        std::printf("\"identical_results\": %s}\n", ok ? "true" : "false");
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
