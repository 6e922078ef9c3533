// shim_demo.cc — drives the drop-in C++ classes exactly like Tracking::GrabImageRGBD_GD does (src/Tracking.cc:238-252):
// ORB on the new gray image (twice, second call memoised), AddNewImage, GetNoGMMmask.  Reads a raw sequence file written by
// tests/test_gpu_shim.py, writes masks / keypoints / descriptors for the test to compare with the oracle.
//   file: int32 w, h, nframes; per frame: bgr (w*h*3 u8), gray (w*h u8), depth (w*h f32), R (9 f32), T (3 f32)
// Third argument "getrt": no pose provider — GeoMaskMaker::GetRt() itself runs (GPU points + the solver).  The image has no
// OpenCV SDK, so the stand-in solver below answers with the pose of the sequence file and logs the points it was given; the
// test compares those with gd_getrt_points (and the real cv2.solvePnPRansac is exercised through the C ABI's pose hook).
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "../GeoMaskMaker.h"
#include "../ORBextractor.h"

static float g_R[9], g_T[3];
static std::vector<float> g_points;  // n, then n x 3 object points, n x 2 image pixels (of the last solver call)

static bool demo_solver(const std::vector<cv::Point3f>& obj, const std::vector<cv::Point2f>& pix, cv::Mat& rvec, cv::Mat& tvec)
{
    g_points.clear();
    g_points.push_back((float)obj.size());
    for (const auto& p : obj) { g_points.push_back(p.x); g_points.push_back(p.y); g_points.push_back(p.z); }
    for (const auto& p : pix) { g_points.push_back(p.x); g_points.push_back(p.y); }
    // rotation matrix of the file -> rotation vector (what solvePnPRansac returns), f64 like OpenCV's outputs
    const double tr = (double)g_R[0] + g_R[4] + g_R[8];
    double c = (tr - 1) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    const double th = std::acos(c);
    double ax[3] = {(double)g_R[7] - g_R[5], (double)g_R[2] - g_R[6], (double)g_R[3] - g_R[1]};
    const double s2 = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    rvec.create(3, 1, CV_64FC1);
    tvec.create(3, 1, CV_64FC1);
    for (int i = 0; i < 3; ++i) {
        rvec.at<double>(i, 0) = s2 > 1e-12 ? ax[i] / s2 * th : 0.0;
        tvec.at<double>(i, 0) = (double)g_T[i];
    }
    return true;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    const bool use_getrt = argc > 3 && std::string(argv[3]) == "getrt";
    FILE* f = std::fopen(argv[1], "rb");
    FILE* o = std::fopen(argv[2], "wb");
    if (!f || !o) return 3;
    int hdr[3];
    if (std::fread(hdr, 4, 3, f) != 3) return 4;
    const int w = hdr[0], h = hdr[1], nf = hdr[2];
    float Kv[9] = {535.4f * w / 640, 0, 320.1f * w / 640, 0, 539.2f * w / 640, 247.6f * w / 640, 0, 0, 1};
    cv::Mat K(3, 3, CV_32FC1, Kv), D(4, 1, CV_32FC1);
    for (int i = 0; i < 4; ++i) D.at<float>(i) = 0.f;
    float* Rv = g_R;
    float* Tv = g_T;
    GeoMaskMaker gm(K, D, 5000.f, w, h, 0);
    if (use_getrt)
        cv::pnp_stand_in() = demo_solver;
    else
        gm.SetPoseProvider([&](cv::Mat& R, cv::Mat& T) {
            R.create(3, 3, CV_32FC1);
            T.create(3, 1, CV_32FC1);
            for (int i = 0; i < 9; ++i) R.at<float>(i / 3, i % 3) = Rv[i];
            for (int i = 0; i < 3; ++i) T.at<float>(i, 0) = Tv[i];
            return true;
        });
    ORB_SLAM2::ORBextractor orb(1500, 1.2f, 8, 20, 7);
    std::vector<unsigned char> bgr((size_t)w * h * 3), gray((size_t)w * h);
    std::vector<float> depth((size_t)w * h);
    double t_orb = 0, t_add = 0, t_mask = 0, t_orb2 = 0;  // steady-state (frame >= 6) wall time per call, ms
    int n_timed = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    for (int i = 0; i < nf; ++i) {
        if (std::fread(bgr.data(), 1, bgr.size(), f) != bgr.size() || std::fread(gray.data(), 1, gray.size(), f) != gray.size() ||
            std::fread(depth.data(), 4, depth.size(), f) != depth.size() || std::fread(Rv, 4, 9, f) != 9 || std::fread(Tv, 4, 3, f) != 3)
            return 5;
        cv::Mat im(h, w, CV_8UC3, bgr.data()), g(h, w, CV_8UC1, gray.data()), d(h, w, CV_32FC1, depth.data()), label, mask;
        std::vector<cv::KeyPoint> kps, kps2;
        cv::Mat desc, desc2;
        const auto t0 = now();
        orb(cv::_InputArray(g), cv::_InputArray(label), kps, cv::_OutputArray(desc));    // Frame(), Tracking.cc:238
        const auto t1 = now();
        gm.AddNewImage(im, d, label, label);                                              // Tracking.cc:242
        const auto t2 = now();
        g_points.clear();
        gm.GetNoGMMmask(mask);                                                            // Tracking.cc:245
        const auto t3 = now();
        orb(cv::_InputArray(g), cv::_InputArray(label), kps2, cv::_OutputArray(desc2));  // Frame(), Tracking.cc:252
        const auto t4 = now();
        if (i >= 6) {
            t_orb += ms(t0, t1); t_add += ms(t1, t2); t_mask += ms(t2, t3); t_orb2 += ms(t3, t4);
            ++n_timed;
        }
        if (kps2.size() != kps.size()) return 6;
        int n = (int)kps.size();
        std::fwrite(&n, 4, 1, o);
        for (int k = 0; k < n; ++k) {
            float rec[5] = {kps[k].pt.x, kps[k].pt.y, kps[k].size, kps[k].angle, kps[k].response};
            std::fwrite(rec, 4, 5, o);
            std::fwrite(&kps[k].octave, 4, 1, o);
            std::fwrite(desc.ptr<unsigned char>(k), 1, 32, o);
        }
        for (int y = 0; y < h; ++y) std::fwrite(mask.ptr<unsigned char>(y), 1, (size_t)w, o);
        if (use_getrt) {  // the solver's input of this frame (empty while GetRt is not called / returns early)
            if (g_points.empty()) g_points.push_back(0.f);
            std::fwrite(g_points.data(), 4, g_points.size(), o);
        }
    }
    std::fclose(f);
    std::fclose(o);
    if (n_timed > 0)
        std::fprintf(stderr, "shim steady state over %d frames, ms per call: ORBextractor() %.3f, AddNewImage %.3f, GetNoGMMmask %.3f, "
                             "ORBextractor() again (memo) %.3f, frame total %.3f\n",
                     n_timed, t_orb / n_timed, t_add / n_timed, t_mask / n_timed, t_orb2 / n_timed,
                     (t_orb + t_add + t_mask + t_orb2) / n_timed);
    return 0;
}
