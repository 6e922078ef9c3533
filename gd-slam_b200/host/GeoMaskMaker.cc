// GeoMaskMaker.cc — host shim: the reference's GeoMaskMaker interface forwarding to libgdslam_cuda (C ABI).
// Replaces GD-SLAM src/GeoMaskMaker.cc:39-76 (ctor), :158-429 (GetFlow / GetNoGMMmask / AddNewImage), :854-964 (GetEdge).
#include "GeoMaskMaker.h"

#include <cstring>
#include <iostream>
#include <stdexcept>

#include "gdslam_cuda.h"

namespace {
void gd_check(int code, const char* what)
{
    if (code != GD_OK) throw std::runtime_error(std::string(what) + ": " + gd_last_error());  // no CPU fallback
}
cv::Mat ones_mask(int h, int w)
{
    cv::Mat m(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) std::memset(m.ptr(y), 1, (size_t)w);
    return m;
}
}  // namespace

GeoMaskMaker::GeoMaskMaker(cv::Mat inst_param, cv::Mat DistCoef, float DepthMapFactor)
    : GeoMaskMaker(inst_param, DistCoef, DepthMapFactor, 640, 480, 0)  // GeoMaskMaker.cc:54-55
{
}

GeoMaskMaker::GeoMaskMaker(cv::Mat inst_param, cv::Mat DistCoef, float DepthMapFactor, int width, int height, int device)
{
    inst_param.copyTo(_inst_param);
    DistCoef.copyTo(_DistCoefParam);
    _DepthMapFactor = DepthMapFactor;
    init(width, height, device);
}

void GeoMaskMaker::init(int width, int height, int device)
{
    mimage_width = width;
    mimage_height = height;
    device_ = device;
    float K[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K[3 * r + c] = _inst_param.at<float>(r, c);
    float D[5] = {0, 0, 0, 0, 0};
    const int nd = _DistCoefParam.empty() ? 0 : std::min(5, _DistCoefParam.rows * _DistCoefParam.cols);
    for (int i = 0; i < nd; ++i) D[i] = _DistCoefParam.at<float>(i);
    gd_check(gd_geomask_create(&handle_, K, nd ? D : nullptr, nd, _DepthMapFactor, width, height, device, 1), "gd_geomask_create");
}

GeoMaskMaker::~GeoMaskMaker() { gd_geomask_destroy(handle_); }

void GeoMaskMaker::AddNewImage(cv::Mat new_Image, cv::Mat new_Depth, cv::Mat /*label*/, cv::Mat /*originlabel*/)
{
    if (new_Image.type() != CV_8UC3 || new_Depth.type() != CV_32FC1 || new_Image.cols != mimage_width ||
        new_Image.rows != mimage_height || new_Depth.cols != mimage_width || new_Depth.rows != mimage_height)
        throw std::invalid_argument("GeoMaskMaker::AddNewImage: expects CV_8UC3 + CV_32FC1 (metres) of the configured size");
    const uint8_t* bgr = new_Image.ptr<uint8_t>(0);
    const float* dep = new_Depth.ptr<float>(0);
#ifndef GD_SHIM_NO_OPENCV_GETRT
    // GetRt() is the pose source unless a provider was installed: its cv::ORB features are computed once per pushed frame on
    // the GPU and kept with the frame in the device ring (the reference re-extracts both images on every call, :82-90)
    if (!pose_provider_ && !getrt_enabled_) {
        gd_check(gd_geomask_enable_getrt(handle_), "gd_geomask_enable_getrt");
        getrt_enabled_ = true;
    }
#endif
    gd_check(gd_geomask_push(handle_, &bgr, (size_t)new_Image.step, &dep, (size_t)new_Depth.step), "gd_geomask_push");
    if (++pushed_ > inter_frame_size) start_flag = true;  // :419-428 (the device ring keeps the six frames)
}

void GeoMaskMaker::GetNoGMMmask(cv::Mat& mask)
{
    std::cout << image_count << "checking" << std::endl;  // GeoMaskMaker.cc:169
    image_count += 1;
    if (!start_flag) {
        mask = ones_mask(mimage_height, mimage_width);  // :171-175
        return;
    }
    cv::Mat R, T;
    const bool ok = pose_provider_ ? pose_provider_(R, T) : GetRt(R, T);
    float Rf[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Tf[3] = {0, 0, 0};
    int valid = ok ? 1 : 0;
    if (ok) {
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rf[3 * r + c] = R.at<float>(r, c);
            Tf[r] = T.at<float>(r, 0);
        }
    } else {
        std::cout << "small feature match." << std::endl;  // :183
    }
    mask.create(mimage_height, mimage_width, CV_8UC1);
    uint8_t* out = mask.ptr<uint8_t>(0);
    gd_check(gd_geomask_mask(handle_, Rf, Tf, &valid, &out, (size_t)mask.step), "gd_geomask_mask");
}

void GeoMaskMaker::GetFlow(cv::Mat& flow)
{
    flow.create(mimage_height, mimage_width, CV_32FC2);
    gd_check(gd_geomask_debug_fetch(handle_, GD_DBG_FLOW, 0, flow.ptr<float>(0), (size_t)mimage_height * mimage_width * 8), "flow");
}

cv::Mat GeoMaskMaker::GetEdge(cv::Mat arg_Depth_image)
{
    cv::Mat e(arg_Depth_image.rows, arg_Depth_image.cols, CV_8UC1);
    cv::Mat d;
    arg_Depth_image.copyTo(d);  // dense rows
    float K[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K[3 * r + c] = _inst_param.at<float>(r, c);
    gd_check(gd_stage_depth_edge(device_, d.ptr<float>(0), d.cols, d.rows, K, e.ptr<uint8_t>(0)), "gd_stage_depth_edge");
    return e;
}

float GeoMaskMaker::depth2std(float depth)
{
    const float fu = _inst_param.at<float>(0, 0), sigma_norm = 0.5f;  // GeoMaskMaker.cc:1386-1391
    return (1 / fu) * (1 / fu) * sigma_norm * sigma_norm * depth * depth * depth * depth;
}

#ifdef GD_SHIM_NO_OPENCV_GETRT
// Build without OpenCV's calib3d (deployments that feed Tracking's pose through SetPoseProvider): no pose of our own -> the
// all-ones mask path of GeoMaskMaker.cc:179-185.
bool GeoMaskMaker::GetRt(cv::Mat&, cv::Mat&) { return false; }
#else
// GeoMaskMaker.cc:77-156.  Everything up to the solver input (:82-141) comes from the GPU in one call; the solver is the
// reference's own: cv::solvePnPRansac with its default arguments and cv::Rodrigues, then the two convertTo (:148-152).
bool GeoMaskMaker::GetRt(cv::Mat& R, cv::Mat& T)
{
    if (!getrt_enabled_) return false;  // frames were pushed while a pose provider was installed: no features to match
    float obj[100 * 3], pix[100 * 2];
    float* po = obj;
    float* pp = pix;
    int n = 0;
    gd_check(gd_geomask_getrt_points(handle_, &po, &pp, &n), "gd_geomask_getrt_points");
    if (n < 20) return false;  // :143-146
    std::vector<cv::Point3f> objectPoints((size_t)n);
    std::vector<cv::Point2f> imagePixels((size_t)n);
    for (int i = 0; i < n; ++i) {
        objectPoints[i].x = obj[3 * i];
        objectPoints[i].y = obj[3 * i + 1];
        objectPoints[i].z = obj[3 * i + 2];
        imagePixels[i].x = pix[2 * i];
        imagePixels[i].y = pix[2 * i + 1];
    }
    cv::Mat rvec;
    cv::solvePnPRansac(objectPoints, imagePixels, _inst_param, _DistCoefParam, rvec, T);  // :148
    cv::Rodrigues(rvec, R);                                                               // :149
    T.convertTo(T, CV_32FC1);
    R.convertTo(R, CV_32FC1);
    return true;
}
#endif
