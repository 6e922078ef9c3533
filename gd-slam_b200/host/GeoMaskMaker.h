// GeoMaskMaker.h — drop-in replacement of GD-SLAM's include/GeoMaskMaker.h (class GeoMaskMaker, :52-116).
//
// Same class name, constructor and the three entry points Tracking.cc uses (src/Tracking.cc:137,242,245):
//   GeoMaskMaker(Mat K, Mat DistCoef, float DepthMapFactor)
//   void AddNewImage(Mat rgb, Mat depth, Mat label, Mat originlabel)
//   void GetNoGMMmask(Mat& mask)
// Everything behind them runs on the GPU through the C ABI of include/gdslam_cuda.h; the prototype code of the
// reference header (graph / segmentation experiments that Tracking never calls) is not carried over.
// GetRt() — the pose between the buffered pair (src/GeoMaskMaker.cc:77-156) — runs its feature / matching half on the GPU
// (gd_geomask_getrt_points: cv::ORB features cached per ring slot, cross-check matching, first-100 selection, undistortPoints,
// depth look-up, back-projection) and calls cv::solvePnPRansac + cv::Rodrigues on the result, exactly like the reference
// (:143-150).  Define GD_SHIM_NO_OPENCV_GETRT to build without OpenCV's calib3d (then install a pose provider with
// SetPoseProvider(), e.g. Tracking's own pose, or every mask is all ones).  See INTEGRATION.md.
#ifndef GEOMASKMAKER_H_
#define GEOMASKMAKER_H_

#include <functional>
#include <opencv2/opencv.hpp>
#include <vector>

struct gd_geomask;

class GeoMaskMaker {
public:
    int inter_frame_size = 5;  // frames between the pair (t-5, t); fixed by the device ring
    cv::Mat _inst_param;       // K (3x3 CV_32F)
    cv::Mat _DistCoefParam;    // k1 k2 p1 p2 [k3]
    float _DepthMapFactor;
    bool start_flag = false;   // true once six frames were pushed
    int mimage_height;
    int mimage_width;
    int image_count = 0;

    GeoMaskMaker(cv::Mat inst_param, cv::Mat DistCoef, float DepthMapFactor);
    // same as above with an explicit image size / device (the reference hard-codes 640x480)
    GeoMaskMaker(cv::Mat inst_param, cv::Mat DistCoef, float DepthMapFactor, int width, int height, int device = 0);
    ~GeoMaskMaker();
    GeoMaskMaker(const GeoMaskMaker&) = delete;
    GeoMaskMaker& operator=(const GeoMaskMaker&) = delete;

    void AddNewImage(cv::Mat new_RGB, cv::Mat new_Depth, cv::Mat label, cv::Mat originlabel);
    void GetNoGMMmask(cv::Mat& mask);
    bool GetRt(cv::Mat& R, cv::Mat& T);  // GPU points + cv::solvePnPRansac; see the header comment

    // optional: replaces GetRt() as the source of (R, T); return false for "no pose" (all-ones mask).  Install it before the
    // first AddNewImage: with a provider the cv::ORB features of GetRt are not computed at all.
    typedef std::function<bool(cv::Mat& R, cv::Mat& T)> PoseProvider;
    void SetPoseProvider(PoseProvider p) { pose_provider_ = p; }

    // debug access to the intermediate products of the last GetNoGMMmask (CV_32FC2 flow, CV_8UC1 edges)
    void GetFlow(cv::Mat& flow);
    cv::Mat GetEdge(cv::Mat arg_Depth_image);
    float depth2std(float depth);

private:
    gd_geomask* handle_ = nullptr;
    PoseProvider pose_provider_;
    int pushed_ = 0;
    int device_ = 0;
    bool getrt_enabled_ = false;
    void init(int width, int height, int device);
};

#endif  // GEOMASKMAKER_H_
