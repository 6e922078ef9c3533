// ORBextractor.h — drop-in replacement of GD-SLAM's include/ORBextractor.h (ORB_SLAM2::ORBextractor, :45-111).
// Same constructor, operator() and getters that Frame.cc / Tracking.cc use (src/Tracking.cc:108, src/Frame.cc:419-425,
// Frame ctor scale tables); the extraction itself runs on the GPU through include/gdslam_cuda.h.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <opencv2/opencv.hpp>
#include <vector>

struct gd_orb;

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Mask is ignored, like in the reference (include/ORBextractor.h:58).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Only the stereo path reads this (src/Frame.cc:645,735,752).  Filled after every extraction when
    // keepImagePyramid is true (one extra D2H of ~0.95 Mpx); left empty otherwise.
    std::vector<cv::Mat> mvImagePyramid;
    bool keepImagePyramid = false;
    // GrabImageRGBD_GD extracts twice from the identical gray image (src/Tracking.cc:238,252): serve the second
    // call from the first one's result (keyed on the exact image content: memcmp with a kept copy).
    bool memoizeLastImage = true;
    // CUDA device of the extractor (set before the first call; the reference has no such notion)
    int device = 0;

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

private:
    gd_orb* handle_ = nullptr;
    int handle_w_ = 0, handle_h_ = 0;
    std::vector<unsigned char> last_image_;  // copy of the last extracted image (memo key: exact content)
    int last_w_ = 0, last_h_ = 0;
    std::vector<cv::KeyPoint> last_kps_;
    cv::Mat last_desc_;
    void ensure_handle(int w, int h);
};

}  // namespace ORB_SLAM2

#endif  // ORBEXTRACTOR_H
