// ORBextractor.cc — host shim: ORB_SLAM2::ORBextractor forwarding to libgdslam_cuda (C ABI).
// Replaces GD-SLAM src/ORBextractor.cc:410-470 (ctor tables) and :1043-1105 (operator()).
#include "ORBextractor.h"

#include <cstring>
#include <stdexcept>

#include "gdslam_cuda.h"

namespace ORB_SLAM2 {

static void gd_check(int code, const char* what)
{
    // GD_ECAPACITY = more keypoints than the caller's buffer holds (clamped, like the reference's retainBest would);
    // everything else, including the internal-overflow report GD_EINTERNAL, is an error
    if (code != GD_OK && code != GD_ECAPACITY) throw std::runtime_error(std::string(what) + ": " + gd_last_error());
}

// exact comparison with the previous image (memcmp runs at memory speed: ~15 us for 640x480, no hash collisions to reason about)
static bool same_image(const cv::Mat& m, const std::vector<unsigned char>& last, int last_w, int last_h)
{
    if (m.cols != last_w || m.rows != last_h || last.size() != (size_t)m.cols * m.rows) return false;
    for (int y = 0; y < m.rows; ++y)
        if (std::memcmp(m.ptr<unsigned char>(y), last.data() + (size_t)y * m.cols, (size_t)m.cols) != 0) return false;
    return true;
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST)
{
    // scale tables handed to Frame (src/ORBextractor.cc:415-431): float products of the double scale factor
    mvScaleFactor.assign(nlevels, 1.0f);
    mvLevelSigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; i++) {
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { gd_orb_destroy(handle_); }

void ORBextractor::ensure_handle(int w, int h)
{
    if (handle_ && w <= handle_w_ && h <= handle_h_) return;
    gd_orb_destroy(handle_);
    handle_ = nullptr;
    handle_w_ = w > handle_w_ ? w : handle_w_;
    handle_h_ = h > handle_h_ ? h : handle_h_;
    int rc = gd_orb_create(&handle_, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, handle_w_, handle_h_, device, 1);
    if (rc != GD_OK) throw std::runtime_error(std::string("gd_orb_create: ") + gd_last_error());
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::invalid_argument("ORBextractor: image must be CV_8UC1");  // :1050
    if (memoizeLastImage) {
        if (!last_desc_.empty() && same_image(image, last_image_, last_w_, last_h_)) {
            _keypoints = last_kps_;
            _descriptors.create(last_desc_.rows, 32, CV_8U);
            last_desc_.copyTo(_descriptors.getMat());
            return;
        }
    }
    ensure_handle(image.cols, image.rows);
    const int cap = nfeatures + 4 * nlevels + 8;
    std::vector<gd_keypoint> kps(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    const unsigned char* gray = image.ptr<unsigned char>(0);
    gd_keypoint* kp = kps.data();
    unsigned char* dp = desc.data();
    int n = 0;
    gd_check(gd_orb_extract(handle_, &gray, (size_t)image.step, image.cols, image.rows, &kp, &dp, cap, &n), "gd_orb_extract");
    if (n > cap) n = cap;
    _keypoints.resize(n);
    static_assert(sizeof(gd_keypoint) == 28, "gd_keypoint mirrors cv::KeyPoint");
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint& k = _keypoints[i];
        k.pt.x = kps[i].x;
        k.pt.y = kps[i].y;
        k.size = kps[i].size;
        k.angle = kps[i].angle;
        k.response = kps[i].response;
        k.octave = kps[i].octave;
        k.class_id = kps[i].class_id;
    }
    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr<unsigned char>(i), desc.data() + (size_t)i * 32, 32);
    }
    if (keepImagePyramid) {
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, h = 0;
            gd_check(gd_orb_level_size(handle_, l, &w, &h), "gd_orb_level_size");
            mvImagePyramid[l].create(h, w, CV_8UC1);
            gd_check(gd_orb_fetch_level(handle_, 0, l, mvImagePyramid[l].ptr<unsigned char>(0), (size_t)mvImagePyramid[l].step, &w, &h),
                     "gd_orb_fetch_level");
        }
    }
    if (memoizeLastImage) {
        last_image_.resize((size_t)image.cols * image.rows);
        for (int y = 0; y < image.rows; ++y)
            std::memcpy(last_image_.data() + (size_t)y * image.cols, image.ptr<unsigned char>(y), (size_t)image.cols);
        last_w_ = image.cols;
        last_h_ = image.rows;
        last_kps_ = _keypoints;
        if (n > 0) _descriptors.getMat().copyTo(last_desc_); else last_desc_ = cv::Mat();
    }
}

}  // namespace ORB_SLAM2
