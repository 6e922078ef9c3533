// Minimal cv:: stand-in used ONLY to compile-check and exercise the host shim in an image without the OpenCV SDK.
// A real deployment compiles gd-slam_b200/host/*.cc against the OpenCV headers GD-SLAM already uses.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include <cmath>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {
struct Point2f {
    float x = 0, y = 0;
};
struct Point3f {
    float x = 0, y = 0, z = 0;
};
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void* ext, size_t st = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(t)
    {
        step = st ? st : (size_t)c * elem();
    }
    void create(int r, int c, int t)
    {
        if (data && rows == r && cols == c && type_ == t) return;
        type_ = t;
        rows = r;
        cols = c;
        step = (size_t)c * elem();
        buf_ = std::make_shared<std::vector<unsigned char>>(step * r);
        data = buf_->data();
    }
    int type() const { return type_; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    size_t elem() const { return (size_t)(((type_ & 7) == CV_64F) ? 8 : ((type_ & 7) == CV_32F) ? 4 : 1) * ((type_ >> 3) + 1); }
    // 32F / 64F single-channel conversions (what GetRt needs)
    void convertTo(Mat& dst, int t) const
    {
        Mat out(rows, cols, t);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const double v = (type_ & 7) == CV_64F ? at<double>(y, x) : (double)at<float>(y, x);
                if ((t & 7) == CV_64F) out.at<double>(y, x) = v; else out.at<float>(y, x) = (float)v;
            }
        dst = out;
    }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    void copyTo(Mat& dst) const
    {
        if (empty()) { dst = Mat(); return; }
        dst.create(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elem());
    }
    void copyTo(Mat&& dst) const { Mat d = dst; copyTo(d); }
private:
    int type_ = 0;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
struct _InputArray {
    const Mat* m = nullptr;
    _InputArray() {}
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
};
struct _OutputArray {
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    void release() const { *m = Mat(); }
    void create(int r, int c, int t) const { m->create(r, c, t); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// calib3d stand-ins.  solvePnPRansac has no implementation here: the demo installs a function that plays the solver's role
// (the parity tests use the real cv2.solvePnPRansac through the C ABI's pose hook).  Rodrigues: rotation vector -> matrix.
typedef bool (*PnPRansacStandIn)(const std::vector<Point3f>& obj, const std::vector<Point2f>& pix, Mat& rvec, Mat& tvec);
inline PnPRansacStandIn& pnp_stand_in()
{
    static PnPRansacStandIn f = nullptr;
    return f;
}
inline bool solvePnPRansac(const std::vector<Point3f>& obj, const std::vector<Point2f>& pix, const Mat&, const Mat&, Mat& rvec, Mat& tvec)
{
    return pnp_stand_in() ? pnp_stand_in()(obj, pix, rvec, tvec) : false;
}
inline void Rodrigues(const Mat& rvec, Mat& R)
{
    const double rx = rvec.at<double>(0), ry = rvec.at<double>(1), rz = rvec.at<double>(2);
    const double th = std::sqrt(rx * rx + ry * ry + rz * rz);
    R.create(3, 3, CV_64FC1);
    double k[3] = {0, 0, 0};
    if (th > 0) { k[0] = rx / th; k[1] = ry / th; k[2] = rz / th; }
    const double c = std::cos(th), s = std::sin(th), c1 = 1 - c;
    const double Kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.at<double>(i, j) = (i == j ? c : 0.0) + c1 * k[i] * k[j] + s * Kx[3 * i + j];
}
}  // namespace cv
