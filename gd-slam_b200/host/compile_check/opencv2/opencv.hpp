// Minimal cv:: stand-in used ONLY to compile-check and exercise the host shim in an image without the OpenCV SDK.
// A real deployment compiles gd-slam_b200/host/*.cc against the OpenCV headers GD-SLAM already uses.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

namespace cv {
struct Point2f {
    float x = 0, y = 0;
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
    size_t elem() const { return (size_t)(((type_ & 7) == CV_32F) ? 4 : 1) * ((type_ >> 3) + 1); }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
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
}  // namespace cv
