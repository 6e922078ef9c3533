// cvstub.hpp — a minimal stand-in for the parts of the OpenCV C++ API that the reference's
// src/ORBextractor.cc touches, so that file can be compiled VERBATIM (from /root/reference, never copied)
// into oracle/_ref/liborbref.so.  TEST INFRASTRUCTURE (see gd_oracle.h).
//
// The image primitives (FAST, resize, GaussianBlur, fastAtan2) are the cv2-pinned restatements of
// oracle/orb_prims.hpp.  std::list is re-pointed at a monotonic (bump) allocator so that the reference's
// pointer-address tie-break in DistributeOctTree (src/ORBextractor.cc:681-684) becomes the canonical
// "later-created node = higher address" rule.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <iterator>
#include <list>
#include <memory>
#include <vector>

#include "../orb_prims.hpp"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795

namespace gdstub {
// thread-local bump arena: addresses only ever grow between resets
void* bump_alloc(size_t bytes);
void bump_reset();
template <class T>
struct BumpAlloc {
    typedef T value_type;
    BumpAlloc() {}
    template <class U>
    BumpAlloc(const BumpAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(bump_alloc(n * sizeof(T))); }
    void deallocate(T*, size_t) {}
    template <class U>
    bool operator==(const BumpAlloc<U>&) const { return true; }
    template <class U>
    bool operator!=(const BumpAlloc<U>&) const { return false; }
};
}  // namespace gdstub

namespace std {
template <class T>
using gd_mlist = std::list<T, gdstub::BumpAlloc<T>>;
}

namespace cv {

inline int cvRound(double v) { return gdo::cv_round(v); }
inline int cvRound(float v) { return gdo::cv_round(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return gdo::cv_floor(v); }
inline int cvCeil(double v) { return gdo::cv_ceil(v); }
inline float fastAtan2(float y, float x) { return gdo::fast_atan2(y, x); }

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(float s)
    {
        x = (T)(x * s);
        y = (T)(y * s);
        return *this;
    }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct Mat {
    int rows, cols;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int /*type*/) { create(r, c, 0); }
    Mat(Size sz, int /*type*/) { create(sz.height, sz.width, 0); }
    Mat(int r, int c, int /*type*/, void* ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)ext) {}
    void create(int r, int c, int /*type*/)
    {
        if (data && rows == r && cols == c) return;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c);
        rows = r;
        cols = c;
        step = (size_t)c;
        data = buf->data();
    }
    // cv::Mat::zeros returns a MatExpr: assigning it to a Mat of the same size fills that Mat IN PLACE
    // (computeDescriptors, src/ORBextractor.cc:1037, relies on this to write into a rowRange of the output)
    struct ZerosExpr {
        int r, c;
    };
    static ZerosExpr zeros(int r, int c, int) { return ZerosExpr{r, c}; }
    Mat(const ZerosExpr& e) : rows(0), cols(0), step(0), data(nullptr) { *this = e; }
    Mat& operator=(const ZerosExpr& e)
    {
        create(e.r, e.c, 0);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, (size_t)cols);
        return *this;
    }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    template <typename T>
    T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T>
    const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat rowRange(int a, int b) const
    {
        Mat m = *this;
        m.data = data + (size_t)a * step;
        m.rows = b - a;
        return m;
    }
    Mat colRange(int a, int b) const
    {
        Mat m = *this;
        m.data = data + a;
        m.cols = b - a;
        return m;
    }
    Mat operator()(const Rect& r) const
    {
        Mat m = *this;
        m.data = data + (size_t)r.y * step + r.x;
        m.rows = r.height;
        m.cols = r.width;
        return m;
    }
    Mat clone() const
    {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
        return m;
    }
    Size size() const { return Size(cols, rows); }
};

struct _InputArray {
    const Mat* m;
    _InputArray() : m(nullptr) {}
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

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) {}  // only reached from ComputeKeyPointsOld (dead code)
};

}  // namespace cv

// the reference says `list<ExtractorNode>` / `std::list<ExtractorNode>::iterator`; route both to the bump-allocated list
#define list gd_mlist
