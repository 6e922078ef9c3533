// cvstub.cpp — implementation of the OpenCV stand-in used to compile the reference's ORBextractor.cc verbatim.
// TEST INFRASTRUCTURE (see cvstub.hpp).
#include "cvstub.hpp"
#undef list

namespace gdstub {
static thread_local std::vector<char>* g_arena = nullptr;
static thread_local size_t g_used = 0;
static const size_t ARENA = (size_t)256 << 20;
void* bump_alloc(size_t bytes)
{
    if (!g_arena) g_arena = new std::vector<char>(ARENA);
    bytes = (bytes + 15) & ~(size_t)15;
    if (g_used + bytes > ARENA) throw std::bad_alloc();
    void* p = g_arena->data() + g_used;
    g_used += bytes;
    return p;
}
void bump_reset() { g_used = 0; }
}  // namespace gdstub

namespace cv {

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax)
{
    assert(nonmax);
    (void)nonmax;
    std::vector<gdo::FastKp> k;
    gdo::fast_detect(image.data, image.cols, image.rows, image.step, threshold, k);
    keypoints.clear();
    for (const auto& p : k) keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1, (float)p.response));
}

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int)
{
    dst.create(dsize.height, dsize.width, 0);
    gdo::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double, int)
{
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2);
    (void)ksize;
    (void)sigmaX;
    Mat s = src.clone();
    dst.create(src.rows, src.cols, 0);
    gdo::gaussian7_u8(s.data, s.cols, s.rows, s.step, dst.data, dst.step);
}

void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int)
{
    Mat s = src.clone();  // src may alias the interior of dst
    dst.create(s.rows + top + bottom, s.cols + left + right, 0);
    for (int y = 0; y < dst.rows; ++y) {
        const int sy = gdo::reflect101(y - top, s.rows);
        for (int x = 0; x < dst.cols; ++x) dst.data[(size_t)y * dst.step + x] = s.data[(size_t)sy * s.step + gdo::reflect101(x - left, s.cols)];
    }
}

}  // namespace cv
