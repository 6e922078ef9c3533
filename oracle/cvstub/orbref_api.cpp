// orbref_api.cpp — C entry point around the reference's own ORB_SLAM2::ORBextractor (compiled verbatim from
// /root/reference/src/ORBextractor.cc against cvstub).  TEST INFRASTRUCTURE.
#include "cvstub.hpp"
#include "ORBextractor.h"
#undef list

extern "C" {

struct orbref_kp {
    float x, y, size, angle, response;
    int octave, class_id;
};

// returns the number of keypoints (may exceed capacity: then only `capacity` are written)
__attribute__((visibility("default"))) int orbref_extract(const unsigned char* gray, int w, int h, size_t step, int nfeatures,
                                                          float scale, int nlevels, int ini_th, int min_th, orbref_kp* kps,
                                                          unsigned char* desc, int capacity, unsigned char* pyramid_out,
                                                          int* level_sizes /* nlevels x 2 */)
{
    gdstub::bump_reset();
    int n = 0;
    {
        ORB_SLAM2::ORBextractor ex(nfeatures, scale, nlevels, ini_th, min_th);
        cv::Mat img(h, w, CV_8UC1, (void*)gray, step);
        cv::Mat mask, d;
        std::vector<cv::KeyPoint> k;
        ex(cv::_InputArray(img), cv::_InputArray(mask), k, cv::_OutputArray(d));
        n = (int)k.size();
        for (int i = 0; i < n && i < capacity; ++i) {
            kps[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id};
            if (desc) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
        size_t off = 0;
        for (int l = 0; l < nlevels; ++l) {
            const cv::Mat& m = ex.mvImagePyramid[l];
            if (level_sizes) {
                level_sizes[2 * l] = m.cols;
                level_sizes[2 * l + 1] = m.rows;
            }
            if (pyramid_out) {
                for (int y = 0; y < m.rows; ++y) std::memcpy(pyramid_out + off + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
                off += (size_t)m.rows * m.cols;
            }
        }
    }
    gdstub::bump_reset();
    return n;
}
}
