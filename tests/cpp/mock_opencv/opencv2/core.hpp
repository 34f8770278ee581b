// Minimal stand-in for <opencv2/core.hpp> -- TEST ONLY.  OpenCV's C++ headers are not installed
// in this image; this mock declares just the members include/sfm_match_opencv.hpp uses, with
// OpenCV's documented layout for cv::DMatch, so that the adapter can be compiled and driven.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_32F 5
namespace cv {
enum { NORM_L2 = 4, NORM_HAMMING = 6 };
struct DMatch {
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    int queryIdx, trainIdx, imgIdx;
    float distance;
};
struct Point2d {
    Point2d() : x(0), y(0) {}
    Point2d(double x_, double y_) : x(x_), y(y_) {}
    double x, y;
};
struct Point2f {
    Point2f() : x(0), y(0) {}
    float x, y;
};
struct KeyPoint {  // OpenCV's documented field order
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
struct Mat {  // continuous row-major matrix header over caller memory (or over storage made by create())
    Mat() : dims(2), rows(0), cols(0), data(nullptr), step(0), depth_(CV_8U) {}
    Mat(int r, int c, int depth, void* d, size_t s = 0)
        : dims(2), rows(r), cols(c), data(static_cast<unsigned char*>(d)), step(s ? s : static_cast<size_t>(c) * (depth == CV_32F ? 4 : 1)), depth_(depth) {}
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    int depth() const { return depth_; }
    int channels() const { return 1; }
    size_t elemSize() const { return depth_ == CV_32F ? 4 : 1; }
    void create(int r, int c, int depth) {
        own_.reset(new std::vector<unsigned char>(static_cast<size_t>(r) * c * (depth == CV_32F ? 4 : 1)));
        rows = r; cols = c; depth_ = depth; data = own_->data(); step = static_cast<size_t>(c) * (depth == CV_32F ? 4 : 1);
    }
    void release() { own_.reset(); rows = cols = 0; data = nullptr; step = 0; }
    std::shared_ptr<std::vector<unsigned char> > own_;
    int dims, rows, cols;
    unsigned char* data;
    size_t step;
    int depth_;
};
}  // namespace cv
