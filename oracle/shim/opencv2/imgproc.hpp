// Minimal stand-in for <opencv2/imgproc.hpp> (see core.hpp in this directory). Not OpenCV code.
// The reference's descriptor sources only reach these in dead or colour-only branches (they throw if ever called); the samples use
// cvtColor(GRAY2BGR) for their display image, a nearest-neighbour resize for the preview, and the drawing calls, of which putText prints
// its text (the only place sample_image_sequence.cpp reports its match count).
#pragma once
#include "core.hpp"
namespace cv
{
enum { COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_BGRA2GRAY = 10 };
enum { INTER_CUBIC = 2, WARP_INVERSE_MAP = 16, BORDER_REPLICATE = 1 };
static inline void cvtColor(const Mat& src, Mat& dst, int code)
{
    if (code != COLOR_GRAY2BGR || src.type() != CV_8UC1) CV_Error(Error::StsBadArg, "shim: colour input not supported");
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int y = 0; y < src.rows; y++) {
        const uchar* s = src.ptr<uchar>(y); uchar* d = out.ptr<uchar>(y);
        for (int x = 0; x < src.cols; x++) d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = s[x];
    }
    dst = out;
}
static inline void warpAffine(const Mat&, Mat&, const Matx23f&, Size, int, int) { CV_Error(Error::StsBadArg, "shim: warpAffine not provided"); }
static inline void GaussianBlur(const Mat&, Mat&, Size, double, double) { CV_Error(Error::StsBadArg, "shim: GaussianBlur not provided"); }
static inline void resize(const Mat& src, Mat& dst, Size sz)
{   // preview only (samples/sample_common.cpp:47-50): nearest neighbour
    CV_Assert(!src.empty() && sz.width > 0 && sz.height > 0);
    Mat out(sz.height, sz.width, src.type());
    const size_t es = src.elemSize();
    for (int y = 0; y < sz.height; y++) {
        const uchar* s = src.ptr<uchar>(std::min(src.rows - 1, (int)((long long)y * src.rows / sz.height)));
        uchar* d = out.ptr<uchar>(y);
        for (int x = 0; x < sz.width; x++) std::memcpy(d + (size_t)x * es, s + (size_t)std::min(src.cols - 1, (int)((long long)x * src.cols / sz.width)) * es, es);
    }
    dst = out;
}
static inline void line(Mat&, Point, Point, const Scalar&, int = 1) {}
static inline void circle(Mat&, Point, int, const Scalar&, int = 1) {}
static inline void putText(Mat&, const String& text, Point, int, double, const Scalar&, int = 1) { std::cout << "[putText] " << text << std::endl; }
} // namespace cv
