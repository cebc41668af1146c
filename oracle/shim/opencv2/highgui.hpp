// Minimal stand-in for <opencv2/highgui.hpp> / imgcodecs (see core.hpp in this directory). Not OpenCV code.
// cv::imread: this image has no JPEG decoder with C headers, so the stand-in reads binary PGM (P5) files; for any other
// path it reads the side-car "<path>.pgm" that tests/test_gpu_adapter.py writes from cv2's decode of the same file (the pixels
// cv::imread(..., IMREAD_GRAYSCALE) returns where OpenCV is installed).  The result is always CV_8UC1, whatever `flags` says.
#pragma once
#include <fstream>
#include "core.hpp"
namespace cv
{
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
static inline Mat imread(const String& filename, int = IMREAD_COLOR)
{
    auto try_pgm = [](const String& path, Mat& out) {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        String magic; int w = 0, h = 0, maxv = 0;
        f >> magic >> w >> h >> maxv;
        if (magic != "P5" || w <= 0 || h <= 0 || maxv != 255) return false;
        f.get();                                              // the single whitespace byte after the header
        out.create(h, w, CV_8UC1);
        f.read((char*)out.data, (std::streamsize)w * h);
        return (bool)f;
    };
    Mat img;
    if (try_pgm(filename, img) || try_pgm(filename + ".pgm", img)) return img;
    return Mat();
}
// no display: imshow reports what it was given, waitKey returns "no key"
static inline void imshow(const String& name, const Mat& img) { std::cout << "[imshow] " << name << " " << img.cols << "x" << img.rows << "x" << img.channels() << std::endl; }
static inline int waitKey(int = 0) { return -1; }
} // namespace cv
