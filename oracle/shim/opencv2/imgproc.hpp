// Minimal stand-in for <opencv2/imgproc.hpp> (see core.hpp in this directory). Not OpenCV code.
// Only dead or colour-only branches of the reference reach these; they throw if ever called.
#pragma once
#include "core.hpp"
namespace cv
{
enum { COLOR_BGR2GRAY = 6, COLOR_BGRA2GRAY = 10 };
enum { INTER_CUBIC = 2, WARP_INVERSE_MAP = 16, BORDER_REPLICATE = 1 };
static inline void cvtColor(const Mat&, Mat&, int) { CV_Error(Error::StsBadArg, "shim: colour input not supported"); }
static inline void warpAffine(const Mat&, Mat&, const Matx23f&, Size, int, int) { CV_Error(Error::StsBadArg, "shim: warpAffine not provided"); }
static inline void GaussianBlur(const Mat&, Mat&, Size, double, double) { CV_Error(Error::StsBadArg, "shim: GaussianBlur not provided"); }
static inline void resize(const Mat&, Mat&, Size) { CV_Error(Error::StsBadArg, "shim: resize not provided"); }
} // namespace cv
