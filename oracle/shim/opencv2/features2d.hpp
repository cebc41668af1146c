// Minimal stand-in for <opencv2/features2d.hpp> (see core.hpp in this directory). Not OpenCV code.
#pragma once
#include "core.hpp"
namespace cv
{
// cv::Feature2D with OpenCV's signatures and defaults: detect / compute forward to detectAndCompute unless overridden
class Feature2D
{
public:
    virtual ~Feature2D() {}
    virtual void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = noArray())
    { detectAndCompute(image, mask, keypoints, noArray(), false); }
    virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors)
    { detectAndCompute(image, noArray(), keypoints, descriptors, true); }
    virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false)
    { CV_Error(Error::StsBadArg, "Feature2D::detectAndCompute is not implemented"); }
    virtual int descriptorSize() const { return 0; }
    virtual int descriptorType() const { return 0; }
    virtual int defaultNorm() const { return 0; }
};
// drawing helpers referenced by samples/sample_common.cpp (never reached by sample_benchmark): declared so that the file compiles
struct DrawMatchesFlags { enum { DEFAULT = 0, DRAW_RICH_KEYPOINTS = 4 }; };
static inline void drawKeypoints(const Mat&, const std::vector<KeyPoint>&, Mat&, const Scalar& = Scalar::all(-1), int = 0)
{ CV_Error(Error::StsBadArg, "shim: drawKeypoints not provided"); }
static inline void drawMatches(const Mat&, const std::vector<KeyPoint>&, const Mat&, const std::vector<KeyPoint>&, const std::vector<DMatch>&, Mat&)
{ CV_Error(Error::StsBadArg, "shim: drawMatches not provided"); }
} // namespace cv
