// Minimal stand-in for <opencv2/features2d.hpp> (see core.hpp in this directory). Not OpenCV code.
#pragma once
#include "core.hpp"
namespace cv
{
class Feature2D
{
public:
    virtual ~Feature2D() {}
    virtual void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) = 0;
    virtual int descriptorSize() const { return 0; }
    virtual int descriptorType() const { return 0; }
    virtual int defaultNorm() const { return 0; }
};
} // namespace cv
