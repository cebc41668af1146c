// Minimal stand-in for <opencv2/features2d.hpp>: the base class the reference's public header derives from.  TEST INFRASTRUCTURE ONLY.
#ifndef EF_SHIM_OPENCV_FEATURES2D_HPP
#define EF_SHIM_OPENCV_FEATURES2D_HPP
#include <opencv2/core/cuda.hpp>
namespace cv { class Feature2D { public: virtual ~Feature2D() {} }; }
#endif
