// Minimal stand-in for <opencv2/core/cuda_stream_accessor.hpp> (see ../core.hpp). Not OpenCV code.
#pragma once
#include "cuda.hpp"
namespace cv { namespace cuda {
struct StreamAccessor { static cudaStream_t getStream(const Stream& s) { return s.raw(); } };
}}
