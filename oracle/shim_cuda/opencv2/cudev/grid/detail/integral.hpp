// Stand-in for <opencv2/cudev/grid/detail/integral.hpp> (third-party, opencv_contrib cudev; not installed): declarations only, so that
// the reference's cuda_bad.cu compiles unmodified.  calcIntegralImage (cuda_bad.cu:350-363) is never called by the test wrapper -- the
// integral image handed to the reference's computeBAD kernel is an exact int32 prefix sum computed by the caller.  TEST INFRASTRUCTURE ONLY.
#ifndef EF_SHIM_CUDEV_INTEGRAL_HPP
#define EF_SHIM_CUDEV_INTEGRAL_HPP
#include <opencv2/core/cuda.hpp>
namespace cv { namespace cudev {
template <class T> struct GlobPtrSz { T* data; size_t step; int rows, cols; };
template <class T> static inline GlobPtrSz<T> globPtr(const cuda::GpuMat& m) { return GlobPtrSz<T>{ (T*)m.data, m.step, m.rows, m.cols }; }
namespace integral_detail {
template <class S, class D> static inline void integral(const GlobPtrSz<S>&, const GlobPtrSz<D>&, int, int, cudaStream_t)
{
    std::fprintf(stderr, "cudev integral is not available in the shim\n");
    std::abort();
}
} } }
#endif
